"""Thin Python face of the C ABI (include/phylocsf_b200.h): numpy in, numpy out. Used by the parity
tests and bench.py; the drop-in host (C++ CLI) links the same library directly."""
import ctypes

import numpy as np

from . import _native as N


class PcsfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pcsf error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One GPU. Mirrors the call order of the reference seam: tree (T.t) -> models (Q.Diag.t + prior)
    -> P(t) tables (PhyloModel.make) -> leaves (pleaves) -> lpr_leaves / maximize_lpr."""

    def __init__(self, device=0):
        self._L = N.load()
        h = ctypes.c_void_p()
        rc = self._L.pcsf_create(device, ctypes.byref(h))
        if rc != N.PCSF_OK:
            raise PcsfError(rc, "pcsf_create(device=%d) failed: no usable CUDA device (there is no CPU fallback)" % device)
        self._h = h
        self.n_leaves = 0
        self.nregions = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.pcsf_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, ok_numeric=False):
        if rc == N.PCSF_OK or (ok_numeric and rc == -4):
            return rc
        raise PcsfError(rc, self._L.pcsf_last_error(self._h).decode())

    def stream_set(self, cuda_stream_ptr):
        self._check(self._L.pcsf_stream_set(self._h, ctypes.c_void_p(cuda_stream_ptr or 0)))

    OPT_RESCALE, OPT_PRUNE_FORM, OPT_CHERRY_TABLES = 1, 2, 3
    FORM_AUTO, FORM_NARROW, FORM_WIDE = 0, 1, 2

    def option_set(self, option, value):
        self._check(self._L.pcsf_option_set(self._h, option, int(value)))

    def tree_set(self, n_leaves, children, branch_len):
        ch = np.ascontiguousarray(children, dtype=np.int32).ravel()
        bl = _f64(branch_len)[: 2 * n_leaves - 2]
        assert ch.size == 2 * (n_leaves - 1) and bl.size == 2 * n_leaves - 2
        self._check(self._L.pcsf_tree_set(self._h, n_leaves, N.ptr(ch), N.ptr(np.ascontiguousarray(bl))))
        self.n_leaves = n_leaves

    def model_set(self, model_id, S, Sinv, lam, prior):
        S, Sinv, lam, prior = _f64(S), _f64(Sinv), _f64(lam), _f64(prior)
        assert S.shape == (64, 64) and Sinv.shape == (64, 64) and lam.shape == (64,) and prior.shape == (64,)
        self._check(self._L.pcsf_model_set(self._h, model_id, N.ptr(S), N.ptr(Sinv), N.ptr(lam), N.ptr(prior)))

    def pt_build(self, model_id, scales, check=True):
        sc = _f64(np.atleast_1d(scales))
        st = np.zeros(sc.size, dtype=np.int32)
        self._check(self._L.pcsf_pt_build(self._h, model_id, sc.size, N.ptr(sc), N.ptr(st)), ok_numeric=not check)
        return st

    def pt_get(self, model_id, scale_idx, branch):
        P = np.empty((64, 64))
        self._check(self._L.pcsf_pt_get(self._h, model_id, scale_idx, branch, N.ptr(P)))
        return P

    def batch_upload(self, region_off, codes):
        """region_off: int64 [nregions+1]; codes: uint8 [total_cols, n_leaves] (numpy) or a raw host
        address (int) of such a buffer, e.g. pinned memory."""
        ro = np.ascontiguousarray(region_off, dtype=np.int64)
        if isinstance(codes, np.ndarray):
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            assert codes.size == int(ro[-1]) * self.n_leaves, "codes must be [total_cols, n_leaves]"
        self._check(self._L.pcsf_batch_upload(self._h, ro.size - 1, N.ptr(ro), N.ptr(codes)))
        self.nregions = ro.size - 1

    def batch_upload_alignments(self, aln_off, aln_len, nt, frames):
        ao = np.ascontiguousarray(aln_off, dtype=np.int64)
        al = np.ascontiguousarray(aln_len, dtype=np.int32)
        if isinstance(nt, np.ndarray):
            nt = np.ascontiguousarray(nt, dtype=np.uint8)
        self._check(self._L.pcsf_batch_upload_alignments(self._h, ao.size, N.ptr(ao), N.ptr(al), N.ptr(nt), frames))
        self.nregions = int(self._L.pcsf_batch_nregions(self._h))

    def batch_upload_alignments_parts(self, aln_off, aln_len, parts, frames):
        """pcsf_batch_upload_alignments_parts: the nucleotide buffer as a list of uint8 arrays; aln_off are
        offsets into their concatenation."""
        import ctypes
        ao = np.ascontiguousarray(aln_off, dtype=np.int64)
        al = np.ascontiguousarray(aln_len, dtype=np.int32)
        parts = [np.ascontiguousarray(q, dtype=np.uint8) for q in parts]
        ptrs = (ctypes.c_void_p * max(1, len(parts)))(*[q.ctypes.data for q in parts])
        nbytes = np.array([q.size for q in parts], dtype=np.int64)
        self._check(self._L.pcsf_batch_upload_alignments_parts(self._h, ao.size, N.ptr(ao), N.ptr(al), len(parts),
                                                                ctypes.cast(ptrs, ctypes.c_void_p), N.ptr(nbytes), frames))
        self.nregions = int(self._L.pcsf_batch_nregions(self._h))

    @property
    def ncols(self):
        return int(self._L.pcsf_batch_ncols(self._h))

    def batch_codes(self):
        """pcsf_batch_codes_get: the staged leaf codes [total_cols, n_leaves] (after batch_upload_alignments: the
        device pleaves' output)."""
        out = np.empty((self.ncols, self.n_leaves), dtype=np.uint8)
        self._check(self._L.pcsf_batch_codes_get(self._h, N.ptr(out)))
        return out

    def lpr_all(self, model_ids, scale_idx=None, out=None):
        mids = np.ascontiguousarray(model_ids, dtype=np.int32)
        sidx = None if scale_idx is None else np.ascontiguousarray(scale_idx, dtype=np.int32)
        m, R = mids.size, self.nregions
        if out is None:
            lpr, elpr = np.empty((m, R)), np.empty((m, R))
        else:
            lpr, elpr = out
        st = np.zeros((m, R), dtype=np.int32)
        self._check(self._L.pcsf_lpr_all(self._h, m, N.ptr(mids), N.ptr(sidx), N.ptr(lpr), N.ptr(elpr), N.ptr(st)))
        return lpr, elpr, st

    def score_alignments(self, aln_off, aln_len, nt, frames, model_ids, scale_idx=None, out=None):
        """Pipelined pleaves + lpr_leaves from host nucleotide rows (H2D of the next chunk overlaps compute)."""
        ao = np.ascontiguousarray(aln_off, dtype=np.int64)
        al = np.ascontiguousarray(aln_len, dtype=np.int32)
        if isinstance(nt, np.ndarray):
            nt = np.ascontiguousarray(nt, dtype=np.uint8)
        mids = np.ascontiguousarray(model_ids, dtype=np.int32)
        sidx = None if scale_idx is None else np.ascontiguousarray(scale_idx, dtype=np.int32)
        m, R = mids.size, ao.size * frames
        lpr, elpr = out if out is not None else (np.empty((m, R)), np.empty((m, R)))
        st = np.zeros((m, R), dtype=np.int32)
        self._check(self._L.pcsf_score_alignments(self._h, ao.size, N.ptr(ao), N.ptr(al), N.ptr(nt), frames, m, N.ptr(mids),
                                                  N.ptr(sidx), N.ptr(lpr), N.ptr(elpr), N.ptr(st)))
        self.nregions = 0
        return lpr, elpr, st

    def lpr(self, eval_model, eval_scale, eval_region):
        em = np.ascontiguousarray(eval_model, dtype=np.int32)
        es = np.ascontiguousarray(eval_scale, dtype=np.int32)
        er = np.ascontiguousarray(eval_region, dtype=np.int64)
        n = em.size
        lpr, elpr, st = np.empty(n), np.empty(n), np.zeros(n, dtype=np.int32)
        self._check(self._L.pcsf_lpr(self._h, n, N.ptr(em), N.ptr(es), N.ptr(er), N.ptr(lpr), N.ptr(elpr), N.ptr(st)))
        return lpr, elpr, st

    def models_set(self, first_id, S, Sinv, lam, prior):
        S, Sinv, lam, prior = _f64(S), _f64(Sinv), _f64(lam), _f64(prior)
        n = S.shape[0]
        assert S.shape == (n, 64, 64) and Sinv.shape == (n, 64, 64) and lam.shape == (n, 64) and prior.shape == (n, 64)
        self._check(self._L.pcsf_models_set(self._h, first_id, n, N.ptr(S), N.ptr(Sinv), N.ptr(lam), N.ptr(prior)))

    def omega_models_set(self, first_id, q_settings, check=True):
        qs = _f64(q_settings).reshape(-1, 12)
        st = np.zeros(qs.shape[0], dtype=np.int32)
        self._check(self._L.pcsf_omega_models_set(self._h, first_id, qs.shape[0], N.ptr(qs), N.ptr(st)), ok_numeric=not check)
        return st

    def omega_cache_reset(self, n_slots):
        self._check(self._L.pcsf_omega_cache_reset(self._h, int(n_slots)))

    def omega_models_set_cached(self, first_id, q_settings, cache_slot, check=True):
        qs = _f64(q_settings).reshape(-1, 12)
        sl = np.ascontiguousarray(cache_slot, dtype=np.int64)
        assert sl.size == qs.shape[0]
        st = np.zeros(qs.shape[0], dtype=np.int32)
        self._check(self._L.pcsf_omega_models_set_cached(self._h, first_id, qs.shape[0], N.ptr(qs), N.ptr(sl), N.ptr(st)), ok_numeric=not check)
        return st

    def counters(self, reset=False):
        names = {0: "pt_slots", 1: "column_evaluations", 2: "eig_matrices", 3: "eig_sweeps", 4: "tiles"}
        out = {v: int(self._L.pcsf_counter(self._h, k)) for k, v in names.items()}
        if reset:
            self._L.pcsf_counter(self._h, -1)
        return out

    def model_get(self, model_id):
        S, Sinv, lam, prior = np.empty((64, 64)), np.empty((64, 64)), np.empty(64), np.empty(64)
        self._check(self._L.pcsf_model_get(self._h, model_id, N.ptr(S), N.ptr(Sinv), N.ptr(lam), N.ptr(prior)))
        return {"S": S, "Sinv": Sinv, "lam": lam, "prior": prior}

    def pt_build_pairs(self, pair_model, pair_scale, check=True):
        pm = np.ascontiguousarray(pair_model, dtype=np.int32)
        sc = _f64(pair_scale)
        st = np.zeros(pm.size, dtype=np.int32)
        self._check(self._L.pcsf_pt_build_pairs(self._h, pm.size, N.ptr(pm), N.ptr(sc), N.ptr(st)), ok_numeric=not check)
        return st

    def lpr_pairs(self, eval_pair, eval_region):
        ep = np.ascontiguousarray(eval_pair, dtype=np.int64)
        er = np.ascontiguousarray(eval_region, dtype=np.int64)
        n = ep.size
        lpr, elpr, st = np.empty(n), np.empty(n), np.zeros(n, dtype=np.int32)
        self._check(self._L.pcsf_lpr_pairs(self._h, n, N.ptr(ep), N.ptr(er), N.ptr(lpr), N.ptr(elpr), N.ptr(st)))
        return lpr, elpr, st

    def posteriors(self, model_id, scale_idx=0, nodes=(), ecounts=True, z=True):
        """Outside algorithm over the staged batch (pcsf_posteriors): -> (node_post [len(nodes), ncols, 64],
        ecounts [n_branches, 64, 64] or None, z [ncols] or None)."""
        nd = np.ascontiguousarray(list(nodes), dtype=np.int32)
        nc, nbr = self.ncols, 2 * self.n_leaves - 2
        post = np.zeros((nd.size, nc, 64))
        ec = np.zeros((nbr, 64, 64)) if ecounts else None
        zz = np.zeros(nc) if z else None
        self._check(self._L.pcsf_posteriors(self._h, model_id, scale_idx, nd.size, N.ptr(nd) if nd.size else None,
                                            N.ptr(post) if nd.size else None, N.ptr(ec), N.ptr(zz)))
        return post, ec, zz

    def column_terms(self, m):
        n = self.ncols
        a, b = np.empty(n), np.empty(n)
        self._check(self._L.pcsf_column_terms(self._h, m, N.ptr(a), N.ptr(b)))
        return a, b

    def maximize_lpr(self, model_id, init=1.0, lo=1e-2, hi=10.0, accuracy=0.01, check=True):
        R = self.nregions
        rho, lpr, elpr = np.empty(R), np.empty(R), np.empty(R)
        st, ne = np.zeros(R, dtype=np.int32), np.zeros(R, dtype=np.int32)
        self._check(self._L.pcsf_maximize_lpr(self._h, model_id, init, lo, hi, accuracy, N.ptr(rho), N.ptr(lpr),
                                              N.ptr(elpr), N.ptr(st), N.ptr(ne)), ok_numeric=not check)
        return rho, lpr, elpr, st, ne

    def maximize_lpr_multi(self, model_ids, init=1.0, lo=1e-2, hi=10.0, accuracy=0.01, check=True):
        mids = np.ascontiguousarray(model_ids, dtype=np.int32)
        m, R = mids.size, self.nregions
        rho, lpr, elpr = np.empty((m, R)), np.empty((m, R)), np.empty((m, R))
        st, ne = np.zeros((m, R), dtype=np.int32), np.zeros((m, R), dtype=np.int32)
        self._check(self._L.pcsf_maximize_lpr_multi(self._h, m, N.ptr(mids), init, lo, hi, accuracy, N.ptr(rho), N.ptr(lpr),
                                                    N.ptr(elpr), N.ptr(st), N.ptr(ne)), ok_numeric=not check)
        return rho, lpr, elpr, st, ne

    def table_level(self, model_id, scale_idx=0):
        """Subtree-table level (0, 2, 3, 4) the P set carries right now (pcsf_table_level)."""
        rc = int(self._L.pcsf_table_level(self._h, model_id, scale_idx))
        if rc < 0:
            raise PcsfError(rc, "pcsf_table_level: unknown model / scale")
        return rc

    def last_launch_info(self):
        """{'form': 'narrow'|'wide', 'table_level', 'tiles', 'grid'} of the most recent pruning launch."""
        f = int(self._L.pcsf_last_launch_info(self._h, 0))
        return {"form": {1: "narrow", 2: "wide"}.get(f, "none"), "table_level": int(self._L.pcsf_last_launch_info(self._h, 1)),
                "tiles": int(self._L.pcsf_last_launch_info(self._h, 2)), "grid": int(self._L.pcsf_last_launch_info(self._h, 3))}

    def last_ms(self, which):
        return float(self._L.pcsf_last_ms(self._h, which))

    def total_ms(self, reset=False):
        """Cumulative device time per kernel class since the last reset (pcsf_total_ms)."""
        names = {0: "prune", 1: "reduce", 2: "pt_build", 5: "subtree_tables", 6: "omega_eig"}
        out = {v: float(self._L.pcsf_total_ms(self._h, k)) for k, v in names.items()}
        if reset:
            self._L.pcsf_total_ms(self._h, -1)
        return out

    @property
    def launch_count(self):
        return int(self._L.pcsf_launch_count(self._h))
