"""Shared helpers for the parity tests: build oracle models and mirror them into a device context."""
import os

import numpy as np

from oracle import oracle as o

DB = 10.0 / np.log(10.0)


def oracle_paramset(params_base, pset, strategy="fixed", species=None):
    opts = o.Options(strategy=strategy, species=species)
    return o.load_paramset(os.path.join(params_base, "PhyloCSF_Parameters", pset), opts)


def push_tree(ctx, tree):
    ctx.tree_set(tree.n_leaves, tree.children_array(), np.array(tree.branches[: tree.root]))


def push_qdiag(ctx, model_id, q):
    """Q.Diag.t + equilibrium prior (what P14n.update leaves in place, PhyloModel.ml:132-146)."""
    ctx.model_set(model_id, q.S, q.Sinv, q.lam, q.equilibrium())


def make_context(ps, device=0):
    import phylocsf_b200 as pb

    ctx = pb.Context(device)
    push_tree(ctx, ps.tree)
    push_qdiag(ctx, 0, ps.model.coding_model.q)
    push_qdiag(ctx, 1, ps.model.noncoding_model.q)
    return ctx


def regions_to_batch(region_codes):
    """list of uint8 [ncols_r, n_leaves] -> (region_off int64, codes uint8 [total, n_leaves])."""
    off = np.zeros(len(region_codes) + 1, dtype=np.int64)
    for i, c in enumerate(region_codes):
        off[i + 1] = off[i] + c.shape[0]
    n_leaves = region_codes[0].shape[1]
    codes = np.concatenate(region_codes, axis=0) if off[-1] > 0 else np.zeros((0, n_leaves), dtype=np.uint8)
    return off, np.ascontiguousarray(codes, dtype=np.uint8)


def oracle_fixed(ps, region_codes, rho=1.0):
    """(lpr[2,R], elpr[2,R]) from the oracle, models in (coding, noncoding) order."""
    R = len(region_codes)
    lpr, elpr = np.zeros((2, R)), np.zeros((2, R))
    for m, inst in enumerate((ps.model.coding_model, ps.model.noncoding_model)):
        mod = inst.model(rho)
        for r, c in enumerate(region_codes):
            if c.shape[0]:
                lpr[m, r], elpr[m, r], _, _ = o.lpr_columns(mod, c)
    return lpr, elpr


def example_codes(ps, fn, frames=1):
    """pleaves of a PhyloCSF_Examples alignment for the AsIs regions of `frames` frames."""
    from tools import golden_params as gp

    species, aln = o.input_mfa(gp.example_lines(fn))
    aln = [s.replace("u", "t").replace("U", "T") for s in aln]
    rc = [o.revcomp(s) for s in aln]
    which = {sp: i for i, sp in enumerate(species)}
    t = ps.tree
    leaf_ord = [which.get(t.labels[i]) for i in range(t.n_leaves)]
    regs = o.candidate_regions(aln[0], o.Options(frames=frames))
    return [o.pleaves(t, leaf_ord, rc if r else aln, lo, hi) for r, lo, hi in regs], (aln, leaf_ord)


def host_cores():
    """Cores this process may use (ignores OMP_NUM_THREADS, which torchrun sets to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_fixed_batch(ps, region_codes, rho=1.0, threads=0):
    """oracle_fixed on all host cores (oracle_lpr_batch: one region per OpenMP task). Same numbers as oracle_fixed."""
    off, codes = regions_to_batch(region_codes)
    R = len(region_codes)
    t = ps.tree
    ch = t.children_array()
    L = o.lib()
    lpr, elpr = np.zeros((2, R)), np.zeros((2, R))
    for m, inst in enumerate((ps.model.coding_model, ps.model.noncoding_model)):
        mod = inst.model(rho)
        pms, prior = np.ascontiguousarray(mod.pms), np.ascontiguousarray(mod.prior())
        a, b = np.empty(R), np.empty(R)
        L.oracle_lpr_batch(t.n_leaves, ch.ctypes.data, o._dp(pms), None, o._dp(prior), 64, R, off.ctypes.data, codes.ctypes.data,
                           o._dp(a), o._dp(b), threads or host_cores())
        lpr[m], elpr[m] = a, b
    return lpr, elpr


def _run_workers(kind, params_base, pset, items, extra, workers):
    """Spread `items` over worker processes running tests/oracle_worker.py (plain subprocesses with pickle files: safe
    whatever CUDA state the parent holds). Returns the per-item results in input order."""
    import pickle
    import subprocess
    import sys
    import tempfile

    workers = max(1, min(workers or host_cores(), len(items)))
    d = tempfile.mkdtemp(prefix="pcsf_oracle_")
    procs = []
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for w in range(workers):
        mine = list(range(w, len(items), workers))  # round robin: neighbours (similar sizes) go to different workers
        fin, fout = os.path.join(d, "in%d.pkl" % w), os.path.join(d, "out%d.pkl" % w)
        with open(fin, "wb") as f:
            pickle.dump({"kind": kind, "params_base": params_base, "pset": pset, "idx": mine, "items": [items[i] for i in mine], "extra": extra}, f)
        procs.append((subprocess.Popen([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_worker.py"), fin, fout], env=env), mine, fout))
    out = [None] * len(items)
    for pr, mine, fout in procs:
        rc = pr.wait()
        assert rc == 0, "oracle worker failed (%d)" % rc
        with open(fout, "rb") as f:
            for i, r in zip(mine, pickle.load(f)):
                out[i] = r
    return out


def oracle_mle_parallel(params_base, pset, region_codes, workers=0):
    """The oracle's find_init + Brent per region under both ECMs, regions spread over worker processes.
    -> per region [(rho, lpr, elpr, iterations, random tries)] x 2"""
    return _run_workers("mle", params_base, pset, [np.ascontiguousarray(c) for c in region_codes], None, workers)


def oracle_lines_parallel(params_base, pset, named_lines, workers=0, **kw):
    """o.process_alignment per alignment in worker processes -> list of line lists (input order)."""
    return _run_workers("lines", params_base, pset, list(named_lines), kw, workers)


def oracle_omega_parallel(params_base, pset, region_codes, omega_H1=0.2, sigma_H1=0.01, workers=0):
    """The oracle's OmegaModel.score per region in worker processes -> per region
    (score dB, lpr_H0, rho_H0, kappa_H0, lpr_H1, rho_H1, kappa_H1), unrounded."""
    return _run_workers("omega", params_base, pset, [np.ascontiguousarray(c) for c in region_codes], (omega_H1, sigma_H1), workers)
