"""ctypes face of the C++ host layer (include/phylocsf_host.h): parameter sets, tree numbering and
rate-matrix diagonalisation — the product's own restatement of the reference's host setup
(src/PhyloCSF.ml:406-467, src/PhyloCSFModel.ml:107-110, lib/CamlPaml/Q.ml:124-177)."""
import ctypes
import os

import numpy as np

from . import _native as N

HOST_SYMBOLS = [
    "pcsf_paramset_load", "pcsf_paramset_free", "pcsf_paramset_n_leaves", "pcsf_paramset_leaf_label",
    "pcsf_paramset_tree", "pcsf_paramset_qdiag", "pcsf_paramset_install", "pcsf_qdiag_reversible", "pcsf_omega_q", "pcsf_omega_score",
    "pcsf_host_tree_program",
]


class HostError(RuntimeError):
    pass


def _lib():
    L = N.load()
    if not getattr(L, "_host_ready", False):
        vp = ctypes.c_void_p
        L.pcsf_paramset_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_int]
        L.pcsf_paramset_free.argtypes = [vp]
        L.pcsf_paramset_free.restype = None
        L.pcsf_paramset_n_leaves.argtypes = [vp]
        L.pcsf_paramset_leaf_label.argtypes = [vp, ctypes.c_int]
        L.pcsf_paramset_leaf_label.restype = ctypes.c_char_p
        L.pcsf_paramset_tree.argtypes = [vp, vp, vp]
        L.pcsf_paramset_qdiag.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp]
        L.pcsf_paramset_install.argtypes = [vp, vp]
        L.pcsf_qdiag_reversible.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_char_p, ctypes.c_int]
        L.pcsf_omega_q.argtypes = [vp, vp, vp, ctypes.c_char_p, ctypes.c_int]
        L.pcsf_omega_score.argtypes = [vp, ctypes.c_int64, vp, vp, ctypes.c_double, ctypes.c_double, vp, vp, vp]
        L.pcsf_host_tree_program.argtypes = [ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp]
        L._host_ready = True
    return L


class ParamSet:
    """<base>/PhyloCSF_Parameters/<name>.nh + ECMs, optionally pruned to `species`."""

    def __init__(self, prefix, species=None, with_ecm=True):
        L = _lib()
        h = ctypes.c_void_p()
        err = ctypes.create_string_buffer(512)
        sp = ",".join(species).encode() if species else None
        rc = L.pcsf_paramset_load(os.fspath(prefix).encode(), sp, 1 if with_ecm else 0, ctypes.byref(h), err, 512)
        if rc != 0:
            raise HostError(err.value.decode())
        self._L, self._h = L, h
        self.n_leaves = L.pcsf_paramset_n_leaves(h)
        self.leaf_labels = [L.pcsf_paramset_leaf_label(h, i).decode() for i in range(self.n_leaves)]
        self.children = np.zeros(2 * (self.n_leaves - 1), dtype=np.int32)
        self.branch_len = np.zeros(2 * self.n_leaves - 2)
        L.pcsf_paramset_tree(h, N.ptr(self.children), N.ptr(self.branch_len))
        self.with_ecm = with_ecm

    def qdiag(self, which):
        Q, S, Sinv = np.empty((64, 64)), np.empty((64, 64)), np.empty((64, 64))
        lam, prior = np.empty(64), np.empty(64)
        rc = self._L.pcsf_paramset_qdiag(self._h, which, N.ptr(Q), N.ptr(S), N.ptr(Sinv), N.ptr(lam), N.ptr(prior))
        if rc != 0:
            raise HostError("pcsf_paramset_qdiag failed (%d)" % rc)
        return {"Q": Q, "S": S, "Sinv": Sinv, "lam": lam, "prior": prior}

    def install(self, ctx):
        rc = self._L.pcsf_paramset_install(ctx._h, self._h)
        ctx._check(rc)
        ctx.n_leaves = self.n_leaves

    def close(self):
        if getattr(self, "_h", None):
            self._L.pcsf_paramset_free(self._h)
            self._h = None

    __del__ = close


def qdiag_reversible(Q, w):
    L = _lib()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    S, Sinv, lam, prior = np.empty((64, 64)), np.empty((64, 64)), np.empty(64), np.empty(64)
    err = ctypes.create_string_buffer(512)
    rc = L.pcsf_qdiag_reversible(N.ptr(Q), N.ptr(w), N.ptr(S), N.ptr(Sinv), N.ptr(lam), N.ptr(prior), err, 512)
    if rc != 0:
        raise HostError(err.value.decode())
    return {"Q": Q, "S": S, "Sinv": Sinv, "lam": lam, "prior": prior}


def omega_q(v):
    L = _lib()
    v = np.ascontiguousarray(v, dtype=np.float64)
    assert v.size == 12
    Q, pi = np.empty((64, 64)), np.empty(64)
    err = ctypes.create_string_buffer(512)
    rc = L.pcsf_omega_q(N.ptr(v), N.ptr(Q), N.ptr(pi), err, 512)
    if rc != 0:
        raise HostError(err.value.decode())
    return Q, pi


OMEGA_DIAG = ("L(H0)", "rho_H0", "kappa_H0", "omega_H0", "sigma_H0", "L(H1)", "rho_H1", "kappa_H1", "omega_H1", "sigma_H1")


def omega_score(ctx, region_off, codes, omega_H1=0.2, sigma_H1=0.01):
    """OmegaModel.score (src/OmegaModel.ml:195-219) for a batch of regions on `ctx` (tree set): -> (score[R] in decibans,
    diag[R, 10] in OMEGA_DIAG order, status[R])."""
    L = _lib()
    ro = np.ascontiguousarray(region_off, dtype=np.int64)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    R = ro.size - 1
    score, diag, st = np.zeros(R), np.zeros((R, 10)), np.zeros(R, dtype=np.int32)
    rc = L.pcsf_omega_score(ctx._h, R, N.ptr(ro), N.ptr(codes), omega_H1, sigma_H1, N.ptr(score), N.ptr(diag), N.ptr(st))
    ctx._check(rc, ok_numeric=True)
    ctx.nregions = R
    return score, diag, st


def tree_program(n_leaves, children, level=0, keep=True):
    """pcsf_host_tree_program: (ops int32 [n_ops, 4], subtabs int32 [n_tabs, 5], (n_tab2, n_tab3, n_tab4, max_levels)).
    Host logic only - works without a GPU."""
    L = _lib()
    ch = np.ascontiguousarray(children, dtype=np.int32).reshape(-1)
    ops = np.zeros((4 * n_leaves + 8, 4), dtype=np.int32)
    tabs = np.zeros((n_leaves + 1, 5), dtype=np.int32)
    info = np.zeros(4, dtype=np.int32)
    n = L.pcsf_host_tree_program(n_leaves, N.ptr(ch), level, 1 if keep else 0, N.ptr(ops), ops.shape[0], N.ptr(tabs), tabs.shape[0],
                                 N.ptr(info))
    if n < 0:
        raise HostError("pcsf_host_tree_program failed (%d)" % n)
    return ops[:n].copy(), tabs[: int(info[:3].sum())].copy(), tuple(int(x) for x in info)
