"""Python mirror of the reference's scoring-model interface, over the C ABI:

    PhyloCSFModel.make s1 pi1 s2 pi2 tree_shape          (src/PhyloCSFModel.ml:107-110)   -> Model.make
    PhyloCSFModel.lpr_leaves inst leaves t                (src/PhyloCSFModel.ml:67-82)     -> Model.lpr_leaves
    PhyloCSFModel.maximize_lpr                            (src/PhyloCSFModel.ml:84-99)     -> Model.maximize_lpr
    PhyloCSFModel.score strategy model leaves             (src/PhyloCSFModel.ml:142-146)   -> Model.score
    PhyloCSF.pleaves ?lo ?hi t leaf_ord aln               (src/PhyloCSF.ml:219-246)        -> Model.pleaves

Same names and argument meaning; the one difference is that `leaves` is a LIST of regions (each a uint8
array [ncols, n_leaves] of `Certain codes 0..63 / 64 = `Marginalize) and every call scores all of them in
one batch. The numbers come from the CUDA kernels; there is no CPU path behind this module."""
import math
import os

import numpy as np

from . import host
from .api import Context

DB = 10.0 / math.log(10.0)
_NT = {c: i for c, i in zip("ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3])}


class Model:
    CODING, NONCODING = 0, 1

    def __init__(self, paramset, ctx):
        self.paramset, self.ctx = paramset, ctx

    @classmethod
    def make(cls, paramset_prefix, species=None, device=0):
        ps = host.ParamSet(os.fspath(paramset_prefix), species=species)
        ctx = Context(device)
        ps.install(ctx)
        return cls(ps, ctx)

    @property
    def leaf_labels(self):
        return self.paramset.leaf_labels

    def pleaves(self, species, rows, lo=0, hi=None):
        """Leaf codes of region [lo, hi] of an alignment given as species names + nucleotide rows."""
        which = {sp: i for i, sp in enumerate(species)}
        hi = len(rows[0]) - 1 if hi is None else hi
        ncols = max(0, (hi - lo + 1) // 3)
        out = np.full((ncols, self.paramset.n_leaves), 64, dtype=np.uint8)
        for l, lab in enumerate(self.leaf_labels):
            r = which.get(lab)
            if r is None:
                continue
            s = rows[r]
            for c in range(ncols):
                p = lo + 3 * c
                try:
                    out[c, l] = 16 * _NT[s[p]] + 4 * _NT[s[p + 1]] + _NT[s[p + 2]]
                except KeyError:
                    pass
        return out

    def _stage(self, leaves):
        """Upload `leaves` as the context's batch. Always uploads: an identity-keyed cache would go stale when a
        caller passes a fresh list that happens to reuse the address of a dead one, or mutates a list in place."""
        off = np.zeros(len(leaves) + 1, dtype=np.int64)
        for i, c in enumerate(leaves):
            off[i + 1] = off[i] + c.shape[0]
        codes = np.concatenate(leaves, axis=0) if off[-1] else np.zeros((0, self.paramset.n_leaves), dtype=np.uint8)
        self.ctx.batch_upload(off, codes)

    def _lpr_staged(self, which, t):
        self.ctx.pt_build(which, [t])
        lpr, elpr, st = self.ctx.lpr_all([which])
        return [{"lpr_leaves": float(a), "elpr_anc": float(b)} for a, b in zip(lpr[0], elpr[0])]

    def _maximize_staged(self, which, init, lo, hi, accuracy):
        rho, lpr, elpr, st, ne = self.ctx.maximize_lpr(which, init, lo, hi, accuracy)
        return [(float(r), {"lpr_leaves": float(a), "elpr_anc": float(b)}) for r, a, b in zip(rho, lpr, elpr)]

    def lpr_leaves(self, which, leaves, t):
        """-> list of {'lpr_leaves', 'elpr_anc'} for instance `which` (CODING / NONCODING) at tree scale t."""
        self._stage(leaves)
        return self._lpr_staged(which, t)

    def maximize_lpr(self, which, leaves, init=1.0, lo=1e-2, hi=10.0, accuracy=0.01):
        """-> list of (rho, {'lpr_leaves', 'elpr_anc'})"""
        self._stage(leaves)
        return self._maximize_staged(which, init, lo, hi, accuracy)

    def score(self, strategy, leaves):
        """strategy 'FixedLik' | 'MaxLik' -> list of {'score', 'anc_comp_score', 'diagnostics'} (decibans)."""
        self._stage(leaves)  # once for both models
        if strategy == "FixedLik":
            c = self._lpr_staged(self.CODING, 1.0)
            n = self._lpr_staged(self.NONCODING, 1.0)
            rho = [(1.0, 1.0)] * len(leaves)
        elif strategy == "MaxLik":
            mc = self._maximize_staged(self.CODING, 1.0, 1e-2, 10.0, 0.01)
            mn = self._maximize_staged(self.NONCODING, 1.0, 1e-2, 10.0, 0.01)
            c, n = [x[1] for x in mc], [x[1] for x in mn]
            rho = [(a[0], b[0]) for a, b in zip(mc, mn)]
        else:
            raise ValueError(strategy)
        return [{"score": DB * (a["lpr_leaves"] - b["lpr_leaves"]), "anc_comp_score": DB * (a["elpr_anc"] - b["elpr_anc"]),
                 "diagnostics": {"rho_C": r[0], "rho_N": r[1], "L(C)": DB * a["lpr_leaves"], "L(NC)": DB * b["lpr_leaves"]}}
                for a, b, r in zip(c, n, rho)]

    def close(self):
        self.ctx.close()
