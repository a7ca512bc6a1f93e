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
