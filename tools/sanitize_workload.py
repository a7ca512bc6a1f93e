#!/usr/bin/env python3
"""Small workload for compute-sanitizer (GPU box): every kernel and both forms of the pruning kernel once.
    compute-sanitizer --tool memcheck|synccheck|initcheck|racecheck python tools/sanitize_workload.py
29mammals, regions of 200/130/17/0/64 columns: pcsf_lpr_all in the narrow and the wide form (two scales,
both models), PCSF_OPT_RESCALE in both forms, pcsf_maximize_lpr_multi, the omega entry points
(pcsf_omega_models_set + pcsf_pt_build_pairs + pcsf_lpr_pairs), pcsf_score_alignments and
pcsf_batch_upload_alignments_parts (6 frames); round 2: pcsf_omega_models_set_cached (warm starts), pcsf_posteriors (K6),
pcsf_omega_score. Results are checked against the oracle where it is cheap."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcsf_helpers as H  # noqa: E402
from oracle import oracle as o  # noqa: E402
from tools import golden_params as gp  # noqa: E402

base = gp.materialize(tempfile.mkdtemp(), sets=["29mammals"])
ps = H.oracle_paramset(base, "29mammals")
n = ps.tree.n_leaves
rng = np.random.default_rng(1)
mc = ps.model.coding_model.model(1.0)
regs = [o.simulate_columns(mc, L, rng) if L else np.zeros((0, n), dtype=np.uint8) for L in (200, 130, 17, 0, 64)]
ctx = H.make_context(ps)
ctx.pt_build(0, [1.0, 0.5])
ctx.pt_build(1, [1.0, 0.5])
off, codes = H.regions_to_batch(regs)
ctx.batch_upload(off, codes)
lo, eo = H.oracle_fixed(ps, regs)
ONLY_NEW = os.environ.get("PCSF_SAN_ONLY") == "round2"  # racecheck: K1 / K5 / K6 only (the pruning kernels take hours under it)
qs = np.tile(np.array([2.5, 1.0, 1.0] + [1.0] * 9), (3, 1))
if not ONLY_NEW:
    res = {}
    for form in (1, 2):
        ctx.option_set(2, form)
        for rescale in (0, 1):
            ctx.option_set(1, rescale)
            lpr, elpr, st = ctx.lpr_all([0, 1])
            assert np.abs(H.DB * (lpr - lo)).max() < 1e-7
            res[(form, rescale)] = lpr
            ctx.lpr_all([0, 1], scale_idx=[1, 1])
    assert (res[(1, 0)] == res[(2, 0)]).all()
    ctx.option_set(1, 0)
    ctx.option_set(2, 0)
    ctx.maximize_lpr_multi([0, 1])
    qs = np.tile(np.array([2.5, 1.0, 1.0] + [1.0] * 9), (3, 1))
    ctx.omega_models_set(4, qs)
    ctx.pt_build_pairs([4, 5, 6], [1.0, 0.7, 1.3])
    ctx.lpr_pairs([0, 1, 2], [0, 1, 2])
    L = 93
    nt = np.frombuffer(b"ACGTacgtN-", dtype=np.uint8)[rng.integers(0, 10, size=(2, n, L))]
    flat = nt.reshape(-1)
    aoff = np.array([0, n * L], dtype=np.int64)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    a = ctx.score_alignments(aoff, [L, L], flat, 6, [0, 1])
    ctx.batch_upload_alignments_parts(aoff, [L, L], [flat[: n * L], flat[n * L:]], 6)
    b = ctx.lpr_all([0, 1])
    assert np.array_equal(a[0], b[0], equal_nan=True)
# round 2: K5 with warm starts (one-pass rotations), K6 (outside algorithm + expected counts), the omega strategy in one call
ctx.omega_cache_reset(3)
for kappa in (2.5, 2.7, 2.71):
    q2 = qs.copy()
    q2[:, 0] = kappa
    ctx.omega_models_set_cached(4, q2, [0, 1, 2])
ctx.batch_upload(off, codes)
ctx.pt_build(0, [1.0])
post, ec, z = ctx.posteriors(0, 0, nodes=[2 * n - 2, n, 0])
zo, po, eo2 = o.posteriors_columns(mc, codes[:40])
assert np.allclose(z[:40], zo, rtol=1e-10) and np.allclose(post[0][:40], po[:, 2 * n - 2], atol=1e-11)
from phylocsf_b200 import host  # noqa: E402
if not ONLY_NEW:
    sc, dg, st = host.omega_score(ctx, off[:3], codes[: off[2]])
    assert (st == 0).all() and np.isfinite(sc).all()
# shared subtree tables: a second context on the same GPU attaches to the first one's block, the first one goes away
ctx.option_set(2, 2)
ctx.option_set(3, 2)
ctx.pt_build(0, [1.0])
ctx.batch_upload(off, codes)
ra = ctx.lpr_all([0])
ctx2 = H.make_context(ps)
ctx2.option_set(2, 2)
ctx2.option_set(3, 2)
ctx2.pt_build(0, [1.0])
ctx2.batch_upload(off, codes)
rb = ctx2.lpr_all([0])
assert ctx2.table_level(0) == ctx.table_level(0) == 3 and np.array_equal(ra[0], rb[0])
ctx.close()
rc = ctx2.lpr_all([0])
assert np.array_equal(rb[0], rc[0])
ctx2.close()
print("sanitize workload ok")
