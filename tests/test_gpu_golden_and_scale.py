"""GPU tests against committed golden vectors (tests/golden/oracle_scores.json, written by
tools/make_golden_scores.py), C-ABI error behaviour, and size-independent properties at the full
BASELINE.json workload size (29.8 M codon columns) where the oracle cannot follow."""
import json
import os

import numpy as np
import pytest

import pcsf_helpers as H
from oracle import oracle as o

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_scores.json")))


def test_golden_oracle_still_agrees_with_committed_vectors(params_base):
    """(runs the CPU oracle; kept with the gpu group because the fixtures exist for the GPU tests)"""
    g = GOLD["examples"]["tal-AA.fa"]
    ps = H.oracle_paramset(params_base, g["paramset"])
    regs, _ = H.example_codes(ps, "tal-AA.fa", frames=g["frames"])
    lpr, elpr = H.oracle_fixed(ps, regs)
    np.testing.assert_allclose(lpr, np.array(g["fixed_lpr"]), rtol=0, atol=1e-9)


@pytest.mark.parametrize("fn", ["tal-AA.fa", "ALDH2.exon5.fa"])
def test_examples_against_committed_golden(params_base, fn):
    from phylocsf_b200 import host
    import phylocsf_b200 as pb

    g = GOLD["examples"][fn]
    ops = H.oracle_paramset(params_base, g["paramset"])
    regs, _ = H.example_codes(ops, fn, frames=g["frames"])  # pleaves only; no scoring by the oracle here
    assert [int(r.shape[0]) for r in regs] == g["ncols"]
    ctx = pb.Context(0)
    host.ParamSet(os.path.join(params_base, "PhyloCSF_Parameters", g["paramset"])).install(ctx)  # product's own eigen
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    assert np.abs(H.DB * (lpr - np.array(g["fixed_lpr"]))).max() < 1e-6
    assert np.abs(H.DB * (elpr - np.array(g["fixed_elpr_anc"]))).max() < 1e-6
    res = [ctx.maximize_lpr(m) for m in (0, 1)]
    for r, row in enumerate(g["mle_rho_lpr_elpr"]):
        for m in (0, 1):
            assert abs(res[m][0][r] - row[m][0]) < 1e-7 * max(1.0, row[m][0])
            assert abs(H.DB * (res[m][1][r] - row[m][1])) < 1e-6
            assert abs(H.DB * (res[m][2][r] - row[m][2])) < 1e-6
    ctx.close()


def test_simulated_batch_against_committed_golden(params_base):
    from phylocsf_b200 import host
    import phylocsf_b200 as pb

    g = GOLD["simulated"]
    regs = [np.frombuffer(bytes.fromhex(h), dtype=np.uint8).reshape(n, 58) for h, n in zip(g["codes_hex"], g["ncols"])]
    ctx = pb.Context(0)
    host.ParamSet(os.path.join(params_base, "PhyloCSF_Parameters", g["paramset"])).install(ctx)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    assert (st == 0).all()
    assert np.abs(H.DB * (lpr - np.array(g["fixed_lpr"]))).max() < 1e-6
    assert np.abs(H.DB * (elpr - np.array(g["fixed_elpr_anc"]))).max() < 1e-6
    ctx.close()


def test_abi_error_behaviour(params_base):
    """Return codes at the points where the reference raises (include/phylocsf_b200.h)."""
    import phylocsf_b200 as pb

    ps = H.oracle_paramset(params_base, "12flies")
    ctx = pb.Context(0)
    with pytest.raises(pb.PcsfError) as e:  # lpr before anything is set
        ctx.nregions = 1
        ctx.lpr_all([0])
    assert e.value.code == -3
    ch = ps.tree.children_array().copy()
    bl = np.array(ps.tree.branches[: ps.tree.root])
    bad = ch.copy()
    bad[0] = 30  # child index >= parent: not a T.t numbering
    with pytest.raises(pb.PcsfError) as e:
        ctx.tree_set(ps.tree.n_leaves, bad, bl)
    assert e.value.code == -1
    nb = bl.copy()
    nb[3] = -0.1  # PhyloModel.make: negative branch length -> Invalid_argument
    with pytest.raises(pb.PcsfError) as e:
        ctx.tree_set(ps.tree.n_leaves, ch, nb)
    assert e.value.code == -1
    H.push_tree(ctx, ps.tree)
    with pytest.raises(pb.PcsfError) as e:  # model never set
        ctx.pt_build(0, [1.0])
    assert e.value.code == -3
    H.push_qdiag(ctx, 0, ps.model.coding_model.q)
    with pytest.raises(pb.PcsfError) as e:  # batch staged, but no P tables for model 0 yet
        ctx.batch_upload(np.array([0, 2]), np.zeros((2, 12), dtype=np.uint8))
        ctx.lpr_all([0])
    assert e.value.code == -3
    ctx.pt_build(0, [1.0])
    with pytest.raises(pb.PcsfError) as e:
        ctx.batch_upload(np.array([0, 3, 2]), np.zeros((2, 12), dtype=np.uint8))  # decreasing offsets
    assert e.value.code == -1
    with pytest.raises(pb.PcsfError) as e:
        ctx.lpr([0], [0], [5])  # region out of range
    assert e.value.code == -1
    with pytest.raises(pb.PcsfError) as e:  # Fit.find_init: lo >= hi
        ctx.maximize_lpr(0, lo=2.0, hi=1.0)
    assert e.value.code == -1
    # codes above 64 are treated as marginalise, like any non-ACGT codon
    ctx.batch_upload(np.array([0, 2]), np.array([[200] * 12, [64] * 12], dtype=np.uint8))
    lpr, _, _ = ctx.lpr_all([0])
    assert abs(lpr[0, 0]) < 1e-9  # all-marginalised columns have likelihood ~1
    ctx.close()
    with pytest.raises(pb.PcsfError):
        pb.Context(10_000)  # no such device


def test_full_size_properties(params_base):
    """BASELINE.json configs[1] at full size (100,000 alignments x 300 nt, 3 frames = 29.8 M codon columns):
    duplicated alignments score bit-identically wherever they sit in the batch; K4's region sums equal the
    sums of K3's per-column terms; simulated-coding alignments outscore simulated-noncoding ones; and a random
    sample of regions matches the oracle."""
    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host, simulate

    A, NC, F = 100_000, 100, 3
    ps = host.ParamSet(os.path.join(params_base, "PhyloCSF_Parameters", "58mammals"))
    ctx = pb.Context(0)
    ps.install(ctx)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    parents = simulate.parents_from_children(ps.n_leaves, ps.children)
    nbr = 2 * ps.n_leaves - 2
    uniq = A - 1000  # the last 1000 alignments are copies of the first 1000
    parts = []
    for w, n in ((0, uniq // 2), (1, uniq - uniq // 2)):
        P = np.stack([ctx.pt_get(w, 0, br) for br in range(nbr)])
        parts.append(simulate.simulate_codes(P, ps.qdiag(w)["prior"], parents, ps.n_leaves, n * NC, gen, dev))
    nt = simulate.codes_to_nt(torch.cat(parts), uniq, NC)
    nt = torch.cat([nt, nt[:1000]]).cpu().numpy()
    L = 3 * NC
    ctx.batch_upload_alignments(np.arange(A, dtype=np.int64) * (ps.n_leaves * L), np.full(A, L, dtype=np.int32), nt, F)
    assert ctx.nregions == A * F and ctx.ncols == A * 298
    lpr, elpr, st = ctx.lpr_all([0, 1])
    assert (st == 0).all() and np.isfinite(lpr).all() and np.isfinite(elpr).all()
    # duplicates: bit-identical, although they fall into different tiles / warps / CTAs
    assert (lpr[:, : 1000 * F] == lpr[:, uniq * F:]).all() and (elpr[:, : 1000 * F] == elpr[:, uniq * F:]).all()
    # K4 against K3's per-column terms
    clz, can = ctx.column_terms(1)
    assert clz.size == A * 298
    assert abs(clz.sum() - lpr[1].sum()) < 1e-9 * abs(lpr[1].sum())
    r = 123456
    lo = (r // F) * 298 + (0, 100, 199)[r % F]
    n = (100, 99, 99)[r % F]
    assert abs(clz[lo:lo + n].sum() - lpr[1, r]) < 1e-9 and abs(can[lo:lo + n].sum() - elpr[1, r]) < 1e-9
    # biology: frame-0 scores separate the two simulated classes
    score = H.DB * (lpr[0] - lpr[1])
    f0 = score[0::F][:uniq]
    assert f0[: uniq // 2].mean() > 500 and f0[uniq // 2:].mean() < -100
    # a random sample against the oracle
    ops = H.oracle_paramset(params_base, "58mammals")
    rng = np.random.default_rng(0)
    lut = np.full(256, 64, dtype=np.int64)
    for ch, i in zip(b"ACGT", range(4)):
        lut[ch] = i
    for r in rng.integers(0, A * F, size=12):
        a, f = divmod(int(r), F)
        idx = lut[nt[a]]  # [n_leaves, L]
        nc = (L - f) // 3
        c = (16 * idx[:, f:f + 3 * nc:3] + 4 * idx[:, f + 1:f + 3 * nc:3] + idx[:, f + 2:f + 3 * nc:3]).T.astype(np.uint8)
        lo_, eo_ = H.oracle_fixed(ops, [np.ascontiguousarray(c)])
        assert abs(H.DB * (lpr[0, r] - lo_[0, 0])) < 1e-6 and abs(H.DB * (lpr[1, r] - lo_[1, 0])) < 1e-6
        assert abs(H.DB * (elpr[0, r] - eo_[0, 0])) < 1e-6
    ctx.close()


def test_mle_batch_properties_120mammals(params_base):
    """BASELINE.json configs[2] at a size the oracle cannot follow (2,000 alignments x 100 codons, 120mammals,
    mle): properties instead. (1) The maximised log-likelihood of every (region, model) is at least the one at
    the starting point rho = 1 (what --strategy=fixed evaluates). (2) A region's result does not depend on what
    else is in the batch: the first 300 regions alone give bit-identical rho and lpr. (3) Duplicated regions give
    bit-identical results. (4) Simulated-coding regions outscore simulated-noncoding ones, and the estimated rho
    follows the scale the region was simulated at."""
    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host, simulate

    N, NC = 2000, 100
    ps = host.ParamSet(os.path.join(params_base, "PhyloCSF_Parameters", "120mammals"))
    ctx = pb.Context(0)
    ps.install(ctx)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    parents = simulate.parents_from_children(ps.n_leaves, ps.children)
    nbr = 2 * ps.n_leaves - 2
    scales = [0.5, 1.6]
    parts = []
    per = (N - 100) // 4
    for w in (0, 1):
        ctx.pt_build(w, scales)
        for si in range(2):
            P = np.stack([ctx.pt_get(w, si, br) for br in range(nbr)])
            parts.append(simulate.simulate_codes(P, ps.qdiag(w)["prior"], parents, ps.n_leaves, per * NC, gen, dev))
    codes = torch.cat(parts).cpu().numpy()
    codes = np.concatenate([codes, codes[: (N - 4 * per) * NC]])  # the tail repeats the head
    off = np.arange(N + 1, dtype=np.int64) * NC
    ctx.batch_upload(off, codes)
    rho, lpr, elpr, st, ne = ctx.maximize_lpr_multi([0, 1])
    assert ((st & ~64) == 0).all() and np.isfinite(lpr).all()
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    l1, _, _ = ctx.lpr_all([0, 1])
    assert (lpr >= l1 - 1e-9 * np.abs(l1)).all()                      # (1)
    dup = N - 4 * per
    assert (rho[:, :dup] == rho[:, 4 * per:]).all() and (lpr[:, :dup] == lpr[:, 4 * per:]).all()   # (3)
    score = H.DB * (lpr[0] - lpr[1])
    assert score[: 2 * per].mean() > 300 and score[2 * per: 4 * per].mean() < -50               # (4)
    for blk, s in ((0, 0.5), (1, 1.6)):
        est = np.median(rho[0, blk * per:(blk + 1) * per])
        assert 0.8 * s < est < 1.25 * s, (s, est)
    ctx.batch_upload(off[:301], codes[: 300 * NC])
    r2, l2, e2, st2, ne2 = ctx.maximize_lpr_multi([0, 1])
    assert (r2 == rho[:, :300]).all() and (l2 == lpr[:, :300]).all() and (ne2 == ne[:, :300]).all()  # (2)
    ctx.close()
