"""GPU parity tests added in round 2 (judge's list): the P(t) failure flags of Q.ml:235-245; the benchmarked
configurations that round 1 left without an oracle comparison at scale - 58mammals with the level-4 subtree tables
in effect (asserted through pcsf_table_level) on >= 500 regions including gapped / missing-species / non-conserved
ones, 120mammals mle on >= 50 full 100-codon regions, omega on the unpruned 100vertebrates tree."""
import os

import numpy as np
import pytest

import pcsf_helpers as H
from oracle import oracle as o

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------------------------
# (e) P(t) failure flags
# ------------------------------------------------------------------------------------------------------------------
def _crafted_models():
    """S = I, lambda = 0 => P(t) = Sinv for every t: any matrix can be put through the fix-ups of Q.ml:226-247."""
    rng = np.random.default_rng(5)
    base = rng.uniform(0.2, 1.0, size=(64, 64))
    base /= base.sum(axis=1, keepdims=True)
    out = {}
    out["ok"] = base.copy()
    m = base.copy()  # clamped entries: inside (-tol, 0) -> 0, no failure
    m[3, 7] = -4e-7
    m[3, 8] += 4e-7 + base[3, 7]
    out["clamp_only"] = m
    m = base.copy()  # entry below -tol -> Failure (Q.ml:235-236); the row still sums to 1
    m[10, 20] = -1e-3
    m[10, 21] += 1e-3 + base[10, 20]
    out["neg_entry"] = m
    m = base.copy()  # row sum off by more than tol -> Failure (Q.ml:243-244); the new diagonal stays inside (0, 1]
    m[40, :] *= 0.99
    out["rowsum"] = m
    m = base.copy()  # off-diagonal mass just above 1 with a diagonal inside (-tol, 0): row sum fine, new diagonal <= 0 -> assert (Q.ml:245)
    row = m[50].copy()
    row[50] = 0.0
    row *= (1.0 + 5e-7) / row.sum()
    row[50] = -5e-7
    m[50] = row
    out["diag_assert"] = m
    return out


def test_pt_failure_flags_match_oracle():
    """PCSF_ST_NEG_ENTRY / ROWSUM / DIAG_ASSERT against oracle_real_to_Pt's codes (Q.ml:235-245), plus the clamp that
    must NOT fail; models that pass give the oracle's P(t) entries."""
    import phylocsf_b200 as pb
    from phylocsf_b200 import _native

    ctx = pb.Context(0)
    # a 3-leaf tree: branches 0..3
    ctx.tree_set(3, np.array([0, 1, 3, 2], dtype=np.int32), np.array([0.1, 0.2, 0.3, 0.4]))
    want_bit = {"ok": 0, "clamp_only": 0, "neg_entry": _native.ST_NEG_ENTRY, "rowsum": _native.ST_ROWSUM, "diag_assert": _native.ST_DIAG_ASSERT}
    oracle_code = {0: 0, 2: _native.ST_NEG_ENTRY, 3: _native.ST_ROWSUM, 4: _native.ST_DIAG_ASSERT}
    eye, lam, prior = np.eye(64), np.zeros(64), np.full(64, 1.0 / 64)
    for name, M in _crafted_models().items():
        ctx.model_set(0, eye, M, lam, prior)
        st = ctx.pt_build(0, [1.0, 2.5], check=False)
        Po = np.empty((64, 64))
        code = o.lib().oracle_real_to_Pt(64, o._dp(np.ascontiguousarray(eye)), o._dp(np.ascontiguousarray(M)), o._dp(lam), 0.25, 1e-6, o._dp(Po))
        assert oracle_code[code] == want_bit[name], (name, code)
        assert (st == want_bit[name]).all(), (name, st)
        if want_bit[name] == 0:
            for br in range(4):
                P = ctx.pt_get(0, 0, br)
                assert np.abs(P - Po).max() < 2e-13 and (P >= 0).all()
        else:  # the error code reaches the caller as PCSF_ERR_NUMERIC (-4) unless it asks for the status words
            with pytest.raises(pb.PcsfError) as e:
                ctx.pt_build(0, [1.0])
            assert e.value.code == -4
    # the pairs entry point (mle / omega candidates) reports per pair
    ms = _crafted_models()
    ctx.models_set(2, np.stack([eye, eye]), np.stack([ms["ok"], ms["rowsum"]]), np.stack([lam, lam]), np.stack([prior, prior]))
    st = ctx.pt_build_pairs([2, 3, 2], [1.0, 1.0, -1.0], check=False)
    assert list(st) == [0, _native.ST_ROWSUM, _native.ST_NEG_T]
    ctx.close()


# ------------------------------------------------------------------------------------------------------------------
# (a)+(b) the headline configuration, level-4 tables asserted, >= 500 regions against the oracle
# ------------------------------------------------------------------------------------------------------------------
def _perturb(nt, rng, lo, hi):
    """Make alignments [lo, hi) of the uint8 [A, n_leaves, L] block look like real data: missing species, gap runs,
    N, lower case, and unrelated substitutions (non-conserved columns)."""
    A, n, L = nt.shape
    for a in range(lo, hi):
        kind = (a - lo) % 5
        blk = nt[a]
        if kind in (0, 4):  # missing species: whole rows of '-'
            rows = rng.choice(np.arange(1, n), size=int(rng.integers(1, n // 2)), replace=False)
            blk[rows] = ord("-")
        if kind in (1, 4):  # gap runs and N
            for _ in range(int(rng.integers(5, 60))):
                r, p, ln = int(rng.integers(0, n)), int(rng.integers(0, L)), int(rng.integers(1, 40))
                blk[r, p:p + ln] = ord("-") if rng.random() < 0.8 else ord("N")
        if kind in (2, 4):  # 10 % random substitutions: code tuples far from the conserved diagonal
            mask = rng.random((n, L)) < 0.10
            blk[mask] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(mask.sum()))]
        if kind == 3:  # lower case + fully random rows for a third of the species
            rows = rng.choice(n, size=n // 3, replace=False)
            blk[rows] = np.frombuffer(b"acgt", dtype=np.uint8)[rng.integers(0, 4, size=(rows.size, L))]


def _frame_codes(nt_a, f):
    lut = np.full(256, -1, dtype=np.int64)
    for ch, i in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut[ch] = i
    idx = lut[nt_a]  # [n_leaves, L]
    L = nt_a.shape[1]
    nc = (L - f) // 3
    i1, i2, i3 = idx[:, f:f + 3 * nc:3], idx[:, f + 1:f + 3 * nc:3], idx[:, f + 2:f + 3 * nc:3]
    c = 16 * i1 + 4 * i2 + i3
    c[(i1 < 0) | (i2 < 0) | (i3 < 0)] = 64
    return np.ascontiguousarray(c.T.astype(np.uint8))


def test_headline_config_level4_against_oracle_540_regions(params_base):
    """BASELINE.json configs[1] at full size through the kernel program bench.py times: wide form, table level 4 -
    ASSERTED through pcsf_table_level / pcsf_last_launch_info, so a silent fallback to level 3 fails the test. 540
    regions against the oracle on all host cores: 300 from alignments with missing species, gap runs, N, lower case and
    10 % unrelated substitutions, 240 drawn from the whole batch."""
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
    gen.manual_seed(19)
    parents = simulate.parents_from_children(ps.n_leaves, ps.children)
    nbr = 2 * ps.n_leaves - 2
    parts = []
    for w, n in ((0, A // 2), (1, A - A // 2)):
        P = np.stack([ctx.pt_get(w, 0, br) for br in range(nbr)])
        parts.append(simulate.simulate_codes(P, ps.qdiag(w)["prior"], parents, ps.n_leaves, n * NC, gen, dev))
    nt = simulate.codes_to_nt(torch.cat(parts), A, NC).cpu().numpy()
    del parts
    torch.cuda.empty_cache()
    rng = np.random.default_rng(23)
    pert = [(2000, 2600), (A // 2 + 3000, A // 2 + 3600)]
    for lo, hi in pert:
        _perturb(nt, rng, lo, hi)
    L = 3 * NC
    ctx.batch_upload_alignments(np.arange(A, dtype=np.int64) * (ps.n_leaves * L), np.full(A, L, dtype=np.int32), nt, F)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    info = ctx.last_launch_info()
    assert ctx.table_level(0) == 4 and ctx.table_level(1) == 4, "level-4 subtree tables not in effect (memory short?)"
    assert info["form"] == "wide" and info["table_level"] == 4 and info["grid"] == 148, info
    assert (st == 0).all() and np.isfinite(lpr).all()
    # the end-to-end entry point bench.py's e2e figure uses gives the same bits
    l2, e2, s2 = ctx.score_alignments(np.arange(A, dtype=np.int64) * (ps.n_leaves * L), np.full(A, L, dtype=np.int32), nt, F, [0, 1])
    assert (l2 == lpr).all() and (e2 == elpr).all()
    sample = np.concatenate([rng.choice(np.arange(lo * F, hi * F), size=150, replace=False) for lo, hi in pert] +
                            [rng.integers(0, A * F, size=240)])
    regs = [_frame_codes(nt[int(r) // F], int(r) % F) for r in sample]
    assert sum((c == 64).any() for c in regs) > 150  # the sample does contain gapped / missing-species regions
    # kinds 2, 3, 4 of _perturb put codons into the columns that are unrelated to their neighbours' (random substitutions)
    unrelated = np.array([any(lo <= int(r) // F < hi and (int(r) // F - lo) % 5 in (2, 3, 4) for lo, hi in pert) for r in sample])
    assert 100 < unrelated.sum() < 300
    ops = H.oracle_paramset(params_base, "58mammals")
    # (1) pruning alone, on all 540 regions: the oracle walks the tree with the very P(t) tables the device holds
    t = ops.tree
    off_s, codes_s = H.regions_to_batch(regs)
    for m in (0, 1):
        pms = np.ascontiguousarray(np.stack([ctx.pt_get(m, 0, br) for br in range(nbr)]))
        prior = np.ascontiguousarray(ps.qdiag(m)["prior"])
        a, b = np.empty(len(regs)), np.empty(len(regs))
        o.lib().oracle_lpr_batch(t.n_leaves, t.children_array().ctypes.data, o._dp(pms), None, o._dp(prior), 64, len(regs),
                                 off_s.ctypes.data, codes_s.ctypes.data, o._dp(a), o._dp(b), H.host_cores())
        d1 = np.abs(H.DB * (lpr[m, sample] - a)).max()
        d2 = np.abs(H.DB * (elpr[m, sample] - b)).max()
        assert d1 < 1e-7 and d2 < 1e-7, ("pruning parity", m, d1, d2)
    # (2) end to end: the oracle with its own P(t) (LAPACK eigensystem, dgemm order) against the product's (Jacobi
    # eigensystem, DMMA order). Regions made of columns the model could have drawn - gaps, missing species and N included -
    # agree to the 1e-6 dB bar. Regions with unrelated substitutions do not and cannot: their likelihood runs through
    # entries of P(t) of 1e-9 and below, which S exp(L t) S^-1 delivers with an ABSOLUTE error of 1e-14, so even the CPU
    # oracle moves by up to 1e-4 dB when only its eigensolver is swapped (conditioning). For those the
    # bar is relative: 2e-9 of |lpr|.
    lo_, eo_ = H.oracle_fixed_batch(ops, regs)
    d = np.abs(H.DB * (lpr[:, sample] - lo_))
    da = np.abs(H.DB * (elpr[:, sample] - eo_))
    assert d[:, ~unrelated].max() < 1e-6 and da[:, ~unrelated].max() < 1e-6, (d[:, ~unrelated].max(), da[:, ~unrelated].max())
    rel = d[:, unrelated] / np.abs(H.DB * lo_[:, unrelated])
    assert rel.max() < 2e-9, rel.max()
    print("headline sample: model-like max |d lpr| %.2e dB; unrelated-substitution regions max %.2e dB (relative %.2e)"
          % (d[:, ~unrelated].max(), d[:, unrelated].max(), rel.max()))
    ctx.close()


# ------------------------------------------------------------------------------------------------------------------
# (c) config 3: 120mammals mle against the oracle on 56 full 100-codon regions
# ------------------------------------------------------------------------------------------------------------------
def test_mle_120mammals_56_full_regions_vs_oracle(params_base):
    ps = H.oracle_paramset(params_base, "120mammals", strategy="mle")
    rng = np.random.default_rng(31)
    regs = []
    for i in range(56):
        inst = ps.model.coding_model if i % 2 == 0 else ps.model.noncoding_model
        rho = float(np.exp(rng.uniform(np.log(0.3), np.log(3.0))))  # log-uniform in [0.3, 3]
        c = o.simulate_columns(inst.model(rho), 100, rng)
        if i % 7 == 3:
            c[:, rng.choice(120, size=30, replace=False)] = 64  # missing species
        if i % 7 == 5:
            c[rng.random(c.shape) < 0.05] = 64  # scattered gaps
        regs.append(c)
        inst.q._memo.clear()
    ora = H.oracle_mle_parallel(params_base, "120mammals", regs)
    ctx = H.make_context(ps)
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    rho, lpr, elpr, st, ne = ctx.maximize_lpr_multi([0, 1])
    worst, same_path, forked = 0.0, 0, []
    for r in range(len(regs)):
        for m in (0, 1):
            ox, olp, oel, oit, otries = ora[r][m]
            assert (st[m, r] & ~64) == 0
            if ne[m, r] != 3 + otries + 3 + 1 + oit + 1:
                # The two searches took different paths: Brent's accept/reject tests compare likelihoods that differ
                # by rounding between any two implementations, so a near-tie can go either way (the reference would
                # fork the same way against another GSL / libm build). Both ends satisfy the stop rule: rho within the
                # 1 % bracket of each other; the maximised likelihood is flat there.
                forked.append((r, m, int(ne[m, r]), oit))
                assert abs(rho[m, r] - ox) < 0.02 * ox and abs(H.DB * (lpr[m, r] - olp)) < 0.02, (r, m, rho[m, r], ox)
                continue
            same_path += 1
            # rho is a Brent iterate: parabolic steps amplify the 1e-13 relative differences between the two likelihood
            # evaluations (measured: up to 3e-7 relative over these regions); the stop rule only asks for 1 % anyway.
            assert abs(rho[m, r] - ox) < 1e-5 * max(1.0, ox), (r, m, rho[m, r], ox)
            worst = max(worst, abs(H.DB * (lpr[m, r] - olp)), abs(H.DB * (elpr[m, r] - oel)))
    assert same_path >= 2 * len(regs) - 3, forked  # at most a few near-ties in 112 searches
    # The returned lpr is the likelihood AT the final iterate, which sits up to 1 % away from the optimum, where the slope is
    # a few nats per unit of rho: a 3e-7 shift of the iterate moves it by ~2e-6 dB (measured worst over these 112 searches:
    # 1.9e-6). That is the conditioning of the reference's own procedure, not evaluation error - shown by (2) below.
    assert worst < 1e-5, worst
    score = H.DB * (lpr[0] - lpr[1])
    want = np.array([H.DB * (ora[r][0][1] - ora[r][1][1]) for r in range(len(regs))])
    ok = np.ones(len(regs), dtype=bool)
    ok[[r for r, _, _, _ in forked]] = False
    assert np.abs(score - want)[ok].max() < 1e-5
    # (2) the same likelihoods evaluated at the ORACLE's final rho: no search in between, the 1e-6 dB bar holds
    R = len(regs)
    ctx.pt_build_pairs(np.repeat([0, 1], R), np.array([ora[r][m][0] for m in (0, 1) for r in range(R)]))
    l2, e2, s2 = ctx.lpr_pairs(np.arange(2 * R), np.tile(np.arange(R), 2))
    olp = np.array([ora[r][m][1] for m in (0, 1) for r in range(R)])
    oel = np.array([ora[r][m][2] for m in (0, 1) for r in range(R)])
    d_at = max(np.abs(H.DB * (l2 - olp)).max(), np.abs(H.DB * (e2 - oel)).max())
    assert (s2 == 0).all() and d_at < 1e-6, d_at
    print("120mammals mle: %d of %d searches on the oracle's path (forked: %s); worst |d lpr| at the device's own iterate %.2e dB, at the oracle's rho %.2e dB"
          % (same_path, 2 * len(regs), forked, worst, d_at))
    ctx.close()


# ------------------------------------------------------------------------------------------------------------------
# (d) config 4: omega on the unpruned 100vertebrates tree, --allScores, 3 frames
# ------------------------------------------------------------------------------------------------------------------
def _full_tree_exons(ops, rng, spec):
    """Simulated exons on all species of the tree: [(codes [ncodons, n], rows)], the last with missing species."""
    out = []
    for k, (ncod, rho) in enumerate(spec):
        inst = ops.model.coding_model if k != 1 else ops.model.noncoding_model
        codes = o.simulate_columns(inst.model(rho), ncod, rng)
        if k == 2:
            codes[:, [i for i in range(1, codes.shape[1]) if i % 9 == 0]] = 64  # some species missing
        out.append((codes, o.codes_to_alignment(codes)))
    return out


def test_omega_full_100vertebrates_tree_cli_vs_oracle(params_base, tmp_path):
    """Two simulated exons (96 and 240 nt) on all 100 species through the command line's omega strategy against the
    oracle's restatement of OmegaModel.score (src/OmegaModel.ml:195-219), line for line, 3 frames, --allScores, with the
    diagnostics (--debug): rho and kappa of both hypotheses to the two printed decimals, scores to 1.5e-4 dB."""
    from test_cli import run_cli, same_lines

    ops = o.load_paramset(os.path.join(params_base, "PhyloCSF_Parameters", "100vertebrates"), o.Options(strategy="fixed"))
    labels = ops.tree.labels[: ops.tree.n_leaves]
    assert len(labels) == 100
    files, named = [], []
    for k, (codes, rows) in enumerate(_full_tree_exons(ops, np.random.default_rng(41), ((32, 1.0), (80, 0.6)))):
        path = tmp_path / ("exon%d.fa" % k)
        path.write_text("".join(">%s\n%s\n" % (l, r) for l, r in zip(labels, rows)))
        files.append(str(path))
        named.append((str(path), path.read_text().split("\n")[:-1]))
    want = H.oracle_lines_parallel(params_base, "100vertebrates", named, strategy="omega", frames=3, all_scores=True, debug=True)
    got = run_cli(params_base, "100vertebrates", files, "--strategy=omega", "--frames=3", "--allScores", "--debug")
    same_lines(got, [l for lines in want for l in lines])


def test_omega_full_100vertebrates_tree_abi_vs_oracle(params_base):
    """config 4 through pcsf_omega_score (what --strategy=omega runs) at full precision: three exons (32, 80, 140 codons,
    the last with missing species) x 3 frames on the unpruned 100-leaf tree. Score and both maximised log-posteriors to
    1e-6 dB; rho and kappa of both hypotheses to 1e-4 relative (they are Brent iterates at the end of six coordinate
    searches: the same path is taken on both sides, but parabolic steps amplify rounding-level differences between the
    device's Jacobi eigensystem and the oracle's LAPACK one - measured up to 3e-6)."""
    import phylocsf_b200 as pb
    from phylocsf_b200 import host

    ops = o.load_paramset(os.path.join(params_base, "PhyloCSF_Parameters", "100vertebrates"), o.Options(strategy="fixed"))
    regs = []
    for codes, rows in _full_tree_exons(ops, np.random.default_rng(43), ((32, 1.0), (80, 0.6), (140, 1.4))):
        nt = np.array([list(r.encode()) for r in rows], dtype=np.uint8)
        for f in range(3):
            regs.append(_frame_codes(nt, f))
    want = H.oracle_omega_parallel(params_base, "100vertebrates", regs)
    ctx = pb.Context(0)
    H.push_tree(ctx, ops.tree)
    off, codes = H.regions_to_batch(regs)
    score, diag, st = host.omega_score(ctx, off, codes)
    assert (st == 0).all()
    for r, w in enumerate(want):
        assert abs(score[r] - w[0]) < 1e-6, (r, score[r], w[0])
        assert abs(diag[r, 0] - H.DB * w[1]) < 1e-6 and abs(diag[r, 5] - H.DB * w[4]) < 1e-6
        for got, exp in ((diag[r, 1], w[2]), (diag[r, 2], w[3]), (diag[r, 6], w[5]), (diag[r, 7], w[6])):
            assert abs(got - exp) < 1e-4 * max(1.0, abs(exp)), (r, got, exp)
        assert diag[r, 3] == 1.0 and diag[r, 4] == 1.0 and diag[r, 8] == 0.2 and diag[r, 9] == 0.01
    ctx.close()


def test_omega_eigen_warm_start_matches_cold(params_base):
    """K5 with warm starts (pcsf_omega_models_set_cached: the previous candidate's eigenvectors start the next
    diagonalisation of the same region) against cold starts: the same eigensystem of the same matrix - S diag(lambda) S^-1
    reproduces the oracle's Q, S S^-1 = I, the equilibrium prior and the likelihood agree to rounding - in fewer sweeps."""
    ps = H.oracle_paramset(params_base, "12flies")
    ctx = H.make_context(ps)
    regs, _ = H.example_codes(ps, "tal-AA.fa")
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    rng = np.random.default_rng(8)
    base = np.array([[2.5, 1.0, 1.0] + [1.0] * 9, [1.7, 0.2, 0.01, 1.3, 0.8, 1.1, 0.9, 1.2, 0.7, 1.05, 0.95, 1.4],
                     [4.0, 0.5, 0.3] + list(rng.uniform(0.3, 3.0, 9))])
    kappas = [2.5, 1.0, 10.0, 4.4377, 3.9, 4.02, 4.011, 4.0101]  # a search closing in, as Brent's does
    ctx.omega_cache_reset(3)
    ctx.counters(reset=True)
    warm_sweeps = []
    for kappa in kappas:
        qs = base.copy()
        qs[:, 0] = kappa
        st = ctx.omega_models_set_cached(3, qs, [0, 1, 2])
        assert (st == 0).all()
        warm = [ctx.model_get(3 + i) for i in range(3)]
        ctx.pt_build_pairs([3, 4, 5], [1.0, 0.6, 1.7])
        lw = ctx.lpr_pairs([0, 1, 2], [0, 0, 0])[0]
        warm_sweeps.append(ctx.counters(reset=True)["eig_sweeps"])
        st = ctx.omega_models_set(3, qs)  # cold
        cold = [ctx.model_get(3 + i) for i in range(3)]
        ctx.pt_build_pairs([3, 4, 5], [1.0, 0.6, 1.7])
        lc = ctx.lpr_pairs([0, 1, 2], [0, 0, 0])[0]
        cold_sweeps = ctx.counters(reset=True)["eig_sweeps"]
        for i in range(3):
            Qo = o.omega_q(list(qs[i]))
            d = warm[i]
            np.testing.assert_allclose(d["S"] @ np.diag(d["lam"]) @ d["Sinv"], Qo, atol=5e-13)
            np.testing.assert_allclose(d["S"] @ d["Sinv"], np.eye(64), atol=1e-12)
            np.testing.assert_allclose(np.sort(d["lam"]), np.sort(cold[i]["lam"]), atol=1e-12)
            np.testing.assert_allclose(d["prior"], cold[i]["prior"], atol=1e-14)
        assert np.abs(H.DB * (lw - lc)).max() < 1e-6, (kappa, lw, lc)
    assert warm_sweeps[0] == cold_sweeps or warm_sweeps[0] >= 3 * 7   # the first solve of a slot is a cold one
    assert warm_sweeps[-1] < 0.6 * cold_sweeps, (warm_sweeps, cold_sweeps)  # nearby candidates converge in a few sweeps
    print("K5 sweeps per 3 matrices: warm", warm_sweeps, "cold", cold_sweeps)
    with pytest.raises(Exception):
        ctx.omega_models_set_cached(3, base, [0, 1, 7])  # slot out of range
    ctx.close()


@pytest.mark.parametrize("kappa", [1.0, 2.7])
def test_omega_pipeline_against_the_beagle_pinned_nucleotide_model(kappa):
    """K5 (Q assembly + Jacobi) -> K1 -> pruning on the BEAGLE test's tree and sequences with omega = sigma = 1, uniform
    F3x4, tree scale 3: the codon likelihood must equal the product of the three nucleotide columns' K80 / JC69 likelihoods
    (oracle, 4 states - the configuration whose lnL the reference pins, lib/CamlPaml/test.ml:81). The one corner of the
    omega pipeline that is tied to a reference-held vector rather than to the oracle's reading of OmegaModel.ml."""
    import json

    import phylocsf_b200 as pb
    from test_oracle_golden import _k80, gapfree_codon_columns

    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "beagle_tiny.json")))
    t = o.Tree.of_newick(o.newick_parse(d["newick"]))
    cod, nuc = gapfree_codon_columns([d["human"], d["chimp"], d["gorilla"]])
    ll_nt = o.lpr_columns(o.PhyloModel(t, o.QDiag(_k80(kappa)), t.branches), nuc)[0]
    ctx = pb.Context(0)
    H.push_tree(ctx, t)
    st = ctx.omega_models_set(0, [[kappa, 1.0, 1.0] + [1.0] * 9])
    assert (st == 0).all()
    np.testing.assert_allclose(ctx.model_get(0)["prior"], np.full(64, 1.0 / 64), atol=1e-14)
    ctx.pt_build_pairs([0], [3.0])
    ctx.batch_upload(np.array([0, cod.shape[0]], dtype=np.int64), cod)
    lpr, _, st = ctx.lpr_pairs([0], [0])
    assert st[0] == 0 and abs(lpr[0] - ll_nt) < 1e-9 * abs(ll_nt), (lpr[0], ll_nt)
    ctx.close()


def test_subtree_tables_are_shared_between_contexts_and_follow_the_cumulative_rule(params_base, monkeypatch):
    """Two contexts on one GPU with the same tree, model and scale use ONE set of subtree tables (the second context
    launches no table kernels), results are bit-identical to a context that builds its own, the block outlives the
    context that built it, and a P set whose batches are each too small for a level gets it once it has scored enough
    columns in total."""
    import phylocsf_b200 as pb
    from phylocsf_b200 import host

    hp = host.ParamSet(os.path.join(params_base, "PhyloCSF_Parameters", "29mammals"))
    ops = H.oracle_paramset(params_base, "29mammals")
    rng = np.random.default_rng(77)
    regs = [o.simulate_columns(ops.model.coding_model.model(1.0), 300, rng), o.simulate_columns(ops.model.noncoding_model.model(1.0), 200, rng)]
    off, codes = H.regions_to_batch(regs)

    def fresh():
        c = pb.Context(0)
        hp.install(c)
        c.pt_build(0, [1.0])
        c.pt_build(1, [1.0])
        c.option_set(pb.Context.OPT_PRUNE_FORM, pb.Context.FORM_WIDE)
        c.batch_upload(off, codes)
        return c

    a, b = fresh(), fresh()
    a.option_set(pb.Context.OPT_CHERRY_TABLES, 2)  # always, up to level 3
    b.option_set(pb.Context.OPT_CHERRY_TABLES, 2)
    l0 = a.launch_count
    ra = a.lpr_all([0, 1])
    built = a.launch_count - l0
    l0 = b.launch_count
    rb = b.lpr_all([0, 1])
    reused = b.launch_count - l0
    assert a.table_level(0) == 3 and b.table_level(0) == 3 and b.last_launch_info()["table_level"] == 3
    assert built > reused and reused == 3, (built, reused)  # pruning + segments + reduction, no table kernels
    assert all(np.array_equal(x, y) for x, y in zip(ra, rb))
    a.close()  # the block lives on while b holds it
    rb2 = b.lpr_all([0, 1])
    assert all(np.array_equal(x, y) for x, y in zip(rb, rb2))
    monkeypatch.setenv("PCSF_SHARE_TABLES", "0")
    c = fresh()
    c.option_set(pb.Context.OPT_CHERRY_TABLES, 2)
    rc = c.lpr_all([0, 1])
    assert c.table_level(0) == 3 and all(np.array_equal(x, y) for x, y in zip(rb, rc))
    c.option_set(pb.Context.OPT_CHERRY_TABLES, 1)  # never: the plain program, same bits
    rc1 = c.lpr_all([0, 1])
    assert c.last_launch_info()["table_level"] == 0 and all(np.array_equal(x, y) for x, y in zip(rb, rc1))
    c.close()
    b.close()
    monkeypatch.delenv("PCSF_SHARE_TABLES")
    # cumulative rule: 500-column batches never reach the 50,000-column threshold of level 2 in one call; after 400,000
    # columns in total the P set gets its cherry tables (and keeps giving the same bits)
    d = fresh()
    first = d.lpr_all([0, 1])
    assert d.table_level(0) == 0
    for _ in range(900):
        d.lpr_all([0])
        if d.table_level(0):
            break
    assert d.table_level(0) == 2
    again = d.lpr_all([0, 1])
    assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])
    d.close()


def _all_sets():
    from tools import golden_params as gp

    return gp.set_names()


@pytest.mark.parametrize("pset", _all_sets())
def test_fixed_and_posteriors_on_every_shipped_parameter_set(params_base, pset):
    """All 14 shipped parameter sets (7 to 120 species; 20flies has a zero-length branch, 23flies / 26worms / 7yeast
    frequencies that sum to 1 +- 2e-6): fixed-strategy scores in both kernel forms and the outside pass against the oracle."""
    import phylocsf_b200 as pb

    ps = H.oracle_paramset(params_base, pset)
    n = ps.tree.n_leaves
    rng = np.random.default_rng(sum(map(ord, pset)))
    regs = [o.simulate_columns(ps.model.coding_model.model(1.0), 21, rng), o.simulate_columns(ps.model.noncoding_model.model(1.0), 12, rng)]
    regs[0][2, n // 2:] = 64
    lo, eo = H.oracle_fixed(ps, regs)
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    for form in (pb.Context.FORM_NARROW, pb.Context.FORM_WIDE):
        ctx.option_set(pb.Context.OPT_PRUNE_FORM, form)
        lpr, elpr, st = ctx.lpr_all([0, 1])
        assert (st == 0).all()
        assert np.abs(H.DB * (lpr - lo)).max() < 1e-7 and np.abs(H.DB * (elpr - eo)).max() < 1e-7
    post, ec, z = ctx.posteriors(1, 0, nodes=[2 * n - 2, n])
    zo, po, eco = o.posteriors_columns(ps.model.noncoding_model.model(1.0), codes)
    np.testing.assert_allclose(z, zo, rtol=1e-10, atol=0)
    np.testing.assert_allclose(post[0], po[:, 2 * n - 2], atol=2e-12)
    np.testing.assert_allclose(ec, eco, atol=1e-10)
    ctx.close()


def test_negative_offsets_are_rejected(params_base):
    """Round-1 advisor finding: a negative aln_off made the library read host memory before the caller's buffer."""
    import phylocsf_b200 as pb

    ps = H.oracle_paramset(params_base, "12flies")
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    nt = np.full((2, 12, 30), ord("A"), dtype=np.uint8)
    for call in (lambda: ctx.batch_upload_alignments([0, -360], [30, 30], nt, 3),
                 lambda: ctx.score_alignments([0, -360], [30, 30], nt, 3, [0]),
                 lambda: ctx.batch_upload_alignments([0, 360], [30, -1], nt, 3)):
        with pytest.raises(pb.PcsfError) as e:
            call()
        assert e.value.code == -1
    ctx.close()


@pytest.mark.parametrize("pset", ["12flies", "29mammals", "120mammals"])
def test_k0_codes_bit_exact_against_oracle_pleaves(params_base, pset):
    """K0 on the device (pcsf_k0.cuh) against the oracle's pleaves (src/PhyloCSF.ml:219-246, Code.ml:39-51), byte for
    byte through pcsf_batch_codes_get: ragged lengths (0, 1, 2 ... several shared-memory tiles), alignments at any byte
    offset of the buffer, 1 / 3 / 6 frames, 12 / 29 / 120 leaves (fewer and more than one 16-byte group per column).
    The same inputs run through the CPU emulation of the kernel in tests/test_k0_emulation.py."""
    import test_k0_emulation as E

    ps = H.oracle_paramset(params_base, pset)
    n = ps.tree.n_leaves
    ctx = H.make_context(ps)
    rng = np.random.default_rng(n)
    alphabet = np.array(list("ACGTacgtNn-"))
    lens = [0, 1, 2, 3, 4, 5, 15, 16, 17, 18, 47, 48, 49, 50, 97, 300, 301, 302, 333, 2500 if n > 100 else 5001]
    alns = [["".join(alphabet[rng.integers(0, len(alphabet), size=L)]) for _ in range(n)] for L in lens]
    nt, off = E.pack(alns, rng)
    for frames in (1, 3, 6):
        want, roff = E.oracle_frames(alns, frames)
        ctx.batch_upload_alignments(off, lens, nt, frames)
        assert ctx.nregions == frames * len(lens) and ctx.ncols == int(roff[-1])
        got = ctx.batch_codes()
        assert np.array_equal(got, want), (pset, frames, np.argwhere(got != want)[:5])
    # a second, shorter batch into the same (larger) buffers, and an empty one
    want, roff = E.oracle_frames(alns[5:9], 6)
    nt2, off2 = E.pack(alns[5:9], rng)
    ctx.batch_upload_alignments(off2, lens[5:9], nt2, 6)
    assert np.array_equal(ctx.batch_codes(), want)
    ctx.batch_upload_alignments(np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.uint8), 3)
    assert ctx.ncols == 0 and ctx.batch_codes().shape == (0, n)
    ctx.close()
