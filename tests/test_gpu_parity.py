"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on the same inputs. Bars: P(t) entries to 2e-13 absolute; region scores to
|delta| <= 1e-6 decibans (BASELINE.json north_star), asserted at 1e-7 to leave margin."""
import numpy as np
import pytest

from oracle import oracle as o
import pcsf_helpers as H

pytestmark = pytest.mark.gpu

TOL_DB = 1e-7


@pytest.fixture(scope="module")
def flies(params_base):
    return H.oracle_paramset(params_base, "12flies")


@pytest.fixture(scope="module")
def mammals58(params_base):
    return H.oracle_paramset(params_base, "58mammals")


@pytest.mark.parametrize("pset", ["12flies", "58mammals", "20flies"])
def test_pt_build_matches_oracle(params_base, pset):
    ps = H.oracle_paramset(params_base, pset)
    ctx = H.make_context(ps)
    scales = np.array([1.0, 0.01, 0.37, 10.0])
    for mid, inst in enumerate((ps.model.coding_model, ps.model.noncoding_model)):
        st = ctx.pt_build(mid, scales)
        assert (st == 0).all()
        for si, rho in enumerate(scales):
            for br in range(ps.tree.root):
                P = ctx.pt_get(mid, si, br)
                Po = inst.q.to_Pt(rho * ps.tree.branches[br])
                assert np.abs(P - Po).max() < 2e-13, (pset, mid, rho, br)
                assert (P >= 0).all()
    ctx.close()


def test_pt_build_negative_scale_flags(flies):
    ctx = H.make_context(flies)
    st = ctx.pt_build(0, [1.0, -1.0], check=False)
    assert st[0] == 0 and (st[1] & 1)
    ctx.close()


def test_fixed_tal_AA(flies):
    regs, _ = H.example_codes(flies, "tal-AA.fa")
    ctx = H.make_context(flies)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    lo, eo = H.oracle_fixed(flies, regs)
    assert (st == 0).all()
    score = H.DB * (lpr[0] - lpr[1])
    assert abs(score[0] - 361.6876) < 1e-4  # BASELINE.json configs[0] (restatement-derived value)
    assert np.abs(H.DB * (lpr - lo)).max() < TOL_DB
    assert np.abs(H.DB * (elpr - eo)).max() < TOL_DB
    # per-column terms against the oracle's
    clz, can = ctx.column_terms(0)
    _, _, oz, oa = o.lpr_columns(flies.model.coding_model.model(1.0), regs[0])
    np.testing.assert_allclose(clz, oz, rtol=0, atol=1e-12)
    np.testing.assert_allclose(can, oa, rtol=0, atol=1e-12)
    ctx.close()


def test_fixed_simulated_58mammals_ragged(mammals58):
    ps = mammals58
    rng = np.random.default_rng(42)
    mc, mn = ps.model.coding_model.model(1.0), ps.model.noncoding_model.model(1.0)
    regs = []
    for i, n in enumerate([100, 99, 99, 1, 0, 7, 128, 129, 300, 16, 15, 17, 0, 64]):
        c = o.simulate_columns(mc if i % 2 == 0 else mn, n, rng) if n else np.zeros((0, 58), dtype=np.uint8)
        regs.append(c)
    # gaps / missing species / whole-column marginalisation
    regs[0][:, 5] = 64
    regs[1][3, :] = 64
    regs[2][10:20, 30:] = 64
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    lo, eo = H.oracle_fixed(ps, regs)
    assert (st == 0).all()
    assert np.abs(H.DB * (lpr - lo)).max() < TOL_DB
    assert np.abs(H.DB * (elpr - eo)).max() < TOL_DB
    assert lpr[0, 4] == 0.0 and elpr[1, 12] == 0.0  # empty regions
    # the explicit-evaluation entry point gives bit-identical numbers
    em = np.repeat([0, 1], len(regs))
    er = np.tile(np.arange(len(regs)), 2)
    l2, e2, _ = ctx.lpr(em, np.zeros_like(em), er)
    assert (l2.reshape(2, -1) == lpr).all() and (e2.reshape(2, -1) == elpr).all()
    ctx.close()


@pytest.mark.parametrize("pset", ["58mammals", "120mammals", "12flies"])
def test_both_kernel_forms_match_oracle_and_each_other(params_base, pset):
    """The narrow (128-column tiles) and the wide (192-column tiles, leaf messages read from the staged P^T
    table, thread-owned parking) form of the pruning kernel, each forced with PCSF_OPT_PRUNE_FORM: both within
    1e-7 dB of the oracle and bit-identical to each other, per region and per column, on ragged region lengths
    around both tile sizes with gaps and missing species; evaluated twice to catch run-to-run differences."""
    import phylocsf_b200 as pb

    ps = H.oracle_paramset(params_base, pset)
    n = ps.tree.n_leaves
    rng = np.random.default_rng(17)
    mc, mn = ps.model.coding_model.model(1.0), ps.model.noncoding_model.model(1.0)
    lens = [191, 192, 193, 1, 0, 127, 128, 129, 385, 16, 15, 17, 640, 64]
    if pset == "120mammals":
        lens = [193, 1, 129, 40]  # keeps the CPU oracle affordable
    regs = []
    for i, L in enumerate(lens):
        regs.append(o.simulate_columns(mc if i % 2 == 0 else mn, L, rng) if L else np.zeros((0, n), dtype=np.uint8))
    regs[0][:, n // 3] = 64
    regs[2][5, :] = 64
    regs[2][20:40, n // 2:] = 64
    lo, eo = H.oracle_fixed(ps, regs)
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    got = {}
    # further variants: the wide form running the table programs (a cherry - or a cherry and the leaf next to it - and
    # the contractions above them as one lookup in a table built with the kernel's own instruction sequence)
    for form, cherry in ((pb.Context.FORM_NARROW, 1), (pb.Context.FORM_WIDE, 1), ("pairs", 3), ("triples", 2)):
        ctx.option_set(pb.Context.OPT_PRUNE_FORM, pb.Context.FORM_WIDE if isinstance(form, str) else form)
        ctx.option_set(pb.Context.OPT_CHERRY_TABLES, cherry)
        for rep in range(2):
            lpr, elpr, st = ctx.lpr_all([0, 1])
            cols = [ctx.column_terms(m) for m in (0, 1)]
            assert (st == 0).all()
            assert np.abs(H.DB * (lpr - lo)).max() < TOL_DB and np.abs(H.DB * (elpr - eo)).max() < TOL_DB
            key = (lpr.copy(), elpr.copy(), [c[0].copy() for c in cols], [c[1].copy() for c in cols])
            if form in got:
                prev = got[form]
                assert (prev[0] == key[0]).all() and (prev[1] == key[1]).all()
                assert all((a == b).all() for a, b in zip(prev[2], key[2]))
            got[form] = key
        # one span per (model, region): tiles start at region starts
        em = np.repeat([0, 1], len(regs))
        er = np.tile(np.arange(len(regs)), 2)
        l2, e2, _ = ctx.lpr(em, np.zeros_like(em), er)
        assert (l2.reshape(2, -1) == got[form][0]).all() and (e2.reshape(2, -1) == got[form][1]).all()
    a = got[pb.Context.FORM_NARROW]
    for other in (pb.Context.FORM_WIDE, "pairs", "triples"):
        b = got[other]
        assert all((x == y).all() for x, y in zip(a[2], b[2])) and all((x == y).all() for x, y in zip(a[3], b[3])), other  # per column
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), other
    ctx.option_set(pb.Context.OPT_PRUNE_FORM, pb.Context.FORM_AUTO)
    ctx.option_set(pb.Context.OPT_CHERRY_TABLES, 0)
    with pytest.raises(Exception):
        ctx.option_set(pb.Context.OPT_PRUNE_FORM, 3)
    ctx.close()


@pytest.mark.parametrize("pset", ["7yeast", "29mammals"])
def test_level4_subtree_tables_bit_identical(params_base, pset):
    """Level 4 of the subtree tables (a caterpillar of four leaves memoised over the 65^4 code quadruples, 9.1 GB per
    subtree and P set - forced here; by default only for P sets that score >= 5 M columns and while memory allows):
    bit-identical to the plain wide form and within 1e-7 dB of the oracle, with gaps and marginalised species."""
    import phylocsf_b200 as pb

    ps = H.oracle_paramset(params_base, pset)
    n = ps.tree.n_leaves
    rng = np.random.default_rng(29)
    mc, mn = ps.model.coding_model.model(1.0), ps.model.noncoding_model.model(1.0)
    regs = [o.simulate_columns(mc, 300, rng), o.simulate_columns(mn, 193, rng), o.simulate_columns(mc, 7, rng)]
    regs[0][::7, :] = rng.integers(0, 65, size=regs[0][::7, :].shape)  # unrelated codes and gaps: rows far from the diagonal
    regs[1][5, :] = 64
    lo, eo = H.oracle_fixed(ps, regs)
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    ctx.option_set(pb.Context.OPT_PRUNE_FORM, pb.Context.FORM_WIDE)
    out = {}
    for mode in (1, 4):
        ctx.option_set(pb.Context.OPT_CHERRY_TABLES, mode)
        lpr, elpr, st = ctx.lpr_all([0, 1])
        out[mode] = (lpr.copy(), elpr.copy(), [ctx.column_terms(m)[0].copy() for m in (0, 1)])
        finite = np.isfinite(lo)
        assert np.abs(H.DB * (lpr - lo))[finite].max() < TOL_DB and (np.isfinite(lpr) == finite).all()
    assert (out[1][0] == out[4][0]).all() or np.array_equal(out[1][0], out[4][0], equal_nan=True)
    assert np.array_equal(out[1][1], out[4][1], equal_nan=True)
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out[1][2], out[4][2]))
    ctx.close()


@pytest.mark.parametrize("pset", ["120mammals", "100vertebrates", "29mammals", "7yeast"])
def test_fixed_other_trees(params_base, pset):
    ps = H.oracle_paramset(params_base, pset)
    rng = np.random.default_rng(7)
    mc, mn = ps.model.coding_model.model(1.0), ps.model.noncoding_model.model(1.0)
    regs = [o.simulate_columns(mc, 40, rng), o.simulate_columns(mn, 33, rng)]
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    lo, eo = H.oracle_fixed(ps, regs)
    assert np.abs(H.DB * (lpr - lo)).max() < TOL_DB
    assert np.abs(H.DB * (elpr - eo)).max() < TOL_DB
    ctx.close()


def test_underflow_matches_reference_semantics(params_base):
    """The reference does not rescale partials (PhyloLik.ml:87-92): uniform-random columns on the
    120-leaf tree underflow to z = 0 => log z = -inf, posterior zeros (PhyloLik.ml:131-132)."""
    ps = H.oracle_paramset(params_base, "120mammals")
    rng = np.random.default_rng(3)
    regs = [rng.integers(0, 64, size=(5, 120)).astype(np.uint8)]
    lo, eo = H.oracle_fixed(ps, regs)
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    assert np.isneginf(lo).all() and np.isneginf(lpr).all()
    assert (st & 16).all()
    assert (elpr == eo).all()
    ctx.close()


def test_device_pleaves_frames(params_base):
    """pcsf_batch_upload_alignments (on-device pleaves, 6 frames) against the oracle's pleaves."""
    ps = H.oracle_paramset(params_base, "29mammals")
    regs, (aln, leaf_ord) = H.example_codes(ps, "ALDH2.exon5.fa", frames=6)
    n, L = ps.tree.n_leaves, len(aln[0])
    nt = np.full((n, L), ord("-"), dtype=np.uint8)
    for l, r in enumerate(leaf_ord):
        if r is not None:
            nt[l] = np.frombuffer(aln[r].encode(), dtype=np.uint8)
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    ctx.batch_upload_alignments([0], [L], nt, 6)
    assert ctx.nregions == 6 and ctx.ncols == sum(r.shape[0] for r in regs)
    lpr, elpr, st = ctx.lpr_all([0, 1])
    lo, eo = H.oracle_fixed(ps, regs)
    assert np.abs(H.DB * (lpr - lo)).max() < TOL_DB
    assert np.abs(H.DB * (elpr - eo)).max() < TOL_DB
    score = H.DB * (lpr[0] - lpr[1])
    assert int(np.argmax(score)) == 1  # frame +1 wins (src/test.ml:41-49 under mle; same frame under fixed)
    ctx.close()


def test_device_pleaves_from_parts_equals_one_buffer(params_base):
    """pcsf_batch_upload_alignments_parts (nucleotide buffer handed over in pieces, as the command line's
    reader threads produce it) stages exactly what pcsf_batch_upload_alignments stages: identical scores,
    for alignments of different lengths spread over pieces of different sizes (one piece empty)."""
    ps = H.oracle_paramset(params_base, "12flies")
    n = ps.tree.n_leaves
    rng = np.random.default_rng(3)
    lens = [30, 7, 2, 91, 45]
    alns = [np.frombuffer(b"ACGTacgtN-", dtype=np.uint8)[rng.integers(0, 10, size=(n, L))] for L in lens]
    flat = np.concatenate([a.reshape(-1) for a in alns])
    off = np.concatenate([[0], np.cumsum([a.size for a in alns])[:-1]])
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    ctx.batch_upload_alignments(off, lens, flat, 6)
    one = ctx.lpr_all([0, 1])
    cut1, cut2 = int(off[2]), int(off[4])
    parts = [flat[:cut1], flat[:0], flat[cut1:cut2], flat[cut2:]]
    ctx.batch_upload_alignments_parts(off, lens, parts, 6)
    assert ctx.nregions == 6 * len(lens)
    many = ctx.lpr_all([0, 1])
    for a, b in zip(one, many):
        assert np.array_equal(a, b, equal_nan=True)
    with pytest.raises(Exception):  # an alignment beyond the pieces
        ctx.batch_upload_alignments_parts(off, lens, parts[:-1], 6)
    ctx.close()


def _oracle_mle(ps, regs):
    out = []
    for c in regs:
        row = []
        for inst in (ps.model.coding_model, ps.model.noncoding_model):
            tr = {}
            x, (lp, el) = o.maximize_lpr(lambda r: o.lpr_leaves(inst, c, r), lambda r: r[0], init=1.0, trace=tr)
            row.append((x, lp, el, tr.get("iterations", 0), tr.get("random_tries", 0)))
        out.append(row)
    return out


def test_mle_examples(params_base):
    """maximize_lpr (find_init + GSL Brent) batched on the device vs the oracle: same iterate
    count, rho to 1e-9, scores to 1e-6 dB; and the reference's golden windows (src/test.ml:27-39)."""
    for pset, fn, window, anc_window in (("12flies", "tal-AA.fa", (297.62, 297.63), (48.25, 48.26)),
                                         ("29mammals", "ALDH2.exon5.fa", (-178.93, -178.92), (-38.29, -38.28))):
        ps = H.oracle_paramset(params_base, pset)
        regs, _ = H.example_codes(ps, fn, frames=3)
        ctx = H.make_context(ps)
        off, codes = H.regions_to_batch(regs)
        ctx.batch_upload(off, codes)
        res = [ctx.maximize_lpr(m) for m in (0, 1)]
        ora = _oracle_mle(ps, regs)
        for r in range(len(regs)):
            for m in (0, 1):
                rho, lpr, elpr, st, ne = (a[r] for a in res[m])
                ox, olp, oel, oit, otries = ora[r][m]
                assert (st & ~64) == 0
                assert abs(rho - ox) < 1e-9 * max(1.0, ox), (pset, r, m, rho, ox)
                assert abs(H.DB * (lpr - olp)) < 1e-6 and abs(H.DB * (elpr - oel)) < 1e-6
                assert ne == 3 + otries + 3 + 1 + oit + 1
        score0 = H.DB * (res[0][1][0] - res[1][1][0])
        anc0 = H.DB * (res[0][2][0] - res[1][2][0])
        assert window[0] < score0 < window[1] and anc_window[0] < anc0 < anc_window[1]
        ctx.close()


def test_mle_simulated_120mammals(params_base):
    """BASELINE.json configs[2] in miniature: 120mammals, mle, alignments simulated at different tree
    scales; batched device Brent vs the oracle's, per region and model."""
    ps = H.oracle_paramset(params_base, "120mammals")
    rng = np.random.default_rng(11)
    regs = []
    for i, rho in enumerate((0.4, 1.0, 2.2)):
        inst = ps.model.coding_model if i % 2 == 0 else ps.model.noncoding_model
        regs.append(o.simulate_columns(inst.model(rho), 24 + 3 * i, rng))
    ctx = H.make_context(ps)
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    res = [ctx.maximize_lpr(m) for m in (0, 1)]
    ora = _oracle_mle(ps, regs)
    for r in range(len(regs)):
        for m in (0, 1):
            rho, lpr, elpr, st, ne = (a[r] for a in res[m])
            ox, olp, oel, oit, otries = ora[r][m]
            assert (st & ~64) == 0
            assert abs(rho - ox) < 1e-8 * max(1.0, ox)
            assert abs(H.DB * (lpr - olp)) < 1e-6 and abs(H.DB * (elpr - oel)) < 1e-6
            assert ne == 3 + otries + 3 + 1 + oit + 1
    ctx.close()


def test_mle_multi_model_equals_per_model_calls(params_base):
    """pcsf_maximize_lpr_multi (coding and noncoding searches advanced in the same rounds, as
    llr_MaxLik needs both, src/PhyloCSFModel.ml:130-136) returns bit-identical results to one
    pcsf_maximize_lpr call per model, in either model order."""
    ps = H.oracle_paramset(params_base, "29mammals")
    rng = np.random.default_rng(23)
    regs = [o.simulate_columns(ps.model.coding_model.model(r), n, rng) for r, n in ((0.5, 31), (1.0, 9), (2.0, 64), (1.3, 1))]
    ctx = H.make_context(ps)
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    single = [ctx.maximize_lpr(m) for m in (0, 1)]
    for order in ((0, 1), (1, 0)):
        multi = ctx.maximize_lpr_multi(list(order))
        for k, m in enumerate(order):
            for a, b in zip(single[m], multi):
                assert (a == b[k]).all(), (order, m)
    ctx.close()


def test_mle_boundary_and_flat_regions(params_base):
    """Regions whose likelihood is monotone in rho: identical sequences (best rho -> lower bound) make
    find_init exhaust its 250 random tries and return the boundary (Fit.ml:42-47); the device driver
    must take the same path as the oracle (same OCaml Random stream, same number of evaluations)."""
    ps = H.oracle_paramset(params_base, "12flies")
    n = ps.tree.n_leaves
    same = np.tile(np.array([[14], [35], [7], [60], [22], [9]], dtype=np.uint8), (1, n))  # every species identical
    regs = [same]
    ctx = H.make_context(ps)
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    res = [ctx.maximize_lpr(m) for m in (0, 1)]
    ora = _oracle_mle(ps, regs)
    for m in (0, 1):
        rho, lpr, elpr, st, ne = (a[0] for a in res[m])
        ox, olp, oel, oit, otries = ora[0][m]
        assert otries == 250 and (st & 64) and (st & ~64) == 0
        assert rho == ox == 0.01
        assert abs(H.DB * (lpr - olp)) < 1e-6
        assert ne == 3 + 250 + 1
    ctx.close()


def test_pairs_api_matches_per_model_api(params_base):
    """pcsf_models_set / pcsf_pt_build_pairs / pcsf_lpr_pairs (the omega strategy's batched form) give
    bit-identical numbers to pcsf_model_set / pcsf_pt_build / pcsf_lpr."""
    ps = H.oracle_paramset(params_base, "29mammals")
    rng = np.random.default_rng(5)
    regs = [o.simulate_columns(ps.model.coding_model.model(1.0), n, rng) for n in (20, 33, 7)]
    ctx = H.make_context(ps)
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    scales = [0.5, 1.0, 1.7]
    ctx.pt_build(0, scales)
    ctx.pt_build(1, scales)
    a = ctx.lpr([0, 1, 0], [0, 1, 2], [0, 1, 2])
    qs = [ps.model.coding_model.q, ps.model.noncoding_model.q]
    ctx.models_set(4, np.stack([q.S for q in qs]), np.stack([q.Sinv for q in qs]), np.stack([q.lam for q in qs]),
                   np.stack([q.equilibrium() for q in qs]))
    st = ctx.pt_build_pairs([4, 5, 4], [0.5, 1.0, 1.7])
    assert (st == 0).all()
    b = ctx.lpr_pairs([0, 1, 2], [0, 1, 2])
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    ctx.close()


def test_device_omega_eigen_matches_host(params_base):
    """K5 (pcsf_omega_models_set: Q assembly + batched Jacobi on the device) against the host layer and
    the oracle: S diag(lambda) S^-1 reproduces the oracle's Q, the prior is the equilibrium, and P(t)
    built from the device model equals the oracle's P(t)."""
    from phylocsf_b200 import host

    ps = H.oracle_paramset(params_base, "12flies")
    ctx = H.make_context(ps)
    rng = np.random.default_rng(2)
    qs = np.array([[2.5, 1.0, 1.0] + [1.0] * 9,
                   [1.7, 0.2, 0.01, 1.3, 0.8, 1.1, 0.9, 1.2, 0.7, 1.05, 0.95, 1.4],
                   [9.9, 3.0, 0.5] + list(rng.uniform(0.3, 3.0, 9))])
    st = ctx.omega_models_set(3, qs)
    assert (st == 0).all()
    for i, v in enumerate(qs):
        Qo = o.omega_q(list(v))
        qd = o.QDiag(Qo)
        d = ctx.model_get(3 + i)
        np.testing.assert_allclose(d["S"] @ np.diag(d["lam"]) @ d["Sinv"], Qo, atol=5e-13)
        np.testing.assert_allclose(d["S"] @ d["Sinv"], np.eye(64), atol=1e-12)
        np.testing.assert_allclose(np.sort(d["lam"]), np.sort(qd.lam), atol=1e-12)
        np.testing.assert_allclose(d["prior"], qd.equilibrium(), atol=1e-14)
        hq, hpi = host.omega_q(v)
        assert (hq == Qo).all()
    ctx.pt_build_pairs([3, 4, 5], [1.0, 0.4, 2.0])
    regs, _ = H.example_codes(ps, "tal-AA.fa")
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    lpr, _, st = ctx.lpr_pairs([0, 1, 2], [0, 0, 0])
    for i, (v, rho) in enumerate(zip(qs, (1.0, 0.4, 2.0))):
        inst = o.OmegaInstance(ps.tree, list(v), rho)
        assert abs(H.DB * (lpr[i] - o.omega_lpr_leaves(inst, regs[0]))) < 1e-6
    ctx.close()


def _lpr_extended_precision(model, codes):
    """Pruning in numpy longdouble (x87 80-bit: exponent range down to 1e-4932), as the yardstick for the
    rescale option where plain FP64 underflows."""
    t = model.tree
    P = model.pms.astype(np.longdouble)
    prior = model.prior().astype(np.longdouble)
    n, ncols = t.n_leaves, codes.shape[0]
    alpha = {}
    for l in range(n):
        a = np.ones((ncols, 64), dtype=np.longdouble)
        known = codes[:, l] < 64
        a[known] = 0
        a[known, codes[known, l]] = 1
        alpha[l] = a
    for i in range(n, t.size):
        lc, rc = t.children[i]
        alpha[i] = (alpha[lc] @ P[lc].T) * (alpha[rc] @ P[rc].T)
    z = alpha[t.root] @ prior
    return float(np.log(z).sum())


@pytest.mark.parametrize("form", [1, 2, 3, 4])
def test_rescale_option_rescues_underflow(params_base, form):
    """(in both forms of the pruning kernel) PCSF_OPT_RESCALE: uniform-random columns on the 120-leaf tree underflow to -inf without it (the
    reference's behaviour); with it the score is finite and equals an extended-precision evaluation.
    Columns that never get near the threshold are unchanged to the last bit."""
    if np.finfo(np.longdouble).minexp > -16000:
        pytest.skip("no extended-precision long double on this platform")
    ps = H.oracle_paramset(params_base, "120mammals")
    rng = np.random.default_rng(3)
    bad = rng.integers(0, 64, size=(9, 120)).astype(np.uint8)
    good = o.simulate_columns(ps.model.coding_model.model(1.0), 21, rng)
    regs = [bad, good]
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    ctx.option_set(2, min(form, 2))
    ctx.option_set(3, {3: 3, 4: 2}.get(form, 1))  # forms 3, 4: the wide form with cherry tables / + 3-leaf tables
    lpr0, elpr0, st0 = ctx.lpr_all([0, 1])
    assert np.isneginf(lpr0[:, 0]).all() and np.isfinite(lpr0[:, 1]).all()
    ctx.option_set(1, 1)
    lpr1, elpr1, st1 = ctx.lpr_all([0, 1])
    assert np.isfinite(lpr1).all() and (st1 == 0).all()
    assert (lpr1[:, 1] == lpr0[:, 1]).all() and (elpr1[:, 1] == elpr0[:, 1]).all()
    for m, inst in enumerate((ps.model.coding_model, ps.model.noncoding_model)):
        want = _lpr_extended_precision(inst.model(1.0), bad)
        assert abs(lpr1[m, 0] - want) < 1e-9 * abs(want), (lpr1[m, 0], want)
    ctx.close()


def test_pipelined_score_alignments_equals_two_step_path(params_base, monkeypatch):
    """pcsf_score_alignments (chunked, H2D on a second stream) gives bit-identical results to
    pcsf_batch_upload_alignments + pcsf_lpr_all, with the chunk size forced small enough for many chunks."""
    monkeypatch.setenv("PCSF_CHUNK_COLS", "300")
    ps = H.oracle_paramset(params_base, "29mammals")
    rng = np.random.default_rng(9)
    n = ps.tree.n_leaves
    lens = [30, 111, 3, 2, 64, 300, 17, 90, 45, 201, 6, 150]
    rows, off = [], [0]
    for L in lens:
        a = rng.choice(np.frombuffer(b"ACGTacgtN-", dtype=np.uint8), size=(n, L), p=[.2, .2, .2, .2, .03, .03, .03, .03, .04, .04])
        rows.append(a)
        off.append(off[-1] + n * L)
    nt = np.concatenate([a.ravel() for a in rows])
    for frames in (1, 3, 6):
        ctx = H.make_context(ps)
        ctx.pt_build(0, [1.0])
        ctx.pt_build(1, [1.0])
        ctx.batch_upload_alignments(off[:-1], lens, nt, frames)
        a = ctx.lpr_all([0, 1])
        b = ctx.score_alignments(off[:-1], lens, nt, frames, [0, 1])
        assert a[0].shape == b[0].shape == (2, len(lens) * frames)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
        ctx.close()


@pytest.mark.parametrize("species", [["dmel", "dvir"], ["dmel", "dana", "dvir"], ["dmel", "dsim", "dsec", "dyak"]])
def test_tiny_trees(params_base, species):
    """Two-, three- and four-leaf trees (after --species pruning, Newick.subtree merges the spliced
    branches): a walk that is a single cherry, a cherry plus one edge, and two cherries joined at the root."""
    ps = H.oracle_paramset(params_base, "12flies", species=species)
    assert ps.tree.n_leaves == len(species)
    rng = np.random.default_rng(1)
    regs = [rng.integers(0, 65, size=(n, len(species))).astype(np.uint8) for n in (1, 19, 130)]
    ctx = H.make_context(ps)
    ctx.pt_build(0, [1.0, 0.2])
    ctx.pt_build(1, [1.0, 0.2])
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    for si, rho in enumerate((1.0, 0.2)):
        lpr, elpr, st = ctx.lpr_all([0, 1], [si, si])
        lo, eo = H.oracle_fixed(ps, regs, rho=rho)
        assert np.abs(H.DB * (lpr - lo)).max() < TOL_DB and np.abs(H.DB * (elpr - eo)).max() < TOL_DB
    ctx.close()


def _random_tree(rng, n_leaves, shape):
    """children[2*(i-n)+k] in T numbering (leaves left to right, internal nodes in post-order) and branch lengths."""
    def build(leaves):
        if len(leaves) == 1:
            return leaves[0]
        if shape == "caterpillar":
            cut = len(leaves) - 1
        elif shape == "balanced":
            cut = len(leaves) // 2
        else:
            cut = int(rng.integers(1, len(leaves)))
        return (build(leaves[:cut]), build(leaves[cut:]))
    nested = build(list(range(n_leaves)))
    children = []
    def number(t):
        if not isinstance(t, tuple):
            return t
        l, r = number(t[0]), number(t[1])
        children.append((l, r))
        return n_leaves + len(children) - 1
    root = number(nested)
    assert root == 2 * n_leaves - 2
    bl = rng.uniform(0.01, 0.6, size=2 * n_leaves - 2)
    return np.array(children, dtype=np.int32).reshape(-1), bl


@pytest.mark.parametrize("seed,n_leaves,shape", [(1, 5, "caterpillar"), (2, 6, "balanced"), (3, 8, "random"), (4, 9, "random"),
                                                 (5, 4, "balanced"), (6, 3, "caterpillar"), (7, 11, "random"), (8, 7, "caterpillar")])
def test_table_programs_on_random_trees(params_base, seed, n_leaves, shape):
    """The table programs are derived from the plain tree program by pattern (cherry / + leaf / + leaf, then the edge
    above). Random tree shapes exercise the corners: cherries right under the root, caterpillars that end at the root,
    balanced trees with no leaf next to a cherry. Every table level must reproduce the narrow form bit for bit."""
    import phylocsf_b200 as pb

    ps = H.oracle_paramset(params_base, "12flies")  # rate matrices only; the tree is replaced
    rng = np.random.default_rng(seed)
    children, bl = _random_tree(rng, n_leaves, shape)
    ctx = pb.Context(0)
    ctx.tree_set(n_leaves, children, bl)
    H.push_qdiag(ctx, 0, ps.model.coding_model.q)
    H.push_qdiag(ctx, 1, ps.model.noncoding_model.q)
    ctx.pt_build(0, [1.0, 0.3])
    ctx.pt_build(1, [1.0, 0.3])
    codes = rng.integers(0, 65, size=(450, n_leaves)).astype(np.uint8)
    codes[:200] = codes[:200, :1]                 # conserved columns
    codes[200:300, ::2] = codes[200:300, :1]      # half conserved
    ctx.batch_upload(np.array([0, 193, 200, 450], dtype=np.int64), codes)
    ref = None
    for form, mode in ((pb.Context.FORM_NARROW, 1), (pb.Context.FORM_WIDE, 1), (pb.Context.FORM_WIDE, 3), (pb.Context.FORM_WIDE, 2),
                       (pb.Context.FORM_WIDE, 4)):
        ctx.option_set(pb.Context.OPT_PRUNE_FORM, form)
        ctx.option_set(pb.Context.OPT_CHERRY_TABLES, mode)
        res = [ctx.lpr_all([0, 1], scale_idx=[si, si]) for si in (0, 1)]
        cols = [ctx.column_terms(m) for m in (0, 1)]
        key = [r[0] for r in res] + [r[1] for r in res] + [c[0] for c in cols] + [c[1] for c in cols]
        if ref is None:
            ref = key
            assert np.isfinite(res[0][0]).any()
        else:
            assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ref, key)), (form, mode)
    ctx.close()
