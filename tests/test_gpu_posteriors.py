"""K6 (pcsf_posteriors: outside algorithm, node posteriors, expected substitution counts - PhyloLik.ml:96-180) against
the oracle's restatement, which tests/test_oracle_golden.py pins to the reference's own 2-state known answers
(lib/CamlPaml/test.ml:8-54 - the only vectors the reference holds for the outside pass)."""
import numpy as np
import pytest

import pcsf_helpers as H
from oracle import oracle as o

pytestmark = pytest.mark.gpu


def _check(ps, regs, rho=1.0, models=(0, 1)):
    ctx = H.make_context(ps)
    off, codes = H.regions_to_batch(regs)
    ctx.batch_upload(off, codes)
    t = ps.tree
    nodes = list(range(t.size))
    for m in models:
        inst = ps.model.coding_model if m == 0 else ps.model.noncoding_model
        ctx.pt_build(m, [0.7, rho])
        post, ec, z = ctx.posteriors(m, 1, nodes=nodes)
        zo, po, eo = o.posteriors_columns(inst.model(rho), codes)
        np.testing.assert_allclose(z, zo, rtol=1e-10, atol=0)
        np.testing.assert_allclose(post, po.transpose(1, 0, 2), rtol=0, atol=2e-12)
        np.testing.assert_allclose(ec, eo, rtol=0, atol=1e-11 * max(1, codes.shape[0]))
        live = int((zo > 0).sum())
        # every branch's expected counts form a joint distribution per column: they sum to the number of columns
        np.testing.assert_allclose(ec.sum(axis=(1, 2)), live, rtol=1e-10)
        # ... whose marginals are the node posteriors of the branch's two ends, summed over the columns
        par = np.zeros(t.size, dtype=int)
        for i in range(t.n_leaves, t.size):
            par[t.children[i][0]] = par[t.children[i][1]] = i
        for br in sorted({0, min(t.n_leaves, t.size - 2), t.size - 2}):
            np.testing.assert_allclose(ec[br].sum(axis=1), post[par[br]].sum(axis=0), atol=1e-9)
            if br >= t.n_leaves:  # (a marginalised leaf's node_posterior is its all-ones leaf vector, not a distribution)
                np.testing.assert_allclose(ec[br].sum(axis=0), post[br].sum(axis=0), atol=1e-9)
        # the subset interface returns the same rows
        sub, _, _ = ctx.posteriors(m, 1, nodes=[t.size - 1, 1, t.n_leaves], ecounts=False, z=False)
        assert (sub[0] == post[t.size - 1]).all() and (sub[1] == post[1]).all() and (sub[2] == post[t.n_leaves]).all()
        # two evaluations give the same bits (per-CTA accumulators reduced in a fixed order)
        post2, ec2, z2 = ctx.posteriors(m, 1, nodes=nodes)
        assert (post2 == post).all() and (ec2 == ec).all() and (z2 == z).all()
    ctx.close()


def test_posteriors_tal_AA(params_base):
    ps = H.oracle_paramset(params_base, "12flies")
    regs, _ = H.example_codes(ps, "tal-AA.fa", frames=3)  # 5 of 12 species present: most leaves marginalised
    _check(ps, regs)


def test_posteriors_simulated_58mammals_with_gaps(params_base):
    ps = H.oracle_paramset(params_base, "58mammals")
    rng = np.random.default_rng(12)
    regs = [o.simulate_columns(ps.model.coding_model.model(1.0), 45, rng), o.simulate_columns(ps.model.noncoding_model.model(1.0), 20, rng)]
    regs[0][:, 7] = 64
    regs[0][3, :] = 64
    regs[1][rng.random(regs[1].shape) < 0.1] = 64
    _check(ps, regs, rho=1.3)


def test_posteriors_impossible_columns_are_zero(params_base):
    """z = 0 (uniform-random columns underflow on the 120-leaf tree): node posteriors are all zeros and the column adds
    nothing to the expected counts (PhyloLik.ml:131-132,152)."""
    ps = H.oracle_paramset(params_base, "120mammals")
    rng = np.random.default_rng(3)
    bad = rng.integers(0, 64, size=(3, 120)).astype(np.uint8)
    good = o.simulate_columns(ps.model.coding_model.model(1.0), 4, rng)
    ctx = H.make_context(ps)
    off, codes = H.regions_to_batch([bad, good])
    ctx.batch_upload(off, codes)
    ctx.pt_build(0, [1.0])
    post, ec, z = ctx.posteriors(0, 0, nodes=[238, 120, 5])
    assert (z[:3] == 0).all() and (z[3:] > 0).all()
    assert (post[:, :3] == 0).all()
    np.testing.assert_allclose(ec.sum(axis=(1, 2)), 4.0, rtol=1e-10)
    zo, po, eo = o.posteriors_columns(ps.model.coding_model.model(1.0), codes)
    np.testing.assert_allclose(ec, eo, atol=1e-10)
    with pytest.raises(Exception):
        ctx.posteriors(0, 0, nodes=[239])
    ctx.close()


@pytest.mark.parametrize("species", [["dmel", "dvir"], ["dmel", "dana", "dvir"], ["dmel", "dsim", "dsec", "dyak", "dere"]])
def test_posteriors_tiny_trees_and_ragged_tiles(params_base, species):
    """Two-, three- and five-leaf trees (a root cherry has no internal non-root node; three leaves have one) and column
    counts around the kernel's 32-column tiles (1, 31, 32, 33, 0 columns)."""
    ps = H.oracle_paramset(params_base, "12flies", species=species)
    rng = np.random.default_rng(4)
    n = len(species)
    regs = [rng.integers(0, 65, size=(c, n)).astype(np.uint8) for c in (1, 31, 0, 32, 33)]
    _check(ps, regs, rho=0.8, models=(0,))


def test_posteriors_empty_batch(params_base):
    ps = H.oracle_paramset(params_base, "12flies")
    ctx = H.make_context(ps)
    ctx.batch_upload(np.array([0, 0], dtype=np.int64), np.zeros((0, 12), dtype=np.uint8))
    ctx.pt_build(0, [1.0])
    post, ec, z = ctx.posteriors(0, 0, nodes=[22, 3])
    assert post.shape == (2, 0, 64) and z.size == 0 and (ec == 0).all()
    ctx.close()


def test_posteriors_both_forms_of_the_kernel_agree(params_base, monkeypatch):
    """pcsf_posteriors runs K6 on the DMMA pipe; PCSF_K6_PLAIN=1 (read when the context is created) selects the plain-FP64
    form. Same walk, different summation order inside the 64-term products: z, posteriors and expected counts agree to
    rounding on 29mammals columns spanning several tiles and CTAs, gaps and an impossible column included."""
    ps = H.oracle_paramset(params_base, "29mammals")
    rng = np.random.default_rng(21)
    regs = [o.simulate_columns(ps.model.coding_model.model(1.0), 700, rng), o.simulate_columns(ps.model.noncoding_model.model(1.0), 333, rng)]
    regs[0][rng.random(regs[0].shape) < 0.05] = 64
    regs[1][5, :] = rng.integers(0, 64, size=29)
    off, codes = H.regions_to_batch(regs)
    nodes = list(range(ps.tree.size))
    res = []
    for plain in ("0", "1"):
        monkeypatch.setenv("PCSF_K6_PLAIN", plain)
        ctx = H.make_context(ps)
        ctx.batch_upload(off, codes)
        ctx.pt_build(0, [1.1])
        res.append(ctx.posteriors(0, 0, nodes=nodes))
        ctx.close()
    (p0, e0, z0), (p1, e1, z1) = res
    np.testing.assert_allclose(z0, z1, rtol=1e-13, atol=0)
    assert ((z0 == 0) == (z1 == 0)).all()
    np.testing.assert_allclose(p0, p1, rtol=0, atol=1e-13)
    np.testing.assert_allclose(e0, e1, rtol=0, atol=1e-10)
    np.testing.assert_allclose(e0.sum(axis=(1, 2)), float((z0 > 0).sum()), rtol=1e-11)
