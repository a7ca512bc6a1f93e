"""Pins the CPU oracle (oracle/) against every known answer the reference's own tests hold for the
scoring path:
  lib/CamlPaml/test.ml:8-54   2-state tree (A,(B,C)), 8 leaf patterns, eps 1e-3
  lib/CamlPaml/test.ml:56-99  JC69 BEAGLE tiny test, lnL = -1574.63623 +- 1e-3
  src/test.ml:27-59           four end-to-end runs on PhyloCSF_Examples (score windows + coordinates)
"""
import math
import os

import numpy as np
import pytest

from oracle import oracle as o
from tools import golden_params as gp


# ---------------------------------------------------------------- lib/CamlPaml/test.ml:8-54
@pytest.mark.parametrize("lvs", [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)])
def test_two_state_pruning(lvs):
    t = o.Tree.of_newick(o.newick_parse("(A,(B,C))"))
    assert t.children[3] == (1, 2) and t.children[4] == (0, 3)
    sm0 = np.array([[0.8, 0.2], [0.25, 0.75]])
    sm1 = np.array([[0.9, 0.1], [0.85, 0.15]])
    prior = np.array([0.6, 0.4])

    class M:  # a PhyloModel with hand-set pms, as PhyloLik.prepare t sms prior (test.ml:31)
        tree = t
        pms = np.ascontiguousarray(np.stack([sm0, sm1, sm1, sm1]))

        @staticmethod
        def prior():
            return prior

    a = [np.array([1.0 if lvs[i] == s else 0.0 for s in (0, 1)]) for i in range(3)]
    a3 = (sm1 @ a[1]) * (sm1 @ a[2])
    a4 = (sm0 @ a[0]) * (sm1 @ a3)
    z = float(prior @ a4)
    z_o, alpha = o.likelihood_column(M, list(lvs))
    assert abs(z_o - z) < 1e-3 and abs(z_o - z) < 1e-15
    post_root = alpha[1] * prior / z_o  # PhyloLik.ml:137-138 at the root
    np.testing.assert_allclose(post_root, prior * a4 / z, atol=1e-3)
    np.testing.assert_allclose(alpha[0], a3, atol=1e-15)
    # outside pass: posterior(root) and posterior(internal) exactly as the reference's test writes them (test.ml:35-43)
    b4 = prior
    b3 = np.array([(b4[0] * sm0[0, 0] * a[0][0] + b4[0] * sm0[0, 1] * a[0][1]) * sm1[0, 0] + (b4[1] * sm0[1, 0] * a[0][0] + b4[1] * sm0[1, 1] * a[0][1]) * sm1[1, 0],
                   (b4[0] * sm0[0, 0] * a[0][0] + b4[0] * sm0[0, 1] * a[0][1]) * sm1[0, 1] + (b4[1] * sm0[1, 0] * a[0][0] + b4[1] * sm0[1, 1] * a[0][1]) * sm1[1, 1]])
    ec = np.zeros((4, 2, 2))
    z2, post = o.posteriors_column(3, t.children_array(), M.pms, prior, list(lvs), 1.0, ec)
    assert z2 == z_o
    np.testing.assert_allclose(post[4], b4 * a4 / z, atol=1e-3)
    np.testing.assert_allclose(post[3], b3 * a3 / z, atol=1e-3)
    np.testing.assert_allclose(post[4], b4 * a4 / z, atol=1e-15)
    np.testing.assert_allclose(post[3], b3 * a3 / z, atol=1e-15)
    for i in range(3):
        np.testing.assert_array_equal(post[i], a[i])  # a leaf's posterior is its leaf vector (PhyloLik.ml:133-134)
    # branch posteriors (PhyloLik.ml:140-180): a joint distribution over (parent state, child state) of every branch
    # whose marginals are the node posteriors of its two ends
    for br in range(4):
        par = 3 if br in (1, 2) else 4
        assert abs(ec[br].sum() - 1.0) < 1e-12
        np.testing.assert_allclose(ec[br].sum(axis=1), post[par], atol=1e-12)
        np.testing.assert_allclose(ec[br].sum(axis=0), post[br], atol=1e-12)


# ---------------------------------------------------------------- lib/CamlPaml/test.ml:56-99
_BEAGLE = None


def _beagle():
    global _BEAGLE
    if _BEAGLE is None:
        import json

        with open(os.path.join(os.path.dirname(__file__), "golden", "beagle_tiny.json")) as f:
            _BEAGLE = json.load(f)
    return _BEAGLE


def test_beagle_tiny_jc69():
    d = _beagle()
    t = o.Tree.of_newick(o.newick_parse(d["newick"]))
    q = o.QDiag(np.array([[-3.0, 1, 1, 1], [1, -3.0, 1, 1], [1, 1, -3.0, 1], [1, 1, 1, -3.0]])).scaled(1.0 / 3.0)
    np.testing.assert_allclose(q.equilibrium(), [0.25] * 4, atol=1e-12)
    m = o.PhyloModel(t, q, t.branches)
    idx = {"A": 0, "C": 1, "G": 2, "T": 3}
    seqs = [d["human"], d["chimp"], d["gorilla"]]
    codes = np.array([[idx.get(s[i], 4) for s in seqs] for i in range(len(seqs[0]))], dtype=np.uint8)
    ll = o.lpr_columns(m, codes)[0]
    assert abs(ll - (-1574.63623)) < 1e-3


# ---------------------------------------------------------------- src/test.ml:27-59
def _run(params_base, pset, fn, **kw):
    opts = o.Options(**kw)
    ps = o.load_paramset(os.path.join(params_base, "PhyloCSF_Parameters", pset), opts)
    lines = o.process_alignment(ps, opts, fn, gp.example_lines(fn))
    return lines[-1].split("\t")


def test_tal_AA(params_base):
    ans = _run(params_base, "12flies", "tal-AA.fa", anc_comp=True)
    assert ans[1] == "score(decibans)"
    assert 297.62 < float(ans[2]) < 297.63
    assert 48.25 < float(ans[3]) < 48.26


def test_aldh2_ex5_out(params_base):
    ans = _run(params_base, "29mammals", "ALDH2.exon5.fa", anc_comp=True)
    assert ans[1] == "score(decibans)"
    assert -178.93 < float(ans[2]) < -178.92
    assert -38.29 < float(ans[3]) < -38.28


def test_aldh2_ex5_in(params_base):
    ans = _run(params_base, "29mammals", "ALDH2.exon5.fa", frames=6)
    assert ans[1] == "max_score(decibans)"
    assert 218.26 < float(ans[2]) < 218.27
    assert int(ans[3]) == 1 and int(ans[4]) == 111 and ans[5] == "+"


@pytest.mark.slow
def test_aldh2_mRNA(params_base):
    ans = _run(params_base, "29mammals", "Aldh2.mRNA.fa", orf="ATGStop", frames=3, remove_ref_gaps=True, aa=True)
    assert ans[1] == "max_score(decibans)"
    assert 2013.92 < float(ans[2]) < 2013.93
    assert int(ans[3]) == 343 and int(ans[4]) == 1899
    assert ans[5].startswith("MLRAALTTVRRGPRLSRLLSAAA")


def test_tal_AA_fixed_restatement_value(params_base):
    """BASELINE.json configs[0]; no reference-published value, 361.6876 is restatement-derived."""
    ans = _run(params_base, "12flies", "tal-AA.fa", anc_comp=True, strategy="fixed")
    assert abs(float(ans[2]) - 361.6876) < 1e-4 and abs(float(ans[3]) - 48.0241) < 1e-4


# ---------------------------------------------------------------- P(t) semantics, Q.ml:211-249
def test_pt_fixups_and_prior(params_base):
    s, pi = o.read_ecm(os.path.join(params_base, "PhyloCSF_Parameters", "23flies_coding.ECM"))
    assert abs(pi.sum() - 1.0) > 5e-7  # the file prior is not normalised
    q = o.QDiag(o.ecm_q(s, pi))
    eq = q.equilibrium()
    assert abs(eq.sum() - 1.0) < 1e-12
    np.testing.assert_allclose(eq, pi / pi.sum(), rtol=1e-8)
    for t in (0.0, 1e-3, 0.1, 1.0, 10.0):
        P = q.to_Pt(t)
        assert (P >= 0).all()
        np.testing.assert_allclose(P.sum(axis=1), 1.0, atol=1e-12)
        if t == 0.0:
            np.testing.assert_allclose(P, np.eye(64), atol=1e-9)
    with pytest.raises(o.OracleFailure):
        q.to_Pt(-1.0)


def test_tree_numbering_58mammals(params_base):
    t = o.Tree.of_newick(o.newick_parse(open(os.path.join(params_base, "PhyloCSF_Parameters", "58mammals.nh")).read()))
    assert t.n_leaves == 58 and t.size == 115 and t.root == 114
    assert t.labels[0] == "Human" and t.labels[1] == "Chimp"
    assert t.children[58] == (0, 1)  # first cherry is the first internal node (post-order)
    for i in range(58, 115):
        lc, rc = t.children[i]
        assert lc < i and rc < i


def test_ocaml_random_is_deterministic():
    a = [o.OCamlRandom(0).rawfloat() for _ in range(3)]
    assert a[0] == a[1] == a[2] and 0.0 <= a[0] < 1.0


def test_find_orfs_modes():
    # hand-traced against src/PhyloCSF.ml:134-196
    assert o.find_orfs("CCATGAAACCCGGGTTTTAGCCATGCCCAAATGAGGG", 2, "ATGStop", 2) == [(2, 16)]
    # nested ATGs share the stop; the later start is reported first (starts is a LIFO list)
    assert o.find_orfs("ATGAAAATGCCCTAAGG", 0, "ATGStop", 2) == [(6, 11), (0, 11)]
    assert o.find_orfs("ATGAAAATGCCCTAAGG", 0, "ATGStop", 3) == [(0, 11)]
    # StopStop: first ORF runs from the frame start to the codon before the stop; the tail ORF
    # runs off the end of the alignment
    assert o.find_orfs("CCCAAATAGGGGTTTCC", 0, "StopStop", 1) == [(0, 5), (9, 14)]
    assert o.find_orfs("CCCAAATAGGGGTTTCC", 0, "ToFirstStop", 1) == [(0, 5)]
    assert o.find_orfs("CCCAAATAGGGGTTTCC", 0, "FromLastStop", 1) == [(9, 14)]


# ---------------------------------------------------------------- an analytic pin for the omega model
def _k80(kappa):
    """Unit-rate K80 (HKY85 with uniform frequencies): transitions A<->G, C<->T at kappa, transversions at 1."""
    m = np.ones((4, 4))
    for a, b in ((0, 2), (2, 0), (1, 3), (3, 1)):
        m[a, b] = kappa
    m /= 4.0
    np.fill_diagonal(m, 0.0)
    np.fill_diagonal(m, -m.sum(axis=1))
    return m * 4.0 / (kappa + 2.0)


def gapfree_codon_columns(seqs):
    """(codon codes [ncols, n], nucleotide codes [3 ncols, n]) of the codon columns in which no sequence has a gap."""
    idx = {"A": 0, "C": 1, "G": 2, "T": 3}
    L = len(seqs[0]) // 3 * 3
    cod, nuc = [], []
    for p in range(0, L, 3):
        trip = [[idx.get(s[p + k], -1) for k in range(3)] for s in seqs]
        if all(min(t) >= 0 for t in trip):
            cod.append([16 * t[0] + 4 * t[1] + t[2] for t in trip])
            for k in range(3):
                nuc.append([t[k] for t in trip])
    return np.array(cod, dtype=np.uint8), np.array(nuc, dtype=np.uint8)


@pytest.mark.parametrize("kappa", [1.0, 2.7])
def test_omega_model_at_omega_1_is_three_independent_nucleotide_processes(kappa):
    """The reference holds no golden for the omega strategy, but one corner of it is pinned analytically: with omega =
    sigma = 1 and uniform F3x4 the codon rate matrix (src/OmegaModel.ml:21-80, scaled to unit rate, PhyloModel.ml:94-106)
    is the Kronecker sum of three K80 nucleotide processes at a third of the rate each, so a codon column's likelihood is
    the product of its three nucleotide columns' likelihoods. On the BEAGLE test's tree and sequences (lib/CamlPaml/test.ml:
    56-99) with kappa = 1 and the tree scaled by 3 the nucleotide side is exactly the configuration whose lnL the
    reference pins (-1574.63623 over all columns)."""
    qs = [kappa, 1.0, 1.0] + [1.0] * 9
    Qc = o.omega_q(qs)
    Qn = _k80(kappa)
    I = np.eye(4)
    ksum = np.kron(np.kron(Qn, I), I) + np.kron(np.kron(I, Qn), I) + np.kron(np.kron(I, I), Qn)
    np.testing.assert_allclose(Qc, ksum / 3.0, atol=1e-15)
    d = _beagle()
    t = o.Tree.of_newick(o.newick_parse(d["newick"]))
    cod, nuc = gapfree_codon_columns([d["human"], d["chimp"], d["gorilla"]])
    assert cod.shape[0] > 200
    inst = o.OmegaInstance(t, qs, 3.0)
    ll_codon = o.omega_lpr_leaves(inst, cod)
    mn = o.PhyloModel(t, o.QDiag(Qn), t.branches)
    ll_nt = o.lpr_columns(mn, nuc)[0]
    assert abs(ll_codon - ll_nt) < 1e-9 * abs(ll_nt), (ll_codon, ll_nt)
