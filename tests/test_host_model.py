"""CPU tests of the product's C++ host layer (no GPU needed): it must agree with the oracle's
independent restatement — tree numbering exactly, rate matrices to rounding, and the Jacobi
diagonalisation must give the same P(t), prior and scores as the oracle's LAPACK one."""
import ctypes
import os

import numpy as np
import pytest

import pcsf_helpers as H
from oracle import oracle as o
from phylocsf_b200 import _native as N
from phylocsf_b200 import host
from tools import golden_params as gp


def test_library_exports_every_declared_symbol():
    L = N.load()
    for s in N.SYMBOLS + host.HOST_SYMBOLS:
        assert hasattr(L, s), s
    # and the headers declare exactly these
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for hdr, syms in (("phylocsf_b200.h", N.SYMBOLS), ("phylocsf_host.h", host.HOST_SYMBOLS)):
        text = open(os.path.join(root, "include", hdr)).read()
        declared = set(re.findall(r"\b(pcsf_[a-z0-9_]+)\s*\(", text))
        assert declared == set(syms), (hdr, declared ^ set(syms))


def test_no_gpu_means_loud_failure_not_fallback():
    L = N.load()
    if L.pcsf_device_count() > 0:
        pytest.skip("a GPU is present")
    import phylocsf_b200 as pb
    with pytest.raises(pb.PcsfError):
        pb.Context(0)


@pytest.mark.parametrize("pset", gp.set_names())
def test_tree_and_q_match_oracle(params_base, pset):
    prefix = os.path.join(params_base, "PhyloCSF_Parameters", pset)
    ps = host.ParamSet(prefix)
    ops = H.oracle_paramset(params_base, pset)
    t = ops.tree
    assert ps.n_leaves == t.n_leaves and ps.leaf_labels == t.labels[: t.n_leaves]
    assert (ps.children == t.children_array()).all()
    assert (ps.branch_len == np.array(t.branches[: t.root])).all()
    for w, inst in enumerate((ops.model.coding_model, ops.model.noncoding_model)):
        d = ps.qdiag(w)
        assert (d["Q"] == inst.q.q).all()  # same evaluation order => identical bits
        np.testing.assert_allclose(d["S"] @ np.diag(d["lam"]) @ d["Sinv"], d["Q"], atol=2e-13)
        np.testing.assert_allclose(d["S"] @ d["Sinv"], np.eye(64), atol=1e-12)
        np.testing.assert_allclose(np.sort(d["lam"]), np.sort(inst.q.lam), atol=1e-12)
        np.testing.assert_allclose(d["prior"], inst.q.equilibrium(), rtol=0, atol=1e-15)
        assert abs(d["prior"].sum() - 1.0) < 1e-14


def test_species_pruning_matches_oracle(params_base):
    sp = ["Human", "Mouse", "Dog", "Cow", "Elephant", "Opossum"]
    prefix = os.path.join(params_base, "PhyloCSF_Parameters", "58mammals")
    ps = host.ParamSet(prefix, species=sp)
    ops = H.oracle_paramset(params_base, "58mammals", species=sp)
    assert ps.leaf_labels == ops.tree.labels[: ops.tree.n_leaves]
    assert (ps.children == ops.tree.children_array()).all()
    np.testing.assert_array_equal(ps.branch_len, np.array(ops.tree.branches[: ops.tree.root]))
    with pytest.raises(host.HostError):
        host.ParamSet(prefix, species=["Human"])


def test_jacobi_basis_gives_same_scores_as_lapack_basis(params_base):
    """Scores through the oracle's own P(t)/pruning code with the product's eigenbasis vs the
    oracle's: the eigenbasis choice must stay far inside the 1e-6 dB bar."""
    ops = H.oracle_paramset(params_base, "12flies")
    ps = host.ParamSet(os.path.join(params_base, "PhyloCSF_Parameters", "12flies"))
    regs, _ = H.example_codes(ops, "tal-AA.fa")
    for w, inst in enumerate((ops.model.coding_model, ops.model.noncoding_model)):
        d = ps.qdiag(w)
        q2 = o.QDiag.__new__(o.QDiag)
        q2.q, q2.tol, q2.S, q2.Sinv, q2.lam = d["Q"], 1e-6, d["S"], d["Sinv"], d["lam"]
        q2._pi, q2._memo = d["prior"], {}
        for rho in (1.0, 0.3):
            m1 = inst.model(rho)
            m2 = o.PhyloModel(ops.tree, q2, [rho * b for b in ops.tree.branches])
            a = o.lpr_columns(m1, regs[0])
            b = o.lpr_columns(m2, regs[0])
            assert abs(H.DB * (a[0] - b[0])) < 1e-8 and abs(H.DB * (a[1] - b[1])) < 1e-8
            assert np.abs(m1.pms - m2.pms).max() < 1e-13


def test_omega_q_matches_oracle():
    v = [2.5, 0.2, 0.01, 1.3, 0.8, 1.1, 0.9, 1.2, 0.7, 1.05, 0.95, 1.4]
    Q, pi = host.omega_q(v)
    Qo = o.omega_q(v)
    assert (Q == Qo).all()
    d = host.qdiag_reversible(Q, pi)
    qo = o.QDiag(Qo)
    np.testing.assert_allclose(d["prior"], qo.equilibrium(), atol=1e-14)
    np.testing.assert_allclose(np.sort(d["lam"]), np.sort(qo.lam), atol=1e-12)


def test_bad_inputs_raise(tmp_path, params_base):
    with pytest.raises(host.HostError):
        host.ParamSet(str(tmp_path / "nope"))
    # an ECM with a codon table out of order
    src = os.path.join(params_base, "PhyloCSF_Parameters", "12flies")
    for suf in (".nh", "_coding.ECM", "_noncoding.ECM"):
        txt = open(src + suf).read()
        if suf == "_coding.ECM":
            txt = txt.replace("AAA AAC", "AAC AAA")
        open(str(tmp_path / ("x" + suf)), "w").write(txt)
    with pytest.raises(host.HostError, match="codon order"):
        host.ParamSet(str(tmp_path / "x"))


def test_reemitted_parameter_files_hold_the_reference_tokens(params_base):
    """tools/golden_params.py re-emits the packed parameter sets in the reference's formats with normalised blanks. Where
    the reference tree is present (the build container), every token - each number as its original text, each Newick
    string - must equal the shipped file's; elsewhere (GPU box) the test has nothing to compare with and is skipped."""
    ref = "/root/reference/PhyloCSF_Parameters"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    mine = os.path.join(params_base, "PhyloCSF_Parameters")
    names = sorted(f for f in os.listdir(ref) if f.endswith((".nh", ".ECM")))
    assert len(names) == 42 and sorted(os.listdir(mine)) == names
    for f in names:
        a = open(os.path.join(ref, f)).read().split()
        b = open(os.path.join(mine, f)).read().split()
        assert a == b, f
