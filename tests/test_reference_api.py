"""The reference's own end-to-end checks (src/test.ml:27-49), written against the Python mirror of its
model interface (phylocsf_b200/model.py: make / pleaves / lpr_leaves / maximize_lpr / score)."""
import os

import pytest

from tools import golden_params as gp

pytestmark = pytest.mark.gpu


def _aln(fn):
    lines = gp.example_lines(fn)
    species = [h[1:].split("|")[0].strip() for h in lines[0::2]]
    return species, lines[1::2]


def test_talAA(params_base):
    from phylocsf_b200.model import Model

    m = Model.make(os.path.join(params_base, "PhyloCSF_Parameters", "12flies"))
    species, rows = _aln("tal-AA.fa")
    leaves = [m.pleaves(species, rows)]
    ans = m.score("MaxLik", leaves)[0]
    assert 297.62 < ans["score"] < 297.63 and 48.25 < ans["anc_comp_score"] < 48.26
    fixed = m.score("FixedLik", leaves)[0]
    assert abs(fixed["score"] - 361.6876) < 1e-3
    # lpr_leaves at the fitted scale reproduces maximize_lpr's value (the reference's final `f x`)
    rho, rec = m.maximize_lpr(m.CODING, leaves)[0]
    again = m.lpr_leaves(m.CODING, leaves, rho)[0]
    assert again["lpr_leaves"] == rec["lpr_leaves"] and again["elpr_anc"] == rec["elpr_anc"]
    m.close()


def test_aldh2_ex5_in_and_out_of_frame(params_base):
    from phylocsf_b200.model import Model

    m = Model.make(os.path.join(params_base, "PhyloCSF_Parameters", "29mammals"))
    species, rows = _aln("ALDH2.exon5.fa")
    hi = len(rows[0]) - 1
    leaves = [m.pleaves(species, rows, lo=f, hi=hi) for f in (0, 1, 2)]
    ans = m.score("MaxLik", leaves)
    assert -178.93 < ans[0]["score"] < -178.92 and -38.29 < ans[0]["anc_comp_score"] < -38.28
    assert 218.26 < ans[1]["score"] < 218.27
    assert max(range(3), key=lambda f: ans[f]["score"]) == 1
    m.close()


def test_fresh_lists_are_never_served_from_a_stale_batch(params_base):
    """Regression (round-1 advisor finding): back-to-back calls with temporary lists, which CPython may give the
    same id(), and in-place mutation of a list must each score what was passed."""
    from phylocsf_b200.model import Model

    m = Model.make(os.path.join(params_base, "PhyloCSF_Parameters", "12flies"))
    species, rows = _aln("tal-AA.fa")
    a = m.pleaves(species, rows)
    b = a[:20].copy()
    sa = m.score("FixedLik", [a])[0]["score"]
    sb = m.score("FixedLik", [b])[0]["score"]
    assert abs(sa - 361.6876) < 1e-3 and abs(sb - sa) > 1.0
    lst = [a]
    s1 = m.lpr_leaves(m.CODING, lst, 1.0)[0]["lpr_leaves"]
    lst[0] = b
    s2 = m.lpr_leaves(m.CODING, lst, 1.0)[0]["lpr_leaves"]
    assert s1 != s2 and s2 == m.lpr_leaves(m.CODING, [b], 1.0)[0]["lpr_leaves"]
    m.close()
