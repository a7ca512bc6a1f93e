"""Re-emit the packed reference input data (tests/golden/*.json, written by
tools/make_golden_params.py) as files in the reference's own formats, so that the file parsers
(.nh / .ECM / multi-FASTA) of both the product and the oracle are exercised on the GPU box, where
/root/reference does not exist.  Formats: SURVEY.md Appendix B."""
import json
import os
import tempfile

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
_cache = {}


def _load(name):
    if name not in _cache:
        with open(os.path.join(_GOLDEN, name)) as f:
            _cache[name] = json.load(f)
    return _cache[name]


def set_names():
    return sorted(_load("phylocsf_parameters.json")["sets"])


def ecm_text(e):
    lines = [" ".join(r) for r in e["s_lower"]]
    lines.append("")
    lines.append(" ".join(e["pi"]))
    lines += ["", ""]
    c = e["codons"]
    lines += [" ".join(c[i:i + 20]) for i in range(0, 64, 20)]
    return "\n".join(lines) + "\n"


def materialize(dirpath=None, sets=None):
    """Write <dirpath>/PhyloCSF_Parameters/<set>.nh, _coding.ECM, _noncoding.ECM; return dirpath
    (usable as $PHYLOCSF_BASE)."""
    data = _load("phylocsf_parameters.json")
    if dirpath is None:
        dirpath = tempfile.mkdtemp(prefix="pcsf_params_")
    pdir = os.path.join(dirpath, "PhyloCSF_Parameters")
    os.makedirs(pdir, exist_ok=True)
    for name in (sets or data["sets"]):
        s = data["sets"][name]
        with open(os.path.join(pdir, name + ".nh"), "w") as f:
            f.write(s["newick"] + "\n")
        for kind in ("coding", "noncoding"):
            with open(os.path.join(pdir, "%s_%s.ECM" % (name, kind)), "w") as f:
                f.write(ecm_text(data["ecm"][s[kind]]))
    return dirpath


def example_lines(fn):
    """Lines of PhyloCSF_Examples/<fn> (header + one sequence line per row)."""
    out = []
    for hdr, seq in _load("examples.json")[fn]:
        out.append(">" + hdr)
        out.append(seq)
    return out


def write_examples(dirpath):
    edir = os.path.join(dirpath, "PhyloCSF_Examples")
    os.makedirs(edir, exist_ok=True)
    for fn in _load("examples.json"):
        with open(os.path.join(edir, fn), "w") as f:
            f.write("\n".join(example_lines(fn)) + "\n")
    return edir
