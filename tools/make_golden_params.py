#!/usr/bin/env python3
"""Pack the reference's shipped *input data* (trees, ECMs, example alignments) into
tests/golden/ so that tests, smoke() and bench.py can run where /root/reference
does not exist (the GPU box).

Run in the build container only:  python tools/make_golden_params.py [/root/reference]

What it writes (data only, no reference source code):
  tests/golden/phylocsf_parameters.json  - all 14 PhyloCSF_Parameters sets; ECMs de-duplicated
                                           by content hash; numbers kept as their original text
                                           tokens, so the re-emitted .ECM / .nh hold exactly the
                                           reference's tokens (blanks are normalised: 33 of the 42
                                           files differ from the originals in whitespace only;
                                           tests/test_host_model.py pins the token identity)
  tests/golden/examples.json             - the three PhyloCSF_Examples/*.fa alignments
                                           (raw header text + sequence per row)
Formats: SURVEY.md Appendix B (src/ECM.ml:18-72, lib/CamlPaml/NewickLexer.mll:5-14).
"""
import hashlib
import json
import os
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))
out_dir = os.path.join(here, "..", "tests", "golden")
pdir = os.path.join(ref, "PhyloCSF_Parameters")

sets = sorted(f[:-3] for f in os.listdir(pdir) if f.endswith(".nh"))
ecms = {}
out_sets = {}


def pack_ecm(path):
    raw = open(path).read()
    h = hashlib.md5(raw.encode()).hexdigest()
    if h not in ecms:
        lines = raw.split("\n")
        s_rows = [ln.split() for ln in lines[:63]]
        assert all(len(r) == i + 1 for i, r in enumerate(s_rows)), path
        assert lines[63].strip() == ""
        pi = lines[64].split()
        assert len(pi) == 64
        codons = " ".join(lines[67:71]).split()
        assert len(codons) == 64
        ecms[h] = {"s_lower": s_rows, "pi": pi, "codons": codons}
    return h


for name in sets:
    newick = " ".join(open(os.path.join(pdir, name + ".nh")).read().split())
    out_sets[name] = {
        "newick": newick,
        "coding": pack_ecm(os.path.join(pdir, name + "_coding.ECM")),
        "noncoding": pack_ecm(os.path.join(pdir, name + "_noncoding.ECM")),
    }

with open(os.path.join(out_dir, "phylocsf_parameters.json"), "w") as f:
    json.dump({"sets": out_sets, "ecm": ecms}, f, separators=(",", ":"))

examples = {}
edir = os.path.join(ref, "PhyloCSF_Examples")
for fn in sorted(os.listdir(edir)):
    rows = []
    for line in open(os.path.join(edir, fn)):
        line = line.rstrip("\n")
        if line.startswith(">"):
            rows.append([line[1:], ""])
        elif line.strip():
            rows[-1][1] += line.strip()
    examples[fn] = rows
with open(os.path.join(out_dir, "examples.json"), "w") as f:
    json.dump(examples, f, indent=0)
print("sets:", len(out_sets), "unique ECMs:", len(ecms), "examples:", list(examples))
