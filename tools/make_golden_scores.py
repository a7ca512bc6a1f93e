#!/usr/bin/env python3
"""Writes tests/golden/oracle_scores.json: known answers of the CPU oracle (oracle/) for the shipped
example alignments and one seeded simulated batch, so the GPU parity tests also check the CUDA path
against committed numbers (not only against a live oracle run). Regenerate with
    python tools/make_golden_scores.py
after any deliberate change to the oracle; the oracle itself is pinned to the reference's own known
answers by tests/test_oracle_golden.py."""
import json
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

base = gp.materialize(tempfile.mkdtemp())
out = {"examples": {}, "simulated": {}}
for pset, fn, frames in (("12flies", "tal-AA.fa", 3), ("29mammals", "ALDH2.exon5.fa", 6)):
    ps = H.oracle_paramset(base, pset)
    regs, _ = H.example_codes(ps, fn, frames=frames)
    lpr, elpr = H.oracle_fixed(ps, regs)
    mle = []
    for c in regs[:3]:
        row = []
        for inst in (ps.model.coding_model, ps.model.noncoding_model):
            x, (lp, el) = o.maximize_lpr(lambda r: o.lpr_leaves(inst, c, r), lambda r: r[0], init=1.0)
            row.append([x, lp, el])
        mle.append(row)
    out["examples"][fn] = {"paramset": pset, "frames": frames, "ncols": [int(r.shape[0]) for r in regs],
                           "fixed_lpr": lpr.tolist(), "fixed_elpr_anc": elpr.tolist(), "mle_rho_lpr_elpr": mle}
ps = H.oracle_paramset(base, "58mammals")
rng = np.random.default_rng(20261017)
regs = []
for i, n in enumerate((100, 99, 99, 31, 128, 257)):
    inst = ps.model.coding_model if i % 2 == 0 else ps.model.noncoding_model
    c = o.simulate_columns(inst.model(1.0), n, rng)
    c[rng.random(c.shape) < 0.08] = 64  # gaps
    regs.append(c)
lpr, elpr = H.oracle_fixed(ps, regs)
out["simulated"] = {"paramset": "58mammals", "seed": 20261017, "codes_hex": [r.tobytes().hex() for r in regs],
                    "ncols": [int(r.shape[0]) for r in regs], "fixed_lpr": lpr.tolist(), "fixed_elpr_anc": elpr.tolist()}
path = os.path.join(ROOT, "tests", "golden", "oracle_scores.json")
json.dump(out, open(path, "w"))
print(path, os.path.getsize(path))
