#!/usr/bin/env python3
"""Secondary measurement (BASELINE.json configs[2]): 120mammals, --strategy=mle, frames=1,
N synthetic alignments x 100 codons simulated at tree scales spread log-uniformly over [0.3, 3].
Reports alignments/s, likelihood evaluations per (alignment, model), and codon-column evaluations/s,
with the library's own CUDA-event split between K1 (P(t) build) and K2-K4 (pruning + reduction).
Not the headline bench (bench.py); run on the GPU box:  python tools/bench_mle.py [N] [paramset]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylocsf_b200 as pb  # noqa: E402
from phylocsf_b200 import host, simulate  # noqa: E402
from tools import golden_params as gp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
PSET = sys.argv[2] if len(sys.argv) > 2 else "120mammals"
NCOD = 100
dev = torch.device("cuda", 0)
base = gp.materialize(tempfile.mkdtemp(), sets=[PSET])
ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", PSET))
ctx = pb.Context(0)
ps.install(ctx)
nbr = 2 * ps.n_leaves - 2
scales = np.exp(np.linspace(np.log(0.3), np.log(3.0), 8))
gen = torch.Generator(device=dev)
gen.manual_seed(42)
parents = simulate.parents_from_children(ps.n_leaves, ps.children)
parts = []
per = N // (2 * len(scales)) + 1
for w in (0, 1):
    ctx.pt_build(w, scales)
    prior = ps.qdiag(w)["prior"]
    for si in range(len(scales)):
        P = np.stack([ctx.pt_get(w, si, br) for br in range(nbr)])
        parts.append(simulate.simulate_codes(P, prior, parents, ps.n_leaves, per * NCOD, gen, dev))
codes = torch.cat(parts)[: N * NCOD].cpu().numpy()
off = np.arange(N + 1, dtype=np.int64) * NCOD
ctx.batch_upload(off, codes)
res = {}
torch.cuda.synchronize()
t0 = time.perf_counter()
tot_evals = 0
ms = {"pt_build": 0.0, "prune": 0.0, "reduce": 0.0}
if os.environ.get("PCSF_MLE_PER_MODEL"):  # one call per model (the older path), for comparison
    for m in (0, 1):
        rho, lpr, elpr, st, ne = ctx.maximize_lpr(m)
        tot_evals += int(ne.sum())
        ms["prune"] += ctx.last_ms(0)
        ms["reduce"] += ctx.last_ms(1)
        ms["pt_build"] += ctx.last_ms(2)
        res[m] = (rho, lpr, st)
else:  # both models' searches advance in the same rounds (what the command line does)
    rho, lpr, elpr, st, ne = ctx.maximize_lpr_multi([0, 1])
    tot_evals += int(ne.sum())
    ms["prune"] += ctx.last_ms(0)
    ms["reduce"] += ctx.last_ms(1)
    ms["pt_build"] += ctx.last_ms(2)
    for m in (0, 1):
        res[m] = (rho[m], lpr[m], st[m])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
score = (10 / np.log(10)) * (res[0][1] - res[1][1])
out = {
    "paramset": PSET, "alignments": N, "codons": NCOD, "seconds": dt, "alignments_per_s": N / dt,
    "evaluations_per_alignment_model": tot_evals / (2 * N), "codon_column_evaluations_per_s": tot_evals * NCOD / dt,
    "device_ms": ms, "failed_regions": int(((res[0][2] | res[1][2]) & ~64).astype(bool).sum()),
    "random_init_regions": int(((res[0][2] | res[1][2]) & 64).astype(bool).sum()),
    "median_rho_coding": float(np.median(res[0][0])), "mean_score_db": float(score.mean()),
}
print(json.dumps(out))
