#!/usr/bin/env python3
"""Throughput of K6 (pcsf_posteriors: outside algorithm + expected substitution counts) on simulated 58mammals columns.
    python tools/bench_posteriors.py [n_columns]"""
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

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
base = gp.materialize(tempfile.mkdtemp(), sets=["58mammals"])
ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", "58mammals"))
ctx = pb.Context(0)
ps.install(ctx)
ctx.pt_build(0, [1.0])
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
nbr = 2 * ps.n_leaves - 2
P = np.stack([ctx.pt_get(0, 0, br) for br in range(nbr)])
codes = simulate.simulate_codes(P, ps.qdiag(0)["prior"], simulate.parents_from_children(ps.n_leaves, ps.children), ps.n_leaves, N, gen, dev).cpu().numpy()
ctx.batch_upload(np.array([0, N], dtype=np.int64), codes)
ctx.posteriors(0, 0, nodes=[], ecounts=True)
t0 = time.perf_counter()
post, ec, z = ctx.posteriors(0, 0, nodes=[], ecounts=True)
dt = time.perf_counter() - t0
ms = ctx.last_ms(0)
flop = N * (4 * ps.n_leaves - 6) * 8192.0
print(json.dumps({"columns": N, "kernel_ms": ms, "wall_s": dt, "columns_per_s": N / (ms * 1e-3), "tflops_plain_fp64": flop / (ms * 1e-3) / 1e12,
                  "ecounts_total": float(ec.sum()), "expected": float(N * nbr)}))
