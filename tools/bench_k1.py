#!/usr/bin/env python3
"""K1 in isolation: P(t) for `jobs` candidate tree scales x every branch of a parameter set (the shape one mle round has),
timed with the library's CUDA events.   python tools/bench_k1.py [jobs] [paramset] [n_models]
PCSF_LIB selects another build of the library (ablation builds: timing only), PCSF_K1_FORM the kernel form."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylocsf_b200 as pb  # noqa: E402
from phylocsf_b200 import host  # noqa: E402
from tools import golden_params as gp  # noqa: E402

jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pset = sys.argv[2] if len(sys.argv) > 2 else "120mammals"
base = gp.materialize(tempfile.mkdtemp(), sets=[pset])
ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", pset))
ctx = pb.Context(0)
ps.install(ctx)
rng = np.random.default_rng(1)
scales = np.exp(rng.uniform(np.log(0.1), np.log(5.0), size=jobs))
models = (np.arange(jobs) % 2).astype(np.int32)
nbr = 2 * ps.n_leaves - 2
ms = []
for it in range(4):
    ctx.pt_build_pairs(models, scales, check=False)
    ms.append(ctx.last_ms(2))
best = min(ms[1:])
print(json.dumps({"lib": os.environ.get("PCSF_LIB", "default"), "k1_form": os.environ.get("PCSF_K1_FORM", "1"), "paramset": pset, "jobs": jobs, "slots": jobs * nbr, "ms": ms,
                  "tflops": jobs * nbr * 524288.0 / (best * 1e-3) / 1e12}))
