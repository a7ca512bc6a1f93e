#!/usr/bin/env python3
"""End-to-end timing of the drop-in command line on simulated alignment files (GPU box):
    python tools/bench_cli.py <paramset> <n_alignments> <codons> [repeat] -- <PhyloCSF flags...>
Writes N multi-FASTA files simulated under the parameter set's coding/noncoding ECMs (half each), lists
each of them `repeat` times (default 1; the files stay in the page cache), runs
phylocsf_b200/bin/PhyloCSF --files on the list and reports wall-clock throughput of the whole process
(start-up and CUDA context creation included). This is the whole drop-in path: file reading, batching,
GPU scoring, report. PCSF_HOST_PROFILE=1 adds the reader / appender split on stderr."""
import json
import os
import subprocess
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

pset, N, ncod = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
REPEAT = int(sys.argv[4]) if sys.argv[4] != "--" else 1
flags = sys.argv[sys.argv.index("--") + 1:]
base = gp.materialize(tempfile.mkdtemp(), sets=[pset])
ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", pset))
ctx = pb.Context(0)
ps.install(ctx)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(42)
nbr = 2 * ps.n_leaves - 2
parents = simulate.parents_from_children(ps.n_leaves, ps.children)
parts = []
for w in (0, 1):
    ctx.pt_build(w, [1.0])
    P = np.stack([ctx.pt_get(w, 0, br) for br in range(nbr)])
    parts.append(simulate.simulate_codes(P, ps.qdiag(w)["prior"], parents, ps.n_leaves, (N // 2 + 1) * ncod, gen, dev))
codes = torch.cat(parts)[: N * ncod]
nt = simulate.codes_to_nt(codes, N, ncod).cpu().numpy()
ctx.close()
d = tempfile.mkdtemp(prefix="pcsf_cli_bench_")
names = []
for a in range(N):
    fn = os.path.join(d, "aln%06d.fa" % a)
    with open(fn, "w") as f:
        for l, lab in enumerate(ps.leaf_labels):
            f.write(">%s\n%s\n" % (lab, nt[a, l].tobytes().decode()))
    names.append(fn)
lst = os.path.join(d, "list.txt")
open(lst, "w").write(("\n".join(names) + "\n") * REPEAT)
N *= REPEAT
env = dict(os.environ, PHYLOCSF_BASE=base)
t0 = time.perf_counter()
wrap = os.environ.get("PCSF_NCU_WRAP", "").split()  # e.g. an ncu command line, for per-kernel time splits
if os.environ.get("PCSF_MULTI"):  # one process per GPU over the same list (tools/phylocsf_multi.py)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "phylocsf_multi.py"), "--gpus", os.environ["PCSF_MULTI"], pset, lst] + flags
else:
    cmd = wrap + [os.path.join(ROOT, "phylocsf_b200", "bin", "PhyloCSF"), pset, lst, "--files"] + flags
r = subprocess.run(cmd, env=env, capture_output=True, text=True)
dt = time.perf_counter() - t0
lines = r.stdout.splitlines()
print(json.dumps({"paramset": pset, "alignments": N, "codons": ncod, "flags": flags, "rc": r.returncode, "seconds": dt,
                  "alignments_per_s": N / dt, "output_lines": len(lines), "first_line": lines[0] if lines else r.stderr[:300],
                  "stderr_tail": r.stderr[-300:]}))
