#!/usr/bin/env python3
"""Debug tool (GPU box): per-warp GEMM begin/end clock marks of CTA 0 of the pruning kernel, to see
how the eight warps of a CTA interleave their DMMA phases and epilogues. Builds a separate
-DPCSF_TIMELINE library (the production library carries no instrumentation).

    python tools/timeline.py [out.json]        (env PCSF_SKEW_NS is honoured)
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phylocsf_b200 import build as B  # noqa: E402

lib = os.path.join(ROOT, "gpurun_out", "libpcsf_timeline.so")
os.makedirs(os.path.dirname(lib), exist_ok=True)
subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + ["-DPCSF_TIMELINE", "-o", lib] + B.sources())

from phylocsf_b200 import _native as N  # noqa: E402

N.LIB_PATH = lib
import phylocsf_b200 as pb  # noqa: E402
from phylocsf_b200 import host  # noqa: E402
from tools import golden_params as gp  # noqa: E402

base = gp.materialize(tempfile.mkdtemp(), sets=["58mammals"])
ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", "58mammals"))
ctx = pb.Context(0)
ps.install(ctx)
ctx.pt_build(0, [1.0])
ctx.pt_build(1, [1.0])
rng = np.random.default_rng(0)
ncols = 148 * 384 * 2
codes = rng.integers(0, 64, size=(ncols, 58)).astype(np.uint8)
codes[:, :] = codes[:, :1]  # conserved columns (no underflow)
ctx.batch_upload(np.array([0, ncols], dtype=np.int64), codes)
ctx.lpr_all([0, 1])
L = N.load()
cap = 4096
L.pcsf_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
assert L.pcsf_debug_timeline(ctx._h, cap, None) == 0
ctx.lpr_all([0, 1])
NW = 12 if os.environ.get('PCSF_WIDE', '0') not in ('', '0') else 8
out = np.zeros((16, cap, 2), dtype=np.int64)
assert L.pcsf_debug_timeline(ctx._h, cap, out.ctypes.data_as(ctypes.c_void_p)) == 0
t0 = out[:, 0, 1].min()
res = {}
for w in range(NW):
    ev = out[w]
    ev = ev[ev[:, 0] >= 0]
    res[w] = [[int(c), int(t - t0)] for c, t in ev[:400]]
# marks per GEMM op: 8*op+0 waiting for the P image, +1 GEMM begin, +2 GEMM end, +3 multiplicand ready
# (LEAF/POP ops only), +4 epilogue end
names = {(0, 1): "wait_P", (1, 2): "gemm", (2, 3): "wait_M", (3, 4): "mul", (2, 4): "push", (4, 0): "between_ops"}
summ = {}
for w in range(NW):
    ev = res[w]
    seg = {}
    for i in range(len(ev) - 1):
        k = (ev[i][0] & 7, ev[i + 1][0] & 7)
        seg.setdefault(names.get(k, str(k)), []).append(ev[i + 1][1] - ev[i][1])
    summ[w] = {k: [round(float(np.mean(v)), 1), len(v)] for k, v in seg.items()}
for w in range(4):
    for w2 in range(w + 4, NW, 4):
        a = np.array([e[1] for e in res[w] if (e[0] & 7) == 1][:90])
        b = np.array([e[1] for e in res[w2] if (e[0] & 7) == 1][:90])
        n = min(len(a), len(b))
        summ["gemm_begin_offset_w%d_w%d" % (w, w2)] = {"mean": float(np.mean(b[:n] - a[:n])), "min": int((b[:n] - a[:n]).min()), "max": int((b[:n] - a[:n]).max())}
print(json.dumps(summ, indent=1))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "timeline.json")
json.dump({"summary": summ, "events": res}, open(path, "w"))
