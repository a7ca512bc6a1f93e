#!/usr/bin/env python3
"""compute-sanitizer workload for the kernels changed in round 2's last session (GPU box):
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitize_k0_k5.py
K0 (pcsf_k0.cuh): ragged alignments at odd byte offsets, 1 / 3 / 6 frames, one alignment spanning several shared-memory
tiles, the nucleotide buffer ending exactly at the last row (the kernel reads whole words: the word guard is what memcheck
checks), codes compared byte for byte with the oracle's pleaves; the pipelined pcsf_score_alignments with many chunks.
K5: pcsf_omega_models_set cold and cached (warm starts), eigensystems checked against the oracle's Q."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcsf_helpers as H  # noqa: E402
import test_k0_emulation as E  # noqa: E402
from oracle import oracle as o  # noqa: E402
from tools import golden_params as gp  # noqa: E402

base = gp.materialize(tempfile.mkdtemp(), sets=["29mammals"])
ps = H.oracle_paramset(base, "29mammals")
n = ps.tree.n_leaves
rng = np.random.default_rng(5)
ctx = H.make_context(ps)
alphabet = np.array(list("ACGTacgtNn-"))
lens = [0, 1, 2, 3, 17, 48, 49, 301, 3001]
alns = [["".join(alphabet[rng.integers(0, len(alphabet), size=L)]) for _ in range(n)] for L in lens]
nt, off = E.pack(alns, rng)
for frames in (1, 3, 6):
    want, roff = E.oracle_frames(alns, frames)
    ctx.batch_upload_alignments(off, lens, nt, frames)
    assert np.array_equal(ctx.batch_codes(), want), frames
print("K0 codes ok")
ctx.pt_build(0, [1.0])
ctx.pt_build(1, [1.0])
os.environ["PCSF_CHUNK_COLS"] = "300"
lpr, elpr, st = ctx.score_alignments(off, lens, nt, 6, [0, 1])
ctx.batch_upload_alignments(off, lens, nt, 6)
two = ctx.lpr_all([0, 1])
assert np.array_equal(lpr, two[0], equal_nan=True)
print("pipelined path ok")
qs = np.array([[2.5, 1.0, 1.0] + [1.0] * 9, [1.7, 0.2, 0.01, 1.3, 0.8, 1.1, 0.9, 1.2, 0.7, 1.05, 0.95, 1.4]])
assert (ctx.omega_models_set(3, qs) == 0).all()
for i, v in enumerate(qs):
    d = ctx.model_get(3 + i)
    assert np.abs(d["S"] @ np.diag(d["lam"]) @ d["Sinv"] - o.omega_q(list(v))).max() < 5e-13
ctx.omega_cache_reset(2)
for k in range(3):  # a cold solve, then two warm ones on nearby matrices
    q2 = qs.copy()
    q2[:, 0] += 0.05 * k
    assert (ctx.omega_models_set_cached(3, q2, [0, 1]) == 0).all()
    for i, v in enumerate(q2):
        d = ctx.model_get(3 + i)
        assert np.abs(d["S"] @ np.diag(d["lam"]) @ d["Sinv"] - o.omega_q(list(v))).max() < 5e-13
print("K5 ok")
ctx.close()
