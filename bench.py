#!/usr/bin/env python3
"""bench.py — headline benchmark of the PhyloCSF scoring hot path on B200.

Metric (BASELINE.json): codon-columns/s, one codon column = one alignment column triplet scored
under BOTH ECMs (coding + noncoding), 58mammals, --strategy=fixed.
Workload (BASELINE.json configs[1]): 100,000 synthetic alignments x 100 codons (300 nt, 58 species),
3 frames => 298 codon columns per alignment, 29.8 M codon columns per GPU per step, simulated
under the shipped 58mammals tree (half from the coding ECM, half from the noncoding ECM, rho = 1).
One step = one full pass of the hot path over that batch. Inputs are larger than L2 (1.7 GB of
leaf codes per pass), so no L2 flush is needed between iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

value   = device-resident throughput (leaf codes already in HBM), CUDA events, max over ranks
e2e     = the same pass through the C ABI from pinned HOST buffers (pcsf_score_alignments): H2D of the
          nucleotide rows in chunks overlapped with on-device pleaves, pruning, region reduction, D2H
roofline= the pruning kernel (wide form, prune_wide_kernel: the one this workload runs) against the FP64 tensor (DMMA) peak
cpu_baseline / --impl reference = the CPU oracle (oracle/, a restatement of the reference's OCaml
          path; the reference itself needs OCaml+GSL, absent here) on a bounded sample, all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PSET = "58mammals"
FLOP_PER_COLUMN = 2 * 56 * 2 * 64 * 64  # 2 ECMs x (n-2) internal edges x 2*64*64 = 917,504
N_CODONS = 100
FRAMES = 3


def dmma_peak_tflops():
    """FP64 tensor peak. MEASURED_PEAKS.json (driver-written) carries only HBM GB/s and bf16 TF/s, so
    the denominator is our own DMMA microbenchmark on this pool (tools/microbench/fp64_peak.cu,
    profiles/r01_fp64_peak_microbench.json); nominal B200 FP64 tensor is 40 TFLOP/s."""
    p = os.path.join(ROOT, "profiles", "r01_fp64_peak_microbench.json")
    try:
        return float(json.load(open(p))["dmma_peak_tflops"]), "measured DMMA microbench (profiles/r01_fp64_peak_microbench.json); MEASURED_PEAKS.json has no FP64 figure; nominal 40"
    except Exception:
        return 40.0, "nominal B200 FP64 tensor peak (no measurement file)"


class ClockSampler:
    def __init__(self, uuid):
        self.uuid = uuid
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--id=" + self.uuid, "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def materialize_params():
    from tools import golden_params as gp
    return gp.materialize(tempfile.mkdtemp(prefix="pcsf_bench_"), sets=[PSET])


def frame_codes_numpy(nt, frames):
    """pleaves for AsIs frames on the host (cpu_baseline leg only). nt: uint8 [A, n_leaves, L] ASCII.
    Returns (region_off, codes [total, n_leaves]) in alignment-major, frame order."""
    A, n, L = nt.shape
    lut = np.full(256, -1, dtype=np.int64)
    for ch, i in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut[ch] = i
    idx = lut[nt]
    chunks, offs = [], [0]
    per_frame = []
    for f in range(frames):
        nc = (L - f) // 3
        i1, i2, i3 = idx[:, :, f:f + 3 * nc:3], idx[:, :, f + 1:f + 3 * nc:3], idx[:, :, f + 2:f + 3 * nc:3]
        c = 16 * i1 + 4 * i2 + i3
        c[(i1 < 0) | (i2 < 0) | (i3 < 0)] = 64
        per_frame.append(np.ascontiguousarray(c.transpose(0, 2, 1)).astype(np.uint8))  # [A, nc, n]
    for a in range(A):
        for f in range(frames):
            chunks.append(per_frame[f][a])
            offs.append(offs[-1] + per_frame[f][a].shape[0])
    return np.array(offs, dtype=np.int64), np.concatenate(chunks, axis=0)


def cpu_oracle_throughput(nt_sample, base, target_seconds, nthreads=0):
    """Time the CPU oracle (restatement of PhyloLik.ensure_alpha's dense ddot form) on a bounded
    sample, one region per OpenMP task, both ECMs. Returns (columns/s, cores, sample description)."""
    import ctypes

    from oracle import oracle as o

    ps = o.load_paramset(os.path.join(base, "PhyloCSF_Parameters", PSET), o.Options(strategy="fixed"))
    t = ps.tree
    L = o.lib()
    cores = L.oracle_max_threads() if nthreads <= 0 else nthreads
    ch = t.children_array()
    models = []
    for inst in (ps.model.coding_model, ps.model.noncoding_model):
        m = inst.model(1.0)
        models.append((np.ascontiguousarray(m.pms), np.ascontiguousarray(m.prior())))

    def run(nalign):
        off, codes = frame_codes_numpy(nt_sample[:nalign], FRAMES)
        R = off.size - 1
        lpr, elpr = np.empty(R), np.empty(R)
        t0 = time.perf_counter()
        for pms, prior in models:
            L.oracle_lpr_batch(t.n_leaves, ch.ctypes.data, o._dp(pms), None, o._dp(prior), 64, R, off.ctypes.data,
                               codes.ctypes.data, o._dp(lpr), o._dp(elpr), cores)
        return int(off[-1]), time.perf_counter() - t0

    probe = min(nt_sample.shape[0], max(2 * cores, 16))
    cols, dt = run(probe)
    rate = cols / dt
    nalign = int(min(nt_sample.shape[0], max(probe, target_seconds * rate / (cols / probe))))
    cols, dt = run(nalign)
    cpu_oracle_throughput.last_seconds = dt
    return cols / dt, cores, "%d of the workload's alignments (%d codon columns, both ECMs), %.1f s" % (nalign, cols, dt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--alignments", type=int, default=100000, help="alignments per GPU (workload default 100000)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "58mammals fixed strategy, 3 frames, %d synthetic alignments x %d codons per GPU" % (args.alignments, N_CODONS),
              "paramset": PSET, "strategy": "fixed", "frames": FRAMES, "alignments_per_gpu": args.alignments,
              "codon_columns_per_gpu_per_step": args.alignments * (3 * N_CODONS // 3 + 2 * ((3 * N_CODONS - 1) // 3)),
              "l2": "inputs larger than L2 (no flush needed)", "sharding": "alignments by rank, no collective on the data path",
              "subtree_tables": "PCSF_CHERRY_TABLES=%s (0 = by batch size: levels 2-4 for this workload; built once per P set outside the timed region, see setup)"
                                % os.environ.get("PCSF_CHERRY_TABLES", "0")}

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args, config)
        return

    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host, simulate

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    base = materialize_params()
    ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", PSET))
    ctx = pb.Context(local_rank)
    ps.install(ctx)
    ctx.stream_set(torch.cuda.current_stream().cuda_stream)
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    nbr = 2 * ps.n_leaves - 2

    # ---- synthetic alignments (not timed) ----
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    parents = simulate.parents_from_children(ps.n_leaves, ps.children)
    A = args.alignments
    halves = [A // 2, A - A // 2]
    parts = []
    for w in (0, 1):
        P = np.stack([ctx.pt_get(w, 0, br) for br in range(nbr)])
        prior = ps.qdiag(w)["prior"]
        parts.append(simulate.simulate_codes(P, prior, parents, ps.n_leaves, halves[w] * N_CODONS, gen, dev))
    codes0 = torch.cat(parts, dim=0)
    nt_dev = simulate.codes_to_nt(codes0, A, N_CODONS)  # [A, n_leaves, 300]
    del codes0, parts
    nt_host = torch.empty(nt_dev.shape, dtype=torch.uint8, pin_memory=True)
    nt_host.copy_(nt_dev)
    del nt_dev
    torch.cuda.synchronize()
    Lnt = 3 * N_CODONS
    aln_off = np.arange(A, dtype=np.int64) * (ps.n_leaves * Lnt)
    aln_len = np.full(A, Lnt, dtype=np.int32)
    nt_np = nt_host.numpy()

    ctx.batch_upload_alignments(aln_off, aln_len, nt_np, FRAMES)
    R, total_cols = ctx.nregions, ctx.ncols
    # one-time work per P set, outside the timed region like the reference's memoised P(t): K1 and the subtree tables
    setup = {"pt_build_ms_per_model": ctx.last_ms(2)}
    out_lpr = torch.empty((2, R), dtype=torch.float64, pin_memory=True)
    out_elpr = torch.empty((2, R), dtype=torch.float64, pin_memory=True)
    outs = (out_lpr.numpy(), out_elpr.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        ctx.lpr_all([0, 1], out=outs)

    def step_e2e():
        ctx.score_alignments(aln_off, aln_len, nt_np, FRAMES, [0, 1], out=outs)

    uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    sampler = ClockSampler(uuid)

    # ---- device-resident: value ----
    step_resident()
    setup["subtree_tables_ms_per_model"] = ctx.last_ms(5)  # built by the first pass that scores enough columns
    for _ in range(args.warmup - 1):
        step_resident()
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prune_ms = []
    e0.record()
    for _ in range(args.steps):
        step_resident()
        prune_ms.append(ctx.last_ms(0))
    e1.record()
    barrier()
    launches = ctx.launch_count - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- end to end from pinned host buffers ----
    step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record()
    for _ in range(args.steps):
        step_e2e()
    e3.record()
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(e2.elapsed_time(e3), wall_e2e)  # host-side staging between launches counts too

    # sanity: the scores are finite and the two halves separate (coding half scores higher)
    score = (10.0 / np.log(10.0)) * (outs[0][0] - outs[0][1])
    if not os.environ.get("PCSF_BENCH_NO_SANITY"):  # unset except for timing-only kernel ablations (tools/ab.sh)
        assert np.isfinite(score).all()
        f0 = score[0::FRAMES]
        assert f0[: A // 2].mean() > f0[A // 2:].mean()

    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(tt[0]), float(tt[1])
    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * total_cols / (ms_step * 1e-3)
        e2e_value = world * total_cols / (ms_e2e / args.steps * 1e-3)
        peak, peak_note = dmma_peak_tflops()
        k_ms = float(np.mean(prune_ms))
        achieved = FLOP_PER_COLUMN * total_cols / (k_ms * 1e-3) / 1e12
        # cherry tables: the contraction above every cherry of the tree is served from a memoised table, so the kernel
        # executes (n-2-cherries) of the (n-2) contractions the algorithmic figure counts
        ch = np.asarray(ps.children).reshape(-1, 2)
        nl = ps.n_leaves
        root = 2 * nl - 2
        parent = {int(c): nl + i for i, pair in enumerate(ch) for c in pair}
        cherries = [nl + i for i, pair in enumerate(ch) if (pair < nl).all() and nl + i != root]
        # a cherry whose sibling is a leaf, under a node that has an edge of its own: one lookup replaces two contractions
        triples = [v for v in cherries if parent[v] != root and min(int(c) for c in ch[parent[v] - nl] if int(c) != v) < nl]
        def sibling(v):
            return [int(c) for c in ch[parent[v] - nl] if int(c) != v][0]
        # ... and a further leaf next to that (a caterpillar of four)
        quads = [parent[v] for v in triples if parent[parent[v]] != root and sibling(parent[v]) < nl]
        mode = os.environ.get("PCSF_CHERRY_TABLES", "0")
        tabled = mode != "1" and os.environ.get("PCSF_WIDE", "-1") != "0"
        level = 0 if not tabled else {"2": 3, "3": 2, "4": 4}.get(mode, 4 if total_cols >= 5000000 else 3 if total_cols >= 1000000 else 2 if total_cols >= 50000 else 0)
        n_lookup_edges = (len(cherries) if level >= 2 else 0) + (len(triples) if level >= 3 else 0) + (len(quads) if level >= 4 else 0)
        executed_share = (nl - 2 - n_lookup_edges) / (nl - 2)
        traffic = None  # dram__bytes_read+write of one launch: ncu-measured bytes per codon column x columns
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "prune_kernel_traffic.json")))["dram_bytes_per_codon_column"] * total_cols
        except Exception:
            pass
        line = {
            "metric": "codon_columns_per_sec", "value": value, "unit": "codon-columns/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (simulated under the shipped 58mammals tree and ECMs, seed 42+rank)",
            "config": config,
            "e2e": {"value": e2e_value, "unit": "codon-columns/s", "h2d_bytes_per_step": int(nt_np.nbytes + aln_off.nbytes + aln_len.nbytes),
                    "d2h_bytes_per_step": int(outs[0].nbytes + outs[1].nbytes), "ms_per_step": ms_e2e / args.steps,
                    "path": "pcsf_score_alignments: pinned host nucleotide rows -> chunked H2D overlapped with on-device pleaves + pruning + reduction -> D2H"},
            "gpu_launches": int(launches),
            "setup": setup,
            "clocks": clocks,
            "roofline": {"kernel": "pcsf::prune_wide_kernel", "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel_ms": k_ms,
                         "flop_per_codon_column": FLOP_PER_COLUMN, "peak_source": peak_note,
                         "executed": {"share_of_algorithmic_flop": executed_share, "tflops": achieved * executed_share,
                                      "frac_of_peak": achieved * executed_share / peak,
                                      "note": "achieved/frac use the algorithmic flop of SURVEY 8(d); with the subtree tables %d of the %d contractions per column are "
                                              "covered by 512-byte lookups of memoised results (bit-identical), so frac can exceed 1; 'executed' is what the DMMA pipe runs"
                                              % (n_lookup_edges, nl - 2)}},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_oracle_throughput(nt_np, base, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "codon-columns/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, config):
    """The reference's own CPU algorithm (dense per-row ddot pruning) on this box's host cores. The
    OCaml+GSL reference cannot be built in this image, so this is the oracle port ("kind": "port"),
    all OpenMP threads, each step a bounded sample of the same workload."""
    base = materialize_params()
    from oracle import oracle as o

    ps = o.load_paramset(os.path.join(base, "PhyloCSF_Parameters", PSET), o.Options(strategy="fixed"))
    rng = np.random.default_rng(42)
    n_align = 400
    mc, mn = ps.model.coding_model.model(1.0), ps.model.noncoding_model.model(1.0)
    codes = np.concatenate([o.simulate_columns(mc, (n_align // 2) * N_CODONS, rng), o.simulate_columns(mn, (n_align - n_align // 2) * N_CODONS, rng)])
    table = np.array([[ord(c) for c in o.codon_of_index(i)] for i in range(64)], dtype=np.uint8)
    nt = table[codes].reshape(n_align, N_CODONS, ps.tree.n_leaves, 3).transpose(0, 2, 1, 3).reshape(n_align, ps.tree.n_leaves, 3 * N_CODONS)
    nt = np.ascontiguousarray(nt)
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals, secs = [], []
    sample = ""
    cores = 1
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_oracle_throughput(nt, base, per_step)
        if i >= args.warmup:
            vals.append(v)
            secs.append(cpu_oracle_throughput.last_seconds)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "codon_columns_per_sec", "value": value, "unit": "codon-columns/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (simulated under the shipped 58mammals tree and ECMs)",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "codon-columns/s", "cores": cores, "kind": "port",
                             "sample": "per step: " + sample + "; OCaml+GSL reference not buildable here, CPU restatement timed instead"},
            "e2e": {"value": value, "unit": "codon-columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
