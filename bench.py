#!/usr/bin/env python3
"""bench.py — headline benchmark of the PhyloCSF scoring hot path on B200.

Metric (BASELINE.json): codon-columns/s, one codon column = one alignment column triplet scored
under BOTH ECMs (coding + noncoding), 58mammals, --strategy=fixed.
Workload (BASELINE.json configs[1]): 100,000 synthetic alignments x 100 codons (300 nt, 58 species),
3 frames => 298 codon columns per alignment, 29.8 M codon columns per GPU per step, simulated
under the shipped 58mammals tree (half from the coding ECM, half from the noncoding ECM, rho = 1).
One step = one full pass of the hot path over that batch. Inputs are larger than L2 (1.7 GB of
leaf codes per pass), so no L2 flush is needed between iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--only headline|cfg5|mle|omega|outside]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

value   = device-resident throughput (leaf codes already in HBM), CUDA events, max over ranks
e2e     = the same pass through the C ABI from pinned HOST buffers (pcsf_score_alignments): H2D of the
          nucleotide rows in chunks overlapped with on-device pleaves, pruning, region reduction, D2H
roofline= the pruning kernel this workload runs (form and table level QUERIED from the library, not re-derived)
          against the FP64 tensor (DMMA) peak: `achieved` / `frac` count the flop the DMMA pipe EXECUTES;
          `algorithmic` is SURVEY 8(d)'s figure (all n-2 contractions per column and model), which exceeds the peak
          when memoised subtree tables serve part of them
cpu_baseline / --impl reference = the CPU oracle (oracle/, a restatement of the reference's OCaml
          path; the reference itself needs OCaml+GSL, absent here) on a bounded sample, all host cores
          (os.sched_getaffinity - torchrun's OMP_NUM_THREADS=1 is ignored on purpose)
extra   = the other BASELINE.json configurations, measured in the same run so that they are driver-witnessed:
          cfg5_strong_scaling (58mammals, fixed, 6 frames, 10 M codon columns in total split over the N ranks: strong
          scaling, end to end from host buffers, subtree tables rebuilt once inside the timed region),
          mle_cfg3 (120mammals, mle, 10,000 x 100 codons; N = 1 only), omega_cfg4 (100vertebrates, omega, 1,000 exon-length
          alignments x 3 frames through pcsf_omega_score; N = 1 only), outside_k6 (58mammals, outside algorithm + expected counts
          of all branches over 1 M resident columns; N = 1 only)
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


def host_cores():
    """Cores this process may run on. torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use
    every core the box gives the process, so the affinity mask decides, not that variable."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def materialize_params(sets):
    from tools import golden_params as gp
    return gp.materialize(tempfile.mkdtemp(prefix="pcsf_bench_"), sets=sets)


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
    from oracle import oracle as o

    ps = o.load_paramset(os.path.join(base, "PhyloCSF_Parameters", PSET), o.Options(strategy="fixed"))
    t = ps.tree
    L = o.lib()
    cores = host_cores() if nthreads <= 0 else nthreads
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
    return cols / dt, cores, "%d of the workload's alignments (%d codon columns, both ECMs), %.1f s on %d threads" % (nalign, cols, dt, cores)


def lookup_edges(ps, level):
    """How many of the n-2 contractions per column the table program of `level` serves by lookup (mirrors the
    library's tree-program builder: cherries not under the root; + a leaf next to them; + a further leaf)."""
    ch = np.asarray(ps.children).reshape(-1, 2)
    nl = ps.n_leaves
    root = 2 * nl - 2
    parent = {int(c): nl + i for i, pair in enumerate(ch) for c in pair}

    def sibling(v):
        return [int(c) for c in ch[parent[v] - nl] if int(c) != v][0]

    cherries = [nl + i for i, pair in enumerate(ch) if (pair < nl).all() and nl + i != root]
    triples = [v for v in cherries if parent[v] != root and sibling(v) < nl]
    quads = [parent[v] for v in triples if parent[parent[v]] != root and sibling(parent[v]) < nl]
    return (len(cherries) if level >= 2 else 0) + (len(triples) if level >= 3 else 0) + (len(quads) if level >= 4 else 0)


def perturb_nt(nt_dev, gen):
    """--data realistic: make the simulated block look like real alignments, on the device: 10 % of all nucleotides replaced
    by random ones (code tuples far from the conserved diagonal), every sixth (alignment, species) row missing, a gap run of
    3-60 nt in every fourth remaining row."""
    import torch

    A, n, L = nt_dev.shape
    dev = nt_dev.device
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    pos = torch.arange(L, device=dev)[None, None, :]
    for a0 in range(0, A, 8192):  # in slices: the random tensors are several times the size of the block itself
        blk = nt_dev[a0:a0 + 8192]
        m = blk.shape[0]
        sub = torch.rand((m, n, L), device=dev, generator=gen) < 0.10
        rnd = torch.randint(0, 4, (m, n, L), device=dev, generator=gen, dtype=torch.uint8)
        blk[sub] = acgt[rnd[sub].long()]
        del sub, rnd
        missing = torch.rand((m, n), device=dev, generator=gen) < (1.0 / 6.0)
        missing[:, 0] = False  # the reference species is always there
        blk[missing] = ord("-")
        start = torch.randint(0, L, (m, n), device=dev, generator=gen)
        length = torch.randint(3, 61, (m, n), device=dev, generator=gen)
        has = torch.rand((m, n), device=dev, generator=gen) < 0.25
        has[:, 0] = False
        gap = has[:, :, None] & (pos >= start[:, :, None]) & (pos < (start + length)[:, :, None])
        blk[gap] = ord("-")
        del gap
    return nt_dev


def simulate_nt(ctx, ps, n_align, n_codons, seed, dev, scales=(1.0,), realistic=False):
    """Synthetic alignments simulated under the context's two ECMs at the given tree scales (equal shares, coding
    first): uint8 [n_align, n_leaves, 3 * n_codons] on the HOST (pinned)."""
    import torch

    from phylocsf_b200 import simulate

    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    parents = simulate.parents_from_children(ps.n_leaves, ps.children)
    nbr = 2 * ps.n_leaves - 2
    groups = [(w, si) for w in (0, 1) for si in range(len(scales))]
    per = [n_align // len(groups) + (1 if g < n_align % len(groups) else 0) for g in range(len(groups))]
    parts = []
    for w in (0, 1):
        ctx.pt_build(w, list(scales))
    for (w, si), n in zip(groups, per):
        if n == 0:
            continue
        P = np.stack([ctx.pt_get(w, si, br) for br in range(nbr)])
        parts.append(simulate.simulate_codes(P, ps.qdiag(w)["prior"], parents, ps.n_leaves, n * n_codons, gen, dev))
    codes0 = torch.cat(parts, dim=0)
    nt_dev = simulate.codes_to_nt(codes0, n_align, n_codons)
    del codes0, parts
    if realistic:
        nt_dev = perturb_nt(nt_dev, gen)
    nt_host = torch.empty(nt_dev.shape, dtype=torch.uint8, pin_memory=True)
    nt_host.copy_(nt_dev)
    del nt_dev
    torch.cuda.synchronize()
    return nt_host


# ======================================================================================================================
# headline: configs[1]
# ======================================================================================================================
def run_headline(args, env):
    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host

    rank, world, local_rank, dev, dist = env["rank"], env["world"], env["local_rank"], env["dev"], env["dist"]
    base = env["base"]
    ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", PSET))
    ctx = pb.Context(local_rank)
    ps.install(ctx)
    ctx.stream_set(torch.cuda.current_stream().cuda_stream)
    A = args.alignments
    nt_host = simulate_nt(ctx, ps, A, N_CODONS, 42 + rank, dev, realistic=args.data == "realistic")
    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
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
    info = ctx.last_launch_info()  # what actually ran: kernel form and table level of the tree program
    level_models = [ctx.table_level(0), ctx.table_level(1)]

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
    info_e2e = ctx.last_launch_info()

    # sanity: the scores are finite and the two halves separate (coding half scores higher)
    score = (10.0 / np.log(10.0)) * (outs[0][0] - outs[0][1])
    if not os.environ.get("PCSF_BENCH_NO_SANITY") and args.data == "simulated":  # unset except for timing-only kernel ablations (tools/ab.sh)
        assert np.isfinite(score).all()
        f0 = score[0::FRAMES]
        assert f0[: A // 2].mean() > f0[A // 2:].mean()

    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(tt[0]), float(tt[1])
    line = None
    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * total_cols / (ms_step * 1e-3)
        e2e_value = world * total_cols / (ms_e2e / args.steps * 1e-3)
        peak, peak_note = dmma_peak_tflops()
        k_ms = float(np.mean(prune_ms))
        algorithmic = FLOP_PER_COLUMN * total_cols / (k_ms * 1e-3) / 1e12
        level = info["table_level"]
        n_lookup = lookup_edges(ps, level)
        executed_share = (ps.n_leaves - 2 - n_lookup) / (ps.n_leaves - 2)
        achieved = algorithmic * executed_share
        traffic, traffic_source = None, None
        for fn in ("r02_prune_wide_tabled_100k_ncu_summary.json", "prune_kernel_traffic.json"):
            try:
                j = json.load(open(os.path.join(ROOT, "profiles", fn)))
                traffic = j["dram_bytes_per_codon_column"] * total_cols
                traffic_source = "ncu dram__bytes_read.sum + dram__bytes_write.sum per codon column (profiles/%s, %s) x the columns of one launch" % (fn, j.get("workload", "20k-alignment run"))
                break
            except Exception:
                pass
        kernel = "pcsf::prune_wide_kernel" if info["form"] == "wide" else "pcsf::prune_kernel"
        config = dict(env["config"])
        config["kernel_form"] = info["form"]
        config["table_level"] = level
        line = {
            "metric": "codon_columns_per_sec", "value": value, "unit": "codon-columns/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (simulated under the shipped 58mammals tree and ECMs, seed 42+rank)" +
                                         (" + 10 % random substitutions, 1/6 of the species rows missing, gap runs (--data realistic)" if args.data == "realistic" else ""),
            "config": config,
            "e2e": {"value": e2e_value, "unit": "codon-columns/s", "h2d_bytes_per_step": int(nt_np.nbytes + aln_off.nbytes + aln_len.nbytes),
                    "d2h_bytes_per_step": int(outs[0].nbytes + outs[1].nbytes), "ms_per_step": ms_e2e / args.steps,
                    "table_level": info_e2e["table_level"], "kernel_form": info_e2e["form"],
                    "path": "pcsf_score_alignments: pinned host nucleotide rows -> chunked H2D overlapped with on-device pleaves + pruning + reduction -> D2H"},
            "gpu_launches": int(launches),
            "baseline_kind": "port: the CPU arm (cpu_baseline, --impl reference) is the oracle's C restatement of the reference's algorithm on all host "
                             "cores; the OCaml + GSL reference itself cannot be built in this image",
            "setup": setup,
            "clocks": clocks,
            "roofline": {"kernel": kernel, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source, "kernel_ms": k_ms,
                         "executed_share": executed_share, "table_level": level, "table_level_per_model": level_models,
                         "lookup_edges": n_lookup, "internal_edges": ps.n_leaves - 2,
                         "algorithmic": algorithmic, "algorithmic_frac": algorithmic / peak,
                         "flop_per_codon_column": FLOP_PER_COLUMN, "peak_source": peak_note,
                         "note": "achieved/frac = DMMA flop the kernel executes: (internal_edges - lookup_edges)/internal_edges of SURVEY 8(d)'s algorithmic flop; "
                                 "the other contractions are 512-byte lookups of memoised, bit-identical results; 'algorithmic' charges all of them"},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_oracle_throughput(nt_np, base, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "codon-columns/s", "cores": cores, "kind": "port", "sample": sample}
    ctx.close()
    del nt_host, out_lpr, out_elpr
    return line


# ======================================================================================================================
# extra: configs[4] as a strong-scaling job (fixed total split over the ranks)
# ======================================================================================================================
def run_cfg5(args, env):
    """58mammals, fixed, --frames=6, `--cfg5-alignments` contiguous alignments of 5,001 nt = 10 M codon columns IN TOTAL,
    split over the ranks by phylocsf_b200.shard.shard_bounds; every step is end to end from pinned host buffers. A fresh
    context, so the subtree-table level is the one a process scoring only this shard gets."""
    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host, shard

    rank, world, local_rank, dev, dist = env["rank"], env["world"], env["local_rank"], env["dev"], env["dist"]
    ps = host.ParamSet(os.path.join(env["base"], "PhyloCSF_Parameters", PSET))
    A_total, Lnt, F = args.cfg5_alignments, 5001, 6
    cols_per_aln = 2 * sum((Lnt - f) // 3 for f in range(3))
    lo, hi = shard.shard_bounds([cols_per_aln] * A_total, world)[rank]
    A = hi - lo
    ctx = pb.Context(local_rank)
    ps.install(ctx)
    ctx.stream_set(torch.cuda.current_stream().cuda_stream)
    nt_host = simulate_nt(ctx, ps, max(A, 1), Lnt // 3, 1000 + rank, dev)
    nt_np = nt_host.numpy()[:A]
    aln_off = np.arange(A, dtype=np.int64) * (ps.n_leaves * Lnt)
    aln_len = np.full(A, Lnt, dtype=np.int32)
    R = A * F
    out_lpr = torch.empty((2, max(R, 1)), dtype=torch.float64, pin_memory=True)
    out_elpr = torch.empty((2, max(R, 1)), dtype=torch.float64, pin_memory=True)
    outs = (out_lpr.numpy(), out_elpr.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if A:  # a rank without alignments (more ranks than alignments) only takes part in the barriers
            ctx.score_alignments(aln_off, aln_len, nt_np, F, [0, 1], out=outs)

    def timed(fn, n):
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt[0])
        return ms

    def cold():  # what a fresh process pays once: K1 for both models, the subtree tables, then the pass itself
        ctx.pt_build(0, [1.0])
        ctx.pt_build(1, [1.0])
        step()

    ctx.pt_build(0, [1.0])
    ctx.pt_build(1, [1.0])
    for _ in range(max(args.warmup, 1)):
        step()
    steady = timed(step, args.steps) / args.steps
    cold_ms = timed(cold, 1)
    info = ctx.last_launch_info() if A else {"form": "none", "table_level": 0}
    tables_ms = ctx.last_ms(5)
    total_cols = A_total * cols_per_aln
    res = None
    if world > 1:
        lv = torch.tensor([info["table_level"]], dtype=torch.int64, device=dev)
        dist.all_reduce(lv, op=dist.ReduceOp.MIN)
        min_level = int(lv[0])
    else:
        min_level = info["table_level"]
    if rank == 0:
        res = {"workload": "58mammals fixed, 6 frames, %d alignments x %d nt = %d codon columns in total, split over %d rank(s)" % (A_total, Lnt, total_cols, world),
               "scaling": "strong", "n_gpus": world, "codon_columns_total": total_cols, "codon_columns_per_gpu": A * cols_per_aln,
               "value": total_cols / (steady * 1e-3), "unit": "codon-columns/s", "ms_per_step": steady,
               "first_pass_ms_with_pt_build_and_tables": cold_ms, "value_first_pass": total_cols / (cold_ms * 1e-3),
               "table_level_rank0": info["table_level"], "table_level_min_over_ranks": min_level, "kernel_form": info["form"],
               "subtree_tables_ms_per_model": tables_ms,
               "path": "pcsf_score_alignments from pinned host buffers (H2D + pleaves + pruning + reduction + D2H), wall clock with barriers, max over ranks"}
    ctx.close()
    return res


# ======================================================================================================================
# extra: configs[2], mle on 120mammals
# ======================================================================================================================
def run_mle(args, env):
    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host

    pset, N, NCOD = "120mammals", args.mle_alignments, 100
    dev = env["dev"]
    base = materialize_params([pset])
    ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", pset))
    ctx = pb.Context(env["local_rank"])
    ps.install(ctx)
    scales = np.exp(np.linspace(np.log(0.3), np.log(3.0), 8))  # rho log-uniform in [0.3, 3]
    nt_host = simulate_nt(ctx, ps, N, NCOD, 4242, dev, scales=scales)
    nt = nt_host.numpy()
    aln_off = np.arange(N, dtype=np.int64) * (ps.n_leaves * 3 * NCOD)
    aln_len = np.full(N, 3 * NCOD, dtype=np.int32)
    best = None
    for it in range(2):  # first pass warms up (allocations); second is reported
        ctx.total_ms(reset=True)
        ctx.counters(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.batch_upload_alignments(aln_off, aln_len, nt, 1)  # host rows -> device pleaves (frames = 1)
        rho, lpr, elpr, st, ne = ctx.maximize_lpr_multi([0, 1])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = (dt, ctx.total_ms(), int(ne.sum()), st, rho, lpr, ctx.counters())
    dt, ms, evals, st, rho, lpr, cnt = best
    peak, _ = dmma_peak_tflops()
    n = ps.n_leaves
    res = {"workload": "120mammals mle (both ECMs maximised in the same rounds), %d alignments x %d codons, frames=1" % (N, NCOD),
           "alignments_per_s": N / dt, "seconds": dt, "evaluations_per_alignment_model": evals / (2.0 * N),
           "codon_column_evaluations_per_s": evals * NCOD / dt, "device_ms": ms,
           "device_share_of_wall": (ms["prune"] + ms["reduce"] + ms["pt_build"] + ms["subtree_tables"]) / (dt * 1e3),
           "failed_regions": int(((st[0] | st[1]) & ~64).astype(bool).sum()),
           "median_rho_coding": float(np.median(rho[0])),
           "pruning_algorithmic_tflops": evals * NCOD * (n - 2) * 8192 / (ms["prune"] * 1e-3) / 1e12 if ms["prune"] > 0 else None,
           "counters": cnt,
           "pt_build_tflops": cnt["pt_slots"] * 524288.0 / (ms["pt_build"] * 1e-3) / 1e12 if ms["pt_build"] > 0 else None,
           "dmma_total_tflops": (cnt["pt_slots"] * 524288.0 + evals * NCOD * (n - 2) * 8192.0) / ((ms["pt_build"] + ms["prune"]) * 1e-3) / 1e12,
           "dmma_peak_tflops": peak,
           "path": "pcsf_batch_upload_alignments (host rows) + pcsf_maximize_lpr_multi, wall clock"}
    res["pruning_algorithmic_frac"] = res["pruning_algorithmic_tflops"] / peak if res["pruning_algorithmic_tflops"] else None
    res["pt_build_frac"] = res["pt_build_tflops"] / peak if res["pt_build_tflops"] else None
    res["dmma_total_frac"] = res["dmma_total_tflops"] / peak
    res["note"] = ("K1 (P(t) build) and pruning both run on the DMMA pipe: dmma_total = (524,288 flop x P(t) slots built + 8,192 flop x (n-2) internal edges x "
                   "codon-column evaluations) / (K1 + pruning device time); pruning_algorithmic counts every evaluation although the rounds whose candidates all "
                   "regions share use cherry tables")
    ctx.close()
    return res


# ======================================================================================================================
# extra: configs[3], omega on 100vertebrates
# ======================================================================================================================
def run_omega(args, env):
    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host

    pset, N = "100vertebrates", args.omega_alignments
    dev = env["dev"]
    base = materialize_params([pset])
    ps = host.ParamSet(os.path.join(base, "PhyloCSF_Parameters", pset))
    ctx = pb.Context(env["local_rank"])
    ps.install(ctx)
    rng = np.random.default_rng(77)
    # exon-like lengths: log-normal, median 120 nt, clipped to 60..1500 nt, multiples of 3
    lens = (np.clip(np.exp(rng.normal(np.log(120.0), 0.7, size=N)), 60, 1500) // 3).astype(np.int64)
    nt_host = simulate_nt(ctx, ps, 1, int(lens.sum()), 777, dev)  # one long coding/noncoding stretch, cut into exons
    nt_all = nt_host.numpy()[0]  # [n_leaves, 3 * sum(lens)]
    lut = np.full(256, -1, dtype=np.int64)
    for ch, i in zip(b"ACGT", range(4)):
        lut[ch] = i
    regs = []
    pos = 0
    for L3 in lens:
        blk = lut[nt_all[:, 3 * pos:3 * (pos + L3)]]
        pos += L3
        for f in range(3):
            nc = (3 * L3 - f) // 3
            c = 16 * blk[:, f:f + 3 * nc:3] + 4 * blk[:, f + 1:f + 3 * nc:3] + blk[:, f + 2:f + 3 * nc:3]
            regs.append(np.ascontiguousarray(c.T.astype(np.uint8)))
    off = np.zeros(len(regs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([r.shape[0] for r in regs])
    codes = np.concatenate(regs, axis=0)
    # --omega-contexts N > 1: N scoring contexts on the one GPU, each with a share of the regions and a host thread of its
    # own (what the command line does with its batches). Measured slower than one context for this strategy, see --help.
    import threading

    ncx = max(1, args.omega_contexts)
    ctxs = [ctx] + [pb.Context(env["local_rank"]) for _ in range(ncx - 1)]
    for c in ctxs[1:]:
        ps.install(c)
    R = len(regs)
    cuts = [R * k // ncx for k in range(ncx + 1)]
    parts = []
    for k in range(ncx):
        o_k = off[cuts[k]:cuts[k + 1] + 1] - off[cuts[k]]
        parts.append((np.ascontiguousarray(o_k), np.ascontiguousarray(codes[off[cuts[k]]:off[cuts[k + 1]]])))
    # warm-up on a slice, then the timed pass over everything
    w = min(R // ncx, 90)
    for c, (o_k, c_k) in zip(ctxs, parts):
        host.omega_score(c, o_k[: w + 1], c_k[: o_k[w]])
    for c in ctxs:
        c.total_ms(reset=True)
        c.counters(reset=True)
    l0 = sum(c.launch_count for c in ctxs)
    out = [None] * ncx

    def work(k):
        out[k] = host.omega_score(ctxs[k], parts[k][0], parts[k][1])

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(ncx)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    score = np.concatenate([o_[0] for o_ in out])
    diag = np.concatenate([o_[1] for o_ in out])
    st = np.concatenate([o_[2] for o_ in out])
    ms = {k: sum(c.total_ms()[k] for c in ctxs) for k in ctxs[0].total_ms()}
    cnt = {k: sum(c.counters()[k] for c in ctxs) for k in ctxs[0].counters()}
    launches = sum(c.launch_count for c in ctxs) - l0
    peak, _ = dmma_peak_tflops()
    n = ps.n_leaves
    res = {"workload": "100vertebrates omega, --allScores --frames=3: %d exon-length alignments (median %d nt) = %d regions, %d codon columns"
                       % (N, int(np.median(lens) * 3), len(regs), int(off[-1])),
           "alignments_per_s": N / dt, "regions_per_s": len(regs) / dt, "seconds": dt, "device_ms": ms,
           "device_share_of_wall": (ms["prune"] + ms["reduce"] + ms["pt_build"] + ms["omega_eig"]) / (dt * 1e3),
           "scoring_contexts": ncx, "gpu_launches": launches, "failed_regions": int((st != 0).sum()), "counters": cnt,
           "evaluations_per_region": cnt["column_evaluations"] / max(1, int(off[-1])),
           "jacobi_sweeps_per_matrix": cnt["eig_sweeps"] / max(1, cnt["eig_matrices"]),
           "pt_build_tflops": cnt["pt_slots"] * 524288.0 / (ms["pt_build"] * 1e-3) / 1e12 if ms["pt_build"] > 0 else None,
           "pruning_tflops": cnt["column_evaluations"] * (n - 2) * 8192.0 / (ms["prune"] * 1e-3) / 1e12 if ms["prune"] > 0 else None,
           "dmma_peak_tflops": peak,
           "median_score_db": float(np.median(score)), "median_rho_H0": float(np.median(diag[:, 1])), "median_kappa_H0": float(np.median(diag[:, 2])),
           "path": "pcsf_omega_score (stages the regions, then kr_map for H0 and H1 in batched Brent rounds: K5 + K1 + K2..K4 per round) on "
                   "%d scoring context(s) of the one GPU, each with a host thread and an equal share of the regions; wall clock; device_ms are "
                   "summed over the contexts (kernels of different contexts overlap, so the share can exceed 1)" % ncx}
    for c in ctxs:
        c.close()
    return res


def run_outside(args, env):
    """K6 (pcsf_posteriors: outside algorithm + expected substitution counts of every branch; PhyloLik.ml:96-180) on simulated
    58mammals columns resident on the device. Not on the command line's path: reported so that the row has a
    driver-witnessed figure. A product is one 64 x 64 x columns contraction; (4 n_leaves - 6) of them per column."""
    import torch

    import phylocsf_b200 as pb
    from phylocsf_b200 import host, simulate

    N = args.outside_columns
    ps = host.ParamSet(os.path.join(env["base"], "PhyloCSF_Parameters", PSET))
    ctx = pb.Context(env["local_rank"])
    ps.install(ctx)
    ctx.pt_build(0, [1.0])
    gen = torch.Generator(device=env["dev"])
    gen.manual_seed(5)
    nbr = 2 * ps.n_leaves - 2
    P = np.stack([ctx.pt_get(0, 0, br) for br in range(nbr)])
    codes = simulate.simulate_codes(P, ps.qdiag(0)["prior"], simulate.parents_from_children(ps.n_leaves, ps.children), ps.n_leaves, N,
                                    gen, env["dev"]).cpu().numpy()
    ctx.batch_upload(np.array([0, N], dtype=np.int64), codes)
    for _ in range(2):
        ctx.posteriors(0, 0, nodes=[], ecounts=True, z=False)
    ms = []
    l0 = ctx.launch_count
    for _ in range(3):
        _, ec, _ = ctx.posteriors(0, 0, nodes=[], ecounts=True, z=False)
        ms.append(ctx.last_ms(0))
    launches = ctx.launch_count - l0
    ctx.close()
    t = float(np.median(ms)) * 1e-3
    flop = N * (4 * ps.n_leaves - 6) * 8192.0
    peak, src = dmma_peak_tflops()
    total = float(ec.sum())
    if abs(total - N * nbr) > 1e-6 * N * nbr:  # every branch's expected counts sum to the number of (possible) columns
        raise RuntimeError("expected counts sum to %r, not %r" % (total, N * nbr))
    return {"workload": "58mammals, %d simulated codon columns, expected counts of all %d branches" % (N, nbr), "columns_per_s": N / t,
            "kernel_ms": t * 1e3, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": flop / t / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flop / t / 1e12 / peak,
                         "peak_source": src, "flop_per_column": (4 * ps.n_leaves - 6) * 8192},
            "path": "pcsf_posteriors on a staged batch (transposed images + outside_dmma_kernel + outside_reduce_kernel), CUDA events "
                    "around the kernels; PCSF_K6_PLAIN=1 selects the plain-FP64 form"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--alignments", type=int, default=100000, help="alignments per GPU (workload default 100000)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline only")
    ap.add_argument("--data", default="simulated", choices=["simulated", "realistic"],
                    help="realistic: the simulated alignments with 10 %% random substitutions, missing species and gap runs (headline only)")
    ap.add_argument("--only", default="", choices=["", "headline", "cfg5", "mle", "omega", "outside"], help="run one part only (prints that part's JSON)")
    ap.add_argument("--outside-columns", type=int, default=1000000)
    ap.add_argument("--cfg5-alignments", type=int, default=1000, help="alignments of 5,001 nt in the strong-scaling job (1000 = 10 M codon columns)")
    ap.add_argument("--mle-alignments", type=int, default=10000)
    ap.add_argument("--omega-alignments", type=int, default=1000)
    ap.add_argument("--omega-contexts", type=int, default=1,
                    help="scoring contexts on the GPU for the omega leg, each with a host thread and a share of the regions (measured: 268 / 240 / 262 "
                         "alignments/s with 1 / 2 / 3 - the persistent pruning kernels of two contexts cannot share the SMs, so there is nothing to overlap)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "58mammals fixed strategy, 3 frames, %d synthetic alignments x %d codons per GPU" % (args.alignments, N_CODONS),
              "paramset": PSET, "strategy": "fixed", "frames": FRAMES, "alignments_per_gpu": args.alignments,
              "codon_columns_per_gpu_per_step": args.alignments * (3 * N_CODONS // 3 + 2 * ((3 * N_CODONS - 1) // 3)),
              "l2": "inputs larger than L2 (no flush needed)", "sharding": "alignments by rank, no collective on the data path",
              "subtree_tables": "PCSF_CHERRY_TABLES=%s (0 = by batch size; the level in effect is reported as table_level, queried from the library; "
                                "built once per P set outside the timed region, see setup)" % os.environ.get("PCSF_CHERRY_TABLES", "0")}

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args, config)
        return

    import torch

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    env = {"rank": rank, "world": world, "local_rank": local_rank, "dev": torch.device("cuda", local_rank), "dist": dist,
           "base": materialize_params([PSET]), "config": config}

    line = None
    if args.only in ("", "headline"):
        line = run_headline(args, env)
    extra = {}

    def guarded(name, fn):
        try:
            r = fn(args, env)
            if r is not None:
                extra[name] = r
        except Exception as e:  # an extra must never take the headline line down with it
            if rank == 0:
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    if not args.no_extra:
        if args.only in ("", "cfg5"):
            guarded("cfg5_strong_scaling", run_cfg5)
        if world == 1 and args.only in ("", "mle"):
            guarded("mle_cfg3", run_mle)
        if world == 1 and args.only in ("", "omega"):
            guarded("omega_cfg4", run_omega)
        if world == 1 and args.only in ("", "outside"):
            guarded("outside_k6", run_outside)
    if rank == 0:
        if line is None:
            line = {"only": args.only, "n_gpus": world}
        line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args, config):
    """The reference's own CPU algorithm (dense per-row ddot pruning) on this box's host cores. The
    OCaml+GSL reference cannot be built in this image, so this is the oracle port ("kind": "port"),
    all cores of the affinity mask (under torchrun too), each step a bounded sample of the same workload."""
    base = materialize_params([PSET])
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
            "config": config, "cores": cores,
            "baseline_kind": "port: the oracle's C restatement of the reference's algorithm; the OCaml + GSL reference itself cannot be built in this image",
            "cpu_baseline": {"value": value, "unit": "codon-columns/s", "cores": cores, "kind": "port",
                             "sample": "per step: " + sample + " (a bounded sample of the 100,000-alignment workload, same alignment shape); "
                                       "OCaml+GSL reference not buildable here, CPU restatement (oracle port) timed instead"},
            "e2e": {"value": value, "unit": "codon-columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
