#!/bin/bash
# Round-2 evidence run (on the GPU box, one GPU): launch lists and ncu --set full captures for every kernel family.
#   bash tools/profile_r02.sh [part ...]     parts: launches prune_sim prune_real k1 k5 k0 tables outside
# Outputs go to gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into profiles/r02_*_ncu_summary.json here.
set -u
mkdir -p gpurun_out
NCU_FULL="ncu --set full --clock-control none --import-source on"
parts=${@:-launches prune_sim prune_real k1 k5 k0 tables outside}
for part in $parts; do
  case $part in
    launches)   # headline workload at full size: which kernels make up a step (cold-cache, serialised times: shares only)
      ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:prune|frame_codes|region_reduce|make_segments|pt_build|subtree_table' -c 200 --csv --log-file gpurun_out/r02_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1 ;;
    prune_sim)  # the headline kernel at the headline size (100,000 alignments, level-4 table program)
      $NCU_FULL -k regex:prune_wide -s 3 -c 1 -f -o gpurun_out/r02_prune_wide_tabled_100k \
        python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_prune_sim.log 2>&1 ;;
    prune_real) # the same on data that does not look like the model's own draws
      $NCU_FULL -k regex:prune_wide -s 3 -c 1 -f -o gpurun_out/r02_prune_wide_tabled_100k_realistic \
        python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --data realistic > gpurun_out/r02_prune_real.log 2>&1 ;;
    k1)         # P(t) build, mle candidates (120mammals)
      $NCU_FULL -k regex:pt_build -s 6 -c 2 -f -o gpurun_out/r02_pt_build \
        python bench.py --only mle --mle-alignments 4000 > gpurun_out/r02_k1.log 2>&1 ;;
    k5)         # omega Q assembly + Jacobi
      $NCU_FULL -k regex:omega_eig -s 6 -c 2 -f -o gpurun_out/r02_omega_eig \
        python bench.py --only omega --omega-alignments 400 > gpurun_out/r02_k5.log 2>&1 ;;
    k0)         # pleaves on the device
      $NCU_FULL -k regex:frame_codes -s 1 -c 2 -f -o gpurun_out/r02_frame_codes \
        python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --alignments 40000 > gpurun_out/r02_k0.log 2>&1 ;;
    tables)     # subtree tables, levels 2-4
      $NCU_FULL -k regex:subtree_table -c 3 -f -o gpurun_out/r02_subtree_tables \
        python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_tables.log 2>&1 ;;
    outside)    # K6
      $NCU_FULL -k regex:outside_dmma -s 1 -c 1 -f -o gpurun_out/r02_outside_dmma \
        python tools/bench_posteriors.py 200000 > gpurun_out/r02_outside.log 2>&1 ;;
  esac
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
