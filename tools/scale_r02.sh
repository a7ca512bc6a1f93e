#!/bin/bash
# Round-2 multi-GPU measurements on one box with N GPUs (run under `gpurun --gpus N`):
#   bash tools/scale_r02.sh N
# 1. bench.py under torchrun: headline (weak scaling, replicas) + extra.cfg5_strong_scaling (10 M columns split over N ranks)
# 2. the command line's own one-process dispatcher (PCSF_DEVICES=all replaces ForkWork.map_list, src/ForkYes.ml:5-8):
#    58mammals fixed, 3 frames, 1,000,000 alignments x 100 codons (10,000 files listed 100 times) from the page cache; and config 5 (6 frames, 5,001 nt)
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    > gpurun_out/r02_scale_bench_n$N.json 2> gpurun_out/r02_scale_bench_n$N.err
else
  python bench.py --no-cpu-baseline > gpurun_out/r02_scale_bench_n$N.json 2> gpurun_out/r02_scale_bench_n$N.err
fi
tail -c 400 gpurun_out/r02_scale_bench_n$N.err
export PCSF_DEVICES=all PCSF_HOST_PROFILE=1
python tools/bench_cli.py 58mammals 10000 100 100 -- --strategy=fixed --frames=3 > gpurun_out/r02_scale_cli_n$N.json 2> gpurun_out/r02_scale_cli_n$N.err
python tools/bench_cli.py 58mammals 250 1667 4 -- --strategy=fixed --frames=6 > gpurun_out/r02_scale_cli_cfg5_n$N.json 2> gpurun_out/r02_scale_cli_cfg5_n$N.err
cut -c1-700 gpurun_out/r02_scale_cli_n$N.json gpurun_out/r02_scale_cli_cfg5_n$N.json
# 3. one process per GPU over the same list (tools/phylocsf_multi.py): the multi-GPU form of the command line that scales
if [ "$N" -gt 1 ]; then
  unset PCSF_DEVICES
  PCSF_MULTI=$N python tools/bench_cli.py 58mammals 10000 100 100 -- --strategy=fixed --frames=3 > gpurun_out/r02_scale_cli_multi_n$N.json 2> gpurun_out/r02_scale_cli_multi_n$N.err
  cut -c1-400 gpurun_out/r02_scale_cli_multi_n$N.json
fi
