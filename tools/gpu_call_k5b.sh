#!/bin/bash
# one GPU call: K5 (unsorted pairs, hoisted lane values, short rotation chain) - omega parity, omega bench leg, ncu of K5
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_cli.py -m gpu -k "omega" -x -q > gpurun_out/k5b_tests.log 2>&1
timeout 900 python bench.py --only omega > gpurun_out/k5b_bench_omega.json 2> gpurun_out/k5b_bench_omega.err
timeout 600 bash tools/profile_r02.sh k5 > gpurun_out/k5b_profile.log 2>&1
tail -5 gpurun_out/k5b_tests.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/k5b_bench_omega.json").read().strip().splitlines()[-1])["extra"]["omega_cfg4"]
print(d["alignments_per_s"], d["seconds"], d["device_ms"], d["jacobi_sweeps_per_matrix"], d["median_score_db"], d["median_kappa_H0"])
PY
