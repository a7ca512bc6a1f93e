#!/bin/bash
# one GPU call: K5 with the short rotation chain - omega parity tests, omega bench leg, headline e2e with the small first chunk
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_cli.py -m gpu -k "omega" -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -k "pipelined" -x -q 2>&1 | tail -3
} > gpurun_out/k5_tests.log 2>&1
timeout 900 python bench.py --only omega > gpurun_out/k5_bench_omega.json 2> gpurun_out/k5_bench_omega.err
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/k5_bench_headline.json 2> gpurun_out/k5_bench_headline.err
cat gpurun_out/k5_tests.log
tail -c 1500 gpurun_out/k5_bench_omega.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/k5_bench_headline.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ms_per_step"])
PY
