#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_round2.py -k "headline or every_shipped or shared" -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_golden_and_scale.py -x -q 2>&1 | tail -3
} > gpurun_out/keep_tests.log 2>&1
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/keep_bench_on.json 2> gpurun_out/keep_bench_on.err
PCSF_NO_KEEP=1 timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/keep_bench_off.json 2> gpurun_out/keep_bench_off.err
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/keep_bench_on2.json 2> gpurun_out/keep_bench_on2.err
cat gpurun_out/keep_tests.log
for f in gpurun_out/keep_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ms_per_step"], d["roofline"]["frac"])
PY
done
