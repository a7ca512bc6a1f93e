#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -k "pleaves or pipelined" -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_round2.py -k "k0_codes or headline" -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_golden_and_scale.py -x -q 2>&1 | tail -3
} > gpurun_out/pipe_tests.log 2>&1
for i in 1 2; do
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/pipe_bench_$i.json 2> gpurun_out/pipe_bench_$i.err
done
cat gpurun_out/pipe_tests.log
for f in gpurun_out/pipe_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ms_per_step"])
PY
done
