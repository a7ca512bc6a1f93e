#!/bin/bash
# one GPU call: K0 parity + the paths around it, headline bench (e2e is where K0 shows), ncu of K0
set -u
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_round2.py -k "k0_codes" -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py -k "pleaves or pipelined or tal_AA" -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_cli.py -x -q -m gpu -k "not omega and not multi_device and not launcher" 2>&1 | tail -5
} > gpurun_out/k0_tests.log 2>&1
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/k0_bench_2m.json 2> gpurun_out/k0_bench_2m.err
timeout 600 bash tools/profile_r02.sh k0 > gpurun_out/k0_profile.log 2>&1
cat gpurun_out/k0_tests.log
for f in gpurun_out/k0_bench_2m.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
