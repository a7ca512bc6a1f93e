#!/bin/bash
# final verification on one GPU: the whole -m gpu suite, smoke(), the default bench line and the reference arm, launch list,
# ncu --set full of the headline kernel at the benchmark's own size
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=8 > gpurun_out/final_tests.log 2>&1
echo "tests rc=$? $(( $(date +%s) - T0 )) s" >> gpurun_out/final_tests.log
T1=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$? $(( $(date +%s) - T1 )) s" >> gpurun_out/final_smoke.log
T2=$(date +%s)
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench rc=$? $(( $(date +%s) - T2 )) s" > gpurun_out/final_times.log
T3=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
echo "reference rc=$? $(( $(date +%s) - T3 )) s" >> gpurun_out/final_times.log
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:prune|frame_codes|region_reduce|make_segments|pt_build|subtree_table' -c 200 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/final_launches_bench.log 2>&1
timeout 900 bash tools/profile_r02.sh prune_sim > gpurun_out/final_profile.log 2>&1
tail -14 gpurun_out/final_tests.log; cat gpurun_out/final_smoke.log | tail -3; cat gpurun_out/final_times.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_bench_n1.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k,v in d.get("extra",{}).items(): print(k, {kk:vv for kk,vv in v.items() if isinstance(vv,(int,float))})
PY
