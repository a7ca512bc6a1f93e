#!/bin/bash
# Kernel A/B on the GPU box: for every build/lib_*.so run a short device-resident bench and print value + roofline.frac.
# usage: tools/ab.sh [alignments] ; results in gpurun_out/ab_<name>.json
N=${1:-30000}
for lib in build/lib_*.so; do
  name=$(basename $lib .so)
  PCSF_LIB=$PWD/$lib python bench.py --alignments $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.3f M/s" % (d["value"] / 1e6), "frac %.4f" % d["roofline"]["frac"], "e2e %.3f" % (d["e2e"]["value"] / 1e6), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name", "FAILED", e)
PY
done
