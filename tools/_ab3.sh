#!/bin/bash
run() { env "$@" timeout 300 python bench.py --alignments 30000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value']/1e6,3), round(d['roofline']['frac'],4), round(d['e2e']['value']/1e6,3))"; }
run PCSF_CHERRY_TABLES=3
run PCSF_CHERRY_TABLES=0
run PCSF_CHERRY_TABLES=2
