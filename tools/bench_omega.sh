#!/bin/bash
# BASELINE.json configs[3] in small: 100vertebrates, --strategy=omega --allScores --frames=3, N simulated exon-length
# alignments; prints CLI throughput and (with NCU=1) the per-kernel device time split.
N=${1:-300}
if [ -n "$NCU" ]; then
  PCSF_NCU_WRAP="ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/omega_launches.csv" python tools/bench_cli.py 100vertebrates $N 60 -- --strategy=omega --frames=3 --allScores | cut -c1-300
  python - <<'PY'
import csv, collections
t = collections.Counter(); n = collections.Counter()
for row in csv.reader(open("gpurun_out/omega_launches.csv")):
    if len(row) > 14 and row[12] == "gpu__time_duration.sum":
        k = row[4].split("(")[0].split("::")[-1]
        t[k] += float(row[14]) / 1e6; n[k] += 1
for k, v in t.most_common(): print("%-28s %9.1f ms %7d launches" % (k, v, n[k]))
PY
else
  python tools/bench_cli.py 100vertebrates $N 60 -- --strategy=omega --frames=3 --allScores | cut -c1-300
fi
