#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`, no GPU needed) into the small JSON files kept
under profiles/: one object per captured launch with the metrics DESIGN.md cites.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x_ncu_summary.json [--columns N] [--workload "..."]
--columns N adds dram bytes per codon column (N = columns one launch processes)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
    "smsp__cycles_active.avg", "sm__cycles_active.avg", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__average_warp_latency_issue_stalled_barrier.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    columns, workload = None, None
    if "--columns" in sys.argv:
        columns = float(sys.argv[sys.argv.index("--columns") + 1])
    if "--workload" in sys.argv:
        workload = sys.argv[sys.argv.index("--workload") + 1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    d[k] = {"value": r[i], "unit": units[i]}
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            tot = rd["value"] * UNIT.get(rd["unit"], 1.0) + wr["value"] * UNIT.get(wr["unit"], 1.0)
            d["dram_bytes_total"] = tot
            t = d.get("gpu__time_duration.sum")
            if t:
                sec = t["value"] * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}.get(t["unit"], 1e-9)
                d["dram_gbytes_per_s"] = tot / sec / 1e9
            if columns:
                d["dram_bytes_per_codon_column"] = tot / columns
        launches.append(d)
    res = {"source": "ncu --set full --clock-control none (read with `ncu -i %s --page raw --csv`)" % rep, "launches": launches}
    if workload:
        res["workload"] = workload
    if columns and launches:
        res["codon_columns_per_launch"] = columns
        res["dram_bytes_per_codon_column"] = launches[-1].get("dram_bytes_per_codon_column")
    json.dump(res, open(out, "w"), indent=1)
    for d in launches:
        print(d["kernel"][:60], {k: v["value"] for k, v in d.items() if isinstance(v, dict) and k in KEEP[:9]})


if __name__ == "__main__":
    main()
