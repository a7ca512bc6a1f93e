#!/usr/bin/env python3
"""Per-kernel SASS evidence from the built library (no GPU needed): cuobjdump -sass, then counts of the mnemonics that
prove which hardware paths a kernel uses (B200_PROFILING.md): DMMA (FP64 tensor), UBLKCP (TMA bulk copy), SYNCS (mbarrier
transaction counts), LDGSTS (cp.async), USETMAXREG (setmaxnreg), UTCMMA / LDTM (tcgen05 - expected absent: no f64 kind).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "phylocsf_b200", "libphylocsf_b200.so")
WATCH = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "UTCMMA", "LDTM", "HMMA", "LDS", "STS", "LDG", "STG", "BAR", "MUFU", "CCTL", "FENCE", "MEMBAR"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    print("library: phylocsf_b200/libphylocsf_b200.so   arch:", ", ".join(arch))
    print("mnemonic counts per kernel (static instruction counts in the SASS, not executed counts)\n")
    cur, counts = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)(\.[A-Za-z0-9_.]+)?", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WATCH:
                if op == w or (w in ("UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "UTCMMA", "LDTM", "DMMA", "HMMA") and op.startswith(w)):
                    counts[cur][w] += 1
            if op == "DMMA":
                counts[cur]["DMMA" + (m.group(2) or "")] += 1
    for k, c in counts.items():
        print(k)
        print("   total %d | " % c["total"] + "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w]))
        shapes = [(n, v) for n, v in c.items() if n.startswith("DMMA.")]
        if shapes:
            print("   " + "  ".join("%s x%d" % nv for nv in shapes))
    print("\nUTCMMA / LDTM (tcgen05) are absent by design: tcgen05.mma has no .kind::f64 (ptxas rejects it), so the FP64 tensor path on sm_100a is warp-level DMMA.8x8x4.")


if __name__ == "__main__":
    main()
