#!/usr/bin/env python3
"""One PhyloCSF process per GPU over one list of alignment files (what the reference's -p N / ForkWork.map_list did with
forked workers, src/ForkYes.ml:5-8; forked children cannot share a CUDA context, and one host process saturates at about
one GPU's worth of FASTA parsing):

    python tools/phylocsf_multi.py [--gpus N] <paramset> <list-of-files> [PhyloCSF options...]

The list is cut into N contiguous shards, shard k is scored by `bin/PhyloCSF --files` on GPU k (PCSF_DEVICE=k), and the
report lines come out on stdout in input order. An abort (return code 255) in shard k ends the output after shard k's
lines, like the single-process run would; the exit status is the first non-zero one."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "phylocsf_b200", "bin", "PhyloCSF")


def main():
    args = sys.argv[1:]
    gpus = 0
    if args and args[0] == "--gpus":
        gpus = int(args[1])
        args = args[2:]
    if len(args) < 2:
        sys.exit(__doc__)
    pset, lst, flags = args[0], args[1], [a for a in args[2:] if a != "--files"]
    if gpus <= 0:
        import ctypes
        gpus = max(1, ctypes.CDLL(os.path.join(ROOT, "phylocsf_b200", "libphylocsf_b200.so")).pcsf_device_count())
    files = [l for l in open(lst).read().split("\n") if l]
    n = max(1, min(gpus, len(files)))
    cuts = [len(files) * k // n for k in range(n + 1)]
    d = tempfile.mkdtemp(prefix="pcsf_multi_")
    procs = []
    for k in range(n):
        part = os.path.join(d, "list%d.txt" % k)
        open(part, "w").write("\n".join(files[cuts[k]:cuts[k + 1]]) + "\n")
        out = open(os.path.join(d, "out%d.txt" % k), "wb")
        env = dict(os.environ, PCSF_DEVICE=str(k))
        env.pop("PCSF_DEVICES", None)
        env.setdefault("PCSF_HOST_THREADS", str(max(2, (os.cpu_count() or 8) // n)))
        procs.append((subprocess.Popen([CLI, pset, part, "--files"] + flags, stdout=out, env=env), out))
    status = 0
    stop = False
    for k, (p, out) in enumerate(procs):
        rc = p.wait()
        out.close()
        if not stop:
            with open(os.path.join(d, "out%d.txt" % k), "rb") as f:
                sys.stdout.buffer.write(f.read())
        if rc != 0 and status == 0:
            status = rc
            stop = True
    sys.stdout.flush()
    sys.exit(status)


if __name__ == "__main__":
    main()
