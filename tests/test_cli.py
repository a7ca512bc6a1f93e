"""The drop-in command line (phylocsf_b200/bin/PhyloCSF, C++ host over the C ABI) against the oracle's
restatement of src/PhyloCSF.ml:process_alignment, line for line. --strategy=nop exercises the whole
host pipeline (MFA reader, species pruning, ORF search, regions, report format, error tiers) without a
GPU; the gpu-marked tests run fixed / mle / omega and the reference's own golden checks (src/test.ml)."""
import os
import subprocess

import pytest

from oracle import oracle as o
from tools import golden_params as gp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "phylocsf_b200", "bin", "PhyloCSF")


def run_cli(params_base, pset, files, *flags, stdin=None, expect_rc=0):
    env = dict(os.environ, PHYLOCSF_BASE=params_base)
    r = subprocess.run([CLI, pset] + list(files) + list(flags), env=env, input=stdin, capture_output=True, text=True, timeout=600)
    assert r.returncode == expect_rc, (r.returncode, r.stdout, r.stderr)
    return r.stdout.splitlines()


def run_oracle(params_base, pset, name, lines, **kw):
    opts = o.Options(**kw)
    ps = o.load_paramset(os.path.join(params_base, "PhyloCSF_Parameters", pset), opts)
    return o.process_alignment(ps, opts, name, lines)


def same_lines(a, b, tol=1.5e-4):
    """Equal up to the last printed digit of floating-point fields."""
    assert len(a) == len(b), (a, b)
    for la, lb in zip(a, b):
        fa, fb = la.split("\t"), lb.split("\t")
        assert len(fa) == len(fb), (la, lb)
        for x, y in zip(fa, fb):
            if x == y:
                continue
            try:
                assert abs(float(x) - float(y)) <= tol, (la, lb)
            except ValueError:
                raise AssertionError((la, lb))


def ex(params_base, fn):
    return os.path.join(params_base, "PhyloCSF_Examples", fn)


NOP_CASES = [
    ("12flies", "tal-AA.fa", [], {}),
    ("12flies", "tal-AA.fa", ["--frames=3", "--allScores", "--dna", "--aa", "--bls"], dict(frames=3, all_scores=True, dna=True, aa=True, bls=True)),
    ("29mammals", "ALDH2.exon5.fa", ["-f", "6", "--allScores", "--bls", "--ancComp"], dict(frames=6, all_scores=True, bls=True, anc_comp=True)),
    ("29mammals", "Aldh2.mRNA.fa", ["--orf=ATGStop", "--frames=3", "--removeRefGaps", "--aa", "--allScores"],
     dict(orf="ATGStop", frames=3, remove_ref_gaps=True, aa=True, all_scores=True)),
    ("29mammals", "Aldh2.mRNA.fa", ["--orf=StopStop", "--frames=6", "--removeRefGaps", "--allScores", "--minCodons=40", "--bls"],
     dict(orf="StopStop", frames=6, remove_ref_gaps=True, all_scores=True, min_codons=40, bls=True)),
    ("29mammals", "Aldh2.mRNA.fa", ["--orf=StopStop3", "--frames=3", "--removeRefGaps", "--allScores"],
     dict(orf="StopStop3", frames=3, remove_ref_gaps=True, all_scores=True)),
    ("29mammals", "Aldh2.mRNA.fa", ["--orf=ToFirstStop", "--frames=3", "--removeRefGaps", "--allScores", "--minCodons=1"],
     dict(orf="ToFirstStop", frames=3, remove_ref_gaps=True, all_scores=True, min_codons=1)),
    ("29mammals", "Aldh2.mRNA.fa", ["--orf=FromLastStop", "--frames=6", "--removeRefGaps", "--allScores", "--minCodons=1"],
     dict(orf="FromLastStop", frames=6, remove_ref_gaps=True, all_scores=True, min_codons=1)),
    ("29mammals", "Aldh2.mRNA.fa", ["--orf=ToOrFromStop", "--frames=6", "--removeRefGaps", "--allScores", "--minCodons=1"],
     dict(orf="ToOrFromStop", frames=6, remove_ref_gaps=True, all_scores=True, min_codons=1)),
    ("29mammals", "ALDH2.exon5.fa", ["--species=Human,Mouse,Rat,Dog,Cow,Horse", "--bls", "--frames=3"],
     dict(species=["Human", "Mouse", "Rat", "Dog", "Cow", "Horse"], bls=True, frames=3)),
    # AsIs regions of an alignment with a gapped reference: the fast reader compacts the rows itself
    ("29mammals", "Aldh2.mRNA.fa", ["--removeRefGaps", "--frames=6", "--bls", "--allScores", "--aa"],
     dict(remove_ref_gaps=True, frames=6, bls=True, all_scores=True, aa=True)),
]


@pytest.mark.parametrize("pset,fn,flags,kw", NOP_CASES)
def test_nop_pipeline_matches_oracle(params_base, pset, fn, flags, kw):
    path = ex(params_base, fn)
    got = run_cli(params_base, pset, [path], "--strategy=nop", *flags)
    want = run_oracle(params_base, pset, path, gp.example_lines(fn), strategy="nop", **kw)
    assert got == want


def test_species_pruned_alignment_with_foreign_species_aborts(params_base):
    path = ex(params_base, "ALDH2.exon5.fa")
    got = run_cli(params_base, "29mammals", [path], "--strategy=nop", "--species=Human,Mouse", expect_rc=255)
    want = run_oracle(params_base, "29mammals", path, gp.example_lines("ALDH2.exon5.fa"), strategy="nop", species=["Human", "Mouse"])
    assert got == want and "\tabort\t" in got[0] and "parameters not available for species" in got[0]


def test_error_tiers(params_base, tmp_path):
    def write(name, text):
        p = tmp_path / name
        p.write_text(text)
        return str(p)

    cases = [
        write("gapref.fa", ">dmel\nATG---GGG\n>dana\nATGCCCGGG\n"),          # reference gapped -> abort
        write("badchar.fa", ">dmel\nATGRCCGGG\n>dana\nATGCCCGGG\n"),        # not ACGTNacgtn- -> abort (revcomp)
        write("ragged.fa", ">dmel\nATGCCCGGG\n>dana\nATGCCC\n"),            # length mismatch -> abort
        write("nohdr.fa", "ATGCCCGGG\n"),                                   # bad header -> abort
        write("alien.fa", ">dmel\nATGCCCGGG\n>human\nATGCCCGGG\n"),         # species not in tree -> abort
    ]
    for path in cases:
        got = run_cli(params_base, "12flies", [path], "--strategy=nop", expect_rc=255)
        want = run_oracle(params_base, "12flies", path, open(path).read().split("\n")[:-1], strategy="nop")
        assert got == want, (got, want)
        assert got[0].split("\t")[1] == "abort"
    # no ORFs -> per-alignment failure, exit code 0, processing continues with the next file
    short = write("short.fa", ">dmel\nCCCGGGAAATAA\n>dana\nCCCGGGAAATAA\n")
    ok = write("ok.fa", ">dmel\nATGCCCGGGAAACCC\n>dana\nATGCCCGGGAAACCC\n")
    got = run_cli(params_base, "12flies", [short, ok], "--strategy=nop", "--orf=ATGStop", "--frames=3", "--minCodons=2")
    assert got[0] == short + '\tfailure\tFailure("no sufficiently long ORFs found")'
    assert got[1].startswith(ok + "\tfailure") or got[1].startswith(ok + "\tmax_score")
    # a missing file aborts after the earlier alignments were reported
    got = run_cli(params_base, "12flies", [ok, str(tmp_path / "missing.fa")], "--strategy=nop", expect_rc=255)
    assert got[0].startswith(ok + "\tscore(decibans)\t0.0000") and "\tabort\tSys_error" in got[1]
    # u -> t, lower case, stdin, --files
    got = run_cli(params_base, "12flies", [], "--strategy=nop", "--dna", stdin=">dmel\nauggcc\n>dana\nAUGGCC\n")
    assert got == ["(STDIN)\tscore(decibans)\t0.0000\tatggcc"]
    lst = write("list.txt", ok + "\n" + ok + "\n")
    got = run_cli(params_base, "12flies", [lst], "--strategy=nop", "--files")
    assert len(got) == 2 and all(l.startswith(ok) for l in got)


def test_usage_and_missing_base(params_base):
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 255 and "usage" in r.stderr
    env = {k: v for k, v in os.environ.items() if k != "PHYLOCSF_BASE"}
    r = subprocess.run([CLI, "12flies", "--strategy=nop"], env=env, input="", capture_output=True, text=True)
    assert r.returncode == 2 and "PHYLOCSF_BASE" in r.stderr
    # a path-like parameter set needs no PHYLOCSF_BASE (src/PhyloCSF.ml:409-419)
    prefix = os.path.join(params_base, "PhyloCSF_Parameters", "12flies")
    r = subprocess.run([CLI, prefix, "--strategy=nop"], env=env, input=">dmel\nATGGCC\n>dana\nATGGCC\n", capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "(STDIN)\tscore(decibans)\t0.0000\n"


def _variants(params_base, tmp_path):
    """tal-AA in the shapes the fast reader takes itself (wrapped lines, CRLF, RNA letters, lower case, a header
    with a |comment) and in shapes it hands to the general reader (a blank line, a padded line)."""
    lines = gp.example_lines("tal-AA.fa")
    recs, cur = [], None
    for l in lines:
        if l.startswith(">"):
            cur = [l[1:].strip(), ""]
            recs.append(cur)
        elif cur is not None:
            cur[1] += l.strip()

    def write(name, text):
        q = tmp_path / name
        q.write_bytes(text.encode())
        return str(q)

    wrap = lambda s, n: "\n".join(s[i:i + n] for i in range(0, len(s), n))
    plain = "".join(">%s\n%s\n" % (h, s) for h, s in recs)
    files = {
        "plain": write("plain.fa", plain),
        "wrapped": write("wrapped.fa", "".join(">%s\n%s\n" % (h, wrap(s, 17)) for h, s in recs)),
        "crlf": write("crlf.fa", plain.replace("\n", "\r\n")),
        "rna_lower": write("rna.fa", "".join(">%s | some comment\n%s\n" % (h, s.replace("T", "u").replace("A", "a")) for h, s in recs)),
        "no_final_newline": write("nonl.fa", plain[:-1]),
        "blank_line": write("blank.fa", plain.replace("\n>", "\n\n>", 1)),
        "padded": write("padded.fa", "".join(">%s\n  %s \n" % (h, s) for h, s in recs)),
    }
    return files, plain.split("\n")[:-1]


def test_reader_variants_nop(params_base, tmp_path):
    """Every shape is accepted and reported in input order; a shape the fast reader declines in the middle of
    the list starts a new batch in the other form without reordering the output."""
    files, _ = _variants(params_base, tmp_path)
    order = ["plain", "wrapped", "blank_line", "crlf", "padded", "rna_lower", "no_final_newline"]
    got = run_cli(params_base, "12flies", [files[k] for k in order], "--strategy=nop", "--frames=6")
    assert [l.split("\t")[0] for l in got] == [files[k] for k in order]
    assert all(l.split("\t")[1:] == got[0].split("\t")[1:] for l in got)


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["fixed", "mle"])
def test_reader_variants_score_like_the_oracle(params_base, tmp_path, strategy):
    """The fast reader (nucleotide rows staged as they are read, pleaves on the device) and the general reader
    (host pleaves) give the oracle's lines for the same alignment in every shape, mixed in one run."""
    files, lines = _variants(params_base, tmp_path)
    order = ["plain", "wrapped", "blank_line", "crlf", "padded", "rna_lower", "no_final_newline"]
    flags = ["--strategy=" + strategy, "--frames=6", "--allScores", "--ancComp", "--dna", "--aa"]
    got = run_cli(params_base, "12flies", [files[k] for k in order], *flags)
    per = len(got) // len(order)
    assert per * len(order) == len(got) and per == 7
    for i, k in enumerate(order):
        src = lines
        if k == "rna_lower":  # the oracle sees the same text the file holds
            src = open(files[k]).read().split("\n")[:-1]
        want = run_oracle(params_base, "12flies", files[k], src, strategy=strategy, frames=6, all_scores=True, anc_comp=True, dna=True, aa=True)
        same_lines(got[i * per:(i + 1) * per], want)


@pytest.mark.gpu
def test_fast_reader_missing_species_and_gaps(params_base, tmp_path):
    """Species absent from the alignment marginalise; gaps and N in non-reference rows marginalise their codons."""
    import numpy as np

    ops = o.load_paramset(os.path.join(params_base, "PhyloCSF_Parameters", "29mammals"), o.Options(strategy="fixed"))
    labels = ops.tree.labels[: ops.tree.n_leaves]
    rng = np.random.default_rng(8)
    rows = o.codes_to_alignment(o.simulate_columns(ops.model.coding_model.model(1.0), 50, rng))
    keep = [0, 3, 4, 9, 17, 28, 11]  # reference first, not in tree order
    text = ""
    for n, i in enumerate(keep):
        r = rows[i]
        if n == 2:
            r = r[:30] + "---" + r[33:60] + "NN" + r[62:]
        if n == 4:
            r = "-" * 40 + r[40:]
        text += ">%s\n%s\n" % (labels[i], r)
    path = tmp_path / "subset.fa"
    path.write_text(text)
    flags = ["--strategy=fixed", "--frames=6", "--allScores", "--ancComp"]
    got = run_cli(params_base, "29mammals", [str(path)] * 2, *flags)
    want = run_oracle(params_base, "29mammals", str(path), text.split("\n")[:-1], strategy="fixed", frames=6, all_scores=True, anc_comp=True)
    same_lines(got, want * 2)



@pytest.mark.gpu
def test_reference_goldens_through_the_cli(params_base):
    """src/test.ml:27-59, the reference's own end-to-end tests (default strategy mle)."""
    ans = run_cli(params_base, "12flies", [ex(params_base, "tal-AA.fa")], "--ancComp")[0].split("\t")
    assert ans[1] == "score(decibans)" and 297.62 < float(ans[2]) < 297.63 and 48.25 < float(ans[3]) < 48.26
    ans = run_cli(params_base, "29mammals", [ex(params_base, "ALDH2.exon5.fa")], "--ancComp")[0].split("\t")
    assert ans[1] == "score(decibans)" and -178.93 < float(ans[2]) < -178.92 and -38.29 < float(ans[3]) < -38.28
    ans = run_cli(params_base, "29mammals", [ex(params_base, "ALDH2.exon5.fa")], "--frames=6", "-p", "8")[0].split("\t")
    assert ans[1] == "max_score(decibans)" and 218.26 < float(ans[2]) < 218.27 and ans[3:6] == ["1", "111", "+"]
    ans = run_cli(params_base, "29mammals", [ex(params_base, "Aldh2.mRNA.fa")], "--orf=ATGStop", "--frames=3", "--removeRefGaps", "--aa")[0].split("\t")
    assert ans[1] == "max_score(decibans)" and 2013.92 < float(ans[2]) < 2013.93 and ans[3:5] == ["343", "1899"]
    assert ans[5].startswith("MLRAALTTVRRGPRLSRLLSAAA")


@pytest.mark.gpu
@pytest.mark.parametrize("pset,fn,flags,kw", [
    ("12flies", "tal-AA.fa", ["--strategy=fixed", "--ancComp", "--debug"], dict(strategy="fixed", anc_comp=True, debug=True)),
    ("29mammals", "ALDH2.exon5.fa", ["--strategy=fixed", "--frames=6", "--allScores", "--ancComp", "--bls"],
     dict(strategy="fixed", frames=6, all_scores=True, anc_comp=True, bls=True)),
    ("29mammals", "ALDH2.exon5.fa", ["--strategy=mle", "--frames=3", "--allScores", "--ancComp", "--debug"],
     dict(strategy="mle", frames=3, all_scores=True, anc_comp=True, debug=True)),
    ("29mammals", "Aldh2.mRNA.fa", ["--strategy=fixed", "--orf=ATGStop", "--frames=3", "--removeRefGaps", "--allScores"],
     dict(strategy="fixed", orf="ATGStop", frames=3, remove_ref_gaps=True, all_scores=True)),
    ("12flies", "tal-AA.fa", ["--strategy=omega", "--frames=3", "--allScores", "--debug"], dict(strategy="omega", frames=3, all_scores=True, debug=True)),
    ("29mammals", "Aldh2.mRNA.fa", ["--strategy=fixed", "--removeRefGaps", "--frames=6", "--bls", "--allScores", "--ancComp"],
     dict(strategy="fixed", remove_ref_gaps=True, frames=6, bls=True, all_scores=True, anc_comp=True)),
    # fixed + ORF search on both strands: scored from per-column terms of whole frames (frame mode)
    ("29mammals", "Aldh2.mRNA.fa", ["--strategy=fixed", "--orf=StopStop3", "--frames=6", "--removeRefGaps", "--allScores", "--ancComp", "--debug", "--minCodons=30"],
     dict(strategy="fixed", orf="StopStop3", frames=6, remove_ref_gaps=True, all_scores=True, anc_comp=True, debug=True, min_codons=30)),
])
def test_scores_match_oracle_lines(params_base, pset, fn, flags, kw):
    path = ex(params_base, fn)
    got = run_cli(params_base, pset, [path], *flags)
    want = run_oracle(params_base, pset, path, gp.example_lines(fn), **kw)
    same_lines(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["2", "3"])
def test_fixed_with_subtree_tables_forced(params_base, monkeypatch, mode):
    """The command line with the memoised subtree tables forced on (they are built only for big batches otherwise) and
    the wide form of the pruning kernel: the same lines as the oracle, 6 frames, ancestral composition included."""
    monkeypatch.setenv("PCSF_CHERRY_TABLES", mode)
    monkeypatch.setenv("PCSF_WIDE", "1")
    path = ex(params_base, "ALDH2.exon5.fa")
    flags = ["--strategy=fixed", "--frames=6", "--allScores", "--ancComp"]
    got = run_cli(params_base, "29mammals", [path] * 2, *flags)
    want = run_oracle(params_base, "29mammals", path, gp.example_lines("ALDH2.exon5.fa"), strategy="fixed", frames=6, all_scores=True, anc_comp=True)
    same_lines(got, want * 2)


@pytest.mark.gpu
def test_many_alignments_one_batch_keep_order(params_base, tmp_path):
    """Several alignments are staged as one GPU batch; output order and values equal one-by-one runs."""
    files = [ex(params_base, "ALDH2.exon5.fa")] * 3
    got = run_cli(params_base, "29mammals", files, "--strategy=fixed", "--frames=3")
    one = run_cli(params_base, "29mammals", files[:1], "--strategy=fixed", "--frames=3")
    assert got == one * 3


@pytest.mark.gpu
def test_omega_100vertebrates_species_subset(params_base, tmp_path):
    """BASELINE.json configs[3] in miniature: omega strategy, --allScores, 3 frames, on a simulated
    exon under the 100vertebrates tree pruned with --species (keeps the CPU oracle affordable)."""
    import numpy as np

    sp = ["Human", "Mouse", "Dog", "Cow", "Elephant", "Opossum", "Chicken", "Lizard", "Zebrafish", "Lamprey"]
    ops = o.load_paramset(os.path.join(params_base, "PhyloCSF_Parameters", "100vertebrates"),
                                                    o.Options(strategy="fixed", species=sp))
    labels = ops.tree.labels[: ops.tree.n_leaves]
    assert len(labels) >= 6, labels
    rng = np.random.default_rng(3)
    codes = o.simulate_columns(ops.model.coding_model.model(1.0), 40, rng)
    rows = o.codes_to_alignment(codes)
    path = tmp_path / "sim.fa"
    path.write_text("".join(">%s\n%s\n" % (l, r) for l, r in zip(labels, rows)))
    flags = ["--strategy=omega", "--frames=3", "--allScores", "--species=" + ",".join(sp)]
    got = run_cli(params_base, "100vertebrates", [str(path)], *flags)
    want = run_oracle(params_base, "100vertebrates", str(path), path.read_text().split("\n")[:-1], strategy="omega", frames=3,
                      all_scores=True, species=sp)
    same_lines(got, want)


@pytest.mark.gpu
def test_multi_device_dispatch_keeps_order(params_base, monkeypatch):
    """Batches go round-robin over the configured devices and are scored asynchronously; the output must
    be identical to a single-context run. PCSF_DEVICES=0,0 exercises the dispatcher on a one-GPU box (two
    contexts on the same device); with more GPUs visible, 'all' is checked as well."""
    import ctypes

    files = [ex(params_base, "ALDH2.exon5.fa")] * 9
    flags = ["--strategy=mle", "--frames=3", "--allScores", "--ancComp"]
    one = run_cli(params_base, "29mammals", files, *flags)
    monkeypatch.setenv("PCSF_BATCH_COLS", "150")  # ~1 alignment per batch
    monkeypatch.setenv("PCSF_DEVICES", "0,0")
    two = run_cli(params_base, "29mammals", files, *flags)
    assert two == one
    from phylocsf_b200 import _native as N

    if N.load().pcsf_device_count() >= 2:
        monkeypatch.setenv("PCSF_DEVICES", "all")
        assert run_cli(params_base, "29mammals", files, *flags) == one


def test_fast_reader_equals_general_reader_on_random_alignments(params_base, tmp_path):
    """Differential test without a GPU (--strategy=nop): random alignments in random shapes (wrapped lines, CRLF, lower
    case, u for t, N and gaps, missing and shuffled species, |comments, a gapped reference, now and then a foreign
    character or a ragged row) through the fast reader and, with PCSF_NO_FAST_READER, through the general reader: the
    same lines for every option set, including --bls, --removeRefGaps, --dna, --aa and the abort messages."""
    import random

    rnd = random.Random(5)
    species = ["dmel", "dsim", "dsec", "dyak", "dere", "dana", "dpse", "dper", "dwil", "dvir", "dmoj", "dgri"]
    files = []
    for k in range(60):
        L = rnd.choice([3, 4, 30, 31, 32, 90, 200])
        present = rnd.sample(species, rnd.randint(2, len(species)))
        ref = "".join(rnd.choice("ACGT") for _ in range(L))
        rows = []
        for i, sp in enumerate(present):
            row = list(ref)
            for j in range(L):
                r = rnd.random()
                if r < 0.08:
                    row[j] = rnd.choice("ACGT")
                elif r < 0.12 and (i > 0 or k % 7 == 3):  # k % 7 == 3: a gapped reference now and then
                    row[j] = rnd.choice("-N")
            row = "".join(row)
            if k % 5 == 1:
                row = row.lower()
            if k % 6 == 2:
                row = row.replace("T", "U").replace("t", "u")
            rows.append(row)
        if k % 19 == 7:
            rows[-1] = rows[-1][:-1] + "R"      # foreign character -> abort
        if k % 23 == 11 and L > 3:
            rows[-1] = rows[-1][:-1]            # ragged -> abort
        text = ""
        for sp, row in zip(present, rows):
            hdr = ">" + sp + (" | chr2L:%d" % k if k % 4 == 0 else "")
            body = row if k % 3 else "\n".join(row[i:i + 13] for i in range(0, len(row), 13))
            text += hdr + "\n" + body + "\n"
        if k % 8 == 5:
            text = text.replace("\n", "\r\n")
        p = tmp_path / ("r%02d.fa" % k)
        p.write_bytes(text.encode())
        files.append(str(p))
    env_general = dict(os.environ, PHYLOCSF_BASE=params_base, PCSF_NO_FAST_READER="1")
    env_fast = dict(os.environ, PHYLOCSF_BASE=params_base, PCSF_HOST_PROFILE="1")
    took_fast = 0
    for flags in (["--frames=6", "--bls", "--dna", "--aa"], ["--frames=3", "--removeRefGaps", "--bls"], ["--allowRefGaps", "--bls", "--allScores", "--frames=3"], []):
        for f in files:  # one alignment per run: an abort ends a run
            cmd = [CLI, "12flies", f, "--strategy=nop"] + flags
            a = subprocess.run(cmd, env=env_fast, capture_output=True, text=True, timeout=60)
            b = subprocess.run(cmd, env=env_general, capture_output=True, text=True, timeout=60)
            assert (a.returncode, a.stdout) == (b.returncode, b.stdout), (f, flags, a.stdout, b.stdout)
            took_fast += "fast reader took 1 alignments" in a.stderr
    assert took_fast > 100  # the fast reader did take most of the well-formed ones


def test_one_process_per_gpu_launcher_keeps_order(params_base, tmp_path):
    """tools/phylocsf_multi.py (the list cut into contiguous shards, one command-line process per GPU, outputs
    concatenated in input order) prints exactly what one process prints - checked without a GPU under --strategy=nop."""
    import sys

    files = [ex(params_base, "ALDH2.exon5.fa"), ex(params_base, "Aldh2.mRNA.fa")] * 5 + [ex(params_base, "ALDH2.exon5.fa")]
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(files) + "\n")
    flags = ["--strategy=nop", "--frames=3", "--removeRefGaps", "--bls"]
    one = run_cli(params_base, "29mammals", [str(lst)], "--files", *flags)
    env = dict(os.environ, PHYLOCSF_BASE=params_base)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "phylocsf_multi.py"), "--gpus", "3", "29mammals", str(lst)] + flags,
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == one and len(one) == len(files)
