"""K0 (pleaves on the device, phylocsf_b200/csrc/pcsf_k0.cuh) without a GPU: the kernel's own indexing functions run
thread by thread in a CPU emulation (tests/k0_emul.cpp, built here with g++) against the oracle's pleaves
(src/PhyloCSF.ml:219-246, Code.ml:39-51) on ragged inputs, for every frame count, tile size and grid shape.
The same comparison runs on the device in tests/test_gpu_round2.py::test_k0_codes_bit_exact_against_oracle_pleaves."""
import ctypes
import os
import subprocess
import types

import numpy as np
import pytest

from oracle import oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("k0") / "k0_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "phylocsf_b200", "csrc"),
                           "-o", so, os.path.join(ROOT, "tests", "k0_emul.cpp")])
    L = ctypes.CDLL(so)
    P = ctypes.c_void_p
    L.k0_emulate.argtypes = [P, ctypes.c_int64, P, P, P, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
    L.k0_emulate.restype = ctypes.c_int
    for f, n in (("k0_decode4", 1), ("k0_codon4", 3), ("k0_revcomp4", 1)):
        getattr(L, f).argtypes = [ctypes.c_uint32] * n
        getattr(L, f).restype = ctypes.c_uint32
    L.k0_choose_tile_pos.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.k0_choose_tile_pos.restype = ctypes.c_int
    return L


def oracle_frames(alns, frames):
    """codes [total_cols][n] and region offsets, regions alignment-major then frame (+0 +1 +2 -0 -1 -2)."""
    out, roff = [], [0]
    for rows in alns:
        n = len(rows)
        fake = types.SimpleNamespace(n_leaves=n)
        for f in range(frames):
            src = [o.revcomp(r) for r in rows] if f >= 3 else rows
            c = o.pleaves(fake, list(range(n)), src, lo=f % 3)
            out.append(c)
            roff.append(roff[-1] + c.shape[0])
    n = len(alns[0]) if alns else 1
    return (np.concatenate(out, axis=0) if out else np.zeros((0, n), np.uint8)), np.array(roff, dtype=np.int64)


def pack(alns, rng, max_pad=7):
    """nucleotide buffer with random padding between alignments (rows start at any byte phase)"""
    parts, offs, at = [], [], 0
    for rows in alns:
        pad = int(rng.integers(0, max_pad + 1))
        parts.append(np.frombuffer(bytes(rng.integers(33, 127, size=pad, dtype=np.uint8)), dtype=np.uint8))
        at += pad
        offs.append(at)
        flat = np.frombuffer("".join(rows).encode(), dtype=np.uint8)
        parts.append(flat)
        at += flat.size
    return (np.concatenate(parts) if parts else np.zeros(0, np.uint8)), np.array(offs, dtype=np.int64)


def run_emul(L, nt, off, lens, roff, frames, n, tile_pos, grid_y):
    total = int(roff[-1])
    codes = np.full(total * n + 64, 0xFF, dtype=np.uint8)
    # the kernel needs a 16-byte aligned output and a 4-byte aligned input; numpy allocations are
    nt = np.ascontiguousarray(nt)
    lens = np.asarray(lens, dtype=np.int32)
    assert codes.ctypes.data % 16 == 0
    rc = L.k0_emulate(nt.ctypes.data, nt.size, off.ctypes.data, lens.ctypes.data, roff.ctypes.data, len(lens), frames, n,
                      tile_pos, grid_y, codes.ctypes.data)
    assert rc == 0
    assert (codes[total * n:] == 0xFF).all()  # nothing written past the batch
    return codes[:total * n].reshape(total, n)


def test_character_decode_is_exhaustive(emul):
    """every byte value: index 0..3 for ACGTacgt, bit 2 set otherwise (what nt_index / Code.ml's tables say)"""
    want = np.full(256, 4, dtype=np.int64)
    for ch, i in o._DNA_INDEX.items():
        want[ord(ch)] = i
    for c in range(256):
        for lane in range(4):
            w = (0x41414141 & ~(0xFF << (8 * lane))) | (c << (8 * lane))
            got = (emul.k0_decode4(w) >> (8 * lane)) & 0xFF
            assert (got if got < 4 else 4) == want[c], (c, lane, got)
            others = emul.k0_decode4(w) & ~(0xFF << (8 * lane)) & 0xFFFFFFFF
            assert others == 0, (c, lane)  # neighbours are 'A' = 0 whatever this byte is


def test_codon_and_revcomp_words(emul):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        idx = rng.integers(0, 8, size=(3, 4))  # 0..3 valid, 4..7 = flagged
        words = [int(sum(int(idx[k, b]) << (8 * b) for b in range(4))) for k in range(3)]
        got = emul.k0_codon4(*words)
        for b in range(4):
            i1, i2, i3 = (int(x) for x in idx[:, b])
            want = 64 if max(i1, i2, i3) >= 4 else 16 * i1 + 4 * i2 + i3
            assert (got >> (8 * b)) & 0xFF == want
    for code in range(65):
        want = 64 if code == 64 else o.codon_code(*o.revcomp(o.codon_of_index(code)))
        for lane in range(4):
            assert (emul.k0_revcomp4(code << (8 * lane)) >> (8 * lane)) & 0xFF == want


@pytest.mark.parametrize("frames", [1, 3, 6])
@pytest.mark.parametrize("n", [1, 3, 58])
def test_emulated_kernel_matches_oracle_pleaves_on_ragged_batches(emul, frames, n):
    rng = np.random.default_rng(100 * frames + n)
    alphabet = np.array(list("ACGTacgtNn-"))
    lens = [0, 1, 2, 3, 4, 5, 15, 16, 17, 18, 31, 47, 48, 49, 50, 95, 96, 97, 300, 301, 302, 333]
    alns = [["".join(alphabet[rng.integers(0, len(alphabet), size=L)]) for _ in range(n)] for L in lens]
    want, roff = oracle_frames(alns, frames)
    nt, off = pack(alns, rng)
    for tile_pos, grid_y in ((16, 1), (48, 1), (48, 3), (64, 2), (emul.k0_choose_tile_pos(max(lens), n, frames), 1), (352, 2)):
        got = run_emul(emul, nt, off, lens, roff, frames, n, tile_pos, grid_y)
        assert np.array_equal(got, want), (tile_pos, grid_y)


def test_emulated_kernel_any_byte_and_wide_trees(emul):
    """bytes outside the alphabet (the reference would have rejected the file; the device must still say `Marginalize`),
    a 120-leaf tree, one long alignment over several tiles"""
    rng = np.random.default_rng(7)
    n, lens = 120, [5001, 7, 64]
    raw = [rng.integers(0, 256, size=(n, L), dtype=np.uint8) for L in lens]
    for r in raw:  # half of the bytes real nucleotides
        m = rng.random(r.shape) < 0.5
        r[m] = np.frombuffer(b"ACGTacgt", dtype=np.uint8)[rng.integers(0, 8, size=int(m.sum()))]
    idx = np.full(256, -1, dtype=np.int64)
    for ch, i in o._DNA_INDEX.items():
        idx[ord(ch)] = i
    want, roff = [], [0]
    for r in raw:
        L = r.shape[1]
        for f in range(6):
            i = idx[r]
            if f >= 3:
                i = np.where(i[:, ::-1] >= 0, 3 - i[:, ::-1], -1)
            ofs = f % 3
            nc = (L - ofs) // 3 if L - ofs >= 3 else 0
            t = i[:, ofs:ofs + 3 * nc].reshape(n, nc, 3)
            code = np.where((t < 0).any(axis=2), 64, 16 * t[:, :, 0] + 4 * t[:, :, 1] + t[:, :, 2]).T
            want.append(code.astype(np.uint8))
            roff.append(roff[-1] + nc)
    want, roff = np.concatenate(want, axis=0), np.array(roff, dtype=np.int64)
    nt = np.concatenate([r.reshape(-1) for r in raw])
    off = np.array([0, raw[0].size, raw[0].size + raw[1].size], dtype=np.int64)
    for tile_pos, grid_y in ((emul.k0_choose_tile_pos(max(lens), n, 6), 16), (48, 5), (1024, 1)):
        got = run_emul(emul, nt, off, lens, roff, 6, n, tile_pos, grid_y)
        assert np.array_equal(got, want), (tile_pos, grid_y)
