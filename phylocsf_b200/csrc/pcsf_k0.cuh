// K0: pleaves on the device - packed leaf codes per codon column and reading frame from nucleotide rows
// (src/PhyloCSF.ml:219-246 `pleaves` over the AsIs `candidate_regions` of :198-205; Code.ml:39-51 for the
// reverse complement). Byte work, bound by HBM on paper: per codon column and strand 3 n nucleotides in (the three
// frames of a strand share them) and n codes out per frame.
//
// Rounds 1-2 ran one thread per (column, leaf) with three byte loads each: 483 us per 2 M-column chunk = 410 GB/s,
// 6 % of the HBM peak, bound by its load instructions (profiles/r02_frame_codes_ncu_summary.json). This kernel moves
// whole words on both sides of a shared-memory tile that already has the OUTPUT's layout. One CTA per (alignment, tile of
// positions); for every frame the codes the tile contributes are one contiguous run of the output (column-major over
// leaves), and the tile keeps that run at the same 16-byte phase as its place in global memory:
//   stage 1  a thread takes 16 consecutive positions of one row: six aligned 32-bit loads (rows start at any byte
//            offset), funnel shifts to the row's phase, four characters decoded per instruction (SIMD within a word),
//            the codes of the 16 forward codons starting at those positions (four per instruction), and - 6 frames -
//            of their reverse complements (complement = 3 - index, bit operations on the codes); position p belongs to
//            column p / 3 of frame p % 3 (and to column (len - 3 - p) / 3 of the reverse frame (len - 3 - p) % 3):
//            the byte goes to run[frame][column * n_leaves + leaf] (STS.U8, three running offsets per strand);
//   stage 2  the runs leave as they are: LDS.128 -> STG.128 for every aligned 16-byte group, bytes for a run's
//            unaligned head and tail.
// Measured (profiles/r02_frame_codes_ncu_summary.json): 123.5 us per chunk = 1.89 TB/s = 29 % of the HBM peak, 3.9 x
// the old kernel; issue slots 78 % busy - instruction-bound at ~28 instructions per nucleotide: per 16 positions ~146 for
// the decode and the codons, ~100 for the 16 range-checked byte stores, ~115 of per-item overhead (item -> row / chunk,
// 64-bit addresses, the three frames' run descriptors), and a seventh of the iterations idle (1,102 items over 256 threads).
// An earlier form of this round kept the tile row-major (codes by leaf and position) and picked the output bytes out of
// it in stage 2 with per-byte column / leaf bookkeeping: 126.8 us - the same, so the simpler stage 2 stayed.
// Everything that indexes is in the PCSF_HD functions below so that the same code runs, thread by thread, in a CPU
// emulation (tests/k0_emul.cpp, tests/test_k0_emulation.py) against the oracle's pleaves on ragged inputs.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PCSF_HD __host__ __device__ __forceinline__
#else
#define PCSF_HD inline
#endif

namespace pcsf {
namespace k0 {

constexpr int THREADS = 256;
constexpr int MAX_FRAMES = 6;

// low word of (hi:lo) >> sh, sh = 0, 8, 16 or 24
PCSF_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}

// four characters -> four bytes: nucleotide index A 0, C 1, G 2, T 3 (either case), bit 2 set for anything else.
// 'A' 0x41, 'C' 0x43, 'G' 0x47, 'T' 0x54: bits 1-2 of the upper-cased character give the index (G and T swapped: the
// xor); a character is one of the four iff its other bits are those of 'A' - or those of 'T' when bits 2:1 are 10.
PCSF_HD uint32_t decode4(uint32_t w) {
    const uint32_t u = w & 0xDFDFDFDFu;
    const uint32_t x = (u >> 1) & 0x03030303u;
    const uint32_t idx = x ^ ((x >> 1) & 0x01010101u);
    const uint32_t t = (u >> 2) & ~(u >> 1) & 0x01010101u;
    const uint32_t bad = ((w & 0xD9D9D9D9u) ^ 0x41414141u) ^ (t * 0x11u);
    const uint32_t nz = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;
    return idx | (nz >> 5);
}

// four codons at once: a, b, c = the indices at positions p, p+1, p+2 (bytes). 16 a + 4 b + c, or 64 if any is not ACGT.
PCSF_HD uint32_t codon4(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t inv = ((a | b | c) & 0x04040404u) << 4;
    const uint32_t code = ((a & 0x03030303u) << 4) | ((b & 0x03030303u) << 2) | (c & 0x03030303u);
    return (code & ~(inv - (inv >> 6))) | inv;
}

// codes of the reverse complements of four forward codons: (i1, i2, i3) -> (3 - i3, 3 - i2, 3 - i1); 64 stays 64
PCSF_HD uint32_t revcomp4(uint32_t w) {
    const uint32_t y = w ^ 0x3F3F3F3Fu;
    const uint32_t r = ((y & 0x03030303u) << 4) | (y & 0x0C0C0C0Cu) | ((y >> 4) & 0x03030303u);
    const uint32_t m = w & 0x40404040u;
    return (r & ~(m - (m >> 6))) | m;
}

PCSF_HD int floordiv3(int x) { return x >= 0 ? x / 3 : -((2 - x) / 3); }

// Positions of a tile: [p0, p0 + tile_pos), tile_pos a multiple of 16. A frame's run has at most tile_pos / 3 + 2 columns;
// its slot in shared memory leaves room for the 16-byte phase in front.
PCSF_HD int slot_bytes(int tile_pos, int n_leaves) { return ((tile_pos / 3 + 2) * n_leaves + 15 + 15) & ~15; }

// The codes a (tile, frame) contributes: columns [cA, cB) of region r, `nbytes` bytes from byte g0 of the output;
// in shared memory the run starts at byte `sbase` (slot of the frame + g0 % 16).
struct Seg {
    int64_t g0;      // (region_off[r] + cA) * n_leaves
    int32_t nbytes;  // (cB - cA) * n_leaves; 0 for frames that were not asked for
    int32_t cA;
    int32_t sbase;
    int32_t ngroups; // aligned 16-byte groups the run touches
};

// frame f of an alignment of `len` positions whose region holds ncols columns starting at output column c0
PCSF_HD Seg make_seg(int f, int len, int64_t c0, int ncols, int p0, int tile_pos, int n_leaves) {
    const int ofs = f % 3;
    int cA, cB;
    if (f < 3) {  // column c reads the codon at ofs + 3 c
        cA = floordiv3(p0 - ofs + 2);              // ceil((p0 - ofs) / 3)
        cB = floordiv3(p0 + tile_pos - ofs + 2);   // ceil((p1 - ofs) / 3)
    } else {      // column c reads the codon at q = len - 3 - ofs - 3 c: q in [p0, p1)  <=>  c in ((top - p1) / 3, (top - p0) / 3]
        const int top = len - 3 - ofs;
        cA = floordiv3(top - (p0 + tile_pos)) + 1;
        cB = floordiv3(top - p0) + 1;
    }
    if (cA < 0) cA = 0;
    if (cB > ncols) cB = ncols;
    if (cB < cA) cB = cA;
    Seg s;
    s.g0 = (c0 + cA) * (int64_t)n_leaves;
    s.nbytes = (cB - cA) * n_leaves;
    s.cA = cA;
    s.sbase = f * slot_bytes(tile_pos, n_leaves) + (int)(s.g0 & 15);
    s.ngroups = s.nbytes > 0 ? (int32_t)(((s.g0 + s.nbytes + 15) >> 4) - (s.g0 >> 4)) : 0;
    return s;
}
PCSF_HD Seg empty_seg() {
    Seg s;
    s.g0 = 0; s.nbytes = 0; s.cA = 0; s.sbase = 0; s.ngroups = 0;
    return s;
}

// stage 1, one item: positions p .. p + 15 (p = p0 + 16 k) of leaf l's row, which starts at byte `row_byte` of the buffer
// (row_byte = aln_off + l * len + p). Reads whole words only, never past word `nwords` - 1. segs[0..5]: all six frames
// (nbytes 0 where a frame is not wanted). Codes of codons that run past the row's end belong to no column and are dropped
// by the range test, like everything else that falls outside the tile's runs.
PCSF_HD void stage1_item(const uint32_t* ntw, int64_t nwords, int64_t row_byte, int l, int p, int len, const Seg* segs, bool both_strands,
                         int n_leaves, uint8_t* sm) {
    const int64_t w0 = row_byte >> 2;
    const uint32_t sh = (uint32_t)(row_byte & 3) * 8u;
    uint32_t w[6];
#pragma unroll
    for (int m = 0; m < 6; m++) w[m] = (w0 + m < nwords) ? ntw[w0 + m] : 0u;
    uint32_t I[5];
#pragma unroll
    for (int m = 0; m < 5; m++) I[m] = decode4(fsr(w[m], w[m + 1], sh));
    uint32_t cc[4];
#pragma unroll
    for (int m = 0; m < 4; m++) cc[m] = codon4(I[m], fsr(I[m], I[m + 1], 8u), fsr(I[m], I[m + 1], 16u));
    {   // forward strand: position p + i is column (p + i) / 3 of frame (p + i) % 3
        const int c0 = p / 3, f0 = p - 3 * c0;
        int rel[3], base[3];
        uint32_t nb[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int t = f0 + j, f = t >= 3 ? t - 3 : t, c = t >= 3 ? c0 + 1 : c0;
            rel[j] = (c - segs[f].cA) * n_leaves;
            nb[j] = (uint32_t)segs[f].nbytes;
            base[j] = segs[f].sbase + l;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int j = i % 3;
            if ((uint32_t)rel[j] < nb[j]) sm[base[j] + rel[j]] = (uint8_t)(cc[i >> 2] >> (8 * (i & 3)));
            rel[j] += n_leaves;
        }
    }
    if (both_strands) {  // reverse strand: with u = len - 3 - (p + i), column u / 3 of frame 3 + u % 3 (none if u < 0)
        const int u0 = len - 3 - p, c0 = floordiv3(u0), f0 = u0 - 3 * c0;
        int rel[3], base[3];
        uint32_t nb[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int t = f0 - j, f = t < 0 ? t + 3 : t, c = t < 0 ? c0 - 1 : c0;
            rel[j] = (c - segs[3 + f].cA) * n_leaves;
            nb[j] = (uint32_t)segs[3 + f].nbytes;
            base[j] = segs[3 + f].sbase + l;
        }
#pragma unroll
        for (int m = 0; m < 4; m++) cc[m] = revcomp4(cc[m]);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int j = i % 3;
            if ((uint32_t)rel[j] < nb[j]) sm[base[j] + rel[j]] = (uint8_t)(cc[i >> 2] >> (8 * (i & 3)));
            rel[j] -= n_leaves;
        }
    }
}

// stage 2, one item: aligned group gi of a run goes out as it is
PCSF_HD void stage2_group(const Seg& s, int gi, const uint8_t* sm, uint8_t* codes) {
    const int64_t A = (s.g0 & ~(int64_t)15) + 16 * (int64_t)gi;
    const int j0 = (int)(A - s.g0);  // >= -15: index, within the run, of the group's first byte
    const uint8_t* src = sm + s.sbase + j0;  // 16-byte aligned: sbase = slot + g0 % 16
    if (j0 >= 0 && j0 + 16 <= s.nbytes) {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint4*>(codes + A) = *reinterpret_cast<const uint4*>(src);
#else
        for (int i = 0; i < 16; i++) codes[A + i] = src[i];
#endif
    } else {
        const int lo = j0 < 0 ? -j0 : 0;
        const int hi = s.nbytes - j0 < 16 ? s.nbytes - j0 : 16;
        for (int i = lo; i < hi; i++) codes[A + i] = src[i];
    }
}

// What one CTA does for one (alignment, tile); `tid` strides by `nthreads`. The two loops are separated by a
// barrier in the kernel (and by running all threads of the first before the second in the emulation).
PCSF_HD void stage1_thread(int tid, int nthreads, const uint32_t* ntw, int64_t nwords, int64_t aln_byte, int len, int n_leaves,
                           int p0, int tile_pos, const Seg* segs, int frames, uint8_t* sm) {
    const int npos = (len - p0 < tile_pos ? len - p0 : tile_pos);  // positions of this tile inside the row
    const int nchunk = (npos + 15) >> 4;
    const int nitems = nchunk * n_leaves;
    for (int it = tid; it < nitems; it += nthreads) {
        const int l = it / nchunk, k = it - l * nchunk;
        const int p = p0 + 16 * k;
        stage1_item(ntw, nwords, aln_byte + (int64_t)l * len + p, l, p, len, segs, frames > 3, n_leaves, sm);
    }
}

PCSF_HD void stage2_thread(int tid, int nthreads, const Seg* segs, int frames, const uint8_t* sm, uint8_t* codes) {
    int total = 0;
    for (int f = 0; f < frames; f++) total += segs[f].ngroups;
    for (int g = tid; g < total; g += nthreads) {
        int f = 0, gi = g;
        while (gi >= segs[f].ngroups) gi -= segs[f++].ngroups;
        stage2_group(segs[f], gi, sm, codes);
    }
}

// tile size (positions) for a batch: the whole alignment when its runs fit `budget` bytes of shared memory, else the
// largest multiple of 16 that does (at least 16)
inline size_t smem_bytes(int tile_pos, int n_leaves, int frames) { return (size_t)frames * slot_bytes(tile_pos, n_leaves); }
inline int choose_tile_pos(int max_len, int n_leaves, int frames, int budget = 40 * 1024) {
    const int whole = ((max_len > 1 ? max_len : 1) + 15) & ~15;
    int fit = 16;
    while (fit + 16 <= whole && smem_bytes(fit + 16, n_leaves, frames) <= (size_t)budget) fit += 16;
    return fit;
}

}  // namespace k0

#if defined(__CUDACC__)
// grid (alignments, tiles per alignment): CTA (a, y) does tiles y, y + gridDim.y, ... of alignment a.
// nt: 4-byte aligned buffer of nt_bytes bytes (reads are whole words inside [0, ceil(nt_bytes / 4))); codes: 16-byte aligned.
__global__ void __launch_bounds__(k0::THREADS) frame_codes_kernel(const uint8_t* __restrict__ nt, int64_t nt_bytes,
                                                                  const int64_t* __restrict__ aln_off,
                                                                  const int32_t* __restrict__ aln_len,
                                                                  const int64_t* __restrict__ region_off, int frames,
                                                                  int n_leaves, int tile_pos, uint8_t* __restrict__ codes) {
    extern __shared__ uint4 k0_smem[];
    __shared__ k0::Seg segs[k0::MAX_FRAMES];
    const int64_t a = blockIdx.x;
    const int len = aln_len[a];
    const int64_t aln_byte = aln_off[a];
    const int64_t nwords = (nt_bytes + 3) >> 2;
    for (int64_t q0 = (int64_t)blockIdx.y * tile_pos; q0 + 3 <= len; q0 += (int64_t)gridDim.y * tile_pos) {
        const int p0 = (int)q0;
        if ((int)threadIdx.x < k0::MAX_FRAMES) {
            k0::Seg s = k0::empty_seg();
            if ((int)threadIdx.x < frames) {
                const int64_t r = a * frames + threadIdx.x;
                const int64_t c0 = region_off[r];
                s = k0::make_seg((int)threadIdx.x, len, c0, (int)(region_off[r + 1] - c0), p0, tile_pos, n_leaves);
            }
            segs[threadIdx.x] = s;
        }
        __syncthreads();
        k0::stage1_thread((int)threadIdx.x, k0::THREADS, reinterpret_cast<const uint32_t*>(nt), nwords, aln_byte, len, n_leaves,
                          p0, tile_pos, segs, frames, reinterpret_cast<uint8_t*>(k0_smem));
        __syncthreads();
        k0::stage2_thread((int)threadIdx.x, k0::THREADS, segs, frames, reinterpret_cast<const uint8_t*>(k0_smem), codes);
        __syncthreads();  // the next tile's runs overwrite the slots
    }
}
#endif

}  // namespace pcsf
