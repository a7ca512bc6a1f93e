// pcsf_kernels.cuh — sm_100a kernels of the PhyloCSF scoring path.
//
//   K1  pt_build_kernel      P(t) = S * (diag(exp(lambda t)) * Sinv) per (scale, branch) on FP64 DMMA,
//                            with the reference's clamp / row-sum / diagonal fix-ups
//                            (lib/CamlPaml/Q.ml:211-249, looped by PhyloModel.ml:17)
//   K2+K3 prune_kernel       Felsenstein pruning of 128 codon columns per CTA over the whole tree:
//                            leaf messages are gathers of P columns, internal edges are
//                            64x64 (P) x 64xNcols (partials) contractions on FP64 DMMA with the
//                            partials held in registers between edges; root dot / log / posterior
//                            (lib/CamlPaml/PhyloLik.ml:73-93,127-138; src/PhyloCSFModel.ml:76-81)
//   K4  region_reduce_kernel per-region sums of the per-column terms (src/PhyloCSFModel.ml:79-81)
//   K0  frame_codes_kernel   pleaves on the device (src/PhyloCSF.ml:219-246) from nucleotide rows
//   K5  omega_eig_kernel     omega-model rate matrix assembly + batched Jacobi diagonalisation
//                            (src/OmegaModel.ml:21-80, lib/CamlPaml/Q.ml:124-177)
//
// FP64 tensor path on sm_100a is warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); tcgen05 has no
// f64 kind. Operand staging uses TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarriers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pcsf_program.hpp"  // Op / Item / SubTab and the host-side builder of the tree programs

namespace pcsf {

constexpr int K = 64;
constexpr int PT_SLOT = 65 * 64;         // doubles per (scale, branch) table slot
constexpr int PT_SLOT_BYTES = PT_SLOT * 8;
constexpr int FRAG_BYTES = 64 * 64 * 8;  // fragment-ordered P image of an internal edge

// ---- work description --------------------------------------------------------------------------
// A span = a run of codon columns scored under one P set (one model at one tree scale). Tiles of
// TILE_COLS columns are cut from spans; tile0 = index of the span's first tile.
struct Span {
    int64_t col0;   // first column in the codes array
    int64_t out0;   // first slot in the per-column output arrays
    int64_t tile0;  // prefix count of tiles
    int32_t ncols;
    int32_t pset;   // index into the PSet table
};
struct PSet {
    const double* tables;    // [n_branches][PT_SLOT]
    const double* prior;     // [64]
    const double* logprior;  // [64]
    const double* cherry;    // subtree tables of this P set (SubTab::off), or null when not built
    long long tab_level;     // 0 none, 2 cherries, 3 + cherry-and-leaf subtrees, 4 + caterpillars of four
};

// ---- small PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// D(8x8) += A(8x4) * B(4x8), FP64. lane = 4g+t: a = A[g][t], b = B[t][g], c = D[g][2t..2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Position of P[a][b] (a = parent state, b = child state) in the fragment-ordered image of an
// internal edge. The image is read as the B operand of D[col][a] += alpha[col][b] * P[a][b]:
// n-tile j = a/8, g = a%8; the contraction index is visited in the order the accumulator
// registers of the previous edge already hold it: k-step s, slot t  <->  b = 8*(s/2) + 2t + (s%2).
__host__ __device__ __forceinline__ int frag_index(int a, int b) {
    const int j = a >> 3, g = a & 7;
    const int s = ((b >> 3) << 1) | (b & 1), t = (b >> 1) & 3;
    return ((j * 16 + s) * 32) + (g * 4 + t);
}

// =================================================================================================
// K1: P(t) build
// =================================================================================================
// Persistent CTAs (two per SM), each owning a contiguous run of (job, branch) items; a job = one
// diagonalised model at one tree scale. Per model the CTA keeps S as register-resident A fragments and
// S^-1 as a fragment-ordered image in shared memory; per item it
//   - evaluates e_k = exp(lambda_k * scale * branch_len) (64 threads),
//   - warp w computes rows 8w..8w+7 of (S diag(e)) S^-1 with 128 DMMAs,
//   - applies the reference's row fix-ups (lib/CamlPaml/Q.ml:226-247: clamp (-tol,0) to 0, row sum within
//     tol of 1, diagonal := 1 - sum of the off-diagonal entries) on the accumulators - a row lives in the
//     four lanes of a quad, so two shuffles finish each sum -,
//   - stores the slot in the layout the pruning kernel consumes: leaves get the transposed gather table
//     PT[code][a] (+ row 64 = row sums, the `Marginalize message), internal edges the fragment-ordered image.
// Versus the reference's evaluation order, diag(e) is folded into S instead of S^-1 and the row sums are
// tree-shaped; both move entries by ~1e-16 (test bar: 2e-13 absolute).
struct PtJob {               // one P set to build: a diagonalised model at one tree scale
    const double* params;   // S | Sinv | lambda | prior | logprior (pcsf_api.cu: Model)
    double scale;
};
// Round-2 measurements on this kernel (tools/bench_k1.py: 4.76 M slots = 20,000 candidate scales x 238 branches;
// profiles/r02_k1_ablation.json; the round-1 figure of 98 % of the DMMA peak divided by a slot count that was 1.5 x too
// high - it was 66 %):
//   round 1: eight warps per slot, each 8 rows; 64 exp by two warps right before a per-slot barrier; 16 + 16 adds and
//     16 FP64 compares per thread for the fix-ups                                                         24.4 TFLOP/s
//   + sign test on the integer pipe, off-diagonal sum = row sum - P[i][i] unless the warp holds a negative entry,
//     exponentials one slot ahead, eight per warp                                                         26.5 = 71 %
//   row sum as a ninth n-tile on the tensor pipe instead of 16 adds                                       25.0
//   timing-only ablations of that form: no stores 29.9; no stores, no exp, no S diag(e) products 30.0 (81 %): what was
//     left was one LDS.64 of S^-1 per DMMA (LSU data pipe 54 % busy) and the barrier per slot
//   a phased form (one CTA of 16 warps per SM, exponentials from a pre-pass kernel, two barriers per pair of slots so
//     that the pipe sees either plain FP64 instructions or DMMAs): 21.7 - the idle pipe at the phase changes costs more
//     than the contention it removes
//   this form: a warp owns 16 rows of a slot (four warps per slot, two slots per CTA step), so every S^-1 fragment it
//     loads feeds two DMMAs; each warp computes the slot's 64 exponentials itself (two per lane) and keeps them in its
//     own scratch line, so there is no barrier per slot at all - warps only meet when the model changes.       30.5 = 82 %
//   the same with the rows leaving through a per-warp staging line and cp.async.bulk shared -> global (one 8 KB copy
//     per warp for an internal slot, 64 copies of 128 B for a leaf slot; one CTA of 16 warps per SM): 29.6 - the
//     stores are not what limits this form. Not kept.
constexpr int K1_THREADS = 256;
constexpr int K1_SMEM = (4096 + 4096 + 64 + 8 * 64) * 8;  // S^-1 image, S image, lambda, per-warp exponentials
__global__ void __launch_bounds__(K1_THREADS, 2) pt_build_kernel(const PtJob* __restrict__ jobs, long long n_items,
                                                                 const double* __restrict__ branch_len, int n_branches,
                                                                 int n_leaves, double* __restrict__ tables,
                                                                 int32_t* __restrict__ status, double tol) {
    extern __shared__ __align__(16) double k1sm[];
    double* Sinv_s = k1sm;          // B fragments: [(j*16+s)*32 + lane] = Sinv[4s+t][8j+g]
    double* S_s = k1sm + 4096;      // A fragments: [(m*16+s)*32 + lane] = S[8m+g][4s+t], m = block of 8 rows
    double* lam_s = S_s + 4096;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* ew = lam_s + 64 + w * 64;  // this warp's exponentials
    const int sub = w >> 2, rb = w & 3;  // which slot of the pair, which block of 16 rows
    // contiguous runs of an even number of items: a pair of items never straddles two jobs (n_branches = 2n-2 is even)
    long long per = (n_items + gridDim.x - 1) / gridDim.x;
    per += per & 1;
    const long long lo = per * blockIdx.x, hi = min(n_items, lo + per);
    const double* cur_params = nullptr;
    for (long long pair = lo; pair < hi; pair += 2) {
        const long long job_i = pair / n_branches;
        const PtJob job = jobs[job_i];
        if (job.params != cur_params) {  // uniform across the CTA
            __syncthreads();             // everybody is done with the previous model's images
            const double* S = job.params;
            const double* Sinv = job.params + 4096;
            for (int idx = tid; idx < 4096; idx += K1_THREADS) {
                const int l = idx & 31, s = (idx >> 5) & 15, j = idx >> 9;
                Sinv_s[idx] = Sinv[(4 * s + (l & 3)) * 64 + 8 * j + (l >> 2)];
                S_s[idx] = S[(8 * j + (l >> 2)) * 64 + 4 * s + (l & 3)];
            }
            if (tid < 64) lam_s[tid] = job.params[8192 + tid];
            cur_params = job.params;
            __syncthreads();
        }
        const long long item = pair + sub;
        if (item >= hi) continue;
        const int br = (int)(item - job_i * n_branches);
        const double tt = job.scale * branch_len[br];  // Mul (Var 0, Val b), src/PhyloCSFModel.ml:33
        __syncwarp();                                  // the previous slot's reads of ew are over
        ew[lane] = exp(tt * lam_s[lane]);              // Q.ml:216-217
        ew[lane + 32] = exp(tt * lam_s[lane + 32]);
        __syncwarp();
        double acc[2][8][2];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[m][j][0] = acc[m][j][1] = 0.0;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const double e = ew[4 * s + t];
            const double a0 = S_s[((2 * rb) * 16 + s) * 32 + lane] * e, a1 = S_s[((2 * rb + 1) * 16 + s) * 32 + lane] * e;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double bf = Sinv_s[(j * 16 + s) * 32 + lane];
                dmma(acc[0][j][0], acc[0][j][1], a0, bf);
                dmma(acc[1][j][0], acc[1][j][1], a1, bf);
            }
        }
        // ---- fix-ups (Q.ml:226-247) on rows i = 16 rb + 8 m + g, whose 64 entries sit in the quad's accumulators ----
        int st = (tt < 0.0) ? 1 : 0;
        double* out = tables + (size_t)item * PT_SLOT;
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const int mt = 2 * rb + m, i = 8 * mt + g;
            bool neg = false;
#pragma unroll
            for (int j = 0; j < 8; j++) neg |= (__double2hiint(acc[m][j][0]) < 0) | (__double2hiint(acc[m][j][1]) < 0);
            double tot = 0.0, off;
#pragma unroll
            for (int j = 0; j < 8; j++) tot += acc[m][j][0] + acc[m][j][1];
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            if (__any_sync(0xffffffffu, neg)) {  // rare: the reference's explicit pre-clamp / post-clamp sums
                tot = 0.0;
                off = 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double v = acc[m][j][e];
                        tot += v;
                        if (v < 0.0) {
                            if (fabs(v) > tol) st |= 2;
                            v = 0.0;
                            acc[m][j][e] = 0.0;
                        }
                        if (8 * j + 2 * t + e != i) off += v;
                    }
                tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                off += __shfl_xor_sync(0xffffffffu, off, 1);
                off += __shfl_xor_sync(0xffffffffu, off, 2);
            } else {  // P[i][i] is entry (n-tile mt, slot g & 1) of lane t = g >> 1
                double pii = 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j == mt) pii = (g & 1) ? acc[m][j][1] : acc[m][j][0];
                pii = __shfl_sync(0xffffffffu, pii, (lane & ~3) | (g >> 1));
                off = tot - pii;
            }
            const double smii = 1.0 - off;
            if (fabs(tot - 1.0) > tol) st |= 4;
            if (!(smii <= 1.0 && smii > 0.0)) st |= 8;
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (8 * j + 2 * t + e == i) acc[m][j][e] = smii;
            if (br < n_leaves) {
#pragma unroll
                for (int j = 0; j < 8; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) out[(8 * j + 2 * t + e) * 64 + i] = acc[m][j][e];
                if (t == 0) out[4096 + i] = off + smii;  // ddot(row, ones): the `Marginalize leaf message
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) out[((mt * 16 + 2 * j + e) * 32) + lane] = acc[m][j][e];
            }
        }
        if (st) atomicOr(&status[job_i], st);
    }
}

// =================================================================================================
// K2+K3: pruning
// =================================================================================================
#ifndef PCSF_PRUNE_T
#define PCSF_PRUNE_T 2
#endif
constexpr int PRUNE_T = PCSF_PRUNE_T;                        // 8-column tiles per compute warp
constexpr int PRUNE_WARPS = 16 / PRUNE_T;                    // compute warps (warpgroups 1..): 16 (T=1) or 8 (T=2)
constexpr int PROD_THREADS = 128;                            // warpgroup 0: data movement only
constexpr int PRUNE_THREADS = PROD_THREADS + PRUNE_WARPS * 32;  // 640 / 384
constexpr int WARP_COLS = 8 * PRUNE_T;                       // 8 / 16
constexpr int TILE_COLS = PRUNE_WARPS * WARP_COLS;           // 128
constexpr int STACK_ENTRY_BYTES = PRUNE_T * 8 * 32 * 16;     // one warp's partial: 8 KB
constexpr int STACK_LEVEL_BYTES = PRUNE_WARPS * STACK_ENTRY_BYTES;  // one CTA's partial: 64 KB
constexpr int P_STAGES = 2;                                  // ring of P images, 32 KB each
constexpr int M_STAGES = 2;                                  // ring of multiplicand buffers, 64 KB each
constexpr int MAX_STACK_LEVELS = 16;
constexpr int PRUNE_BAR_BYTES = 256;                         // pfull[2] pempty[2] mfull[2] mempty[2] pushed[16]
// setmaxnreg targets. Launch allocation is 65536 / threads (96 for 640 threads, 168 for 384).
constexpr int PROD_REGS = PRUNE_T == 1 ? 32 : 56;  // inc can only take what dec released (CTA pool): (launch-PROD)*128 >= (COMPUTE-launch)*32*PRUNE_WARPS
constexpr int COMPUTE_REGS = PRUNE_T == 1 ? 112 : 224;

struct PruneParams {
    const Op* ops;
    int n_ops;
    const Item* items;
    int n_items;
    int n_leaves;
    const Span* spans;
    int n_spans;
    int64_t n_tiles;
    const PSet* psets;
    const uint8_t* codes;  // [total_cols][n_leaves]
    double* out_logz;
    double* out_anc;
    uint8_t* global_stack;   // [gridDim.x][n_levels][STACK_LEVEL_BYTES]
    int n_levels;
    int ops_bytes;           // n_ops * sizeof(Op) rounded up to 16
    int items_bytes;         // n_items * sizeof(Item) rounded up to 16
    int skew_ns;             // start-up offset of the second compute warp of every SM sub-partition
    int32_t* global_exp;     // [gridDim.x][n_levels][TILE_COLS] exponents of parked partials (rescale only)
    const long long* tab_off;  // table program: offset of every table in a P set's table block
    long long* timeline;     // PCSF_TIMELINE builds only: [warp][event][2] = (code, clock64) of CTA 0
    int timeline_cap;
};

#ifdef PCSF_TIMELINE
#define TL_MARK(code)                                                                        \
    do {                                                                                     \
        if (p.timeline && blockIdx.x == 0 && lane == 0 && tl_n < p.timeline_cap) {            \
            p.timeline[((size_t)cw * p.timeline_cap + tl_n) * 2] = (long long)(code);         \
            p.timeline[((size_t)cw * p.timeline_cap + tl_n) * 2 + 1] = clock64();             \
            tl_n++;                                                                          \
        }                                                                                    \
    } while (0)
#else
#define TL_MARK(code) do { } while (0)
#endif

__device__ __forceinline__ int find_span(const Span* __restrict__ spans, int n_spans, int64_t tile) {
    int lo = 0, hi = n_spans - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (spans[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// arrive on `bar` once all cp.async of this thread issued so far have landed (counts as a normal arrival)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// A ring stage is read through the generic proxy (LDS) and refilled through the async proxy (TMA). The
// empty-barrier arrive alone does not order the two: without a proxy fence between a warp's last read of a
// stage and its arrive, a refill was observed to overtake reads still queued in the load/store unit (rare
// single-entry corruptions, run-to-run differences of 1e-13 in log z). Every lane fences its own reads.
__device__ __forceinline__ void release_stage(uint64_t* empty_bar, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar);
}

// Underflow rescue (optional; the reference has none, lib/CamlPaml/PhyloLik.ml:87-92): when the largest
// of a column's 64 partials drops below 2^-256 the column is multiplied by the exact power of two that
// brings it back to [1,2) and the exponent is remembered; log z gets it back at the root. Columns that
// never come near the threshold are computed exactly as without the option.
__device__ __forceinline__ void rescale_columns(double (&cur)[PRUNE_T][8][2], int (&esum)[PRUNE_T]) {
#pragma unroll
    for (int T = 0; T < PRUNE_T; T++) {
        double m = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) m = fmax(m, fmax(cur[T][j][0], cur[T][j][1]));
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
        if (m > 0.0 && m < 0x1p-256) {  // uniform within the quad
            const int e = ilogb(m);
            const double f = scalbn(1.0, -e);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                cur[T][j][0] *= f;
                cur[T][j][1] *= f;
            }
            esum[T] += e;
        }
    }
}

// One CTA per SM, warp-specialised:
//   warpgroup 0 (128 threads, registers trimmed to 56) moves data only. In program order it stages
//     - the P image of every internal edge into a 2 x 32 KB ring (TMA bulk copy, one thread),
//     - every multiplicand an epilogue needs into a 2 x 64 KB ring, already in the register layout
//       of the compute warps: leaf messages as a cp.async gather of P^T rows selected by the leaf
//       codes (K2), parked partials as a TMA bulk copy back from the L2-resident stack,
//     running up to one tree edge ahead of the compute warps, across tile boundaries.
//   warpgroups 1-2 (8 warps, registers raised to 224) each own 16 codon columns of the 128-column
//     tile and walk the tree with the partials in registers: DMMA contraction against the staged P
//     image (K3), then one pass of LDS.128 + DMUL over the staged multiplicand.
// All hand-offs are mbarriers; there is no CTA-wide barrier after start-up.
// shared memory map: [P ring | M ring | barriers | ops | items | tile codes]
template <bool RESCALE>
__global__ void __launch_bounds__(PRUNE_THREADS, 1) prune_kernel(const PruneParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* Pring = smem;
    uint8_t* Mring = smem + P_STAGES * FRAG_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Mring + M_STAGES * STACK_LEVEL_BYTES);
    uint64_t* pfull = bars;
    uint64_t* pempty = bars + 2;
    uint64_t* mfull = bars + 4;
    uint64_t* mempty = bars + 6;
    uint64_t* pushed = bars + 8;  // [MAX_STACK_LEVELS]
    Op* ops_s = reinterpret_cast<Op*>(reinterpret_cast<uint8_t*>(bars) + PRUNE_BAR_BYTES);
    Item* items_s = reinterpret_cast<Item*>(reinterpret_cast<uint8_t*>(ops_s) + p.ops_bytes);
    uint8_t* codes_s = reinterpret_cast<uint8_t*>(items_s) + p.items_bytes;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < P_STAGES; i++) {
            mbar_init(&pfull[i], 1);
            mbar_init(&pempty[i], PRUNE_WARPS);
        }
        for (int i = 0; i < M_STAGES; i++) {
            mbar_init(&mfull[i], PROD_THREADS);
            mbar_init(&mempty[i], PRUNE_WARPS);
        }
        for (int i = 0; i < MAX_STACK_LEVELS; i++) mbar_init(&pushed[i], PRUNE_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = tid; i < p.n_ops; i += PRUNE_THREADS) ops_s[i] = p.ops[i];
    for (int i = tid; i < p.n_items; i += PRUNE_THREADS) items_s[i] = p.items[i];
    __syncthreads();

    uint8_t* gstack = p.global_stack + (size_t)blockIdx.x * p.n_levels * STACK_LEVEL_BYTES;

    if (tid < PROD_THREADS) {
        // ====================================== producer warpgroup ======================================
        reg_dealloc<PROD_REGS>();
        uint32_t pq = 0, mq = 0;  // P images / multiplicands staged so far
        uint32_t pops = 0;        // bit l = parity of the next pop of stack level l
        // element e = tid + 128 i of a multiplicand: compute warp e/512, column tile (e/256)%2, state
        // block (e/32)%8, lane e%32 -> 16 bytes at offset 16 e of the stage
        for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const Span sp = p.spans[find_span(p.spans, p.n_spans, tile)];
            const double* tables = p.psets[sp.pset].tables;
            const int64_t tcol0 = (tile - sp.tile0) * TILE_COLS;
            const int ncols = (int)min((int64_t)TILE_COLS, (int64_t)sp.ncols - tcol0);
            {   // leaf codes of the tile (columns past the end of the span marginalise)
                asm volatile("bar.sync 1, 128;" ::: "memory");  // every gather of the previous tile has been issued
                const uint8_t* src = p.codes + (size_t)(sp.col0 + tcol0) * p.n_leaves;
                const int nbytes = ncols * p.n_leaves, total = TILE_COLS * p.n_leaves;
                for (int i = tid; i < total; i += PROD_THREADS) codes_s[i] = (i < nbytes) ? src[i] : (uint8_t)64;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            for (int ii = 0; ii < p.n_items; ii++) {
                const Item it = items_s[ii];
                if (it.kind == ITEM_P) {
                    if (tid == 0) {
                        const uint32_t st = pq % P_STAGES, u = pq / P_STAGES;
                        if (u >= 1) mbar_wait(&pempty[st], (u - 1) & 1);
                        mbar_expect_tx(&pfull[st], FRAG_BYTES);
                        tma_bulk_g2s(Pring + st * FRAG_BYTES, tables + (size_t)it.a * PT_SLOT, FRAG_BYTES, &pfull[st]);
                    }
                    pq++;
                    continue;
                }
                const uint32_t st = mq % M_STAGES, u = mq / M_STAGES;
                if (u >= 1) mbar_wait(&mempty[st], (u - 1) & 1);
                uint8_t* dst = Mring + st * STACK_LEVEL_BYTES;
                if (it.kind == ITEM_LEAF) {
                    const double* tab = tables + (size_t)it.a * PT_SLOT;
                    const int lane = tid & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 4
                    for (int i = 0; i < 32; i++) {
                        const int e = tid + PROD_THREADS * i;  // 16-byte element of the stage
                        // e = ((warp * T + tile) * 8 + j) * 32 + lane, so (e >> 8) = warp * T + tile = column / 8
                        const int j = (e >> 5) & 7, col = ((e >> 8) << 3) + g;
                        int code = codes_s[col * p.n_leaves + it.a];
                        code = code > 64 ? 64 : code;
                        cp_async_16(dst + 16 * e, tab + code * 64 + 8 * j + 2 * t);
                    }
                    cp_async_arrive(&mfull[st]);
                } else {  // ITEM_POP: the compute warps' stores to this level must be visible first
                    mbar_wait(&pushed[it.a], (pops >> it.a) & 1);
                    pops ^= 1u << it.a;
                    if (tid == 0) {
                        mbar_expect_tx(&mfull[st], STACK_LEVEL_BYTES);
                        tma_bulk_g2s(dst, gstack + (size_t)it.a * STACK_LEVEL_BYTES, STACK_LEVEL_BYTES, &mfull[st]);
                    } else {
                        mbar_arrive(&mfull[st]);
                    }
                }
                mq++;
            }
        }
        return;
    }

    // ========================================= compute warps =========================================
    reg_alloc<COMPUTE_REGS>();
    const int ctid = tid - PROD_THREADS, lane = ctid & 31, cw = ctid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wcol = cw * WARP_COLS;
    uint32_t pq = 0, mq = 0;
    // The two compute warps of an SM sub-partition start half an edge apart so that one is in its
    // epilogue while the other keeps the DMMA pipe busy; nothing re-synchronises them afterwards.
    bool skew_pending = cw >= PRUNE_WARPS / 2 && p.skew_ns > 0;  // applied once the first data has arrived
#ifdef PCSF_TIMELINE
    int tl_n = 0;
#endif

    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const Span sp = p.spans[find_span(p.spans, p.n_spans, tile)];
        const PSet ps = p.psets[sp.pset];
        const int64_t tcol0 = (tile - sp.tile0) * TILE_COLS;  // first column of the tile within the span
        const int ncols = (int)min((int64_t)TILE_COLS, (int64_t)sp.ncols - tcol0);
        const bool warp_active = wcol < ncols;

        double cur[PRUNE_T][8][2];
        int esum[PRUNE_T];  // power-of-two exponent taken out of each column so far (rescale option)
#pragma unroll
        for (int T = 0; T < PRUNE_T; T++) {
            esum[T] = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) cur[T][j][0] = cur[T][j][1] = 1.0;
        }
        int32_t* gexp = p.global_exp + ((size_t)blockIdx.x * p.n_levels) * TILE_COLS + wcol + g;

        for (int oi = 0; oi < p.n_ops; oi++) {
            const Op op = ops_s[oi];
            if (op.kind == OP_CHERRY) {
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const uint32_t st = mq % M_STAGES;
                    mbar_wait(&mfull[st], (mq / M_STAGES) & 1);
                    if (warp_active) {
                        const double2* m = reinterpret_cast<const double2*>(Mring + st * STACK_LEVEL_BYTES + cw * STACK_ENTRY_BYTES) + lane;
#pragma unroll
                        for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const double2 v = m[(T * 8 + j) * 32];
                                cur[T][j][0] = k ? cur[T][j][0] * v.x : v.x;
                                cur[T][j][1] = k ? cur[T][j][1] * v.y : v.y;
                            }
                    }
                    release_stage(&mempty[st], lane);
                    mq++;
                }
                continue;
            }
            if (op.kind == OP_ROOT) {
                if (warp_active) {
                    const double2* pr = reinterpret_cast<const double2*>(ps.prior + 2 * t);
                    const double2* lp = reinterpret_cast<const double2*>(ps.logprior + 2 * t);
#pragma unroll
                    for (int T = 0; T < PRUNE_T; T++) {
                        double zp = 0.0;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const double2 pj = __ldg(pr + 4 * j);
                            cur[T][j][0] *= pj.x;  // alpha_root[x] * prior[x]
                            cur[T][j][1] *= pj.y;
                            zp += cur[T][j][0];
                            zp += cur[T][j][1];
                        }
                        zp += __shfl_xor_sync(0xffffffffu, zp, 1);
                        zp += __shfl_xor_sync(0xffffffffu, zp, 2);
                        double ap = 0.0;
                        if (zp != 0.0) {  // PhyloLik.ml:131-132: impossible data => zero posterior
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const double2 lj = __ldg(lp + 4 * j);
                                ap += (cur[T][j][0] / zp) * lj.x;
                                ap += (cur[T][j][1] / zp) * lj.y;
                            }
                        }
                        ap += __shfl_xor_sync(0xffffffffu, ap, 1);
                        ap += __shfl_xor_sync(0xffffffffu, ap, 2);
                        const int c = wcol + 8 * T + g;
                        if (t == 0 && c < ncols) {
                            p.out_logz[sp.out0 + tcol0 + c] = (RESCALE && esum[T]) ? log(zp) + (double)esum[T] * 0.6931471805599453 : log(zp);
                            p.out_anc[sp.out0 + tcol0 + c] = ap;
                        }
                    }
                }
                continue;
            }
            // ------------------------------ K3: contraction over one internal edge ------------------------------
            TL_MARK(oi * 8 + 0);
            const uint32_t pst = pq % P_STAGES;
            mbar_wait(&pfull[pst], (pq / P_STAGES) & 1);
            if (skew_pending) {
                __nanosleep(p.skew_ns);
                skew_pending = false;
            }
            TL_MARK(oi * 8 + 1);
            double acc[PRUNE_T][8][2];
            if (warp_active) {
                const double* Pb = reinterpret_cast<const double*>(Pring + pst * FRAG_BYTES) + lane;
#pragma unroll
                for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[T][j][0] = acc[T][j][1] = 0.0;
#pragma unroll
                for (int s = 0; s < 16; s++) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double bf = Pb[(j * 16 + s) * 32];
#pragma unroll
                        for (int T = 0; T < PRUNE_T; T++) dmma(acc[T][j][0], acc[T][j][1], cur[T][s >> 1][s & 1], bf);
                    }
                }
            }
            release_stage(&pempty[pst], lane);
            pq++;
            TL_MARK(oi * 8 + 2);
            // ------------------------------ epilogue ------------------------------
            if (op.kind == OP_GEMM_PUSH) {
                if (warp_active) {
                    double2* slot = reinterpret_cast<double2*>(gstack + (size_t)op.c * STACK_LEVEL_BYTES + cw * STACK_ENTRY_BYTES) + lane;
#pragma unroll
                    for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) slot[(T * 8 + j) * 32] = make_double2(acc[T][j][0], acc[T][j][1]);
                    if (RESCALE) {  // the parked message keeps its exponent; the sibling subtree starts at 0
#pragma unroll
                        for (int T = 0; T < PRUNE_T; T++) {
                            if (t == 0) gexp[(size_t)op.c * TILE_COLS + 8 * T] = esum[T];
                            esum[T] = 0;
                        }
                    }
                    // The producer reads the level back with a TMA bulk copy (async proxy). Writer-side
                    // ordering: CTA-scope fence (writer and reader share the SM), generic->async proxy fence,
                    // then the release-arrive on pushed[] that the producer acquires before issuing the copy.
                    __threadfence_block();
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&pushed[op.c]);
            } else {  // OP_GEMM_LEAF / OP_GEMM_POP: multiply by the staged leaf message / parked partial
                const uint32_t st = mq % M_STAGES;
                mbar_wait(&mfull[st], (mq / M_STAGES) & 1);
                TL_MARK(oi * 8 + 3);
                if (warp_active) {
                    const double2* m = reinterpret_cast<const double2*>(Mring + st * STACK_LEVEL_BYTES + cw * STACK_ENTRY_BYTES) + lane;
#pragma unroll
                    for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const double2 v = m[(T * 8 + j) * 32];
                            cur[T][j][0] = acc[T][j][0] * v.x;
                            cur[T][j][1] = acc[T][j][1] * v.y;
                        }
                    if (RESCALE) {
                        if (op.kind == OP_GEMM_POP) {
                            __syncwarp();  // lane t == 0 of the quad wrote the exponents at the push
#pragma unroll
                            for (int T = 0; T < PRUNE_T; T++) esum[T] += gexp[(size_t)op.c * TILE_COLS + 8 * T];
                        }
                        rescale_columns(cur, esum);
                    }
                }
                release_stage(&mempty[st], lane);
                mq++;
            }
            TL_MARK(oi * 8 + 4);
        }
    }
}

// =================================================================================================
// K2+K3, wide form: three compute warps per SM sub-partition
// =================================================================================================
// Two warps inside the contraction saturate a sub-partition's DMMA pipe (98.6 %); one alone reaches 84-89 %
// (it has too few independent accumulator chains in flight). With two compute warps per sub-partition the
// pipe therefore idles whenever either of them is outside the contraction (barrier round trips, multiplicand
// pass, cherries, pushes: ~14 % of an edge). This form keeps THREE compute warps per sub-partition (12 per
// CTA, 192-column tiles), which needs the compute warps down at 160 registers and the shared memory freed
// of the 2 x 64 KB multiplicand ring. Both come from the same change: nothing is staged per column any more.
//   - A leaf message is read straight out of the leaf's P^T table (65 rows x 512 B, one TMA bulk copy per
//     leaf into a 2-stage ring) with the column's code as the row index: 16 LDS.128 per thread, into the
//     registers of the partials the contraction has just consumed.
//   - A parked partial is written to and read back from the CTA's global (L2-resident) stack by the very
//     thread that owns it, so parking needs no barrier, fence or staging at all.
//   - The producer warpgroup shrinks to one thread issuing TMA copies (P images into a 3-stage ring, leaf
//     tables) and three warps that load the next tile's leaf codes into a double buffer.
// A tile of <= 128 columns keeps warps 8..11 idle and costs what it costs in the narrow form.
constexpr int W_WARPS = 12;
constexpr int W_THREADS = 128 + W_WARPS * 32;                   // 512: launch allocation 128 registers per thread
constexpr int W_TILE_COLS = W_WARPS * 16;                       // 192
constexpr int W_LEVEL_BYTES = W_WARPS * STACK_ENTRY_BYTES;      // one CTA's parked partial: 96 KB
constexpr int W_P_STAGES = 3;
constexpr int W_L_STAGES = 2;
constexpr int W_BAR_BYTES = 128;                                // pfull[3] pempty[3] lfull[2] lempty[2] cfull[2] cempty[2]
constexpr int W_CODE_THREADS = 96;                              // warps 1..3 of the producer warpgroup
constexpr int W_PROD_REGS = 32;                                 // (128 - 32) * 128 released = (160 - 128) * 384 taken
constexpr int W_COMPUTE_REGS = 160;
static_assert(PRUNE_T == 2, "the wide form is written for two 8-column tiles per warp");

__device__ __forceinline__ int w_codes_bytes(int n_leaves) { return (W_TILE_COLS * n_leaves + 15) & ~15; }

#ifndef PCSF_TABLE_PREFETCH
#define PCSF_TABLE_PREFETCH 1
#endif
template <bool RESCALE>
__global__ void __launch_bounds__(W_THREADS, 1) prune_wide_kernel(const PruneParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* Pring = smem;
    uint8_t* Lring = smem + W_P_STAGES * FRAG_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Lring + W_L_STAGES * PT_SLOT_BYTES);
    uint64_t* pfull = bars;
    uint64_t* pempty = bars + 3;
    uint64_t* lfull = bars + 6;
    uint64_t* lempty = bars + 8;
    uint64_t* cfull = bars + 10;
    uint64_t* cempty = bars + 12;
    Op* ops_s = reinterpret_cast<Op*>(reinterpret_cast<uint8_t*>(bars) + W_BAR_BYTES);
    Item* items_s = reinterpret_cast<Item*>(reinterpret_cast<uint8_t*>(ops_s) + p.ops_bytes);
    uint8_t* codes_s = reinterpret_cast<uint8_t*>(items_s) + p.items_bytes;  // two buffers of w_codes_bytes
    const int cbytes = w_codes_bytes(p.n_leaves);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < W_P_STAGES; i++) {
            mbar_init(&pfull[i], 1);
            mbar_init(&pempty[i], W_WARPS);
        }
        for (int i = 0; i < W_L_STAGES; i++) {
            mbar_init(&lfull[i], 1);
            mbar_init(&lempty[i], W_WARPS);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&cfull[i], W_CODE_THREADS);
            mbar_init(&cempty[i], W_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = tid; i < p.n_ops; i += W_THREADS) ops_s[i] = p.ops[i];
    for (int i = tid; i < p.n_items; i += W_THREADS) items_s[i] = p.items[i];
    __syncthreads();

    if (tid < 128) {
        // ====================================== producer warpgroup ======================================
        reg_dealloc<W_PROD_REGS>();
        if (tid == 0) {  // the TMA thread: P images and leaf tables, in program order, as far ahead as the rings allow
            uint32_t pq = 0, lq = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const Span sp = p.spans[find_span(p.spans, p.n_spans, tile)];
                const double* tables = p.psets[sp.pset].tables;
                for (int ii = 0; ii < p.n_items; ii++) {
                    const Item it = items_s[ii];
                    if (it.kind == ITEM_P) {
                        const uint32_t st = pq % W_P_STAGES, u = pq / W_P_STAGES;
                        if (u >= 1) mbar_wait(&pempty[st], (u - 1) & 1);
                        mbar_expect_tx(&pfull[st], FRAG_BYTES);
                        tma_bulk_g2s(Pring + st * FRAG_BYTES, tables + (size_t)it.a * PT_SLOT, FRAG_BYTES, &pfull[st]);
                        pq++;
                    } else if (it.kind == ITEM_LEAF) {
                        const uint32_t st = lq % W_L_STAGES, u = lq / W_L_STAGES;
                        if (u >= 1) mbar_wait(&lempty[st], (u - 1) & 1);
                        mbar_expect_tx(&lfull[st], PT_SLOT_BYTES);
                        tma_bulk_g2s(Lring + st * PT_SLOT_BYTES, tables + (size_t)it.a * PT_SLOT, PT_SLOT_BYTES, &lfull[st]);
                        lq++;
                    }  // ITEM_POP: nothing to stage, the owning thread reads its parked partial back itself
                }
            }
        } else if (tid >= 32) {  // leaf codes of the CTA's k-th tile into buffer k % 2 (columns past the end of the span marginalise)
            const int ctid = tid - 32;
            uint32_t k = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, k++) {
                const Span sp = p.spans[find_span(p.spans, p.n_spans, tile)];
                const int64_t tcol0 = (tile - sp.tile0) * W_TILE_COLS;
                const int ncols = (int)min((int64_t)W_TILE_COLS, (int64_t)sp.ncols - tcol0);
                if (k >= 2) mbar_wait(&cempty[k & 1], ((k >> 1) - 1) & 1);
                uint8_t* dst = codes_s + (k & 1) * cbytes;
                const uint8_t* src = p.codes + (size_t)(sp.col0 + tcol0) * p.n_leaves;
                const int nbytes = ncols * p.n_leaves, total = W_TILE_COLS * p.n_leaves;
                for (int i = ctid; i < total; i += W_CODE_THREADS) dst[i] = (i < nbytes) ? src[i] : (uint8_t)64;
                mbar_arrive(&cfull[k & 1]);
            }
        }
        return;
    }

    // ========================================= compute warps =========================================
    // 160 registers hold 128 of partials and accumulators; everything that is needed once per tile only
    // (where the tile's results go, which prior to use) lives in a per-warp scratch line of shared memory,
    // and addresses are kept as 32-bit shared-memory offsets, so that the loop scalars stay in registers.
    reg_alloc<W_COMPUTE_REGS>();
    const int ctid = tid - 128, lane = ctid & 31, cw = ctid >> 5;
    const int g = lane >> 2, t = lane & 3;
    uint32_t pq = 0, lq = 0;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t lring0 = smem0 + W_P_STAGES * FRAG_BYTES + t * 16;
    const uint32_t codes00 = smem_u32(codes_s) + (cw * WARP_COLS + g) * p.n_leaves;  // this lane's first column in buffer 0
    int64_t* const scratch = reinterpret_cast<int64_t*>(codes_s + 2 * cbytes) + cw * 4;  // {out index of column 0 of the tile, pset, cherry tables}
    // No start-up offsets and no barrier among the warps of a sub-partition: left alone they spread out over
    // the edge by themselves (timeline: thousands of clocks apart). Both were measured: offsets of 0.8-3 us
    // change nothing, re-aligning the three warps after every contraction costs 6 points (83.3 %).
#ifdef PCSF_TIMELINE
    int tl_n = 0;
#endif

    for (uint32_t k = 0; (int64_t)blockIdx.x + (int64_t)k * gridDim.x < p.n_tiles; k++) {
        bool warp_active;
        int ncols;
        {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
            const Span sp = p.spans[find_span(p.spans, p.n_spans, tile)];
            const int64_t tcol0 = (tile - sp.tile0) * W_TILE_COLS;  // first column of the tile within the span
            ncols = (int)min((int64_t)W_TILE_COLS, (int64_t)sp.ncols - tcol0);
            warp_active = cw * WARP_COLS < ncols;
            if (lane == 0) {
                scratch[0] = sp.out0 + tcol0;
                scratch[1] = sp.pset;
                scratch[2] = (int64_t)p.psets[sp.pset].cherry;
            }
            __syncwarp();
        }
        mbar_wait(&cfull[k & 1], (k >> 1) & 1);
        const uint32_t codes0 = codes00 + (k & 1) * cbytes;  // code row of this lane's column g (column 8+g: + 8 n_leaves)

        double cur[2][8][2];
        int esum[2] = {0, 0};  // power-of-two exponent taken out of each column so far (rescale option)
#pragma unroll
        for (int T = 0; T < 2; T++)
#pragma unroll
            for (int j = 0; j < 8; j++) cur[T][j][0] = cur[T][j][1] = 1.0;

        // leaf message of leaf `leaf` for this lane's columns, out of the staged P^T table: row = code, 16 B at state 8j+2t
        auto leaf_rows = [&](uint32_t st, int leaf, uint32_t& a0, uint32_t& a1) {
            uint32_t c0 = lds_u8(codes0 + leaf), c1 = lds_u8(codes0 + 8 * p.n_leaves + leaf);
            c0 = c0 > 64 ? 64 : c0;
            c1 = c1 > 64 ? 64 : c1;
            const uint32_t base = lring0 + st * PT_SLOT_BYTES;
            a0 = base + c0 * 512;
            a1 = base + c1 * 512;
        };
        // rows of this lane's two columns in the table of a table op: the codes of the subtree's leaves, base 65
        auto table_rows = [&](const Op& op, uint32_t& row0, uint32_t& row1) {
            auto code = [&](uint32_t leaf, int col8) {
                const uint32_t c = lds_u8(codes0 + col8 * p.n_leaves + leaf);
                return c > 64 ? 64u : c;
            };
            const uint32_t la = op.a & 0xffff, lb = (uint32_t)op.a >> 16;
            row0 = code(la, 0) * 65 + code(lb, 0);
            row1 = code(la, 8) * 65 + code(lb, 8);
            const uint32_t lc = op.b & 0xffff, ld = (uint32_t)op.b >> 16;
            if (lc != 0xffff) {
                row0 = row0 * 65 + code(lc, 0);
                row1 = row1 * 65 + code(lc, 8);
                if (ld != 0xffff) {
                    row0 = row0 * 65 + code(ld, 0);
                    row1 = row1 * 65 + code(ld, 8);
                }
            }
        };
#if PCSF_TABLE_PREFETCH
        // The large tables (3 and 4 leaves) do not stay in L2 as a whole: ask for this tile's rows now, thousands of
        // clocks before the ops that gather them (lane t of a quad asks for the t-th 128-byte line of the row).
        if (warp_active)
            for (int oi = 0; oi < p.n_ops; oi++) {
                const Op op = ops_s[oi];
                if (!op_is_table(op.kind) || (op.b & 0xffff) == 0xffff) continue;
                const double* W = reinterpret_cast<const double*>(scratch[2]) + p.tab_off[op.kind >> 8] + 16 * t;
                uint32_t row0, row1;
                table_rows(op, row0, row1);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(W + (size_t)row0 * 64));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(W + (size_t)row1 * 64));
            }
#endif
        auto park = [&](int level) {  // this thread's 256 bytes of stack level `level`
            return p.global_stack + ((size_t)blockIdx.x * p.n_levels + level) * W_LEVEL_BYTES + (size_t)cw * STACK_ENTRY_BYTES + lane * 16;
        };
        auto park_exp = [&](int level) {
            return p.global_exp + ((size_t)blockIdx.x * p.n_levels + level) * W_TILE_COLS + cw * WARP_COLS + g;
        };

        for (int oi = 0; oi < p.n_ops; oi++) {
            const Op op = ops_s[oi];
            if (op.kind == OP_CHERRY) {
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    const uint32_t st = lq % W_L_STAGES;
                    mbar_wait(&lfull[st], (lq / W_L_STAGES) & 1);
                    if (warp_active) {
                        uint32_t a0, a1;
                        leaf_rows(st, kk ? op.b : op.a, a0, a1);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const double2 v0 = lds_f64x2(a0 + j * 64), v1 = lds_f64x2(a1 + j * 64);
                            cur[0][j][0] = kk ? cur[0][j][0] * v0.x : v0.x;
                            cur[0][j][1] = kk ? cur[0][j][1] * v0.y : v0.y;
                            cur[1][j][0] = kk ? cur[1][j][0] * v1.x : v1.x;
                            cur[1][j][1] = kk ? cur[1][j][1] * v1.y : v1.y;
                        }
                    }
                    release_stage(&lempty[st], lane);
                    lq++;
                }
                continue;
            }
            if (op.kind == OP_ROOT) {
                if (warp_active) {
                    const int64_t out0 = scratch[0];
                    const PSet ps = p.psets[scratch[1]];
                    const double2* pr = reinterpret_cast<const double2*>(ps.prior + 2 * t);
                    const double2* lp = reinterpret_cast<const double2*>(ps.logprior + 2 * t);
#pragma unroll
                    for (int T = 0; T < 2; T++) {
                        double zp = 0.0;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const double2 pj = __ldg(pr + 4 * j);
                            cur[T][j][0] *= pj.x;  // alpha_root[x] * prior[x]
                            cur[T][j][1] *= pj.y;
                            zp += cur[T][j][0];
                            zp += cur[T][j][1];
                        }
                        zp += __shfl_xor_sync(0xffffffffu, zp, 1);
                        zp += __shfl_xor_sync(0xffffffffu, zp, 2);
                        double ap = 0.0;
                        if (zp != 0.0) {  // PhyloLik.ml:131-132: impossible data => zero posterior
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const double2 lj = __ldg(lp + 4 * j);
                                ap += (cur[T][j][0] / zp) * lj.x;
                                ap += (cur[T][j][1] / zp) * lj.y;
                            }
                        }
                        ap += __shfl_xor_sync(0xffffffffu, ap, 1);
                        ap += __shfl_xor_sync(0xffffffffu, ap, 2);
                        const int c = cw * WARP_COLS + 8 * T + g;
                        if (t == 0 && c < ncols) {
                            p.out_logz[out0 + c] = (RESCALE && esum[T]) ? log(zp) + (double)esum[T] * 0.6931471805599453 : log(zp);
                            p.out_anc[out0 + c] = ap;
                        }
                    }
                }
                continue;
            }
            double acc[2][8][2];
            int ekind = op.kind;
            if (op_is_table(op.kind)) {
                // -------------------- cherry + the edge above it: one 512-byte row of the cherry's table per column --------------------
                const int tk = op.kind & 0xff;  // LEAF / PUSH / POP epilogue as after a contraction; KEEP / MUL: see OpKind
                ekind = tk == OP_TAB_KEEP ? OP_GEMM_KEEP : tk == OP_TAB_MUL ? OP_TAB_MUL : tk - OP_TAB_LEAF + OP_GEMM_LEAF;
                if (warp_active) {
                    const double* W = reinterpret_cast<const double*>(scratch[2]) + p.tab_off[op.kind >> 8] + 2 * t;
                    uint32_t row0, row1;
                    table_rows(op, row0, row1);
                    const double2* r0 = reinterpret_cast<const double2*>(W + (size_t)row0 * 64);
                    const double2* r1 = reinterpret_cast<const double2*>(W + (size_t)row1 * 64);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double2 v0 = __ldg(r0 + 4 * j), v1 = __ldg(r1 + 4 * j);
                        acc[0][j][0] = v0.x; acc[0][j][1] = v0.y;
                        acc[1][j][0] = v1.x; acc[1][j][1] = v1.y;
                    }
                }
            } else {
            // ------------------------------ K3: contraction over one internal edge ------------------------------
            TL_MARK(oi * 8 + 0);
            const uint32_t pst = pq % W_P_STAGES;
            mbar_wait(&pfull[pst], (pq / W_P_STAGES) & 1);
            TL_MARK(oi * 8 + 1);
            if (warp_active) {
                const double* Pb = reinterpret_cast<const double*>(Pring + pst * FRAG_BYTES) + lane;
#pragma unroll
                for (int T = 0; T < 2; T++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[T][j][0] = acc[T][j][1] = 0.0;
#pragma unroll
                for (int s = 0; s < 16; s++) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double bf = Pb[(j * 16 + s) * 32];
#pragma unroll
                        for (int T = 0; T < 2; T++) dmma(acc[T][j][0], acc[T][j][1], cur[T][s >> 1][s & 1], bf);
                    }
                }
            }
            release_stage(&pempty[pst], lane);
            pq++;
            TL_MARK(oi * 8 + 2);
            }
            // ------------------------------ epilogue ------------------------------
            if (ekind == OP_GEMM_KEEP) {  // the sibling is the next op's lookup: the message stays where it is
                if (warp_active) {
#pragma unroll
                    for (int T = 0; T < 2; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            cur[T][j][0] = acc[T][j][0];
                            cur[T][j][1] = acc[T][j][1];
                        }
                }
            } else if (ekind == OP_TAB_MUL) {  // ... and the lookup multiplies into it
                if (warp_active) {
#pragma unroll
                    for (int T = 0; T < 2; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            cur[T][j][0] = acc[T][j][0] * cur[T][j][0];
                            cur[T][j][1] = acc[T][j][1] * cur[T][j][1];
                        }
                    if (RESCALE) rescale_columns(cur, esum);
                }
            } else if (ekind == OP_GEMM_PUSH) {  // park the message; only this thread ever touches these addresses
                if (warp_active) {
                    double2* slot = reinterpret_cast<double2*>(park(op.c));
#pragma unroll
                    for (int T = 0; T < 2; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            slot[(T * 8 + j) * 32] = make_double2(acc[T][j][0], acc[T][j][1]);
                        }
                    if (RESCALE) {  // the parked message keeps its exponent; the sibling subtree starts at 0
#pragma unroll
                        for (int T = 0; T < 2; T++) {
                            if (t == 0) park_exp(op.c)[8 * T] = esum[T];
                            esum[T] = 0;
                        }
                    }
                }
            } else if (ekind == OP_GEMM_POP) {
                if (warp_active) {
                    const double2* slot = reinterpret_cast<const double2*>(park(op.c));
#pragma unroll
                    for (int T = 0; T < 2; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const double2 v = slot[(T * 8 + j) * 32];
                            cur[T][j][0] = acc[T][j][0] * v.x;
                            cur[T][j][1] = acc[T][j][1] * v.y;
                        }
                    if (RESCALE) {
                        __syncwarp();  // lane t == 0 of the quad wrote the exponents at the push
#pragma unroll
                        for (int T = 0; T < 2; T++) esum[T] += park_exp(op.c)[8 * T];
                        rescale_columns(cur, esum);
                    }
                }
            } else {  // OP_GEMM_LEAF
                const uint32_t st = lq % W_L_STAGES;
                mbar_wait(&lfull[st], (lq / W_L_STAGES) & 1);
                TL_MARK(oi * 8 + 3);
                if (warp_active) {
                    uint32_t a0, a1;
                    leaf_rows(st, ekind == op.kind ? op.b : op.c, a0, a1);  // the sibling leaf
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double2 v0 = lds_f64x2(a0 + j * 64), v1 = lds_f64x2(a1 + j * 64);
                        cur[0][j][0] = acc[0][j][0] * v0.x;
                        cur[0][j][1] = acc[0][j][1] * v0.y;
                        cur[1][j][0] = acc[1][j][0] * v1.x;
                        cur[1][j][1] = acc[1][j][1] * v1.y;
                    }
                    if (RESCALE) rescale_columns(cur, esum);
                }
                release_stage(&lempty[st], lane);
                lq++;
            }
            TL_MARK(oi * 8 + 4);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cempty[k & 1]);  // this warp is done with the tile's codes
    }
}

// =================================================================================================
// Subtree tables. For a cherry (leaves a, b under node v) the partial likelihood that leaves the edge above v,
// W2[code_a][code_b][y] = sum_x P_v[y][x] (G_a[code_a][x] G_b[code_b][x]), depends on the column only through the code
// pair; if the cherry's sibling is a leaf c (parent u), W3[code_a][code_b][code_c] = P_u x (W2[code_a][code_b] * G_c[code_c])
// only through the triple, and so on for a further leaf d. They are memoised over all 65^k code tuples with the very instruction sequence of
// the pruning kernels (same products, same fragment-ordered image, same DMMA accumulation order), 16 tuples per warp as
// if they were 16 columns, so a lookup is bit-identical to the computation it replaces.
// =================================================================================================
__global__ void __launch_bounds__(128) subtree_table_kernel(const double* __restrict__ tables, const SubTab* __restrict__ tabs,
                                                            int first, int count, int rows, double* __restrict__ base) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warps_per_table = (rows + 15) / 16;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)count * warps_per_table) return;
    const SubTab tb = tabs[first + (int)(wid / warps_per_table)];
    const int q0 = (int)(wid % warps_per_table) * 16;
    const double* Pb = tables + (size_t)tb.edge * PT_SLOT + lane;
    double cur[2][8][2], acc[2][8][2];
#pragma unroll
    for (int T = 0; T < 2; T++) {
        const int q = min(q0 + 8 * T + g, rows - 1);  // padding tuples repeat the last one (never stored)
        if (tb.lnew < 0) {  // cherry: the product of the two leaf messages, as in the pruning kernels
            const double2* ra = reinterpret_cast<const double2*>(tables + (size_t)tb.la * PT_SLOT + (q / 65) * 64 + 2 * t);
            const double2* rb = reinterpret_cast<const double2*>(tables + (size_t)tb.lb * PT_SLOT + (q % 65) * 64 + 2 * t);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double2 u = ra[4 * j], v = rb[4 * j];
                cur[T][j][0] = u.x * v.x;
                cur[T][j][1] = u.y * v.y;
            }
        } else {  // subtree + leaf: the source table's row times the leaf message, as in the GEMM_LEAF epilogue
            const double2* rw = reinterpret_cast<const double2*>(base + tabs[tb.src].off + (size_t)(q / 65) * 64 + 2 * t);
            const double2* rc = reinterpret_cast<const double2*>(tables + (size_t)tb.lnew * PT_SLOT + (q % 65) * 64 + 2 * t);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double2 w = rw[4 * j], v = rc[4 * j];
                cur[T][j][0] = w.x * v.x;
                cur[T][j][1] = w.y * v.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc[T][j][0] = acc[T][j][1] = 0.0;
    }
#pragma unroll
    for (int s = 0; s < 16; s++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double bf = Pb[(j * 16 + s) * 32];
#pragma unroll
            for (int T = 0; T < 2; T++) dmma(acc[T][j][0], acc[T][j][1], cur[T][s >> 1][s & 1], bf);
        }
    }
#pragma unroll
    for (int T = 0; T < 2; T++) {
        const int q = q0 + 8 * T + g;
        if (q < rows) {
            double2* w = reinterpret_cast<double2*>(base + tb.off + (size_t)q * 64 + 2 * t);
#pragma unroll
            for (int j = 0; j < 8; j++) w[4 * j] = make_double2(acc[T][j][0], acc[T][j][1]);
        }
    }
}

// =================================================================================================
// K4: per-region sums of the per-column terms (deterministic order: lane-strided partial sums,
// then a fixed xor tree).  One warp per segment.
// =================================================================================================
__global__ void region_reduce_kernel(const double* __restrict__ col_logz, const double* __restrict__ col_anc,
                                     const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_end,
                                     int64_t n_segs, double* __restrict__ lpr, double* __restrict__ elpr) {
    const int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (seg >= n_segs) return;
    const int64_t b = seg_begin[seg], e = seg_end[seg];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = b + lane; i < e; i += 32) {
        s0 += col_logz[i];
        s1 += col_anc[i];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
        lpr[seg] = s0;
        elpr[seg] = s1;
    }
}

// Segments of "every region under the m-th listed model": slot m*total_cols + region_off[r].
__global__ void make_segments_kernel(const int64_t* __restrict__ region_off, int64_t nregions, int n_models,
                                     int64_t total_cols, int64_t* __restrict__ seg_begin,
                                     int64_t* __restrict__ seg_end) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nregions * n_models) return;
    const int64_t m = i / nregions, r = i - m * nregions;
    seg_begin[i] = m * total_cols + region_off[r];
    seg_end[i] = m * total_cols + region_off[r + 1];
}

// =================================================================================================
// K5: omega-model rate matrices and their diagonalisation, one CTA per (region, candidate)
// =================================================================================================
// Replaces, for the omega strategy, what the reference does on the host for every kappa candidate of
// every region (src/OmegaModel.ml:166-170 -> PhyloModel.P14n.update ~q_settings -> Q.Diag.of_Q):
//   1. assemble Q(kappa, omega, sigma, F3x4) in the evaluation order of the reference's Expr trees
//      (OmegaModel.ml:21-80; __d*_rn intrinsics keep nvcc from contracting into FMAs, so Q is
//      bit-identical to the host assembly),
//   2. symmetrise with the codon frequencies (the model is reversible) and run a parallel-ordered
//      cyclic Jacobi (round-robin pairing, 32 disjoint rotations per round, 63 rounds per sweep),
//   3. write S | S^-1 | lambda | equilibrium prior | log prior straight into the model block K1 reads.
__device__ __forceinline__ bool omega_is_stop(int c) { return c == 48 || c == 50 || c == 56; }
__constant__ char kOmegaAA[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";

constexpr int EIG_THREADS = 256;
// Warm start (cache_slot[i] >= 0 and the slot holds eigenvectors of an earlier solve): the kappa search of a region
// visits a sequence of nearby rate matrices, so the previous candidate's eigenvectors V0 nearly diagonalise the new
// matrix; the sweeps then run on V0^T A V0 with the rotations accumulated onto V0 and converge in 2-4 sweeps instead
// of ~9. The result is an eigensystem of the same matrix to the same stopping threshold (it differs from a cold
// solve by rounding, like any two eigensolvers do); it depends only on the region's own sequence of candidates.
__global__ void __launch_bounds__(EIG_THREADS) omega_eig_kernel(const double* __restrict__ qs_all, double* __restrict__ params_all,
                                                                  int32_t* __restrict__ status, const int64_t* __restrict__ cache_slot,
                                                                  double* __restrict__ cache, int32_t* __restrict__ cache_valid,
                                                                  int32_t* __restrict__ sweeps_out) {
    extern __shared__ double esm[];
    double* A = esm;              // 64 x 65 (padded)
    double* V = esm + 64 * 65;    // 64 x 65
    double* pi = V + 64 * 65;     // 64
    double* sw = pi + 64;         // 64
    double* rc = sw + 64;         // 32 cosines
    double* rs = rc + 32;         // 32 sines
    __shared__ int pp[32], pq[32];
    __shared__ double red[EIG_THREADS / 32], red2[EIG_THREADS / 32];
    __shared__ double s_factor, s_off, s_diag;
    const int tid = threadIdx.x;
    const double* v = qs_all + (size_t)blockIdx.x * 12;
    double* out = params_all + (size_t)blockIdx.x * (8192 + 192);
    const double kappa = v[0], omega = v[1];
    // ---- codon frequencies, OmegaModel.ml:24-42 ----
    if (tid < 64) {
        auto sc = [&](int i1, int i2, int i3) {
            const double f1 = __ddiv_rn(i1 == 3 ? 1.0 : v[3 + i1], __dadd_rn(v[3], __dadd_rn(v[4], __dadd_rn(v[5], 1.0))));
            const double f2 = __ddiv_rn(i2 == 3 ? 1.0 : v[6 + i2], __dadd_rn(v[6], __dadd_rn(v[7], __dadd_rn(v[8], 1.0))));
            const double f3 = __ddiv_rn(i3 == 3 ? 1.0 : v[9 + i3], __dadd_rn(v[9], __dadd_rn(v[10], __dadd_rn(v[11], 1.0))));
            return __dmul_rn(f1, __dmul_rn(f2, f3));
        };
        const double denom = __dsub_rn(1.0, __dmul_rn(__dsub_rn(1.0, v[2]), __dadd_rn(sc(3, 0, 0), __dadd_rn(sc(3, 0, 2), sc(3, 2, 0)))));
        pi[tid] = __ddiv_rn(sc(tid >> 4, (tid >> 2) & 3, tid & 3), denom);
    }
    __syncthreads();
    // ---- off-diagonal rates, OmegaModel.ml:44-72 ----
    for (int idx = tid; idx < 4096; idx += EIG_THREADS) {
        const int i = idx >> 6, j = idx & 63;
        const int ii[3] = {i >> 4, (i >> 2) & 3, i & 3}, jj[3] = {j >> 4, (j >> 2) & 3, j & 3};
        int nd = 0, da = 0, dbb = 0;
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (ii[k] != jj[k]) { nd++; da = ii[k]; dbb = jj[k]; }
        double q = 0.0;
        if (nd == 1) {
            const bool transition = (da == 0 && dbb == 2) || (da == 2 && dbb == 0) || (da == 1 && dbb == 3) || (da == 3 && dbb == 1);
            const double kp = transition ? kappa : 1.0;
            const double op = (!omega_is_stop(i) && !omega_is_stop(j) && kOmegaAA[i] != kOmegaAA[j]) ? omega : 1.0;
            q = __dmul_rn(pi[j], __dmul_rn(kp, op));
        }
        A[i * 65 + j] = q;
    }
    __syncthreads();
    // ---- diagonal and unit-rate scale, PhyloModel.ml:76-84,94-104 / OmegaModel.ml:76-80 ----
    if (tid < 64) {
        double tot = 0.0;
        for (int j = 0; j < 64; j++)
            if (j != tid) tot = __dadd_rn(A[tid * 65 + j], tot);
        A[tid * 65 + tid] = __dsub_rn(0.0, tot);
    }
    __syncthreads();
    if (tid == 0) {
        double factor = 0.0;
        for (int i = 0; i < 64; i++) factor = __dsub_rn(factor, __dmul_rn(pi[i], A[i * 65 + i]));
        s_factor = factor;
    }
    __syncthreads();
    const double factor = s_factor;
    int st = 0;
    if (!(factor > 0.0)) st |= 128;  // "Q scale evaluated to a non-positive value"
    if (tid < 64) {
        if (!(pi[tid] > 0.0)) st |= 128;
        sw[tid] = sqrt(pi[tid]);
    }
    __syncthreads();
    // ---- scaled Q, symmetrised: A = W^1/2 Q W^-1/2 ----
    for (int idx = tid; idx < 4096; idx += EIG_THREADS) {
        const int i = idx >> 6, j = idx & 63;
        V[i * 65 + j] = sw[i] * __ddiv_rn(A[i * 65 + j], factor) / sw[j];
    }
    __syncthreads();
    for (int idx = tid; idx < 4096; idx += EIG_THREADS) {
        const int i = idx >> 6, j = idx & 63;
        A[i * 65 + j] = (i == j) ? V[i * 65 + i] : 0.5 * (V[i * 65 + j] + V[j * 65 + i]);
    }
    __syncthreads();
    const int64_t slot = cache_slot ? cache_slot[blockIdx.x] : -1;
    const bool warm = slot >= 0 && cache_valid[slot] != 0;  // uniform across the CTA
    if (!warm) {
        for (int idx = tid; idx < 4096; idx += EIG_THREADS) V[(idx >> 6) * 65 + (idx & 63)] = ((idx >> 6) == (idx & 63)) ? 1.0 : 0.0;
        __syncthreads();
    } else {
        // A <- V0^T A V0, V <- V0. The intermediate A V0 goes through this model's own output block (global, L2;
        // overwritten with S at the end), so the kernel keeps its two shared-memory matrices and three CTAs per SM.
        const double* V0 = cache + (size_t)slot * 4096;
        double* Tg = out;
        for (int idx = tid; idx < 4096; idx += EIG_THREADS) V[(idx >> 6) * 65 + (idx & 63)] = V0[idx];
        __syncthreads();
        for (int idx = tid; idx < 4096; idx += EIG_THREADS) {
            const int i = idx >> 6, j = idx & 63;
            double acc = 0.0;
#pragma unroll 4
            for (int k = 0; k < 64; k++) acc += A[i * 65 + k] * V[k * 65 + j];
            Tg[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < 4096; idx += EIG_THREADS) {
            const int i = idx >> 6, j = idx & 63;
            double acc = 0.0;
#pragma unroll 4
            for (int k = 0; k < 64; k++) acc += V[k * 65 + i] * Tg[k * 64 + j];
            A[i * 65 + j] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < 4096; idx += EIG_THREADS) {  // exactly symmetric again
            const int i = idx >> 6, j = idx & 63;
            if (i < j) {
                const double m = 0.5 * (A[i * 65 + j] + A[j * 65 + i]);
                A[i * 65 + j] = m;
                A[j * 65 + i] = m;
            }
        }
        __syncthreads();
    }
    // ---- parallel cyclic Jacobi ----
    int sweeps = 0;
    for (int sweep = 0; sweep < 40; sweep++) {
        double o2 = 0.0, d2 = 0.0;
        for (int idx = tid; idx < 4096; idx += EIG_THREADS) {
            const int i = idx >> 6, j = idx & 63;
            const double a = A[i * 65 + j];
            if (i == j) d2 += a * a; else o2 += a * a;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { o2 += __shfl_xor_sync(0xffffffffu, o2, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o); }
        if ((tid & 31) == 0) { red[tid >> 5] = o2; red2[tid >> 5] = d2; }
        __syncthreads();
        if (tid == 0) {
            double t = 0, u = 0;
            for (int k = 0; k < EIG_THREADS / 32; k++) { t += red[k]; u += red2[k]; }
            s_off = t;
            s_diag = u;
        }
        __syncthreads();
        // Stop at an off-diagonal norm of 1e-15 of the diagonal's (squares compared). The residual E is a perturbation
        // of the MATRIX (A = V (L + E) V^T holds exactly), so it moves P(t) = exp(tA) by <= t |E| ~ 1e-15 t whatever the
        // eigenvalue gaps are - the level every P(t) entry carries anyway. The convergence is
        // quadratic (1e-8 -> 1e-16 -> 1e-32 per sweep): round 1's 1e-17 bought nothing and cost most matrices a sweep.
        if (s_off <= 1e-30 * s_diag || s_off == 0.0) break;
        sweeps++;
        for (int r = 0; r < 63; r++) {
            if (tid < 32) {  // pairing of round r and its rotations
                // Pair t = ((r + t) mod 63, (r - t) mod 63), pair 0 = (63, r) - kept in that order, NOT sorted: over the lanes
                // of a warp the first members are then consecutive columns and the second members consecutive descending ones,
                // so the lane-indexed accesses A[row][p(lane)], V[row][q(lane)] below fall into distinct banks (row stride 65).
                // Sorted pairs scattered them (~3-way conflicts on every access), and this kernel is bound by shared-memory
                // wavefronts: ncu had the pipe 50-65 % busy with ONE CTA per SM, which is why three CTAs per SM bought so little.
                const int p_ = tid == 0 ? 63 : (r + tid) % 63, q_ = tid == 0 ? r : (r - tid + 63) % 63;
                pp[tid] = p_;
                pq[tid] = q_;
                const double apq = A[p_ * 65 + q_], app = A[p_ * 65 + p_], aqq = A[q_ * 65 + q_];
                double c = 1.0, s = 0.0;
                // the second test: after a warm start most pairs are already negligible in the first sweeps
                if (apq != 0.0 && !((sweep > 4 || warm) && fabs(apq) <= 1e-20 * fabs(app) && fabs(apq) <= 1e-20 * fabs(aqq))) {
                    // The rotation's angle phi (|phi| <= pi/4, tan 2 phi = 2 a_pq / (a_qq - a_pp)) from two reciprocal square
                    // roots instead of the textbook's three divisions and two square roots: the other seven warps of the CTA wait for
                    // this dependent chain 63 times per sweep.
                    //   cos 2phi = |d| / r, sin 2phi = +-b / r (r = hypot(d, b));  cos^2 phi = (1 + cos 2phi) / 2 =: h;
                    //   cos phi = h / sqrt(h), sin phi = sin 2phi / (2 cos phi).   c^2 + s^2 = 1 to rounding, like the textbook's.
                    const double d = aqq - app, b = 2.0 * apq;
                    const double r2 = d * d + b * b;
                    if (r2 > 1e-280 && r2 < 1e280) {
                        const double ir = rsqrt(r2);
                        const double c2 = fabs(d) * ir;
                        const double s2 = (d >= 0.0 ? b : -b) * ir;
                        const double h = 0.5 + 0.5 * c2;
                        const double ic = rsqrt(h);
                        c = h * ic;
                        s = 0.5 * s2 * ic;
                    } else {  // squares out of range: the textbook form, which only divides
                        const double theta = d / b;
                        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                }
                rc[tid] = c;
                rs[tid] = s;
            }
            __syncthreads();
            // A <- J^T A J in one pass: the 2 x 2 block (rows of pair ki, columns of pair kj) only needs the two pairs'
            // rotations and belongs to one thread; columns first, then rows, as two separate passes would do it.
            // A thread's column pair is the same in all its blocks and in its share of V (kj = lane): loaded once per round.
            const int lane = tid & 31;
            const int pj_ = pp[lane], qj_ = pq[lane];
            const double cj = rc[lane], sj = rs[lane];
#pragma unroll 2
            for (int ki = tid >> 5; ki < 32; ki += EIG_THREADS / 32) {
                const int pi_ = pp[ki], qi_ = pq[ki];
                const double ci = rc[ki], si = rs[ki];
                const double a_pp = A[pi_ * 65 + pj_], a_pq = A[pi_ * 65 + qj_], a_qp = A[qi_ * 65 + pj_], a_qq = A[qi_ * 65 + qj_];
                const double t_pp = cj * a_pp - sj * a_pq, t_pq = sj * a_pp + cj * a_pq;
                const double t_qp = cj * a_qp - sj * a_qq, t_qq = sj * a_qp + cj * a_qq;
                double n_pp = ci * t_pp - si * t_qp, n_qp = si * t_pp + ci * t_qp;
                double n_pq = ci * t_pq - si * t_qq, n_qq = si * t_pq + ci * t_qq;
                // the rotated pair itself ends up exactly zero (annihilated, or flushed when it no longer registers
                // against the diagonal), as in the host's Jacobi: without this the off-diagonal norm stalls at
                // rounding level above the stopping threshold and every matrix runs all 40 sweeps instead of ~9
                if (ki == lane) { n_pq = 0.0; n_qp = 0.0; }
                A[pi_ * 65 + pj_] = n_pp;
                A[pi_ * 65 + qj_] = n_pq;
                A[qi_ * 65 + pj_] = n_qp;
                A[qi_ * 65 + qj_] = n_qq;
            }
            // V <- V J (pair `lane` touches columns p, q of every row)
#pragma unroll 4
            for (int row = tid >> 5; row < 64; row += EIG_THREADS / 32) {
                const double vkp = V[row * 65 + pj_], vkq = V[row * 65 + qj_];
                V[row * 65 + pj_] = cj * vkp - sj * vkq;
                V[row * 65 + qj_] = sj * vkp + cj * vkq;
            }
            __syncthreads();
        }
    }
    if (sweeps_out && tid == 0) sweeps_out[blockIdx.x] = sweeps;
    if (slot >= 0) {  // this solve's eigenvectors start the region's next one
        double* V0 = cache + (size_t)slot * 4096;
        for (int idx = tid; idx < 4096; idx += EIG_THREADS) V0[idx] = V[(idx >> 6) * 65 + (idx & 63)];
        if (tid == 0) cache_valid[slot] = 1;
    }
    // ---- outputs: S = W^-1/2 U, S^-1 = U^T W^1/2, lambda, equilibrium (Q.ml:153-177) ----
    for (int idx = tid; idx < 4096; idx += EIG_THREADS) {
        const int i = idx >> 6, k = idx & 63;
        out[i * 64 + k] = V[i * 65 + k] / sw[i];
        out[4096 + k * 64 + i] = V[i * 65 + k] * sw[i];
    }
    if (tid < 64) out[8192 + tid] = A[tid * 65 + tid];
    __syncthreads();
    if (tid == 0) {
        int p_ = 0;
        double best = INFINITY;
        for (int i = 0; i < 64; i++)
            if (fabs(A[i * 65 + i]) < best) { best = fabs(A[i * 65 + i]); p_ = i; }
        if (best > 1e-6) st |= 256;  // "smallest-magnitude eigenvalue is unacceptably large"
        double mass = 0.0;
        for (int i = 0; i < 64; i++) mass += V[i * 65 + p_] * sw[i];
        for (int i = 0; i < 64; i++) {
            const double pr = V[i * 65 + p_] * sw[i] / mass;
            out[8192 + 64 + i] = pr;
            out[8192 + 128 + i] = log(pr);
        }
    }
    if (st) atomicOr(&status[blockIdx.x], st);
}

// =================================================================================================
// K6: outside algorithm, node posteriors and expected substitution counts
//     (lib/CamlPaml/PhyloLik.ml:96-180: ensure_beta, node_posterior, add_branch_posteriors)
// =================================================================================================
// Not on the scoring path of the command line: this is the E step PhyloEM-style ECM training needs.
// Two forms: outside_kernel (this one, plain FP64, PCSF_K6_PLAIN=1) and outside_dmma_kernel (below, what pcsf_posteriors runs).
// CTAs walk tiles of 64 codon columns. Per tile, with msg_i = P_i x alpha_i the message of node i to its parent
// (a leaf's message is a row of its P^T table, as in the pruning kernels):
//   inside   (i ascending, internal nodes): alpha_i = msg_lc * msg_rc; msg_i kept for the way down
//   root:    z = alpha_root . prior; beta_root = prior
//   outside  (i descending, internal nodes below the root): inter = beta_p * msg_sib; beta_i[b] = sum_a inter[a] P_i[a][b]
//   branch br: G_br[a][b] += sum over the tile's columns with z > 0 of (beta_p[a] msg_sib[a] / z) alpha_br[b];
//              the expected counts are P_br[a][b] * G_br[a][b] (the P factor does not depend on the column)
// alpha, msg and beta of the tile live in a per-CTA global scratch block (L2), G_br in a per-CTA accumulator block that
// outside_reduce_kernel sums over the CTAs in a fixed order: results do not depend on scheduling.
// Every 64 x 64 product runs on the plain FP64 pipe as Y[c][o] = sum_k X[c][k] L[k][o] with L staged in shared memory in
// the orientation the product needs (B200's plain-FP64 peak equals its DMMA peak, and this kernel is bound by the L2
// traffic of its scratch blocks, not by arithmetic).
constexpr int OUT_TC = 64;          // columns per tile
constexpr int OUT_THREADS = 256;
constexpr int OUT_XS = 65;          // padded row stride of the column-major operand tiles
struct OutsideParams {
    const double* tables;    // P set: [n_branches][PT_SLOT]
    const double* prior;
    const uint8_t* codes;    // [total_cols][n_leaves]
    int64_t total_cols;
    int n_leaves;
    const int32_t* children; // [2 * (n_leaves - 1)]
    const int32_t* parent;   // [2 n_leaves - 1]
    const int32_t* sibling;  // [2 n_leaves - 1]
    double* scratch;         // [grid][3 * n_internal][OUT_TC][64]: alpha | msg | beta of internal nodes
    double* gacc;            // [grid][n_branches][64][64] or null (no expected counts wanted)
    int n_post;              // node posteriors wanted for these nodes
    const int32_t* post_nodes;
    double* post_out;        // [n_post][total_cols][64]
    double* z_out;           // [total_cols] or null
    const double* pt_images; // [n_leaves - 2][4096]: transposed fragment images of the internal edges (DMMA form)
    const int32_t* steps;    // [2 n_leaves - 2][3]: (node, OUT_STEP_* flags, the first internal node at or after this step or -1):
                             // the outside pass of the DMMA form in depth-first order; children | parent | sibling | steps are one array
};
constexpr int OUT_STEP_LOADB = 1;   // beta of the node's parent is not in registers: read it back (the root's is the prior)
constexpr int OUT_STEP_STOREB = 2;  // park beta_i: a later step (or a node posterior) reads it
constexpr int OUT_STEP_CARRY = 4;   // the walk visits this node's children next: beta_i stays in registers

__device__ __forceinline__ double out_p_entry(const double* slot, bool leaf, int a, int b) {
    return leaf ? slot[b * 64 + a] : slot[frag_index(a, b)];  // leaf slots hold P^T, internal ones the fragment-ordered image
}

// Y[c][o] = sum_k X[c][k] * L[k][o] for the tile's 64 columns; thread -> columns 2 (tid / 8), + 1 and outputs
// 8 (tid % 8) .. + 7: two X loads and four 16-byte loads of L feed sixteen FMAs
__device__ __forceinline__ void out_product(const double* __restrict__ X, const double* __restrict__ L, double (&y)[2][8], int tid) {
    const int c0 = (tid >> 3) * 2, o0 = (tid & 7) * 8;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) y[i][j] = 0.0;
    const double* x0 = X + c0 * OUT_XS;
    const double* x1 = x0 + OUT_XS;
#pragma unroll 4
    for (int k = 0; k < 64; k++) {
        const double xa = x0[k], xb = x1[k];
        const double2* l = reinterpret_cast<const double2*>(L + k * 64 + o0);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double2 v = l[j];
            y[0][2 * j] = fma(xa, v.x, y[0][2 * j]);
            y[0][2 * j + 1] = fma(xa, v.y, y[0][2 * j + 1]);
            y[1][2 * j] = fma(xb, v.x, y[1][2 * j]);
            y[1][2 * j + 1] = fma(xb, v.y, y[1][2 * j + 1]);
        }
    }
}

__global__ void __launch_bounds__(OUT_THREADS, 2) outside_kernel(const OutsideParams p) {
    extern __shared__ __align__(16) double osm[];
    double* L = osm;                      // 64 x 64 operand matrix
    double* X = L + 4096;                 // OUT_TC x OUT_XS
    double* U = X + OUT_TC * OUT_XS;      // OUT_TC x OUT_XS
    double* zs = U + OUT_TC * OUT_XS;     // OUT_TC
    uint8_t* codes_s = reinterpret_cast<uint8_t*>(zs + OUT_TC);  // OUT_TC x n_leaves
    const int tid = threadIdx.x, nl = p.n_leaves, n = 2 * nl - 1, ni = nl - 1;
    const int c0 = (tid >> 3) * 2, o0 = (tid & 7) * 8;  // this thread's two columns and eight states
    double* sc = p.scratch + (size_t)blockIdx.x * 3 * ni * OUT_TC * 64;
    auto alpha_at = [&](int node) { return sc + ((size_t)(node - nl) * OUT_TC) * 64; };
    auto msg_at = [&](int node) { return sc + ((size_t)(ni + node - nl) * OUT_TC) * 64; };
    auto beta_at = [&](int node) { return sc + ((size_t)(2 * ni + node - nl) * OUT_TC) * 64; };
    double* gacc = p.gacc ? p.gacc + (size_t)blockIdx.x * (n - 1) * 4096 : nullptr;
    if (gacc)
        for (size_t i = tid; i < (size_t)(n - 1) * 4096; i += OUT_THREADS) gacc[i] = 0.0;
    // message of node `node` for this thread's (columns, 8 states): a leaf gathers its P^T row, an internal node reads msg
    auto load_msg = [&](int node, double (&m)[2][8]) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double* row;
            if (node < nl) {
                int code = codes_s[(c0 + i) * nl + node];
                code = code > 64 ? 64 : code;
                row = p.tables + (size_t)node * PT_SLOT + code * 64 + o0;
            } else {
                row = msg_at(node) + (c0 + i) * 64 + o0;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) m[i][j] = row[j];
        }
    };
    // L[k][o] = P[o][k] (inside pass) or P[k][o] (outside pass): the slot is read linearly (coalesced) and scattered
    auto stage_L = [&](int node, bool transposed) {
        const double* slot = p.tables + (size_t)node * PT_SLOT;
        const bool leaf = node < nl;
        for (int idx = tid; idx < 4096; idx += OUT_THREADS) {
            int a, b;
            if (leaf) {
                a = idx & 63;
                b = idx >> 6;
            } else {  // inverse of frag_index
                const int j = idx >> 9, s = (idx >> 5) & 15, ln = idx & 31;
                a = 8 * j + (ln >> 2);
                b = 8 * (s >> 1) + 2 * (ln & 3) + (s & 1);
            }
            L[transposed ? b * 64 + a : a * 64 + b] = slot[idx];
        }
    };
    const int64_t n_tiles = (p.total_cols + OUT_TC - 1) / OUT_TC;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t col0 = tile * OUT_TC;
        const int ncols = (int)min((int64_t)OUT_TC, p.total_cols - col0);
        __syncthreads();
        for (int i = tid; i < OUT_TC * nl; i += OUT_THREADS) codes_s[i] = i < ncols * nl ? p.codes[(size_t)col0 * nl + i] : (uint8_t)64;
        __syncthreads();
        // ---------------- inside ----------------
        for (int i = nl; i < n; i++) {
            double ml[2][8], mr[2][8];
            load_msg(p.children[2 * (i - nl)], ml);
            load_msg(p.children[2 * (i - nl) + 1], mr);
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double a = ml[q][j] * mr[q][j];
                    alpha_at(i)[(c0 + q) * 64 + o0 + j] = a;
                    X[(c0 + q) * OUT_XS + o0 + j] = a;
                }
            if (i == n - 1) break;
            stage_L(i, true);
            __syncthreads();
            double y[2][8];
            out_product(X, L, y, tid);
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int j = 0; j < 8; j++) msg_at(i)[(c0 + q) * 64 + o0 + j] = y[q][j];
            __syncthreads();  // L and X free again (msg_i is only ever read back by the thread that wrote it)
        }
        // ---------------- root ----------------
        __syncthreads();
        if (tid < OUT_TC) {
            double z = 0.0;
            for (int a = 0; a < 64; a++) z += X[tid * OUT_XS + a] * p.prior[a];
            zs[tid] = z;
            if (p.z_out && tid < ncols) p.z_out[col0 + tid] = z;
        }
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int j = 0; j < 8; j++) beta_at(n - 1)[(c0 + q) * 64 + o0 + j] = p.prior[o0 + j];
        __syncthreads();
        // ---------------- outside + expected counts, node by node from the top ----------------
        for (int i = n - 2; i >= 0; i--) {
            const int par = p.parent[i], sib = p.sibling[i];
            double ms[2][8], inter[2][8];
            load_msg(sib, ms);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const double* bp = beta_at(par) + (c0 + q) * 64 + o0;
#pragma unroll
                for (int j = 0; j < 8; j++) inter[q][j] = bp[j] * ms[q][j];
            }
            if (i >= nl) {  // beta_i[b] = sum_a inter[a] P_i[a][b]
#pragma unroll
                for (int q = 0; q < 2; q++)
#pragma unroll
                    for (int j = 0; j < 8; j++) X[(c0 + q) * OUT_XS + o0 + j] = inter[q][j];
                stage_L(i, false);
                __syncthreads();
                double y[2][8];
                out_product(X, L, y, tid);
#pragma unroll
                for (int q = 0; q < 2; q++)
#pragma unroll
                    for (int j = 0; j < 8; j++) beta_at(i)[(c0 + q) * 64 + o0 + j] = y[q][j];
                __syncthreads();
            }
            if (gacc) {  // G_i[a][b] += sum_c u[c][a] v[c][b], u = inter / z (columns with z > 0), v = alpha_i
                const int i0 = i;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int c = c0 + q;
                    const double z = zs[c];
                    const bool live = c < ncols && z > 0.0;
#pragma unroll
                    for (int j = 0; j < 8; j++) U[c * OUT_XS + o0 + j] = live ? inter[q][j] / z : 0.0;
                    if (i < nl) {
                        const int code = codes_s[c * nl + i];
#pragma unroll
                        for (int j = 0; j < 8; j++) X[c * OUT_XS + o0 + j] = (code >= 64 || code == o0 + j) ? 1.0 : 0.0;
                    } else {
                        const double* ai = alpha_at(i) + c * 64 + o0;
#pragma unroll
                        for (int j = 0; j < 8; j++) X[c * OUT_XS + o0 + j] = ai[j];
                    }
                }
                __syncthreads();
                // 4 x 4 register tile per thread: four u and four v loads feed sixteen FMAs (a 1 x 16 tile needed
                // seventeen loads for the same work and sat on the shared-memory pipe: 67 % short-scoreboard stalls)
                const int a0 = (tid >> 4) * 4, b0 = (tid & 15) * 4;
                double g[4][4];
#pragma unroll
                for (int ii = 0; ii < 4; ii++)
#pragma unroll
                    for (int j = 0; j < 4; j++) g[ii][j] = 0.0;
#pragma unroll 4
                for (int cc = 0; cc < OUT_TC; cc++) {
                    double u[4], v[4];
#pragma unroll
                    for (int ii = 0; ii < 4; ii++) u[ii] = U[cc * OUT_XS + a0 + ii];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = X[cc * OUT_XS + b0 + j];
#pragma unroll
                    for (int ii = 0; ii < 4; ii++)
#pragma unroll
                        for (int j = 0; j < 4; j++) g[ii][j] = fma(u[ii], v[j], g[ii][j]);
                }
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    double* G = gacc + (size_t)i0 * 4096 + (a0 + ii) * 64 + b0;
#pragma unroll
                    for (int j = 0; j < 4; j++) G[j] += g[ii][j];
                }
                __syncthreads();
            }
        }
        // ---------------- node posteriors (PhyloLik.ml:127-138) ----------------
        for (int q = 0; q < p.n_post; q++) {
            const int node = p.post_nodes[q];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int c = c0 + r;
                if (c >= ncols) continue;
                double* out = p.post_out + ((size_t)q * p.total_cols + col0 + c) * 64 + o0;
                const double z = zs[c];
                if (node < nl) {
                    const int code = codes_s[c * nl + node];
#pragma unroll
                    for (int j = 0; j < 8; j++) out[j] = z == 0.0 ? 0.0 : ((code >= 64 || code == o0 + j) ? 1.0 : 0.0);
                } else {
                    const double* ai = alpha_at(node) + c * 64 + o0;
                    const double* bi = beta_at(node) + c * 64 + o0;
#pragma unroll
                    for (int j = 0; j < 8; j++) out[j] = z == 0.0 ? 0.0 : ai[j] * bi[j] / z;
                }
            }
        }
    }
}

// ---- K6 on the DMMA pipe ---------------------------------------------------------------------------
// Same walk, same scratch blocks, four warps per CTA with 16 columns each held in the accumulator layout of the pruning
// kernels (column 8T + g, states 8j + 2t, 8j + 2t + 1), two CTAs per SM:
//   inside  msg_i = alpha_i x P_i^T   the pruning contraction: A fragments are the alpha registers, B the fragment-ordered
//                                     image of the slot, brought into shared memory by one bulk copy (TMA)
//   outside beta_i = inter x P_i      the same loop over the transposed image (outside_transpose_kernel, once per call)
//   G_i    += U^T V                   warp w owns rows 16w .. 16w+15 of G: U = inter / z and V = alpha_i (or the leaf's
//                                     indicator rows) go through shared memory (row stride 68: fragment loads are
//                                     conflict-free), 256 DMMAs per warp like a contraction
// alpha, msg, beta and the G blocks are stored in the owning thread's register order (32 consecutive double2 per warp
// instruction); outside_reduce_kernel undoes the order of G. Two barriers per node: one before the image / U / V are
// overwritten, one before G's fragments are read.
#ifndef PCSF_OD_T
#define PCSF_OD_T 2
#endif
constexpr int OD_T = PCSF_OD_T;          // 8-column m-tiles per warp: 2 = four warps per CTA, 1 = eight
constexpr int OD_WC = 8 * OD_T;          // columns per warp (= rows of G per warp)
constexpr int OD_THREADS = 32 * 8 / OD_T;
constexpr int OD_TC = 64;
constexpr int OD_TS = 68;
constexpr int OD_SMEM_FIXED = FRAG_BYTES + 2 * OD_TC * OD_TS * 8 + 16;  // image, U, V, one mbarrier
__host__ __device__ constexpr int od_tree_ints(int nl) {  // children | parent | sibling | steps as int16, rounded to 16 bytes
    return (2 * (nl - 1) + 2 * (2 * nl - 1) + 3 * (2 * nl - 2) + 7) & ~7;
}

// transposed fragment images of the internal edges: out[i][frag_index(x, y)] = P_i[y][x]
__global__ void outside_transpose_kernel(const double* __restrict__ tables, int n_leaves, int n_images, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_images * 4096) return;
    const int i = (int)(idx >> 12), e = (int)(idx & 4095);
    const int j = e >> 9, s = (e >> 5) & 15, ln = e & 31;  // inverse of frag_index: e = frag_index(x, y)
    const int x = 8 * j + (ln >> 2), y = 8 * (s >> 1) + 2 * (ln & 3) + (s & 1);
    out[idx] = tables[(size_t)(n_leaves + i) * PT_SLOT + frag_index(y, x)];
}

__global__ void __launch_bounds__(OD_THREADS, 2) outside_dmma_kernel(const OutsideParams p) {
    extern __shared__ __align__(128) unsigned char odsm[];
    double* Pimg = reinterpret_cast<double*>(odsm);
    double* U = Pimg + 4096;
    double* V = U + OD_TC * OD_TS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(V + OD_TC * OD_TS);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int nl = p.n_leaves, n = 2 * nl - 1, ni = nl - 1;
    // the tree and the outside program next to the operands: a step's loads hang off a chain of these lookups
    int16_t* children_s = reinterpret_cast<int16_t*>(bar + 2);
    int16_t* parent_s = children_s + 2 * ni;
    int16_t* sibling_s = parent_s + n;
    int16_t* steps_s = sibling_s + n;
    uint8_t* codes_s = reinterpret_cast<uint8_t*>(children_s + od_tree_ints(nl));
    for (int i = tid; i < 2 * ni + 2 * n + 3 * (n - 1); i += OD_THREADS) children_s[i] = (int16_t)p.children[i];
    // scratch blocks of this CTA in register order: kind 0 alpha, 1 msg, 2 beta; element (T, j) at + (T * 8 + j) * 32
    double2* sc = reinterpret_cast<double2*>(p.scratch + (size_t)blockIdx.x * 3 * ni * OD_TC * 64) + w * (OD_T * 256) + lane;
    auto blk = [&](int kind, int node) { return sc + (size_t)(kind * ni + node - nl) * 2048; };
    double2* gacc = p.gacc ? reinterpret_cast<double2*>(p.gacc + (size_t)blockIdx.x * (n - 1) * 4096) + w * (OD_T * 256) + lane : nullptr;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (p.gacc) {
        double2* z2 = reinterpret_cast<double2*>(p.gacc + (size_t)blockIdx.x * (n - 1) * 4096);
        for (size_t i = tid; i < (size_t)(n - 1) * 2048; i += OD_THREADS) z2[i] = make_double2(0.0, 0.0);
    }
    uint32_t phase = 0;
    const double2* pri = reinterpret_cast<const double2*>(p.prior + 2 * t);  // prior at this lane's states: pri[4 j]

    // every thread is past its reads of the image, U and V; thread 0 starts the next image's copy
    auto turn = [&](const double* image) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (image && tid == 0) {
            mbar_expect_tx(bar, FRAG_BYTES);
            tma_bulk_g2s(Pimg, image, FRAG_BYTES, bar);
        }
    };
    auto load_msg_half = [&](int node, int T, double2 (&m)[8]) {  // the message of `node` for columns 8T + g of this warp
        if (node < nl) {
            int code = codes_s[(OD_WC * w + 8 * T + g) * nl + node];
            code = code > 64 ? 64 : code;
            const double2* row = reinterpret_cast<const double2*>(p.tables + (size_t)node * PT_SLOT + code * 64 + 2 * t);
#pragma unroll
            for (int j = 0; j < 8; j++) m[j] = __ldg(row + 4 * j);
        } else {
            const double2* b = blk(1, node);
#pragma unroll
            for (int j = 0; j < 8; j++) m[j] = b[(T * 8 + j) * 32];
        }
    };
    auto load_msg = [&](int node, double2 (&m)[OD_T][8]) {
#pragma unroll
        for (int T = 0; T < OD_T; T++) load_msg_half(node, T, m[T]);
    };
    auto store_blk = [&](double2* b, const double2 (&v)[OD_T][8]) {
#pragma unroll
        for (int T = 0; T < OD_T; T++)
#pragma unroll
            for (int j = 0; j < 8; j++) b[(T * 8 + j) * 32] = v[T][j];
    };
    // ask L2 for a block this thread will read a node later (blocks written during the inside pass have long left L2 when
    // the outside pass comes back for them: 296 CTAs x 9 MB of scratch against 126 MB)
    auto prefetch_blk = [&](const double2* b) {
        if ((lane & 7) == 0) {
#pragma unroll
            for (int k = 0; k < 16; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + k * 32));
        }
    };
    // y[c][o] = sum_k x[c][k] * M[o][k], M's fragment-ordered image in shared memory
    auto contract = [&](const double2 (&x)[OD_T][8], double2 (&y)[OD_T][8]) {
        const double* Pb = Pimg + lane;
#pragma unroll
        for (int T = 0; T < OD_T; T++)
#pragma unroll
            for (int j = 0; j < 8; j++) y[T][j] = make_double2(0.0, 0.0);
#pragma unroll
        for (int s = 0; s < 16; s++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double bf = Pb[(j * 16 + s) * 32];
#pragma unroll
                for (int T = 0; T < OD_T; T++) dmma(y[T][j].x, y[T][j].y, (s & 1) ? x[T][s >> 1].y : x[T][s >> 1].x, bf);
            }
        }
    };

    const int64_t n_tiles = (p.total_cols + OD_TC - 1) / OD_TC;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t col0 = tile * OD_TC;
        const int ncols = (int)min((int64_t)OD_TC, p.total_cols - col0);
        __syncthreads();
        for (int i = tid; i < OD_TC * nl; i += OD_THREADS) codes_s[i] = i < ncols * nl ? p.codes[(size_t)col0 * nl + i] : (uint8_t)64;
        // ---------------- inside ----------------
        double2 cur[OD_T][8];
        for (int i = nl; i < n; i++) {
            turn(i < n - 1 ? p.tables + (size_t)i * PT_SLOT : nullptr);  // (the first turn also publishes the tile's codes)
            double2 ml[OD_T][8], mr[OD_T][8];
            load_msg(children_s[2 * (i - nl)], ml);
            load_msg(children_s[2 * (i - nl) + 1], mr);
#pragma unroll
            for (int T = 0; T < OD_T; T++)
#pragma unroll
                for (int j = 0; j < 8; j++) cur[T][j] = make_double2(ml[T][j].x * mr[T][j].x, ml[T][j].y * mr[T][j].y);
            if (p.n_post) store_blk(blk(0, i), cur);  // only the node posteriors read alpha back
            if (i == n - 1) break;
            mbar_wait(bar, phase);
            phase ^= 1;
            double2 y[OD_T][8];
            contract(cur, y);
            store_blk(blk(1, i), y);
        }
        // ---------------- root: z = alpha_root . prior ----------------
        double z[OD_T], zinv[OD_T];
        bool live[OD_T];
#pragma unroll
        for (int T = 0; T < OD_T; T++) {
            double zp = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double2 pj = __ldg(pri + 4 * j);
                zp += cur[T][j].x * pj.x;
                zp += cur[T][j].y * pj.y;
            }
            zp += __shfl_xor_sync(0xffffffffu, zp, 1);
            zp += __shfl_xor_sync(0xffffffffu, zp, 2);
            z[T] = zp;
            const int c = OD_WC * w + 8 * T + g;
            live[T] = c < ncols && zp > 0.0;
            zinv[T] = live[T] ? 1.0 / zp : 0.0;  // one division per column: 32 per thread and branch cost a third of the kernel
            if (p.z_out && t == 0 && c < ncols) p.z_out[col0 + c] = zp;
        }
        // ---------------- outside + expected counts: the host's depth-first program (OUT_STEP_*) ----------------
        // beta of the node whose children are being visited stays in registers: a node's leaf children come first, then an
        // internal child whose beta is parked in scratch for later, last the internal child the walk descends into next
        double2 bv[OD_T][8];
        bool pending = false;  // an image is in the buffer (or on its way) and not yet used
        for (int k = 0; k < n - 1; k++) {
            const int i = steps_s[3 * k], fl = steps_s[3 * k + 1];
            const int par = parent_s[i], sib = sibling_s[i];
            // the image buffer is free from the barrier after a contraction: the next internal step's image is asked for at
            // once, steps ahead of its use when leaf steps come in between
            const int ahead = pending ? -1 : steps_s[3 * k + 2];
            turn(ahead >= 0 ? p.pt_images + (size_t)(ahead - nl) * 4096 : nullptr);
            pending |= ahead >= 0;
            if (fl & OUT_STEP_LOADB) {
                if (par == n - 1) {
#pragma unroll
                    for (int T = 0; T < OD_T; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) bv[T][j] = __ldg(pri + 4 * j);
                } else {
                    const double2* bp = blk(2, par);
#pragma unroll
                    for (int T = 0; T < OD_T; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) bv[T][j] = bp[(T * 8 + j) * 32];
                }
            }
            double2 inter[OD_T][8];
            load_msg(sib, inter);
#pragma unroll
            for (int T = 0; T < OD_T; T++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    inter[T][j].x *= bv[T][j].x;
                    inter[T][j].y *= bv[T][j].y;
                }
            if (gacc) {  // U[c][a] = inter / z on live columns (else 0), V[c][b] = alpha_i (a leaf: its indicator row)
#pragma unroll
                for (int T = 0; T < OD_T; T++) {
                    const int c = OD_WC * w + 8 * T + g;
                    double2* ur = reinterpret_cast<double2*>(U + c * OD_TS + 2 * t);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        ur[4 * j] = make_double2(inter[T][j].x * zinv[T], inter[T][j].y * zinv[T]);
                }
                if (i < nl) {
#pragma unroll
                    for (int T = 0; T < OD_T; T++) {
                        const int c = OD_WC * w + 8 * T + g;
                        double2* vr = reinterpret_cast<double2*>(V + c * OD_TS + 2 * t);
                        const int code = codes_s[c * nl + i];
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            vr[4 * j] = make_double2((code >= 64 || code == 8 * j + 2 * t) ? 1.0 : 0.0,
                                                     (code >= 64 || code == 8 * j + 2 * t + 1) ? 1.0 : 0.0);
                    }
                } else {
                    // alpha_i again from its children's messages (the same product, the same bits): the children's blocks
                    // are the ones the next steps read as sibling messages anyway, alpha's own block need not exist
                    const int lc = children_s[2 * (i - nl)], rc = children_s[2 * (i - nl) + 1];
#pragma unroll 1
                    for (int T = 0; T < OD_T; T++) {
                        double2 ml[8], mr[8];
                        load_msg_half(lc, T, ml);
                        load_msg_half(rc, T, mr);
                        double2* vr = reinterpret_cast<double2*>(V + (OD_WC * w + 8 * T + g) * OD_TS + 2 * t);
#pragma unroll
                        for (int j = 0; j < 8; j++) vr[4 * j] = make_double2(ml[j].x * mr[j].x, ml[j].y * mr[j].y);
                    }
                }
            }
            if (k + 1 < n - 1) {  // the next step's sibling message
                const int sn = sibling_s[steps_s[3 * k + 3]];
                if (sn >= nl) prefetch_blk(blk(1, sn));
            }
            if (i >= nl) {  // beta_i[b] = sum_a inter[a] P_i[a][b]
                mbar_wait(bar, phase);
                phase ^= 1;
                double2 y[OD_T][8];
                contract(inter, y);
                if (fl & OUT_STEP_STOREB) store_blk(blk(2, i), y);
                pending = false;
                if (fl & OUT_STEP_CARRY) {
#pragma unroll
                    for (int T = 0; T < OD_T; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) bv[T][j] = y[T][j];
                }
            }
            if (gacc) {
                __syncthreads();
                double2 ga[OD_T][8];
#pragma unroll
                for (int M = 0; M < OD_T; M++)
#pragma unroll
                    for (int j = 0; j < 8; j++) ga[M][j] = make_double2(0.0, 0.0);
                const double* ua = U + t * OD_TS + OD_WC * w + g;
                const double* vb = V + t * OD_TS + g;
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    double af[OD_T];
#pragma unroll
                    for (int M = 0; M < OD_T; M++) af[M] = ua[4 * s * OD_TS + 8 * M];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double bf = vb[4 * s * OD_TS + 8 * j];
#pragma unroll
                        for (int M = 0; M < OD_T; M++) dmma(ga[M][j].x, ga[M][j].y, af[M], bf);
                    }
                }
                // this CTA's block, this thread's elements, tile after tile: reductions without a return value keep the
                // order (and the bits) of a load-add-store and do not wait for the line
                double* G = reinterpret_cast<double*>(gacc + (size_t)i * 2048);
#pragma unroll
                for (int M = 0; M < OD_T; M++)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        atomicAdd(G + (M * 8 + j) * 64, ga[M][j].x);
                        atomicAdd(G + (M * 8 + j) * 64 + 1, ga[M][j].y);
                    }
            }
        }
        // ---------------- node posteriors (PhyloLik.ml:127-138) ----------------
        for (int q = 0; q < p.n_post; q++) {
            const int node = p.post_nodes[q];
#pragma unroll
            for (int T = 0; T < OD_T; T++) {
                const int c = OD_WC * w + 8 * T + g;
                if (c >= ncols) continue;
                double2* out = reinterpret_cast<double2*>(p.post_out + ((size_t)q * p.total_cols + col0 + c) * 64 + 2 * t);
                const double zc = z[T];
                if (node < nl) {
                    const int code = codes_s[c * nl + node];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        out[4 * j] = zc == 0.0 ? make_double2(0.0, 0.0)
                                               : make_double2((code >= 64 || code == 8 * j + 2 * t) ? 1.0 : 0.0,
                                                              (code >= 64 || code == 8 * j + 2 * t + 1) ? 1.0 : 0.0);
                } else {
                    const double2* ai = blk(0, node);
                    const double2* bi = blk(2, node);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double2 a = ai[(T * 8 + j) * 32];
                        const double2 b = node == n - 1 ? __ldg(pri + 4 * j) : bi[(T * 8 + j) * 32];
                        out[4 * j] = zc == 0.0 ? make_double2(0.0, 0.0) : make_double2(a.x * b.x / zc, a.y * b.y / zc);
                    }
                }
            }
        }
    }
}

// ecounts[br][a][b] = P_br[a][b] * sum over CTAs (ascending) of G[cta][br][a][b]; reg_order: the G blocks are in the
// register order of outside_dmma_kernel (m-tile a / 8, n-tile b / 8, lane 4 (a % 8) + (b / 2) % 4, b % 2)
__global__ void outside_reduce_kernel(const double* __restrict__ gacc, int n_cta, int n_branches, int n_leaves,
                                      const double* __restrict__ tables, double* __restrict__ ecounts, int reg_order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_branches * 4096) return;
    const int br = (int)(i >> 12), a = (int)((i >> 6) & 63), b = (int)(i & 63);
    const int e = reg_order ? ((((a >> 3) * 8 + (b >> 3)) * 32 + ((a & 7) * 4 + ((b >> 1) & 3))) * 2 + (b & 1)) : a * 64 + b;
    double s = 0.0;
    for (int k = 0; k < n_cta; k++) s += gacc[((size_t)k * n_branches + br) * 4096 + e];
    ecounts[i] = out_p_entry(tables + (size_t)br * PT_SLOT, br < n_leaves, a, b) * s;
}

}  // namespace pcsf

#include "pcsf_k0.cuh"  // K0: pleaves on the device (frame_codes_kernel)
