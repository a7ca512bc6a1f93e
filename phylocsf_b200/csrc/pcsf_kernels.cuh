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
//
// FP64 tensor path on sm_100a is warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); tcgen05 has no
// f64 kind. Operand staging uses TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarriers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcsf {

constexpr int K = 64;
constexpr int PT_SLOT = 65 * 64;         // doubles per (scale, branch) table slot
constexpr int PT_SLOT_BYTES = PT_SLOT * 8;
constexpr int FRAG_BYTES = 64 * 64 * 8;  // fragment-ordered P image of an internal edge

// ---- tree program (built on the host from T.children) -----------------------------------------
enum OpKind : int32_t {
    OP_CHERRY = 0,     // cur = G(a) * G(b)                       a, b leaves
    OP_GEMM_LEAF = 1,  // cur = (P_a x cur) * G(b)                a internal child, b sibling leaf
    OP_GEMM_PUSH = 2,  // stack[c] = P_a x cur                    sibling subtree still to come
    OP_GEMM_POP = 3,   // cur = (P_a x cur) * stack[c]
    OP_ROOT = 4        // z = cur . prior, log z, root posterior . log prior
};
struct Op {
    int32_t kind, a, b, c;
};

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
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
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
// grid = (n_branches, nscales), block = 256. One CTA builds one 64x64 P(t):
//   warp w computes rows 8w..8w+7 with 128 DMMAs (A = S rows, B = exp(lambda_k t) * Sinv[k][.]),
//   then 64 threads apply the reference's row fix-ups sequentially (bit-faithful order), then the
//   CTA writes the table in the layout the pruning kernel wants: leaves get the transposed gather
//   table PT[code][a] (+ row 64 = row sums, the `Marginalize message); internal edges get the
//   fragment-ordered image.
__global__ void __launch_bounds__(256) pt_build_kernel(const double* __restrict__ S, const double* __restrict__ Sinv,
                                                       const double* __restrict__ lambda,
                                                       const double* __restrict__ branch_len,
                                                       const double* __restrict__ scales, int n_leaves,
                                                       double* __restrict__ tables, int32_t* __restrict__ status,
                                                       double tol) {
    __shared__ double Psm[64][65];
    __shared__ double e_s[64];
    __shared__ double rowsum_s[64];
    const int br = blockIdx.x, sc = blockIdx.y;
    const int n_branches = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const double tt = scales[sc] * branch_len[br];  // Mul (Var 0, Val b), src/PhyloCSFModel.ml:33
    if (tid < 64) e_s[tid] = exp(tt * lambda[tid]);  // Q.ml:216-217
    __syncthreads();
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll 4
    for (int s = 0; s < 16; s++) {
        const int k = 4 * s + t;
        const double a = S[(8 * w + g) * 64 + k];
        const double ek = e_s[k];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double b = Sinv[k * 64 + 8 * j + g] * ek;  // diagm: row k of S' scaled, Q.ml:61-64
            dmma(acc[j][0], acc[j][1], a, b);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        Psm[8 * w + g][8 * j + 2 * t] = acc[j][0];
        Psm[8 * w + g][8 * j + 2 * t + 1] = acc[j][1];
    }
    __syncthreads();
    if (tid < 64) {  // Q.ml:226-247, one row per thread, j ascending
        const int i = tid;
        int st = (tt < 0.0) ? 1 : 0;
        double tot = 0.0, smii = 1.0;
        for (int j = 0; j < 64; j++) {
            double v = Psm[i][j];
            tot += v;
            if (v < 0.0) {
                if (fabs(v) > tol) st |= 2;
                v = 0.0;
                Psm[i][j] = 0.0;
            }
            if (i != j) smii -= v;
        }
        if (fabs(tot - 1.0) > tol) st |= 4;
        if (!(smii <= 1.0 && smii > 0.0)) st |= 8;
        Psm[i][i] = smii;
        double rs = 0.0;  // ddot(row, ones): the `Marginalize leaf message, PhyloLik.ml:90 with raw_marg
        for (int j = 0; j < 64; j++) rs += Psm[i][j];
        rowsum_s[i] = rs;
        if (st) atomicOr(&status[sc], st);
    }
    __syncthreads();
    double* out = tables + ((size_t)sc * n_branches + br) * PT_SLOT;
    if (br < n_leaves) {
        for (int idx = tid; idx < 64 * 64; idx += 256) {
            const int b = idx >> 6, a = idx & 63;
            out[idx] = Psm[a][b];
        }
        if (tid < 64) out[64 * 64 + tid] = rowsum_s[tid];
    } else {
        for (int idx = tid; idx < 64 * 64; idx += 256) {
            // idx = ((j*16+s)*32 + 4g+t)
            const int l = idx & 31, s = (idx >> 5) & 15, j = idx >> 9;
            const int gg = l >> 2, tq = l & 3;
            const int a = 8 * j + gg, b = 8 * (s >> 1) + 2 * tq + (s & 1);
            out[idx] = Psm[a][b];
        }
        if (tid < 64) out[64 * 64 + tid] = rowsum_s[tid];
    }
}

// =================================================================================================
// K2+K3: pruning
// =================================================================================================
constexpr int PRUNE_WARPS = 8;
constexpr int PRUNE_T = 2;                                   // 8-column tiles per warp
constexpr int PRUNE_THREADS = PRUNE_WARPS * 32;
constexpr int WARP_COLS = 8 * PRUNE_T;                       // 16
constexpr int TILE_COLS = PRUNE_WARPS * WARP_COLS;           // 128
constexpr int STACK_ENTRY_BYTES = PRUNE_T * 8 * 32 * 16;     // per warp per level: 8 KB
constexpr int STACK_LEVEL_BYTES = PRUNE_WARPS * STACK_ENTRY_BYTES;  // 64 KB per level per CTA

struct PruneParams {
    const Op* ops;
    int n_ops;
    int n_leaves;
    int n_gemm;  // GEMM ops per tile
    const Span* spans;
    int n_spans;
    int64_t n_tiles;
    const PSet* psets;
    const uint8_t* codes;  // [total_cols][n_leaves]
    double* out_logz;
    double* out_anc;
    int smem_levels;         // stack levels kept in shared memory
    uint8_t* global_stack;   // [gridDim.x][max_levels - smem_levels][STACK_LEVEL_BYTES]
    int global_levels;
    int codes_smem_bytes;    // TILE_COLS * n_leaves rounded up to 16
};

// shared memory map: [P buf 0 | P buf 1 | barriers(64 B) | ops | codes | stack levels]
__device__ __forceinline__ void load_leaf_mul(double (&cur)[PRUNE_T][8][2], const double* __restrict__ tab,
                                              const uint8_t* codes_s, int n_leaves, int leaf, int wcol, int g, int t,
                                              bool init) {
#pragma unroll
    for (int T = 0; T < PRUNE_T; T++) {
        int code = codes_s[(wcol + 8 * T + g) * n_leaves + leaf];
        code = code > 64 ? 64 : code;
        const double2* src = reinterpret_cast<const double2*>(tab + code * 64 + 2 * t);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double2 v = __ldg(src + 4 * j);
            if (init) {
                cur[T][j][0] = v.x;
                cur[T][j][1] = v.y;
            } else {
                cur[T][j][0] *= v.x;
                cur[T][j][1] *= v.y;
            }
        }
    }
}

__global__ void __launch_bounds__(PRUNE_THREADS, 1) prune_kernel(const PruneParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    double* Pbuf0 = reinterpret_cast<double*>(smem);
    double* Pbuf1 = reinterpret_cast<double*>(smem + FRAG_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * FRAG_BYTES);  // [2]
    uint64_t* empty = full + 2;                                           // [2]
    Op* ops_s = reinterpret_cast<Op*>(smem + 2 * FRAG_BYTES + 64);
    const int ops_bytes = ((p.n_ops * (int)sizeof(Op)) + 15) & ~15;
    uint8_t* codes_s = smem + 2 * FRAG_BYTES + 64 + ops_bytes;
    uint8_t* stack_s = codes_s + p.codes_smem_bytes;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wcol = w * WARP_COLS;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], PRUNE_WARPS);
        mbar_init(&empty[1], PRUNE_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = tid; i < p.n_ops; i += PRUNE_THREADS) ops_s[i] = p.ops[i];
    __syncthreads();

    uint32_t gq = 0;  // running count of GEMMs this CTA has gone through (selects buffer and parity)
    uint8_t* gstack = p.global_stack + (size_t)blockIdx.x * p.global_levels * STACK_LEVEL_BYTES;

    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        // ---- locate the span of this tile (binary search over tile0) ----
        int lo = 0, hi = p.n_spans - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.spans[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
        }
        const Span sp = p.spans[lo];
        const PSet ps = p.psets[sp.pset];
        const int64_t tcol0 = (tile - sp.tile0) * TILE_COLS;  // first column of the tile within the span
        const int ncols = (int)min((int64_t)TILE_COLS, (int64_t)sp.ncols - tcol0);
        const bool warp_active = wcol < ncols;

        __syncthreads();  // previous tile fully retired: codes_s and both P buffers are free

        // ---- producer: first P image of the tile ----
        int gemm_seen = 0;  // GEMM ops met so far in this tile (uniform across the CTA)
        // branch ids of GEMM ops are found by scanning ops_s; next_gemm_op = index of the next GEMM op to prefetch
        int next_gemm_op = 0;
        while (next_gemm_op < p.n_ops && (ops_s[next_gemm_op].kind == OP_CHERRY || ops_s[next_gemm_op].kind == OP_ROOT))
            next_gemm_op++;
        if (tid == 0 && next_gemm_op < p.n_ops) {
            const uint32_t q = gq, b = q & 1, u = q >> 1;
            if (u >= 1) mbar_wait(&empty[b], (u - 1) & 1);
            mbar_expect_tx(&full[b], FRAG_BYTES);
            tma_bulk_g2s(b ? Pbuf1 : Pbuf0, ps.tables + (size_t)ops_s[next_gemm_op].a * PT_SLOT, FRAG_BYTES, &full[b]);
        }
        // ---- codes of the tile -> shared ----
        {
            const uint8_t* src = p.codes + (size_t)(sp.col0 + tcol0) * p.n_leaves;
            const int nbytes = ncols * p.n_leaves;
            const int total = TILE_COLS * p.n_leaves;
            for (int i = tid; i < total; i += PRUNE_THREADS) codes_s[i] = (i < nbytes) ? src[i] : (uint8_t)64;
        }
        __syncthreads();

        double cur[PRUNE_T][8][2];
#pragma unroll
        for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
            for (int j = 0; j < 8; j++) cur[T][j][0] = cur[T][j][1] = 1.0;

        for (int oi = 0; oi < p.n_ops; oi++) {
            const Op op = ops_s[oi];
            if (op.kind == OP_CHERRY) {
                if (warp_active) {
                    load_leaf_mul(cur, ps.tables + (size_t)op.a * PT_SLOT, codes_s, p.n_leaves, op.a, wcol, g, t, true);
                    load_leaf_mul(cur, ps.tables + (size_t)op.b * PT_SLOT, codes_s, p.n_leaves, op.b, wcol, g, t, false);
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
                            p.out_logz[sp.out0 + tcol0 + c] = log(zp);
                            p.out_anc[sp.out0 + tcol0 + c] = ap;
                        }
                    }
                }
                continue;
            }
            // ------------------------------ GEMM ops ------------------------------
            const uint32_t q = gq + gemm_seen, b = q & 1, u = q >> 1;
            // producer: prefetch the P image of the next GEMM op of this tile into the other buffer
            {
                int nxt = oi + 1;
                while (nxt < p.n_ops && (ops_s[nxt].kind == OP_CHERRY || ops_s[nxt].kind == OP_ROOT)) nxt++;
                if (tid == 0 && nxt < p.n_ops) {
                    const uint32_t q1 = q + 1, b1 = q1 & 1, u1 = q1 >> 1;
                    if (u1 >= 1) mbar_wait(&empty[b1], (u1 - 1) & 1);
                    mbar_expect_tx(&full[b1], FRAG_BYTES);
                    tma_bulk_g2s(b1 ? Pbuf1 : Pbuf0, ps.tables + (size_t)ops_s[nxt].a * PT_SLOT, FRAG_BYTES, &full[b1]);
                }
                __syncwarp();
            }
            mbar_wait(&full[b], u & 1);
            double acc[PRUNE_T][8][2];
            if (warp_active) {
                const double* Pb = (b ? Pbuf1 : Pbuf0) + lane;
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b]);
            gemm_seen++;
            if (!warp_active) continue;
            // ------------------------------ epilogues ------------------------------
            if (op.kind == OP_GEMM_LEAF) {
#pragma unroll
                for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        cur[T][j][0] = acc[T][j][0];
                        cur[T][j][1] = acc[T][j][1];
                    }
                load_leaf_mul(cur, ps.tables + (size_t)op.b * PT_SLOT, codes_s, p.n_leaves, op.b, wcol, g, t, false);
            } else {
                uint8_t* base = (op.c < p.smem_levels)
                                    ? stack_s + (size_t)op.c * STACK_LEVEL_BYTES
                                    : gstack + (size_t)(op.c - p.smem_levels) * STACK_LEVEL_BYTES;
                double2* slot = reinterpret_cast<double2*>(base + (size_t)w * STACK_ENTRY_BYTES) + lane;
                if (op.kind == OP_GEMM_PUSH) {
#pragma unroll
                    for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) slot[(T * 8 + j) * 32] = make_double2(acc[T][j][0], acc[T][j][1]);
                } else {  // OP_GEMM_POP
#pragma unroll
                    for (int T = 0; T < PRUNE_T; T++)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const double2 v = slot[(T * 8 + j) * 32];
                            cur[T][j][0] = acc[T][j][0] * v.x;
                            cur[T][j][1] = acc[T][j][1] * v.y;
                        }
                }
            }
        }
        gq += (uint32_t)gemm_seen;
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
// K0: pleaves on the device. One thread per (region column, leaf).
// =================================================================================================
__device__ __forceinline__ int nt_index(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}
__device__ __forceinline__ uint8_t nt_comp(uint8_t c) {  // Code.ml:39-51 (validated on the host)
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return c;
    }
}
// region r = alignment a, frame f: columns at lo = f%3, strand = f/3 (src/PhyloCSF.ml:198-205,219-246)
__global__ void frame_codes_kernel(const uint8_t* __restrict__ nt, const int64_t* __restrict__ aln_off,
                                   const int32_t* __restrict__ aln_len, const int64_t* __restrict__ region_off,
                                   int64_t nregions, int frames, int n_leaves, uint8_t* __restrict__ codes) {
    const int64_t r = blockIdx.x;
    if (r >= nregions) return;
    const int64_t a = r / frames;
    const int f = (int)(r - a * frames);
    const int ofs = f % 3;
    const bool rc = f >= 3;
    const int len = aln_len[a];
    const uint8_t* base = nt + aln_off[a];
    const int64_t c0 = region_off[r];
    const int ncols = (int)(region_off[r + 1] - c0);
    for (int idx = threadIdx.x; idx < ncols * n_leaves; idx += blockDim.x) {
        const int c = idx / n_leaves, l = idx - c * n_leaves;
        const int pos = ofs + 3 * c;
        const uint8_t* row = base + (size_t)l * len;
        uint8_t n1, n2, n3;
        if (!rc) {
            n1 = row[pos]; n2 = row[pos + 1]; n3 = row[pos + 2];
        } else {
            n1 = nt_comp(row[len - 1 - pos]); n2 = nt_comp(row[len - 2 - pos]); n3 = nt_comp(row[len - 3 - pos]);
        }
        const int i1 = nt_index(n1), i2 = nt_index(n2), i3 = nt_index(n3);
        codes[(size_t)(c0 + c) * n_leaves + l] = (i1 < 0 || i2 < 0 || i3 < 0) ? (uint8_t)64 : (uint8_t)(16 * i1 + 4 * i2 + i3);
    }
}

}  // namespace pcsf
