// pcsf_api.cu — the C ABI declared in include/phylocsf_b200.h: context, resident tables, batch
// staging, launch sequencing, and the batched Brent driver. No torch types; plain CUDA runtime.
#include "../../include/phylocsf_b200.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pcsf_brent.hpp"
#include "pcsf_kernels.cuh"

using namespace pcsf;

#define PCSF_VERSION_STRING "phylocsf_b200 0.1 (sm_100a)"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Model {
    bool set = false;
    bool owns_params = true;     // false: d_params points into the context's batch block (pcsf_models_set)
    double* d_params = nullptr;  // S | Sinv | lambda | prior | logprior  (64*64*2 + 3*64 doubles)
    DevBuf tables;               // [nscales][n_branches][PT_SLOT]
    DevBuf d_scales, d_status;
    int nscales = 0;
    std::vector<int32_t> status;
    DevBuf cherry;                  // [nscales][subtree tables of one P set], built on demand (fixed strategy)
    std::vector<char> cherry_built; // per scale: 0, or the level built (2 cherries, 3 cherries + cherry-and-leaf subtrees)
    int cherry_level = 0;           // level the buffer's per-scale stride was sized for
    struct TabBlock* tab = nullptr; // one-scale models (fixed strategy): subtree tables shared by the process's contexts on this GPU
    uint64_t params_hash = 0;       // of S | Sinv | lambda as given to pcsf_model_set (0: not shareable)
    std::vector<double> scales;     // the tree scales of the P tables
    const double* prior() const { return d_params + 8192 + 64; }
    const double* logprior() const { return d_params + 8192 + 128; }
};

// Subtree tables of one P set, shared by every context of the process on the same GPU that has the same tree, the same
// diagonalised model and the same scale (the command line runs two contexts per device; 58mammals level 4 is 47 GB per
// model - two private copies do not fit in 180 GB). A block is immutable once built. A higher level is a NEW block that
// becomes the key's current one; a context keeps the block it holds until its next call and a block is freed when the
// last context lets go of it, so nobody's kernels ever run on freed memory.
struct TabBlock {
    int device = 0;
    uint64_t key = 0;
    int level = 0;
    void* p = nullptr;
    size_t bytes = 0;
    cudaEvent_t ready = nullptr;  // recorded after the building kernels: other contexts' streams wait for it
    int refs = 0;
    bool current = false;
};
struct TabKey {  // per (GPU, P set): how many columns have been scored under it by all contexts (cumulative level rule)
    int device;
    uint64_t key;
    int64_t cols_seen;
};
std::mutex g_tab_mu;
std::vector<TabBlock*> g_tab_blocks;
std::vector<TabKey> g_tab_keys;

uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = (const unsigned char*)data;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

void tab_release_locked(TabBlock* b) {  // g_tab_mu held
    if (!b) return;
    if (--b->refs > 0) return;
    for (size_t i = 0; i < g_tab_blocks.size(); i++)
        if (g_tab_blocks[i] == b) { g_tab_blocks.erase(g_tab_blocks.begin() + i); break; }
    bool other = false;  // the P set's history goes with its last block: a later context starts from zero, like a new process
    for (TabBlock* q : g_tab_blocks) other |= q->device == b->device && q->key == b->key;
    if (!other)
        for (size_t i = 0; i < g_tab_keys.size(); i++)
            if (g_tab_keys[i].device == b->device && g_tab_keys[i].key == b->key) { g_tab_keys.erase(g_tab_keys.begin() + i); break; }
    cudaSetDevice(b->device);
    if (b->p) cudaFree(b->p);  // (synchronises the device: nothing in flight still reads the block)
    if (b->ready) cudaEventDestroy(b->ready);
    delete b;
}

void tab_release(Model& m) {
    if (!m.tab) return;
    std::lock_guard<std::mutex> lk(g_tab_mu);
    tab_release_locked(m.tab);
    m.tab = nullptr;
}

}  // namespace

struct pcsf_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    // tree
    int n_leaves = 0, n_branches = 0;
    std::vector<int32_t> children;
    std::vector<double> branch_len;
    double* d_branch_len = nullptr;
    Op* d_ops = nullptr;
    Item* d_items = nullptr;
    std::vector<Op> ops;
    std::vector<Item> items;          // what the producer warpgroup stages per tile, in program order
    int n_gemm = 0, max_levels = 0;
    // models
    std::vector<Model> models;
    // batch
    int64_t nregions = -1, total_cols = 0;
    std::vector<int64_t> region_off;
    DevBuf d_region_off, d_codes, d_nt, d_aln_off, d_aln_len;
    // work + outputs
    DevBuf d_spans, d_psets, d_out_logz, d_out_anc, d_seg_begin, d_seg_end, d_lpr, d_elpr, d_gstack;
    DevBuf d_jobs, d_batch_params, d_pair_tables, d_pair_status, d_qs, d_gexp, d_pair_cherry;
    DevBuf d_eig_cache, d_eig_valid, d_eig_slots, d_eig_sweeps;  // K5 warm starts (pcsf_omega_models_set_cached)
    int64_t eig_slots = 0;
    int64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // pcsf_counter: 0 K1 slots built, 1 codon-column evaluations pruned, 2 K5 matrices, 3 K5 sweeps, 4 tiles
    DevBuf d_out_tree, d_out_scratch, d_out_gacc, d_out_post, d_out_ecounts, d_out_z, d_out_nodes, d_out_images;  // K6 (pcsf_posteriors)
    // pcsf_score_alignments: double-buffered chunk staging on a second stream
    DevBuf pipe_nt[2], pipe_aln_off[2], pipe_aln_len[2], pipe_codes[2], pipe_roff[2];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    std::vector<int32_t> pair_model, pair_status;  // P sets built by pcsf_pt_build_pairs
    int last_all_models = 0;
    // timing
    cudaEvent_t ev[12];
    double ms[6] = {-1, -1, -1, -1, -1, -1};  // [5]: the most recent build of subtree tables
    double ms_total[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // cumulative device time per kernel class (pcsf_total_ms); [6] = K5
    int64_t launches = 0;
    int prune_smem_optin = 0;
    int skew_ns = 1500;
    // cherry-table program: the tree program with every (cherry, edge above it) pair folded into one lookup
    std::vector<Op> ops_t, ops_t3, ops_t4;  // level 2 (cherries), 3 (+ cherry-and-leaf subtrees), 4 (+ caterpillars of four) programs
    std::vector<Item> items_t, items_t3, items_t4;
    std::vector<SubTab> subtabs;        // cherries first (n_tab2), then the 3-leaf (n_tab3) and the 4-leaf subtrees (n_tab4)
    int n_tab2 = 0, n_tab3 = 0, n_tab4 = 0;
    Op *d_ops_t = nullptr, *d_ops_t3 = nullptr, *d_ops_t4 = nullptr;
    Item *d_items_t = nullptr, *d_items_t3 = nullptr, *d_items_t4 = nullptr;
    SubTab* d_subtabs = nullptr;
    long long* d_tab_off = nullptr;
    int cherry_mode = 0;  // PCSF_OPT_CHERRY_TABLES: 0 by the number of columns a P set scores, 1 never, 2 always levels 2-3, 3 always cherries only, 4 always levels 2-4
    int wide = -1;  // pruning kernel form: 0 narrow (128-column tiles), 1 wide (192), -1 chosen per launch (PCSF_WIDE overrides)
    int rescale = 0;  // PCSF_OPT_RESCALE
    uint64_t tree_hash = 0;
    int share_tables = 1;  // PCSF_SHARE_TABLES=0: every context builds its own subtree tables (as in round 1)
    int k6_plain = 0;      // PCSF_K6_PLAIN=1: pcsf_posteriors runs the plain-FP64 form of K6 instead of the DMMA form
    int last_form = 0, last_level = 0, last_grid = 0;  // what the most recent pruning launch ran (pcsf_last_launch_info)
    int64_t last_tiles = 0;
    void* timeline = nullptr;  // PCSF_TIMELINE debug builds (tools/timeline.py)
    int timeline_cap = 0;
};

namespace {

int fail(pcsf_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, PCSF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

int reserve(pcsf_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return PCSF_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + std::min<size_t>(bytes / 8, (size_t)256 << 20) + 256;  // head room for slowly growing batches, capped
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, PCSF_ERR_NOMEM, "cudaMalloc of " + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
    }
    b.cap = want;
    return PCSF_OK;
}
#define TRY(x)                    \
    do {                          \
        int r_ = (x);             \
        if (r_ != PCSF_OK) return r_; \
    } while (0)

int r16(int x) { return (x + 15) & ~15; }
int prune_smem_for(int n_ops, int n_items, int n_leaves) {
    return P_STAGES * FRAG_BYTES + M_STAGES * STACK_LEVEL_BYTES + PRUNE_BAR_BYTES + r16(n_ops * (int)sizeof(Op)) +
           r16(n_items * (int)sizeof(Item)) + r16(TILE_COLS * n_leaves);
}

// Launch K2+K3 over `spans`, then K4 over the given segments. Outputs land in ctx->d_lpr/d_elpr.
int prune_wide_smem_for(int n_ops, int n_items, int n_leaves) {
    return W_P_STAGES * FRAG_BYTES + W_L_STAGES * PT_SLOT_BYTES + W_BAR_BYTES + r16(n_ops * (int)sizeof(Op)) +
           r16(n_items * (int)sizeof(Item)) + 2 * r16(W_TILE_COLS * n_leaves) + W_WARPS * 32;  // + per-warp scratch lines
}

int run_prune(pcsf_ctx* ctx, const std::vector<Span>& spans_in, const std::vector<PSet>& psets, int64_t out_cols,
              const uint8_t* codes = nullptr) {
    std::vector<Span> spans = spans_in;
    int64_t tiles = 0;
    // Which form: a sub-partition's time for a tile is (its warps with columns) x (one contraction), so in units
    // of 64 columns a narrow tile costs ceil(n/64) <= 2 and a wide tile ceil(n/64) <= 3. Full wide tiles run at
    // 89 % of the DMMA peak against 86.5 % for narrow ones; part-filled wide tiles run ~4 % behind narrow ones
    // (measured with 100-column regions). Long spans (fixed strategy) go wide, one short region per P set
    // (mle, omega) stays narrow.
    bool wide = ctx->wide > 0;
    if (ctx->wide < 0) {
        double cost_n = 0, cost_w = 0;
        for (const Span& s : spans) {
            const int64_t n = s.ncols;
            cost_n += ((n / TILE_COLS) * 2 + ((n % TILE_COLS) + 63) / 64) / 0.865;
            cost_w += (n / W_TILE_COLS) * 3 / 0.89 + (((n % W_TILE_COLS) + 63) / 64) / 0.83;
        }
        wide = cost_w < cost_n && prune_wide_smem_for((int)ctx->ops.size(), (int)ctx->items.size(), ctx->n_leaves) <= ctx->prune_smem_optin;
    }
    const int tile_cols = wide ? W_TILE_COLS : TILE_COLS;
    const size_t level_bytes = wide ? W_LEVEL_BYTES : STACK_LEVEL_BYTES;
    for (auto& s : spans) {
        s.tile0 = tiles;
        tiles += (s.ncols + tile_cols - 1) / tile_cols;
    }
    // spans with zero columns would break the tile0 search (duplicate tile0): drop them
    spans.erase(std::remove_if(spans.begin(), spans.end(), [](const Span& s) { return s.ncols == 0; }), spans.end());
    TRY(reserve(ctx, ctx->d_out_logz, sizeof(double) * std::max<int64_t>(out_cols, 1)));
    TRY(reserve(ctx, ctx->d_out_anc, sizeof(double) * std::max<int64_t>(out_cols, 1)));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (tiles > 0) {
        TRY(reserve(ctx, ctx->d_spans, sizeof(Span) * spans.size()));
        TRY(reserve(ctx, ctx->d_psets, sizeof(PSet) * psets.size()));
        CU(cudaMemcpyAsync(ctx->d_spans.p, spans.data(), sizeof(Span) * spans.size(), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_psets.p, psets.data(), sizeof(PSet) * psets.size(), cudaMemcpyHostToDevice, ctx->stream));
        const int grid = (int)std::min<int64_t>(tiles, ctx->num_sms);
        if (ctx->max_levels > 0) TRY(reserve(ctx, ctx->d_gstack, (size_t)grid * ctx->max_levels * level_bytes));
        // the cherry-table program runs when the wide form does and every P set of the launch carries its tables
        long long level = (wide && ctx->n_tab2 > 0) ? 4 : 0;
        for (const PSet& q : psets) level = std::min(level, q.cherry ? q.tab_level : 0LL);
        if (level == 4 && ctx->n_tab4 == 0) level = 3;
        if (level == 3 && ctx->n_tab3 == 0) level = 2;
        const std::vector<Op>& prog_ops = level == 4 ? ctx->ops_t4 : level == 3 ? ctx->ops_t3 : level == 2 ? ctx->ops_t : ctx->ops;
        const std::vector<Item>& prog_items = level == 4 ? ctx->items_t4 : level == 3 ? ctx->items_t3 : level == 2 ? ctx->items_t : ctx->items;
        PruneParams p;
        memset(&p, 0, sizeof(p));
        p.ops = level == 4 ? ctx->d_ops_t4 : level == 3 ? ctx->d_ops_t3 : level == 2 ? ctx->d_ops_t : ctx->d_ops;
        p.n_ops = (int)prog_ops.size();
        p.items = level == 4 ? ctx->d_items_t4 : level == 3 ? ctx->d_items_t3 : level == 2 ? ctx->d_items_t : ctx->d_items;
        p.n_items = (int)prog_items.size();
        p.tab_off = ctx->d_tab_off;
        p.n_leaves = ctx->n_leaves;
        p.spans = (const Span*)ctx->d_spans.p;
        p.n_spans = (int)spans.size();
        p.n_tiles = tiles;
        p.psets = (const PSet*)ctx->d_psets.p;
        p.codes = codes ? codes : (const uint8_t*)ctx->d_codes.p;
        p.out_logz = (double*)ctx->d_out_logz.p;
        p.out_anc = (double*)ctx->d_out_anc.p;
        p.global_stack = (uint8_t*)ctx->d_gstack.p;
        p.n_levels = ctx->max_levels;
        p.ops_bytes = r16((int)(prog_ops.size() * sizeof(Op)));
        p.items_bytes = r16((int)(prog_items.size() * sizeof(Item)));
        p.timeline = (long long*)ctx->timeline;
        p.timeline_cap = ctx->timeline_cap;
        p.skew_ns = ctx->skew_ns;
        TRY(reserve(ctx, ctx->d_gexp, sizeof(int32_t) * (size_t)grid * std::max(1, ctx->max_levels) * tile_cols));
        p.global_exp = (int32_t*)ctx->d_gexp.p;
        const int smem = wide ? prune_wide_smem_for(p.n_ops, p.n_items, ctx->n_leaves) : prune_smem_for(p.n_ops, p.n_items, ctx->n_leaves);
        if (smem > ctx->prune_smem_optin)
            return fail(ctx, PCSF_ERR_INVALID_ARG, "tree too large for the pruning kernel's shared memory (" + std::to_string(smem) + " bytes)");
        if (wide) {
            if (ctx->rescale) {
                CU(cudaFuncSetAttribute(prune_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                prune_wide_kernel<true><<<grid, W_THREADS, smem, ctx->stream>>>(p);
            } else {
                CU(cudaFuncSetAttribute(prune_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                prune_wide_kernel<false><<<grid, W_THREADS, smem, ctx->stream>>>(p);
            }
        } else if (ctx->rescale) {  // PCSF_OPT_RESCALE: separate instantiation, the default path carries no extra code
            CU(cudaFuncSetAttribute(prune_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            prune_kernel<true><<<grid, PRUNE_THREADS, smem, ctx->stream>>>(p);
        } else {
            CU(cudaFuncSetAttribute(prune_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            prune_kernel<false><<<grid, PRUNE_THREADS, smem, ctx->stream>>>(p);
        }
        CU(cudaGetLastError());
        ctx->launches++;
        ctx->last_form = wide ? 2 : 1;
        ctx->last_level = (int)level;
        ctx->last_grid = grid;
        ctx->last_tiles = tiles;
        ctx->counters[4] += tiles;
        for (const Span& sp : spans) ctx->counters[1] += sp.ncols;
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    return PCSF_OK;
}

int run_reduce(pcsf_ctx* ctx, int64_t n_segs) {
    TRY(reserve(ctx, ctx->d_lpr, sizeof(double) * std::max<int64_t>(n_segs, 1)));
    TRY(reserve(ctx, ctx->d_elpr, sizeof(double) * std::max<int64_t>(n_segs, 1)));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (n_segs > 0) {
        const int threads = 256;
        const int64_t blocks = (n_segs * 32 + threads - 1) / threads;
        region_reduce_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
            (const double*)ctx->d_out_logz.p, (const double*)ctx->d_out_anc.p, (const int64_t*)ctx->d_seg_begin.p,
            (const int64_t*)ctx->d_seg_end.p, n_segs, (double*)ctx->d_lpr.p, (double*)ctx->d_elpr.p);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    return PCSF_OK;
}

int finish_timing(pcsf_ctx* ctx) {
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
    ctx->ms[0] = t;
    ctx->ms_total[0] += t;
    CU(cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3]));
    ctx->ms[1] = t;
    ctx->ms_total[1] += t;
    return PCSF_OK;
}

int check_ready(pcsf_ctx* ctx, bool need_batch) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (ctx->n_leaves == 0) return fail(ctx, PCSF_ERR_STATE, "pcsf_tree_set has not been called");
    if (need_batch && ctx->nregions < 0) return fail(ctx, PCSF_ERR_STATE, "no batch staged (pcsf_batch_upload)");
    return PCSF_OK;
}

int check_model(pcsf_ctx* ctx, int model_id, int scale) {
    if (model_id < 0 || model_id >= (int)ctx->models.size() || !ctx->models[model_id].set)
        return fail(ctx, PCSF_ERR_STATE, "model " + std::to_string(model_id) + " has not been set (pcsf_model_set)");
    const Model& m = ctx->models[model_id];
    if (scale < 0 || scale >= m.nscales)
        return fail(ctx, PCSF_ERR_STATE, "model " + std::to_string(model_id) + " has no P tables for scale index " + std::to_string(scale) + " (pcsf_pt_build)");
    return PCSF_OK;
}

size_t table_block_doubles(const pcsf_ctx* ctx, int level) {  // subtree tables of one P set
    return (size_t)ctx->n_tab2 * CHERRY_TABLE + (level >= 3 ? (size_t)ctx->n_tab3 * (size_t)TRIPLE_TABLE : 0) +
           (level >= 4 ? (size_t)ctx->n_tab4 * (size_t)QUAD_TABLE : 0);
}

PSet make_pset(const pcsf_ctx* ctx, int model_id, int scale) {
    const Model& m = ctx->models[model_id];
    PSet ps;
    ps.tables = (const double*)m.tables.p + (size_t)scale * ctx->n_branches * PT_SLOT;
    ps.prior = m.prior();
    ps.logprior = m.logprior();
    ps.cherry = nullptr;
    ps.tab_level = 0;
    if (m.tab && ctx->cherry_mode != 1) {  // shared block (one-scale models); a block of level L holds the tables of all levels <= L
        const int cap = ctx->cherry_mode == 2 ? 3 : ctx->cherry_mode == 3 ? 2 : 4;  // PCSF_OPT_CHERRY_TABLES: "always, up to ..."
        ps.cherry = (const double*)m.tab->p;
        ps.tab_level = std::min(m.tab->level, cap);
    } else if (!m.tab && ctx->cherry_mode != 1 && scale < (int)m.cherry_built.size() && m.cherry_built[scale]) {
        ps.cherry = (const double*)m.cherry.p + (size_t)scale * table_block_doubles(ctx, m.cherry_level);
        ps.tab_level = m.cherry_built[scale];
    }
    return ps;
}

// Which subtree tables a P set that scores `cols_per_pset` columns should carry: cherries (2.16 MB each, built in
// well under a millisecond) pay from a few ten thousand columns; the 140.6 MB tables of cherry-and-leaf subtrees
// from about a million, and only while they fit comfortably in device memory.
// `seen`: columns already scored under this P set (by all contexts that share its tables). A level pays for itself over
// a run, not only over one call: level 4 costs 45 ms to build for both models and saves 2.2 ns per column against
// level 3, i.e. breaks even at 20 M columns - a P set that has seen twice that gets it even if no single batch is big.
int want_table_level(pcsf_ctx* ctx, int64_t cols_per_pset, int64_t seen = 0) {
    if (ctx->cherry_mode == 1 || ctx->wide == 0 || ctx->n_tab2 == 0) return 0;
    if (ctx->cherry_mode >= 2) return ctx->cherry_mode == 2 ? 3 : ctx->cherry_mode == 3 ? 2 : 4;
    const int64_t cum = cols_per_pset + seen;
    return (cols_per_pset >= 5000000 || cum >= 40000000) ? 4 : (cols_per_pset >= 1000000 || cum >= 8000000) ? 3
           : (cols_per_pset >= 50000 || cum >= 400000) ? 2 : 0;
}

// Build (once) the subtree tables of model `m` at scale index `scale` up to `level`. Only for models with a handful
// of scales (the fixed strategy's): per-candidate P sets of mle / omega score too few columns to pay for them.
int build_table_levels(pcsf_ctx* ctx, const double* tables, double* base, int from_level, int level) {
    if (from_level < 2) {
        const long long warps = (long long)ctx->n_tab2 * ((CHERRY_ROWS + 15) / 16);
        subtree_table_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(tables, ctx->d_subtabs, 0, ctx->n_tab2, CHERRY_ROWS, base);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    if (level >= 3 && from_level < 3 && ctx->n_tab3 > 0) {
        const long long warps = (long long)ctx->n_tab3 * ((TRIPLE_ROWS + 15) / 16);
        subtree_table_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(tables, ctx->d_subtabs, ctx->n_tab2, ctx->n_tab3, TRIPLE_ROWS, base);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    if (level >= 4 && ctx->n_tab4 > 0) {
        const long long warps = (long long)ctx->n_tab4 * ((QUAD_ROWS + 15) / 16);
        subtree_table_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(tables, ctx->d_subtabs, ctx->n_tab2 + ctx->n_tab3, ctx->n_tab4, QUAD_ROWS, base);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    return PCSF_OK;
}

// The shared path: one-scale models whose inputs are known by hash (pcsf_model_set). `est_cols` columns are about to be
// scored under the P set.
int ensure_tables_shared(pcsf_ctx* ctx, Model& m, int64_t est_cols) {
    const double scale_v = m.scales[0];
    uint64_t key = fnv1a(&ctx->tree_hash, sizeof(uint64_t), m.params_hash);
    key = fnv1a(&scale_v, sizeof(double), key);
    std::lock_guard<std::mutex> lk(g_tab_mu);
    TabKey* ks = nullptr;
    for (TabKey& k : g_tab_keys)
        if (k.device == ctx->device && k.key == key) ks = &k;
    if (!ks) {
        g_tab_keys.push_back(TabKey{ctx->device, key, 0});
        ks = &g_tab_keys.back();
    }
    int level = want_table_level(ctx, est_cols, ks->cols_seen);
    ks->cols_seen += est_cols;
    if (level == 4 && ctx->n_tab4 == 0) level = 3;
    if (level == 3 && ctx->n_tab3 == 0) level = 2;
    TabBlock* cur = nullptr;
    for (TabBlock* b : g_tab_blocks)
        if (b->device == ctx->device && b->key == key && b->current) cur = b;
    auto attach = [&](TabBlock* b) {
        if (m.tab != b) {
            b->refs++;
            tab_release_locked(m.tab);
            m.tab = b;
            if (b->ready) cudaStreamWaitEvent(ctx->stream, b->ready, 0);
        }
        if ((int)m.cherry_built.size() != m.nscales) m.cherry_built.assign(m.nscales, 0);
        m.cherry_built[0] = (char)b->level;
    };
    const int have = cur ? cur->level : 0;
    if (level <= have) {
        if (cur) attach(cur);
        return PCSF_OK;
    }
    // a new block at `level`: large levels only while they fit comfortably next to everything else on the device
    void* p = nullptr;
    size_t bytes = 0;
    for (; level > have; level--) {
        if (level == 3 && ctx->n_tab3 == 0) continue;
        bytes = sizeof(double) * table_block_doubles(ctx, level);
        if (level >= 3) {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
            // level 4 (tens of GB): room for the other model's block too and 16 GB to spare; level 3: a few blocks' worth
            if ((double)free_b < (level == 3 ? 6.0 : 2.0) * (double)bytes + (level == 4 ? 16e9 : 0.0)) continue;
        }
        CU(cudaStreamSynchronize(ctx->stream));
        if (cudaMalloc(&p, bytes) == cudaSuccess) break;
        cudaGetLastError();
        p = nullptr;
    }
    if (!p) {
        if (cur) attach(cur);
        return PCSF_OK;  // no (better) tables is fine too
    }
    TabBlock* nb = new TabBlock();
    nb->device = ctx->device;
    nb->key = key;
    nb->level = level;
    nb->p = p;
    nb->bytes = bytes;
    CU(cudaEventCreateWithFlags(&nb->ready, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev[10], ctx->stream));
    TRY(build_table_levels(ctx, (const double*)m.tables.p, (double*)p, 0, level));
    CU(cudaEventRecord(nb->ready, ctx->stream));
    CU(cudaEventRecord(ctx->ev[11], ctx->stream));
    CU(cudaEventSynchronize(ctx->ev[11]));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[10], ctx->ev[11]));
    ctx->ms[5] = t;
    ctx->ms_total[5] += t;
    if (cur) cur->current = false;
    nb->current = true;
    g_tab_blocks.push_back(nb);
    attach(nb);
    return PCSF_OK;
}

int ensure_tables(pcsf_ctx* ctx, Model& m, int scale, int level) {
    if (level == 0 || ctx->n_tab2 == 0 || m.nscales > 8) return PCSF_OK;
    if (level == 4 && ctx->n_tab4 == 0) level = 3;
    if (level == 3 && ctx->n_tab3 == 0) level = 2;
    if ((int)m.cherry_built.size() != m.nscales) m.cherry_built.assign(m.nscales, 0);
    if (m.cherry_built[scale] >= level) return PCSF_OK;
    while (level >= 3 && m.cherry_level < level) {  // about to allocate large tables: only while they fit comfortably
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
        const double need = sizeof(double) * (double)table_block_doubles(ctx, level) * m.nscales;
        // level 3 (a GB or so): room for a few such blocks and the batch; level 4 (tens of GB): room for the other
        // model's block too and 16 GB to spare
        if ((double)free_b + (double)m.cherry.cap >= (level == 3 ? 6.0 * need : 2.0 * need + 16e9)) break;
        level--;
        if (level == 3 && ctx->n_tab3 == 0) level = 2;
        if (m.cherry_built[scale] >= level) return PCSF_OK;
    }
    if (m.cherry_level < level || m.cherry.cap < sizeof(double) * table_block_doubles(ctx, level) * m.nscales) {
        CU(cudaStreamSynchronize(ctx->stream));
        std::fill(m.cherry_built.begin(), m.cherry_built.end(), 0);
        // the memory check above is only a snapshot (another scoring context of the process may allocate at the same
        // moment): when the allocation fails after all, settle for the level below instead of failing the call
        for (;;) {
            m.cherry_level = level;
            const int rc = reserve(ctx, m.cherry, sizeof(double) * table_block_doubles(ctx, level) * m.nscales);
            if (rc == PCSF_OK) break;
            m.cherry_level = 0;
            if (rc != PCSF_ERR_NOMEM || level <= 2) return level <= 2 && rc == PCSF_ERR_NOMEM ? PCSF_OK : rc;  // no tables at all is fine too
            level--;
            if (level == 3 && ctx->n_tab3 == 0) level = 2;
        }
    }
    const double* tables = (const double*)m.tables.p + (size_t)scale * ctx->n_branches * PT_SLOT;
    double* base = (double*)m.cherry.p + (size_t)scale * table_block_doubles(ctx, m.cherry_level);
    CU(cudaEventRecord(ctx->ev[10], ctx->stream));
    if (m.cherry_built[scale] < 2) {
        const long long warps = (long long)ctx->n_tab2 * ((CHERRY_ROWS + 15) / 16);
        subtree_table_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(tables, ctx->d_subtabs, 0, ctx->n_tab2, CHERRY_ROWS, base);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    if (level >= 3 && m.cherry_built[scale] < 3 && ctx->n_tab3 > 0) {
        const long long warps = (long long)ctx->n_tab3 * ((TRIPLE_ROWS + 15) / 16);
        subtree_table_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(tables, ctx->d_subtabs, ctx->n_tab2, ctx->n_tab3, TRIPLE_ROWS, base);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    if (level >= 4 && ctx->n_tab4 > 0) {
        const long long warps = (long long)ctx->n_tab4 * ((QUAD_ROWS + 15) / 16);
        subtree_table_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(tables, ctx->d_subtabs, ctx->n_tab2 + ctx->n_tab3, ctx->n_tab4, QUAD_ROWS, base);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    CU(cudaEventRecord(ctx->ev[11], ctx->stream));
    CU(cudaEventSynchronize(ctx->ev[11]));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[10], ctx->ev[11]));
    ctx->ms[5] = t;
    ctx->ms_total[5] += t;
    m.cherry_built[scale] = (char)level;
    return PCSF_OK;
}

// K1 over a list of jobs: tables[job][branch][PT_SLOT], status[job]
int pt_build_jobs(pcsf_ctx* ctx, const std::vector<PtJob>& jobs, DevBuf& tables, DevBuf& d_status, std::vector<int32_t>& status) {
    const int64_t n = (int64_t)jobs.size();
    TRY(reserve(ctx, tables, sizeof(double) * (size_t)n * ctx->n_branches * PT_SLOT));
    TRY(reserve(ctx, ctx->d_jobs, sizeof(PtJob) * n));
    TRY(reserve(ctx, d_status, sizeof(int32_t) * n));
    CU(cudaMemcpyAsync(ctx->d_jobs.p, jobs.data(), sizeof(PtJob) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_status.p, 0, sizeof(int32_t) * n, ctx->stream));
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    {
        const long long n_items = (long long)n * ctx->n_branches;
        const int grid = (int)std::min<long long>((n_items + 1) / 2, 2LL * ctx->num_sms);
        CU(cudaFuncSetAttribute(pt_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM));
        pt_build_kernel<<<grid, K1_THREADS, K1_SMEM, ctx->stream>>>((const PtJob*)ctx->d_jobs.p, n_items, ctx->d_branch_len, ctx->n_branches,
                                                                   ctx->n_leaves, (double*)tables.p, (int32_t*)d_status.p, 1e-6);
        CU(cudaGetLastError());
        ctx->launches++;
        ctx->counters[0] += n_items;
    }
    CU(cudaEventRecord(ctx->ev[5], ctx->stream));
    status.assign(n, 0);
    CU(cudaMemcpyAsync(status.data(), d_status.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]));
    ctx->ms[2] = t;
    ctx->ms_total[2] += t;
    return PCSF_OK;
}

int pt_build_device(pcsf_ctx* ctx, Model& m, int nscales, const double* scales) {
    std::vector<PtJob> jobs(nscales);
    for (int i = 0; i < nscales; i++) jobs[i] = PtJob{m.d_params, scales[i]};
    TRY(pt_build_jobs(ctx, jobs, m.tables, m.d_status, m.status));
    m.nscales = nscales;
    m.scales.assign(scales, scales + nscales);
    tab_release(m);
    m.cherry_built.assign(nscales, 0);  // the cherry tables belonged to the previous P(t)
    return PCSF_OK;
}

// Shared tail of the evaluation entry points: spans + P sets -> K2/K3 -> K4 -> host
int eval_spans(pcsf_ctx* ctx, const std::vector<Span>& spans, const std::vector<PSet>& psets, int64_t out_cols,
               const std::vector<int64_t>& seg_b, const std::vector<int64_t>& seg_e, double* out_lpr, double* out_elpr_anc) {
    const int64_t n_evals = (int64_t)seg_b.size();
    TRY(run_prune(ctx, spans, psets, out_cols));
    TRY(reserve(ctx, ctx->d_seg_begin, sizeof(int64_t) * std::max<int64_t>(n_evals, 1)));
    TRY(reserve(ctx, ctx->d_seg_end, sizeof(int64_t) * std::max<int64_t>(n_evals, 1)));
    if (n_evals > 0) {
        CU(cudaMemcpyAsync(ctx->d_seg_begin.p, seg_b.data(), sizeof(int64_t) * n_evals, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_seg_end.p, seg_e.data(), sizeof(int64_t) * n_evals, cudaMemcpyHostToDevice, ctx->stream));
    }
    TRY(run_reduce(ctx, n_evals));
    if (n_evals > 0) {
        CU(cudaMemcpyAsync(out_lpr, ctx->d_lpr.p, sizeof(double) * n_evals, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_elpr_anc) CU(cudaMemcpyAsync(out_elpr_anc, ctx->d_elpr.p, sizeof(double) * n_evals, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    TRY(finish_timing(ctx));
    ctx->last_all_models = 0;
    return PCSF_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* pcsf_version(void) { return PCSF_VERSION_STRING; }

int pcsf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pcsf_create(int device_id, pcsf_ctx** out) {
    if (!out) return PCSF_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device_id < 0 || device_id >= n) {
        cudaGetLastError();
        return PCSF_ERR_CUDA;  // no CPU fallback by design
    }
    pcsf_ctx* ctx = new pcsf_ctx();
    ctx->device = device_id;
    auto bail = [&](cudaError_t er) {
        (void)er;
        delete ctx;
        return PCSF_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return bail(e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return bail(e);
    ctx->num_sms = prop.multiProcessorCount;
    ctx->prune_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (const char* e = getenv("PCSF_SKEW_NS")) ctx->skew_ns = atoi(e);  // tuning knob, see prune_kernel
    if (const char* e = getenv("PCSF_WIDE")) ctx->wide = atoi(e);
    if (const char* e = getenv("PCSF_CHERRY_TABLES")) ctx->cherry_mode = atoi(e);
    if (const char* e = getenv("PCSF_RESCALE")) ctx->rescale = atoi(e) ? 1 : 0;
    if (const char* e = getenv("PCSF_SHARE_TABLES")) ctx->share_tables = atoi(e) ? 1 : 0;
    if (const char* e = getenv("PCSF_K6_PLAIN")) ctx->k6_plain = atoi(e) ? 1 : 0;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    ctx->stream = ctx->own_stream;
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e);
    *out = ctx;
    return PCSF_OK;
}

void pcsf_destroy(pcsf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    auto fr = [](DevBuf& b) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
    };
    for (auto& m : ctx->models) {
        tab_release(m);
        if (m.d_params && m.owns_params) cudaFree(m.d_params);
        fr(m.tables);
        fr(m.d_scales);
        fr(m.d_status);
        fr(m.cherry);
    }
    DevBuf* bufs[] = {&ctx->d_region_off, &ctx->d_codes, &ctx->d_nt, &ctx->d_aln_off, &ctx->d_aln_len, &ctx->d_spans,
                      &ctx->d_psets, &ctx->d_out_logz, &ctx->d_out_anc, &ctx->d_seg_begin, &ctx->d_seg_end,
                      &ctx->d_lpr, &ctx->d_elpr, &ctx->d_gstack, &ctx->d_jobs, &ctx->d_batch_params, &ctx->d_pair_tables,
                      &ctx->d_pair_status, &ctx->d_qs, &ctx->d_gexp, &ctx->d_pair_cherry, &ctx->d_out_tree, &ctx->d_out_scratch, &ctx->d_out_gacc,
                      &ctx->d_out_post, &ctx->d_out_ecounts, &ctx->d_out_z, &ctx->d_out_nodes, &ctx->d_eig_cache, &ctx->d_eig_valid, &ctx->d_eig_slots,
                      &ctx->d_eig_sweeps, &ctx->pipe_nt[0], &ctx->pipe_nt[1], &ctx->pipe_aln_off[0],
                      &ctx->pipe_aln_off[1], &ctx->pipe_aln_len[0], &ctx->pipe_aln_len[1], &ctx->pipe_codes[0], &ctx->pipe_codes[1],
                      &ctx->pipe_roff[0], &ctx->pipe_roff[1]};
    for (auto* b : bufs) fr(*b);
    if (ctx->d_branch_len) cudaFree(ctx->d_branch_len);
    if (ctx->d_ops) cudaFree(ctx->d_ops);
    if (ctx->d_items) cudaFree(ctx->d_items);
    for (void* q : {(void*)ctx->d_ops_t, (void*)ctx->d_items_t, (void*)ctx->d_ops_t3, (void*)ctx->d_items_t3, (void*)ctx->d_ops_t4,
                    (void*)ctx->d_items_t4, (void*)ctx->d_subtabs, (void*)ctx->d_tab_off})
        if (q) cudaFree(q);
    for (auto& ev : ctx->ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* pcsf_last_error(const pcsf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pcsf_stream_set(pcsf_ctx* ctx, void* cuda_stream) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PCSF_OK;
}

int pcsf_option_set(pcsf_ctx* ctx, int option, int64_t value) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (option == PCSF_OPT_RESCALE) {
        ctx->rescale = value ? 1 : 0;
        return PCSF_OK;
    }
    if (option == PCSF_OPT_PRUNE_FORM) {
        if (value < 0 || value > 2) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_option_set: PCSF_OPT_PRUNE_FORM takes 0, 1 or 2");
        ctx->wide = value == 0 ? -1 : (int)value - 1;
        return PCSF_OK;
    }
    if (option == PCSF_OPT_CHERRY_TABLES) {
        if (value < 0 || value > 4) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_option_set: PCSF_OPT_CHERRY_TABLES takes 0 .. 4");
        ctx->cherry_mode = (int)value;
        return PCSF_OK;
    }
    return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_option_set: unknown option");
}

int pcsf_tree_set(pcsf_ctx* ctx, int n_leaves, const int32_t* children, const double* branch_len) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (n_leaves < 2 || !children || !branch_len) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_tree_set: need n_leaves >= 2");
    CU(cudaSetDevice(ctx->device));
    const int n = 2 * n_leaves - 1;
    std::vector<int> seen(n, 0);
    for (int i = n_leaves; i < n; i++)
        for (int k = 0; k < 2; k++) {
            const int c = children[2 * (i - n_leaves) + k];
            if (c < 0 || c >= i || seen[c]++) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_tree_set: children do not follow the T numbering (child < parent, each node one parent)");
        }
    for (int i = 0; i < n - 1; i++) {
        if (!seen[i]) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_tree_set: node without a parent");
        if (!(branch_len[i] >= 0.0)) return fail(ctx, PCSF_ERR_INVALID_ARG, "CamlPaml.PhyloModel.make: negative or NaN branch length");  // PhyloModel.ml:16
    }
    if (prune_smem_for(n_leaves + 1, 3 * n_leaves, n_leaves) > ctx->prune_smem_optin)
        return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_tree_set: too many leaves for the pruning kernel's shared memory");
    // every check comes before the context is touched: a call that fails leaves the previous tree in place
    const std::vector<int32_t> new_children(children, children + 2 * (n_leaves - 1));
    TreePrograms tp = build_tree_programs(n_leaves, new_children, getenv("PCSF_NO_KEEP") == nullptr);  // pcsf_program.hpp
    if (tp.max_levels > MAX_STACK_LEVELS) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_tree_set: tree needs more than 16 parked partials");
    ctx->n_leaves = 0;  // from here to the last upload the context has no tree (check_ready fails if an upload does)
    for (auto& m : ctx->models) { m.nscales = 0; m.cherry_built.clear(); tab_release(m); }  // tables belong to the previous tree
    ctx->tree_hash = fnv1a(branch_len, sizeof(double) * (n - 1), fnv1a(children, sizeof(int32_t) * 2 * (n_leaves - 1)));
    ctx->n_branches = n - 1;
    ctx->children = new_children;
    ctx->branch_len.assign(branch_len, branch_len + (n - 1));
    ctx->ops = std::move(tp.ops);
    ctx->items = std::move(tp.items);
    ctx->ops_t = std::move(tp.ops_t);
    ctx->items_t = std::move(tp.items_t);
    ctx->ops_t3 = std::move(tp.ops_t3);
    ctx->items_t3 = std::move(tp.items_t3);
    ctx->ops_t4 = std::move(tp.ops_t4);
    ctx->items_t4 = std::move(tp.items_t4);
    ctx->subtabs = std::move(tp.subtabs);
    ctx->n_tab2 = tp.n_tab2;
    ctx->n_tab3 = tp.n_tab3;
    ctx->n_tab4 = tp.n_tab4;
    ctx->n_gemm = tp.n_gemm;
    ctx->max_levels = tp.max_levels;
    const std::vector<long long>& tab_off = tp.tab_off;
    if (n_leaves > 0xfffe) { ctx->n_tab2 = ctx->n_tab3 = ctx->n_tab4 = 0; }  // leaf ids are packed into 16 bits, 0xffff = none
    CU(cudaStreamSynchronize(ctx->stream));
    for (void** q : {(void**)&ctx->d_branch_len, (void**)&ctx->d_ops, (void**)&ctx->d_items, (void**)&ctx->d_ops_t, (void**)&ctx->d_items_t,
                     (void**)&ctx->d_ops_t3, (void**)&ctx->d_items_t3, (void**)&ctx->d_ops_t4, (void**)&ctx->d_items_t4, (void**)&ctx->d_subtabs,
                     (void**)&ctx->d_tab_off}) {
        if (*q) CU(cudaFree(*q));
        *q = nullptr;
    }
    auto upload = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess || bytes == 0) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    CU(upload((void**)&ctx->d_ops_t, ctx->ops_t.data(), sizeof(Op) * ctx->ops_t.size()));
    CU(upload((void**)&ctx->d_items_t, ctx->items_t.data(), sizeof(Item) * ctx->items_t.size()));
    CU(upload((void**)&ctx->d_ops_t3, ctx->ops_t3.data(), sizeof(Op) * ctx->ops_t3.size()));
    CU(upload((void**)&ctx->d_items_t3, ctx->items_t3.data(), sizeof(Item) * ctx->items_t3.size()));
    CU(upload((void**)&ctx->d_ops_t4, ctx->ops_t4.data(), sizeof(Op) * ctx->ops_t4.size()));
    CU(upload((void**)&ctx->d_items_t4, ctx->items_t4.data(), sizeof(Item) * ctx->items_t4.size()));
    CU(upload((void**)&ctx->d_subtabs, ctx->subtabs.data(), sizeof(SubTab) * ctx->subtabs.size()));
    CU(upload((void**)&ctx->d_tab_off, tab_off.data(), sizeof(long long) * tab_off.size()));
    CU(cudaMalloc(&ctx->d_branch_len, sizeof(double) * (n - 1)));
    CU(cudaMalloc(&ctx->d_ops, sizeof(Op) * ctx->ops.size()));
    CU(cudaMalloc(&ctx->d_items, sizeof(Item) * std::max<size_t>(1, ctx->items.size())));
    CU(cudaMemcpy(ctx->d_items, ctx->items.data(), sizeof(Item) * ctx->items.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_branch_len, branch_len, sizeof(double) * (n - 1), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_ops, ctx->ops.data(), sizeof(Op) * ctx->ops.size(), cudaMemcpyHostToDevice));
    ctx->n_leaves = n_leaves;  // committed
    ctx->nregions = -1;        // a staged batch belonged to the previous tree's leaf set
    return PCSF_OK;
}

int pcsf_tree_n_leaves(const pcsf_ctx* ctx) { return ctx ? ctx->n_leaves : 0; }

int pcsf_model_set(pcsf_ctx* ctx, int model_id, const double* S, const double* Sinv, const double* lambda,
                   const double* prior) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (model_id < 0 || model_id > (1 << 20) || !S || !Sinv || !lambda || !prior)
        return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_model_set: bad argument");
    CU(cudaSetDevice(ctx->device));
    if ((int)ctx->models.size() <= model_id) ctx->models.resize(model_id + 1);
    Model& m = ctx->models[model_id];
    std::vector<double> h(8192 + 192);
    memcpy(h.data(), S, 4096 * 8);
    memcpy(h.data() + 4096, Sinv, 4096 * 8);
    memcpy(h.data() + 8192, lambda, 64 * 8);
    memcpy(h.data() + 8192 + 64, prior, 64 * 8);
    for (int i = 0; i < 64; i++) h[8192 + 128 + i] = log(prior[i]);  // anc_lprior, src/PhyloCSFModel.ml:74
    if (!m.owns_params) { m.d_params = nullptr; m.owns_params = true; }
    if (!m.d_params) CU(cudaMalloc(&m.d_params, h.size() * 8));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(m.d_params, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    m.set = true;
    m.nscales = 0;
    tab_release(m);
    m.params_hash = fnv1a(h.data(), sizeof(double) * (8192 + 64));
    if (m.params_hash == 0) m.params_hash = 1;
    return PCSF_OK;
}

int pcsf_pt_build(pcsf_ctx* ctx, int model_id, int nscales, const double* scales, int32_t* status) {
    TRY(check_ready(ctx, false));
    if (model_id < 0 || model_id >= (int)ctx->models.size() || !ctx->models[model_id].set)
        return fail(ctx, PCSF_ERR_STATE, "pcsf_pt_build: model not set");
    if (nscales < 1 || !scales) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_pt_build: need at least one scale");
    CU(cudaSetDevice(ctx->device));
    Model& m = ctx->models[model_id];
    TRY(pt_build_device(ctx, m, nscales, scales));
    bool bad = false;
    for (int i = 0; i < nscales; i++) {
        if (status) status[i] = m.status[i];
        bad |= m.status[i] != 0;
    }
    if (bad) return fail(ctx, PCSF_ERR_NUMERIC, "CamlPaml.Q.real_to_Pt: P(t) failed its checks for at least one scale (see status)");
    return PCSF_OK;
}

int pcsf_pt_get(pcsf_ctx* ctx, int model_id, int scale_idx, int branch, double* P_out) {
    TRY(check_ready(ctx, false));
    TRY(check_model(ctx, model_id, scale_idx));
    if (branch < 0 || branch >= ctx->n_branches || !P_out) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_pt_get: bad branch");
    CU(cudaSetDevice(ctx->device));
    std::vector<double> slot(PT_SLOT);
    const double* src = (const double*)ctx->models[model_id].tables.p + ((size_t)scale_idx * ctx->n_branches + branch) * PT_SLOT;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(slot.data(), src, PT_SLOT_BYTES, cudaMemcpyDeviceToHost));
    for (int a = 0; a < 64; a++)
        for (int b = 0; b < 64; b++)
            P_out[a * 64 + b] = branch < ctx->n_leaves ? slot[b * 64 + a] : slot[frag_index(a, b)];
    return PCSF_OK;
}

int pcsf_batch_upload(pcsf_ctx* ctx, int64_t nregions, const int64_t* region_off, const uint8_t* codes) {
    TRY(check_ready(ctx, false));
    if (nregions < 0 || !region_off || region_off[0] != 0) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_upload: bad region offsets");
    for (int64_t r = 0; r < nregions; r++)
        if (region_off[r + 1] < region_off[r]) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_upload: region offsets must be non-decreasing");
    const int64_t total = region_off[nregions];
    if (total > 0 && !codes) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_upload: null codes");
    CU(cudaSetDevice(ctx->device));
    TRY(reserve(ctx, ctx->d_codes, (size_t)std::max<int64_t>(total, 1) * ctx->n_leaves));
    TRY(reserve(ctx, ctx->d_region_off, sizeof(int64_t) * (nregions + 1)));
    CU(cudaEventRecord(ctx->ev[6], ctx->stream));
    if (total > 0) CU(cudaMemcpyAsync(ctx->d_codes.p, codes, (size_t)total * ctx->n_leaves, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_region_off.p, region_off, sizeof(int64_t) * (nregions + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev[7], ctx->stream));
    ctx->region_off.assign(region_off, region_off + nregions + 1);
    ctx->nregions = nregions;
    ctx->total_cols = total;
    CU(cudaStreamSynchronize(ctx->stream));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]));
    ctx->ms[3] = t;
    return PCSF_OK;
}

// K0 (pcsf_k0.cuh): pleaves for `nalign` staged alignments. Grid = (alignments, tiles of the longest one), dynamic shared
// memory = the runs of codes the tile contributes to each frame. d_nt and d_codes are device allocations (aligned); the kernel
// reads d_nt in whole words up to ceil(nt_bytes / 4), which reserve()'s head room covers.
static int launch_frame_codes(pcsf_ctx* ctx, cudaStream_t stream, void* d_nt, int64_t nt_bytes, const void* d_aln_off,
                              const void* d_aln_len, const void* d_roff, int64_t nalign, int max_len, int frames, void* d_codes) {
    if (nalign <= 0 || max_len < 3) return PCSF_OK;  // no codon anywhere
    const int tile = k0::choose_tile_pos(max_len, ctx->n_leaves, frames);
    const size_t smem = k0::smem_bytes(tile, ctx->n_leaves, frames);
    if (smem > (size_t)227 * 1024) return fail(ctx, PCSF_ERR_INVALID_ARG, "pleaves on the device: too many leaves for one shared-memory tile");
    if (smem > (size_t)48 * 1024)
        CU(cudaFuncSetAttribute(frame_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (nalign > INT32_MAX) return fail(ctx, PCSF_ERR_INVALID_ARG, "pleaves on the device: too many alignments in one batch");
    // the kernel reads the buffer's last word whole: give its bytes past the end a value (they never reach a code)
    if (nt_bytes & 3) CU(cudaMemsetAsync((uint8_t*)d_nt + nt_bytes, 0, 4 - (nt_bytes & 3), stream));
    const int64_t tiles = ((int64_t)max_len + tile - 1) / tile;
    const dim3 grid((unsigned)nalign, (unsigned)std::min<int64_t>(tiles, 65535));
    frame_codes_kernel<<<grid, k0::THREADS, smem, stream>>>((const uint8_t*)d_nt, nt_bytes, (const int64_t*)d_aln_off,
                                                                  (const int32_t*)d_aln_len, (const int64_t*)d_roff, frames,
                                                                  ctx->n_leaves, tile, (uint8_t*)d_codes);
    CU(cudaGetLastError());
    ctx->launches++;
    return PCSF_OK;
}

int pcsf_batch_upload_alignments(pcsf_ctx* ctx, int64_t nalign, const int64_t* aln_off, const int32_t* aln_len,
                                 const uint8_t* nt, int frames) {
    int64_t nt_bytes = 0;
    if (ctx && nalign > 0 && aln_off && aln_len)
        for (int64_t a = 0; a < nalign; a++)
            nt_bytes = std::max<int64_t>(nt_bytes, aln_off[a] + (int64_t)std::max(aln_len[a], 0) * ctx->n_leaves);
    if (nt_bytes > 0 && !nt) return fail(ctx, PCSF_ERR_INVALID_ARG, "null nucleotide buffer");
    return pcsf_batch_upload_alignments_parts(ctx, nalign, aln_off, aln_len, nt_bytes > 0 ? 1 : 0, &nt, &nt_bytes, frames);
}

int pcsf_batch_upload_alignments_parts(pcsf_ctx* ctx, int64_t nalign, const int64_t* aln_off, const int32_t* aln_len,
                                       int64_t nparts, const uint8_t* const* part_ptr, const int64_t* part_bytes, int frames) {
    TRY(check_ready(ctx, false));
    if (nalign < 0 || !aln_off || !aln_len || nparts < 0 || (nparts > 0 && (!part_ptr || !part_bytes)) ||
        (frames != 1 && frames != 3 && frames != 6))
        return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_upload_alignments: bad argument");
    int64_t have = 0;
    for (int64_t i = 0; i < nparts; i++) {
        if (part_bytes[i] < 0 || (part_bytes[i] > 0 && !part_ptr[i])) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_upload_alignments: bad part");
        have += part_bytes[i];
    }
    CU(cudaSetDevice(ctx->device));
    const int64_t nregions = nalign * frames;
    std::vector<int64_t> roff(nregions + 1, 0);
    int64_t nt_bytes = 0;
    int max_len = 0;
    for (int64_t a = 0; a < nalign; a++) {
        if (aln_len[a] < 0) return fail(ctx, PCSF_ERR_INVALID_ARG, "negative alignment length");
        if (aln_off[a] < 0) return fail(ctx, PCSF_ERR_INVALID_ARG, "negative alignment offset");
        nt_bytes = std::max<int64_t>(nt_bytes, aln_off[a] + (int64_t)aln_len[a] * ctx->n_leaves);
        max_len = std::max(max_len, aln_len[a]);
        for (int f = 0; f < frames; f++) {
            const int rem = aln_len[a] - (f % 3);
            roff[a * frames + f + 1] = roff[a * frames + f] + (rem >= 3 ? rem / 3 : 0);  // pos+2 <= hi, PhyloCSF.ml:226
        }
    }
    const int64_t total = roff[nregions];
    if (nt_bytes > have) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_upload_alignments: an alignment lies outside the nucleotide buffer");
    if (nparts > 1) {  // an alignment must lie inside one piece (the pieces are copied back to back, so it would still be read correctly, but the contract says so)
        std::vector<int64_t> ends(nparts);
        int64_t at = 0;
        for (int64_t i = 0; i < nparts; i++) ends[i] = (at += part_bytes[i]);
        for (int64_t a = 0; a < nalign; a++) {
            const int64_t nb = (int64_t)aln_len[a] * ctx->n_leaves;
            if (nb == 0) continue;
            const int64_t i = std::upper_bound(ends.begin(), ends.end(), aln_off[a]) - ends.begin();
            if (i >= nparts || aln_off[a] + nb > ends[i]) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_upload_alignments_parts: an alignment straddles two pieces");
        }
    }
    TRY(reserve(ctx, ctx->d_nt, std::max<int64_t>(nt_bytes, 1)));
    TRY(reserve(ctx, ctx->d_aln_off, sizeof(int64_t) * std::max<int64_t>(nalign, 1)));
    TRY(reserve(ctx, ctx->d_aln_len, sizeof(int32_t) * std::max<int64_t>(nalign, 1)));
    TRY(reserve(ctx, ctx->d_codes, (size_t)std::max<int64_t>(total, 1) * ctx->n_leaves));
    TRY(reserve(ctx, ctx->d_region_off, sizeof(int64_t) * (nregions + 1)));
    CU(cudaEventRecord(ctx->ev[6], ctx->stream));
    {
        int64_t at = 0;
        for (int64_t i = 0; i < nparts && at < nt_bytes; i++) {
            const int64_t nb = std::min(part_bytes[i], nt_bytes - at);
            if (nb > 0) CU(cudaMemcpyAsync((uint8_t*)ctx->d_nt.p + at, part_ptr[i], nb, cudaMemcpyHostToDevice, ctx->stream));
            at += part_bytes[i];
        }
    }
    if (nalign > 0) {
        CU(cudaMemcpyAsync(ctx->d_aln_off.p, aln_off, sizeof(int64_t) * nalign, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_aln_len.p, aln_len, sizeof(int32_t) * nalign, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaMemcpyAsync(ctx->d_region_off.p, roff.data(), sizeof(int64_t) * (nregions + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev[7], ctx->stream));
    TRY(launch_frame_codes(ctx, ctx->stream, ctx->d_nt.p, nt_bytes, ctx->d_aln_off.p, ctx->d_aln_len.p, ctx->d_region_off.p, nalign, max_len,
                           frames, ctx->d_codes.p));
    ctx->region_off = roff;
    ctx->nregions = nregions;
    ctx->total_cols = total;
    CU(cudaStreamSynchronize(ctx->stream));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]));
    ctx->ms[3] = t;
    return PCSF_OK;
}

int pcsf_batch_codes_get(pcsf_ctx* ctx, uint8_t* codes_out) {
    TRY(check_ready(ctx, false));
    if (ctx->total_cols > 0 && !codes_out) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_batch_codes_get: null buffer");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->total_cols > 0)
        CU(cudaMemcpy(codes_out, ctx->d_codes.p, (size_t)ctx->total_cols * ctx->n_leaves, cudaMemcpyDeviceToHost));
    return PCSF_OK;
}

int64_t pcsf_batch_nregions(const pcsf_ctx* ctx) { return ctx ? ctx->nregions : -1; }
int64_t pcsf_batch_ncols(const pcsf_ctx* ctx) { return ctx ? ctx->total_cols : -1; }

int pcsf_lpr_all(pcsf_ctx* ctx, int n_models, const int32_t* model_ids, const int32_t* scale_idx, double* out_lpr,
                 double* out_elpr_anc, int32_t* out_status) {
    TRY(check_ready(ctx, true));
    if (n_models < 1 || !model_ids || !out_lpr) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_lpr_all: bad argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<Span> spans;
    std::vector<PSet> psets;
    for (int m = 0; m < n_models; m++) {
        const int sc = scale_idx ? scale_idx[m] : 0;
        TRY(check_model(ctx, model_ids[m], sc));
        {
            Model& mm = ctx->models[model_ids[m]];
            if (ctx->share_tables && mm.nscales == 1 && mm.params_hash && ctx->cherry_mode != 1 && ctx->wide != 0 && ctx->n_tab2 > 0)
                TRY(ensure_tables_shared(ctx, mm, ctx->total_cols));
            else
                TRY(ensure_tables(ctx, mm, sc, want_table_level(ctx, ctx->total_cols)));
        }
        psets.push_back(make_pset(ctx, model_ids[m], sc));
        spans.push_back(Span{0, (int64_t)m * ctx->total_cols, 0, (int32_t)0, m});
        spans.back().ncols = (int32_t)ctx->total_cols;
    }
    if (ctx->total_cols > 0x7fffffffLL) return fail(ctx, PCSF_ERR_INVALID_ARG, "batch too large: more than 2^31-1 codon columns");
    const int64_t n_segs = ctx->nregions * n_models;
    TRY(run_prune(ctx, spans, psets, ctx->total_cols * n_models));
    TRY(reserve(ctx, ctx->d_seg_begin, sizeof(int64_t) * std::max<int64_t>(n_segs, 1)));
    TRY(reserve(ctx, ctx->d_seg_end, sizeof(int64_t) * std::max<int64_t>(n_segs, 1)));
    if (n_segs > 0) {
        make_segments_kernel<<<(unsigned)((n_segs + 255) / 256), 256, 0, ctx->stream>>>(
            (const int64_t*)ctx->d_region_off.p, ctx->nregions, n_models, ctx->total_cols, (int64_t*)ctx->d_seg_begin.p,
            (int64_t*)ctx->d_seg_end.p);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    TRY(run_reduce(ctx, n_segs));
    CU(cudaEventRecord(ctx->ev[8], ctx->stream));
    if (n_segs > 0) {
        CU(cudaMemcpyAsync(out_lpr, ctx->d_lpr.p, sizeof(double) * n_segs, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_elpr_anc) CU(cudaMemcpyAsync(out_elpr_anc, ctx->d_elpr.p, sizeof(double) * n_segs, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev[9], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    TRY(finish_timing(ctx));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[8], ctx->ev[9]));
    ctx->ms[4] = t;
    ctx->last_all_models = n_models;
    if (out_status)
        for (int m = 0; m < n_models; m++) {
            const int st = ctx->models[model_ids[m]].status[scale_idx ? scale_idx[m] : 0];
            for (int64_t r = 0; r < ctx->nregions; r++)
                out_status[m * ctx->nregions + r] = st | (std::isfinite(out_lpr[m * ctx->nregions + r]) ? 0 : PCSF_ST_NOT_FINITE);
        }
    return PCSF_OK;
}

int pcsf_score_alignments(pcsf_ctx* ctx, int64_t nalign, const int64_t* aln_off, const int32_t* aln_len, const uint8_t* nt,
                          int frames, int n_models, const int32_t* model_ids, const int32_t* scale_idx, double* out_lpr,
                          double* out_elpr_anc, int32_t* out_status) {
    TRY(check_ready(ctx, false));
    if (nalign < 0 || !aln_off || !aln_len || (frames != 1 && frames != 3 && frames != 6) || n_models < 1 || !model_ids || !out_lpr)
        return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_score_alignments: bad argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<PSet> psets;
    int64_t est_cols = 0;
    for (int64_t a = 0; a < nalign; a++) est_cols += (int64_t)(aln_len[a] > 0 ? aln_len[a] / 3 : 0) * frames;
    for (int m = 0; m < n_models; m++) {
        const int sc = scale_idx ? scale_idx[m] : 0;
        TRY(check_model(ctx, model_ids[m], sc));
        {
            Model& mm = ctx->models[model_ids[m]];
            if (ctx->share_tables && mm.nscales == 1 && mm.params_hash && ctx->cherry_mode != 1 && ctx->wide != 0 && ctx->n_tab2 > 0)
                TRY(ensure_tables_shared(ctx, mm, est_cols));
            else
                TRY(ensure_tables(ctx, mm, sc, want_table_level(ctx, est_cols)));
        }
        psets.push_back(make_pset(ctx, model_ids[m], sc));
    }
    if (!ctx->copy_stream) {
        CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CU(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
        }
    }
    const int64_t R_total = nalign * frames;
    // chunks of whole alignments, ~2 M codon columns each (tens of ms of compute, a few ms of copy)
    int64_t chunk_cols = 2000000;
    if (const char* e = getenv("PCSF_CHUNK_COLS")) chunk_cols = std::max<int64_t>(1, atoll(e));  // tests force many chunks
    std::vector<int64_t> chunk_begin{0};
    {
        int64_t cols = 0;
        for (int64_t a = 0; a < nalign; a++) {
            if (aln_len[a] < 0) return fail(ctx, PCSF_ERR_INVALID_ARG, "negative alignment length");
            if (aln_off[a] < 0) return fail(ctx, PCSF_ERR_INVALID_ARG, "negative alignment offset");
            int64_t c = 0;
            for (int f = 0; f < frames; f++) { const int rem = aln_len[a] - (f % 3); c += rem >= 3 ? rem / 3 : 0; }
            // the first chunk is an eighth of the others: its host->device copy is the only one nothing overlaps
            const int64_t limit = chunk_begin.size() == 1 ? std::max<int64_t>(1, chunk_cols / 8) : chunk_cols;
            if (cols > 0 && cols + c > limit) { chunk_begin.push_back(a); cols = 0; }
            cols += c;
        }
        chunk_begin.push_back(nalign);
    }
    CU(cudaEventRecord(ctx->ev[6], ctx->stream));
    std::vector<int64_t> h_aoff[2], h_roff[2];
    for (size_t k = 0; k + 1 < chunk_begin.size(); k++) {
        const int b = (int)(k & 1);
        const int64_t a0 = chunk_begin[k], a1 = chunk_begin[k + 1], na = a1 - a0;
        if (na == 0) continue;
        int64_t lo = INT64_MAX, hi = 0;
        int max_len = 0;
        for (int64_t a = a0; a < a1; a++) {
            lo = std::min(lo, aln_off[a]);
            hi = std::max(hi, aln_off[a] + (int64_t)aln_len[a] * ctx->n_leaves);
            max_len = std::max(max_len, aln_len[a]);
        }
        if (hi > lo && !nt) return fail(ctx, PCSF_ERR_INVALID_ARG, "null nucleotide buffer");
        // the host-side staging vectors of buffer b may still feed an in-flight pageable copy of chunk k-2
        if (k >= 2) CU(cudaEventSynchronize(ctx->ev_copied[b]));
        h_aoff[b].resize(na);
        h_roff[b].assign(na * frames + 1, 0);
        for (int64_t a = a0; a < a1; a++) {
            h_aoff[b][a - a0] = aln_off[a] - lo;
            for (int f = 0; f < frames; f++) {
                const int rem = aln_len[a] - (f % 3);
                h_roff[b][(a - a0) * frames + f + 1] = h_roff[b][(a - a0) * frames + f] + (rem >= 3 ? rem / 3 : 0);
            }
        }
        const int64_t nreg = na * frames, total = h_roff[b][nreg];
        TRY(reserve(ctx, ctx->pipe_nt[b], std::max<int64_t>(hi - lo, 1)));
        TRY(reserve(ctx, ctx->pipe_aln_off[b], sizeof(int64_t) * na));
        TRY(reserve(ctx, ctx->pipe_aln_len[b], sizeof(int32_t) * na));
        TRY(reserve(ctx, ctx->pipe_roff[b], sizeof(int64_t) * (nreg + 1)));
        TRY(reserve(ctx, ctx->pipe_codes[b], (size_t)std::max<int64_t>(total, 1) * ctx->n_leaves));
        // ---- copy stream: stage chunk k once the kernels of chunk k-2 have released buffer b ----
        if (k >= 2) CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
        if (hi > lo) CU(cudaMemcpyAsync(ctx->pipe_nt[b].p, nt + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaMemcpyAsync(ctx->pipe_aln_off[b].p, h_aoff[b].data(), sizeof(int64_t) * na, cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaMemcpyAsync(ctx->pipe_aln_len[b].p, aln_len + a0, sizeof(int32_t) * na, cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaMemcpyAsync(ctx->pipe_roff[b].p, h_roff[b].data(), sizeof(int64_t) * (nreg + 1), cudaMemcpyHostToDevice, ctx->copy_stream));
        // pleaves (K0) rides on the copy stream behind its input: its CTAs cannot share an SM with the pruning kernel's (that
        // one takes the whole register file), so they run on the SMs the previous chunk's pruning launch frees as it drains -
        // inside its tail instead of between two pruning launches. Buffer b's codes were last read by the kernels of chunk k-2,
        // which the wait on ev_done[b] above covers.
        TRY(launch_frame_codes(ctx, ctx->copy_stream, ctx->pipe_nt[b].p, hi - lo, ctx->pipe_aln_off[b].p, ctx->pipe_aln_len[b].p,
                               ctx->pipe_roff[b].p, na, max_len, frames, ctx->pipe_codes[b].p));
        CU(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
        // ---- compute stream: pruning, reduction, results back ----
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
        std::vector<Span> spans;
        for (int m = 0; m < n_models; m++) spans.push_back(Span{0, (int64_t)m * total, 0, (int32_t)total, m});
        const int64_t n_segs = nreg * n_models;
        TRY(run_prune(ctx, spans, psets, total * n_models, (const uint8_t*)ctx->pipe_codes[b].p));
        TRY(reserve(ctx, ctx->d_seg_begin, sizeof(int64_t) * n_segs));
        TRY(reserve(ctx, ctx->d_seg_end, sizeof(int64_t) * n_segs));
        make_segments_kernel<<<(unsigned)((n_segs + 255) / 256), 256, 0, ctx->stream>>>(
            (const int64_t*)ctx->pipe_roff[b].p, nreg, n_models, total, (int64_t*)ctx->d_seg_begin.p, (int64_t*)ctx->d_seg_end.p);
        CU(cudaGetLastError());
        ctx->launches++;
        TRY(run_reduce(ctx, n_segs));
        for (int m = 0; m < n_models; m++) {
            CU(cudaMemcpyAsync(out_lpr + (size_t)m * R_total + a0 * frames, (const double*)ctx->d_lpr.p + (size_t)m * nreg,
                               sizeof(double) * nreg, cudaMemcpyDeviceToHost, ctx->stream));
            if (out_elpr_anc)
                CU(cudaMemcpyAsync(out_elpr_anc + (size_t)m * R_total + a0 * frames, (const double*)ctx->d_elpr.p + (size_t)m * nreg,
                                   sizeof(double) * nreg, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CU(cudaEventRecord(ctx->ev_done[b], ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev[7], ctx->stream));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]));
    ctx->ms[0] = t;  // whole pipelined pass
    ctx->last_all_models = 0;
    ctx->nregions = -1;  // no single staged batch remains
    if (out_status)
        for (int m = 0; m < n_models; m++) {
            const int st = ctx->models[model_ids[m]].status[scale_idx ? scale_idx[m] : 0];
            for (int64_t r = 0; r < R_total; r++)
                out_status[m * R_total + r] = st | (std::isfinite(out_lpr[m * R_total + r]) ? 0 : PCSF_ST_NOT_FINITE);
        }
    return PCSF_OK;
}

int pcsf_lpr(pcsf_ctx* ctx, int64_t n_evals, const int32_t* eval_model, const int32_t* eval_scale,
             const int64_t* eval_region, double* out_lpr, double* out_elpr_anc, int32_t* out_status) {
    TRY(check_ready(ctx, true));
    if (n_evals < 0 || !eval_model || !eval_region || !out_lpr) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_lpr: bad argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<Span> spans;
    std::vector<PSet> psets;
    std::vector<int64_t> seg_b(n_evals), seg_e(n_evals);
    spans.reserve(n_evals);
    psets.reserve(n_evals);
    int64_t out = 0;
    int prev_model = -1, prev_scale = -1;
    for (int64_t e = 0; e < n_evals; e++) {
        const int sc = eval_scale ? eval_scale[e] : 0;
        const int64_t r = eval_region[e];
        if (r < 0 || r >= ctx->nregions) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_lpr: region index out of range");
        if (eval_model[e] != prev_model || sc != prev_scale) {
            TRY(check_model(ctx, eval_model[e], sc));
            psets.push_back(make_pset(ctx, eval_model[e], sc));
            prev_model = eval_model[e];
            prev_scale = sc;
        }
        const int64_t c0 = ctx->region_off[r], nc = ctx->region_off[r + 1] - c0;
        seg_b[e] = out;
        seg_e[e] = out + nc;
        // merge with the previous span when it continues the same columns under the same P set
        if (!spans.empty() && spans.back().pset == (int32_t)psets.size() - 1 &&
            spans.back().col0 + spans.back().ncols == c0 && spans.back().out0 + spans.back().ncols == out &&
            (int64_t)spans.back().ncols + nc < 0x7fffffffLL) {
            spans.back().ncols += (int32_t)nc;
        } else {
            Span s{c0, out, 0, (int32_t)nc, (int32_t)psets.size() - 1};
            spans.push_back(s);
        }
        out += nc;
    }
    TRY(eval_spans(ctx, spans, psets, out, seg_b, seg_e, out_lpr, out_elpr_anc));
    if (out_status)
        for (int64_t e = 0; e < n_evals; e++)
            out_status[e] = ctx->models[eval_model[e]].status[eval_scale ? eval_scale[e] : 0] |
                            (std::isfinite(out_lpr[e]) ? 0 : PCSF_ST_NOT_FINITE);
    return PCSF_OK;
}

int pcsf_models_set(pcsf_ctx* ctx, int first_id, int n, const double* S, const double* Sinv, const double* lambda,
                    const double* prior) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (first_id < 0 || n < 1 || first_id + n > (1 << 22) || !S || !Sinv || !lambda || !prior)
        return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_models_set: bad argument");
    CU(cudaSetDevice(ctx->device));
    const size_t per = 8192 + 192;
    std::vector<double> h(per * n);
    for (int i = 0; i < n; i++) {
        double* d = h.data() + per * i;
        memcpy(d, S + (size_t)4096 * i, 4096 * 8);
        memcpy(d + 4096, Sinv + (size_t)4096 * i, 4096 * 8);
        memcpy(d + 8192, lambda + (size_t)64 * i, 64 * 8);
        memcpy(d + 8192 + 64, prior + (size_t)64 * i, 64 * 8);
        for (int k = 0; k < 64; k++) d[8192 + 128 + k] = log(prior[(size_t)64 * i + k]);
    }
    CU(cudaStreamSynchronize(ctx->stream));
    // models of the previous batch block become unset: the block is about to be reused
    for (auto& m : ctx->models)
        if (!m.owns_params) { m.set = false; m.d_params = nullptr; m.nscales = 0; }
    TRY(reserve(ctx, ctx->d_batch_params, h.size() * 8));
    CU(cudaMemcpy(ctx->d_batch_params.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    if ((int)ctx->models.size() < first_id + n) ctx->models.resize(first_id + n);
    for (int i = 0; i < n; i++) {
        Model& m = ctx->models[first_id + i];
        if (m.d_params && m.owns_params) CU(cudaFree(m.d_params));
        m.owns_params = false;
        m.d_params = (double*)ctx->d_batch_params.p + per * i;
        m.set = true;
        m.nscales = 0;
        m.params_hash = 0;  // not shareable: the block of parameters is rewritten per round
        tab_release(m);
    }
    ctx->pair_model.clear();
    return PCSF_OK;
}

int pcsf_omega_cache_reset(pcsf_ctx* ctx, int64_t n_slots) {
    if (!ctx || n_slots < 0) return PCSF_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    TRY(reserve(ctx, ctx->d_eig_cache, sizeof(double) * 4096 * (size_t)std::max<int64_t>(n_slots, 1)));
    TRY(reserve(ctx, ctx->d_eig_valid, sizeof(int32_t) * (size_t)std::max<int64_t>(n_slots, 1)));
    CU(cudaMemsetAsync(ctx->d_eig_valid.p, 0, sizeof(int32_t) * (size_t)std::max<int64_t>(n_slots, 1), ctx->stream));
    ctx->eig_slots = n_slots;
    return PCSF_OK;
}

int pcsf_omega_models_set(pcsf_ctx* ctx, int first_id, int n, const double* q_settings, int32_t* status) {
    return pcsf_omega_models_set_cached(ctx, first_id, n, q_settings, nullptr, status);
}

int pcsf_omega_models_set_cached(pcsf_ctx* ctx, int first_id, int n, const double* q_settings, const int64_t* cache_slot, int32_t* status) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (cache_slot)
        for (int i = 0; i < n; i++)
            if (cache_slot[i] >= ctx->eig_slots) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_omega_models_set_cached: cache slot out of range (pcsf_omega_cache_reset)");
    if (first_id < 0 || n < 1 || first_id + n > (1 << 22) || !q_settings)
        return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_omega_models_set: bad argument");
    for (int i = 0; i < n * 12; i++)
        if (!(q_settings[i] >= 0.0)) return fail(ctx, PCSF_ERR_INVALID_ARG, "CamlPaml.P14n.instantiate_q: domain violation");  // Fit.NonNeg
    CU(cudaSetDevice(ctx->device));
    const size_t per = 8192 + 192;
    CU(cudaStreamSynchronize(ctx->stream));
    for (auto& m : ctx->models)
        if (!m.owns_params) { m.set = false; m.d_params = nullptr; m.nscales = 0; }
    TRY(reserve(ctx, ctx->d_batch_params, per * 8 * (size_t)n));
    TRY(reserve(ctx, ctx->d_qs, sizeof(double) * 12 * (size_t)n));
    TRY(reserve(ctx, ctx->d_pair_status, sizeof(int32_t) * (size_t)n));
    CU(cudaMemcpyAsync(ctx->d_qs.p, q_settings, sizeof(double) * 12 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_pair_status.p, 0, sizeof(int32_t) * (size_t)n, ctx->stream));
    const int smem = (2 * 64 * 65 + 64 + 64 + 32 + 32) * 8;
    CU(cudaFuncSetAttribute(omega_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    const int64_t* d_slots = nullptr;
    if (cache_slot) {
        TRY(reserve(ctx, ctx->d_eig_slots, sizeof(int64_t) * (size_t)n));
        CU(cudaMemcpyAsync(ctx->d_eig_slots.p, cache_slot, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        d_slots = (const int64_t*)ctx->d_eig_slots.p;
    }
    TRY(reserve(ctx, ctx->d_eig_sweeps, sizeof(int32_t) * (size_t)n));
    omega_eig_kernel<<<n, EIG_THREADS, smem, ctx->stream>>>((const double*)ctx->d_qs.p, (double*)ctx->d_batch_params.p,
                                                            (int32_t*)ctx->d_pair_status.p, d_slots, (double*)ctx->d_eig_cache.p,
                                                            (int32_t*)ctx->d_eig_valid.p, (int32_t*)ctx->d_eig_sweeps.p);
    CU(cudaGetLastError());
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev[5], ctx->stream));
    std::vector<int32_t> st(n), sw(n);
    CU(cudaMemcpyAsync(st.data(), ctx->d_pair_status.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(sw.data(), ctx->d_eig_sweeps.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]));
    ctx->ms[2] = t;
    ctx->ms_total[6] += t;
    for (int i = 0; i < n; i++) ctx->counters[3] += sw[i];
    ctx->counters[2] += n;
    if ((int)ctx->models.size() < first_id + n) ctx->models.resize(first_id + n);
    bool bad = false;
    for (int i = 0; i < n; i++) {
        Model& m = ctx->models[first_id + i];
        if (m.d_params && m.owns_params) CU(cudaFree(m.d_params));
        m.owns_params = false;
        m.d_params = (double*)ctx->d_batch_params.p + per * i;
        m.set = true;  // a failed model keeps its slot (contents undefined); the caller sees its status
        m.nscales = 0;
        m.params_hash = 0;
        tab_release(m);
        if (status) status[i] = st[i];
        bad |= st[i] != 0;
    }
    ctx->pair_model.clear();
    if (bad) return fail(ctx, PCSF_ERR_NUMERIC, "omega rate matrix could not be scaled or diagonalised for at least one model (see status)");
    return PCSF_OK;
}

int pcsf_model_get(pcsf_ctx* ctx, int model_id, double* S, double* Sinv, double* lambda, double* prior) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (model_id < 0 || model_id >= (int)ctx->models.size() || !ctx->models[model_id].set)
        return fail(ctx, PCSF_ERR_STATE, "pcsf_model_get: model not set");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const double* d = ctx->models[model_id].d_params;
    if (S) CU(cudaMemcpy(S, d, 4096 * 8, cudaMemcpyDeviceToHost));
    if (Sinv) CU(cudaMemcpy(Sinv, d + 4096, 4096 * 8, cudaMemcpyDeviceToHost));
    if (lambda) CU(cudaMemcpy(lambda, d + 8192, 64 * 8, cudaMemcpyDeviceToHost));
    if (prior) CU(cudaMemcpy(prior, d + 8192 + 64, 64 * 8, cudaMemcpyDeviceToHost));
    return PCSF_OK;
}

int pcsf_pt_build_pairs(pcsf_ctx* ctx, int64_t npairs, const int32_t* pair_model, const double* pair_scale, int32_t* status) {
    TRY(check_ready(ctx, false));
    if (npairs < 1 || !pair_model || !pair_scale) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_pt_build_pairs: bad argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<PtJob> jobs(npairs);
    for (int64_t i = 0; i < npairs; i++) {
        const int mid = pair_model[i];
        if (mid < 0 || mid >= (int)ctx->models.size() || !ctx->models[mid].set)
            return fail(ctx, PCSF_ERR_STATE, "pcsf_pt_build_pairs: model " + std::to_string(mid) + " not set");
        jobs[i] = PtJob{ctx->models[mid].d_params, pair_scale[i]};
    }
    ctx->pair_model.clear();
    TRY(pt_build_jobs(ctx, jobs, ctx->d_pair_tables, ctx->d_pair_status, ctx->pair_status));
    ctx->pair_model.assign(pair_model, pair_model + npairs);
    bool bad = false;
    for (int64_t i = 0; i < npairs; i++) {
        if (status) status[i] = ctx->pair_status[i];
        bad |= ctx->pair_status[i] != 0;
    }
    if (bad) return fail(ctx, PCSF_ERR_NUMERIC, "CamlPaml.Q.real_to_Pt: P(t) failed its checks for at least one pair (see status)");
    return PCSF_OK;
}

int pcsf_lpr_pairs(pcsf_ctx* ctx, int64_t n_evals, const int64_t* eval_pair, const int64_t* eval_region, double* out_lpr,
                   double* out_elpr_anc, int32_t* out_status) {
    TRY(check_ready(ctx, true));
    if (n_evals < 0 || !eval_pair || !eval_region || !out_lpr) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_lpr_pairs: bad argument");
    CU(cudaSetDevice(ctx->device));
    const int64_t npairs = (int64_t)ctx->pair_model.size();
    std::vector<Span> spans;
    std::vector<PSet> psets;
    std::vector<int64_t> seg_b(n_evals), seg_e(n_evals);
    spans.reserve(n_evals);
    psets.reserve(n_evals);
    int64_t out = 0, prev_pair = -1;
    for (int64_t e = 0; e < n_evals; e++) {
        const int64_t pr = eval_pair[e], r = eval_region[e];
        if (pr < 0 || pr >= npairs) return fail(ctx, PCSF_ERR_STATE, "pcsf_lpr_pairs: pair index out of range (pcsf_pt_build_pairs)");
        if (r < 0 || r >= ctx->nregions) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_lpr_pairs: region index out of range");
        if (pr != prev_pair) {
            const Model& m = ctx->models[ctx->pair_model[pr]];
            PSet ps;
            ps.tables = (const double*)ctx->d_pair_tables.p + (size_t)pr * ctx->n_branches * PT_SLOT;
            ps.prior = m.prior();
            ps.logprior = m.logprior();
            ps.cherry = nullptr;
            ps.tab_level = 0;
            psets.push_back(ps);
            prev_pair = pr;
        }
        const int64_t c0 = ctx->region_off[r], nc = ctx->region_off[r + 1] - c0;
        seg_b[e] = out;
        seg_e[e] = out + nc;
        if (!spans.empty() && spans.back().pset == (int32_t)psets.size() - 1 && spans.back().col0 + spans.back().ncols == c0 &&
            spans.back().out0 + spans.back().ncols == out && (int64_t)spans.back().ncols + nc < 0x7fffffffLL) {
            spans.back().ncols += (int32_t)nc;
        } else {
            spans.push_back(Span{c0, out, 0, (int32_t)nc, (int32_t)psets.size() - 1});
        }
        out += nc;
    }
    // A round of candidates that all regions share (the bracket ends, the starting point, Brent's first golden-section
    // point, find_init's fixed random stream) scores the whole batch under a handful of P sets: worth their cherry
    // tables, built into a scratch block that lives until the next call.
    if (!psets.empty() && psets.size() <= 8 && ctx->cherry_mode != 1 && ctx->wide != 0 && ctx->n_tab2 > 0) {
        std::vector<int64_t> cols(psets.size(), 0);
        for (const Span& sp : spans) cols[sp.pset] += sp.ncols;
        bool all_long = true;
        for (int64_t c : cols) all_long = all_long && (ctx->cherry_mode >= 2 || c >= 50000);
        if (all_long) {
            const size_t block = table_block_doubles(ctx, 2);
            TRY(reserve(ctx, ctx->d_pair_cherry, sizeof(double) * block * psets.size()));
            const long long warps = (long long)ctx->n_tab2 * ((CHERRY_ROWS + 15) / 16);
            for (size_t i = 0; i < psets.size(); i++) {
                double* base = (double*)ctx->d_pair_cherry.p + i * block;
                subtree_table_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(psets[i].tables, ctx->d_subtabs, 0, ctx->n_tab2, CHERRY_ROWS, base);
                CU(cudaGetLastError());
                ctx->launches++;
                psets[i].cherry = base;
                psets[i].tab_level = 2;
            }
        }
    }
    TRY(eval_spans(ctx, spans, psets, out, seg_b, seg_e, out_lpr, out_elpr_anc));
    if (out_status)
        for (int64_t e = 0; e < n_evals; e++)
            out_status[e] = ctx->pair_status[eval_pair[e]] | (std::isfinite(out_lpr[e]) ? 0 : PCSF_ST_NOT_FINITE);
    return PCSF_OK;
}

// K6. Outside algorithm over the staged batch (PhyloLik.ml:96-180); see outside_kernel.
int pcsf_posteriors(pcsf_ctx* ctx, int model_id, int scale_idx, int n_nodes, const int32_t* nodes, double* out_node_post,
                    double* out_ecounts, double* out_z) {
    TRY(check_ready(ctx, true));
    TRY(check_model(ctx, model_id, scale_idx));
    const int nl = ctx->n_leaves, n = 2 * nl - 1, ni = nl - 1, nbr = n - 1;
    if (n_nodes < 0 || (n_nodes > 0 && (!nodes || !out_node_post))) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_posteriors: bad node list");
    for (int q = 0; q < n_nodes; q++)
        if (nodes[q] < 0 || nodes[q] >= n) return fail(ctx, PCSF_ERR_INVALID_ARG, "CamlPaml.PhyloLik.node_posterior: node index out of range");
    CU(cudaSetDevice(ctx->device));
    const int64_t total = ctx->total_cols;
    const Model& m = ctx->models[model_id];
    if (m.status[scale_idx] != 0) return fail(ctx, PCSF_ERR_NUMERIC, "pcsf_posteriors: the P set failed its checks (see pcsf_pt_build)");
    std::vector<int32_t> tree(2 * (nl - 1) + 2 * n + 3 * (n - 1), -1);  // children | parent | sibling | steps of the outside pass
    int32_t* par = tree.data() + 2 * (nl - 1);
    int32_t* sib = par + n;
    for (int i = nl; i < n; i++) {
        const int l = ctx->children[2 * (i - nl)], r = ctx->children[2 * (i - nl) + 1];
        tree[2 * (i - nl)] = l;
        tree[2 * (i - nl) + 1] = r;
        par[l] = par[r] = i;
        sib[l] = r;
        sib[r] = l;
    }
    {
        // outside pass of the DMMA form, depth first: under each node its leaf children, then an internal child whose beta is
        // parked, last the internal child the walk descends into with beta in registers (OUT_STEP_*, pcsf_kernels.cuh)
        int32_t* steps = sib + n;
        int ns = 0;
        std::vector<std::pair<int, bool>> stack{{n - 1, false}};  // (node, its beta is in registers)
        while (!stack.empty()) {
            const auto [v, carried] = stack.back();
            stack.pop_back();
            int kids[2] = {tree[2 * (v - nl)], tree[2 * (v - nl) + 1]};
            if (kids[0] >= nl && kids[1] < nl) std::swap(kids[0], kids[1]);  // leaves first
            for (int q = 0; q < 2; q++) {
                const int c = kids[q];
                int fl = (q == 0 && !carried) ? OUT_STEP_LOADB : 0;
                if (c >= nl) {
                    const bool last = q == 1;
                    fl |= last ? OUT_STEP_CARRY : OUT_STEP_STOREB;
                    if (n_nodes > 0) fl |= OUT_STEP_STOREB;
                }
                steps[3 * ns] = c;
                steps[3 * ns + 1] = fl;
                ns++;
            }
            if (kids[0] >= nl) stack.push_back({kids[0], false});  // popped after the carried subtree is done
            if (kids[1] >= nl) stack.push_back({kids[1], true});
        }
        for (int k = n - 2, next = -1; k >= 0; k--) {  // the first internal node at or after each step
            if (steps[3 * k] >= nl) next = steps[3 * k];
            steps[3 * k + 2] = next;
        }
    }
    const int64_t n_tiles = (total + OUT_TC - 1) / OUT_TC;
    // two CTAs per SM (~105 KB of shared memory each): the walk is a chain of short products separated by barriers, so
    // co-resident CTAs are what hides its latencies
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, 2LL * ctx->num_sms));
    static_assert(OUT_TC == OD_TC, "both forms of K6 walk 64-column tiles and share the scratch blocks");
    const bool plain = ctx->k6_plain != 0;  // PCSF_K6_PLAIN=1: the plain-FP64 form (A/B runs, differential tests)
    TRY(reserve(ctx, ctx->d_out_tree, tree.size() * sizeof(int32_t)));
    TRY(reserve(ctx, ctx->d_out_scratch, sizeof(double) * (size_t)grid * 3 * ni * OUT_TC * 64));
    if (out_ecounts) {
        TRY(reserve(ctx, ctx->d_out_gacc, sizeof(double) * (size_t)grid * nbr * 4096));
        TRY(reserve(ctx, ctx->d_out_ecounts, sizeof(double) * (size_t)nbr * 4096));
    }
    if (n_nodes > 0) {
        TRY(reserve(ctx, ctx->d_out_post, sizeof(double) * (size_t)n_nodes * std::max<int64_t>(total, 1) * 64));
        TRY(reserve(ctx, ctx->d_out_nodes, sizeof(int32_t) * n_nodes));
        CU(cudaMemcpyAsync(ctx->d_out_nodes.p, nodes, sizeof(int32_t) * n_nodes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (out_z) TRY(reserve(ctx, ctx->d_out_z, sizeof(double) * std::max<int64_t>(total, 1)));
    if (!plain) TRY(reserve(ctx, ctx->d_out_images, sizeof(double) * (size_t)std::max(1, nl - 2) * 4096));
    CU(cudaMemcpyAsync(ctx->d_out_tree.p, tree.data(), tree.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    OutsideParams p;
    memset(&p, 0, sizeof(p));
    p.tables = (const double*)m.tables.p + (size_t)scale_idx * ctx->n_branches * PT_SLOT;
    p.prior = m.prior();
    p.codes = (const uint8_t*)ctx->d_codes.p;
    p.total_cols = total;
    p.n_leaves = nl;
    p.children = (const int32_t*)ctx->d_out_tree.p;
    p.parent = p.children + 2 * (nl - 1);
    p.sibling = p.parent + n;
    p.scratch = (double*)ctx->d_out_scratch.p;
    p.gacc = out_ecounts ? (double*)ctx->d_out_gacc.p : nullptr;
    p.n_post = n_nodes;
    p.post_nodes = (const int32_t*)ctx->d_out_nodes.p;
    p.post_out = (double*)ctx->d_out_post.p;
    p.z_out = out_z ? (double*)ctx->d_out_z.p : nullptr;
    p.pt_images = (const double*)ctx->d_out_images.p;
    p.steps = p.sibling + n;
    const int smem = plain ? (4096 + 2 * OUT_TC * OUT_XS + OUT_TC) * (int)sizeof(double) + r16(OUT_TC * nl) : OD_SMEM_FIXED + od_tree_ints(nl) * 2 + r16(OD_TC * nl);
    if (smem > ctx->prune_smem_optin) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_posteriors: tree too large for the kernel's shared memory");
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (plain) {
        CU(cudaFuncSetAttribute(outside_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        outside_kernel<<<grid, OUT_THREADS, smem, ctx->stream>>>(p);
    } else {
        if (nl > 2) {
            outside_transpose_kernel<<<(unsigned)(((size_t)(nl - 2) * 4096 + 255) / 256), 256, 0, ctx->stream>>>(p.tables, nl, nl - 2, (double*)ctx->d_out_images.p);
            CU(cudaGetLastError());
            ctx->launches++;
        }
        CU(cudaFuncSetAttribute(outside_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        outside_dmma_kernel<<<grid, OD_THREADS, smem, ctx->stream>>>(p);
    }
    CU(cudaGetLastError());
    ctx->launches++;
    if (out_ecounts) {
        outside_reduce_kernel<<<(unsigned)(((size_t)nbr * 4096 + 255) / 256), 256, 0, ctx->stream>>>(
            (const double*)ctx->d_out_gacc.p, grid, nbr, nl, p.tables, (double*)ctx->d_out_ecounts.p, plain ? 0 : 1);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (out_ecounts) CU(cudaMemcpyAsync(out_ecounts, ctx->d_out_ecounts.p, sizeof(double) * (size_t)nbr * 4096, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_nodes > 0 && total > 0)
        CU(cudaMemcpyAsync(out_node_post, ctx->d_out_post.p, sizeof(double) * (size_t)n_nodes * total * 64, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_z && total > 0) CU(cudaMemcpyAsync(out_z, ctx->d_out_z.p, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float t;
    CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
    ctx->ms[0] = t;
    return PCSF_OK;
}

int pcsf_column_terms(pcsf_ctx* ctx, int m, double* col_logz, double* col_anc) {
    TRY(check_ready(ctx, true));
    if (m < 0 || m >= ctx->last_all_models) return fail(ctx, PCSF_ERR_STATE, "pcsf_column_terms: no matching pcsf_lpr_all result");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t off = (size_t)m * ctx->total_cols, n = (size_t)ctx->total_cols;
    if (col_logz && n) CU(cudaMemcpy(col_logz, (const double*)ctx->d_out_logz.p + off, n * 8, cudaMemcpyDeviceToHost));
    if (col_anc && n) CU(cudaMemcpy(col_anc, (const double*)ctx->d_out_anc.p + off, n * 8, cudaMemcpyDeviceToHost));
    return PCSF_OK;
}

// -------------------------------------------------------------------------------------------------
// Batched maximize_lpr. Every region runs the reference's find_init + GSL Brent state machine
// (pcsf_brent.hpp); each round gathers one candidate rho per live region, builds their P tables and
// scores them in one K1 + K2/K3 + K4 sequence. Candidates that are the same for every region
// (lo, hi, init, Brent's first golden-section point, find_init's fixed random stream) share one
// P set and run as a single span over the whole batch.
// -------------------------------------------------------------------------------------------------
int pcsf_maximize_lpr_multi(pcsf_ctx* ctx, int n_models, const int32_t* model_ids, double init, double lo, double hi,
                            double accuracy, double* out_rho, double* out_lpr, double* out_elpr_anc, int32_t* out_status,
                            int32_t* out_nevals) {
    TRY(check_ready(ctx, true));
    if (n_models < 1 || !model_ids) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_maximize_lpr: no models");
    for (int m = 0; m < n_models; m++)
        if (model_ids[m] < 0 || model_ids[m] >= (int)ctx->models.size() || !ctx->models[model_ids[m]].set)
            return fail(ctx, PCSF_ERR_STATE, "pcsf_maximize_lpr: model not set");
    if (!out_rho || !out_lpr) return fail(ctx, PCSF_ERR_INVALID_ARG, "pcsf_maximize_lpr: null output");
    if (lo >= hi || lo <= 0.0) return fail(ctx, PCSF_ERR_INVALID_ARG, "CamlPaml.Fit.find_init");  // Fit.ml:28
    CU(cudaSetDevice(ctx->device));
    const int64_t R = ctx->nregions, N = R * n_models;
    std::vector<MaximizeLpr> st((size_t)N, MaximizeLpr(init, lo, hi, accuracy));  // [model][region]
    // candidates of a round, model-major so that candidates common to all regions of a model share one P set
    std::vector<int64_t> live, e_pair, e_region;
    std::vector<double> xs, pair_scale, r_lpr, r_elpr;
    std::vector<int32_t> pair_model, r_status, p_status;
    // bound the P tables of one launch sequence to ~16 GiB
    const size_t pset_bytes = (size_t)ctx->n_branches * PT_SLOT_BYTES;
    const int64_t max_sets = std::max<int64_t>(1, (int64_t)((16ull << 30) / pset_bytes));
    double ms_prune = 0, ms_reduce = 0, ms_pt = 0;
    for (;;) {
        live.clear();
        xs.clear();
        for (int64_t i = 0; i < N; i++)
            if (!st[i].done()) {
                live.push_back(i);
                xs.push_back(st[i].candidate());
            }
        if (live.empty()) break;
        size_t c0 = 0;
        while (c0 < live.size()) {
            pair_model.clear();
            pair_scale.clear();
            e_pair.clear();
            e_region.clear();
            size_t c1 = c0;
            for (; c1 < live.size(); c1++) {
                const int32_t mid = model_ids[live[c1] / R];
                const double x = xs[c1];
                if (pair_model.empty() || pair_model.back() != mid || pair_scale.back() != x) {
                    if ((int64_t)pair_model.size() >= max_sets) break;
                    pair_model.push_back(mid);
                    pair_scale.push_back(x);
                }
                e_pair.push_back((int64_t)pair_model.size() - 1);
                e_region.push_back(live[c1] % R);
            }
            const int64_t np = (int64_t)pair_model.size(), ne = (int64_t)e_region.size();
            p_status.assign(np, 0);
            const int rc = pcsf_pt_build_pairs(ctx, np, pair_model.data(), pair_scale.data(), p_status.data());
            if (rc != PCSF_OK && rc != PCSF_ERR_NUMERIC) return rc;
            ms_pt += ctx->ms[2];
            r_lpr.resize(ne);
            r_elpr.resize(ne);
            r_status.resize(ne);
            TRY(pcsf_lpr_pairs(ctx, ne, e_pair.data(), e_region.data(), r_lpr.data(), r_elpr.data(), r_status.data()));
            ms_prune += ctx->ms[0];
            ms_reduce += ctx->ms[1];
            for (int64_t i = 0; i < ne; i++) st[live[c0 + i]].feed(r_lpr[i], r_elpr[i], r_status[i]);
            c0 = c1;
        }
    }
    bool bad = false;
    for (int64_t i = 0; i < N; i++) {
        out_rho[i] = st[i].result_x;
        out_lpr[i] = st[i].result_f;
        if (out_elpr_anc) out_elpr_anc[i] = st[i].result_elpr;
        if (out_status) out_status[i] = st[i].status;
        if (out_nevals) out_nevals[i] = st[i].nevals;
        bad |= (st[i].status & ~PCSF_ST_RANDOM_INIT) != 0;
    }
    ctx->ms[0] = ms_prune;
    ctx->ms[1] = ms_reduce;
    ctx->ms[2] = ms_pt;
    if (bad) return fail(ctx, PCSF_ERR_NUMERIC, "maximize_lpr failed for at least one region (see status)");
    return PCSF_OK;
}

int pcsf_maximize_lpr(pcsf_ctx* ctx, int model_id, double init, double lo, double hi, double accuracy,
                      double* out_rho, double* out_lpr, double* out_elpr_anc, int32_t* out_status,
                      int32_t* out_nevals) {
    const int32_t mid = model_id;
    return pcsf_maximize_lpr_multi(ctx, 1, &mid, init, lo, hi, accuracy, out_rho, out_lpr, out_elpr_anc, out_status, out_nevals);
}

#ifdef PCSF_TIMELINE
// debug builds only (tools/timeline.py): per-warp (op, clock64) marks of CTA 0
int pcsf_debug_timeline(pcsf_ctx* ctx, int cap, long long* out) {
    if (!ctx) return PCSF_ERR_INVALID_ARG;
    if (!out) {
        if (ctx->timeline) cudaFree(ctx->timeline);
        ctx->timeline = nullptr;
        ctx->timeline_cap = cap;
        CU(cudaMalloc(&ctx->timeline, (size_t)cap * 16 * 16));
        CU(cudaMemset(ctx->timeline, 0xff, (size_t)cap * 16 * 16));
        return PCSF_OK;
    }
    CU(cudaMemcpy(out, ctx->timeline, (size_t)ctx->timeline_cap * 16 * 16, cudaMemcpyDeviceToHost));
    return PCSF_OK;
}
#endif

int pcsf_table_level(const pcsf_ctx* ctx, int model_id, int scale_idx) {
    if (!ctx || model_id < 0 || model_id >= (int)ctx->models.size() || !ctx->models[model_id].set) return PCSF_ERR_INVALID_ARG;
    const Model& m = ctx->models[model_id];
    if (scale_idx < 0 || scale_idx >= m.nscales) return PCSF_ERR_INVALID_ARG;
    if (m.tab) return ctx->cherry_mode == 1 ? 0 : std::min(m.tab->level, ctx->cherry_mode == 2 ? 3 : ctx->cherry_mode == 3 ? 2 : 4);
    return scale_idx < (int)m.cherry_built.size() ? (int)m.cherry_built[scale_idx] : 0;
}

int64_t pcsf_last_launch_info(const pcsf_ctx* ctx, int which) {
    if (!ctx) return -1;
    switch (which) {
        case 0: return ctx->last_form;
        case 1: return ctx->last_level;
        case 2: return ctx->last_tiles;
        case 3: return ctx->last_grid;
        default: return -1;
    }
}

double pcsf_last_ms(const pcsf_ctx* ctx, int which) {
    if (!ctx || which < 0 || which > 5) return -1.0;
    return ctx->ms[which];
}

int64_t pcsf_counter(pcsf_ctx* ctx, int which) {
    if (!ctx) return -1;
    if (which < 0) {
        for (int64_t& v : ctx->counters) v = 0;
        return 0;
    }
    return which < 8 ? ctx->counters[which] : -1;
}

double pcsf_total_ms(pcsf_ctx* ctx, int which) {
    if (!ctx) return -1.0;
    if (which < 0) {  // reset
        for (double& v : ctx->ms_total) v = 0.0;
        return 0.0;
    }
    return which < 8 ? ctx->ms_total[which] : -1.0;
}

int64_t pcsf_launch_count(const pcsf_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
