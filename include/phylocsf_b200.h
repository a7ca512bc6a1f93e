/*
 * phylocsf_b200.h — C ABI of the B200-native PhyloCSF scoring path.
 *
 * The reference (mlin/PhyloCSF, OCaml + GSL) has no FFI of its own; the seam this library replaces
 * is the pair of closures
 *     PhyloCSFModel.lpr_leaves : instance -> leaf array array -> float -> {lpr_leaves; elpr_anc; inst}
 *                                                          (src/PhyloCSFModel.ml:67-82)
 *     OmegaModel.lpr_leaves    : instance -> leaf array array -> float   (src/OmegaModel.ml:137-144)
 * together with the CamlPaml calls underneath them: PhyloModel.make's P(t) loop
 * (lib/CamlPaml/PhyloModel.ml:12-25 -> Q.Diag.real_to_Pt, lib/CamlPaml/Q.ml:211-249) and the
 * pruning pass PhyloLik.prepare / likelihood / node_posterior (lib/CamlPaml/PhyloLik.ml:46-138).
 * INTEGRATION.md shows the OCaml C stubs a maintainer would add to bind these entry points.
 *
 * Conventions: every array argument is a caller-owned HOST pointer (pinned or pageable); matrices
 * are FP64 row-major; all entry points return PCSF_OK (0) or a negative PCSF_ERR_* code, with a
 * message available from pcsf_last_error(). One context per GPU; a context is not re-entrant.
 * There is no CPU fallback: pcsf_create fails if no CUDA device is usable.
 */
#ifndef PHYLOCSF_B200_H
#define PHYLOCSF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCSF_K 64                 /* Codon64.dim, lib/CamlPaml/Code.ml:135 */
#define PCSF_CODE_MARGINALIZE 64  /* `Marginalize (lib/CamlPaml/PhyloLik.ml:9); codes 0..63 = `Certain i */

/* Return codes. The reference raises OCaml exceptions at the same points. */
#define PCSF_OK 0
#define PCSF_ERR_INVALID_ARG (-1) /* Invalid_argument (PhyloLik.ml:51-58, PhyloModel.ml:14-23, Q.ml:212) */
#define PCSF_ERR_CUDA (-2)        /* CUDA runtime / driver error, or no device */
#define PCSF_ERR_STATE (-3)       /* call order violated (e.g. lpr before tree_set / batch_upload) */
#define PCSF_ERR_NUMERIC (-4)     /* a Failure of the numeric path; see the status arrays */
#define PCSF_ERR_NOMEM (-5)

/* Bits of the per-P-set / per-evaluation status words (0 = fine). */
#define PCSF_ST_NEG_T 1        /* t < 0                          -> Invalid_argument, Q.ml:212 */
#define PCSF_ST_NEG_ENTRY 2    /* expm entry < -1e-6             -> Failure, Q.ml:235-236 */
#define PCSF_ST_ROWSUM 4       /* |row sum - 1| > 1e-6           -> Failure, Q.ml:243-244 */
#define PCSF_ST_DIAG_ASSERT 8  /* not (0 < new diagonal <= 1)    -> Assert_failure, Q.ml:245 */
#define PCSF_ST_NOT_FINITE 16  /* lpr is -inf/nan (z underflowed to 0; the reference prints -inf/nan) */
#define PCSF_ST_BRACKET 32     /* Brent: endpoints do not enclose a minimum -> Gsl_exn via Fit.ml:13 */
#define PCSF_ST_RANDOM_INIT 64 /* Fit.find_init took its random branch (Fit.ml:33-41); informational */

typedef struct pcsf_ctx pcsf_ctx;

const char *pcsf_version(void);
int pcsf_device_count(void);

/* Context = one GPU: stream, resident P tables, staged batch.  (No reference analogue.) */
int pcsf_create(int device_id, pcsf_ctx **out);
void pcsf_destroy(pcsf_ctx *ctx);
const char *pcsf_last_error(const pcsf_ctx *ctx);
/* Run all work of this context on an existing cudaStream_t (so a caller can bracket calls with its
 * own CUDA events). NULL restores the context's own stream. */
int pcsf_stream_set(pcsf_ctx *ctx, void *cuda_stream);

/*
 * Options (no reference analogue). PCSF_OPT_RESCALE = 1: rescue per-column partials from underflow by
 * exact powers of two and add the exponents back into log z. The reference never rescales
 * (lib/CamlPaml/PhyloLik.ml:87-92) and returns log 0 = -inf for such columns, which is what the default
 * (0) reproduces; with the option on, columns that stay clear of 2^-256 are computed exactly as before.
 * The environment variable PCSF_RESCALE=1 sets it for every new context (used by the command line).
 */
#define PCSF_OPT_RESCALE 1
/*
 * PCSF_OPT_PRUNE_FORM: which form of the pruning kernel a launch uses. 0 (default) = chosen per launch from
 * the lengths of its spans; 1 = narrow (128-column tiles, two compute warps per SM sub-partition);
 * 2 = wide (192-column tiles, three). Both forms compute every column with the same arithmetic in the same
 * order: results are bit-identical (tested). The environment variable PCSF_WIDE=0/1 forces narrow/wide
 * for every new context.
 */
#define PCSF_OPT_PRUNE_FORM 2
/*
 * PCSF_OPT_CHERRY_TABLES: memoise, per P set, the partial likelihood above every cherry of the tree over the
 * 65 x 65 code pairs of its two leaves (2.16 MB per cherry); one level up, above every cherry-and-leaf subtree over
 * the 65^3 code triples (140.6 MB each); and above every caterpillar of four leaves over the 65^4 quadruples (9.14 GB
 * each). The tables are computed with the pruning kernel's own instruction sequence, so a lookup is bit-identical
 * to the computation it replaces. 0 (default) = built by pcsf_lpr_all / pcsf_score_alignments when the wide form runs
 * and a P set scores >= 50,000 columns (cherries) / >= 1,000,000 (3 leaves) / >= 5,000,000 (4 leaves) in one call, or
 * has scored eight times that in total (a level pays for itself over a run, not only over one batch), the larger
 * ones only while device memory allows; 1 = never; 2 = always, up to 3 leaves; 3 = always, cherries only;
 * 4 = always, up to 4 leaves. PCSF_CHERRY_TABLES in the environment sets it for every new context.
 * The tables of a one-scale model set with pcsf_model_set are shared by all contexts of the process on the same GPU
 * that hold the same tree, model and scale (PCSF_SHARE_TABLES=0 in the environment: one private copy per context).
 */
#define PCSF_OPT_CHERRY_TABLES 3
int pcsf_option_set(pcsf_ctx *ctx, int option, int64_t value);
/*
 * What is actually in effect (no reference analogue; lets tests and bench.py state which kernel program produced a
 * number instead of re-deriving the library's heuristics). pcsf_table_level: the subtree-table level (0 none, 2, 3, 4)
 * the P set (model_id, scale_idx) carries right now - it can be lower than requested, because the large levels are
 * skipped when device memory is short - or a negative PCSF_ERR_* for an unknown model / scale.
 * pcsf_last_launch_info: the most recent pruning launch of the context: which = 0 kernel form (1 narrow, 2 wide),
 * 1 table level of the tree program it ran (0, 2, 3, 4), 2 number of tiles, 3 grid size (CTAs).
 */
int pcsf_table_level(const pcsf_ctx *ctx, int model_id, int scale_idx);
int64_t pcsf_last_launch_info(const pcsf_ctx *ctx, int which);

/*
 * Tree shape = T.t (lib/CamlPaml/T.mli:3, T.ml:57-112): leaves are nodes 0..n_leaves-1 in
 * left-to-right order, internal nodes follow in post-order, root = 2*n_leaves-2.
 *   children[2*(i-n_leaves)+{0,1}] = (left,right) child of internal node i; both < i.
 *   branch_len[i], i < 2*n_leaves-2 = length of the branch above node i (T.branch), >= 0.
 */
int pcsf_tree_set(pcsf_ctx *ctx, int n_leaves, const int32_t *children, const double *branch_len);
/* Leaves of the tree currently set (0 = none). */
int pcsf_tree_n_leaves(const pcsf_ctx *ctx);

/*
 * A diagonalised rate matrix = Q.Diag.t, real path (Q.ml:96-141): Q = S * diag(lambda) * Sinv, plus
 * the root prior the caller wants (PhyloModel.prior: the Q equilibrium when prior=None,
 * PhyloModel.ml:30-32 / Q.ml:153-177). model_id is a small non-negative slot number chosen by the
 * caller (e.g. 0 = coding ECM, 1 = noncoding ECM); setting a slot again replaces it.
 */
int pcsf_model_set(pcsf_ctx *ctx, int model_id, const double *S, const double *Sinv, const double *lambda,
                   const double *prior);

/*
 * K1. P(t) for every branch and every tree scale: P[scale][br] = real_to_Pt(Q, scales[scale] *
 * branch_len[br]) (PhyloModel.ml:17 looping Q.ml:211-249, with the clamp / row-sum / diagonal
 * fix-ups and tol = 1e-6). status[scale] (optional) receives PCSF_ST_* bits. Returns
 * PCSF_ERR_NUMERIC if any status is non-zero (the tables are still usable for the others).
 * The tables stay resident on the device until the next pt_build of the same model.
 */
int pcsf_pt_build(pcsf_ctx *ctx, int model_id, int nscales, const double *scales, int32_t *status);

/* PhyloModel.p (PhyloModel.ml:29): copy one built P(t) back, 64x64 row-major (row = parent state,
 * column = child state, PhyloLik.mli:17). For tests and diagnostics. */
int pcsf_pt_get(pcsf_ctx *ctx, int model_id, int scale_idx, int branch, double *P_out);

/*
 * Stage a batch of regions = many `leaf array array`s (the output of pleaves, src/PhyloCSF.ml:219-246).
 *   region_off[r] .. region_off[r+1]-1 = codon columns of region r (nregions+1 offsets, region_off[0]=0);
 *   codes[col * n_leaves + leaf]       = 0..63 (`Certain) or PCSF_CODE_MARGINALIZE.
 * Empty regions are allowed (their lpr is 0, like an empty Array.iter). Replaces any previous batch.
 */
int pcsf_batch_upload(pcsf_ctx *ctx, int64_t nregions, const int64_t *region_off, const uint8_t *codes);

/*
 * Stage a batch of alignments and let the device do pleaves: one region per requested frame.
 *   nt[a]: alignment a = n_leaves rows (tree-leaf order; absent species = a row of '-') of aln_len[a]
 *   bytes each, at nt + aln_off[a]; bytes are the alignment characters after u->t.
 *   frames = 1, 3 or 6 (AsIs candidate_regions, src/PhyloCSF.ml:198-205). Regions are numbered
 *   alignment-major, then frame (+0,+1,+2,-0,-1,-2).
 *   Preconditions checked: aln_off[a] >= 0, aln_len[a] >= 0. NOT checkable here: the buffer at nt must
 *   extend to max_a(aln_off[a] + aln_len[a] * n_leaves) bytes (the _parts form takes explicit sizes and checks).
 */
int pcsf_batch_upload_alignments(pcsf_ctx *ctx, int64_t nalign, const int64_t *aln_off, const int32_t *aln_len,
                                 const uint8_t *nt, int frames);
/* The same with the nucleotide buffer given as `nparts` host pieces; aln_off[] are offsets into their
 * concatenation (an alignment must lie inside one piece). Lets a multi-threaded reader hand over the
 * buffers its threads filled without first copying them into one. */
int pcsf_batch_upload_alignments_parts(pcsf_ctx *ctx, int64_t nalign, const int64_t *aln_off, const int32_t *aln_len,
                                       int64_t nparts, const uint8_t *const *part_ptr, const int64_t *part_bytes,
                                       int frames);
/* The staged batch's leaf codes, codes_out[col * n_leaves + leaf] for pcsf_batch_ncols() columns: what pleaves
 * (src/PhyloCSF.ml:219-246) returned for each region, back on the host - after pcsf_batch_upload_alignments this is
 * the output of the device pleaves (K0), which the parity tests compare byte for byte with the reference's. */
int pcsf_batch_codes_get(pcsf_ctx *ctx, uint8_t *codes_out);
int64_t pcsf_batch_nregions(const pcsf_ctx *ctx);
int64_t pcsf_batch_ncols(const pcsf_ctx *ctx);

/*
 * K2-K4. lpr_leaves for every region of the staged batch under each listed (model, scale):
 *   out_lpr[m*nregions + r]      = sum over columns of log z          (PhyloCSFModel.ml:79)
 *   out_elpr_anc[m*nregions + r] = sum over columns of post_root . log prior (PhyloCSFModel.ml:80-81)
 * scale_idx may be NULL (= scale 0 of each model). out_elpr_anc / out_status may be NULL.
 */
int pcsf_lpr_all(pcsf_ctx *ctx, int n_models, const int32_t *model_ids, const int32_t *scale_idx,
                 double *out_lpr, double *out_elpr_anc, int32_t *out_status);

/*
 * pcsf_batch_upload_alignments + pcsf_lpr_all in one call, pipelined: the alignments are processed in
 * chunks of ~2 M codon columns, and the host->device copy of chunk k+1 runs on a second stream while
 * chunk k is being scored. out_* are indexed [m * (nalign*frames) + region], regions numbered as in
 * pcsf_batch_upload_alignments. No batch remains staged afterwards.
 */
int pcsf_score_alignments(pcsf_ctx *ctx, int64_t nalign, const int64_t *aln_off, const int32_t *aln_len,
                          const uint8_t *nt, int frames, int n_models, const int32_t *model_ids,
                          const int32_t *scale_idx, double *out_lpr, double *out_elpr_anc, int32_t *out_status);

/*
 * The same for an explicit list of evaluations: evaluation e scores region eval_region[e] under
 * model eval_model[e] with P tables of scale index eval_scale[e]. This is the shape one round of
 * batched maximize_lpr candidates has (one candidate rho per (region, model)).
 */
int pcsf_lpr(pcsf_ctx *ctx, int64_t n_evals, const int32_t *eval_model, const int32_t *eval_scale,
             const int64_t *eval_region, double *out_lpr, double *out_elpr_anc, int32_t *out_status);

/*
 * Many models at once (the omega strategy has one rate matrix per region and per kappa candidate,
 * src/OmegaModel.ml:166-170). pcsf_models_set fills slots first_id .. first_id+n-1 from arrays of n
 * consecutive S, Sinv (64x64), lambda, prior (64) blocks; the block lives until the next
 * pcsf_models_set call, which unsets those slots. pcsf_pt_build_pairs builds one P set per
 * (model, scale) pair in one K1 launch sequence (replacing the previous pairs); pcsf_lpr_pairs scores
 * evaluation e = region eval_region[e] under pair eval_pair[e].
 */
int pcsf_models_set(pcsf_ctx *ctx, int first_id, int n, const double *S, const double *Sinv, const double *lambda,
                    const double *prior);
/*
 * K5. The omega model's Q assembly and diagonalisation on the device: model first_id+i is built from
 * q_settings[12*i .. 12*i+11] = kappa, omega, sigma, 9 x F3x4 ratios (src/OmegaModel.ml:21-80, then
 * Q.Diag.of_Q + equilibrium, lib/CamlPaml/Q.ml:124-177). Same slot/lifetime rules as pcsf_models_set.
 * status bits: 128 = Q scale non-positive (PhyloModel.ml:99-100), 256 = no zero eigenvalue (Q.ml:168-169).
 */
int pcsf_omega_models_set(pcsf_ctx *ctx, int first_id, int n, const double *q_settings, int32_t *status);
/*
 * The same with warm starts for the diagonalisation. A search over kappa (src/OmegaModel.ml:177-188) diagonalises a
 * sequence of nearby rate matrices per region; cache_slot[i] >= 0 names a per-region slot (0 .. n_slots-1 of the last
 * pcsf_omega_cache_reset) that holds the eigenvectors of the region's previous candidate: the Jacobi sweeps then start
 * from them (2-4 sweeps instead of ~9) and leave the new eigenvectors in the slot. cache_slot[i] < 0, a NULL cache_slot
 * or an empty slot = cold start. Slots within one call must be distinct. The eigensystem is that of the same matrix to
 * the same stopping threshold; against a cold start it differs by rounding only.
 * pcsf_omega_cache_reset empties all slots (call it when a new search begins: bounds the accumulation of rotations).
 */
int pcsf_omega_cache_reset(pcsf_ctx *ctx, int64_t n_slots);
int pcsf_omega_models_set_cached(pcsf_ctx *ctx, int first_id, int n, const double *q_settings, const int64_t *cache_slot,
                                 int32_t *status);
/* Read a model slot back (S, Sinv 64x64; lambda, prior 64); any output may be NULL. For tests. */
int pcsf_model_get(pcsf_ctx *ctx, int model_id, double *S, double *Sinv, double *lambda, double *prior);
int pcsf_pt_build_pairs(pcsf_ctx *ctx, int64_t npairs, const int32_t *pair_model, const double *pair_scale,
                        int32_t *status);
int pcsf_lpr_pairs(pcsf_ctx *ctx, int64_t n_evals, const int64_t *eval_pair, const int64_t *eval_region,
                   double *out_lpr, double *out_elpr_anc, int32_t *out_status);

/* Per-column terms of the most recent pcsf_lpr_all call for its m-th listed model:
 * col_logz[c] = log z(c), col_anc[c] = post_root(c) . log prior (either may be NULL). */
int pcsf_column_terms(pcsf_ctx *ctx, int m, double *col_logz, double *col_anc);

/*
 * K6. Outside algorithm over the staged batch under the P set (model_id, scale_idx): PhyloLik.ensure_beta,
 * node_posterior and add_branch_posteriors (lib/CamlPaml/PhyloLik.ml:96-180) - the E step of PhyloEM-style training;
 * the command line's scoring strategies do not need it.
 *   nodes[n_nodes]: nodes (T numbering: leaves first, root = 2 n_leaves - 2) whose posterior is wanted;
 *   out_node_post[(q * total_cols + col) * 64 + x] = P(node nodes[q] in state x | column col): alpha_x beta_x / z for an
 *     internal node, the leaf vector itself for a leaf (one-hot, or all ones when marginalised), all zeros when z = 0
 *     (PhyloLik.ml:131-138). May be NULL when n_nodes = 0. Mind the size: 512 bytes per (node, column).
 *   out_ecounts[(br * 64 + a) * 64 + b] (optional) = sum over all columns with z > 0 of
 *     beta_parent[a] (P_sib alpha_sib)[a] P_br[a][b] alpha_br[b] / z, br = 0 .. 2 n_leaves - 3: branch_posteriors summed over
 *     the batch with weight 1 (PhyloLik.ml:140-180) - the expected number of a -> b substitutions on branch br.
 *   out_z[col] (optional) = likelihood of every column (PhyloLik.likelihood).
 * Sums over columns are formed per CTA and then over CTAs in a fixed order (deterministic); they differ from the
 * reference's column-by-column accumulation by rounding only.
 */
int pcsf_posteriors(pcsf_ctx *ctx, int model_id, int scale_idx, int n_nodes, const int32_t *nodes, double *out_node_post,
                    double *out_ecounts, double *out_z);

/*
 * Batched PhyloCSFModel.maximize_lpr (src/PhyloCSFModel.ml:84-99): for every region of the batch,
 * Fit.find_init (lib/CamlPaml/Fit.ml:27-48) then GSL Brent on -lpr(rho) until (ub-lb)/x <= accuracy,
 * then the final re-evaluation f(x). Each round builds P(t) for all live candidates (K1) and scores
 * them (K2-K4) in one launch sequence. out_* have nregions entries; out_nevals (optional) counts
 * likelihood evaluations per region.
 */
int pcsf_maximize_lpr(pcsf_ctx *ctx, int model_id, double init, double lo, double hi, double accuracy,
                      double *out_rho, double *out_lpr, double *out_elpr_anc, int32_t *out_status,
                      int32_t *out_nevals);

/* The same for several models at once (llr_MaxLik maximises the coding and the noncoding model,
 * src/PhyloCSFModel.ml:130-136): all (region, model) searches advance in the same rounds, so every
 * launch sequence carries n_models times the work. out_* are indexed [m * nregions + region]. */
int pcsf_maximize_lpr_multi(pcsf_ctx *ctx, int n_models, const int32_t *model_ids, double init, double lo, double hi,
                            double accuracy, double *out_rho, double *out_lpr, double *out_elpr_anc,
                            int32_t *out_status, int32_t *out_nevals);

/* Timing of the most recent call, measured with CUDA events on the context's stream:
 * which = 0 pruning kernel (K2+K3 fused), 1 region reduction (K4), 2 P(t) build (K1),
 * 3 H2D copies, 4 D2H copies; 5 = the most recent build of subtree tables (PCSF_OPT_CHERRY_TABLES; once per
 * P set). Returns milliseconds, or a negative value if not recorded. */
double pcsf_last_ms(const pcsf_ctx *ctx, int which);
/* Device time summed over every call since the context was created or last reset (which < 0 resets and returns 0):
 * which = 0 pruning (K2+K3), 1 region reduction (K4), 2 P(t) build (K1), 5 subtree tables, 6 omega Q assembly +
 * diagonalisation (K5). Lets a caller split a whole strategy run (mle, omega: hundreds of launch sequences) by kernel. */
double pcsf_total_ms(pcsf_ctx *ctx, int which);
/* Work counters since creation or last reset (which < 0 resets): 0 = P(t) slots built by K1 (branch x candidate),
 * 1 = codon-column evaluations pruned (columns x P sets), 2 = matrices diagonalised by K5, 3 = Jacobi sweeps they took,
 * 4 = pruning tiles. With pcsf_total_ms they give a strategy run's achieved FLOP/s per kernel. */
int64_t pcsf_counter(pcsf_ctx *ctx, int which);
/* Kernel launches issued by this context since creation (for bench.py's gpu_launches). */
int64_t pcsf_launch_count(const pcsf_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PHYLOCSF_B200_H */
