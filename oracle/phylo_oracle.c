/*
 * phylo_oracle.c — CPU restatement of the PhyloCSF scoring hot path (TEST INFRASTRUCTURE).
 *
 * This file is the parity oracle for the CUDA kernels in phylocsf_b200/csrc. It is NOT part of
 * the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it. The product path (phylocsf_b200) never calls into it and has no CPU fallback.
 *
 * It restates, loop for loop, the numerical core of the reference (OCaml + GSL, which cannot be
 * built here: no ocaml/opam/GSL in the image — see DESIGN.md):
 *
 *   oracle_real_to_Pt      <- lib/CamlPaml/Q.ml:211-249 (Diag.real_to_Pt, real path) with
 *                             diagm (Q.ml:53-66) and gemm (Q.ml:33-41 -> cblas_dgemm)
 *   oracle_ensure_alpha    <- lib/CamlPaml/PhyloLik.ml:73-93 (inside pass, dense ddot form,
 *                             explicit one-hot / all-ones leaf vectors of PhyloLik.ml:11-19)
 *   oracle_lpr_leaves      <- src/PhyloCSFModel.ml:67-82 (per-column loop: log z, root posterior
 *                             from PhyloLik.ml:127-138, dot with log prior)
 *   oracle_lpr_batch       <- the same, one region per OpenMP task (the reference's only
 *                             parallelism is process-level over regions, src/ForkYes.ml:5-8)
 *
 * Third-party arithmetic restated (GSL, unpinned version, not in /root/reference): cblas_ddot as a
 * sequential sum r += x[i]*y[i], i ascending; cblas_dgemm (RowMajor, NoTrans, NoTrans, beta=0) as
 * C[i][j] = sum_k A[i][k]*B[k][j] accumulated for k ascending starting from 0.
 *
 * Parity status: pinned to the reference's own known-answer tests to their stated windows
 * (lib/CamlPaml/test.ml:8-54 eps 1e-3, test.ml:81 lnL -1574.63623 +-1e-3, src/test.ml:27-59
 * +-0.005 dB); the 1e-6 dB bar between this oracle and the CUDA path is pinned by this file only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Status codes of oracle_real_to_Pt, mirroring the reference's exceptions. */
#define ORACLE_OK 0
#define ORACLE_ERR_NEG_T 1        /* Q.ml:212  invalid_arg */
#define ORACLE_ERR_NEG_ENTRY 2    /* Q.ml:235-236 failwith (entry < -tol) */
#define ORACLE_ERR_ROWSUM 3       /* Q.ml:243-244 failwith (|rowsum-1| > tol) */
#define ORACLE_ERR_DIAG_ASSERT 4  /* Q.ml:245 assert (0 < smii <= 1) */

/* Q.ml:211-249. S, Sinv, P are k x k row-major; lambda has k entries. */
int oracle_real_to_Pt(int k, const double *S, const double *Sinv, const double *lambda, double t,
                      double tol, double *P) {
    if (t < 0.) return ORACLE_ERR_NEG_T;
    double *expLt = (double *)malloc(sizeof(double) * k);
    double *DS = (double *)malloc(sizeof(double) * k * k);
    for (int i = 0; i < k; i++) expLt[i] = exp(t * lambda[i]); /* Q.ml:216-217 */
    /* diagm: row i of S' scaled by expLt[i]  (Q.ml:61-64) */
    for (int i = 0; i < k; i++)
        for (int j = 0; j < k; j++) DS[i * k + j] = Sinv[i * k + j] * expLt[i];
    /* gemm r_s (diagm ...)  (Q.ml:218): reference cblas_dgemm accumulation order */
    for (int i = 0; i < k * k; i++) P[i] = 0.;
    for (int kk = 0; kk < k; kk++)
        for (int i = 0; i < k; i++) {
            const double temp = S[i * k + kk];
            for (int j = 0; j < k; j++) P[i * k + j] += temp * DS[kk * k + j];
        }
    free(expLt);
    free(DS);
    /* fix-ups, Q.ml:226-247 */
    for (int i = 0; i < k; i++) {
        double tot = 0.;
        double smii = 1.;
        for (int j = 0; j < k; j++) {
            tot += P[i * k + j]; /* pre-clamp value */
            if (P[i * k + j] < 0.) {
                if (fabs(P[i * k + j]) > tol) return ORACLE_ERR_NEG_ENTRY;
                P[i * k + j] = 0.;
            }
            if (i != j) smii -= P[i * k + j]; /* post-clamp value */
        }
        if (fabs(tot - 1.) > tol) return ORACLE_ERR_ROWSUM;
        if (!(smii <= 1. && smii > 0.)) return ORACLE_ERR_DIAG_ASSERT;
        P[i * k + i] = smii;
    }
    return ORACLE_OK;
}

/* cblas_ddot, unit stride */
static inline double ddot(int k, const double *x, const double *y) {
    double r = 0.;
    for (int i = 0; i < k; i++) r += x[i] * y[i];
    return r;
}

/*
 * PhyloLik.ml:73-93 for one column.
 *   n_leaves, children[2*(i-n_leaves)+{0,1}] for internal node i (T.ml numbering: leaves first,
 *   internals in post-order, root last); pms[br] = k x k matrix of the branch above node br
 *   (row = parent state, col = child state, PhyloLik.mli:17); leaf code c < k => `Certain c,
 *   c >= k => `Marginalize (PhyloLik.ml:11-19).
 *   alpha: workspace (n_internal x k). leafvec: workspace (n_leaves x k).
 * Returns z = alpha_root . prior  (PhyloLik.ml:92).
 */
static double ensure_alpha(int n_leaves, const int32_t *children, const double *pms, const double *prior,
                           int k, const uint8_t *codes, double *alpha, double *leafvec) {
    const int n = 2 * n_leaves - 1;
    for (int l = 0; l < n_leaves; l++) {
        double *v = leafvec + (size_t)l * k;
        if (codes[l] < k) {
            for (int j = 0; j < k; j++) v[j] = 0.;
            v[codes[l]] = 1.;
        } else {
            for (int j = 0; j < k; j++) v[j] = 1.;
        }
    }
    for (int i = n_leaves; i < n; i++) {
        const int lc = children[2 * (i - n_leaves)], rc = children[2 * (i - n_leaves) + 1];
        const double *ls = pms + (size_t)lc * k * k;
        const double *rs = pms + (size_t)rc * k * k;
        const double *alc = lc < n_leaves ? leafvec + (size_t)lc * k : alpha + (size_t)(lc - n_leaves) * k;
        const double *arc = rc < n_leaves ? leafvec + (size_t)rc * k : alpha + (size_t)(rc - n_leaves) * k;
        double *ai = alpha + (size_t)(i - n_leaves) * k;
        for (int a = 0; a < k; a++) ai[a] = ddot(k, ls + (size_t)a * k, alc) * ddot(k, rs + (size_t)a * k, arc);
    }
    return ddot(k, alpha + (size_t)(n - 1 - n_leaves) * k, prior);
}

double oracle_ensure_alpha(int n_leaves, const int32_t *children, const double *pms, const double *prior,
                           int k, const uint8_t *codes, double *alpha_out /* n_internal*k */) {
    double *leafvec = (double *)malloc(sizeof(double) * (size_t)n_leaves * k);
    double z = ensure_alpha(n_leaves, children, pms, prior, k, codes, alpha_out, leafvec);
    free(leafvec);
    return z;
}

/*
 * src/PhyloCSFModel.ml:67-82 for one region: codes is [ncols][n_leaves] (one `leaf array` per
 * codon column, as produced by pleaves, src/PhyloCSF.ml:219-246).
 * Outputs lpr = sum log z, elpr_anc = sum_cols sum_x post_root[x] * log prior[x].
 * col_logz / col_anc (optional, may be NULL): the per-column terms.
 */
void oracle_lpr_leaves(int n_leaves, const int32_t *children, const double *pms, const double *prior, int k,
                       int64_t ncols, const uint8_t *codes, double *lpr_out, double *elpr_anc_out,
                       double *col_logz, double *col_anc) {
    const int n_internal = n_leaves - 1;
    double *alpha = (double *)malloc(sizeof(double) * (size_t)n_internal * k);
    double *leafvec = (double *)malloc(sizeof(double) * (size_t)n_leaves * k);
    double *anc_lprior = (double *)malloc(sizeof(double) * k);
    double *pr_root = (double *)malloc(sizeof(double) * k);
    for (int x = 0; x < k; x++) anc_lprior[x] = log(prior[x]); /* PhyloCSFModel.ml:74 */
    double lpr = 0., elpr = 0.;
    const double *aroot = alpha + (size_t)(n_internal - 1) * k;
    for (int64_t c = 0; c < ncols; c++) {
        const double z = ensure_alpha(n_leaves, children, pms, prior, k, codes + (size_t)c * n_leaves, alpha, leafvec);
        const double lz = log(z);
        lpr += lz; /* PhyloCSFModel.ml:79 */
        /* PhyloLik.ml:127-138, root */
        if (z == 0.) {
            for (int x = 0; x < k; x++) pr_root[x] = 0.;
        } else {
            for (int x = 0; x < k; x++) pr_root[x] = aroot[x] * prior[x] / z;
        }
        double d = 0.; /* PhyloCSFModel.ml:45-50 `dot` */
        for (int x = 0; x < k; x++) d += pr_root[x] * anc_lprior[x];
        elpr += d;
        if (col_logz) col_logz[c] = lz;
        if (col_anc) col_anc[c] = d;
    }
    *lpr_out = lpr;
    *elpr_anc_out = elpr;
    free(alpha);
    free(leafvec);
    free(anc_lprior);
    free(pr_root);
}

/*
 * Many regions, one model instance per "P set": region r uses pms + pset[r]*(2n-2)*k*k and
 * prior + pset_prior... (prior is per model, shared). Regions are independent (one OpenMP task
 * each), which is how the reference parallelises (-p N forks per region).
 * region_off has nregions+1 entries (column offsets into codes [total_cols][n_leaves]).
 */
void oracle_lpr_batch(int n_leaves, const int32_t *children, const double *pms, const int32_t *region_pset,
                      const double *prior, int k, int64_t nregions, const int64_t *region_off,
                      const uint8_t *codes, double *lpr_out, double *elpr_anc_out, int nthreads) {
    const size_t pset_stride = (size_t)(2 * n_leaves - 2) * k * k;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int64_t r = 0; r < nregions; r++) {
        const double *p = pms + (region_pset ? (size_t)region_pset[r] * pset_stride : 0);
        oracle_lpr_leaves(n_leaves, children, p, prior, k, region_off[r + 1] - region_off[r],
                          codes + (size_t)region_off[r] * n_leaves, lpr_out + r, elpr_anc_out + r, NULL, NULL);
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Outside algorithm and posteriors  <- lib/CamlPaml/PhyloLik.ml:96-180 (ensure_beta, node_posterior,
 * add_branch_posteriors), for one column. Parent / sibling come from the children array (T.parent, T.sibling).
 *   beta: (2*n_leaves-1) x k, row i = beta of node i; beta[root] = prior (PhyloLik.ml:62).
 *   For i = root-1 downto 0 (PhyloLik.ml:103-119):
 *     inter[a] = beta_p[a] * ddot(P_s[a,:], alpha_s)        p = parent(i), s = sibling(i)
 *     beta_i[b] = ddot(inter, P_i[:,b])
 * Pinned by the reference's own 2-state known answers (lib/CamlPaml/test.ml:8-54, eps 1e-3) in
 * tests/test_oracle_golden.py.
 */
static void ensure_beta(int n_leaves, const int32_t *children, const double *pms, const double *prior, int k,
                        const double *alpha, const double *leafvec, double *beta, double *inter) {
    const int n = 2 * n_leaves - 1;
    int *parent = (int *)malloc(sizeof(int) * n);
    int *sibling = (int *)malloc(sizeof(int) * n);
    for (int i = n_leaves; i < n; i++) {
        const int lc = children[2 * (i - n_leaves)], rc = children[2 * (i - n_leaves) + 1];
        parent[lc] = parent[rc] = i;
        sibling[lc] = rc;
        sibling[rc] = lc;
    }
    for (int a = 0; a < k; a++) beta[(size_t)(n - 1) * k + a] = prior[a];
    for (int i = n - 2; i >= 0; i--) {
        const int p = parent[i], s = sibling[i];
        const double *ps = pms + (size_t)i * k * k, *ss = pms + (size_t)s * k * k;
        const double *bp = beta + (size_t)p * k;
        const double *xas = s < n_leaves ? leafvec + (size_t)s * k : alpha + (size_t)(s - n_leaves) * k;
        for (int a = 0; a < k; a++) inter[a] = bp[a] * ddot(k, ss + (size_t)a * k, xas);
        for (int b = 0; b < k; b++) {
            double r = 0.;
            for (int a = 0; a < k; a++) r += inter[a] * ps[(size_t)a * k + b]; /* ddot inter ps_colb */
            beta[(size_t)i * k + b] = r;
        }
    }
    free(parent);
    free(sibling);
}

/*
 * PhyloLik.node_posterior for every node and PhyloLik.branch_posteriors for every branch of one column.
 *   node_post: (2*n_leaves-1) x k (may be NULL). z = 0 => zeros (PhyloLik.ml:131-132); leaf => its leaf vector (:133-134).
 *   ecounts:   (2*n_leaves-2) x k x k, ACCUMULATED with `weight` (add_branch_posteriors, :140-174; may be NULL):
 *              ecounts[br][a][b] += weight * beta_p[a] * ddot(P_sib[a,:], alpha_sib) * P_br[a][b] * alpha_br[b] / z,
 *              only for z > 0 and beta_p[a] > 0.
 * Returns z.
 */
double oracle_posteriors(int n_leaves, const int32_t *children, const double *pms, const double *prior, int k,
                         const uint8_t *codes, double weight, double *node_post, double *ecounts) {
    const int n = 2 * n_leaves - 1;
    double *alpha = (double *)malloc(sizeof(double) * (size_t)(n_leaves - 1) * k);
    double *leafvec = (double *)malloc(sizeof(double) * (size_t)n_leaves * k);
    double *beta = (double *)malloc(sizeof(double) * (size_t)n * k);
    double *inter = (double *)malloc(sizeof(double) * k);
    const double z = ensure_alpha(n_leaves, children, pms, prior, k, codes, alpha, leafvec);
    ensure_beta(n_leaves, children, pms, prior, k, alpha, leafvec, beta, inter);
    if (node_post) {
        for (int i = 0; i < n; i++) {
            const double *ai = i < n_leaves ? leafvec + (size_t)i * k : alpha + (size_t)(i - n_leaves) * k;
            for (int x = 0; x < k; x++) {
                double v;
                if (z == 0.) v = 0.;
                else if (i < n_leaves) v = ai[x];
                else v = ai[x] * beta[(size_t)i * k + x] / z;
                node_post[(size_t)i * k + x] = v;
            }
        }
    }
    if (ecounts && z > 0.) {
        int *parent = (int *)malloc(sizeof(int) * n);
        int *sibling = (int *)malloc(sizeof(int) * n);
        for (int i = n_leaves; i < n; i++) {
            const int lc = children[2 * (i - n_leaves)], rc = children[2 * (i - n_leaves) + 1];
            parent[lc] = parent[rc] = i;
            sibling[lc] = rc;
            sibling[rc] = lc;
        }
        for (int br = 0; br < n - 1; br++) {
            const int p = parent[br], sib = sibling[br];
            const double *sm = pms + (size_t)br * k * k, *sms = pms + (size_t)sib * k * k;
            const double *bp = beta + (size_t)p * k;
            const double *ab = br < n_leaves ? leafvec + (size_t)br * k : alpha + (size_t)(br - n_leaves) * k;
            const double *xas = sib < n_leaves ? leafvec + (size_t)sib * k : alpha + (size_t)(sib - n_leaves) * k;
            for (int a = 0; a < k; a++) {
                const double bpa = bp[a];
                if (bpa > 0.) {
                    const double bpa_sibtot = bpa * ddot(k, sms + (size_t)a * k, xas);
                    double *ea = ecounts + ((size_t)br * k + a) * k;
                    for (int b = 0; b < k; b++) {
                        const double pr = bpa_sibtot * sm[(size_t)a * k + b] * ab[b] / z;
                        ea[b] += weight * pr;
                    }
                }
            }
        }
        free(parent);
        free(sibling);
    }
    free(alpha);
    free(leafvec);
    free(beta);
    free(inter);
    return z;
}
