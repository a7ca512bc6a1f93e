/*
 * phylocsf_host.h — C ABI of the host-side half of the drop-in (C++ in phylocsf_b200/csrc/host):
 * parameter loading, tree numbering and rate-matrix diagonalisation, i.e. what the reference's OCaml
 * host does in PhyloCSF.initialize_strategy (src/PhyloCSF.ml:406-467), PhyloCSFModel.make
 * (src/PhyloCSFModel.ml:107-110) and Q.Diag.of_Q / equilibrium (lib/CamlPaml/Q.ml:124-177) before
 * any likelihood is evaluated. The OCaml toolchain is absent from this image, so this layer is C++
 * (DESIGN.md "Host language"); with OCaml present, the reference's own modules would call the
 * compute ABI in phylocsf_b200.h directly.
 */
#ifndef PHYLOCSF_HOST_H
#define PHYLOCSF_HOST_H

#include <stdint.h>

#include "phylocsf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcsf_paramset pcsf_paramset;

/*
 * Load <prefix>.nh, <prefix>_coding.ECM, <prefix>_noncoding.ECM (src/PhyloCSF.ml:423-445).
 * species_csv: optional "--species" list (comma separated) -> Newick.subtree pruning (:425-433).
 * with_ecm = 0 loads only the tree (omega strategy). On failure returns a negative PCSF_ERR_* and
 * writes the reference-style exception text into err (if errlen > 0).
 */
int pcsf_paramset_load(const char *prefix, const char *species_csv, int with_ecm, pcsf_paramset **out, char *err,
                       int errlen);
void pcsf_paramset_free(pcsf_paramset *ps);
int pcsf_paramset_n_leaves(const pcsf_paramset *ps);
const char *pcsf_paramset_leaf_label(const pcsf_paramset *ps, int leaf);
/* children[2*(n_leaves-1)], branch_len[2*n_leaves-2] in the pcsf_tree_set layout */
int pcsf_paramset_tree(const pcsf_paramset *ps, int32_t *children, double *branch_len);
/* which: 0 = coding ECM, 1 = noncoding ECM. Any output may be NULL. Q is the scaled rate matrix. */
int pcsf_paramset_qdiag(const pcsf_paramset *ps, int which, double *Q, double *S, double *Sinv, double *lambda,
                        double *prior);
/* sets the tree and both models (slot 0 = coding, slot 1 = noncoding) on a compute context */
int pcsf_paramset_install(pcsf_ctx *ctx, const pcsf_paramset *ps);

/* Diagonalise an arbitrary reversible 64x64 rate matrix with stationary weights w (Q.Diag.of_Q +
 * equilibrium for callers that assemble their own Q, e.g. the omega model). */
int pcsf_qdiag_reversible(const double *Q, const double *w, double *S, double *Sinv, double *lambda, double *prior,
                          char *err, int errlen);
/* OmegaModel rate matrix at settings v[12] = kappa, omega, sigma, 9 x F3x4 (src/OmegaModel.ml:21-80) */
int pcsf_omega_q(const double *v, double *Q, double *pi, char *err, int errlen);

/*
 * OmegaModel.score (src/OmegaModel.ml:195-219) for a batch of regions, in one call: stages the regions
 * (pcsf_batch_upload layout: region_off[nregions+1], codes[col * n_leaves + leaf] on the host) on `ctx`, whose tree must be
 * set, and runs the omega strategy for all of them together - H0 (omega = sigma = 1) against H1 (omega_H1, sigma_H1),
 * each by kr_map (three rounds of maximize_lpr over rho and kappa with their priors, :160-190) after update_f3x4
 * (:102-134). Model slots 0 .. nregions-1 of the context are overwritten.
 *   out_score[r]     = 10 (lpr_H1 - lpr_H0) / ln 10 (decibans)
 *   out_diag[10 r..] = L(H0), rho_H0, kappa_H0, omega_H0, sigma_H0, L(H1), rho_H1, kappa_H1, omega_H1, sigma_H1 - the
 *                      reference's diagnostics (--debug), unrounded
 *   out_status[r]    = 0, or non-zero when the region raised in the reference's terms (its score is then undefined);
 *                      may be NULL
 * This is what the command line's --strategy=omega runs (csrc/host/omega_strategy.hpp).
 */
int pcsf_omega_score(pcsf_ctx *ctx, int64_t nregions, const int64_t *region_off, const uint8_t *codes, double omega_H1,
                     double sigma_H1, double *out_score, double *out_diag, int32_t *out_status);

/*
 * The tree program pcsf_tree_set derives from T.children for the pruning kernels (csrc/pcsf_program.hpp), for inspection and
 * for the CPU tests that interpret it against the reference's pruning (lib/CamlPaml/PhyloLik.ml:73-93). No GPU needed.
 *   level     0 = the plain post-order program (Sethi-Ullman order), 2 / 3 / 4 = the table programs of the wide kernel
 *   keep      non-zero: with the KEEP / MUL rewrite (the default of the library), 0: every push and pop in place
 *   ops_out   4 ints per op: kind (| table << 8), a, b, c as in csrc/pcsf_program.hpp (OpKind)
 *   tabs_out  5 ints per memoised subtree: leaf a, leaf b, joining leaf (-1: a cherry), node whose upward edge the table
 *             includes, index of the table it extends (-1: a cherry); cherries first, then 3-leaf, then 4-leaf subtrees
 *   info_out  n cherries, n 3-leaf subtrees, n 4-leaf subtrees, parked partials needed (4 ints; may be NULL)
 * Returns the number of ops, or a negative error (bad tree, buffers too small).
 */
int pcsf_host_tree_program(int n_leaves, const int32_t *children, int level, int keep, int32_t *ops_out, int max_ops,
                           int32_t *tabs_out, int max_tabs, int32_t *info_out);

#ifdef __cplusplus
}
#endif
#endif
