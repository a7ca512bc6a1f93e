// pcsf_host_api.cu — C ABI of include/phylocsf_host.h over the C++ host layer (csrc/host).
#include "../../include/phylocsf_host.h"

#include <cstring>
#include <set>
#include <sstream>

#include "pcsf_program.hpp"
#include "host/omega_strategy.hpp"
#include "host/paramset.hpp"

using namespace pcsf::host;

struct pcsf_paramset : ParamSet {};

namespace {
void put_err(char* err, int errlen, const std::string& m) {
    if (err && errlen > 0) {
        std::strncpy(err, m.c_str(), (size_t)errlen - 1);
        err[errlen - 1] = 0;
    }
}
}  // namespace

extern "C" {

int pcsf_paramset_load(const char* prefix, const char* species_csv, int with_ecm, pcsf_paramset** out, char* err,
                       int errlen) {
    if (!prefix || !out) return PCSF_ERR_INVALID_ARG;
    *out = nullptr;
    try {
        auto ps = new pcsf_paramset();
        std::unique_ptr<pcsf_paramset> guard(ps);
        static_cast<ParamSet&>(*ps) = load_paramset(prefix, species_csv ? species_csv : "", with_ecm != 0);
        *out = guard.release();
        return PCSF_OK;
    } catch (const std::exception& e) {
        put_err(err, errlen, e.what());
        return PCSF_ERR_INVALID_ARG;
    }
}

void pcsf_paramset_free(pcsf_paramset* ps) { delete ps; }
int pcsf_paramset_n_leaves(const pcsf_paramset* ps) { return ps ? ps->tree.n_leaves : 0; }
const char* pcsf_paramset_leaf_label(const pcsf_paramset* ps, int leaf) {
    if (!ps || leaf < 0 || leaf >= ps->tree.n_leaves) return "";
    return ps->tree.labels[leaf].c_str();
}

int pcsf_paramset_tree(const pcsf_paramset* ps, int32_t* children, double* branch_len) {
    if (!ps) return PCSF_ERR_INVALID_ARG;
    const auto ch = ps->tree.children_array();
    if (children) std::memcpy(children, ch.data(), ch.size() * sizeof(int32_t));
    if (branch_len)
        for (int i = 0; i < ps->tree.root(); i++) branch_len[i] = ps->tree.branches[i];
    return PCSF_OK;
}

int pcsf_paramset_qdiag(const pcsf_paramset* ps, int which, double* Q, double* S, double* Sinv, double* lambda,
                        double* prior) {
    if (!ps || !ps->have_ecm || which < 0 || which > 1) return PCSF_ERR_INVALID_ARG;
    const QDiag& d = ps->qd[which];
    if (Q) std::memcpy(Q, d.q.data(), 4096 * 8);
    if (S) std::memcpy(S, d.S.data(), 4096 * 8);
    if (Sinv) std::memcpy(Sinv, d.Sinv.data(), 4096 * 8);
    if (lambda) std::memcpy(lambda, d.lam.data(), 64 * 8);
    if (prior) std::memcpy(prior, d.pi_eq.data(), 64 * 8);
    return PCSF_OK;
}

int pcsf_paramset_install(pcsf_ctx* ctx, const pcsf_paramset* ps) {
    if (!ctx || !ps) return PCSF_ERR_INVALID_ARG;
    const auto ch = ps->tree.children_array();
    std::vector<double> bl(ps->tree.branches.begin(), ps->tree.branches.begin() + ps->tree.root());
    int rc = pcsf_tree_set(ctx, ps->tree.n_leaves, ch.data(), bl.data());
    if (rc != PCSF_OK) return rc;
    if (ps->have_ecm)
        for (int w = 0; w < 2; w++) {
            const QDiag& d = ps->qd[w];
            rc = pcsf_model_set(ctx, w, d.S.data(), d.Sinv.data(), d.lam.data(), d.pi_eq.data());
            if (rc != PCSF_OK) return rc;
        }
    return PCSF_OK;
}

int pcsf_qdiag_reversible(const double* Q, const double* w, double* S, double* Sinv, double* lambda, double* prior,
                          char* err, int errlen) {
    if (!Q || !w) return PCSF_ERR_INVALID_ARG;
    try {
        QDiag d = QDiag::of_reversible_Q(std::vector<double>(Q, Q + 4096), std::vector<double>(w, w + 64));
        if (S) std::memcpy(S, d.S.data(), 4096 * 8);
        if (Sinv) std::memcpy(Sinv, d.Sinv.data(), 4096 * 8);
        if (lambda) std::memcpy(lambda, d.lam.data(), 64 * 8);
        if (prior) std::memcpy(prior, d.pi_eq.data(), 64 * 8);
        return PCSF_OK;
    } catch (const std::exception& e) {
        put_err(err, errlen, e.what());
        return PCSF_ERR_NUMERIC;
    }
}

int pcsf_omega_q(const double* v, double* Q, double* pi, char* err, int errlen) {
    if (!v || !Q) return PCSF_ERR_INVALID_ARG;
    try {
        std::vector<double> p;
        std::vector<double> q = omega_q(v, &p);
        std::memcpy(Q, q.data(), 4096 * 8);
        if (pi) std::memcpy(pi, p.data(), 64 * 8);
        return PCSF_OK;
    } catch (const std::exception& e) {
        put_err(err, errlen, e.what());
        return PCSF_ERR_NUMERIC;
    }
}

int pcsf_omega_score(pcsf_ctx* ctx, int64_t nregions, const int64_t* region_off, const uint8_t* codes, double omega_H1,
                     double sigma_H1, double* out_score, double* out_diag, int32_t* out_status) {
    if (!ctx || nregions < 0 || !region_off || !out_score || !out_diag) return PCSF_ERR_INVALID_ARG;
    int rc = pcsf_batch_upload(ctx, nregions, region_off, codes);
    if (rc != PCSF_OK) return rc;
    try {
        // the tree's leaf count, from the staged batch: codes has ncols * n_leaves bytes; the context knows it
        OmegaStrategy om(ctx, pcsf_tree_n_leaves(ctx));
        std::vector<std::string> exn;
        om.score(nregions, region_off, codes, omega_H1, sigma_H1, out_score, out_diag, exn);
        bool bad = false;
        for (int64_t r = 0; r < nregions; r++) {
            if (out_status) out_status[r] = exn[r].empty() ? 0 : 1;
            bad |= !exn[r].empty();
        }
        return bad ? PCSF_ERR_NUMERIC : PCSF_OK;
    } catch (const std::exception&) {
        return PCSF_ERR_CUDA;  // pcsf_last_error(ctx) holds the message of the failing call
    }
}

int pcsf_host_tree_program(int n_leaves, const int32_t* children, int level, int keep, int32_t* ops_out, int max_ops,
                           int32_t* tabs_out, int max_tabs, int32_t* info_out) {
    if (n_leaves < 2 || !children || (level != 0 && level != 2 && level != 3 && level != 4)) return PCSF_ERR_INVALID_ARG;
    const int n = 2 * n_leaves - 1;
    std::vector<int> seen(n, 0);
    for (int i = n_leaves; i < n; i++)
        for (int k = 0; k < 2; k++) {
            const int c = children[2 * (i - n_leaves) + k];
            if (c < 0 || c >= i || seen[c]++) return PCSF_ERR_INVALID_ARG;  // the checks of pcsf_tree_set
        }
    const pcsf::TreePrograms tp =
        pcsf::build_tree_programs(n_leaves, std::vector<int32_t>(children, children + 2 * (n_leaves - 1)), keep != 0);
    const std::vector<pcsf::Op>& ops = level == 4 ? tp.ops_t4 : level == 3 ? tp.ops_t3 : level == 2 ? tp.ops_t : tp.ops;
    if (info_out) {
        info_out[0] = tp.n_tab2;
        info_out[1] = tp.n_tab3;
        info_out[2] = tp.n_tab4;
        info_out[3] = tp.max_levels;
    }
    if ((int)ops.size() > max_ops || (int)tp.subtabs.size() > max_tabs || (!ops_out && max_ops > 0) || (!tabs_out && max_tabs > 0))
        return PCSF_ERR_INVALID_ARG;
    for (size_t i = 0; i < ops.size(); i++) {
        ops_out[4 * i] = ops[i].kind;
        ops_out[4 * i + 1] = ops[i].a;
        ops_out[4 * i + 2] = ops[i].b;
        ops_out[4 * i + 3] = ops[i].c;
    }
    for (size_t k = 0; k < tp.subtabs.size(); k++) {
        tabs_out[5 * k] = tp.subtabs[k].la;
        tabs_out[5 * k + 1] = tp.subtabs[k].lb;
        tabs_out[5 * k + 2] = tp.subtabs[k].lnew;
        tabs_out[5 * k + 3] = tp.subtabs[k].edge;
        tabs_out[5 * k + 4] = tp.subtabs[k].src;
    }
    return (int)ops.size();
}

}  // extern "C"
