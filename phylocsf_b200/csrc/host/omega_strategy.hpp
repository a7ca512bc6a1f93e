// omega_strategy.hpp — the omega strategy over the C ABI, for all regions of a batch at once:
//   OmegaModel.score        <- src/OmegaModel.ml:195-219 (H0: omega = sigma = 1 against H1: omega_H1, sigma_H1)
//   kr_map                  <- src/OmegaModel.ml:160-190 (three rounds of maximize_lpr over rho, then over kappa)
//   update_f3x4             <- src/OmegaModel.ml:102-134 (codon-position nucleotide counts, pseudocount 1)
//   lpr_rho / lpr_kappa     <- src/OmegaModel.ml:147-157 (half-Cauchy and gamma log-priors)
// Every maximize_lpr (src/PhyloCSFModel.ml:84-99 -> lib/CamlPaml/Fit.ml) is a MaximizeLpr state machine per region
// (pcsf_brent.hpp); a round gathers one candidate per live region, builds the candidates' rate matrices (K5,
// pcsf_omega_models_set) and P(t) (K1, pcsf_pt_build_pairs) and scores them (pcsf_lpr_pairs) in one launch sequence.
// Used by the command line (driver.hpp: DeviceScorer::score_omega) and by pcsf_omega_score (phylocsf_host.h).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/phylocsf_b200.h"
#include "../pcsf_brent.hpp"
#include "newick_tree.hpp"

namespace pcsf {
namespace host {

inline double db(double x) { return 10.0 * x / std::log(10.0); }

inline std::string status_exn(int32_t st) {
    if (st & PCSF_ST_NEG_T) return "Invalid_argument(\"CamlPaml.Q.to_Pt\")";
    if (st & (PCSF_ST_NEG_ENTRY | PCSF_ST_ROWSUM)) return "Failure(\"CamlPaml.Q.substitution matrix: expm(t*Q) failed its checks\")";
    if (st & PCSF_ST_DIAG_ASSERT) return "Assert_failure(\"lib/CamlPaml/Q.ml\", 245, 3)";
    if (st & PCSF_ST_BRACKET) return "Gsl.Error.Gsl_exn(Gsl.Error.EINVAL, \"endpoints do not enclose a minimum\")";
    if (st & PCSF_ST_NOT_FINITE) return "Gsl.Error.Gsl_exn(Gsl.Error.EBADFUNC, \"computed function value is infinite or NaN\")";
    return "";
}


struct OmegaStrategy {
    pcsf_ctx* ctx;
    int n_leaves;
    int64_t evaluations = 0;  // likelihood evaluations (region x candidate) issued
    bool use_cache = std::getenv("PCSF_OMEGA_COLD") == nullptr;  // PCSF_OMEGA_COLD=1: every diagonalisation from the identity (A/B runs)
    OmegaStrategy(pcsf_ctx* c, int nl) : ctx(c), n_leaves(nl) {}

    void check(int rc) {
        if (rc != PCSF_OK && rc != PCSF_ERR_NUMERIC) throw failure(std::string("phylocsf_b200: ") + pcsf_last_error(ctx));
    }

    struct OmegaInst {
        double qs[12];
        double rho;
    };
    static double lpr_rho(double x) {  // half_cauchy_lpdf ~mode:1.0 ~scale:0.5, OmegaModel.ml:148-156
        const double pi = std::acos(-1.0), mode = 1.0, scale = 0.5;
        const double numer = 1.0 / (pi * scale * (1.0 + std::pow((x - mode) / scale, 2.0)));
        const double denom = 1.0 - (std::atan((0.0 - mode) / scale) / pi + 0.5);
        return std::log(numer) - std::log(denom);
    }
    static double lpr_kappa(double k) {  // log (gsl_ran_gamma_pdf ~a:7 ~b:0.25 (k - 1 + epsilon_float)), :157
        const double x = k - 1.0 + 2.220446049250313e-16, a = 7.0, b = 0.25;
        double p;
        if (x < 0) p = 0;
        else if (x == 0) p = 0;
        else p = std::exp((a - 1) * std::log(x / b) - x / b - std::lgamma(a)) / b;
        return std::log(p);
    }

    // Assemble and diagonalise Q(qs) for every listed region on the device (K5) as models slot = index.
    // `warm`: region r's slot of the eigenvector cache starts (and keeps) its diagonalisation (kappa searches)
    void omega_install_models(const std::vector<OmegaInst>& inst, const std::vector<int64_t>& which, std::vector<std::string>& exn, bool warm = false) {
        const size_t n = which.size();
        std::vector<double> qs(n * 12);
        for (size_t i = 0; i < n; i++) std::memcpy(&qs[i * 12], inst[which[i]].qs, 12 * sizeof(double));
        std::vector<int32_t> st(n, 0);
        check(pcsf_omega_models_set_cached(ctx, 0, (int)n, qs.data(), warm && use_cache ? which.data() : nullptr, st.data()));
        for (size_t i = 0; i < n; i++) {
            if (st[i] & 128) exn[which[i]] = "Failure(\"CamlPaml.P14n.instantiate_q: Q scale evaluated to a non-positive value\")";
            else if (st[i]) exn[which[i]] = "Failure(\"CamlPaml.Q.equilibrium: smallest-magnitude eigenvalue is unacceptably large; check rate matrix validity or increase tol\")";
        }
    }

    // One coordinate of kr_map (OmegaModel.ml:171-188) for all regions at once: maximize_lpr over rho
    // (kappa_phase = false: rate matrices fixed, tree scale varies) or over kappa (new Q per candidate).
    void omega_maximize(std::vector<OmegaInst>& inst, bool kappa_phase, std::vector<double>& lpr_out, std::vector<std::string>& exn) {
        const int64_t R = (int64_t)inst.size();
        std::vector<MaximizeLpr> st;
        st.reserve(R);
        for (int64_t r = 0; r < R; r++)
            st.emplace_back(kappa_phase ? inst[r].qs[0] : inst[r].rho, kappa_phase ? 1.0 : 0.001, 10.0, 0.01);
        std::vector<int64_t> all(R);
        for (int64_t r = 0; r < R; r++) all[r] = r;
        if (!kappa_phase) omega_install_models(inst, all, exn);  // slot r = region r
        else if (use_cache) check(pcsf_omega_cache_reset(ctx, R));  // every kappa search starts cold: bounds the accumulated rotations
        std::vector<int64_t> live, eval_pair;
        std::vector<int32_t> pair_model, pstat, estat;
        std::vector<double> pair_scale, lpr, xs;
        for (;;) {
            live.clear();
            xs.clear();
            for (int64_t r = 0; r < R; r++)
                if (!st[r].done() && exn[r].empty()) { live.push_back(r); xs.push_back(st[r].candidate()); }
            if (live.empty()) break;
            const int64_t n = (int64_t)live.size();
            pair_model.resize(n);
            pair_scale.resize(n);
            eval_pair.resize(n);
            if (kappa_phase) {
                std::vector<OmegaInst> cand(inst);
                for (int64_t i = 0; i < n; i++) cand[live[i]].qs[0] = xs[i];
                omega_install_models(cand, live, exn, true);  // slot i = i-th live region
                for (int64_t i = 0; i < n; i++) { pair_model[i] = (int32_t)i; pair_scale[i] = inst[live[i]].rho; }
            } else {
                for (int64_t i = 0; i < n; i++) { pair_model[i] = (int32_t)live[i]; pair_scale[i] = xs[i]; }
            }
            pstat.assign(n, 0);
            estat.assign(n, 0);
            lpr.assign(n, 0.0);
            // one P set per candidate: keep the tables of one launch sequence under ~16 GiB
            const int64_t max_sets = std::max<int64_t>(1, (int64_t)((16ull << 30) / ((size_t)(2 * n_leaves - 2) * 65 * 64 * 8)));
            for (int64_t c0 = 0; c0 < n; c0 += max_sets) {
                const int64_t nc = std::min(max_sets, n - c0);
                for (int64_t i = 0; i < nc; i++) eval_pair[c0 + i] = i;
                check(pcsf_pt_build_pairs(ctx, nc, pair_model.data() + c0, pair_scale.data() + c0, pstat.data() + c0));
                check(pcsf_lpr_pairs(ctx, nc, eval_pair.data() + c0, live.data() + c0, lpr.data() + c0, nullptr, estat.data() + c0));
            }
            evaluations += n;
            for (int64_t i = 0; i < n; i++) {
                const int64_t r = live[i];
                if (!exn[r].empty()) continue;  // diagonalisation failed for this candidate
                const double prior = kappa_phase ? lpr_kappa(xs[i]) : lpr_rho(xs[i]);
                st[r].feed(prior + lpr[i], 0.0, estat[i] & ~PCSF_ST_NOT_FINITE);
            }
        }
        for (int64_t r = 0; r < R; r++) {
            if (!exn[r].empty()) continue;
            const int32_t bad = st[r].status & ~PCSF_ST_RANDOM_INIT;
            if (bad) { exn[r] = status_exn(bad); continue; }
            (kappa_phase ? inst[r].qs[0] : inst[r].rho) = st[r].result_x;
            lpr_out[r] = st[r].result_f;
        }
    }

    void omega_kr_map(std::vector<OmegaInst>& inst, std::vector<double>& lpr, std::vector<std::string>& exn) {
        for (int round = 0; round < 3; round++) {  // OmegaModel.ml:189-190
            omega_maximize(inst, false, lpr, exn);
            omega_maximize(inst, true, lpr, exn);
        }
    }

    // OmegaModel.score for regions r = 0..R-1: codes[region_off[r] .. region_off[r+1])[n_leaves] (host). The batch
    // must already be staged on the context (pcsf_batch_upload of the same regions).
    // out_score[r] = 10 (lpr_H1 - lpr_H0) / ln 10; out_diag[10 r ..] = L(H0) rho_H0 kappa_H0 omega_H0 sigma_H0 L(H1) rho_H1
    // kappa_H1 omega_H1 sigma_H1 (the reference's diagnostics, unrounded); exn[r] non-empty = the region raised.
    void score(int64_t R, const int64_t* region_off, const uint8_t* codes, double omega_H1, double sigma_H1, double* out_score,
               double* out_diag, std::vector<std::string>& exn) {
        std::vector<OmegaInst> inst(R);
        exn.assign(R, std::string());
        for (int64_t r = 0; r < R; r++) {  // new_instance ~kappa:2.5 + update_f3x4 (OmegaModel.ml:95-134,197)
            OmegaInst& in = inst[r];
            in.qs[0] = 2.5; in.qs[1] = 1.0; in.qs[2] = 1.0;
            in.rho = 1.0;
            long counts[3][4];
            for (auto& row : counts) for (auto& c : row) c = 1;
            for (int64_t i = region_off[r] * n_leaves; i < region_off[r + 1] * n_leaves; i++) {
                const int c = codes[i];
                if (c < 64) { counts[0][c / 16]++; counts[1][(c / 4) % 4]++; counts[2][c % 4]++; }
            }
            for (int p = 0; p < 3; p++)
                for (int n = 0; n < 3; n++) in.qs[3 + 3 * p + n] = (double)counts[p][n] / (double)counts[p][3];
        }
        std::vector<double> lpr0(R, 0.0), lpr1(R, 0.0);
        omega_kr_map(inst, lpr0, exn);
        std::vector<OmegaInst> inst0(inst);
        for (int64_t r = 0; r < R; r++) { inst[r].qs[1] = omega_H1; inst[r].qs[2] = sigma_H1; }
        omega_kr_map(inst, lpr1, exn);
        for (int64_t r = 0; r < R; r++) {
            if (!exn[r].empty()) continue;
            out_score[r] = 10.0 * (lpr1[r] - lpr0[r]) / std::log(10.0);
            const OmegaInst &a = inst0[r], &b = inst[r];
            const double d[10] = {db(lpr0[r]), a.rho, a.qs[0], a.qs[1], a.qs[2], db(lpr1[r]), b.rho, b.qs[0], b.qs[1], b.qs[2]};
            std::memcpy(out_diag + 10 * r, d, sizeof d);
        }
    }
};

}  // namespace host
}  // namespace pcsf
