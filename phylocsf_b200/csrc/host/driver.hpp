// driver.hpp — the drop-in PhyloCSF driver over the C ABI: option surface, batching of regions from
// many alignments onto the GPU, the three scoring strategies, region selection and the report lines.
//   options / main loop      <- src/PhyloCSF.ml:17-49,469-491
//   process_alignment        <- src/PhyloCSF.ml:280-389 (three-tier error policy, report format)
//   PhyloCSFModel.score      <- src/PhyloCSFModel.ml:113-146 (fixed / mle)
//   OmegaModel.score, kr_map <- src/OmegaModel.ml:102-219
// The reference evaluates one region at a time; here the regions of many alignments are staged as
// one batch (pcsf_batch_upload) and every strategy advances all of them together.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <future>
#include <memory>
#include <functional>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/phylocsf_b200.h"
#include "../pcsf_brent.hpp"
#include "alignment.hpp"
#include "paramset.hpp"

namespace pcsf {
namespace host {

enum Strategy { STRAT_MLE, STRAT_FIXED, STRAT_OMEGA, STRAT_NOP };

struct Options {
    Strategy strategy = STRAT_MLE;
    bool filenames = false, remove_ref_gaps = false, allow_ref_gaps = false;
    std::string species;  // csv, empty = all
    int frames = 1;
    OrfMode orf = AsIs;
    int min_codons = 25;
    bool all_scores = false;
    int procs = 1;
    bool bls = false, anc_comp = false, dna = false, aa = false, debug = false;
    double omega_H1 = 0.2, sigma_H1 = 0.01;
    // not in the reference: how many codon columns to stage per GPU batch
    int64_t batch_cols = 4000000;
    int device = 0;
};

struct ScoreRecord {
    double score = 0.0, anc_comp = 0.0;
    std::vector<std::pair<std::string, std::string>> diag;
    std::string exn;  // non-empty: this region raised (Printexc text)
};

inline double db(double x) { return 10.0 * x / std::log(10.0); }
inline std::string sf2(double x) {
    char b[64];
    snprintf(b, sizeof b, "%.2f", x);
    return b;
}

struct AlnJob {
    std::string name;
    std::vector<std::string> aln, rc_aln;
    std::map<std::string, int> which_row;
    std::vector<Region> regions;
    int64_t first_region = 0;  // index of this alignment's first staged region in the GPU batch
    int64_t first_rec = 0;     // index of its first candidate region in the batch's score records
    // frame mode (fixed strategy + ORF search): the batch holds whole reading frames, and candidate
    // region k is columns [reg_col0[k], reg_col0[k] + reg_ncols[k]) of staged frame reg_frame[k]
    std::vector<int> reg_frame, reg_col0, reg_ncols;
    std::string failure;  // per-alignment failure text (no ORFs)
};

// The regions of many alignments, staged and scored together on one GPU.
struct Batch {
    std::vector<AlnJob> jobs;
    std::vector<uint8_t> codes;          // [column][leaf]
    std::vector<int64_t> region_off{0};  // staged regions (whole frames in frame mode)
    int64_t n_rec = 0;                   // candidate regions (= staged regions except in frame mode)
};

inline std::string status_exn(int32_t st) {
    if (st & PCSF_ST_NEG_T) return "Invalid_argument(\"CamlPaml.Q.to_Pt\")";
    if (st & (PCSF_ST_NEG_ENTRY | PCSF_ST_ROWSUM)) return "Failure(\"CamlPaml.Q.substitution matrix: expm(t*Q) failed its checks\")";
    if (st & PCSF_ST_DIAG_ASSERT) return "Assert_failure(\"lib/CamlPaml/Q.ml\", 245, 3)";
    if (st & PCSF_ST_BRACKET) return "Gsl.Error.Gsl_exn(Gsl.Error.EINVAL, \"endpoints do not enclose a minimum\")";
    if (st & PCSF_ST_NOT_FINITE) return "Gsl.Error.Gsl_exn(Gsl.Error.EBADFUNC, \"computed function value is infinite or NaN\")";
    return "";
}

// One GPU: a compute context with the tree and models installed, and the three strategies over a Batch.
class DeviceScorer {
  public:
    DeviceScorer(const Options& o, const ParamSet& pset, int device) : opt(o), ps(pset) {
        n_leaves = ps.tree.n_leaves;
        if (o.strategy == STRAT_NOP) return;
        if (pcsf_create(device, &ctx) != PCSF_OK)
            throw failure("phylocsf_b200: no usable CUDA device " + std::to_string(device) + " (there is no CPU fallback)");
        const auto ch = ps.tree.children_array();
        std::vector<double> bl(ps.tree.branches.begin(), ps.tree.branches.begin() + ps.tree.root());
        check(pcsf_tree_set(ctx, n_leaves, ch.data(), bl.data()));
        if (ps.have_ecm)
            for (int w = 0; w < 2; w++)
                check(pcsf_model_set(ctx, w, ps.qd[w].S.data(), ps.qd[w].Sinv.data(), ps.qd[w].lam.data(), ps.qd[w].pi_eq.data()));
        if (o.strategy == STRAT_FIXED) {
            const double one = 1.0;
            check(pcsf_pt_build(ctx, 0, 1, &one, nullptr));
            check(pcsf_pt_build(ctx, 1, 1, &one, nullptr));
        }
    }
    ~DeviceScorer() {
        if (ctx) pcsf_destroy(ctx);
    }
    DeviceScorer(const DeviceScorer&) = delete;

    bool frame_mode() const { return opt.strategy == STRAT_FIXED && opt.orf != AsIs; }

    // Score every candidate region of the batch and render the report lines of its alignments.
    std::string run(const Batch& b) {
        cur = &b;
        const int64_t R = (int64_t)b.region_off.size() - 1;
        std::vector<ScoreRecord> rec(b.n_rec);
        if (R > 0) {
            if (opt.strategy != STRAT_NOP) check(pcsf_batch_upload(ctx, R, b.region_off.data(), b.codes.data()));
            switch (opt.strategy) {
                case STRAT_FIXED: if (frame_mode()) score_fixed_frames(rec); else score_fixed(rec); break;
                case STRAT_MLE: score_mle(rec); break;
                case STRAT_OMEGA: score_omega(rec); break;
                case STRAT_NOP: break;
            }
        }
        std::ostringstream out;
        for (const AlnJob& j : b.jobs) report(j, rec, out);
        cur = nullptr;
        return out.str();
    }

    int64_t evaluations = 0;  // likelihood evaluations (region x model) issued

  private:
    Options opt;
    const ParamSet& ps;
    pcsf_ctx* ctx = nullptr;
    int n_leaves = 0;
    const Batch* cur = nullptr;

    void check(int rc) {
        if (rc != PCSF_OK && rc != PCSF_ERR_NUMERIC) throw failure(std::string("phylocsf_b200: ") + pcsf_last_error(ctx));
    }

    // ---- PhyloCSFModel.llr_FixedLik (src/PhyloCSFModel.ml:122-128) ----
    void score_fixed(std::vector<ScoreRecord>& rec) {
        const int64_t R = (int64_t)rec.size();
        std::vector<double> lpr(2 * R), elpr(2 * R);
        std::vector<int32_t> st(2 * R);
        const int32_t mids[2] = {0, 1};
        check(pcsf_lpr_all(ctx, 2, mids, nullptr, lpr.data(), elpr.data(), st.data()));
        evaluations += 2 * R;
        for (int64_t r = 0; r < R; r++) {
            const int32_t bad = (st[r] | st[R + r]) & ~PCSF_ST_NOT_FINITE;  // log 0 = -inf is not an exception here
            if (bad) { rec[r].exn = status_exn(bad); continue; }
            rec[r].score = db(lpr[r] - lpr[R + r]);
            rec[r].anc_comp = db(elpr[r] - elpr[R + r]);
            rec[r].diag = {{"rho", sf2(1.0)}, {"L(C)", sf2(db(lpr[r]))}, {"L(NC)", sf2(db(lpr[R + r]))}};
        }
    }

    // llr_FixedLik for ORF candidates from per-column terms of whole frames. The column sums run in
    // column order, as the reference's `lpr := !lpr +. log ...` does (src/PhyloCSFModel.ml:76-81).
    void score_fixed_frames(std::vector<ScoreRecord>& rec) {
        const int64_t R = (int64_t)cur->region_off.size() - 1, C = cur->region_off.back();
        std::vector<double> lpr(2 * R), elpr(2 * R), clz[2], can[2];
        std::vector<int32_t> st(2 * R);
        const int32_t mids[2] = {0, 1};
        check(pcsf_lpr_all(ctx, 2, mids, nullptr, lpr.data(), elpr.data(), st.data()));
        evaluations += 2 * R;
        for (int m = 0; m < 2; m++) {
            clz[m].resize(C);
            can[m].resize(C);
            check(pcsf_column_terms(ctx, m, clz[m].data(), can[m].data()));
        }
        for (const AlnJob& j : cur->jobs)
            for (size_t k = 0; k < j.regions.size(); k++) {
                ScoreRecord& rc = rec[j.first_rec + k];
                const int64_t b = j.first_region + j.reg_frame[k];
                const int32_t bad = (st[b] | st[R + b]) & ~PCSF_ST_NOT_FINITE;
                if (bad) { rc.exn = status_exn(bad); continue; }
                const int64_t c0 = cur->region_off[b] + j.reg_col0[k];
                double l[2] = {0.0, 0.0}, e[2] = {0.0, 0.0};
                for (int m = 0; m < 2; m++)
                    for (int c = 0; c < j.reg_ncols[k]; c++) {
                        l[m] += clz[m][c0 + c];
                        e[m] += can[m][c0 + c];
                    }
                rc.score = db(l[0] - l[1]);
                rc.anc_comp = db(e[0] - e[1]);
                rc.diag = {{"rho", sf2(1.0)}, {"L(C)", sf2(db(l[0]))}, {"L(NC)", sf2(db(l[1]))}};
            }
    }

    // ---- PhyloCSFModel.llr_MaxLik ~init:1. (src/PhyloCSFModel.ml:130-136) ----
    void score_mle(std::vector<ScoreRecord>& rec) {
        const int64_t R = (int64_t)rec.size();
        std::vector<double> rho2(2 * R), lpr2(2 * R), elpr2(2 * R);
        std::vector<int32_t> st2(2 * R), ne2(2 * R);
        const int32_t mids[2] = {0, 1};
        check(pcsf_maximize_lpr_multi(ctx, 2, mids, 1.0, 1e-2, 10.0, 0.01, rho2.data(), lpr2.data(), elpr2.data(), st2.data(), ne2.data()));
        for (int64_t i = 0; i < 2 * R; i++) evaluations += ne2[i];
        const double *rho[2] = {rho2.data(), rho2.data() + R}, *lpr[2] = {lpr2.data(), lpr2.data() + R},
                     *elpr[2] = {elpr2.data(), elpr2.data() + R};
        const int32_t* st[2] = {st2.data(), st2.data() + R};
        for (int64_t r = 0; r < R; r++) {
            const int32_t b0 = st[0][r] & ~PCSF_ST_RANDOM_INIT, b1 = st[1][r] & ~PCSF_ST_RANDOM_INIT;
            if (b0 || b1) { rec[r].exn = status_exn(b0 ? b0 : b1); continue; }  // coding model is maximised first
            rec[r].score = db(lpr[0][r] - lpr[1][r]);
            rec[r].anc_comp = db(elpr[0][r] - elpr[1][r]);
            rec[r].diag = {{"rho_0", sf2(1.0)}, {"rho_C", sf2(rho[0][r])}, {"rho_N", sf2(rho[1][r])},
                           {"L(C)", sf2(db(lpr[0][r]))}, {"L(NC)", sf2(db(lpr[1][r]))}};
        }
    }

    // ---- OmegaModel (src/OmegaModel.ml) ----
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
    void omega_install_models(const std::vector<OmegaInst>& inst, const std::vector<int64_t>& which, std::vector<std::string>& exn) {
        const size_t n = which.size();
        std::vector<double> qs(n * 12);
        for (size_t i = 0; i < n; i++) std::memcpy(&qs[i * 12], inst[which[i]].qs, 12 * sizeof(double));
        std::vector<int32_t> st(n, 0);
        check(pcsf_omega_models_set(ctx, 0, (int)n, qs.data(), st.data()));
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
                omega_install_models(cand, live, exn);  // slot i = i-th live region
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

    void score_omega(std::vector<ScoreRecord>& rec) {
        const int64_t R = (int64_t)rec.size();
        std::vector<OmegaInst> inst(R);
        std::vector<std::string> exn(R);
        for (int64_t r = 0; r < R; r++) {  // new_instance ~kappa:2.5 + update_f3x4 (OmegaModel.ml:95-134,197)
            OmegaInst& in = inst[r];
            in.qs[0] = 2.5; in.qs[1] = 1.0; in.qs[2] = 1.0;
            in.rho = 1.0;
            long counts[3][4];
            for (auto& row : counts) for (auto& c : row) c = 1;
            for (int64_t i = cur->region_off[r] * n_leaves; i < cur->region_off[r + 1] * n_leaves; i++) {
                const int c = cur->codes[i];
                if (c < 64) { counts[0][c / 16]++; counts[1][(c / 4) % 4]++; counts[2][c % 4]++; }
            }
            for (int p = 0; p < 3; p++)
                for (int n = 0; n < 3; n++) in.qs[3 + 3 * p + n] = (double)counts[p][n] / (double)counts[p][3];
        }
        std::vector<double> lpr0(R, 0.0), lpr1(R, 0.0);
        omega_kr_map(inst, lpr0, exn);
        std::vector<OmegaInst> inst0(inst);
        for (int64_t r = 0; r < R; r++) { inst[r].qs[1] = opt.omega_H1; inst[r].qs[2] = opt.sigma_H1; }
        omega_kr_map(inst, lpr1, exn);
        for (int64_t r = 0; r < R; r++) {
            if (!exn[r].empty()) { rec[r].exn = exn[r]; continue; }
            rec[r].score = 10.0 * (lpr1[r] - lpr0[r]) / std::log(10.0);
            rec[r].anc_comp = std::nan("");
            const OmegaInst &a = inst0[r], &b = inst[r];
            rec[r].diag = {{"L(H0)", sf2(db(lpr0[r]))}, {"rho_H0", sf2(a.rho)}, {"kappa_H0", sf2(a.qs[0])}, {"omega_H0", sf2(a.qs[1])},
                           {"sigma_H0", sf2(a.qs[2])}, {"L(H1)", sf2(db(lpr1[r]))}, {"rho_H1", sf2(b.rho)}, {"kappa_H1", sf2(b.qs[0])},
                           {"omega_H1", sf2(b.qs[1])}, {"sigma_H1", sf2(b.qs[2])}};
        }
    }

    // ---- report (src/PhyloCSF.ml:340-380) ----
    static bool ocaml_ge(const ScoreRecord& a, const Region& ra, const ScoreRecord& b, const Region& rb) {
        // structural compare of (score record, rc, lo, hi); a NaN makes the comparison false
        const double ka[2] = {a.score, a.anc_comp}, kb[2] = {b.score, b.anc_comp};
        for (int i = 0; i < 2; i++) {
            if (std::isnan(ka[i]) || std::isnan(kb[i])) return false;
            if (ka[i] > kb[i]) return true;
            if (ka[i] < kb[i]) return false;
        }
        if (a.diag != b.diag) return a.diag > b.diag;
        if (ra.rc != rb.rc) return ra.rc;
        if (ra.lo != rb.lo) return ra.lo > rb.lo;
        return ra.hi >= rb.hi;
    }

    void report(const AlnJob& j, const std::vector<ScoreRecord>& rec, std::ostream& out) {
        char buf[64];
        auto fmt4 = [&](double x) { snprintf(buf, sizeof buf, "%.4f", x); return std::string(buf); };
        std::string fail = j.failure;
        std::vector<int> ok;
        if (fail.empty()) {
            for (size_t k = 0; k < j.regions.size(); k++) {
                const ScoreRecord& s = rec[j.first_rec + k];
                const Region& rg = j.regions[k];
                if (s.exn.empty()) { ok.push_back((int)k); continue; }
                out << j.name << "\texception\t" << rg.lo << "\t" << rg.hi;
                if (opt.frames == 6) out << (rg.rc ? "\t-" : "\t+");
                out << "\t" << s.exn << "\n";
            }
            if (ok.empty()) fail = "Failure(\"no regions successfully evaluated\")";
        }
        if (!fail.empty()) {
            out << j.name << "\tfailure\t" << fail << "\n";
            out.flush();
            return;
        }
        auto line = [&](const char* ty, int k) {
            const ScoreRecord& s = rec[j.first_rec + k];
            const Region& rg = j.regions[k];
            const std::vector<std::string>& rows = rg.rc ? j.rc_aln : j.aln;
            out << j.name << "\t" << ty << "\t" << fmt4(s.score);
            if (opt.frames != 1 || opt.orf != AsIs) out << "\t" << rg.lo << "\t" << rg.hi;
            if (opt.frames == 6) out << "\t" << (rg.rc ? '-' : '+');
            if (opt.bls) out << "\t" << fmt4(bls_score(ps.nt, rows, j.which_row, rg.lo, rg.hi));
            if (opt.anc_comp) out << "\t" << fmt4(s.anc_comp);
            const std::string refdna = rows[0].substr(rg.lo, rg.hi - rg.lo + 1);
            if (opt.dna) out << "\t" << refdna;
            if (opt.aa) out << "\t" << translate(refdna);
            if (opt.debug) {
                out << "\t#";
                for (auto& kv : s.diag) out << " " << kv.first << "=" << kv.second;
            }
            out << "\n";
        };
        if (opt.all_scores)
            for (int k : ok) line("orf_score(decibans)", k);
        int best = ok[0];
        for (size_t i = 1; i < ok.size(); i++)
            if (!ocaml_ge(rec[j.first_rec + best], j.regions[best], rec[j.first_rec + ok[i]], j.regions[ok[i]])) best = ok[i];
        line((opt.orf != AsIs || opt.frames != 1) ? "max_score(decibans)" : "score(decibans)", best);
        out.flush();
    }

};

// The driver proper: prepares alignments (host threads), groups them into batches, hands each batch to a
// GPU (round-robin over the configured devices, scoring asynchronously while the next batch is being
// prepared) and prints the report lines in input order.
class Driver {
  public:
    Driver(const Options& o, const std::string& paramset_prefix, std::vector<int> devices = {}) : opt(o) {
        ps = load_paramset(paramset_prefix, o.species, o.strategy == STRAT_MLE || o.strategy == STRAT_FIXED);
        n_leaves = ps.tree.n_leaves;
        leaf_labels.assign(ps.tree.labels.begin(), ps.tree.labels.begin() + n_leaves);
        leaf_set.insert(leaf_labels.begin(), leaf_labels.end());
        if (devices.empty()) devices.push_back(o.device);
        if (o.strategy == STRAT_NOP) devices.resize(1);
        for (int d : devices) dev.emplace_back(new DeviceScorer(o, ps, d));
    }

    bool frame_mode() const { return opt.strategy == STRAT_FIXED && opt.orf != AsIs; }

    // Everything of process_alignment that precedes scoring (src/PhyloCSF.ml:283-308,320), for one
    // alignment: parse, sanity checks, leaf order, candidate regions, leaf codes. Thread-safe (reads
    // only the immutable parts of the driver), so many alignments can be prepared in parallel.
    struct Prepared {
        AlnJob job;
        std::vector<uint8_t> codes;
        std::vector<int> region_cols;
        std::string abort;  // non-empty: the alignment aborted with this Printexc text
    };
    Prepared prepare(const std::string& name, const std::vector<std::string>& lines) const {
        Prepared p;
        AlnJob& job = p.job;
        job.name = name;
        try {
            Alignment a = input_mfa(lines);
            if (opt.remove_ref_gaps) remove_ref_gaps(a.seqs);
            for (auto& s : a.seqs)
                for (auto& c : s) c = c == 'u' ? 't' : (c == 'U' ? 'T' : c);
            if (!opt.allow_ref_gaps && a.seqs[0].find('-') != std::string::npos)
                throw failure("the reference sequence (first alignment row) must be ungapped");
            job.aln = a.seqs;
            for (auto& s : a.seqs) job.rc_aln.push_back(revcomp(s));
            std::set<std::string> wtf;
            for (auto& sp : a.species)
                if (!leaf_set.count(sp)) wtf.insert(sp);
            if (!wtf.empty()) {
                std::string m = "parameters not available for species:";
                for (auto& s : wtf) m += " " + s;
                throw failure(m);
            }
            for (size_t i = 0; i < a.species.size(); i++) job.which_row[a.species[i]] = (int)i;
            job.regions = candidate_regions(job.aln[0], opt.orf, opt.frames, opt.min_codons);
            if (job.regions.empty()) job.failure = "Failure(\"no sufficiently long ORFs found\")";
        } catch (const HostError& e) {
            p.abort = e.what();
            return p;
        }
        std::vector<int> leaf_ord(n_leaves, -1);
        for (int l = 0; l < n_leaves; l++) {
            auto it = job.which_row.find(leaf_labels[l]);
            if (it != job.which_row.end()) leaf_ord[l] = it->second;
        }
        if (!frame_mode()) {
            for (const Region& r : job.regions)
                p.region_cols.push_back(pleaves(n_leaves, leaf_ord, r.rc ? job.rc_aln : job.aln, r.lo, r.hi, p.codes));
            return p;
        }
        // Frame mode: under the fixed strategy every column's log-likelihood is independent of the region
        // it is scored in, so nested / overlapping ORFs (ATGStop emits one ORF per upstream ATG of a stop)
        // share columns. Stage each reading frame that holds a candidate once; ORF scores become
        // segment sums of the per-column terms.
        int frame_slot[6] = {-1, -1, -1, -1, -1, -1};
        const int hi_all = (int)job.aln[0].size() - 1;
        for (const Region& r : job.regions) {
            const int ofs = r.lo % 3, f = (r.rc ? 3 : 0) + ofs;
            if (frame_slot[f] < 0) {
                frame_slot[f] = (int)p.region_cols.size();
                p.region_cols.push_back(pleaves(n_leaves, leaf_ord, r.rc ? job.rc_aln : job.aln, ofs, hi_all, p.codes));
            }
            job.reg_frame.push_back(frame_slot[f]);
            job.reg_col0.push_back((r.lo - ofs) / 3);
            job.reg_ncols.push_back((r.hi - r.lo + 1) / 3);
        }
        return p;
    }

    // Appends a prepared alignment to the current batch (in input order). Returns false when the run
    // must stop (the alignment aborted: src/PhyloCSF.ml:381-388 exits -1).
    bool append(Prepared&& p, std::ostream& out) {
        if (!p.abort.empty()) {
            finish(out);
            out << p.job.name << "\tabort\t" << p.abort << "\n";
            out.flush();
            return false;
        }
        p.job.first_region = (int64_t)batch.region_off.size() - 1;
        p.job.first_rec = batch.n_rec;
        batch.n_rec += (int64_t)p.job.regions.size();
        for (int nc : p.region_cols) batch.region_off.push_back(batch.region_off.back() + nc);
        batch.codes.insert(batch.codes.end(), p.codes.begin(), p.codes.end());
        batch.jobs.push_back(std::move(p.job));
        if (batch.region_off.back() >= opt.batch_cols) flush(out);
        return true;
    }

    bool add_alignment(const std::string& name, const std::vector<std::string>& lines, std::ostream& out) {
        return append(prepare(name, lines), out);
    }

    // Hand the current batch to the next GPU; at most one batch per device is in flight.
    void flush(std::ostream& out) {
        if (batch.jobs.empty()) return;
        if (inflight.size() >= dev.size()) drain_one(out);
        auto b = std::make_shared<Batch>(std::move(batch));
        batch = Batch();
        DeviceScorer* d = dev[next_dev++ % dev.size()].get();  // the oldest in-flight batch ran on this device: it is free
        inflight.push_back(std::async(std::launch::async, [d, b]() { return d->run(*b); }));
    }
    // Wait for everything in flight and print it (end of input, or before an abort line).
    void finish(std::ostream& out) {
        flush(out);
        while (!inflight.empty()) drain_one(out);
    }
    int64_t evaluations() const {
        int64_t n = 0;
        for (auto& d : dev) n += d->evaluations;
        return n;
    }

  private:
    Options opt;
    ParamSet ps;
    int n_leaves = 0;
    std::vector<std::string> leaf_labels;
    std::set<std::string> leaf_set;
    std::vector<std::unique_ptr<DeviceScorer>> dev;
    Batch batch;
    std::deque<std::future<std::string>> inflight;
    size_t next_dev = 0;

    void drain_one(std::ostream& out) {
        out << inflight.front().get();
        out.flush();
        inflight.pop_front();
    }
};

}  // namespace host
}  // namespace pcsf
