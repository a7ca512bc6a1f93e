// driver.hpp — the drop-in PhyloCSF driver over the C ABI: option surface, batching of regions from
// many alignments onto the GPU, the three scoring strategies, region selection and the report lines.
//   options / main loop      <- src/PhyloCSF.ml:17-49,469-491
//   process_alignment        <- src/PhyloCSF.ml:280-389 (three-tier error policy, report format)
//   PhyloCSFModel.score      <- src/PhyloCSFModel.ml:113-146 (fixed / mle)
//   OmegaModel.score, kr_map <- src/OmegaModel.ml:102-219
// The reference evaluates one region at a time; here the regions of many alignments are staged as
// one batch (pcsf_batch_upload) and every strategy advances all of them together.
#pragma once
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <future>
#include <memory>
#include <mutex>
#include <functional>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../../include/phylocsf_b200.h"
#include "../pcsf_brent.hpp"
#include "alignment.hpp"
#include "omega_strategy.hpp"
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

inline std::string sf2(double x) {
    char b[64];
    snprintf(b, sizeof b, "%.2f", x);
    return b;
}

struct ScoreRecord {
    double score = 0.0, anc_comp = 0.0;
    // the strategy's diagnostics (the reference's (string*string) list), kept as numbers and rendered only
    // when printed (--debug) or when two records tie on both scores (the structural compare reaches them)
    enum DiagKind : uint8_t { D_NONE, D_FIXED, D_MLE, D_OMEGA } dk = D_NONE;
    double dv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::string exn;  // non-empty: this region raised (Printexc text)
    bool random_init = false;  // Fit.find_init took its random branch (PCSF_ST_RANDOM_INIT): the restated OCaml Random stream decided the start

    std::vector<std::pair<std::string, std::string>> diag() const {
        static const char* const kFixed[] = {"rho", "L(C)", "L(NC)"};
        static const char* const kMle[] = {"rho_0", "rho_C", "rho_N", "L(C)", "L(NC)"};
        static const char* const kOmega[] = {"L(H0)", "rho_H0", "kappa_H0", "omega_H0", "sigma_H0", "L(H1)", "rho_H1", "kappa_H1", "omega_H1", "sigma_H1"};
        const char* const* keys = dk == D_FIXED ? kFixed : dk == D_MLE ? kMle : kOmega;
        const int n = dk == D_FIXED ? 3 : dk == D_MLE ? 5 : dk == D_OMEGA ? 10 : 0;
        std::vector<std::pair<std::string, std::string>> d;
        for (int i = 0; i < n; i++) d.emplace_back(keys[i], sf2(dv[i]));
        return d;
    }
    void set_diag(DiagKind k, std::initializer_list<double> v) {
        dk = k;
        int i = 0;
        for (double x : v) dv[i++] = x;
    }
};

// Recycles the readers' nucleotide buffers: a buffer goes back to the pool when the last batch using it is
// done, so after the first few batches no buffer is allocated (or page-faulted in) again.
class BufferPool {
  public:
    std::shared_ptr<std::vector<uint8_t>> get() {
        std::vector<uint8_t>* v = nullptr;
        {
            std::lock_guard<std::mutex> lk(st->mu);
            if (!st->free.empty()) {
                v = st->free.back().release();
                st->free.pop_back();
            }
        }
        if (!v) v = new std::vector<uint8_t>();
        v->clear();
        std::shared_ptr<State> keep = st;
        return std::shared_ptr<std::vector<uint8_t>>(v, [keep](std::vector<uint8_t>* q) {
            std::lock_guard<std::mutex> lk(keep->mu);
            keep->free.emplace_back(q);
        });
    }

  private:
    struct State {
        std::mutex mu;
        std::vector<std::unique_ptr<std::vector<uint8_t>>> free;
    };
    std::shared_ptr<State> st = std::make_shared<State>();
};

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
    std::vector<double> bls;  // --bls of every candidate region, when the reader already computed it
};

// The regions of many alignments, staged and scored together on one GPU.
struct Batch {
    std::vector<AlnJob> jobs;
    std::vector<uint8_t> codes;          // [column][leaf]
    std::vector<int64_t> region_off{0};  // staged regions (whole frames in frame mode)
    int64_t n_rec = 0;                   // candidate regions (= staged regions except in frame mode)
    // nucleotide form (AsIs regions): the batch carries leaf-ordered nucleotide rows instead of codon codes
    // and the device does pleaves for every frame (pcsf_batch_upload_alignments)
    // The rows are not copied into one buffer: the batch keeps the readers' buffers (shared with the reader
    // side only until it lets go of them) as the pieces of pcsf_batch_upload_alignments_parts.
    bool nt_form = false;
    struct NtPart {
        std::shared_ptr<std::vector<uint8_t>> buf;
        size_t begin, end;               // the bytes of `buf` this batch uses
    };
    std::vector<NtPart> parts;
    int64_t nt_bytes = 0;                // total size of the pieces = next alignment's offset
    std::vector<int64_t> aln_off;        // alignment a: n_leaves rows of aln_len[a] bytes at aln_off[a] of the pieces' concatenation
    std::vector<int32_t> aln_len;
};

// One GPU: a compute context with the tree and models installed, and the three strategies over a Batch.
class DeviceScorer {
  public:
    DeviceScorer(const Options& o, const ParamSet& pset, int device) : opt(o), ps(pset) {
        n_leaves = ps.tree.n_leaves;
        leaf_labels.assign(ps.tree.labels.begin(), ps.tree.labels.begin() + n_leaves);
        if (o.bls) bls_table = BlsTable(ps.nt, leaf_labels);
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
    // The device is held only while scoring: the report lines of a batch are rendered while the next batch of
    // this device is already being uploaded and scored.
    std::string run(const Batch& b) {
        const int64_t R = (int64_t)b.region_off.size() - 1;
        std::vector<ScoreRecord> rec(b.n_rec);
        if (R > 0) {
            std::lock_guard<std::mutex> device_lock(mu);
            cur = &b;
            if (opt.strategy != STRAT_NOP) {
                if (b.nt_form) {
                    std::vector<const uint8_t*> pp;
                    std::vector<int64_t> pb;
                    for (const Batch::NtPart& q : b.parts) {
                        pp.push_back(q.buf->data() + q.begin);
                        pb.push_back((int64_t)(q.end - q.begin));
                    }
                    check(pcsf_batch_upload_alignments_parts(ctx, (int64_t)b.aln_len.size(), b.aln_off.data(), b.aln_len.data(),
                                                             (int64_t)pp.size(), pp.data(), pb.data(), opt.frames));
                } else
                    check(pcsf_batch_upload(ctx, R, b.region_off.data(), b.codes.data()));
            }
            switch (opt.strategy) {
                case STRAT_FIXED: if (frame_mode()) score_fixed_frames(rec); else score_fixed(rec); break;
                case STRAT_MLE: score_mle(rec); break;
                case STRAT_OMEGA: score_omega(rec); break;
                case STRAT_NOP: break;
            }
            cur = nullptr;
        }
        std::ostringstream out;
        for (const AlnJob& j : b.jobs) report(j, rec, out);
        return out.str();
    }

    int64_t evaluations = 0;  // likelihood evaluations (region x model) issued

  private:
    Options opt;
    const ParamSet& ps;
    pcsf_ctx* ctx = nullptr;
    int n_leaves = 0;
    std::vector<std::string> leaf_labels;
    BlsTable bls_table;
    const Batch* cur = nullptr;
    std::mutex mu;  // one batch at a time on the device

    // --bls of region k of alignment j (src/PhyloCSF.ml:252-262)
    double region_bls(const AlnJob& j, size_t k) const {
        if (!j.bls.empty()) return j.bls[k];
        const Region& rg = j.regions[k];
        const std::vector<std::string>& rows = rg.rc ? j.rc_aln : j.aln;
        if (!bls_table.usable()) return bls_score(ps.nt, rows, j.which_row, rg.lo, rg.hi);
        std::vector<int> leaf_row(n_leaves, -1);
        for (int l = 0; l < n_leaves; l++) {
            auto it = j.which_row.find(leaf_labels[l]);
            if (it != j.which_row.end()) leaf_row[l] = it->second;
        }
        return bls_table.region(rg.lo, rg.hi, [&](int i) {
            BlsTable::Mask m;
            for (int l = 0; l < n_leaves; l++)
                if (leaf_row[l] >= 0 && BlsTable::counts(rows[leaf_row[l]][i])) m.set(l);
            return m;
        });
    }

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
            rec[r].set_diag(ScoreRecord::D_FIXED, {1.0, db(lpr[r]), db(lpr[R + r])});
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
                rc.set_diag(ScoreRecord::D_FIXED, {1.0, db(l[0]), db(l[1])});
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
            rec[r].set_diag(ScoreRecord::D_MLE, {1.0, rho[0][r], rho[1][r], db(lpr[0][r]), db(lpr[1][r])});
            rec[r].random_init = ((st[0][r] | st[1][r]) & PCSF_ST_RANDOM_INIT) != 0;
        }
    }

    // ---- OmegaModel.score (src/OmegaModel.ml:195-219): host/omega_strategy.hpp ----
    void score_omega(std::vector<ScoreRecord>& rec) {
        const int64_t R = (int64_t)rec.size();
        OmegaStrategy om(ctx, n_leaves);
        std::vector<double> score(R), diag(10 * R);
        std::vector<std::string> exn(R);
        om.score(R, cur->region_off.data(), cur->codes.data(), opt.omega_H1, opt.sigma_H1, score.data(), diag.data(), exn);
        evaluations += om.evaluations;
        for (int64_t r = 0; r < R; r++) {
            if (!exn[r].empty()) { rec[r].exn = exn[r]; continue; }
            rec[r].score = score[r];
            rec[r].anc_comp = std::nan("");
            const double* d = &diag[10 * r];
            rec[r].set_diag(ScoreRecord::D_OMEGA, {d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9]});
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
        const auto da = a.diag(), dbb = b.diag();
        if (da != dbb) return da > dbb;
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
            if (opt.bls) out << "\t" << fmt4(region_bls(j, (size_t)k));
            if (opt.anc_comp) out << "\t" << fmt4(s.anc_comp);
            const std::string refdna = (opt.dna || opt.aa) ? rows[0].substr(rg.lo, rg.hi - rg.lo + 1) : std::string();
            if (opt.dna) out << "\t" << refdna;
            if (opt.aa) out << "\t" << translate(refdna);
            if (opt.debug) {
                out << "\t#";
                for (auto& kv : s.diag()) out << " " << kv.first << "=" << kv.second;
                // not part of the reference's report (stdout stays a drop-in): regions whose search started from
                // Fit.find_init's random branch depend on the restated OCaml Random stream, which no reference test pins
                if (s.random_init) std::cerr << j.name << "\t" << rg.lo << "\t" << rg.hi << "\tnote: find_init took its random branch (PCSF_ST_RANDOM_INIT)\n";
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
        for (int l = 0; l < n_leaves; l++) leaf_index[leaf_labels[l]] = l;
        if (o.bls) bls_table = BlsTable(ps.nt, leaf_labels);
        for (int c = 0; c < 256; c++) nt_lut[c] = 0;
        for (const char* q = "ACGTacgtNn-"; *q; q++) nt_lut[(unsigned char)*q] = (uint8_t)*q;  // Code.ml:39-51
        nt_lut[(unsigned char)'u'] = 't';                                                        // src/PhyloCSF.ml:266,285
        nt_lut[(unsigned char)'U'] = 'T';
        if (devices.empty()) devices.push_back(o.device);
        if (o.strategy == STRAT_NOP) devices.resize(1);
        // The scoring contexts (CUDA initialisation, tree and model upload, P(t) tables) come up in parallel, one
        // thread each, and are all up before reading starts.
        for (int d : devices)
            dev_init.push_back(std::async(std::launch::async, [this, d]() { return std::unique_ptr<DeviceScorer>(new DeviceScorer(opt, ps, d)); }));
        devices_ready();
    }
    Driver(const Driver&) = delete;
    ~Driver() {
        for (auto& f : dev_init)
            if (f.valid()) f.wait();
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
        bool nt_form = false;     // prepared by prepare_fast: n_leaves rows of aln_len nucleotides at offset
        size_t nt_off = 0;        // nt_off of the buffer the caller passed (shared by a chunk of alignments)
        int32_t aln_len = 0;
    };

    // Fast form of prepare() for the bulk case: AsIs regions of a plain, well-formed multi-FASTA text. One pass
    // over the bytes finds the records; the rows go straight into a leaf-ordered nucleotide block (u->t applied,
    // alphabet checked with the revcomp table) that the device turns into codon codes for every frame, so the
    // host builds neither per-row strings nor reverse complements nor codes. Anything unusual - an option that
    // needs the rows on the host, blank or padded lines, a character outside ACGTacgtNn-uU, an unknown or
    // repeated species, ragged rows, a gapped reference - returns false WITHOUT judging it: the caller then
    // runs prepare(), which applies the reference's checks in the reference's order with its messages.
    bool prepare_fast(const std::string& name, const char* data, size_t n, Prepared& p, std::vector<uint8_t>& nt_buf) const {
        if (opt.orf != AsIs || opt.strategy == STRAT_OMEGA || no_fast_reader) return false;
        if (opt.bls && !bls_table.usable()) return false;
        if (n == 0 || data[0] != '>') return false;
        struct Seg { const char *b, *e; };
        struct Rec { int leaf; size_t seg0, seg1; };
        static thread_local std::vector<Seg> segs;  // scratch, reused across calls
        static thread_local std::vector<Rec> recs;
        static thread_local std::vector<char> seen;
        segs.clear();
        recs.clear();
        seen.assign(n_leaves, 0);
        auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\014'; };
        const char *q = data, *end = data + n;
        std::string key;
        while (q < end) {
            const char* nl = (const char*)memchr(q, '\n', (size_t)(end - q));
            const char* le = nl ? nl : end;
            if (*q == '>') {  // header: species = text up to the first '|', trimmed (src/PhyloCSF.ml:112-114)
                const char* b = q + 1;
                const char* bar = (const char*)memchr(b, '|', (size_t)(le - b));
                const char* e = bar ? bar : le;
                while (b < e && ws(*b)) b++;
                while (e > b && ws(e[-1])) e--;
                key.assign(b, e);
                auto it = leaf_index.find(key);
                if (it == leaf_index.end() || seen[it->second]) return false;
                seen[it->second] = 1;
                if (!recs.empty()) recs.back().seg1 = segs.size();
                recs.push_back({it->second, segs.size(), segs.size()});
            } else {
                const char *b = q, *e = le;
                while (e > b && ws(e[-1])) e--;  // trailing CR / blanks; anything else odd fails the alphabet check
                if (b == e || ws(*b)) return false;
                segs.push_back({b, e});
            }
            q = nl ? nl + 1 : end;
        }
        recs.back().seg1 = segs.size();
        auto rec_len = [&](const Rec& r) {
            size_t L = 0;
            for (size_t k = r.seg0; k < r.seg1; k++) L += (size_t)(segs[k].e - segs[k].b);
            return L;
        };
        size_t len = rec_len(recs[0]);
        if (len == 0 || len > (size_t)INT32_MAX) return false;
        for (const Rec& r : recs)
            if (rec_len(r) != len) return false;
        const size_t off0 = nt_buf.size();
        nt_buf.resize(off0 + (size_t)n_leaves * len);
        uint8_t* const block = nt_buf.data() + off0;
        if ((int)recs.size() < n_leaves)
            for (int l = 0; l < n_leaves; l++)
                if (!seen[l]) memset(block + (size_t)l * len, '-', len);  // absent species marginalise
        auto reject = [&]() { nt_buf.resize(off0); return false; };
        for (const Rec& r : recs) {
            uint8_t* dst = block + (size_t)r.leaf * len;
            unsigned bad = 0;
            for (size_t k = r.seg0; k < r.seg1; k++) {
                const unsigned char* sb = (const unsigned char*)segs[k].b;
                const size_t L = (size_t)(segs[k].e - segs[k].b);
                size_t i = 0;
#if defined(__SSE2__)
                // 16 characters at a time: copied as they are when all of them are in ACGTNacgtn- (a letter's
                // upper case is the byte with bit 5 cleared); a block holding anything else - u/U to map, or a
                // character to reject - goes through the table below
                const __m128i cA = _mm_set1_epi8('A'), cC = _mm_set1_epi8('C'), cG = _mm_set1_epi8('G'), cT = _mm_set1_epi8('T'),
                              cN = _mm_set1_epi8('N'), cDash = _mm_set1_epi8('-'), up = _mm_set1_epi8((char)0xDF);
                for (; i + 16 <= L; i += 16) {
                    const __m128i x = _mm_loadu_si128((const __m128i*)(sb + i));
                    const __m128i u = _mm_and_si128(x, up);
                    __m128i ok = _mm_or_si128(_mm_cmpeq_epi8(u, cA), _mm_cmpeq_epi8(u, cC));
                    ok = _mm_or_si128(ok, _mm_or_si128(_mm_cmpeq_epi8(u, cG), _mm_cmpeq_epi8(u, cT)));
                    ok = _mm_or_si128(ok, _mm_or_si128(_mm_cmpeq_epi8(u, cN), _mm_cmpeq_epi8(x, cDash)));
                    if (_mm_movemask_epi8(ok) != 0xFFFF) break;
                    _mm_storeu_si128((__m128i*)(dst + i), x);
                }
#endif
                for (; i < L; i++) {  // the tail, or the rest of a segment that left the fast loop
                    const uint8_t m = nt_lut[sb[i]];
                    bad |= (m == 0);
                    dst[i] = m;
                }
                dst += L;
            }
            if (bad) return reject();
        }
        if (opt.remove_ref_gaps && memchr(block + (size_t)recs[0].leaf * len, '-', len)) {
            // --removeRefGaps (src/PhyloCSF.ml:122-132): drop the columns that are gapped in the reference row, in
            // place: rows are compacted in ascending order, so a row never overwrites one still to be read
            static thread_local std::vector<uint32_t> keep;
            keep.clear();
            const uint8_t* ref0 = block + (size_t)recs[0].leaf * len;
            for (size_t i = 0; i < len; i++)
                if (ref0[i] != '-') keep.push_back((uint32_t)i);
            const size_t nl = keep.size();
            if (nl == 0) return reject();
            for (int l = 0; l < n_leaves; l++) {
                const uint8_t* src = block + (size_t)l * len;
                uint8_t* dst = block + (size_t)l * nl;
                for (size_t j = 0; j < nl; j++) dst[j] = src[keep[j]];
            }
            len = nl;
            nt_buf.resize(off0 + (size_t)n_leaves * len);
        }
        const uint8_t* ref = block + (size_t)recs[0].leaf * len;
        if (!opt.allow_ref_gaps && memchr(ref, '-', len)) return reject();
        AlnJob& job = p.job;
        job.name = name;
        job.aln.assign(1, std::string((const char*)ref, len));  // the report reads only the reference row (--dna / --aa)
        if (opt.frames == 6 && (opt.dna || opt.aa)) job.rc_aln.assign(1, revcomp(job.aln[0]));
        job.regions = candidate_regions(job.aln[0], AsIs, opt.frames, opt.min_codons);
        for (const Region& r : job.regions) p.region_cols.push_back(r.hi - r.lo + 1 >= 3 ? (r.hi - r.lo + 1) / 3 : 0);
        if (opt.bls) {  // per position: which leaves carry a nucleotide -> the pruned tree's length (memoised)
            static thread_local std::vector<BlsTable::Mask> masks;
            masks.assign(len, BlsTable::Mask());
            for (int l = 0; l < n_leaves; l++) {
                if (!seen[l]) continue;
                const uint8_t* row = block + (size_t)l * len;
                for (size_t i = 0; i < len; i++)
                    if (BlsTable::counts((char)row[i])) masks[i].set(l);
            }
            for (const Region& r : job.regions)  // a '-' strand region reads the rows backwards
                job.bls.push_back(bls_table.region(r.lo, r.hi, [&](int i) { return masks[r.rc ? len - 1 - (size_t)i : (size_t)i]; }));
        }
        p.nt_form = true;
        p.nt_off = off0;
        p.aln_len = (int32_t)len;
        return true;
    }
    // prepare() on raw text: the fast form when it applies, else the general one on the text's lines
    Prepared prepare_text(const std::string& name, const char* data, size_t n, std::vector<uint8_t>& nt_buf) const {
        Prepared p;
        if (prepare_fast(name, data, n, p, nt_buf)) {
            n_fast++;
            return p;
        }
        n_general++;
        std::vector<std::string> lines;
        const char *q = data, *end = data + n;
        while (q < end) {  // like std::getline
            const char* nl = (const char*)memchr(q, '\n', (size_t)(end - q));
            lines.emplace_back(q, nl ? nl : end);
            q = nl ? nl + 1 : end;
        }
        return prepare(name, lines);
    }
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
    // `nt_buf`: the buffer prepare_text filled for this alignment (nucleotide form only).
    bool append(Prepared&& p, std::ostream& out, const std::shared_ptr<std::vector<uint8_t>>& nt_buf = nullptr) {
        if (!p.abort.empty()) {
            finish(out);
            out << p.job.name << "\tabort\t" << p.abort << "\n";
            out.flush();
            return false;
        }
        if (!batch.jobs.empty() && batch.nt_form != p.nt_form) flush(out);  // a batch is staged in one form
        batch.nt_form = p.nt_form;
        if (batch.jobs.empty() && !p.nt_form) {  // size the staging buffer once instead of growing (and re-faulting) it by doubling
            const size_t want = (size_t)(opt.batch_cols + 4096) * n_leaves + p.codes.size();
            if (batch.codes.capacity() < want) batch.codes.reserve(want);
        }
        p.job.first_region = (int64_t)batch.region_off.size() - 1;
        p.job.first_rec = batch.n_rec;
        batch.n_rec += (int64_t)p.job.regions.size();
        for (int nc : p.region_cols) batch.region_off.push_back(batch.region_off.back() + nc);
        if (p.nt_form) {
            const size_t nb = (size_t)p.aln_len * n_leaves;
            if (batch.parts.empty() || batch.parts.back().buf != nt_buf || batch.parts.back().end != p.nt_off)
                batch.parts.push_back({nt_buf, p.nt_off, p.nt_off});
            batch.parts.back().end += nb;
            batch.aln_off.push_back(batch.nt_bytes);
            batch.aln_len.push_back(p.aln_len);
            batch.nt_bytes += (int64_t)nb;
        } else {
            batch.codes.insert(batch.codes.end(), p.codes.begin(), p.codes.end());
        }
        batch.jobs.push_back(std::move(p.job));
        if (batch.region_off.back() >= opt.batch_cols) flush(out);
        return true;
    }

    bool add_alignment(const std::string& name, const std::vector<std::string>& lines, std::ostream& out) {
        return append(prepare(name, lines), out);
    }
    size_t jobs_in_batch() const { return batch.jobs.size(); }
    mutable std::atomic<int64_t> n_fast{0}, n_general{0};  // alignments taken by the fast / the general reader

    // Hand the current batch to the next GPU (round-robin). Up to two batches per device are in flight: one
    // being scored, and either one waiting for the device or one whose report is being rendered.
    void flush(std::ostream& out) {
        if (batch.jobs.empty()) return;
        devices_ready();
        if (inflight.size() >= 2 * dev.size()) drain_one(out);
        auto b = std::make_shared<Batch>(std::move(batch));
        batch = Batch();
        if (!spare.empty()) {  // staging buffers of a finished batch: their pages are already faulted in
            batch.codes.swap(spare.back().codes);
            batch.region_off.swap(spare.back().region_off);
            batch.aln_off.swap(spare.back().aln_off);
            batch.aln_len.swap(spare.back().aln_len);
            spare.pop_back();
        }
        DeviceScorer* d = dev[next_dev++ % dev.size()].get();
        inflight.push_back({std::async(std::launch::async, [d, b]() { return d->run(*b); }), b});
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
    std::unordered_map<std::string, int> leaf_index;  // species -> tree leaf
    BlsTable bls_table;
    uint8_t nt_lut[256];                              // alignment character -> staged character, 0 = not allowed
    const bool no_fast_reader = std::getenv("PCSF_NO_FAST_READER") != nullptr;  // tests: everything through the general reader
    std::vector<std::unique_ptr<DeviceScorer>> dev;
    std::vector<std::future<std::unique_ptr<DeviceScorer>>> dev_init;
    Batch batch;
    std::deque<std::pair<std::future<std::string>, std::shared_ptr<Batch>>> inflight;
    std::vector<Batch> spare;  // emptied staging buffers, capacity kept
    size_t next_dev = 0;

    void devices_ready() {
        for (auto& f : dev_init) dev.push_back(f.get());  // rethrows a context's start-up failure
        dev_init.clear();
    }

    void drain_one(std::ostream& out) {
        out << inflight.front().first.get();
        out.flush();
        Batch& b = *inflight.front().second;
        Batch keep;
        keep.codes.swap(b.codes);
        keep.region_off.swap(b.region_off);
        keep.aln_off.swap(b.aln_off);
        keep.aln_len.swap(b.aln_len);
        keep.codes.clear();
        keep.region_off.assign(1, 0);
        keep.aln_off.clear();
        keep.aln_len.clear();
        spare.push_back(std::move(keep));
        inflight.pop_front();
    }
};

}  // namespace host
}  // namespace pcsf
