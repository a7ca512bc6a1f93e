// alignment.hpp — host-side alignment handling of the drop-in driver: multi-FASTA reader, reference
// gap removal, ORF search, candidate regions, leaf coding, branch-length score, translation.
//   input_mfa              <- src/PhyloCSF.ml:83-118
//   maybe_remove_ref_gaps  <- src/PhyloCSF.ml:122-132
//   find_orfs              <- src/PhyloCSF.ml:134-196
//   candidate_regions      <- src/PhyloCSF.ml:198-217
//   pleaves                <- src/PhyloCSF.ml:219-246
//   bls                    <- src/PhyloCSF.ml:252-262
//   translate              <- src/PhyloCSF.ml:268-278
#pragma once
#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "codon_model.hpp"
#include "newick_tree.hpp"

namespace pcsf {
namespace host {

enum OrfMode { AsIs, ATGStop, StopStop, StopStop3, ToFirstStop, FromLastStop, ToOrFromStop };

struct Region {
    bool rc;
    int lo, hi;
};

struct Alignment {
    std::vector<std::string> species, seqs;
};

// Deviation (documented in DESIGN.md): the reference spins forever on a blank line inside a record
// (PhyloCSF.ml:89-90 peeks without consuming); blank lines are skipped here.
inline Alignment input_mfa(const std::vector<std::string>& lines) {
    Alignment a;
    try {
        for (const std::string& raw : lines) {
            const std::string line = ocaml_trim(raw);
            if (!a.species.empty() && line.empty()) continue;
            if (a.species.empty() || (!line.empty() && line[0] == '>')) {
                if (raw.empty()) continue;
                if (raw[0] != '>') throw failure("bad header");
                std::string hdr = raw.substr(1);
                const size_t bar = hdr.find('|');
                a.species.push_back(ocaml_trim(bar == std::string::npos ? hdr : hdr.substr(0, bar)));
                a.seqs.emplace_back();
            } else {
                a.seqs.back() += line;
            }
        }
        if (a.species.empty()) throw HostError("No_value");
        const size_t seqlen = a.seqs[0].size();
        bool bad = seqlen == 0;
        for (auto& s : a.species) bad = bad || s.empty();
        for (auto& s : a.seqs) bad = bad || s.size() != seqlen;
        if (bad) throw failure("empty species name or sequence, or sequence length mismatch");
    } catch (const HostError& e) {
        std::string m = e.what();
        // the reference re-wraps: Failure msg -> "invalid MFA alignment: msg"; others -> Printexc text
        if (m.rfind("Failure(\"", 0) == 0) m = m.substr(9, m.size() - 11);
        throw failure("invalid MFA alignment: " + m);
    }
    return a;
}

inline void remove_ref_gaps(std::vector<std::string>& aln) {
    std::vector<size_t> keep;
    for (size_t j = 0; j < aln[0].size(); j++)
        if (aln[0][j] != '-') keep.push_back(j);
    for (auto& row : aln) {
        std::string r;
        r.reserve(keep.size());
        for (size_t j : keep) r.push_back(row[j]);
        row.swap(r);
    }
}

inline std::vector<std::pair<int, int>> find_orfs(const std::string& dna, int ofs, OrfMode mode, int min_codons) {
    const bool atg = mode == ATGStop;
    auto up = [&](int p) { return (char)std::toupper((unsigned char)dna[p]); };
    auto is_start = [&](int p) { return up(p) == 'A' && up(p + 1) == 'T' && up(p + 2) == 'G'; };
    auto is_stop = [&](int p) {
        return up(p) == 'T' && ((up(p + 1) == 'A' && (up(p + 2) == 'A' || up(p + 2) == 'G')) || (up(p + 1) == 'G' && up(p + 2) == 'A'));
    };
    const int len = (int)dna.size();
    std::vector<std::pair<int, int>> orfs;  // most recent first, like the OCaml list
    std::vector<int> starts;                // most recent first
    for (int codon_lo = ofs; codon_lo <= len - 3; codon_lo++) {
        if ((codon_lo - ofs) % 3 != 0) continue;
        const int codon_hi = codon_lo + 2;
        if ((!atg && starts.empty() && !is_stop(codon_lo)) || (atg && is_start(codon_lo))) starts.insert(starts.begin(), codon_lo);
        if (codon_hi + 3 < len && is_stop(codon_hi + 1)) {
            for (int start : starts)
                if (codon_hi > start + 2) orfs.insert(orfs.begin(), {start, codon_hi});
            starts.clear();
        }
    }
    if (!atg)
        for (int start : starts) {
            const int rem = len - start;
            orfs.insert(orfs.begin(), {start, start + (rem / 3) * 3 - 1});
        }
    if (mode == StopStop3) {
        std::vector<std::pair<int, int>> all;
        for (auto& o : orfs) {
            std::vector<std::pair<int, int>> sub{o};
            const int lo = o.first, hi = o.second, codons = (hi - lo + 1) / 3;
            const int lo2 = lo + (codons / 3) * 3;
            if (lo2 > lo) sub.insert(sub.begin(), {lo2, hi});
            const int lo3 = lo + (2 * codons / 3) * 3;
            if (lo3 > lo2 && lo3 > lo) sub.insert(sub.begin(), {lo3, hi});
            all.insert(all.end(), sub.begin(), sub.end());
        }
        orfs.swap(all);
    }
    if (mode == ToFirstStop && !orfs.empty()) {
        const auto first = orfs.back();
        orfs.clear();
        if (first.first == ofs) orfs.push_back(first);
    }
    if (mode == FromLastStop && !orfs.empty()) {
        const auto last = orfs.front();
        orfs.clear();
        if (len - last.second <= 3) orfs.push_back(last);
    }
    if (mode == ToOrFromStop && !orfs.empty()) {
        const auto first = orfs.back(), last = orfs.front();
        orfs.clear();
        if (first.first == ofs) orfs.push_back(first);
        if (len - last.second <= 3 && first != last) orfs.insert(orfs.begin(), last);
    }
    std::vector<std::pair<int, int>> out;
    for (auto it = orfs.rbegin(); it != orfs.rend(); ++it)
        if ((it->second - it->first + 1) / 3 >= min_codons) out.push_back(*it);
    return out;
}

inline std::vector<Region> candidate_regions(const std::string& dna, OrfMode mode, int frames, int min_codons) {
    std::vector<Region> r;
    if (mode == AsIs) {
        const int hi = (int)dna.size() - 1;
        r.push_back({false, 0, hi});
        if (frames != 1) { r.push_back({false, 1, hi}); r.push_back({false, 2, hi}); }
        if (frames == 6) { r.push_back({true, 0, hi}); r.push_back({true, 1, hi}); r.push_back({true, 2, hi}); }
        return r;
    }
    auto add = [&](bool rc, const std::string& s, int ofs) {
        for (auto& o : find_orfs(s, ofs, mode, min_codons)) r.push_back({rc, o.first, o.second});
    };
    add(false, dna, 0);
    if (frames != 1) { add(false, dna, 1); add(false, dna, 2); }
    if (frames == 6) {
        const std::string rcdna = revcomp(dna);
        add(true, rcdna, 0);
        add(true, rcdna, 1);
        add(true, rcdna, 2);
    }
    return r;
}

// Appends the codon codes of region [lo,hi] to `codes` ([col][leaf]); returns the number of columns.
inline int pleaves(int n_leaves, const std::vector<int>& leaf_ord, const std::vector<std::string>& aln, int lo, int hi,
                   std::vector<uint8_t>& codes) {
    int ncols = 0;
    for (int pos = lo; pos + 2 <= hi; pos += 3, ncols++)
        for (int l = 0; l < n_leaves; l++) {
            const int r = leaf_ord[l];
            codes.push_back(r < 0 ? (uint8_t)CODE_MARG : (uint8_t)codon_code(aln[r][pos], aln[r][pos + 1], aln[r][pos + 2]));
        }
    return ncols;
}

inline double bls_score(const NewickPtr& nt, const std::vector<std::string>& aln, const std::map<std::string, int>& which_row,
                        int lo, int hi) {
    double total = 0.0;
    for (int i = lo; i <= hi; i++) {
        NewickPtr st = newick_subtree(
            [&](const std::string& sp) {
                auto it = which_row.find(sp);
                if (it == which_row.end()) return false;
                const char c = aln[it->second][i];
                return !(c == '-' || c == '.' || c == 'N');
            },
            nt);
        total += st ? newick_total_length(*st) : 0.0;
    }
    return total / (newick_total_length(*nt) * (double)(hi - lo + 1));
}

// The same score, column by column, as a function of WHICH tree leaves are present in the column: the pruned
// tree's total length is computed with the reference's own recipe (Newick.subtree + total_length, as above) once
// per distinct set of leaves and memoised per thread, so every region of every alignment sums precomputed
// doubles in the reference's order. Bit-identical to bls_score.
class BlsTable {
  public:
    struct Mask {
        uint64_t w[4] = {0, 0, 0, 0};
        void set(int l) { w[l >> 6] |= 1ull << (l & 63); }
        bool operator==(const Mask& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
    };
    static constexpr int kMaxLeaves = 256;
    BlsTable() = default;
    BlsTable(NewickPtr tree, const std::vector<std::string>& leaf_labels) : nt_(std::move(tree)) {
        for (size_t l = 0; l < leaf_labels.size(); l++) index_[leaf_labels[l]] = (int)l;
        total_ = newick_total_length(*nt_);
        usable_ = (int)leaf_labels.size() <= kMaxLeaves;
    }
    bool usable() const { return usable_; }
    static bool counts(char c) { return !(c == '-' || c == '.' || c == 'N'); }  // src/PhyloCSF.ml:257
    double column(const Mask& m) const {
        static thread_local const BlsTable* owner = nullptr;
        static thread_local std::unordered_map<Mask, double, Hash> memo;
        if (owner != this) {
            memo.clear();
            owner = this;
        }
        auto it = memo.find(m);
        if (it != memo.end()) return it->second;
        NewickPtr st = newick_subtree(
            [&](const std::string& sp) {
                auto f = index_.find(sp);
                return f != index_.end() && ((m.w[f->second >> 6] >> (f->second & 63)) & 1);
            },
            nt_);
        const double v = st ? newick_total_length(*st) : 0.0;
        if (memo.size() > (1u << 20)) memo.clear();
        memo.emplace(m, v);
        return v;
    }
    // score of positions lo..hi given a way to get the mask of position i
    template <class MaskAt>
    double region(int lo, int hi, MaskAt&& mask_at) const {
        double total = 0.0;
        for (int i = lo; i <= hi; i++) total += column(mask_at(i));
        return total / (total_ * (double)(hi - lo + 1));
    }

  private:
    struct Hash {
        size_t operator()(const Mask& m) const {
            uint64_t h = 0x9e3779b97f4a7c15ull;
            for (uint64_t x : m.w) h = (h ^ x) * 0xff51afd7ed558ccdull, h ^= h >> 32;
            return (size_t)h;
        }
    };
    NewickPtr nt_;
    std::unordered_map<std::string, int> index_;
    double total_ = 0.0;
    bool usable_ = false;
};

inline std::string translate(const std::string& dna) {
    std::string pp(dna.size() / 3, '?');
    for (size_t i = 0; i < pp.size(); i++) {
        const int c = codon_code(dna[3 * i], dna[3 * i + 1], dna[3 * i + 2]);
        if (c != CODE_MARG) pp[i] = kTranslation[c];
    }
    return pp;
}

}  // namespace host
}  // namespace pcsf
