// paramset.hpp — PhyloCSF.initialize_strategy's loading half (src/PhyloCSF.ml:406-445): parameter
// file resolution, tree (+ --species pruning), ECMs and their diagonalisation.
#pragma once
#include <set>
#include <sstream>
#include <string>

#include "codon_model.hpp"
#include "newick_tree.hpp"

namespace pcsf {
namespace host {

struct ParamSet {
    NewickPtr nt;  // after --species pruning
    Tree tree;
    bool have_ecm = false;
    ECM ecm[2];
    QDiag qd[2];  // 0 = coding, 1 = noncoding
};

inline std::string slurp_required(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw failure("could not find required parameter file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

inline ParamSet load_paramset(const std::string& prefix, const std::string& species_csv, bool with_ecm) {
    ParamSet ps;
    NewickPtr nt = newick_parse(slurp_required(prefix + ".nh"));
    if (!species_csv.empty()) {
        std::set<std::string> want;
        std::stringstream ss(species_csv);
        std::string tok;
        while (std::getline(ss, tok, ',')) want.insert(tok);
        NewickPtr snt = newick_subtree([&](const std::string& s) { return want.count(s) > 0; }, nt);
        if (!snt || newick_leaves(*snt) <= 1) throw failure("specify at least two available --species");
        nt = snt;
    }
    ps.nt = nt;
    ps.tree = Tree::of_newick(*nt);
    if (with_ecm) {
        {
            std::ifstream c(prefix + "_coding.ECM"), n(prefix + "_noncoding.ECM");
            if (!c) throw failure("could not find required parameter file " + prefix + "_coding.ECM");
            if (!n) throw failure("could not find required parameter file " + prefix + "_noncoding.ECM");
        }
        ps.ecm[0] = read_ecm(prefix + "_coding.ECM");
        ps.ecm[1] = read_ecm(prefix + "_noncoding.ECM");
        for (int w = 0; w < 2; w++) ps.qd[w] = QDiag::of_reversible_Q(ecm_q(ps.ecm[w]), ps.ecm[w].pi);
        ps.have_ecm = true;
    }
    return ps;
}

}  // namespace host
}  // namespace pcsf
