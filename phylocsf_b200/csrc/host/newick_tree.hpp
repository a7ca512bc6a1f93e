// newick_tree.hpp — host-side tree plumbing of the drop-in: Newick parsing/pruning and the T.t
// node numbering that is part of the kernels' input contract.
//   Newick lexer/parser  <- lib/CamlPaml/NewickLexer.mll:5-14, NewickParser.mly:7-24
//   subtree/total_length <- lib/CamlPaml/Newick.ml:34-41,52-61
//   Tree::of_newick      <- lib/CamlPaml/T.ml:57-112 (leaves left-to-right, internals post-order)
#pragma once
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pcsf {
namespace host {

// Mirrors the reference's exception tiers: what() carries Printexc.to_string's text.
struct HostError : std::runtime_error {
    explicit HostError(const std::string& m) : std::runtime_error(m) {}
};
inline HostError failure(const std::string& m) { return HostError("Failure(\"" + m + "\")"); }
inline HostError invalid_arg(const std::string& m) { return HostError("Invalid_argument(\"" + m + "\")"); }

struct NewickNode {
    std::vector<std::shared_ptr<NewickNode>> children;
    std::string label;
    bool has_bl = false;
    double bl = 0.0;
};
using NewickPtr = std::shared_ptr<NewickNode>;

class NewickParser {
  public:
    explicit NewickParser(const std::string& text) : s_(text) { next(); }
    NewickPtr parse() {
        NewickPtr r = node();
        if (tok_ != T_EOF) throw failure("Newick: parse error (trailing input)");
        return r;
    }

  private:
    enum Tok { T_LP, T_RP, T_COMMA, T_COLON, T_BL, T_LABEL, T_EOF };
    const std::string& s_;
    size_t pos_ = 0;
    Tok tok_ = T_EOF;
    std::string lex_;
    static bool is_bl(char c) { return (c >= '0' && c <= '9') || c == '.'; }
    static bool is_lbl(char c) { return is_bl(c) || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_'; }
    void next() {
        while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\t' || s_[pos_] == '\r' || s_[pos_] == '\n' || s_[pos_] == ';')) pos_++;
        if (pos_ >= s_.size()) { tok_ = T_EOF; return; }
        const char c = s_[pos_];
        if (c == '(') { tok_ = T_LP; pos_++; return; }
        if (c == ')') { tok_ = T_RP; pos_++; return; }
        if (c == ',') { tok_ = T_COMMA; pos_++; return; }
        if (c == ':') { tok_ = T_COLON; pos_++; return; }
        if (!is_lbl(c)) throw failure("Newick: illegal character");
        size_t e = pos_;
        bool all_bl = true;
        while (e < s_.size() && is_lbl(s_[e])) { all_bl = all_bl && is_bl(s_[e]); e++; }
        lex_ = s_.substr(pos_, e - pos_);
        pos_ = e;
        tok_ = all_bl ? T_BL : T_LABEL;  // longest match; BRANCHLEN wins ties (NewickLexer.mll:11-12)
    }
    double take_bl() {
        if (tok_ != T_BL) throw failure("Newick: parse error (branch length expected)");
        size_t used = 0;
        double v;
        try { v = std::stod(lex_, &used); } catch (...) { throw failure("float_of_string"); }
        if (used != lex_.size()) throw failure("float_of_string");
        next();
        return v;
    }
    void label(NewickNode& n) {  // NewickParser.mly:16-20
        if (tok_ == T_LABEL) {
            n.label = lex_;
            next();
            if (tok_ == T_COLON) { next(); n.bl = take_bl(); n.has_bl = true; }
        } else if (tok_ == T_COLON) {
            next();
            n.bl = take_bl();
            n.has_bl = true;
        } else throw failure("Newick: parse error (label expected)");
    }
    NewickPtr node() {  // NewickParser.mly:11-15
        auto n = std::make_shared<NewickNode>();
        if (tok_ == T_LP) {
            next();
            n->children.push_back(node());
            while (tok_ == T_COMMA) { next(); n->children.push_back(node()); }
            if (tok_ != T_RP) throw failure("Newick: parse error (')' expected)");
            next();
            if (tok_ == T_LABEL || tok_ == T_COLON) label(*n);
        } else label(*n);
        return n;
    }
};

inline NewickPtr newick_parse(const std::string& text) { return NewickParser(text).parse(); }

inline int newick_size(const NewickNode& n) {
    int s = 1;
    for (auto& c : n.children) s += newick_size(*c);
    return s;
}
inline int newick_leaves(const NewickNode& n) {
    if (n.children.empty()) return 1;
    int s = 0;
    for (auto& c : n.children) s += newick_leaves(*c);
    return s;
}

// Newick.ml:34-41: keep leaves (and labelled internals) satisfying `keep`; splice unary nodes,
// adding branch lengths (None if either is None).
inline NewickPtr newick_subtree(const std::function<bool(const std::string&)>& keep, const NewickPtr& nd) {
    if (nd->children.empty()) return (nd->label.empty() || keep(nd->label)) ? nd : nullptr;
    if (!(nd->label.empty() || keep(nd->label))) return nullptr;
    std::vector<NewickPtr> st;
    for (auto& c : nd->children) {
        NewickPtr s = newick_subtree(keep, c);
        if (s) st.push_back(s);
    }
    if (st.empty()) return nullptr;
    auto out = std::make_shared<NewickNode>();
    if (st.size() == 1) {
        out->children = st[0]->children;
        out->label = st[0]->label;
        out->has_bl = nd->has_bl && st[0]->has_bl;
        out->bl = out->has_bl ? nd->bl + st[0]->bl : 0.0;
    } else {
        out->children = st;
        out->label = nd->label;
        out->has_bl = nd->has_bl;
        out->bl = nd->bl;
    }
    return out;
}

inline double newick_total_length_rec(const NewickNode& n, bool top) {  // Newick.ml:52-61
    if (!top && !n.has_bl) throw invalid_arg("CamlPaml.Newick.total_length: unspecified branch length");
    double acc = 0.0;
    for (auto& c : n.children) acc = acc + newick_total_length_rec(*c, false);
    return (top ? 0.0 : n.bl) + acc;
}
inline double newick_total_length(const NewickNode& n) { return newick_total_length_rec(n, true); }

struct Tree {
    int n_leaves = 0;
    std::vector<std::string> labels;
    std::vector<int> parents;
    std::vector<std::pair<int, int>> children;  // by node id
    std::vector<double> branches;               // NaN when unspecified (root)
    int size() const { return (int)parents.size(); }
    int root() const { return size() - 1; }
    std::vector<int32_t> children_array() const {
        std::vector<int32_t> out;
        for (int i = n_leaves; i < size(); i++) { out.push_back(children[i].first); out.push_back(children[i].second); }
        return out;
    }
    static Tree of_newick(const NewickNode& nt) {
        const char* bitch = "CamlPaml.T.of_newick: input is not a rooted, bifurcating tree";
        const int n = newick_size(nt);
        if (n < 3 || n % 2 == 0) throw invalid_arg(bitch);
        std::vector<const NewickNode*> leaves;
        std::function<void(const NewickNode&)> find = [&](const NewickNode& x) {
            if (x.children.empty()) leaves.push_back(&x);
            else if (x.children.size() == 2) { find(*x.children[0]); find(*x.children[1]); }
            else throw invalid_arg(bitch);
        };
        find(nt);
        Tree t;
        t.n_leaves = (int)leaves.size();
        t.labels.assign(n, "");
        t.parents.assign(n, -1);
        t.children.assign(n, {-1, -1});
        t.branches.assign(n, std::nan(""));
        int next_leaf = 0, next_internal = t.n_leaves;
        std::function<int(const NewickNode&)> fill = [&](const NewickNode& x) -> int {
            int i;
            if (x.children.empty()) i = next_leaf++;  // leaves are met left to right
            else {
                const int lc = fill(*x.children[0]);
                const int rc = fill(*x.children[1]);
                i = next_internal++;
                t.parents[lc] = i;
                t.parents[rc] = i;
                t.children[i] = {lc, rc};
            }
            t.labels[i] = x.label;
            if (x.has_bl) t.branches[i] = x.bl;
            return i;
        };
        fill(nt);
        return t;
    }
};

}  // namespace host
}  // namespace pcsf
