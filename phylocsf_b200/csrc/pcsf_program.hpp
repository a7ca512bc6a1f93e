// pcsf_program.hpp - the tree programs the pruning kernels run, and the host-side code that builds them from T.children.
// No CUDA in here: pcsf_tree_set (pcsf_api.cu) calls build_tree_programs() and uploads the result; pcsf_host_tree_program
// (pcsf_host_api.cu, include/phylocsf_host.h) hands the same programs to the tests, which interpret them on the CPU against
// the oracle's pruning (tests/test_tree_programs.py) - every variant: plain, table levels 2-4, with and without the
// KEEP / MUL rewrite.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#if defined(__CUDACC__)
#define PCSF_PROG_HD __host__ __device__
#else
#define PCSF_PROG_HD
#endif

namespace pcsf {

// ---- tree program (built on the host from T.children) -----------------------------------------
enum OpKind : int32_t {
    OP_CHERRY = 0,     // cur = G(a) * G(b)                       a, b leaves
    OP_GEMM_LEAF = 1,  // cur = (P_a x cur) * G(b)                a internal child, b sibling leaf
    OP_GEMM_PUSH = 2,  // stack[c] = P_a x cur                    sibling subtree still to come
    OP_GEMM_POP = 3,   // cur = (P_a x cur) * stack[c]
    OP_ROOT = 4,       // z = cur . prior, log z, root posterior . log prior
    // table program (wide form, P sets that carry subtree tables): a cherry (leaves a, b) - or a cherry plus the
    // leaf c next to it - and the contractions up to and including the edge above that subtree are one lookup,
    //   W2[code_a][code_b][.]         = P_v x (G(a) * G(b))
    //   W3[code_a][code_b][code_c][.] = P_u x (W2[code_a][code_b] * G(c))
    //   W4[code_a][code_b][code_c][code_d] = P_w x (W3[code_a][code_b][code_c] * G(d))     (caterpillar of four)
    // Encoding: kind | table << 8, a = leaf a | leaf b << 16, b = leaf c | leaf d << 16 (0xffff = none), c = operand
    // of the usual epilogue
    OP_TAB_LEAF = 5,   // cur = W * G(c)
    OP_TAB_PUSH = 6,   // stack[c] = W
    OP_TAB_POP = 7,    // cur = W * stack[c]
    // A push whose pop is the very next op - the sibling subtree is one table lookup - parks nothing: the message stays in
    // registers and the lookup multiplies into it (same two factors, same product: bit-identical to push + pop).
    OP_TAB_KEEP = 8,   // cur = W                                  (was OP_TAB_PUSH)
    OP_TAB_MUL = 9,    // cur = cur * W                            (was OP_TAB_POP, right after a ..._KEEP)
    OP_GEMM_KEEP = 10  // cur = P_a x cur                          (was OP_GEMM_PUSH)
};
PCSF_PROG_HD inline bool op_is_table(int kind) {
    const int k = kind & 0xff;
    return k >= OP_TAB_LEAF && k <= OP_TAB_MUL;
}
struct Op {
    int32_t kind, a, b, c;
};
constexpr int CHERRY_ROWS = 65 * 65;              // code pairs, 64 = marginalise
constexpr int CHERRY_TABLE = CHERRY_ROWS * 64;    // doubles per cherry table (2.16 MB)
constexpr int TRIPLE_ROWS = 65 * 65 * 65;         // code triples
constexpr long long TRIPLE_TABLE = (long long)TRIPLE_ROWS * 64;  // doubles per 3-leaf table (140.6 MB)
constexpr int QUAD_ROWS = 65 * 65 * 65 * 65;      // code quadruples
constexpr long long QUAD_TABLE = (long long)QUAD_ROWS * 64;      // doubles per 4-leaf table (9.14 GB)
struct SubTab {          // one memoised subtree: a cherry, or the subtree of table `src` plus one more leaf
    int32_t la, lb;      // cherries: the two leaves
    int32_t lnew;        // deeper tables: the leaf that joins the subtree of table `src` (-1 for a cherry)
    int32_t edge;        // the node whose upward edge the table includes
    int32_t src;         // deeper tables: index of the table they are built from
    int32_t pad;
    long long off;       // offset of the table (doubles) in the P set's table block
};

// What the producer warpgroup stages for the compute warps, in program order (built on the host).
enum ItemKind : int32_t {
    ITEM_P = 0,     // fragment-ordered P image of internal edge `a`            -> P ring (TMA bulk copy)
    ITEM_LEAF = 1,  // leaf message of leaf `a`: gathered rows of its P^T table  -> M ring (cp.async gather)
    ITEM_POP = 2    // parked partial of stack level `a`                         -> M ring (TMA bulk copy)
};
struct Item {
    int32_t kind, a;
};


// ---- tree program ------------------------------------------------------------------------------
// Post-order schedule that keeps the partial of the current subtree in registers. At a node with two
// internal children the child needing more live partials goes first (Sethi-Ullman), its message is
// parked on the stack while the other subtree is evaluated.
struct ProgramBuilder {
    int nl;
    const std::vector<int32_t>& ch;
    std::vector<Op> ops;
    std::vector<int> need, leaves;
    int height = 0, max_height = 0, n_gemm = 0;
    ProgramBuilder(int n_leaves, const std::vector<int32_t>& children)
        : nl(n_leaves), ch(children), need(2 * n_leaves - 1, 0), leaves(2 * n_leaves - 1, 1) {}
    int lc(int i) const { return ch[2 * (i - nl)]; }
    int rc(int i) const { return ch[2 * (i - nl) + 1]; }
    void compute_need() {
        for (int i = nl; i < 2 * nl - 1; i++) {  // children precede parents in T numbering
            const int l = lc(i), r = rc(i);
            const bool li = l >= nl, ri = r >= nl;
            leaves[i] = leaves[l] + leaves[r];
            if (!li && !ri) need[i] = 1;
            else if (li && ri) {
                const int a = std::max(need[l], need[r]), b = std::min(need[l], need[r]);
                need[i] = std::max(a, b + 1);
            } else need[i] = need[li ? l : r];
        }
    }
    void emit(int i) {
        // iterative post-order would also do; depth is bounded by the tree height (<= n_leaves)
        const int l = lc(i), r = rc(i);
        const bool li = l >= nl, ri = r >= nl;
        if (!li && !ri) {
            ops.push_back({OP_CHERRY, l, r, 0});
        } else if (li && ri) {
            // Sethi-Ullman: the child that needs more parked partials goes first. On a tie the one with more leaves does: with
            // equal needs of 1 both are caterpillars, and the smaller one is the likelier to be a single table lookup, which as
            // the SECOND subtree costs no stack round trip at all (OP_..._KEEP + OP_TAB_MUL below).
            const int first = need[l] != need[r] ? (need[l] > need[r] ? l : r) : (leaves[l] >= leaves[r] ? l : r);
            const int second = first == l ? r : l;
            emit(first);
            ops.push_back({OP_GEMM_PUSH, first, 0, height});
            n_gemm++;
            height++;
            max_height = std::max(max_height, height);
            emit(second);
            height--;
            ops.push_back({OP_GEMM_POP, second, 0, height});
            n_gemm++;
        } else {
            const int inner = li ? l : r, leaf = li ? r : l;
            emit(inner);
            ops.push_back({OP_GEMM_LEAF, inner, leaf, 0});
            n_gemm++;
        }
    }
};


// Everything pcsf_tree_set derives from the tree's shape: the plain program (both kernel forms), the three table programs
// of the wide form, the producer's item lists and the subtree tables they refer to.
struct TreePrograms {
    std::vector<Op> ops, ops_t, ops_t3, ops_t4;        // plain; level 2 (cherries), 3 (+ cherry-and-leaf), 4 (+ caterpillars of four)
    std::vector<Item> items, items_t, items_t3, items_t4;
    std::vector<SubTab> subtabs;                       // cherries first (n_tab2), then the 3-leaf (n_tab3) and 4-leaf subtrees (n_tab4)
    std::vector<long long> tab_off;                    // SubTab::off again, as the array the kernels index
    int n_tab2 = 0, n_tab3 = 0, n_tab4 = 0, n_gemm = 0, max_levels = 0;
};

// keep = false leaves every push / pop in place (PCSF_NO_KEEP=1: the A/B switch for the measurement in DESIGN section 3)
inline TreePrograms build_tree_programs(int n_leaves, const std::vector<int32_t>& children, bool keep) {
    TreePrograms out;
    ProgramBuilder pb(n_leaves, children);
    pb.compute_need();
    pb.emit(2 * n_leaves - 2);
    pb.ops.push_back({OP_ROOT, 0, 0, 0});
    out.ops = pb.ops;
    out.n_gemm = pb.n_gemm;
    out.max_levels = pb.max_height;
    out.items.clear();
    for (const Op& op : out.ops) {
        if (op.kind == OP_CHERRY) { out.items.push_back({ITEM_LEAF, op.a}); out.items.push_back({ITEM_LEAF, op.b}); }
        else if (op.kind == OP_GEMM_LEAF) { out.items.push_back({ITEM_P, op.a}); out.items.push_back({ITEM_LEAF, op.b}); }
        else if (op.kind == OP_GEMM_PUSH) out.items.push_back({ITEM_P, op.a});
        else if (op.kind == OP_GEMM_POP) { out.items.push_back({ITEM_P, op.a}); out.items.push_back({ITEM_POP, op.c}); }
    }
    // The table programs. Level 2: (OP_CHERRY, the contraction over the edge above the cherry) -> one lookup in the
    // cherry's table. Level 3: when the cherry's sibling is a leaf and the node above them has an edge of its own,
    // (OP_CHERRY, OP_GEMM_LEAF, the contraction over that edge) -> one lookup in the 3-leaf table.
    auto is_gemm = [](const Op& o) { return o.kind == OP_GEMM_LEAF || o.kind == OP_GEMM_PUSH || o.kind == OP_GEMM_POP; };
    auto plain_items = [](const Op& op, std::vector<Item>& items) {
        if (op.kind == OP_CHERRY) { items.push_back({ITEM_LEAF, op.a}); items.push_back({ITEM_LEAF, op.b}); }
        else if (op.kind == OP_GEMM_LEAF) { items.push_back({ITEM_P, op.a}); items.push_back({ITEM_LEAF, op.b}); }
        else if (op.kind == OP_GEMM_PUSH || op.kind == OP_GEMM_POP) items.push_back({ITEM_P, op.a});
    };
    auto table_op = [](const Op& g, int table, int la, int lb, int lc, int ld) {  // g: the contraction the lookup replaces last
        const int kind = (g.kind == OP_GEMM_LEAF ? OP_TAB_LEAF : g.kind == OP_GEMM_PUSH ? OP_TAB_PUSH : OP_TAB_POP) | (table << 8);
        const uint32_t b = (uint32_t)(lc < 0 ? 0xffff : lc) | ((uint32_t)(ld < 0 ? 0xffff : ld) << 16);
        return Op{kind, la | (lb << 16), (int32_t)b, g.kind == OP_GEMM_LEAF ? g.b : g.c};
    };
    out.ops_t.clear();
    out.items_t.clear();
    out.ops_t3.clear();
    out.items_t3.clear();
    out.subtabs.clear();
    std::vector<SubTab> triples, quads;
    std::vector<int> cherry_table_at(out.ops.size(), -1), triple_table_at(out.ops.size(), -1);
    out.ops_t4.clear();
    out.items_t4.clear();
    for (size_t i = 0; i + 1 < out.ops.size(); i++)
        if (out.ops[i].kind == OP_CHERRY && is_gemm(out.ops[i + 1])) {
            cherry_table_at[i] = (int)out.subtabs.size();
            out.subtabs.push_back(SubTab{out.ops[i].a, out.ops[i].b, -1, out.ops[i + 1].a, -1, 0, 0});
        }
    out.n_tab2 = (int)out.subtabs.size();
    for (size_t i = 0; i < out.ops.size(); i++) {  // level 2
        const Op& op = out.ops[i];
        if (cherry_table_at[i] >= 0) {
            const Op& g = out.ops[i + 1];
            out.ops_t.push_back(table_op(g, cherry_table_at[i], op.a, op.b, -1, -1));
            if (g.kind == OP_GEMM_LEAF) out.items_t.push_back({ITEM_LEAF, g.b});
            i++;
            continue;
        }
        out.ops_t.push_back(op);
        plain_items(op, out.items_t);
    }
    for (size_t i = 0; i < out.ops.size(); i++) {  // level 3
        const Op& op = out.ops[i];
        if (cherry_table_at[i] >= 0) {
            const Op& g = out.ops[i + 1];
            if (g.kind == OP_GEMM_LEAF && i + 2 < out.ops.size() && is_gemm(out.ops[i + 2])) {
                const Op& g2 = out.ops[i + 2];  // the edge above the node that joins the cherry and the leaf g.b
                const int ti = out.n_tab2 + (int)triples.size();
                triple_table_at[i] = ti;
                triples.push_back(SubTab{op.a, op.b, g.b, g2.a, cherry_table_at[i], 0, 0});
                out.ops_t3.push_back(table_op(g2, ti, op.a, op.b, g.b, -1));
                if (g2.kind == OP_GEMM_LEAF) out.items_t3.push_back({ITEM_LEAF, g2.b});
                i += 2;
                continue;
            }
            out.ops_t3.push_back(table_op(g, cherry_table_at[i], op.a, op.b, -1, -1));
            if (g.kind == OP_GEMM_LEAF) out.items_t3.push_back({ITEM_LEAF, g.b});
            i++;
            continue;
        }
        out.ops_t3.push_back(op);
        plain_items(op, out.items_t3);
    }
    out.n_tab3 = (int)triples.size();
    for (size_t i = 0; i < out.ops.size(); i++) {  // level 4: a further leaf d joins the 3-leaf subtree
        const Op& op = out.ops[i];
        if (cherry_table_at[i] >= 0) {
            const Op& g = out.ops[i + 1];
            if (triple_table_at[i] >= 0) {
                const Op& g2 = out.ops[i + 2];
                if (g2.kind == OP_GEMM_LEAF && i + 3 < out.ops.size() && is_gemm(out.ops[i + 3])) {
                    const Op& g3 = out.ops[i + 3];  // the edge above the node that joins the 3-leaf subtree and the leaf g2.b
                    const int ti = out.n_tab2 + out.n_tab3 + (int)quads.size();
                    quads.push_back(SubTab{op.a, op.b, g2.b, g3.a, triple_table_at[i], 0, 0});
                    out.ops_t4.push_back(table_op(g3, ti, op.a, op.b, g.b, g2.b));
                    if (g3.kind == OP_GEMM_LEAF) out.items_t4.push_back({ITEM_LEAF, g3.b});
                    i += 3;
                    continue;
                }
                out.ops_t4.push_back(table_op(g2, triple_table_at[i], op.a, op.b, g.b, -1));
                if (g2.kind == OP_GEMM_LEAF) out.items_t4.push_back({ITEM_LEAF, g2.b});
                i += 2;
                continue;
            }
            out.ops_t4.push_back(table_op(g, cherry_table_at[i], op.a, op.b, -1, -1));
            if (g.kind == OP_GEMM_LEAF) out.items_t4.push_back({ITEM_LEAF, g.b});
            i++;
            continue;
        }
        out.ops_t4.push_back(op);
        plain_items(op, out.items_t4);
    }
    out.n_tab4 = (int)quads.size();
    // A push directly followed by its pop means the second subtree is a single lookup: nothing is parked, the first message
    // stays in the registers and the lookup multiplies into it (OP_..._KEEP + OP_TAB_MUL). Same factors, same product.
    // The producer's items do not change: a contraction still needs its P image, lookups never had items.
    auto keep_in_registers = [](std::vector<Op>& ops) {
        for (size_t i = 0; i + 1 < ops.size(); i++) {
            const int k0 = ops[i].kind & 0xff, k1 = ops[i + 1].kind & 0xff;
            if (k1 == OP_TAB_POP && (k0 == OP_GEMM_PUSH || k0 == OP_TAB_PUSH) && ops[i].c == ops[i + 1].c) {
                ops[i].kind = k0 == OP_GEMM_PUSH ? (int)OP_GEMM_KEEP : (OP_TAB_KEEP | (ops[i].kind & ~0xff));
                ops[i + 1].kind = OP_TAB_MUL | (ops[i + 1].kind & ~0xff);
            }
        }
    };
    if (keep) {
        keep_in_registers(out.ops_t);
        keep_in_registers(out.ops_t3);
        keep_in_registers(out.ops_t4);
    }
    out.subtabs.insert(out.subtabs.end(), triples.begin(), triples.end());
    out.subtabs.insert(out.subtabs.end(), quads.begin(), quads.end());
    out.tab_off.assign(out.subtabs.size(), 0);
    for (size_t k = 0; k < out.subtabs.size(); k++) {
        const long long k2 = std::min<long long>(k, out.n_tab2), k3 = std::min<long long>(std::max<long long>((long long)k - out.n_tab2, 0), out.n_tab3),
                        k4 = std::max<long long>((long long)k - out.n_tab2 - out.n_tab3, 0);
        out.tab_off[k] = k2 * CHERRY_TABLE + k3 * TRIPLE_TABLE + k4 * QUAD_TABLE;
        out.subtabs[k].off = out.tab_off[k];
    }
    return out;
}

}  // namespace pcsf
