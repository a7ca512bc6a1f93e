"""The tree programs the pruning kernels run (phylocsf_b200/csrc/pcsf_program.hpp, built by pcsf_tree_set), checked without
a GPU: pcsf_host_tree_program hands out the very op lists the device gets; a small interpreter executes them in numpy - the
plain post-order program and the table programs of levels 2-4, with and without the KEEP / MUL rewrite - and the root
likelihood must equal the oracle's pruning (lib/CamlPaml/PhyloLik.ml:73-93 restated in oracle/phylo_oracle.c) on random
trees, random transition matrices and random leaf codes (gaps included). Also structural invariants: stack discipline,
KEEP directly followed by MUL, every internal edge contracted or tabulated exactly once."""
import numpy as np
import pytest

from oracle import oracle as o
from phylocsf_b200 import host

OP_CHERRY, OP_GEMM_LEAF, OP_GEMM_PUSH, OP_GEMM_POP, OP_ROOT = 0, 1, 2, 3, 4
OP_TAB_LEAF, OP_TAB_PUSH, OP_TAB_POP, OP_TAB_KEEP, OP_TAB_MUL, OP_GEMM_KEEP = 5, 6, 7, 8, 9, 10
K = 8  # states: the programs do not depend on the alphabet size, and 8 keeps the test fast


def random_tree(n, rng, shape):
    """children [n-1, 2] in the T numbering (leaves 0..n-1, a child's id below its parent's, root 2n-2)"""
    roots = list(range(n))
    ch = []
    nxt = n
    while len(roots) > 1:
        if shape == "caterpillar":
            i, j = len(roots) - 1, 0  # the growing spine joins the next leaf
        elif shape == "balanced":
            i, j = 0, 1
        else:
            i, j = rng.choice(len(roots), size=2, replace=False)
        a, b = roots[i], roots[j]
        if rng.random() < 0.5:
            a, b = b, a
        for x in sorted((i, j), reverse=True):
            roots.pop(x)
        ch.append((a, b))
        if shape == "balanced":
            roots.append(nxt)
        else:
            roots.insert(int(rng.integers(0, len(roots) + 1)) if shape == "random" else len(roots), nxt)
        nxt += 1
    return np.array(ch, dtype=np.int32)


def leaf_message(P, code):
    """G(leaf)[x] = P[x][code], or the row sum for `Marginalize (PhyloLik.ml:68-71)"""
    return P.sum(axis=1) if code >= K else P[:, code]


def table_value(tabs, k, P, codes):
    """W of memoised subtree k for this column: P_edge x (G(a) * G(b)) for a cherry, P_edge x (W_src * G(new)) above it"""
    la, lb, lnew, edge, src = (int(x) for x in tabs[k])
    if src < 0:
        inner = leaf_message(P[la], codes[la]) * leaf_message(P[lb], codes[lb])
    else:
        inner = table_value(tabs, src, P, codes) * leaf_message(P[lnew], codes[lnew])
    return P[edge] @ inner


def table_leaves(tabs, k):
    la, lb, lnew, edge, src = (int(x) for x in tabs[k])
    return [la, lb] if src < 0 else table_leaves(tabs, src) + [lnew]


def interpret(ops, tabs, P, prior, codes):
    """runs one column through a program; returns (z, contracted edges, leaves read)"""
    cur, stack, edges, leaves = None, {}, [], []
    for kind, a, b, c in (tuple(int(x) for x in op) for op in ops):
        k, table = kind & 0xFF, kind >> 8
        if k in (OP_TAB_LEAF, OP_TAB_PUSH, OP_TAB_POP, OP_TAB_KEEP, OP_TAB_MUL):
            la, lb, lc, ld = a & 0xFFFF, (a >> 16) & 0xFFFF, b & 0xFFFF, (b >> 16) & 0xFFFF
            want = [la, lb] + ([lc] if lc != 0xFFFF else []) + ([ld] if ld != 0xFFFF else [])
            assert table_leaves(tabs, table) == want  # the op names the leaves whose codes index its table, in table order
            W = table_value(tabs, table, P, codes)
            leaves += want
            t = table
            while t >= 0:
                edges.append(int(tabs[t][3]))
                t = int(tabs[t][4])
            if k == OP_TAB_LEAF:
                cur = W * leaf_message(P[c], codes[c])
                leaves.append(c)
            elif k == OP_TAB_PUSH:
                assert c not in stack
                stack[c] = W
                cur = None
            elif k == OP_TAB_POP:
                cur = W * stack.pop(c)
            elif k == OP_TAB_KEEP:
                cur = W
            else:
                cur = cur * W
        elif k == OP_CHERRY:
            cur = leaf_message(P[a], codes[a]) * leaf_message(P[b], codes[b])
            leaves += [a, b]
        elif k == OP_ROOT:
            return float(cur @ prior), edges, leaves, stack
        else:
            acc = P[a] @ cur
            edges.append(a)
            if k == OP_GEMM_LEAF:
                cur = acc * leaf_message(P[b], codes[b])
                leaves.append(b)
            elif k == OP_GEMM_PUSH:
                assert c not in stack
                stack[c] = acc
                cur = None
            elif k == OP_GEMM_POP:
                cur = acc * stack.pop(c)
            else:
                assert k == OP_GEMM_KEEP
                cur = acc
    raise AssertionError("program without OP_ROOT")


@pytest.mark.parametrize("shape", ["random", "caterpillar", "balanced", "spiky"])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 12, 29, 58])
def test_every_program_variant_computes_the_reference_likelihood(n, shape):
    rng = np.random.default_rng(1000 * n + len(shape))
    for rep in range(3):
        ch = random_tree(n, rng, shape)
        P = rng.dirichlet(np.full(K, 0.3), size=(2 * n - 2, K))  # a stochastic matrix per branch
        prior = rng.dirichlet(np.ones(K))
        cols = rng.integers(0, K + 1, size=(4, n))  # code K = gap / missing: marginalised
        cols[0] = K  # a column of gaps only: z = 1
        # the oracle's C code takes the alphabet size from the matrices; any code >= k marginalises (PhyloLik.ml:11-19)
        want = [o.posteriors_column(n, ch, P, prior, [int(x) for x in c])[0] for c in cols]
        assert want[0] == pytest.approx(1.0, rel=1e-12)
        n_keep_rewrites = 0
        for level in (0, 2, 3, 4):
            for keep in (False, True):
                ops, tabs, (t2, t3, t4, levels) = host.tree_program(n, ch, level, keep)
                kinds = ops[:, 0] & 0xFF
                assert kinds[-1] == OP_ROOT and (kinds[:-1] != OP_ROOT).all()
                if not keep or level == 0:
                    assert not np.isin(kinds, (OP_TAB_KEEP, OP_TAB_MUL, OP_GEMM_KEEP)).any()
                if level == 0:
                    assert (kinds <= OP_ROOT).all()
                for i in np.flatnonzero(np.isin(kinds, (OP_TAB_KEEP, OP_GEMM_KEEP))):
                    assert kinds[i + 1] == OP_TAB_MUL  # what is kept is consumed by the very next op
                assert (kinds == OP_TAB_MUL).sum() == np.isin(kinds, (OP_TAB_KEEP, OP_GEMM_KEEP)).sum()
                n_keep_rewrites += int((kinds == OP_TAB_MUL).sum())
                pushes = np.isin(kinds, (OP_GEMM_PUSH, OP_TAB_PUSH))
                assert (ops[pushes, 3] < max(levels, 1)).all() and levels <= 16
                for c, w in zip(cols, want):
                    z, edges, leaves, stack = interpret(ops, tabs, P, prior, c)
                    assert not stack  # everything parked was un-parked
                    assert sorted(edges) == list(range(n, 2 * n - 2)), (level, keep)  # each internal edge once
                    assert sorted(leaves) == list(range(n))  # each leaf once
                    assert z == pytest.approx(w, rel=1e-12, abs=1e-300), (n, shape, level, keep)
                if level == 0:
                    assert len(tabs) == t2 + t3 + t4
        if n >= 12 and shape in ("random", "spiky"):
            assert n_keep_rewrites > 0  # the rewrite is exercised on trees of this size


def test_program_of_the_58mammals_tree_keeps_eleven_of_sixteen_messages_in_registers(params_base):
    """the numbers DESIGN section 3 quotes for the headline configuration"""
    import pcsf_helpers as H

    ps = H.oracle_paramset(params_base, "58mammals")
    n, ch = ps.tree.n_leaves, ps.tree.children_array()
    ops, tabs, (t2, t3, t4, levels) = host.tree_program(n, ch, 4, True)
    kinds = ops[:, 0] & 0xFF
    assert (t2, t3, t4) == (17, 9, 5)
    assert np.isin(kinds, (OP_GEMM_LEAF, OP_GEMM_PUSH, OP_GEMM_POP, OP_GEMM_KEEP)).sum() == 25  # of the 56 internal edges
    assert (kinds == OP_TAB_MUL).sum() == 11 and np.isin(kinds, (OP_GEMM_PUSH, OP_TAB_PUSH)).sum() == 5
    old, _, _ = host.tree_program(n, ch, 4, False)
    assert np.isin(old[:, 0] & 0xFF, (OP_GEMM_PUSH, OP_TAB_PUSH)).sum() == 16


def test_bad_trees_are_rejected():
    with pytest.raises(host.HostError):
        host.tree_program(3, [0, 3, 1, 2], 0)  # a node that is its own child
    with pytest.raises(host.HostError):
        host.tree_program(3, [0, 1, 0, 3], 0)  # a node with two parents
