"""
oracle.py — CPU restatement of PhyloCSF's scoring path (TEST INFRASTRUCTURE, not the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The product (phylocsf_b200) never does; it has no CPU fallback.

It restates the reference's OCaml host logic in numpy and calls oracle/liboracle.so
(phylo_oracle.c) for the two numeric loops (P(t) and pruning). Every function cites the reference
file:line it follows (paths relative to /root/reference). It is written independently of the
C++/CUDA product: eigendecomposition here is LAPACK's general non-symmetric solver
(numpy.linalg.eig, like GSL's gsl_eigen_nonsymmv at lib/CamlPaml/Q.ml:126) and the inverse is an LU
inverse (Q.ml:83-93), whereas the product symmetrises Q and runs its own Jacobi solver. (When LAPACK
splits a degenerate real eigenvalue of a reversible Q into a complex pair - the omega model's Q does
that - the oracle falls back to LAPACK's symmetric solver on the symmetrised matrix, QDiag._symmetric_path.)

Third-party pieces restated because they are not in /root/reference (GSL, version unpinned by the
reference; OCaml stdlib Random): gsl_min_fminimizer_brent + gsl_min_fminimizer_set (GSL min/brent.c,
min/fsolver.c — restated from the published algorithm), gsl_ran_gamma_pdf, cblas ddot/dgemm (in
phylo_oracle.c), OCaml 4.x Random.init/Random.float (lagged-Fibonacci generator).

Parity status: pinned against the reference's own known answers (tests/test_oracle_golden.py):
lib/CamlPaml/test.ml:8-54 (2-state, 8 patterns), test.ml:56-99 (JC69 lnL -1574.63623),
src/test.ml:27-59 (four end-to-end scores to +-0.005 dB with region coordinates). Brent's iterate
sequence and OCaml's Random stream are pinned only through those windows.
"""
from __future__ import annotations

import ctypes
import hashlib
import math
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
K = 64  # Codon64.dim, lib/CamlPaml/Code.ml:135
MARG = 64  # packed leaf code for `Marginalize


class OracleFailure(Exception):
    """Stands for the reference's Failure / Invalid_argument / Gsl_exn in the scoring path."""


# --------------------------------------------------------------------------------------------
# C library
# --------------------------------------------------------------------------------------------
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_real_to_Pt.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp]
        L.oracle_real_to_Pt.restype = ctypes.c_int
        L.oracle_ensure_alpha.argtypes = [ctypes.c_int, ctypes.c_void_p, dp, dp, ctypes.c_int, ctypes.c_void_p, dp]
        L.oracle_ensure_alpha.restype = ctypes.c_double
        L.oracle_lpr_leaves.argtypes = [ctypes.c_int, ctypes.c_void_p, dp, dp, ctypes.c_int, ctypes.c_int64,
                                        ctypes.c_void_p, dp, dp, dp, dp]
        L.oracle_lpr_leaves.restype = None
        L.oracle_lpr_batch.argtypes = [ctypes.c_int, ctypes.c_void_p, dp, ctypes.c_void_p, dp, ctypes.c_int,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, dp, dp, ctypes.c_int]
        L.oracle_lpr_batch.restype = None
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_posteriors.argtypes = [ctypes.c_int, ctypes.c_void_p, dp, dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_double, dp, dp]
        L.oracle_posteriors.restype = ctypes.c_double
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------------------------
# Newick  (lib/CamlPaml/NewickLexer.mll:5-14, NewickParser.mly:7-24, Newick.ml)
# --------------------------------------------------------------------------------------------
@dataclass
class Node:
    children: List["Node"]
    label: str
    bl: Optional[float]


_BL_CHARS = set("0123456789.")
_LBL_CHARS = set("ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_.")


def _newick_tokens(text: str):
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c in " \t\r\n;":  # NewickLexer.mll:6 (';' is skipped as whitespace)
            i += 1
        elif c in "(),:":
            yield (c, None)
            i += 1
        elif c in _LBL_CHARS:
            j = i
            while j < n and text[j] in _LBL_CHARS:
                j += 1
            lex = text[i:j]
            # longest match; on a tie the BRANCHLEN rule comes first (NewickLexer.mll:11-12)
            if all(ch in _BL_CHARS for ch in lex):
                yield ("BL", float(lex))
            else:
                yield ("LABEL", lex)
            i = j
        else:
            raise OracleFailure("newick: illegal character %r" % c)
    yield ("EOF", None)


def newick_parse(text: str) -> Node:
    toks = list(_newick_tokens(text))
    pos = [0]

    def peek():
        return toks[pos[0]][0]

    def take(kind):
        k, v = toks[pos[0]]
        if k != kind:
            raise OracleFailure("newick: parse error at token %d (%s, wanted %s)" % (pos[0], k, kind))
        pos[0] += 1
        return v

    def label():  # NewickParser.mly:16-20
        if peek() == "LABEL":
            lbl = take("LABEL")
            if peek() == ":":
                take(":")
                return lbl, take("BL")
            return lbl, None
        take(":")
        return "", take("BL")

    def node():  # NewickParser.mly:11-15
        if peek() == "(":
            take("(")
            sub = [node()]
            while peek() == ",":
                take(",")
                sub.append(node())
            take(")")
            if peek() in ("LABEL", ":"):
                lbl, bl = label()
                return Node(sub, lbl, bl)
            return Node(sub, "", None)
        lbl, bl = label()
        return Node([], lbl, bl)

    root = node()
    take("EOF")
    return root


def newick_size(nd: Node) -> int:  # Newick.ml:19
    return 1 + sum(newick_size(c) for c in nd.children)


def newick_leaves(nd: Node) -> int:  # Newick.ml:21-23
    return 1 if not nd.children else sum(newick_leaves(c) for c in nd.children)


def newick_subtree(keep: Callable[[str], bool], nd: Node) -> Optional[Node]:  # Newick.ml:34-41
    if not nd.children:
        return nd if (nd.label == "" or keep(nd.label)) else None
    if nd.label == "" or keep(nd.label):
        st = [s for s in (newick_subtree(keep, c) for c in nd.children) if s is not None]
        if not st:
            return None
        if len(st) == 1:
            s = st[0]
            bl = (nd.bl + s.bl) if (nd.bl is not None and s.bl is not None) else None  # maybe_add
            return Node(s.children, s.label, bl)
        return Node(st, nd.label, nd.bl)
    return None


def newick_total_length(nd: Node, count_root: bool = False) -> float:  # Newick.ml:52-61
    def tl(x: Node) -> float:
        if x.bl is None:
            raise OracleFailure("CamlPaml.Newick.total_length: unspecified branch length")
        acc = 0.0  # fold_left (+.) 0. (map total_length' st)
        for c in x.children:
            acc = acc + tl(c)
        return x.bl + acc

    if count_root:
        return tl(nd)
    return tl(Node(nd.children, nd.label, 0.0))


# --------------------------------------------------------------------------------------------
# T  (lib/CamlPaml/T.ml:57-112): leaves 0..n-1 left to right, internals post-order, root last
# --------------------------------------------------------------------------------------------
@dataclass
class Tree:
    labels: List[str]
    parents: List[int]
    children: List[Tuple[int, int]]  # indexed by node id; (-1,-1) for leaves
    branches: List[float]

    @property
    def size(self) -> int:
        return len(self.parents)

    @property
    def n_leaves(self) -> int:
        return (self.size + 1) // 2

    @property
    def root(self) -> int:
        return self.size - 1

    def children_array(self) -> np.ndarray:
        """int32 [(n-1)*2]: children of internal nodes n..2n-2 (the C-ABI / oracle C layout)."""
        nl = self.n_leaves
        return np.array([c for i in range(nl, self.size) for c in self.children[i]], dtype=np.int32)

    @staticmethod
    def of_newick(nt: Node) -> "Tree":
        n = newick_size(nt)
        if n < 3 or n % 2 == 0:
            raise OracleFailure("CamlPaml.T.of_newick: input is not a rooted, bifurcating tree")
        leaves: List[Node] = []

        def find_leaves(x: Node):
            if not x.children:
                leaves.append(x)
            elif len(x.children) == 2:
                find_leaves(x.children[0])
                find_leaves(x.children[1])
            else:
                raise OracleFailure("CamlPaml.T.of_newick: input is not a rooted, bifurcating tree")

        find_leaves(nt)
        nl = len(leaves)
        assert nl == (n + 1) // 2
        leaf_id = {id(x): i for i, x in enumerate(leaves)}
        labels = [""] * n
        parents = [-1] * n
        children = [(-1, -1)] * n
        branches = [float("nan")] * n
        nxt = [nl - 1]

        def fill(x: Node) -> int:
            if not x.children:
                i = leaf_id[id(x)]
            else:
                lc = fill(x.children[0])
                rc = fill(x.children[1])
                nxt[0] += 1
                i = nxt[0]
                parents[lc] = i
                parents[rc] = i
                children[i] = (lc, rc)
            labels[i] = x.label
            branches[i] = x.bl if x.bl is not None else float("nan")
            return i

        r = fill(nt)
        assert r == n - 1
        return Tree(labels, parents, children, branches)


# --------------------------------------------------------------------------------------------
# Code  (lib/CamlPaml/Code.ml:11-58, 133-181)
# --------------------------------------------------------------------------------------------
_DNA_INDEX = {"A": 0, "a": 0, "C": 1, "c": 1, "G": 2, "g": 2, "T": 3, "t": 3}
_COMP = {"A": "T", "G": "C", "C": "G", "T": "A", "a": "t", "g": "c", "c": "g", "t": "a", "N": "N", "n": "n", "-": "-"}
TRANSLATION = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"  # Code.ml:159-178
STOPS = (48, 50, 56)  # TAA TAG TGA, Code.ml:149-151


def revcomp(s: str) -> str:  # Code.ml:39-58; raises on anything but ACGTacgtNn-
    try:
        return "".join(_COMP[c] for c in reversed(s))
    except KeyError as e:
        raise OracleFailure("Invalid_argument(\"unrecognized nucleotide %s\")" % e.args[0])


def codon_code(c1: str, c2: str, c3: str) -> int:
    """`Certain (16 i1 + 4 i2 + i3) or MARG (src/PhyloCSF.ml:233-240, Code.ml:138-141)."""
    try:
        return 16 * _DNA_INDEX[c1] + 4 * _DNA_INDEX[c2] + _DNA_INDEX[c3]
    except KeyError:
        return MARG


def codon_of_index(i: int) -> str:
    return "ACGT"[i // 16] + "ACGT"[(i // 4) % 4] + "ACGT"[i % 4]


def translate(dna: str) -> str:  # src/PhyloCSF.ml:268-278
    out = []
    for i in range(len(dna) // 3):
        c = codon_code(dna[3 * i], dna[3 * i + 1], dna[3 * i + 2])
        out.append(TRANSLATION[c] if c != MARG else "?")
    return "".join(out)


# --------------------------------------------------------------------------------------------
# ECM files  (src/ECM.ml:18-72)
# --------------------------------------------------------------------------------------------
def read_ecm(path: str) -> Tuple[np.ndarray, np.ndarray]:
    lines = open(path).read().split("\n")
    raw = [[float(x) for x in ln.split(" ") if x.strip() != ""] for ln in lines[: K - 1]]
    s = np.zeros((K, K))
    for i in range(K):
        for j in range(K):
            if i > j:
                s[i, j] = raw[i - 1][j]
            elif i < j:
                s[i, j] = raw[j - 1][i]
    if lines[K - 1].strip() != "":
        raise OracleFailure("ECM.import_parameters")
    pi = np.array([float(x) for x in lines[K].split(" ") if x.strip() != ""])
    codons = [x.strip() for i in (3, 4, 5, 6) for x in lines[K + i].split(" ") if x.strip() != ""]
    if len(codons) != K or any(codon_code(*c) != i for i, c in enumerate(codons)):
        raise OracleFailure("ECM.import_parameters: incorrect codon order")
    return s, pi


# --------------------------------------------------------------------------------------------
# Q  (lib/CamlPaml/Q.ml)
# --------------------------------------------------------------------------------------------
def _fill_diag_and_scale(q: np.ndarray, pi: np.ndarray, skip_zero_diag: bool) -> np.ndarray:
    """fill_q_diagonal (PhyloModel.ml:76-84) + qscale (PhyloCSFModel.ml:24-30 / OmegaModel.ml:76-80)
    + division by the scale (PhyloModel.ml:94-104). Accumulation orders as the Expr trees evaluate."""
    k = q.shape[0]
    q = q.copy()
    for i in range(k):
        tot = 0.0
        for j in range(k):
            if i != j:
                tot = q[i, j] + tot
        q[i, i] = 0.0 - tot
    factor = 0.0
    for i in range(k):
        if skip_zero_diag and q[i, i] == 0.0:
            continue
        factor = factor - pi[i] * q[i, i]
    if factor <= 0.0:
        raise OracleFailure("CamlPaml.P14n.instantiate_q: Q scale evaluated to a non-positive value")
    return q / factor


def ecm_q(s: np.ndarray, pi: np.ndarray) -> np.ndarray:
    """q_ij = s_ij * pi_j, diagonal, unit-rate scale (src/PhyloCSFModel.ml:11-30)."""
    q = s * pi[None, :]
    np.fill_diagonal(q, 0.0)
    return _fill_diag_and_scale(q, pi, skip_zero_diag=True)


def _check_real(z: complex, tol: float = 1e-6) -> bool:  # Q.ml:20
    return z.imag == 0.0 or abs(z.imag) * 1000.0 < abs(z) or (abs(z.real) < tol and abs(z.imag) < tol)


class QDiag:
    """Q.Diag.t, real (reversible) path only (Q.ml:96-141)."""

    def __init__(self, qm: np.ndarray, tol: float = 1e-6):
        self.q = np.ascontiguousarray(qm, dtype=np.float64)
        self.tol = tol
        lam, s = np.linalg.eig(self.q)  # general non-symmetric, like gsl_eigen_nonsymmv (Q.ml:126)
        if np.iscomplexobj(lam) and np.any(lam.imag != 0.0):
            # LAPACK split a (near-)degenerate real eigenvalue of a reversible Q into a conjugate pair whose
            # eigenvectors are genuinely complex; the reference's real_of_complex would reject them. For a
            # reversible Q the spectrum is real: redo it on the symmetrised matrix with LAPACK's symmetric
            # solver (dsyevd - still independent of the product's Jacobi code). The omega model's Q has such
            # degeneracies; the ECMs do not.
            lam, s, sinv = self._symmetric_path()
        else:
            sinv = np.linalg.inv(s)  # LU inverse, like zinvm (Q.ml:83-93)
        if not all(_check_real(complex(z), tol) for z in lam):
            raise OracleFailure("oracle: non-reversible model (complex eigenvalues) is out of scope")
        for m in (s, sinv):
            if np.iscomplexobj(m) and not all(_check_real(complex(z), tol) for z in m.ravel()):
                raise OracleFailure("CamlPaml.Q.real_of_complex")
        self.S = np.ascontiguousarray(np.real(s), dtype=np.float64)
        self.Sinv = np.ascontiguousarray(np.real(sinv), dtype=np.float64)
        self.lam = np.ascontiguousarray(np.real(lam), dtype=np.float64)
        self._pi = None
        self._memo = {}

    def _symmetric_path(self):
        k = self.q.shape[0]
        u_, sv, vt = np.linalg.svd(self.q.T)  # stationary weights = null vector of Q^T
        w = np.abs(vt[-1])
        w = w / w.sum()
        sw = np.sqrt(w)
        a = sw[:, None] * self.q / sw[None, :]
        if np.abs(a - a.T).max() > 1e-9 * max(1.0, np.abs(a).max()):
            raise OracleFailure("oracle: non-reversible model (complex eigenvalues) is out of scope")
        lam, u = np.linalg.eigh(0.5 * (a + a.T))
        return lam, u / sw[:, None], u.T * sw[None, :]

    def scaled(self, x: float) -> "QDiag":  # Q.ml:193-209
        if x <= 0.0:
            raise OracleFailure("CamlPaml.Q.scale: nonpositive scale factor")
        o = QDiag.__new__(QDiag)
        o.q, o.tol, o.S, o.Sinv, o.lam = self.q * x, self.tol, self.S, self.Sinv, self.lam * x
        o._pi, o._memo = None, {}
        return o

    def equilibrium(self) -> np.ndarray:  # Q.ml:153-177
        if self._pi is None:
            mags = np.abs(self.lam)
            p = int(np.argmin(mags))  # first minimum, as the strict '<' scan finds
            if mags[p] > self.tol:
                raise OracleFailure("CamlPaml.Q.equilibrium: smallest-magnitude eigenvalue %e is unacceptably large" % mags[p])
            lev = self.Sinv[p, :]
            mass = 0.0
            for v in lev:
                mass += v
            self._pi = np.array([v / mass for v in lev])
        return self._pi.copy()

    def to_Pt(self, t: float) -> np.ndarray:  # Q.ml:211-256
        key = float(t)
        if key not in self._memo:
            k = self.q.shape[0]
            P = np.empty((k, k))
            st = lib().oracle_real_to_Pt(k, _dp(self.S), _dp(self.Sinv), _dp(self.lam), key, self.tol, _dp(P))
            if st != 0:
                raise OracleFailure("CamlPaml.Q.real_to_Pt status %d at t=%g" % (st, t))
            self._memo[key] = P
        return self._memo[key]


# --------------------------------------------------------------------------------------------
# PhyloModel / PhyloLik  (lib/CamlPaml/PhyloModel.ml:12-34, PhyloLik.ml)
# --------------------------------------------------------------------------------------------
class PhyloModel:
    def __init__(self, tree: Tree, q: QDiag, branches: Sequence[float], prior: Optional[np.ndarray] = None):
        self.tree = tree
        self.q = q
        self.branches = list(branches)
        for b in self.branches[: tree.root]:
            if b < 0.0:
                raise OracleFailure("CamlPaml.PhyloModel.make")
        self.pms = np.ascontiguousarray(np.stack([q.to_Pt(b) for b in self.branches[: tree.root]]))  # PhyloModel.ml:17
        self._prior = None if prior is None else np.array(prior, dtype=np.float64)

    def prior(self) -> np.ndarray:  # PhyloModel.ml:30-32
        return self._prior.copy() if self._prior is not None else self.q.equilibrium()


def likelihood_column(model: PhyloModel, codes: Sequence[int]):
    """PhyloLik.prepare + likelihood for one column; returns (z, alpha[n_internal,k])."""
    t = model.tree
    k = model.pms.shape[1]
    ch = t.children_array()
    pr = np.ascontiguousarray(model.prior())
    c = np.ascontiguousarray(np.array(codes, dtype=np.uint8))
    alpha = np.empty((t.n_leaves - 1, k))
    z = lib().oracle_ensure_alpha(t.n_leaves, ch.ctypes.data, _dp(model.pms), _dp(pr), k, c.ctypes.data, _dp(alpha))
    return z, alpha


def posteriors_column(n_leaves: int, children: np.ndarray, pms: np.ndarray, prior: np.ndarray, codes: Sequence[int],
                      weight: float = 1.0, ecounts: Optional[np.ndarray] = None):
    """PhyloLik.node_posterior for every node and (accumulating into `ecounts` when given) add_branch_posteriors for
    every branch of one column (PhyloLik.ml:96-180). Returns (z, node_post [2n-1, k])."""
    k = pms.shape[1]
    ch = np.ascontiguousarray(children, dtype=np.int32)
    pms = np.ascontiguousarray(pms, dtype=np.float64)
    pr = np.ascontiguousarray(prior, dtype=np.float64)
    c = np.ascontiguousarray(np.array(codes, dtype=np.uint8))
    post = np.empty((2 * n_leaves - 1, k))
    if ecounts is not None:
        assert ecounts.shape == (2 * n_leaves - 2, k, k) and ecounts.flags["C_CONTIGUOUS"] and ecounts.dtype == np.float64
    z = lib().oracle_posteriors(n_leaves, ch.ctypes.data, _dp(pms), _dp(pr), k, c.ctypes.data, weight, _dp(post),
                                _dp(ecounts) if ecounts is not None else None)
    return z, post


def posteriors_columns(model: PhyloModel, codes: np.ndarray):
    """All columns of a region: (z [ncols], node_post [ncols, 2n-1, k], ecounts [2n-2, k, k] summed over the columns in
    column order with weight 1 - what PhyloEM's E step accumulates)."""
    t = model.tree
    k = model.pms.shape[1]
    ch = t.children_array()
    pr = model.prior()
    ecounts = np.zeros((2 * t.n_leaves - 2, k, k))
    zs, posts = [], []
    for c in np.ascontiguousarray(codes, dtype=np.uint8):
        z, post = posteriors_column(t.n_leaves, ch, model.pms, pr, c, 1.0, ecounts)
        zs.append(z)
        posts.append(post)
    return np.array(zs), np.array(posts), ecounts


def lpr_columns(model: PhyloModel, codes: np.ndarray):
    """The per-column loop of src/PhyloCSFModel.ml:72-82. codes: uint8 [ncols][n_leaves].
    Returns (lpr, elpr_anc, col_logz, col_anc)."""
    t = model.tree
    k = model.pms.shape[1]
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    ncols = codes.shape[0]
    assert codes.shape[1] == t.n_leaves
    ch = t.children_array()
    pr = np.ascontiguousarray(model.prior())
    lpr = ctypes.c_double()
    elpr = ctypes.c_double()
    clz = np.empty(ncols)
    can = np.empty(ncols)
    lib().oracle_lpr_leaves(t.n_leaves, ch.ctypes.data, _dp(model.pms), _dp(pr), k, ncols, codes.ctypes.data,
                            ctypes.cast(ctypes.byref(lpr), ctypes.POINTER(ctypes.c_double)),
                            ctypes.cast(ctypes.byref(elpr), ctypes.POINTER(ctypes.c_double)), _dp(clz), _dp(can))
    return lpr.value, elpr.value, clz, can


# --------------------------------------------------------------------------------------------
# Fit  (lib/CamlPaml/Fit.ml:3-48) + GSL Brent + OCaml Random
# --------------------------------------------------------------------------------------------
class OCamlRandom:
    """OCaml 4.x Stdlib.Random (lagged Fibonacci, 55 words): Random.init / Random.float.
    Used only by Fit.find_init's rarely taken random branch (Fit.ml:33-41). The reference pins no
    OCaml version (5.x switched to LXM), so this stream is parity-unpinned."""

    def __init__(self, seed: int = 0):
        st = list(range(55))
        accu = b"x"
        for i in range(55 + 55):
            j = i % 55
            accu = hashlib.md5(accu + str(seed).encode()).digest()
            ext = accu[0] + (accu[1] << 8) + (accu[2] << 16) + (accu[3] << 24)
            st[j] = (st[j] ^ ext) & 0x3FFFFFFF
        self.st = st
        self.idx = 0

    def bits(self) -> int:
        self.idx = (self.idx + 1) % 55
        cur = self.st[self.idx]
        new = self.st[(self.idx + 24) % 55] + (cur ^ ((cur >> 25) & 0x1F))
        new30 = new & 0x3FFFFFFF
        self.st[self.idx] = new30
        return new30

    def rawfloat(self) -> float:
        scale = 1073741824.0
        r1 = float(self.bits())
        r2 = float(self.bits())
        return (r1 / scale + r2) / scale

    def float(self, bound: float) -> float:
        return self.rawfloat() * bound


def find_init(f: Callable[[float], float], init: float, lo: float, hi: float, maxtries: int = 1000,
              logspace: bool = False, trace: Optional[dict] = None) -> float:  # Fit.ml:27-48
    if lo >= hi or (logspace and lo <= 0.0):
        raise OracleFailure("CamlPaml.Fit.find_init")
    width = (math.log(hi) - math.log(lo)) if logspace else (hi - lo)
    flo = f(lo)
    fhi = f(hi)
    x = init
    fx = f(init)
    i = 0
    rng = OCamlRandom(0)
    while i < maxtries and (fx <= flo or fx <= fhi):
        if logspace:
            x = math.exp(math.log(lo) + rng.float(width))
        else:
            x = lo + rng.float(width)
        fx = f(x)
        i += 1
    if trace is not None:
        trace["random_tries"] = trace.get("random_tries", 0) + i
    if i == maxtries:
        x = lo if flo > fhi else hi
    return x


_GOLDEN = 0.3819660
_SQRT_DBL_EPSILON = 1.4901161193847656e-08


class BrentMinimizer:
    """gsl_min_fminimizer (brent) as driven through Gsl.Min.make/iterate/minimum/interval
    (Fit.ml:13-19). Restated from GSL min/fsolver.c + min/brent.c (GSL is not in the image)."""

    def __init__(self, f: Callable[[float], float], x_minimum: float, x_lower: float, x_upper: float):
        self.f = f
        # gsl_min_fminimizer_set: compute_f_values then range checks
        f_lower = self._call(x_lower)
        f_upper = self._call(x_upper)
        f_minimum = self._call(x_minimum)
        if x_lower > x_upper:
            raise OracleFailure("Gsl_exn: invalid interval (lower > upper)")
        if x_minimum >= x_upper or x_minimum <= x_lower:
            raise OracleFailure("Gsl_exn: x_minimum must lie inside interval (lower < x < upper)")
        if f_minimum >= f_lower or f_minimum >= f_upper:
            raise OracleFailure("Gsl_exn: endpoints do not enclose a minimum")
        self.x_minimum, self.f_minimum = x_minimum, f_minimum
        self.x_lower, self.f_lower = x_lower, f_lower
        self.x_upper, self.f_upper = x_upper, f_upper
        # brent_init
        v = x_lower + _GOLDEN * (x_upper - x_lower)
        self.v = v
        self.w = v
        self.d = 0.0
        self.e = 0.0
        f_vw = self._call(v)
        self.f_v = f_vw
        self.f_w = f_vw

    def _call(self, x: float) -> float:  # SAFE_FUNC_CALL
        y = self.f(x)
        if not math.isfinite(y):
            raise OracleFailure("Gsl_exn: computed function value is infinite or NaN")
        return y

    def iterate(self):
        x_left, x_right = self.x_lower, self.x_upper
        z = self.x_minimum
        d = self.e  # sic: GSL loads d from state->e and e from state->d
        e = self.d
        v, w, f_v, f_w, f_z = self.v, self.w, self.f_v, self.f_w, self.f_minimum
        w_lower = z - x_left
        w_upper = x_right - z
        tolerance = _SQRT_DBL_EPSILON * abs(z)
        p = q = r = 0.0
        midpoint = 0.5 * (x_left + x_right)
        if abs(e) > tolerance:
            r = (z - w) * (f_z - f_v)
            q = (z - v) * (f_z - f_w)
            p = (z - v) * q - (z - w) * r
            q = 2 * (q - r)
            if q > 0:
                p = -p
            else:
                q = -q
            r = e
            e = d
        if abs(p) < abs(0.5 * q * r) and p < q * w_lower and p < q * w_upper:
            t2 = 2 * tolerance
            d = p / q
            u = z + d
            if (u - x_left) < t2 or (x_right - u) < t2:
                d = tolerance if z < midpoint else -tolerance
        else:
            e = (x_right - z) if z < midpoint else -(z - x_left)
            d = _GOLDEN * e
        if abs(d) >= tolerance:
            u = z + d
        else:
            u = z + (tolerance if d > 0 else -tolerance)
        self.e = e
        self.d = d
        f_u = self._call(u)
        if f_u <= f_z:
            if u < z:
                self.x_upper, self.f_upper = z, f_z
            else:
                self.x_lower, self.f_lower = z, f_z
            self.v, self.f_v = w, f_w
            self.w, self.f_w = z, f_z
            self.x_minimum, self.f_minimum = u, f_u
        else:
            if u < z:
                self.x_lower, self.f_lower = u, f_u
            else:
                self.x_upper, self.f_upper = u, f_u
            if f_u <= f_w or w == z:
                self.v, self.f_v = w, f_w
                self.w, self.f_w = u, f_u
            elif f_u <= f_v or v == z or v == w:
                self.v, self.f_v = u, f_u


def maximize_lpr(f: Callable[[float], object], g: Callable[[object], float], init: float = 1.0, lo: float = 1e-2,
                 hi: float = 10.0, accuracy: float = 0.01, trace: Optional[dict] = None):
    """src/PhyloCSFModel.ml:84-99. Returns (x, f x)."""
    good_init = find_init(lambda x: g(f(x)), init, lo, hi, maxtries=250, logspace=True, trace=trace)
    if lo < good_init < hi:
        m = BrentMinimizer(lambda x: 0.0 - g(f(x)), good_init, lo, hi)  # Fit.ml:5-13
        go = True
        iters = 0
        while go:
            m.iterate()
            iters += 1
            x = m.x_minimum
            go = ((m.x_upper - m.x_lower) / x) > accuracy
        if trace is not None:
            trace["iterations"] = trace.get("iterations", 0) + iters
        x = m.x_minimum
        return x, f(x)
    return good_init, f(good_init)


# --------------------------------------------------------------------------------------------
# PhyloCSFModel  (src/PhyloCSFModel.ml)
# --------------------------------------------------------------------------------------------
def db(x: float) -> float:  # PhyloCSFModel.ml:113
    return 10.0 * x / math.log(10.0)


@dataclass
class Score:
    score: float
    anc_comp_score: float
    diagnostics: List[Tuple[str, str]] = field(default_factory=list)


class CodonInstance:
    """PM.P14n.instance for one ECM (new_instance, PhyloCSFModel.ml:52-56)."""

    def __init__(self, s: np.ndarray, pi: np.ndarray, tree_shape: Tree):
        self.q = QDiag(ecm_q(s, pi))
        self.tree_shape = tree_shape
        self.file_pi = pi

    def model(self, tree_scale: float) -> PhyloModel:
        # instantiate_tree: branch = Mul (Var 0, Val b) (PhyloCSFModel.ml:33, PhyloModel.ml:86-92);
        # P14n.update ~tree_settings rebuilds with prior=None => equilibrium prior (PhyloModel.ml:132-146)
        if not tree_scale > 0.0:
            raise OracleFailure("CamlPaml.P14n.instantiate_tree: domain violation on variable 0")
        br = [tree_scale * b for b in self.tree_shape.branches]
        return PhyloModel(self.tree_shape, self.q, br, prior=None)


def lpr_leaves(inst: CodonInstance, codes: np.ndarray, t: float):
    """src/PhyloCSFModel.ml:67-82 -> (lpr_leaves, elpr_anc)."""
    m = inst.model(t)
    lpr, elpr, _, _ = lpr_columns(m, codes)
    return lpr, elpr


class PhyloCSFModel:
    def __init__(self, s1, pi1, s2, pi2, tree_shape: Tree):  # PhyloCSFModel.ml:107-110
        self.coding_model = CodonInstance(s1, pi1, tree_shape)
        self.noncoding_model = CodonInstance(s2, pi2, tree_shape)
        self.tree = tree_shape

    def llr_fixed(self, t: float, codes: np.ndarray) -> Score:  # PhyloCSFModel.ml:122-128
        lpr_c, anc_c = lpr_leaves(self.coding_model, codes, t)
        lpr_n, anc_n = lpr_leaves(self.noncoding_model, codes, t)
        diag = [("rho", "%.2f" % t), ("L(C)", "%.2f" % db(lpr_c)), ("L(NC)", "%.2f" % db(lpr_n))]
        return Score(db(lpr_c - lpr_n), db(anc_c - anc_n), diag)

    def llr_maxlik(self, codes: np.ndarray, init: float = 1.0, trace: Optional[dict] = None) -> Score:  # :130-136
        rho_c, (lpr_c, anc_c) = maximize_lpr(lambda x: lpr_leaves(self.coding_model, codes, x), lambda r: r[0], init=init, trace=trace)
        rho_n, (lpr_n, anc_n) = maximize_lpr(lambda x: lpr_leaves(self.noncoding_model, codes, x), lambda r: r[0], init=init, trace=trace)
        diag = [("rho_0", "%.2f" % init), ("rho_C", "%.2f" % rho_c), ("rho_N", "%.2f" % rho_n),
                ("L(C)", "%.2f" % db(lpr_c)), ("L(NC)", "%.2f" % db(lpr_n))]
        if trace is not None:
            trace["rho_c"], trace["rho_n"] = rho_c, rho_n
        return Score(db(lpr_c - lpr_n), db(anc_c - anc_n), diag)

    def score(self, strategy: str, codes: np.ndarray, trace: Optional[dict] = None) -> Score:  # :142-146
        if strategy == "fixed":
            return self.llr_fixed(1.0, codes)
        if strategy == "mle":
            return self.llr_maxlik(codes, 1.0, trace)
        raise ValueError(strategy)


# --------------------------------------------------------------------------------------------
# OmegaModel  (src/OmegaModel.ml)
# --------------------------------------------------------------------------------------------
def _omega_pi(v: Sequence[float]) -> np.ndarray:  # OmegaModel.ml:24-42
    sigma = v[2]

    def sc(i1, i2, i3):
        f1 = (1.0 if i1 == 3 else v[3 + i1]) / (v[3] + (v[4] + (v[5] + 1.0)))
        f2 = (1.0 if i2 == 3 else v[6 + i2]) / (v[6] + (v[7] + (v[8] + 1.0)))
        f3 = (1.0 if i3 == 3 else v[9 + i3]) / (v[9] + (v[10] + (v[11] + 1.0)))
        return f1 * (f2 * f3)

    denom = 1.0 - (1.0 - sigma) * (sc(3, 0, 0) + (sc(3, 0, 2) + sc(3, 2, 0)))
    return np.array([sc(i // 16, (i // 4) % 4, i % 4) / denom for i in range(K)])


_TRANSITIONS = {(0, 2), (2, 0), (1, 3), (3, 1)}  # A<->G, C<->T


def omega_q(v: Sequence[float]) -> np.ndarray:
    """q_p14n + q_scale evaluated at settings v (OmegaModel.ml:44-80, PhyloModel.ml:94-106)."""
    kappa, omega = v[0], v[1]
    pi = _omega_pi(v)
    q = np.zeros((K, K))
    for i in range(K):
        ii = (i // 16, (i // 4) % 4, i % 4)
        for j in range(K):
            jj = (j // 16, (j // 4) % 4, j % 4)
            diffs = [(a, b) for a, b in zip(ii, jj) if a != b]
            if len(diffs) == 1:
                kp = kappa if diffs[0] in _TRANSITIONS else 1.0
                op = omega if (i not in STOPS and j not in STOPS and TRANSLATION[i] != TRANSLATION[j]) else 1.0
                q[i, j] = pi[j] * (kp * op)
    return _fill_diag_and_scale(q, pi, skip_zero_diag=False)


def gsl_ran_gamma_pdf(x: float, a: float, b: float) -> float:
    if x < 0:
        return 0.0
    if x == 0:
        return 1.0 / b if a == 1 else 0.0
    if a == 1:
        return math.exp(-x / b) / b
    return math.exp((a - 1) * math.log(x / b) - x / b - math.lgamma(a)) / b


def _cauchy_cdf(scale, x):
    return math.atan(x / scale) / math.acos(-1.0) + 0.5


def half_cauchy_lpdf(x: float, mode: float, scale: float) -> float:  # OmegaModel.ml:149-154
    if x < 0.0 or scale <= 0.0 or mode < 0.0:
        raise OracleFailure("half_cauchy_lpdf")
    pi_ = math.acos(-1.0)
    numer = 1.0 / (pi_ * scale * (1.0 + ((x - mode) / scale) ** 2.0))
    denom = 1.0 - _cauchy_cdf(scale, 0.0 - mode)
    return math.log(numer) - math.log(denom)


def lpr_rho(rho: float) -> float:  # OmegaModel.ml:156
    return half_cauchy_lpdf(rho, mode=1.0, scale=0.5)


def _log(x: float) -> float:
    return math.log(x) if x > 0 else (float("-inf") if x == 0 else float("nan"))


def lpr_kappa(k: float) -> float:  # OmegaModel.ml:157
    return _log(gsl_ran_gamma_pdf(k - 1.0 + np.finfo(float).eps, 7.0, 0.25))


class OmegaInstance:
    def __init__(self, tree_shape: Tree, q_settings: Sequence[float], tree_scale: float, q: Optional[QDiag] = None):
        self.tree_shape = tree_shape
        self.q_settings = list(q_settings)
        self.tree_scale = tree_scale
        for x in self.q_settings:
            if not x >= 0.0:
                raise OracleFailure("CamlPaml.P14n.instantiate_q: domain violation")
        self.q = q if q is not None else QDiag(omega_q(self.q_settings))
        if not tree_scale > 0.0:
            raise OracleFailure("CamlPaml.P14n.instantiate_tree: domain violation on variable 0")
        self.model = PhyloModel(tree_shape, self.q, [tree_scale * b for b in tree_shape.branches], prior=None)

    def with_rho(self, rho: float) -> "OmegaInstance":
        return OmegaInstance(self.tree_shape, self.q_settings, rho, q=self.q)

    def with_q(self, qs: Sequence[float]) -> "OmegaInstance":
        return OmegaInstance(self.tree_shape, qs, self.tree_scale)


def omega_update_f3x4(inst: OmegaInstance, codes: np.ndarray) -> OmegaInstance:  # OmegaModel.ml:102-134
    counts = [[1] * 4 for _ in range(3)]
    flat = codes.ravel()
    for c in flat[flat < K]:
        c = int(c)
        counts[0][c // 16] += 1
        counts[1][(c // 4) % 4] += 1
        counts[2][c % 4] += 1
    qs = list(inst.q_settings)
    for p in range(3):
        for n in range(3):
            qs[3 + 3 * p + n] = float(counts[p][n]) / float(counts[p][3])
    return inst.with_q(qs)


def omega_lpr_leaves(inst: OmegaInstance, codes: np.ndarray) -> float:  # OmegaModel.ml:137-144
    return lpr_columns(inst.model, codes)[0]


def omega_kr_map(codes: np.ndarray, inst: OmegaInstance, trace: Optional[dict] = None):  # OmegaModel.ml:160-190
    def f_rho(i: OmegaInstance, rho: float):
        ir = i.with_rho(rho)
        return (lpr_rho(rho) + omega_lpr_leaves(ir, codes)), ir

    def f_kappa(i: OmegaInstance, kappa: float):
        qs = list(i.q_settings)
        qs[0] = kappa
        ik = i.with_q(qs)
        return (lpr_kappa(kappa) + omega_lpr_leaves(ik, codes)), ik

    def rnd(i: OmegaInstance):
        _, (_, inst_rho) = maximize_lpr(lambda x: f_rho(i, x), lambda r: r[0], init=i.tree_scale, lo=0.001, hi=10.0, accuracy=0.01, trace=trace)
        _, (lpr, inst_kappa) = maximize_lpr(lambda x: f_kappa(inst_rho, x), lambda r: r[0], init=inst_rho.q_settings[0], lo=1.0, hi=10.0, accuracy=0.01, trace=trace)
        return inst_kappa, lpr

    return rnd(rnd(rnd(inst)[0])[0])


def omega_score(tree_shape: Tree, codes: np.ndarray, omega_H1: float = 0.2, sigma_H1: float = 0.01,
                trace: Optional[dict] = None) -> Score:  # OmegaModel.ml:195-219
    i0 = OmegaInstance(tree_shape, [2.5, 1.0, 1.0] + [1.0] * 9, 1.0)
    inst0, lpr_H0 = omega_kr_map(codes, omega_update_f3x4(i0, codes), trace)
    qs = list(inst0.q_settings)
    qs[1] = omega_H1
    qs[2] = sigma_H1
    inst1, lpr_H1 = omega_kr_map(codes, inst0.with_q(qs), trace)
    q0, q1 = inst0.q_settings, inst1.q_settings
    diag = [("L(H0)", "%.2f" % db(lpr_H0)), ("rho_H0", "%.2f" % inst0.tree_scale), ("kappa_H0", "%.2f" % q0[0]),
            ("omega_H0", "%.2f" % q0[1]), ("sigma_H0", "%.2f" % q0[2]), ("L(H1)", "%.2f" % db(lpr_H1)),
            ("rho_H1", "%.2f" % inst1.tree_scale), ("kappa_H1", "%.2f" % q1[0]), ("omega_H1", "%.2f" % q1[1]),
            ("sigma_H1", "%.2f" % q1[2])]
    return Score(10.0 * (lpr_H1 - lpr_H0) / math.log(10.0), float("nan"), diag)


# --------------------------------------------------------------------------------------------
# Driver  (src/PhyloCSF.ml)
# --------------------------------------------------------------------------------------------
@dataclass
class Options:
    strategy: str = "mle"  # mle|fixed|omega|nop
    remove_ref_gaps: bool = False
    allow_ref_gaps: bool = False
    species: Optional[List[str]] = None
    frames: int = 1
    orf: str = "AsIs"
    min_codons: int = 25
    all_scores: bool = False
    bls: bool = False
    anc_comp: bool = False
    dna: bool = False
    aa: bool = False
    debug: bool = False
    omega_H1: Optional[float] = None
    sigma_H1: Optional[float] = None


def input_mfa(lines: Sequence[str]):  # src/PhyloCSF.ml:83-118
    """Deviation: a blank line while reading a sequence makes the reference spin forever
    (PhyloCSF.ml:89-90 peeks without consuming); blank lines are skipped here."""
    rows: List[List[str]] = []
    for line in lines:
        if rows and line.strip() == "":
            continue
        if not rows or line.strip()[0:1] == ">":
            if line == "":
                continue
            if line[0] != ">":
                raise OracleFailure("invalid MFA alignment: bad header")
            hdr = line[1:]
            sp = (hdr[: hdr.index("|")] if "|" in hdr else hdr).strip()
            rows.append([sp, ""])
        else:
            rows[-1][1] += line.strip()
    if not rows:
        raise OracleFailure("invalid MFA alignment: No_value")
    species = [r[0] for r in rows]
    seqs = [r[1] for r in rows]
    seqlen = len(seqs[0])
    if seqlen == 0 or any(s == "" for s in species) or any(len(s) != seqlen for s in seqs):
        raise OracleFailure("invalid MFA alignment: empty species name or sequence, or sequence length mismatch")
    return species, seqs


def remove_ref_gaps(aln: Sequence[str]) -> List[str]:  # src/PhyloCSF.ml:122-132
    keep = [j for j, c in enumerate(aln[0]) if c != "-"]
    return ["".join(row[j] for j in keep) for row in aln]


def find_orfs(dna: str, ofs: int, orf_mode: str, min_codons: int):  # src/PhyloCSF.ml:134-196
    atg = orf_mode == "ATGStop"
    up = dna.upper()

    def is_start(p):
        return up[p:p + 3] == "ATG"

    def is_stop(p):
        return up[p:p + 3] in ("TAA", "TAG", "TGA")

    ln = len(dna)
    orfs: List[Tuple[int, int]] = []  # most recent first, like the OCaml list
    starts: List[int] = []
    for codon_lo in range(ofs, ln - 2):
        if (codon_lo - ofs) % 3 == 0:
            codon_hi = codon_lo + 2
            if (not atg and not starts and not is_stop(codon_lo)) or (atg and is_start(codon_lo)):
                starts.insert(0, codon_lo)
            if codon_hi + 3 < ln and is_stop(codon_hi + 1):
                for start in starts:
                    if codon_hi > start + 2:
                        orfs.insert(0, (start, codon_hi))
                starts = []
    if not atg:
        for start in starts:
            rem = ln - start
            orfs.insert(0, (start, start + (rem // 3) * 3 - 1))
    if orf_mode == "StopStop3":
        allsub = []
        for lo, hi in orfs:
            sub = [(lo, hi)]
            codons = (hi - lo + 1) // 3
            lo2 = lo + (codons // 3) * 3
            if lo2 > lo:
                sub.insert(0, (lo2, hi))
            lo3 = lo + (2 * codons // 3) * 3
            if lo3 > lo2 and lo3 > lo:
                sub.insert(0, (lo3, hi))
            allsub.extend(sub)
        orfs = allsub
    if orf_mode == "ToFirstStop" and orfs:
        first = orfs[-1]
        orfs = [first] if first[0] == ofs else []
    if orf_mode == "FromLastStop" and orfs:
        last = orfs[0]
        orfs = [last] if ln - last[1] <= 3 else []
    if orf_mode == "ToOrFromStop" and orfs:
        first = orfs[-1]
        last = orfs[0]
        orfs = [first] if first[0] == ofs else []
        if ln - last[1] <= 3 and first != last:
            orfs = [last] + orfs
    return [(lo, hi) for lo, hi in reversed(orfs) if (hi - lo + 1) // 3 >= min_codons]


def candidate_regions(dna: str, opts: Options):  # src/PhyloCSF.ml:198-217
    if opts.orf == "AsIs":
        hi = len(dna) - 1
        r = [(False, 0, hi)]
        if opts.frames != 1:
            r += [(False, 1, hi), (False, 2, hi)]
        if opts.frames == 6:
            r += [(True, 0, hi), (True, 1, hi), (True, 2, hi)]
        return r
    out = [(False, lo, hi) for lo, hi in find_orfs(dna, 0, opts.orf, opts.min_codons)]
    if opts.frames != 1:
        for o in (1, 2):
            out += [(False, lo, hi) for lo, hi in find_orfs(dna, o, opts.orf, opts.min_codons)]
    if opts.frames == 6:
        rc = revcomp(dna)
        for o in (0, 1, 2):
            out += [(True, lo, hi) for lo, hi in find_orfs(rc, o, opts.orf, opts.min_codons)]
    return out


def pleaves(tree: Tree, leaf_ord: Sequence[Optional[int]], aln: Sequence[str], lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
    """src/PhyloCSF.ml:219-246 -> uint8 [ncols][n_leaves], MARG for `Marginalize."""
    if hi is None:
        hi = len(aln[0]) - 1
    cols = []
    pos = lo
    while pos + 2 <= hi:
        cols.append([MARG if r is None else codon_code(aln[r][pos], aln[r][pos + 1], aln[r][pos + 2]) for r in leaf_ord])
        pos += 3
    if not cols:
        return np.zeros((0, tree.n_leaves), dtype=np.uint8)
    return np.array(cols, dtype=np.uint8)


def bls_score(nt: Node, aln: Sequence[str], which_row: dict, lo: int, hi: int) -> float:  # src/PhyloCSF.ml:252-262
    total = 0.0
    for i in range(lo, hi + 1):
        def keep(sp):
            r = which_row.get(sp)
            return r is not None and aln[r][i] not in "-.N"
        st = newick_subtree(keep, nt)
        total += newick_total_length(st) if st is not None else 0.0
    return total / (newick_total_length(nt) * float(hi - lo + 1))


def _ocaml_ge(a, b) -> bool:
    """a >= b under OCaml's structural comparison for (score record, rc, lo, hi) tuples, as used by
    List.reduce max (src/PhyloCSF.ml:376). NaN compares unordered => false. Diagnostics (a string
    list that only differs when the floats already differ) are skipped."""
    for x, y in zip(a, b):
        if isinstance(x, float) and (math.isnan(x) or math.isnan(y)):
            return False
        if x > y:
            return True
        if x < y:
            return False
    return True


@dataclass
class ParamSet:
    nt: Node  # after --species pruning
    tree: Tree
    model: Optional[PhyloCSFModel]


def load_paramset(prefix: str, opts: Options) -> ParamSet:  # src/PhyloCSF.ml:406-467
    nt = newick_parse(open(prefix + ".nh").read())
    if opts.species:
        want = set(opts.species)
        snt = newick_subtree(lambda s: s in want, nt)
        if snt is None or newick_leaves(snt) <= 1:
            raise OracleFailure("specify at least two available --species")
    else:
        snt = nt
    t = Tree.of_newick(snt)
    model = None
    if opts.strategy in ("mle", "fixed"):
        s1, pi1 = read_ecm(prefix + "_coding.ECM")
        s2, pi2 = read_ecm(prefix + "_noncoding.ECM")
        model = PhyloCSFModel(s1, pi1, s2, pi2, t)
    return ParamSet(snt, t, model)


def evaluate(ps: ParamSet, opts: Options, codes: np.ndarray, trace: Optional[dict] = None) -> Score:
    if opts.strategy in ("mle", "fixed"):
        return ps.model.score(opts.strategy, codes, trace)
    if opts.strategy == "omega":
        kw = {}
        if opts.omega_H1 is not None:
            kw["omega_H1"] = opts.omega_H1
        if opts.sigma_H1 is not None:
            kw["sigma_H1"] = opts.sigma_H1
        return omega_score(ps.tree, codes, trace=trace, **kw)
    return Score(0.0, 0.0, [])


def process_alignment(ps: ParamSet, opts: Options, name: str, lines: Sequence[str]) -> List[str]:
    """src/PhyloCSF.ml:280-389 -> the stdout lines for one alignment."""
    out: List[str] = []
    try:
        species, aln = input_mfa(lines)
        if opts.remove_ref_gaps:
            aln = remove_ref_gaps(aln)
        aln = [s.replace("u", "t").replace("U", "T") for s in aln]
        if not opts.allow_ref_gaps and "-" in aln[0]:
            raise OracleFailure("the reference sequence (first alignment row) must be ungapped")
        rc_aln = [revcomp(s) for s in aln]
        t = ps.tree
        t_species = set(t.labels[: t.n_leaves])
        wtf = sorted(set(species) - t_species)
        if wtf:
            raise OracleFailure("parameters not available for species: " + " ".join(wtf))
        which_row = {}
        for i, sp in enumerate(species):
            which_row[sp] = i
        leaf_ord = [which_row.get(t.labels[i]) for i in range(t.n_leaves)]
        rgns = candidate_regions(aln[0], opts)
    except OracleFailure as e:
        return ["%s\tabort\t%s" % (name, _exn_string(e))]
    try:
        if not rgns:
            raise OracleFailure("no sufficiently long ORFs found")
        results = []
        for rc, lo, hi in rgns:
            try:
                codes = pleaves(t, leaf_ord, rc_aln if rc else aln, lo, hi)
                results.append((evaluate(ps, opts, codes), rc, lo, hi))
            except OracleFailure as e:
                strand = ("\t-" if rc else "\t+") if opts.frames == 6 else ""
                out.append("%s\texception\t%d\t%d%s\t%s" % (name, lo, hi, strand, _exn_string(e)))
        if not results:
            raise OracleFailure("no regions successfully evaluated")

        def report(ty, item):
            rslt, rc, lo, hi = item
            s = "%s\t%s\t%.4f" % (name, ty, rslt.score)
            if opts.frames != 1 or opts.orf != "AsIs":
                s += "\t%d\t%d" % (lo, hi)
            if opts.frames == 6:
                s += "\t%s" % ("-" if rc else "+")
            if opts.bls:
                s += "\t%.4f" % bls_score(ps.nt, rc_aln if rc else aln, which_row, lo, hi)
            if opts.anc_comp:
                s += "\t%.4f" % rslt.anc_comp_score
            refdna = (rc_aln if rc else aln)[0][lo:hi + 1]
            if opts.dna:
                s += "\t%s" % refdna
            if opts.aa:
                s += "\t%s" % translate(refdna)
            if opts.debug:
                s += "\t#" + "".join(" %s=%s" % kv for kv in rslt.diagnostics)
            out.append(s)

        if opts.all_scores:
            for item in results:
                report("orf_score(decibans)", item)
        best = results[0]
        for item in results[1:]:
            ka = (best[0].score, best[0].anc_comp_score, best[1], best[2], best[3])
            kb = (item[0].score, item[0].anc_comp_score, item[1], item[2], item[3])
            best = best if _ocaml_ge(ka, kb) else item
        report("max_score(decibans)" if (opts.orf != "AsIs" or opts.frames != 1) else "score(decibans)", best)
    except OracleFailure as e:
        out.append("%s\tfailure\t%s" % (name, _exn_string(e)))
    return out


def _exn_string(e: Exception) -> str:
    """Printexc.to_string of a Failure is 'Failure("msg")'. Messages of Invalid_argument / Gsl_exn
    are kept verbatim (they already carry their constructor in the text where it matters)."""
    msg = str(e)
    if msg.startswith(("Invalid_argument", "Gsl_exn", "Failure")):
        return msg
    return 'Failure("%s")' % msg


# --------------------------------------------------------------------------------------------
# Simulator (lib/CamlPaml/PhyloModel.ml:38-54, Tools.ml:22-41; as used by src/testSim.ml:52-65)
# --------------------------------------------------------------------------------------------
def simulate_columns(model: PhyloModel, ncols: int, rng: np.random.Generator, redraw_ref_stops: bool = True) -> np.ndarray:
    """Vectorised restatement: root ~ prior, each child ~ row parent of P via inverse-CDF on the
    cumulative sums. Returns uint8 codes [ncols][n_leaves]. Columns whose reference-species (leaf 0)
    codon is a stop are redrawn (testSim.ml:54-57)."""
    t = model.tree
    k = model.pms.shape[1]
    cum_prior = np.cumsum(model.prior())
    cum_p = np.cumsum(model.pms, axis=2)
    out = np.empty((ncols, t.n_leaves), dtype=np.uint8)
    todo = np.arange(ncols)
    while todo.size:
        m = todo.size
        a = np.empty((t.size, m), dtype=np.int64)
        a[t.root] = np.minimum(np.searchsorted(cum_prior, rng.random(m) * cum_prior[-1], side="left"), k - 1)
        for i in range(t.root - 1, -1, -1):
            par = a[t.parents[i]]
            cdf = cum_p[i][par]  # [m, k]
            u = rng.random(m) * cdf[:, -1]
            a[i] = np.minimum((cdf < u[:, None]).sum(axis=1), k - 1)
        leaves = a[: t.n_leaves].T.astype(np.uint8)
        ok = ~np.isin(leaves[:, 0], STOPS) if redraw_ref_stops else np.ones(m, dtype=bool)
        out[todo[ok]] = leaves[ok]
        todo = todo[~ok]
    return out


def codes_to_alignment(codes: np.ndarray) -> List[str]:
    """[ncols][n_leaves] codon codes -> one nucleotide string per leaf."""
    table = np.array([codon_of_index(i) for i in range(K)] + ["---"])
    return ["".join(table[codes[:, l]]) for l in range(codes.shape[1])]
