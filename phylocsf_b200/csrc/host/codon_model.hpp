// codon_model.hpp — host-side model setup of the drop-in: codon alphabet, .ECM reader, rate-matrix
// assembly for the ECM and omega models, and the diagonalisation Q = S diag(lambda) S^-1 that feeds
// the P(t) kernel (pcsf_model_set).
//   Codon64 / DNA            <- lib/CamlPaml/Code.ml:11-58,133-181
//   read_ecm                 <- src/ECM.ml:18-72
//   ecm_q                    <- src/PhyloCSFModel.ml:11-30 + lib/CamlPaml/PhyloModel.ml:76-106
//   omega_q                  <- src/OmegaModel.ml:21-80
//   QDiag (of_Q/equilibrium) <- lib/CamlPaml/Q.ml:124-177
// The reference diagonalises with GSL's general non-symmetric solver + a complex LU inverse. Both
// model families are reversible (q_ij = w_j * sym_ij), so here Q is symmetrised with its stationary
// weights w and solved with a cyclic Jacobi sweep: A = W^1/2 Q W^-1/2 = U L U^T, S = W^-1/2 U,
// S^-1 = U^T W^1/2 (no inversion). P(t) is invariant to the eigenbasis up to rounding; measured
// effect on scores ~1e-10 dB (tests/test_host_model.py).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "newick_tree.hpp"

namespace pcsf {
namespace host {

constexpr int KC = 64;
constexpr int CODE_MARG = 64;

inline int dna_index(char c) {  // Code.ml:25-30
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}
inline int codon_code(char a, char b, char c) {  // src/PhyloCSF.ml:233-240
    const int i1 = dna_index(a), i2 = dna_index(b), i3 = dna_index(c);
    return (i1 < 0 || i2 < 0 || i3 < 0) ? CODE_MARG : 16 * i1 + 4 * i2 + i3;
}
static const char kTranslation[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";  // Code.ml:159-178
inline bool is_stop_index(int c) { return c == 48 || c == 50 || c == 56; }                               // TAA TAG TGA

// Code.ml:39-58: raises on anything but ACGTacgtNn-
inline std::string revcomp(const std::string& s) {
    std::string out(s.size(), ' ');
    for (size_t i = 0; i < s.size(); i++) {
        char c = s[i], r;
        switch (c) {
            case 'A': r = 'T'; break; case 'G': r = 'C'; break; case 'C': r = 'G'; break; case 'T': r = 'A'; break;
            case 'a': r = 't'; break; case 'g': r = 'c'; break; case 'c': r = 'g'; break; case 't': r = 'a'; break;
            case 'N': r = 'N'; break; case 'n': r = 'n'; break; case '-': r = '-'; break;
            default: throw invalid_arg(std::string("unrecognized nucleotide ") + c);
        }
        out[s.size() - 1 - i] = r;
    }
    return out;
}

inline std::string ocaml_trim(const std::string& s) {
    size_t b = 0, e = s.size();
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\014'; };
    while (b < e && ws(s[b])) b++;
    while (e > b && ws(s[e - 1])) e--;
    return s.substr(b, e - b);
}
inline double float_of_string(const std::string& s) {
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (s.empty() || end != s.c_str() + s.size()) throw failure("float_of_string");
    return v;
}
inline std::vector<std::string> split_sp(const std::string& line) {  // Str.split (regexp " ") + trim + drop empties
    std::vector<std::string> out;
    std::stringstream ss(line);
    std::string tok;
    while (std::getline(ss, tok, ' ')) {
        tok = ocaml_trim(tok);
        if (!tok.empty()) out.push_back(tok);
    }
    return out;
}

struct ECM {
    std::vector<double> s = std::vector<double>(KC * KC, 0.0);
    std::vector<double> pi = std::vector<double>(KC, 0.0);
};

inline ECM read_ecm(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw HostError("Sys_error(\"" + path + ": No such file or directory\")");
    std::vector<std::string> lines;
    std::string ln;
    while (std::getline(f, ln)) lines.push_back(ln);
    if ((int)lines.size() < KC + 7) throw failure("ECM.import_parameters");
    ECM e;
    for (int i = 1; i < KC; i++) {
        const auto toks = split_sp(lines[i - 1]);
        if ((int)toks.size() < i) throw invalid_arg("index out of bounds");
        for (int j = 0; j < i; j++) {
            const double v = float_of_string(toks[j]);
            e.s[i * KC + j] = v;
            e.s[j * KC + i] = v;
        }
    }
    if (!ocaml_trim(lines[KC - 1]).empty()) throw failure("ECM.import_parameters");
    const auto pt = split_sp(lines[KC]);
    if ((int)pt.size() < KC) throw invalid_arg("index out of bounds");
    for (int i = 0; i < KC; i++) e.pi[i] = float_of_string(pt[i]);
    std::vector<std::string> codons;
    for (int i = 3; i <= 6; i++)
        for (auto& t : split_sp(lines[KC + i])) codons.push_back(t);
    if ((int)codons.size() != KC) throw failure("ECM.import_parameters: incorrect codon order");
    for (int i = 0; i < KC; i++)
        if (codons[i].size() != 3 || codon_code(codons[i][0], codons[i][1], codons[i][2]) != i)
            throw failure("ECM.import_parameters: incorrect codon order");
    return e;
}

// fill_q_diagonal (PhyloModel.ml:76-84), the scale expression (PhyloCSFModel.ml:24-30 /
// OmegaModel.ml:76-80) and the division by it (PhyloModel.ml:94-104), in the order the Expr trees
// evaluate.
inline void fill_diag_and_scale(std::vector<double>& q, const std::vector<double>& pi, bool skip_zero_diag) {
    for (int i = 0; i < KC; i++) {
        double tot = 0.0;
        for (int j = 0; j < KC; j++)
            if (i != j) tot = q[i * KC + j] + tot;
        q[i * KC + i] = 0.0 - tot;
    }
    double factor = 0.0;
    for (int i = 0; i < KC; i++) {
        if (skip_zero_diag && q[i * KC + i] == 0.0) continue;
        factor = factor - pi[i] * q[i * KC + i];
    }
    if (!(factor > 0.0)) throw failure("CamlPaml.P14n.instantiate_q: Q scale evaluated to a non-positive value");
    for (auto& v : q) v = v / factor;
}

inline std::vector<double> ecm_q(const ECM& e) {
    std::vector<double> q(KC * KC, 0.0);
    for (int i = 0; i < KC; i++)
        for (int j = 0; j < KC; j++)
            if (i != j) q[i * KC + j] = e.s[i * KC + j] * e.pi[j];
    fill_diag_and_scale(q, e.pi, true);
    return q;
}

// OmegaModel.ml:24-42: codon frequencies from F3x4 settings v[3..11] and sigma = v[2]
inline std::vector<double> omega_pi(const double* v) {
    auto sc = [&](int i1, int i2, int i3) {
        const double f1 = (i1 == 3 ? 1.0 : v[3 + i1]) / (v[3] + (v[4] + (v[5] + 1.0)));
        const double f2 = (i2 == 3 ? 1.0 : v[6 + i2]) / (v[6] + (v[7] + (v[8] + 1.0)));
        const double f3 = (i3 == 3 ? 1.0 : v[9 + i3]) / (v[9] + (v[10] + (v[11] + 1.0)));
        return f1 * (f2 * f3);
    };
    const double denom = 1.0 - (1.0 - v[2]) * (sc(3, 0, 0) + (sc(3, 0, 2) + sc(3, 2, 0)));
    std::vector<double> pi(KC);
    for (int i = 0; i < KC; i++) pi[i] = sc(i / 16, (i / 4) % 4, i % 4) / denom;
    return pi;
}

// OmegaModel.ml:44-80 evaluated at settings v = [kappa; omega; sigma; 9 x F3x4]
inline std::vector<double> omega_q(const double* v, std::vector<double>* pi_out = nullptr) {
    const double kappa = v[0], omega = v[1];
    const std::vector<double> pi = omega_pi(v);
    std::vector<double> q(KC * KC, 0.0);
    for (int i = 0; i < KC; i++) {
        const int ii[3] = {i / 16, (i / 4) % 4, i % 4};
        for (int j = 0; j < KC; j++) {
            const int jj[3] = {j / 16, (j / 4) % 4, j % 4};
            int nd = 0, da = 0, dbb = 0;
            for (int p = 0; p < 3; p++)
                if (ii[p] != jj[p]) { nd++; da = ii[p]; dbb = jj[p]; }
            if (nd != 1) continue;
            const bool transition = (da == 0 && dbb == 2) || (da == 2 && dbb == 0) || (da == 1 && dbb == 3) || (da == 3 && dbb == 1);
            const double kp = transition ? kappa : 1.0;
            const double op = (!is_stop_index(i) && !is_stop_index(j) && kTranslation[i] != kTranslation[j]) ? omega : 1.0;
            q[i * KC + j] = pi[j] * (kp * op);
        }
    }
    fill_diag_and_scale(q, pi, false);
    if (pi_out) *pi_out = pi;
    return q;
}

// Cyclic Jacobi for a symmetric n x n matrix (row-major). On return a's diagonal holds the
// eigenvalues and u's columns the eigenvectors.
inline void jacobi_eigen(std::vector<double>& a, std::vector<double>& u, int n) {
    u.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) u[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) (i == j ? diag : off) += a[i * n + j] * a[i * n + j];
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a[p * n + q];
                if (apq == 0.0) continue;
                const double app = a[p * n + p], aqq = a[q * n + q];
                // after a few sweeps, off-diagonals that no longer register against the diagonal are
                // flushed instead of rotated (classical Jacobi threshold rule)
                if (sweep > 4 && std::fabs(apq) <= 1e-20 * std::fabs(app) && std::fabs(apq) <= 1e-20 * std::fabs(aqq)) {
                    a[p * n + q] = a[q * n + p] = 0.0;
                    continue;
                }
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {  // rotate columns p, q
                    const double akp = a[k * n + p], akq = a[k * n + q];
                    a[k * n + p] = c * akp - s * akq;
                    a[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {  // rotate rows p, q
                    const double apk = a[p * n + k], aqk = a[q * n + k];
                    a[p * n + k] = c * apk - s * aqk;
                    a[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double ukp = u[k * n + p], ukq = u[k * n + q];
                    u[k * n + p] = c * ukp - s * ukq;
                    u[k * n + q] = s * ukp + c * ukq;
                }
            }
    }
}

struct QDiag {
    std::vector<double> q, S, Sinv, lam, pi_eq;
    double tol = 1e-6;

    // w: positive stationary weights of the reversible Q (w_i q_ij = w_j q_ji), any normalisation.
    static QDiag of_reversible_Q(const std::vector<double>& qm, const std::vector<double>& w) {
        const int n = KC;
        QDiag d;
        d.q = qm;
        std::vector<double> sw(n), a((size_t)n * n), u;
        for (int i = 0; i < n; i++) {
            if (!(w[i] > 0.0)) throw failure("CamlPaml.Q: non-positive stationary weight; cannot symmetrise the rate matrix");
            sw[i] = std::sqrt(w[i]);
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) a[i * n + j] = sw[i] * qm[i * n + j] / sw[j];
        double asym = 0.0, scale = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                asym = std::max(asym, std::fabs(a[i * n + j] - a[j * n + i]));
                scale = std::max(scale, std::fabs(a[i * n + j]));
                const double m = 0.5 * (a[i * n + j] + a[j * n + i]);
                a[i * n + j] = a[j * n + i] = m;
            }
        if (asym > 1e-9 * std::max(scale, 1.0)) throw failure("CamlPaml.Q: rate matrix is not reversible (complex eigen path is out of scope)");
        jacobi_eigen(a, u, n);
        d.lam.resize(n);
        d.S.resize((size_t)n * n);
        d.Sinv.resize((size_t)n * n);
        for (int k = 0; k < n; k++) d.lam[k] = a[k * n + k];
        for (int i = 0; i < n; i++)
            for (int k = 0; k < n; k++) {
                d.S[i * n + k] = u[i * n + k] / sw[i];
                d.Sinv[k * n + i] = u[i * n + k] * sw[i];
            }
        // equilibrium, Q.ml:153-177
        int p = 0;
        double best = INFINITY;
        for (int i = 0; i < n; i++)
            if (std::fabs(d.lam[i]) < best) { best = std::fabs(d.lam[i]); p = i; }
        if (best > d.tol) throw failure("CamlPaml.Q.equilibrium: smallest-magnitude eigenvalue is unacceptably large; check rate matrix validity or increase tol");
        double mass = 0.0;
        for (int i = 0; i < n; i++) mass += d.Sinv[p * n + i];
        d.pi_eq.resize(n);
        for (int i = 0; i < n; i++) d.pi_eq[i] = d.Sinv[p * n + i] / mass;
        return d;
    }
};

}  // namespace host
}  // namespace pcsf
