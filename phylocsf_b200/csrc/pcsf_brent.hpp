// pcsf_brent.hpp — PhyloCSFModel.maximize_lpr as a resumable state machine.
//
// The reference runs, per region and per model (src/PhyloCSFModel.ml:84-99):
//     Fit.find_init ~maxtries:250 ~logspace:true   (lib/CamlPaml/Fit.ml:27-48)
//     Gsl.Min.make BRENT on -f, iterate until (ub - lb)/x <= accuracy, then f x
// with one likelihood sweep per function evaluation. Here the control flow is turned inside out so
// that thousands of regions advance in lock step and every round's evaluations run as one batched
// kernel sequence: candidate() names the next abscissa this region needs, feed() delivers f there.
// The iterate sequence is the reference's; evaluations the reference repeats at an abscissa it has
// already visited (fminimizer_set re-evaluates lo/hi/init, and the final `f x`) reuse the stored
// value, which is bit-identical because an evaluation is a pure function of (region, rho).
//
// GSL's Brent (min/brent.c, min/fsolver.c) and OCaml's Random are not in the reference tree; they
// are restated from their published algorithms (DESIGN.md "parity unpinned").
#pragma once
#include <cmath>
#include <cstdint>

namespace pcsf {

static const double kOcamlRandom250[250] = {
#include "pcsf_ocaml_random.inc"
};

struct MaximizeLpr {
    enum Phase { FI_LO, FI_HI, FI_INIT, FI_RANDOM, BR_V, BR_ITER, DONE };
    Phase phase = FI_LO;
    double init, lo, hi, accuracy;
    // find_init
    double flo = 0, fhi = 0, elo = 0, ehi = 0, x = 0, fx = 0, ex = 0;
    int tries = 0;
    // brent (on g = -f)
    double x_lower = 0, g_lower = 0, x_upper = 0, g_upper = 0, x_min = 0, g_min = 0, e_min = 0;
    double v = 0, w = 0, d = 0, e = 0, g_v = 0, g_w = 0, u = 0;
    // results
    double result_x = NAN, result_f = NAN, result_elpr = NAN;
    int32_t status = 0, nevals = 0, iterations = 0;

    MaximizeLpr(double init_, double lo_, double hi_, double acc_) : init(init_), lo(lo_), hi(hi_), accuracy(acc_) {}

    bool done() const { return phase == DONE; }

    double candidate() const {
        switch (phase) {
            case FI_LO: return lo;
            case FI_HI: return hi;
            case FI_INIT: return init;
            case FI_RANDOM: return x;
            case BR_V: return v;
            case BR_ITER: return u;
            default: return NAN;
        }
    }

    void finish(double rx, double rf, double re) {
        result_x = rx;
        result_f = rf;
        result_elpr = re;
        nevals++;  // the reference's trailing `f x` / `f good_init`
        phase = DONE;
    }
    void failed(int32_t st) {
        status |= st;
        phase = DONE;
    }

    // Fit.ml:33-48 loop header, entered after f(init) and after every random try
    void find_init_step() {
        const int maxtries = 250;
        if (tries < maxtries && (fx <= flo || fx <= fhi)) {
            status |= 64;  // PCSF_ST_RANDOM_INIT
            const double width = std::log(hi) - std::log(lo);
            x = std::exp(std::log(lo) + kOcamlRandom250[tries] * width);
            phase = FI_RANDOM;
            return;
        }
        double good = x, fgood = fx, egood = ex;
        if (tries == maxtries) {
            if (flo > fhi) { good = lo; fgood = flo; egood = elo; }
            else { good = hi; fgood = fhi; egood = ehi; }
        }
        if (lo < good && good < hi) {
            // gsl_min_fminimizer_set: f(lo), f(hi), f(min) again (3 evaluations), then the checks
            nevals += 3;
            if (!std::isfinite(flo) || !std::isfinite(fhi) || !std::isfinite(fgood)) return failed(16);
            x_lower = lo; g_lower = -flo;
            x_upper = hi; g_upper = -fhi;
            x_min = good; g_min = -fgood; e_min = egood;
            if (g_min >= g_lower || g_min >= g_upper) return failed(32);  // endpoints do not enclose a minimum
            const double golden = 0.3819660;
            v = x_lower + golden * (x_upper - x_lower);  // brent_init
            w = v;
            d = 0;
            e = 0;
            phase = BR_V;
        } else {
            finish(good, fgood, egood);  // PhyloCSFModel.ml:98-99
        }
    }

    // first half of brent_iterate: choose u (GSL min/brent.c)
    void brent_propose() {
        const double x_left = x_lower, x_right = x_upper, z = x_min;
        double dd = e, ee = d;  // sic: GSL reads d from state->e and e from state->d
        const double g_z = g_min;
        const double golden = 0.3819660;
        const double w_lower = z - x_left, w_upper = x_right - z;
        const double tolerance = 1.4901161193847656e-08 * std::fabs(z);
        double p = 0, q = 0, r = 0;
        const double midpoint = 0.5 * (x_left + x_right);
        if (std::fabs(ee) > tolerance) {
            r = (z - w) * (g_z - g_v);
            q = (z - v) * (g_z - g_w);
            p = (z - v) * q - (z - w) * r;
            q = 2 * (q - r);
            if (q > 0) p = -p; else q = -q;
            r = ee;
            ee = dd;
        }
        if (std::fabs(p) < std::fabs(0.5 * q * r) && p < q * w_lower && p < q * w_upper) {
            const double t2 = 2 * tolerance;
            dd = p / q;
            const double uu = z + dd;
            if ((uu - x_left) < t2 || (x_right - uu) < t2) dd = (z < midpoint) ? tolerance : -tolerance;
        } else {
            ee = (z < midpoint) ? x_right - z : -(z - x_left);
            dd = golden * ee;
        }
        if (std::fabs(dd) >= tolerance) u = z + dd;
        else u = z + ((dd > 0) ? tolerance : -tolerance);
        e = ee;
        d = dd;
    }

    // second half of brent_iterate, with g_u = -f(u)
    void brent_update(double g_u, double e_u) {
        const double z = x_min, g_z = g_min;
        if (g_u <= g_z) {
            if (u < z) { x_upper = z; g_upper = g_z; } else { x_lower = z; g_lower = g_z; }
            v = w; g_v = g_w;
            w = z; g_w = g_z;
            x_min = u; g_min = g_u; e_min = e_u;
        } else {
            if (u < z) { x_lower = u; g_lower = g_u; } else { x_upper = u; g_upper = g_u; }
            if (g_u <= g_w || w == z) {
                v = w; g_v = g_w;
                w = u; g_w = g_u;
            } else if (g_u <= g_v || v == z || v == w) {
                v = u; g_v = g_u;
            }
        }
    }

    // f = lpr at candidate(), elpr = elpr_anc there, st = PCSF_ST_* bits of that evaluation
    void feed(double f, double elpr, int32_t st) {
        nevals++;
        if (st & ~16) return failed(st & ~16);  // P(t) failure inside f raises in the reference
        switch (phase) {
            case FI_LO: flo = f; elo = elpr; phase = FI_HI; break;
            case FI_HI: fhi = f; ehi = elpr; phase = FI_INIT; break;
            case FI_INIT:
                x = init; fx = f; ex = elpr;
                find_init_step();
                break;
            case FI_RANDOM:
                fx = f; ex = elpr;
                tries++;
                find_init_step();
                break;
            case BR_V:
                if (!std::isfinite(f)) return failed(16);  // SAFE_FUNC_CALL
                g_v = -f; g_w = -f;
                phase = BR_ITER;
                brent_propose();
                break;
            case BR_ITER: {
                if (!std::isfinite(f)) return failed(16);
                brent_update(-f, elpr);
                iterations++;
                if (((x_upper - x_lower) / x_min) > accuracy) brent_propose();  // PhyloCSFModel.ml:89-94
                else finish(x_min, -g_min, e_min);                              // :96-97
                break;
            }
            default: break;
        }
    }
};

}  // namespace pcsf
