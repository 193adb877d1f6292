// ls_automaton.cuh — the scalar control logic of the reference's line searches as a resumable
// state machine, compiled for host (host-driven engine) and device (device-resident engine,
// batched kernels).  The machine never touches vectors: the caller evaluates the objective at
// x + t*d for the step `t` the machine requests and feeds back the scalars
//     f_t = f(x + t d),  gd_t = g(x + t d) . d,  dn = ||P(x + t d) - x||^2   (BackTrackingB only).
//
// Follows, statement by statement:
//   BackTracking   src/line_search/backtracking.rs:19-59
//   BackTrackingB  src/line_search/backtracking_b.rs:24-34,52-90
//   MoreThuente    src/line_search/morethuente.rs:64-149,164-298 (MoreThuenteB: morethuente_b.rs:173-325)
//   GLLQuadratic   src/line_search/gll_quadratic.rs:30-99
//   NoSearch       src/line_search/nosearch.rs:3-15
// including the quirks catalogued in SURVEY §3.4 (update_interval receives the NEW t; tl is
// re-evaluated every inner iteration; case 4 evaluates at tu, possibly +inf; NaN -> t_min clamp;
// exact t == tl / t == tu exits; a NaN/inf trial in backtracking does not consume an iteration).
#pragma once
#include "common.cuh"

namespace osb {

enum LSKind : int { LS_BACKTRACKING = 0, LS_BACKTRACKING_B = 1, LS_MORETHUENTE = 2, LS_MORETHUENTE_B = 3, LS_GLL = 4, LS_NOSEARCH = 5 };

constexpr int GLL_MAX_M = 64;

// Constructor parameters + the state the reference keeps INSIDE the line-search object across
// outer iterations (GLLQuadratic.f_previous, MoreThuenteB.t_max).
struct LSParams {
  int kind;
  double c1, c2, beta;
  double t_min, t_max, delta_min, delta, delta_max;
  double sigma1, sigma2;
  int m;                    // GLL window length
  int n_prev;               // GLL: valid entries of f_prev
  double f_prev[GLL_MAX_M]; // GLL: oldest first (Vec::remove(0) + push, gll_quadratic.rs:30-35)
};

enum MTPhase : int { MT_EVAL_T = 0, MT_EVAL_TL = 1, MT_EVAL_TU = 2 };

struct LSMachine {
  // per-search state
  double f0, gd0;       // f(x_k), g(x_k).d
  double t;             // current trial
  double result;        // valid when done
  int64_t i, max_iter;
  bool done;
  bool last_eval_is_result;  // the last fed evaluation was taken at `result` (and unprojected)
  // More–Thuente
  double tl, tu, f_max;
  bool modified, interval_converged;
  int phase;
  double phi_t_f, phi_t_g, psi_t_f, psi_t_g;
  double sv_f_tl, sv_g_tl, sv_f_t, sv_g_t;

  HD void finish(double t_res, bool current) {
    result = t_res;
    done = true;
    last_eval_is_result = current;
  }

  // `tmax_candidate`: MoreThuenteB's feasible step bound (morethuente_b.rs:185-197), ignored otherwise.
  // FK >= 0 fixes the line-search kind at compile time (device head kernels are specialised per kind so
  // that each carries only its own automaton: a cold, large kernel body is bound by instruction fetch)
  template <int FK = -1>
  HD void begin(LSParams& p, double f0_, double gd0_, int64_t max_iter_, double tmax_candidate) {
    f0 = f0_;
    gd0 = gd0_;
    max_iter = max_iter_;
    i = 0;
    done = false;
    last_eval_is_result = false;
    modified = false;
    interval_converged = false;
    phase = MT_EVAL_T;
    f_max = f0_;
    switch (FK >= 0 ? FK : p.kind) {
      case LS_NOSEARCH:
        finish(1.0, false);
        return;
      case LS_GLL: {
        // append_new_f + f_max (gll_quadratic.rs:30-43,62-64)
        if (p.n_prev == p.m) {
          for (int q = 1; q < p.n_prev; ++q) p.f_prev[q - 1] = p.f_prev[q];
          p.n_prev -= 1;
        }
        p.f_prev[p.n_prev] = f0_;
        p.n_prev += 1;
        double mx = -INFINITY;
        for (int q = 0; q < p.n_prev; ++q) mx = rmax(p.f_prev[q], mx);
        f_max = mx;
        t = 1.0;
        break;
      }
      case LS_BACKTRACKING:
      case LS_BACKTRACKING_B:
        t = 1.0;
        break;
      case LS_MORETHUENTE_B:
        p.t_max = rmin(p.t_max, tmax_candidate);  // morethuente_b.rs:201 — permanent
        // fallthrough
      case LS_MORETHUENTE:
        t = rmin(rmax(1.0, p.t_min), p.t_max);  // morethuente.rs:176
        tl = p.t_min;
        tu = p.t_max;
        break;
    }
    if (max_iter <= 0) finish(t, false);  // loops do not execute; the initial t is returned
  }

  // step at which the objective must be evaluated next (only when !done)
  HD double request(const LSParams& p) const {
    (void)p;
    if (phase == MT_EVAL_TL) return tl;
    if (phase == MT_EVAL_TU) return tu;
    return t;
  }
  // BackTrackingB evaluates the objective at the PROJECTED trial (backtracking_b.rs:65-67)
  template <int FK = -1>
  HD bool wants_projection(const LSParams& p) const { return (FK >= 0 ? FK : p.kind) == LS_BACKTRACKING_B; }

  static HD double cubic_minimizer(double ta, double tb, double f_ta, double f_tb, double g_ta, double g_tb) {
    double s = 3. * (f_tb - f_ta) / (tb - ta);  // morethuente.rs:103
    double z = s - g_ta - g_tb;
    double w = sqrt(z * z - g_ta * g_tb);
    return ta + ((tb - ta) * ((w - g_ta - z) / (g_tb - g_ta + 2. * w)));
  }
  static HD double quadratic_minimizer_1(double ta, double tb, double f_ta, double f_tb, double g_ta) {
    double lin_int = (f_ta - f_tb) / (ta - tb);  // morethuente.rs:118
    return ta - 0.5 * ((ta - tb) * g_ta / (g_ta - lin_int));
  }
  static HD double quadratic_minimizer_2(double ta, double tb, double g_ta, double g_tb) {
    return ta - g_ta * ((ta - tb) / (g_ta - g_tb));  // morethuente.rs:131
  }
  static HD bool update_interval(double f_tl, double f_t, double g_t, double& tl_, double t_, double& tu_) {
    if (f_t > f_tl) {  // morethuente.rs:72-75
      tu_ = t_;
      return false;
    } else if (g_t * (tl_ - t_) > 0.) {
      tl_ = t_;
      return false;
    } else if (g_t * (tl_ - t_) < 0.) {
      tu_ = tl_;
      tl_ = t_;
      return false;
    }
    return true;
  }

  HD void mt_finish_iteration(const LSParams& p, double t_new) {
    t = rmin(rmax(t_new, p.t_min), p.t_max);  // morethuente.rs:290
    interval_converged = update_interval(sv_f_tl, sv_f_t, sv_g_t, tl, t, tu);  // :293 (NEW t)
    i += 1;
    phase = MT_EVAL_T;
    if (i >= max_iter) finish(t, false);  // :295-296
  }

  template <int FK = -1>
  HD void feed(LSParams& p, double f_t, double gd_t, double dn) {
    const int kind = FK >= 0 ? FK : p.kind;
    switch (kind) {
      case LS_BACKTRACKING:
      case LS_BACKTRACKING_B: {
        if (is_bad(f_t)) {  // backtracking.rs:37-41: shrink, do NOT count the iteration
          t *= p.beta;
          return;
        }
        bool ok = (kind == LS_BACKTRACKING) ? (f_t - f0 <= p.c1 * t * gd0)        // line_search/mod.rs:35
                                              : (f_t - f0 <= (-p.c1 / t) * dn);     // backtracking_b.rs:33
        if (ok) {
          finish(t, kind == LS_BACKTRACKING);
          return;
        }
        t *= p.beta;
        i += 1;
        if (i >= max_iter) finish(t, false);
        return;
      }
      case LS_GLL: {
        if (f_t - f_max <= p.c1 * t * gd0) {  // gll_quadratic.rs:72
          finish(t, true);
          return;
        }
        if (t <= 0.1) {
          t *= 0.5;
        } else {
          double t_tmp = -0.5 * t * t * gd0 / (f_t - f0 - t * gd0);  // :83-84
          if (t_tmp > p.sigma1 && t_tmp < p.sigma2 * t) t = t_tmp;
          else t = t_tmp * 0.5;
        }
        i += 1;
        if (i >= max_iter) finish(t, false);
        return;
      }
      case LS_MORETHUENTE:
      case LS_MORETHUENTE_B: {
        if (phase == MT_EVAL_T) {
          bool armijo = f_t - f0 <= p.c1 * t * gd0;                 // mod.rs:35
          bool curv = fabs(gd_t) <= p.c2 * fabs(gd0);               // mod.rs:55
          if (armijo && curv) { finish(t, true); return; }          // morethuente.rs:184-193
          else if (interval_converged) { finish(t, true); return; } // :194-196
          else if (t == tl) { finish(t, true); return; }            // :198-200
          else if (t == tu) { finish(t, true); return; }            // :202-204
          phi_t_f = f_t;
          phi_t_g = gd_t;
          psi_t_f = f_t - f0 - p.c1 * t * gd0;                      // :140-149
          psi_t_g = gd_t - p.c1 * gd0;
          if (!modified && psi_t_f <= 0. && phi_t_g > 0.) modified = true;  // :212-215
          phase = MT_EVAL_TL;
          return;
        }
        if (phase == MT_EVAL_TL) {
          double f_tl, g_tl, ft, gt;
          if (modified) {
            f_tl = f_t; g_tl = gd_t; ft = phi_t_f; gt = phi_t_g;
          } else {
            f_tl = f_t - f0 - p.c1 * tl * gd0;
            g_tl = gd_t - p.c1 * gd0;
            ft = psi_t_f; gt = psi_t_g;
          }
          sv_f_tl = f_tl; sv_g_tl = g_tl; sv_f_t = ft; sv_g_t = gt;
          double t_new;
          if (ft > f_tl) {  // case 1 :230-241
            double tc = cubic_minimizer(tl, t, f_tl, ft, g_tl, gt);
            double tq = quadratic_minimizer_1(tl, t, f_tl, ft, g_tl);
            if (fabs(tc - tl) < fabs(tq - tl)) t_new = tc;
            else t_new = 0.5 * (tq + tc);
          } else if (gt * g_tl < 0.) {  // case 2 :243-254
            double tc = cubic_minimizer(tl, t, f_tl, ft, g_tl, gt);
            double ts = quadratic_minimizer_2(tl, t, g_tl, gt);
            if (fabs(tc - t) >= fabs(ts - t)) t_new = tc;
            else t_new = ts;
          } else if (fabs(gt) <= fabs(g_tl)) {  // case 3 :256-272
            double tc = cubic_minimizer(tl, t, f_tl, ft, g_tl, gt);
            double ts = quadratic_minimizer_2(tl, t, g_tl, gt);
            double t_plus = (fabs(tc - t) < fabs(ts - t)) ? tc : ts;
            if (t > tl) t_new = rmin(t_plus, t + p.delta * (tu - t));
            else t_new = rmax(t_plus, t + p.delta * (tu - t));
          } else {  // case 4 :274-287 needs the oracle at tu
            phase = MT_EVAL_TU;
            return;
          }
          mt_finish_iteration(p, t_new);
          return;
        }
        // MT_EVAL_TU
        double f_tu, g_tu;
        if (modified) {
          f_tu = f_t; g_tu = gd_t;
        } else {
          f_tu = f_t - f0 - p.c1 * tu * gd0;
          g_tu = gd_t - p.c1 * gd0;
        }
        mt_finish_iteration(p, cubic_minimizer(tu, t, sv_f_t, f_tu, sv_g_t, g_tu));  // :286
        return;
      }
      default:
        finish(1.0, false);
    }
  }
};

inline LSParams ls_defaults(int kind) {
  LSParams p{};
  p.kind = kind;
  p.c1 = 1e-4; p.c2 = 0.9; p.beta = 0.5;
  p.t_min = 0.0; p.t_max = INFINITY; p.delta_min = 0.58333333; p.delta = 0.66; p.delta_max = 1.1;
  p.sigma1 = 0.1; p.sigma2 = 0.9;
  p.m = 1; p.n_prev = 0;
  return p;
}

}  // namespace osb
