// functors.cuh — built-in O(n) device objective functors ("block functors").
//
// A block functor evaluates BS consecutive coordinates at a time: given the BS values of x it
// returns the block's contribution to f and writes the BS gradient entries.  The same functor
// body is used by the multi-CTA trial kernel (objectives.cu), the single-CTA device-resident
// line-search kernel (qn_device.cu) and the one-warp-per-problem batched solver (batched.cu).
// A user who wants a new built-in objective adds a functor here and one dispatch line; arbitrary
// user kernels plug in through osb_objective_create_user instead.
#pragma once
#include "common.cuh"

namespace osb {

// Extended Rosenbrock: f = sum_{i<n/2} 100 (x_{2i+1} - x_{2i}^2)^2 + (1 - x_{2i})^2.
// Not in the reference crate (only its 2-D form appears in wasm/demo/index.html:441-453); the
// arithmetic below is the statement-for-statement twin of oracle/oracle.cpp `Rosenbrock`.
struct RosenbrockFn {
  static constexpr int BS = 2;
  HD double block(int64_t /*i0*/, const double* xb, double* gb) const {
    const double a = xb[0], b = xb[1];
    const double t1 = b - a * a;
    const double t2 = 1.0 - a;
    gb[0] = -400.0 * (a * t1) - 2.0 * t2;
    gb[1] = 200.0 * t1;
    return 100.0 * (t1 * t1) + t2 * t2;
  }
};

// Separable quadratic f = sum 0.5 c_i (x_i - a_i)^2  (twin of oracle `SeparableQuadratic`)
struct SepQuadFn {
  static constexpr int BS = 1;
  const double* c;
  const double* a;
  HD double block(int64_t i0, const double* xb, double* gb) const {
    const double dlt = xb[0] - a[i0];
    const double cd = c[i0] * dlt;
    gb[0] = cd;
    return 0.5 * (cd * dlt);
  }
};

}  // namespace osb
