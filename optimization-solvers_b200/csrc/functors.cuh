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

// Separable quadratic f = sum 0.5 c_i (x_i - a_i)^2  (twin of oracle `SeparableQuadratic`).
// c == nullptr: the GENERATED problem — c_i = 1 + (hash(7, i) & 255) / 16, a_i = int16(hash(8, i)) 2^-14 are recomputed
// from the integer hash instead of being read (two splitmix64 per coordinate are free next to an HBM-bound stream, and
// at n = 2^28 the two 2 GiB coefficient vectors were a third of the traffic of a trial step); index0 is the global
// index of local coordinate 0 (index-range sharding).
struct SepQuadFn {
  static constexpr int BS = 1;
  const double* c;
  const double* a;
  int64_t index0;
  HD double block(int64_t i0, const double* xb, double* gb) const {
    double ci, ai;
    if (c != nullptr) {
      ci = c[i0];
      ai = a[i0];
    } else {
      const uint64_t gi = (uint64_t)(index0 + i0);
      ci = 1.0 + (double)(hash3(7, gi, 0) & 0xFF) / 16.0;
      ai = (double)h16(8, gi, 0) * 6.103515625e-05;  // 2^-14
    }
    const double dlt = xb[0] - ai;
    const double cd = ci * dlt;
    gb[0] = cd;
    return 0.5 * (cd * dlt);
  }
};

}  // namespace osb
