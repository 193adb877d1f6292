// batched.cu — batched mode: many independent small BFGS problems, one warp per problem.
#include "engine.cuh"
#include "functors.cuh"

namespace osb {

int batched_bfgs_rosenbrock(Ctx*, int64_t, int64_t, const double*, bool, int64_t, double, int64_t, int64_t, double, double,
                            double*, double*, int32_t*, int32_t*, int32_t*, double*) {
  throw Error(OSB_ERR_UNSUPPORTED, "batched BFGS: not built yet");
}

}  // namespace osb
