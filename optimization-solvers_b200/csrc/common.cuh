// common.cuh — shared host/device helpers of the sm_100a library.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/optsolv_b200.h"

#define HD __host__ __device__ __forceinline__

namespace osb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define OSB_CUDA(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      throw ::osb::Error(OSB_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                                           __FILE__ + ":" + std::to_string(__LINE__) + ")");           \
  } while (0)

#define OSB_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) throw ::osb::Error((code), (msg)); \
  } while (0)

// Rust f64::max / f64::min drop NaN (number.rs:19, morethuente.rs:290); fmax/fmin have the same
// contract on host and device.
HD double rmax(double a, double b) { return fmax(a, b); }
HD double rmin(double a, double b) { return fmin(a, b); }
HD bool is_bad(double f) { return isnan(f) || isinf(f); }  // ls_solver.rs:37, backtracking.rs:37

// splitmix64-based integer hash: the specification of every synthetic input (DESIGN.md §inputs,
// SURVEY §8d).  Values are small integers times a power of two, so host and device agree bit-for-bit.
HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
HD uint64_t hash3(uint64_t seed, uint64_t i, uint64_t j) { return splitmix64(seed ^ (i * 0x9E3779B97F4A7C15ULL + j)); }
HD int h16(uint64_t seed, uint64_t i, uint64_t j) { return (int)(int16_t)(hash3(seed, i, j) & 0xFFFF); }

#ifdef __CUDACC__
// ---- device-only helpers ------------------------------------------------------------------
// The library is compiled with -fmad=false: a*b+c stays two roundings everywhere (the reference
// is Rust, which never contracts; the active set is defined by exact == on x + t*d).  Fused
// multiply-adds appear only where written explicitly (dot-product accumulators).
__device__ __forceinline__ double2 ld_stream(const double* p) {  // 128-bit streaming load, no L1 allocation
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream_nc(const double* p) {  // read-only for the kernel's lifetime
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
// O(n) column vectors of the streaming kernels: an ordinary cached load, but volatile so that it keeps its program
// position relative to the (volatile) H loads and stores.  The schedule "all loads of a column step, then the
// arithmetic, then the stores" is what the measured bandwidth depends on; left free, the compiler rotated the loop
// differently from build to build (0.73 ms vs 0.78 ms per pass at n = 16384 for unrelated source changes).
__device__ __forceinline__ double2 ld_vec2(const double* p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(double* p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
// Programmatic dependent launch: a kernel launched with launch_pdl may be scheduled while the previous kernel of the
// stream is still draining; its first statement is pdl_wait(), which returns once that kernel has completed and its
// writes are visible.  What overlaps is the launch latency and the ramp of the CTAs (2-3 us per boundary), three
// boundaries per iteration on the one-launch-per-phase path.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the NEXT kernel of the stream be scheduled as soon as SM resources free up (it still blocks in its pdl_wait until
// this grid has completed).  Only for kernels whose whole grid is resident at once: a waiting dependent must never hold
// the slot an unscheduled CTA of this grid needs.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  OSB_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}

// L2 evict-first policy for the H stream: 4 GB pass through L2 every iteration and would otherwise evict
// what is re-used across launches (the O(n) vectors and, notably, the instruction lines of the small
// head kernel, whose cold start is bound by instruction fetch).
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ld_stream_ef(const double* p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ld_stream_nc_ef(const double* p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_stream_ef(double* p, double2 v, unsigned long long pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace osb
