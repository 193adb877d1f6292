// exp_kernels.cu — access-pattern experiments for the in-place read-modify-write of H (bench hook only).
// They answer one question with measurements: how close to the device-to-device copy rate can an in-place
// RMW stream get on B200, and with which thread-to-address mapping (DESIGN.md "RMW access pattern").
#include "engine.cuh"

namespace osb {

// (a) 1-D grid-stride, 16 B per thread per step, U steps in flight
template <int U>
__global__ void __launch_bounds__(256) exp_rmw_gridstride(double* __restrict__ M, int64_t total2, double f) {
  const unsigned long long pol = l2_evict_first_policy();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += stride * U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < total2) v[u] = ld_stream_ef(M + 2 * (i + u * stride), pol);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < total2) {
        v[u].x = fma(v[u].x, f, 1e-30);
        v[u].y = fma(v[u].y, f, 1e-30);
        st_stream_ef(M + 2 * (i + u * stride), v[u], pol);
      }
  }
}

// (b) each CTA owns a contiguous slab; T threads x 16 B x U per step (a step = T*U*16 B contiguous)
template <int U>
__global__ void __launch_bounds__(512) exp_rmw_slab(double* __restrict__ M, int64_t total2, double f) {
  const unsigned long long pol = l2_evict_first_policy();
  const int64_t per = (total2 + gridDim.x - 1) / gridDim.x;
  const int64_t b = per * blockIdx.x, e = (b + per < total2) ? b + per : total2;
  for (int64_t i = b + threadIdx.x; i < e; i += (int64_t)blockDim.x * U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * blockDim.x < e) v[u] = ld_stream_ef(M + 2 * (i + u * blockDim.x), pol);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * blockDim.x < e) {
        v[u].x = fma(v[u].x, f, 1e-30);
        v[u].y = fma(v[u].y, f, 1e-30);
        st_stream_ef(M + 2 * (i + u * blockDim.x), v[u], pol);
      }
  }
}

// (c) out-of-place variant of (b): read M, write N (what a copy does)
template <int U>
__global__ void __launch_bounds__(512) exp_copy_slab(const double* __restrict__ M, double* __restrict__ N, int64_t total2, double f) {
  const unsigned long long pol = l2_evict_first_policy();
  const int64_t per = (total2 + gridDim.x - 1) / gridDim.x;
  const int64_t b = per * blockIdx.x, e = (b + per < total2) ? b + per : total2;
  for (int64_t i = b + threadIdx.x; i < e; i += (int64_t)blockDim.x * U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * blockDim.x < e) v[u] = ld_stream_nc_ef(M + 2 * (i + u * blockDim.x), pol);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * blockDim.x < e) {
        v[u].x = fma(v[u].x, f, 1e-30);
        v[u].y = fma(v[u].y, f, 1e-30);
        st_stream_ef(N + 2 * (i + u * blockDim.x), v[u], pol);
      }
  }
}

void exp_launch(Ctx* ctx, int which, double* M, double* N, int64_t total_doubles) {
  const int64_t t2 = total_doubles / 2;
  const int sms = ctx->num_sms;
  switch (which) {
    case 10: exp_rmw_gridstride<4><<<sms * 8, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 11: exp_rmw_gridstride<8><<<sms * 8, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 12: exp_rmw_slab<8><<<sms, 512, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 13: exp_rmw_slab<8><<<sms * 2, 512, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 14: exp_rmw_slab<4><<<sms * 4, 512, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 15: exp_copy_slab<8><<<sms * 2, 512, 0, ctx->stream>>>(M, N, t2, 1.0000001); break;
    case 16: exp_rmw_gridstride<8><<<sms * 16, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 17: exp_rmw_slab<16><<<sms, 512, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 18: exp_rmw_gridstride<8><<<sms * 2, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 19: exp_rmw_gridstride<8><<<sms * 4, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 20: exp_rmw_gridstride<16><<<sms * 4, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 21: exp_rmw_gridstride<16><<<sms * 8, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    case 22: exp_rmw_gridstride<4><<<sms * 16, 256, 0, ctx->stream>>>(M, t2, 1.0000001); break;
    default: break;
  }
  ctx->counters[0]++;
}

}  // namespace osb
