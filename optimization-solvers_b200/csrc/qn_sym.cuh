// qn_sym.cuh — packed symmetric storage of the inverse-Hessian approximation: layout, the streaming pass body and the
// peer-memory synchronisation helpers, shared by qn_kernels.cu (one launch per phase) and qn_iter.cu (one cooperative
// kernel per outer iteration).
#pragma once
#include "engine.cuh"

namespace osb {

constexpr int QN_T = 512;          // threads per CTA
constexpr int QN_R = 8;            // rows per tile
constexpr int QN_CHUNK = QN_T * 2; // columns per sweep step (one 16-byte vector per thread)


__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Bounded wait for a peer's flag: a dead or diverged peer must not wedge this GPU.  After ~5 s without progress the
// run is terminated (DevState.done, status AbnormalTermination); every later launch of the solve is then a no-op.
__device__ __forceinline__ bool wait_flag_sys(const unsigned long long* p, unsigned long long seq) {
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  while (ld_acquire_sys(p) < seq) {
    if ((++spins & 1023u) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 5000000000ULL) return false;
    }
  }
  return true;
}
__device__ __forceinline__ void peer_timeout(DevState* st) {
  st->done = 1;
  st->status = OSB_ABNORMAL_TERMINATION;
}


__host__ __device__ inline int64_t sym_lpad(int64_t tile) { return 16 * (tile / 2) + 16; }  // roundup16(8 tile + 8)
__host__ __device__ inline int64_t sym_tile_offset(int64_t tile) {                          // doubles before tile
  const int64_t m = tile / 2;
  int64_t cnt = 16 * m * (m + 1);      // sum of lpad over tiles < 2m
  if (tile & 1) cnt += 16 * (m + 1);
  return 8 * cnt;
}

// Sharded layout (world > 1, ntiles even): tile PAIRS (p, T-1-p) are dealt round-robin over the ranks (pair p belongs
// to rank p % world) and stored pair by pair; lpad(p) + lpad(T-1-p) = 8 T + 16 for every p when T is even, so the
// local offset is a closed form: local pair index * 8 (8 T + 16), the short tile p first.
__host__ __device__ inline int64_t symsh_pair_doubles(int64_t T) { return 8 * (8 * T + 16); }
__host__ __device__ inline int64_t symsh_tile_offset(int64_t tile, int64_t T, int world) {
  const int64_t half = T / 2;
  const int64_t pairi = tile < half ? tile : T - 1 - tile;
  return (pairi / world) * symsh_pair_doubles(T) + (tile < half ? 0 : 8 * sym_lpad(pairi));
}
__host__ __device__ inline int64_t symsh_local_pairs(int64_t T, int world, int rank) {
  const int64_t half = T / 2;
  return half > rank ? (half - rank + world - 1) / world : 0;
}

struct QNSymArgs {
  double* P;         // packed matrix (this rank's tiles when sharded)
  double* Pout;      // ping-pong variant: the pass reads P and writes Pout (== P in place)
  double* colpart;   // gridDim x 2 x ld per-CTA column partials (h then w)
  int64_t n, ld;
  int world, rank;   // sharded: pairs p = rank, rank + world, ...
  double* const* peers;        // exchange regions (fold kernel, sharded)
  unsigned long long* seq;     // exchange sequence number
  int pass_grid;     // grid of the streaming pass (= number of column-partial vectors)
  int zeroed;        // 1: the pass zeroes its partial vector first (legacy variant); 0: the first (longest) tile of a CTA
                     //    WRITES the partial, columns at or beyond that tile's first row are never read (sym_first_row)
};

// 16 per-thread values -> warp sums with 16 double shuffles (recursive halving) instead of 80; after the call
// lane l holds the warp sum of value l >> 1.
__device__ __forceinline__ double warp_sum16(double (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int hs = 8, o = 16; hs >= 1; hs >>= 1, o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < hs; ++i) {
      const double send = upper ? v[i] : v[i + hs];
      const double keep = upper ? v[i + hs] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Static assignment of tiles to the CTAs of the pass (deterministic: a row's sums are formed inside one CTA, the column
// partials are folded in CTA order).
//   one GPU: tile t has 8 (t + 1) stored columns; tiles are dealt in PAIRS (T-1-p, p) of equal total length, pair p to
//            CTA p % grid, the LONG tile of a pair first (plain round-robin leaves the CTA with the longest tiles 7 % above
//            the average);
//   sharded: the rank owns the pairs p = rank, rank + world, ... = 2 nlp tiles; they are dealt ONE BY ONE in decreasing
//            length in snake order (round k left-to-right for even k, right-to-left for odd k): position q < nlp is tile
//            T-1-(rank + q world), position q >= nlp is tile rank + (2 nlp - 1 - q) world (whole pairs would quantise
//            the work to ceil(nlp / grid) rounds: 4 instead of 3.46 at 2 GPUs).
// Either way the FIRST tile a CTA processes is its longest one: its column contributions initialise the CTA's partial
// vector (no zeroing pass, no 39 MB of zeros through L2), and every later tile adds into a prefix of it.
template <bool SHARDED>
__host__ __device__ inline int64_t sym_cta_tile(int64_t ntiles, int world, int rank, int grid, int cta, int64_t step) {
  if (SHARDED) {
    const int64_t nlp = symsh_local_pairs(ntiles, world, rank);
    const int64_t pos = (step & 1) ? (int64_t)grid - 1 - cta : cta;
    const int64_t uq = step * grid + pos;
    if (uq >= 2 * nlp) return -1;
    const int64_t pairi = uq < nlp ? rank + uq * world : rank + (2 * nlp - 1 - uq) * world;
    return uq < nlp ? ntiles - 1 - pairi : pairi;
  } else {
    const int64_t nhalf = (ntiles + 1) / 2;
    const int64_t pairi = (step >> 1) * grid + cta;
    if (pairi >= nhalf) return -1;
    const int64_t longt = ntiles - 1 - pairi;
    if ((step & 1) == 0) return longt;
    return longt == pairi ? -1 : pairi;  // odd tile count: the middle tile only once
  }
}
// first row of the first tile of CTA `cta`: the columns [0, that) of its partial vector are valid after the pass
template <bool SHARDED>
__host__ __device__ inline int64_t sym_first_row(int64_t ntiles, int world, int rank, int grid, int cta) {
  const int64_t t = sym_cta_tile<SHARDED>(ntiles, world, rank, grid, cta, 0);
  return t < 0 ? 0 : t * QN_R;
}

// SHARDED is a template parameter so that the single-GPU instantiation keeps exactly its own loop structure (the
// 128-register streaming loop is sensitive to anything that stays live across it).  NT = threads per CTA: 512 (one
// CTA per SM) or 256 (two independent CTAs per SM: one streams while the other drains into its tile-end reduction).
// OOP: ping-pong storage, the pass reads sa.P and writes sa.Pout.  ZERO: legacy zero-first partials.
// IDENT: the stored matrix is the identity and has NOT been written yet (H_0 = I, bfgs.rs:30-33): the elements are
// generated instead of loaded, so the first pass of a solve writes the triangle once instead of memset + read + write.
template <int KIND, bool SHARDED, int NT, bool OOP, bool ZERO, bool IDENT = false>
__device__ __forceinline__ void sym_pass_body(const QNLazyArgs& a, const QNSymArgs& sa, const double c0, const double c1, const double c2,
                                              const int pp, const int grid, const int cta) {
  const unsigned long long pol = l2_evict_first_policy();
  __shared__ double red2[2][NT / 32][16];  // double-buffered by tile parity: two barriers per tile instead of four
  __shared__ double4 rowv2[2][QN_R];       // p_i, q_i, y_i, g_i of the tile's rows
  int tpar = 0;
  const int64_t n = sa.n, ld = sa.ld;
  const double* __restrict__ p = a.ps;
  const double* __restrict__ q = a.ph;
  const double* __restrict__ yv = a.y;
  const double* __restrict__ gv = a.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* __restrict__ cph = sa.colpart + (int64_t)cta * 2 * ld;
  double* __restrict__ cpw = cph + ld;
  if (ZERO) {
    for (int64_t j = 2 * threadIdx.x; j < ld; j += 2 * NT) {
      *reinterpret_cast<double2*>(cph + j) = make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(cpw + j) = make_double2(0.0, 0.0);
    }
    __syncthreads();
  }
  const int64_t ntiles = (n + QN_R - 1) / QN_R;
  bool first = !ZERO;
  for (int64_t step = 0;; ++step) {
    const int64_t tile = sym_cta_tile<SHARDED>(ntiles, sa.world, sa.rank, grid, cta, step);
    if (tile < 0) {
      if (SHARDED || (step & 1) == 0) break;  // (one GPU, odd step: only the middle tile of an odd count is skipped)
      continue;
    }
    const int64_t r0 = tile * QN_R;
    const int rows_here = (int)((n - r0) < QN_R ? (n - r0) : QN_R);
    const int64_t lpad = sym_lpad(tile);
    const int ncols = (int)(r0 + QN_R < n ? r0 + QN_R : n);  // stored columns: 0 .. ncols-1
    double ah[QN_R], aw[QN_R];
#pragma unroll
    for (int r = 0; r < QN_R; ++r) ah[r] = aw[r] = 0.0;
    tpar ^= 1;
    double4* rowv = rowv2[tpar];
    double (*red)[16] = red2[tpar];
    if (threadIdx.x < QN_R) {
      const bool ok = (int)threadIdx.x < rows_here;
      const int64_t i = r0 + threadIdx.x;
      rowv[threadIdx.x] = ok ? make_double4(p[i], q[i], yv[i], gv[i]) : make_double4(0.0, 0.0, 0.0, 0.0);
    }
    __syncthreads();  // (A) rowv[tpar] visible; also orders the previous tile's red[tpar^1] readers before its next reuse
    const int64_t toff = SHARDED ? symsh_tile_offset(tile, ntiles, sa.world) : sym_tile_offset(tile);
    const double* __restrict__ base = (OOP && pp ? sa.Pout : sa.P) + toff;
    double* __restrict__ obase = (OOP && !pp ? sa.Pout : sa.P) + toff;
    for (int col = 2 * threadIdx.x; col < (int)lpad; col += 2 * NT) {
      // element validity: columns >= ncols are padding (never stored, never updated)
      const bool v0 = col < ncols, v1 = col + 1 < ncols;
      if (!v0) continue;
      const double2 gj = ld_vec2(gv + col);
      const double2 yj = ld_vec2(yv + col);
      const double2 pj = ld_vec2(p + col);
      const double2 qj = ld_vec2(q + col);
      double2 hv[QN_R];
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        if (IDENT) hv[r] = make_double2(col == (int)r0 + r ? 1.0 : 0.0, col + 1 == (int)r0 + r ? 1.0 : 0.0);
        else hv[r] = r < rows_here ? ld_stream_ef(base + r * lpad + col, pol) : make_double2(0.0, 0.0);
      }
      // column contributions only strictly left of the diagonal block (r0 is a multiple of 8 and col is even: the pair
      // (col, col + 1) is on the same side)
      const bool cok = col < (int)r0;
      double ch0 = 0.0, ch1 = 0.0, cw0 = 0.0, cw1 = 0.0;
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        if (r < rows_here) {
          const double4 rv = rowv[r];
          const double pi = rv.x, qi = rv.y;
          double2 hn;
          if (KIND == QN_BFGS) {
            const double cx = pi * qj.x + qi * pj.x, cy = pi * qj.y + qi * pj.y;
            hn.x = fma(c0, pi * pj.x, fma(c1, cx, hv[r].x));
            hn.y = fma(c0, pi * pj.y, fma(c1, cy, hv[r].y));
          } else if (KIND == QN_DFP) {
            hn.x = fma(c2, qi * qj.x, fma(c0, pi * pj.x, hv[r].x));
            hn.y = fma(c2, qi * qj.y, fma(c0, pi * pj.y, hv[r].y));
          } else {
            // both rules in one expression (BFGS: c2 = 0, DFP: c1 = 0): fma(0, finite, v) == v, so each rule keeps its own bits
            const double cx = pi * qj.x + qi * pj.x, cy = pi * qj.y + qi * pj.y;
            hn.x = fma(c2, qi * qj.x, fma(c0, pi * pj.x, fma(c1, cx, hv[r].x)));
            hn.y = fma(c2, qi * qj.y, fma(c0, pi * pj.y, fma(c1, cy, hv[r].y)));
          }
          if (!v1) hn.y = 0.0;
          ah[r] = fma(hn.x, yj.x, ah[r]);
          ah[r] = fma(hn.y, yj.y, ah[r]);
          aw[r] = fma(hn.x, gj.x, aw[r]);
          aw[r] = fma(hn.y, gj.y, aw[r]);
          ch0 = fma(hn.x, rv.z, ch0);
          ch1 = fma(hn.y, rv.z, ch1);
          cw0 = fma(hn.x, rv.w, cw0);
          cw1 = fma(hn.y, rv.w, cw1);
          st_stream_ef(obase + r * lpad + col, hn, pol);
        }
      }
      if (cok) {
        double2 oh = make_double2(0.0, 0.0), ow = make_double2(0.0, 0.0);
        if (!first) {
          oh = *reinterpret_cast<double2*>(cph + col);
          ow = *reinterpret_cast<double2*>(cpw + col);
        }
        oh.x += ch0;
        ow.x += cw0;
        oh.y += ch1;
        ow.y += cw1;
        *reinterpret_cast<double2*>(cph + col) = oh;
        *reinterpret_cast<double2*>(cpw + col) = ow;
      }
    }
    first = false;
    {
      double v16[16];
#pragma unroll
      for (int r = 0; r < QN_R; ++r) {
        v16[r] = ah[r];
        v16[QN_R + r] = aw[r];
      }
      const double ws = warp_sum16(v16);
      if ((lane & 1) == 0) red[warp][lane >> 1] = ws;
    }
    __syncthreads();  // (B)
    if (threadIdx.x < 2 * QN_R) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) v = v + red[w][threadIdx.x];
      const int r = threadIdx.x % QN_R;
      if (r < rows_here) {
        if (threadIdx.x < QN_R) a.h[r0 + r] = v;
        else a.w[r0 + r] = v;
      }
    }
  }
}


}  // namespace osb
