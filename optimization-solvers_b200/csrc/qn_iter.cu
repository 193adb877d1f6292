// qn_iter.cu — whole outer iterations of dense BFGS / DFP in ONE cooperative kernel.
//
// LineSearchSolver::minimize (src/ls_solver.rs:78-107) for the packed-symmetric lazy schedule, with nothing but grid
// barriers between the steps of an iteration (no launch boundary, no host round trip, no cluster-only head):
//
//   epilogue of the previous pass   y.h, s.g, h.g -> coefficients, u = H+ g (O(n), distributed over all CTAs)
//   has_converged                   s_norm / y_norm / ||g||_2                                       bfgs.rs:64-76
//   compute_direction               d = -u | P(x - u) - x                                           bfgs.rs:47, bfgs_b.rs:72-75
//   compute_step_len                the complete line search (ls_automaton.cuh); BackTracking evaluates its first 16
//                                   trial steps in one sweep and ONE grid reduction                 backtracking.rs, morethuente.rs
//   update_next_iterate             x+ = x + t d, oracle(x+), s, y, norms, y.s, k += 1             bfgs.rs:94-112
//   H pass                          pending rank-2 update + h = H y + w = H g over the packed triangle (qn_sym.cuh)
//   fold + exchange                 column partials -> h, w; on several GPUs every CTA pushes ITS chunk of the rank's
//                                   contribution straight into every peer's slot over NVLink and waits for the same
//                                   chunk of every peer (per-chunk flags: no rank-wide barrier, no NCCL call)
//
// Every CTA owns one contiguous chunk of the O(n) vectors (n / grid elements: 112 at n = 16384 on 148 SMs) for all
// O(n) work, so the line search runs on every SM instead of on one cluster, and a grid-wide sum is: CTA partial ->
// global, grid barrier, every CTA adds the partials in CTA order (same bits in every CTA and on every rank).
// Fixed cost per iteration: four grid barriers (~1.5 us each) + the fold, against three launches, a 20 us cluster head
// and a 13-38 us fold + exchange kernel before.
#include <cooperative_groups.h>

#include "engine.cuh"
#include "functors.cuh"
#include "qn_sym.cuh"

namespace osb {

namespace cg = cooperative_groups;

constexpr int IT_NT = 512;     // threads per CTA, one CTA per SM
constexpr int IT_NW = IT_NT / 32;
constexpr int IT_NSPEC = 16;   // backtracking trials per grid reduction
constexpr int IT_GPK = 64;     // doubles per CTA in one grid-reduction buffer
constexpr int IT_MAXG = 256;   // largest grid

struct IterSmem {
  double w[IT_GPK * IT_NW];
  double res[IT_GPK];
  double2 part[2][IT_NT / 2];
  int ext[IT_MAXG];
};

// grid-wide sum of K values; result in every thread of every CTA, identical bits everywhere
template <int K>
__device__ __forceinline__ void grid_sum(double (&acc)[K], const QNIterArgs& a, int& gbuf, IterSmem& sm, cg::grid_group& grid) {
  static_assert(K <= IT_GPK, "grid reduction buffer too small");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (K <= 4) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double v = warp_sum(acc[k]);
      if (lane == 0) sm.w[k * IT_NW + warp] = v;
    }
  } else {
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 16) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (k0 + i < K) ? acc[k0 + i] : 0.0;
      const double r = warp_sum16(v);  // lane l holds the warp sum of value l >> 1
      if ((lane & 1) == 0 && k0 + (lane >> 1) < K) sm.w[(k0 + (lane >> 1)) * IT_NW + warp] = r;
    }
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double v = 0.0;
#pragma unroll
    for (int wq = 0; wq < IT_NW; ++wq) v = v + sm.w[threadIdx.x * IT_NW + wq];
    a.gpart[((size_t)gbuf * gridDim.x + blockIdx.x) * IT_GPK + threadIdx.x] = v;
  }
  grid.sync();
  const double* base = a.gpart + (size_t)gbuf * gridDim.x * IT_GPK;
  for (int k = warp; k < K; k += IT_NW) {
    double v = 0.0;
    for (int c = lane; c < (int)gridDim.x; c += 32) v = v + __ldcg(base + (size_t)c * IT_GPK + k);
    v = warp_sum(v);
    if (lane == 0) sm.res[k] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = sm.res[k];
  gbuf ^= 1;
}

__device__ __forceinline__ double grid_min(double v, const QNIterArgs& a, int& gbuf, IterSmem& sm, cg::grid_group& grid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_min(v);
  if (lane == 0) sm.w[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = INFINITY;
    for (int wq = 0; wq < IT_NW; ++wq) m = fmin(m, sm.w[wq]);
    a.gpart[((size_t)gbuf * gridDim.x + blockIdx.x) * IT_GPK] = m;
  }
  grid.sync();
  const double* base = a.gpart + (size_t)gbuf * gridDim.x * IT_GPK;
  if (warp == 0) {
    double m = INFINITY;
    for (int c = lane; c < (int)gridDim.x; c += 32) m = fmin(m, __ldcg(base + (size_t)c * IT_GPK));
    m = warp_min(m);
    if (lane == 0) sm.res[0] = m;
  }
  __syncthreads();
  const double r = sm.res[0];
  gbuf ^= 1;
  return r;
}

// the streaming pass, out of line: its 120-register loop is compiled once per sharding mode and does not share its
// register allocation with the head code around it
template <bool SHARDED>
__device__ __noinline__ void iter_pass(const QNLazyArgs& la, const QNSymArgs& sa, double c0, double c1, double c2) {
  sym_pass_body<-1, SHARDED, IT_NT, false, false>(la, sa, c0, c1, c2, 0, (int)gridDim.x, (int)blockIdx.x);
}

__device__ __forceinline__ unsigned long long* iter_flags(double* region, int world) {
  return reinterpret_cast<unsigned long long*>(region + xflag2_off(world));
}

// h, w of the fold: columns [j0, j0 + cw) of this CTA.  One GPU: written in place over the row sums.  Sharded: this
// rank's contribution goes into slot `rank` of every rank's exchange region, then the chunk flags.
template <bool SHARDED>
__device__ __forceinline__ void iter_fold(const QNIterArgs& a, IterSmem& sm, int64_t j0, int cw, unsigned long long seq, DevState* st) {
  const int G = (int)gridDim.x;
  const int cwp = cw / 2;
  int cwpp = 1;
  while (cwpp < cwp) cwpp <<= 1;  // pairs per vector, padded to a power of two (<= 256)
  const int ng = IT_NT / (2 * cwpp) > 8 ? 8 : IT_NT / (2 * cwpp);  // groups of partial vectors summed in parallel
  const int tid = threadIdx.x;
  const int pr = tid % cwpp, vec = (tid / cwpp) % 2, grp = tid / (2 * cwpp);
  const int64_t j = j0 + 2 * pr;
  const bool active = pr < cwp && j < a.n && grp < ng;
  const int per = (G + ng - 1) / ng;
  const int cb = grp * per, ce = cb + per < G ? cb + per : G;
  double2 acc = make_double2(0.0, 0.0);
  if (active) {
    const double* src = a.colpart + (int64_t)vec * a.ld + j;
    int c = cb;
    for (; c + 8 <= ce; c += 8) {
      double2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v[k] = j < sm.ext[c + k] ? __ldcg(reinterpret_cast<const double2*>(src + (int64_t)(c + k) * 2 * a.ld)) : make_double2(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc.x = acc.x + v[k].x;
        acc.y = acc.y + v[k].y;
      }
    }
    for (; c < ce; ++c) {
      const double2 v = j < sm.ext[c] ? __ldcg(reinterpret_cast<const double2*>(src + (int64_t)c * 2 * a.ld)) : make_double2(0.0, 0.0);
      acc.x = acc.x + v.x;
      acc.y = acc.y + v.y;
    }
  }
  __syncthreads();
  if (grp < ng) sm.part[vec][grp * cwpp + pr] = acc;
  __syncthreads();
  const int par = (int)(seq & 1ULL);
  if (active && grp == 0) {
    double2 tot = sm.part[vec][pr];
    for (int q = 1; q < ng; ++q) {
      tot.x = tot.x + sm.part[vec][q * cwpp + pr].x;
      tot.y = tot.y + sm.part[vec][q * cwpp + pr].y;
    }
    double* rowarr = vec == 0 ? a.h : a.w;
    bool local = true;
    if (SHARDED) {  // the row sums of rows j, j + 1 exist on this rank only when it owns their tile
      const int64_t T = (a.n + QN_R - 1) / QN_R, tile = j / QN_R;
      const int64_t pairi = tile < T / 2 ? tile : T - 1 - tile;
      local = (pairi % a.world) == a.rank;
    }
    double2 rs = make_double2(0.0, 0.0);
    if (local) rs = __ldcg(reinterpret_cast<const double2*>(rowarr + j));
    const double vx = rs.x + tot.x;
    const double vy = (j + 1 < a.n) ? rs.y + tot.y : 0.0;
    if (SHARDED) {
      const int64_t off = XSLOT_OFF + ((int64_t)(par * a.world + a.rank) * 2 + vec) * XSLOT_LD + j;
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<double2*>(a.peers[p] + off) = make_double2(vx, vy);
      __threadfence_system();  // this thread's peer stores are performed before the flag below is raised
    } else {
      *reinterpret_cast<double2*>(rowarr + j) = make_double2(vx, vy);
    }
  }
  if (SHARDED) {
    __syncthreads();
    if (tid < a.world) {
      st_release_sys(iter_flags(a.peers[tid], a.world) + (int64_t)a.rank * XFLAG2_LD + blockIdx.x, seq);
      if (!wait_flag_sys(iter_flags(a.peers[a.rank], a.world) + (int64_t)tid * XFLAG2_LD + blockIdx.x, seq)) peer_timeout(st);
    }
  }
  __syncthreads();
}

template <class Fn, bool BOUNDED, bool BT, bool SHARDED>
__global__ void __launch_bounds__(IT_NT, 1) qn_iter_kernel(const __grid_constant__ QNIterArgs a, const Fn fn) {
  constexpr int BS = Fn::BS;
  cg::grid_group grid = cg::this_grid();
  __shared__ IterSmem sm;
  DevState* st = a.st;
  if (st->done) return;
  const int tid = threadIdx.x, G = (int)gridDim.x, cta = (int)blockIdx.x;
  const bool leader = cta == 0 && tid == 0;
  const int64_t n = a.n;
  // this CTA's chunk of every O(n) vector: cw elements (even, a multiple of the functor block), one block per thread
  const int64_t nb = (n + BS - 1) / BS;
  int64_t bpc = (nb + G - 1) / G;
  if ((bpc * BS) & 1) bpc += 1;
  const int cw = (int)(bpc * BS);
  const int64_t j0 = (int64_t)cta * cw;
  const int64_t i0 = j0 + (int64_t)tid * BS;          // first element of this thread's block
  const bool own = tid < bpc && i0 + BS <= n;         // (n is a multiple of BS for every functor)
  // ---- state carried across the iterations of this launch (identical in every thread of every CTA)
  double f0 = st->f;
  int has_s = st->has_s, has_y = st->has_y;
  double s_norm = st->s_norm, y_norm = st->y_norm, ys_prev = st->ys, yh = st->yh;
  int skip_prev = st->skip, pending = st->pending, epi_owed = st->epi;
  double pc0 = st->pc0, pc1 = st->pc1, pc2 = st->pc2;
  double cc0 = st->c0, cc1 = st->c1, cc2 = st->c2;
  long long k = st->k;
  int ls_evals = st->ls_evals;
  double t_last = st->t_last, gd0_last = st->gd0;
  double ss = st->ss, yy = st->yy;
  unsigned long long seq = SHARDED ? *a.seq : 0ULL;
  int status = st->status, reason = st->reason, done = 0;
  LSParams p = *a.lsp;
  {
    const int64_t T = (n + QN_R - 1) / QN_R;
    for (int c = tid; c < G; c += IT_NT) sm.ext[c] = (int)sym_first_row<SHARDED>(T, a.world, a.rank, G, c);
  }
  int gbuf = 0;
  // every CTA has read the entry state before anybody can write it
  grid.sync();
  QNLazyArgs la{};
  la.ps = a.ps;
  la.ph = a.ph;
  la.y = a.y;
  la.g = a.g;
  la.h = a.h;
  la.w = a.w;
  la.st = st;
  QNSymArgs sa{a.P, a.P, a.colpart, n, a.ld, SHARDED ? a.world : 1, SHARDED ? a.rank : 0, a.peers, a.seq, G, 0};
  long long t_head = 0, t_pass = 0, t_fold = 0, t_mark = 0;
  const bool timing = a.prof != nullptr && leader;
  auto stamp = [&]() -> long long {
    long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
  };
  if (timing) t_mark = stamp();
  int it = 0;
  for (;; ++it) {
    const bool epi_only = a.epi_only != 0;
    if (!epi_only && it >= a.iters) break;
    if (!epi_only && is_bad(f0)) {  // ls_solver.rs:37-40
      done = 1;
      status = OSB_OUT_OF_DOMAIN;
      break;
    }
    // ---- the epilogue the previous pass left owed: coefficients of the update that becomes pending, u = H+ g
    double ub_[BS];
    if (epi_owed) {
      const int par = epi_owed - 1;  // sharded: parity of the exchange buffers holding h, w
      double hb[BS], wb[BS], sb[BS];
      double e3[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int jq = 0; jq < BS; ++jq) {
        hb[jq] = wb[jq] = sb[jq] = 0.0;
        if (own) {
          const int64_t i = i0 + jq;
          if (SHARDED) {  // rank-ordered sum of the per-rank slots: the same numbers in the same order on every rank
            const double* bh = a.peers[a.rank] + XSLOT_OFF + ((int64_t)(par * a.world) * 2 + 0) * XSLOT_LD + i;
            double vh = __ldcg(bh), vw = __ldcg(bh + XSLOT_LD);
            for (int r = 1; r < a.world; ++r) {
              vh = vh + __ldcg(bh + (int64_t)r * 2 * XSLOT_LD);
              vw = vw + __ldcg(bh + (int64_t)r * 2 * XSLOT_LD + XSLOT_LD);
            }
            hb[jq] = vh;
            wb[jq] = vw;
          } else {
            hb[jq] = __ldcg(a.h + i);
            wb[jq] = __ldcg(a.w + i);
          }
          sb[jq] = __ldcg(a.s + i);
          if (!skip_prev) {
            const double yi = __ldcg(a.y + i), gi = __ldcg(a.g + i);
            e3[0] = fma(yi, hb[jq], e3[0]);     // y.h
            e3[1] = fma(sb[jq], gi, e3[1]);     // s.g
            e3[2] = fma(hb[jq], gi, e3[2]);     // h.g
          }
        }
      }
      double ca = 0.0, cb = 0.0;
      if (!skip_prev) {
        grid_sum<3>(e3, a, gbuf, sm, grid);
        yh = e3[0];
        const double sg = e3[1], hg = e3[2];
        if (a.kind == QN_BFGS) {  // bfgs.rs:115-124, no curvature safeguard
          const double rho = 1.0 / ys_prev;
          cc0 = rho * rho * yh + rho;
          cc1 = -rho;
          cc2 = 0.0;
        } else {  // dfp.rs:115-120
          cc0 = 1.0 / ys_prev;
          cc1 = 0.0;
          cc2 = -1.0 / yh;
        }
        ca = cc0 * sg + cc1 * hg;
        cb = cc1 * sg + cc2 * hg;
        pc0 = cc0;
        pc1 = cc1;
        pc2 = cc2;
        pending = 1;
      } else {  // bfgs.rs:106-112: no new update; the stored matrix is exact once the pending one is applied
        pending = 0;
        pc0 = pc1 = pc2 = 0.0;
      }
#pragma unroll
      for (int jq = 0; jq < BS; ++jq) {
        ub_[jq] = 0.0;
        if (own) {
          const int64_t i = i0 + jq;
          if (skip_prev) {
            ub_[jq] = wb[jq];
          } else {
            ub_[jq] = wb[jq] + (sb[jq] * ca + hb[jq] * cb);
            a.ps[i] = sb[jq];
            a.ph[i] = hb[jq];
          }
          a.u[i] = ub_[jq];
        }
      }
      epi_owed = 0;
    } else {
#pragma unroll
      for (int jq = 0; jq < BS; ++jq) ub_[jq] = own ? __ldcg(a.u + i0 + jq) : 0.0;
    }
    if (epi_only) break;
    // ---- has_converged (bfgs.rs:64-76), first two tests
    if (has_s && s_norm < a.tol) {
      done = 1;
      status = OSB_OK;
      reason = OSB_REASON_S_NORM;
      break;
    }
    if (has_y && y_norm < a.tol) {
      done = 1;
      status = OSB_OK;
      reason = OSB_REASON_Y_NORM;
      break;
    }
    // ---- compute_direction, ||g||^2, g.d — and, for BackTracking, the first IT_NSPEC trial steps in the same sweep
    double xb[BS], db[BS], gb0[BS];
    double gg = 0.0, gd = 0.0, tm = INFINITY;
    const bool need_tmax = !BT && p.kind == LS_MORETHUENTE_B;
#pragma unroll
    for (int jq = 0; jq < BS; ++jq) {
      xb[jq] = db[jq] = gb0[jq] = 0.0;
      if (own) {
        const int64_t i = i0 + jq;
        const double xi = __ldcg(a.x + i), gi = __ldcg(a.g + i);
        double di;
        if (BOUNDED) di = fmin(fmax(xi - ub_[jq], a.lb[i]), a.ub[i]) - xi;  // bfgs_b.rs:72-75
        else di = -ub_[jq];                                                 // bfgs.rs:47
        xb[jq] = xi;
        db[jq] = di;
        gb0[jq] = gi;
        gg = gg + gi * gi;
        gd = gd + gi * di;
        if (need_tmax) {  // morethuente_b.rs:185-197
          double cand;
          if (di > 0.0) cand = (a.ls_ub[i] - xi) / di;
          else if (di < 0.0) cand = (a.ls_lb[i] - xi) / di;
          else cand = INFINITY;
          tm = fmin(cand, tm);
        }
      }
    }
    LSMachine m;
    int evals = 0;
    double gd0;
    if (BT) {
      // backtracking visits 1, beta, beta^2, ... whatever the outcome of a trial (backtracking.rs:37-55): the first
      // IT_NSPEC steps are evaluated before g.d is even known, and fed to the automaton in order afterwards
      double ts[IT_NSPEC];
      ts[0] = 1.0;
#pragma unroll
      for (int q = 1; q < IT_NSPEC; ++q) ts[q] = ts[q - 1] * p.beta;
      double aS[2 + 3 * IT_NSPEC];
      aS[0] = gg;
      aS[1] = gd;
#pragma unroll
      for (int q = 0; q < 3 * IT_NSPEC; ++q) aS[2 + q] = 0.0;
      bool first_round = true;
      for (;;) {
        if (own) {
#pragma unroll
          for (int q = 0; q < IT_NSPEC; ++q) {
            double xt[BS], gt[BS];
#pragma unroll
            for (int jq = 0; jq < BS; ++jq) {
              const double td = ts[q] * db[jq];
              xt[jq] = xb[jq] + td;
              const double df = xt[jq] - xb[jq];
              aS[2 + 3 * q + 2] = aS[2 + 3 * q + 2] + df * df;
            }
            const double fb = fn.block(i0, xt, gt);
#pragma unroll
            for (int jq = 0; jq < BS; ++jq) aS[2 + 3 * q + 1] = aS[2 + 3 * q + 1] + gt[jq] * db[jq];
            aS[2 + 3 * q] = aS[2 + 3 * q] + fb;
          }
        }
        grid_sum<2 + 3 * IT_NSPEC>(aS, a, gbuf, sm, grid);
        if (first_round) {
          if (sqrt(aS[0]) < a.tol) {  // bfgs.rs:74
            done = 1;
            status = OSB_OK;
            reason = OSB_REASON_GRAD_TOL;
            break;
          }
          gd0 = aS[1];
          m.template begin<LS_BACKTRACKING>(p, f0, gd0, a.max_ls, INFINITY);
          first_round = false;
        }
#pragma unroll
        for (int q = 0; q < IT_NSPEC; ++q) {
          if (!m.done && m.request(p) == ts[q]) {
            m.template feed<LS_BACKTRACKING>(p, aS[2 + 3 * q], aS[2 + 3 * q + 1], aS[2 + 3 * q + 2]);
            ++evals;
          }
        }
        if (m.done) break;
        ts[0] = m.request(p);
#pragma unroll
        for (int q = 1; q < IT_NSPEC; ++q) ts[q] = ts[q - 1] * p.beta;
        aS[0] = aS[1] = 0.0;
#pragma unroll
        for (int q = 0; q < 3 * IT_NSPEC; ++q) aS[2 + q] = 0.0;
      }
      if (done) break;
    } else {
      double a2[2] = {gg, gd};
      grid_sum<2>(a2, a, gbuf, sm, grid);
      if (sqrt(a2[0]) < a.tol) {  // bfgs.rs:74
        done = 1;
        status = OSB_OK;
        reason = OSB_REASON_GRAD_TOL;
        break;
      }
      gd0 = a2[1];
      double tmaxc = INFINITY;
      if (need_tmax) tmaxc = grid_min(tm, a, gbuf, sm, grid);
      m.begin(p, f0, gd0, a.max_ls, tmaxc);
      while (!m.done) {
        const double t = m.request(p);
        const bool proj = m.wants_projection(p);
        double a3[3] = {0.0, 0.0, 0.0};
        if (own) {
          double xt[BS], gt[BS];
#pragma unroll
          for (int jq = 0; jq < BS; ++jq) {
            const double td = t * db[jq];
            double v = xb[jq] + td;
            if (proj) v = fmin(fmax(v, a.ls_lb[i0 + jq]), a.ls_ub[i0 + jq]);  // backtracking_b.rs:65-67
            xt[jq] = v;
            const double df = v - xb[jq];
            a3[2] = a3[2] + df * df;
          }
          const double fb = fn.block(i0, xt, gt);
#pragma unroll
          for (int jq = 0; jq < BS; ++jq) a3[1] = a3[1] + gt[jq] * db[jq];
          a3[0] = a3[0] + fb;
        }
        grid_sum<3>(a3, a, gbuf, sm, grid);
        m.feed(p, a3[0], a3[1], a3[2]);
        ++evals;
      }
    }
    const double t = m.result;
    // ---- next = x + t d (ls_solver.rs:60); oracle(next) (bfgs.rs:98); s, y, norms, y.s
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
    if (own) {
      double xn[BS], gn[BS];
#pragma unroll
      for (int jq = 0; jq < BS; ++jq) {
        const double td = t * db[jq];
        xn[jq] = xb[jq] + td;
      }
      const double fb = fn.block(i0, xn, gn);
      a4[3] = fb;
#pragma unroll
      for (int jq = 0; jq < BS; ++jq) {
        const int64_t i = i0 + jq;
        const double si = xn[jq] - xb[jq];
        const double yi = gn[jq] - gb0[jq];
        a.s[i] = si;
        a.y[i] = yi;
        a.x[i] = xn[jq];
        a.g[i] = gn[jq];
        a4[0] = a4[0] + si * si;
        a4[1] = a4[1] + yi * yi;
        a4[2] = a4[2] + yi * si;
      }
    }
    grid_sum<4>(a4, a, gbuf, sm, grid);  // (its grid barrier also publishes s, y, x, g, ps, ph to the pass below)
    ss = a4[0];
    yy = a4[1];
    ys_prev = a4[2];
    f0 = a4[3];
    s_norm = sqrt(ss);
    y_norm = sqrt(yy);
    has_s = has_y = 1;
    skip_prev = (s_norm < a.tol || y_norm < a.tol) ? 1 : 0;  // bfgs.rs:106-112
    t_last = t;
    gd0_last = gd0;
    k += 1;
    ls_evals += evals + 1;
    if (timing) {
      const long long now = stamp();
      t_head += now - t_mark;
      t_mark = now;
    }
    // ---- the H pass: pending update + h = H y + w = H g over (this rank's share of) the packed triangle
    iter_pass<SHARDED>(la, sa, pc0, pc1, pc2);
    grid.sync();
    if (timing) {
      const long long now = stamp();
      t_pass += now - t_mark;
      t_mark = now;
    }
    // ---- fold of the column partials (+ exchange): h, w of this CTA's chunk, complete
    seq += 1ULL;
    iter_fold<SHARDED>(a, sm, j0, cw, seq, st);
    epi_owed = SHARDED ? 1 + (int)(seq & 1ULL) : 1;
    if (timing) {
      const long long now = stamp();
      t_fold += now - t_mark;
      t_mark = now;
    }
  }
  // ---- persist the carried state (one thread; every CTA holds the same values)
  if (leader) {
    st->f = f0;
    st->ft = f0;
    st->gd0 = gd0_last;
    st->ss = ss;
    st->yy = yy;
    st->ys = ys_prev;
    st->yh = yh;
    st->s_norm = s_norm;
    st->y_norm = y_norm;
    st->has_s = has_s;
    st->has_y = has_y;
    st->skip = skip_prev;
    st->c0 = cc0;
    st->c1 = cc1;
    st->c2 = cc2;
    st->pc0 = pc0;
    st->pc1 = pc1;
    st->pc2 = pc2;
    st->pending = pending;
    st->epi = epi_owed;
    st->t_last = t_last;
    st->k = k;
    st->ls_evals = ls_evals;
    if (done) {
      st->done = 1;
      st->status = status;
      st->reason = reason;
    }
    if (SHARDED) *a.seq = seq;
    if (p.kind == LS_GLL || p.kind == LS_MORETHUENTE_B) *a.lsp = p;  // only f_previous / t_max persist across iterations
    if (timing) {
      a.prof[0] += t_head;
      a.prof[1] += t_pass;
      a.prof[2] += t_fold;
      a.prof[3] += it;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
int qn_iter_grid(Ctx* ctx) { return ctx->num_sms < IT_MAXG ? ctx->num_sms : IT_MAXG; }
int64_t qn_iter_gpart_doubles(Ctx* ctx) { return 2 * (int64_t)qn_iter_grid(ctx) * IT_GPK; }

bool qn_iter_supported(Ctx* ctx, int functor_kind, int64_t n, int world) {
  if (functor_kind != FN_ROSENBROCK && functor_kind != FN_SEPQUAD) return false;
  if (n <= QN_SMALL_N) return false;
  const int G = qn_iter_grid(ctx);
  const int bs = functor_kind == FN_ROSENBROCK ? 2 : 1;
  if (n % bs != 0) return false;
  int64_t bpc = ((n + bs - 1) / bs + G - 1) / G;
  if ((bpc * bs) & 1) bpc += 1;
  if (bpc * bs > IT_NT || bpc > IT_NT) return false;  // one functor block per thread, one column pair per fold thread
  if (world > 1 && (n > XSLOT_LD || G > XFLAG2_LD)) return false;
  static int coop = -1;
  if (coop < 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, ctx->device);
    coop = v;
  }
  return coop != 0;
}

template <class Fn, bool BOUNDED, bool BT, bool SHARDED>
static void launch_iter_k(Ctx* ctx, const QNIterArgs& a, const Fn& fn) {
  auto kern = qn_iter_kernel<Fn, BOUNDED, BT, SHARDED>;
  QNIterArgs aa = a;
  Fn f = fn;
  void* params[] = {(void*)&aa, (void*)&f};
  OSB_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(qn_iter_grid(ctx)), dim3(IT_NT), params, 0, ctx->stream));
  ctx->counters[0]++;
}

template <class Fn>
static void launch_iter_fn(Ctx* ctx, const QNIterArgs& a, const Fn& fn, bool bounded, bool bt, bool sharded) {
#define OSB_IT(B, T, S) launch_iter_k<Fn, B, T, S>(ctx, a, fn)
  if (bounded) {
    if (bt) { if (sharded) OSB_IT(true, true, true); else OSB_IT(true, true, false); }
    else { if (sharded) OSB_IT(true, false, true); else OSB_IT(true, false, false); }
  } else {
    if (bt) { if (sharded) OSB_IT(false, true, true); else OSB_IT(false, true, false); }
    else { if (sharded) OSB_IT(false, false, true); else OSB_IT(false, false, false); }
  }
#undef OSB_IT
}

void qn_launch_iter(Ctx* ctx, int functor_kind, const double* fn_a, const double* fn_b, bool bounded, int ls_kind, const QNIterArgs& a) {
  const bool bt = ls_kind == LS_BACKTRACKING;
  const bool sharded = a.world > 1;
  if (functor_kind == FN_ROSENBROCK) {
    launch_iter_fn(ctx, a, RosenbrockFn{}, bounded, bt, sharded);
  } else if (functor_kind == FN_SEPQUAD) {
    SepQuadFn fn;
    fn.c = fn_a;
    fn.a = fn_b;
    launch_iter_fn(ctx, a, fn, bounded, bt, sharded);
  } else {
    throw Error(OSB_ERR_UNSUPPORTED, "objective has no block functor for the device-resident engine");
  }
}

}  // namespace osb
