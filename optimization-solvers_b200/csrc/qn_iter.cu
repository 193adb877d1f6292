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

int qn_iter_grid(Ctx* ctx);

// State carried across the iterations of one launch: identical in every thread of every CTA (every thread writes the
// same values).  It lives in SHARED memory: not in registers, because nothing should be live across the streaming pass
// (with the scalars in registers the 120-register loop of the pass re-materialised its row addresses every step: 447 us
// per pass instead of 400), and not in local memory, because every grid barrier invalidates L1 (fence -> CCTL.IVALL)
// and each first touch of a stack line after a barrier then costs an L2 / DRAM round trip: with the state on the stack
// the three grid sums of the head took 12 us each instead of 2.
struct IterCarry {
  double f0, s_norm, y_norm, ys_prev, yh;
  double pc0, pc1, pc2, cc0, cc1, cc2;
  double t_last, gd0_last, ss, yy;
  long long k;
  unsigned long long seq;
  int has_s, has_y, skip_prev, pending, epi_owed, ls_evals, status, reason, done, gbuf;
  unsigned int llseq;  // sequence number of the flagged (barrier-free) grid reductions; persists in DevState.ll_seq
  int snap_it;         // host-callback snapshots published so far in this launch (QNIterArgs.snap_*)
  int ds_owed;         // the step sums (s.s, y.y, y.s, f+) of the last iteration are still per-CTA partials (IterSmem.ds_part)
  LSParams p;
};

struct IterSmem {
  IterCarry c;
  double w[IT_GPK * IT_NW];
  double res[IT_GPK];
  double2 part[2][IT_NT / 2];
  double xs[IT_NT], ds[IT_NT];  // this CTA's chunk of x and d (backtracking: one warp per trial step sweeps it)
  int ext[IT_MAXG];
  double ds_part[4];   // this CTA's partials of the deferred step sums (IterCarry.ds_owed)
  long long prof[16];  // leader only: [0..2] accumulated ns in head / pass / fold, [3] last stamp, [4..15] head sub-phases
  int prof_skip;       // iterations left out of the profile (the first one of a launch: it absorbs the ranks' launch skew)
};

// Grid-wide sums, compact on purpose: the head runs once per 400 us pass, i.e. with a cold instruction cache, and a
// cold straight-line body costs ~8 cycles per instruction (round 1 measured 86 us for an 11k-instruction head; the first
// version of this kernel, with the 16 trial steps and a 50-value shuffle reduction fully unrolled, spent 38 us per
// iteration in the head).  Everything here is a rolled loop.
//   step 1 (caller): every warp leaves its partial of value k in sm.w[k * IT_NW + warp]  (warp_put)
//   step 2: thread k adds the IT_NW warp partials in warp order -> this CTA's partial, to global memory
//   step 3: grid barrier; warp w adds the G CTA partials of the values k = w, w + IT_NW, ... in CTA order -> sm.res[k]
// Result in sm.res[0..K), identical bits in every CTA and on every rank.
__device__ __forceinline__ long long iter_stamp();
__device__ __noinline__ void iter_submark(const QNIterArgs& a, IterSmem& sm, int slot) {
  if (a.prof == nullptr || blockIdx.x != 0 || threadIdx.x != 0) return;
  long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  sm.prof[slot] += now - sm.prof[15];
  sm.prof[15] = now;
}
// (`active`: this warp holds data — warps beyond the CTA's chunk contribute an exact zero without the 10 shuffles; with
//  112 elements per CTA only 2 of the 16 warps are active, and 50 values x 16 warps x 10 SHFL was 4 us of the head)
__device__ __forceinline__ void warp_put(IterSmem& sm, int k, double v, bool active = true) {
  if (active) v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm.w[k * IT_NW + (threadIdx.x >> 5)] = active ? v : 0.0;
}
// FLAGGED = false: the sums travel through plain stores + ONE real grid barrier (which also publishes whatever the CTAs
//   wrote before it: needed in front of the H pass).
// FLAGGED = true: a pure reduction needs no barrier at all.  Every value is split in two 32-bit halves, each stored
//   together with the 32-bit sequence number of this reduction in ONE 64-bit word (single-copy atomic); a reader polls
//   the word until the sequence number matches (the LL protocol of NCCL, between the CTAs of one GPU through L2).  Cost:
//   one store-to-load propagation instead of fence + barrier + load (5.6 us -> ~1.5 us per reduction, measured).  Two
//   buffers alternate: a CTA can be at most one reduction ahead of the slowest one, because finishing reduction r means
//   having read every CTA's words of r.
// Ksm: the values [0, Ksm) come from the warp partials in sm.w; the values [Ksm, K) were written by the caller straight
//   into this CTA's row (grid_row).
__device__ __forceinline__ double* grid_row(const QNIterArgs& a, int gbuf) {
  return a.gpart + ((size_t)gbuf * gridDim.x + blockIdx.x) * IT_GPK;
}
__device__ __forceinline__ unsigned long long* ll_row(const QNIterArgs& a, int buf, int cta) {
  return reinterpret_cast<unsigned long long*>(a.gpart + 2 * (size_t)gridDim.x * IT_GPK + IT_MAXG) + ((size_t)buf * gridDim.x + cta) * (2 * IT_GPK);
}
__device__ __forceinline__ void ll_put(unsigned long long* row, int k, double v, unsigned int seq) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (bits & 0xffffffffULL) | ((unsigned long long)seq << 32);
  const unsigned long long w1 = (bits >> 32) | ((unsigned long long)seq << 32);
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(row + 2 * k), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ double ll_get(const unsigned long long* row, int k, unsigned int seq) {
  unsigned long long w0, w1;
  do {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(row + 2 * k) : "memory");
  } while ((unsigned int)(w0 >> 32) != seq || (unsigned int)(w1 >> 32) != seq);
  return __longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32)));
}

template <bool FLAGGED>
__device__ __forceinline__ void grid_reduce(const QNIterArgs& a, IterSmem& sm, int K, int Ksm, int* gbuf_io, unsigned int* llseq_io, bool is_min) {
  const int lane = threadIdx.x & 31;
  const int gbuf = *gbuf_io;
  const unsigned int seq = *llseq_io + 1u;
  __syncthreads();
  if ((int)threadIdx.x < Ksm) {
    double v = is_min ? INFINITY : 0.0;
#pragma unroll 1
    for (int wq = 0; wq < IT_NW; ++wq) {
      const double x = sm.w[threadIdx.x * IT_NW + wq];
      v = is_min ? fmin(v, x) : v + x;
    }
    if (FLAGGED) ll_put(ll_row(a, (int)(seq & 1u), (int)blockIdx.x), (int)threadIdx.x, v, seq);
    else grid_row(a, gbuf)[threadIdx.x] = v;
  }
  if (!FLAGGED) cg::this_grid().sync();
  // 8 lanes per value: lane q of the group adds the partials of the CTAs q, q + 8, ... in that order (loads issued in
  // batches of 8 before the adds), then an 8-lane tree: fixed shape, hence the same bits in every CTA and on every rank
  const double* base = a.gpart + (size_t)gbuf * gridDim.x * IT_GPK;
  const int sub = lane & 7;
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += IT_NT / 8) {
    const int k = k0 + (int)(threadIdx.x >> 3);
    double v = is_min ? INFINITY : 0.0;
    if (k < K) {
#pragma unroll 1
      for (int cq = sub; cq < (int)gridDim.x; cq += 80) {
        double x[10];
        if (FLAGGED) {
          // all ten loads are issued before any flag is looked at; a round is repeated until every word carries this
          // reduction's sequence number (polling one word after the other serialised ten L2 round trips)
          unsigned long long w0[10], w1[10];
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int e = 0; e < 10; ++e) {
              w0[e] = w1[e] = (unsigned long long)seq << 32;
              if (cq + 8 * e < (int)gridDim.x)
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[e]), "=l"(w1[e]) : "l"(ll_row(a, (int)(seq & 1u), cq + 8 * e) + 2 * k) : "memory");
            }
#pragma unroll
            for (int e = 0; e < 10; ++e) ok = ok && (unsigned int)(w0[e] >> 32) == seq && (unsigned int)(w1[e] >> 32) == seq;
          } while (!ok);
#pragma unroll
          for (int e = 0; e < 10; ++e)
            x[e] = (cq + 8 * e < (int)gridDim.x) ? __longlong_as_double((long long)((w0[e] & 0xffffffffULL) | (w1[e] << 32))) : (is_min ? INFINITY : 0.0);
        } else {
#pragma unroll
          for (int e = 0; e < 10; ++e)
            x[e] = (cq + 8 * e < (int)gridDim.x) ? __ldcg(base + (size_t)(cq + 8 * e) * IT_GPK + k) : (is_min ? INFINITY : 0.0);
        }
#pragma unroll
        for (int e = 0; e < 10; ++e) v = is_min ? fmin(v, x[e]) : v + x[e];
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_min ? fmin(v, ov) : v + ov;
    }
    if (k < K && sub == 0) sm.res[k] = v;
  }
  __syncthreads();
  if (FLAGGED) *llseq_io = seq;
  else *gbuf_io = gbuf ^ 1;
}

__device__ __forceinline__ unsigned long long* iter_flags(double* region, int world) {
  return reinterpret_cast<unsigned long long*>(region + xflag2_off(world));
}

// h, w of the fold: columns [j0, j0 + cw) of this CTA.  One GPU: written in place over the row sums.  Sharded: this
// rank's contribution goes into slot `rank` of every rank's exchange region, then the chunk flags.
template <int SH>  // 0: one GPU, 1: packed triangle sharded by tile pairs over the ranks
__device__ __forceinline__ void iter_fold(const QNIterArgs& a, IterSmem& sm, int64_t j0, int cw, unsigned long long seq, DevState* st) {
  constexpr bool SHARDED = SH != 0;
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
    double2 rs = make_double2(0.0, 0.0);
    if (SHARDED) {  // the row sums of rows j, j + 1 exist on this rank only when it owns their tile: 
      const int64_t T = (a.n + QN_R - 1) / QN_R, tile = j / QN_R;
      const int64_t pairi = tile < T / 2 ? tile : T - 1 - tile;
      if ((pairi % a.world) == a.rank) rs = __ldcg(reinterpret_cast<const double2*>(rowarr + j));
    } else {
      rs = __ldcg(reinterpret_cast<const double2*>(rowarr + j));
    }
    const double vx = rs.x + tot.x;
    const double vy = (j + 1 < a.n) ? rs.y + tot.y : 0.0;
    if (SHARDED) {
      const int64_t off = XSLOT_OFF + ((int64_t)(par * a.world + a.rank) * 2 + vec) * XSLOT_LD + j;
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<double2*>(a.peers[p] + off) = make_double2(vx, vy);
    } else {
      *reinterpret_cast<double2*>(rowarr + j) = make_double2(vx, vy);
    }
  }
  if (SHARDED) {
    // the CTA barrier orders every thread's peer stores before the flag threads; their release stores (system scope,
    // cumulative) then publish the chunk — one fence per flag thread instead of one per storing thread
    __syncthreads();
    if (tid < a.world) {
      st_release_sys(iter_flags(a.peers[tid], a.world) + (int64_t)a.rank * XFLAG2_LD + blockIdx.x, seq);
      if (!wait_flag_sys(iter_flags(a.peers[a.rank], a.world) + (int64_t)tid * XFLAG2_LD + blockIdx.x, seq)) peer_timeout(st);
    }
  }
  __syncthreads();
}

// scalars of the snapshot (device ring) + its flag (the ONLY store to host memory: SM stores of the 256 KB of x and g to
// pinned memory plus a system fence per CTA cost 45 us per iteration, profiles/r02_callbacks.md; the host copies the
// slot out with a DMA on a side stream once it sees the flag).  One thread of the leader CTA; x and g were stored by
// their owners before the grid barrier.
__device__ __noinline__ void iter_publish(const QNIterArgs& a, const IterCarry& c, int slot) {
  DevState* h = a.snap_st + slot;
  h->f = c.f0;
  h->k = c.k;
  h->s_norm = c.s_norm;
  h->y_norm = c.y_norm;
  h->has_s = c.has_s;
  h->has_y = c.has_y;
  h->t_last = c.t_last;
  h->ls_evals = c.ls_evals;
  h->done = c.done;
  h->status = c.status;
  h->reason = c.reason;
  st_release_sys(a.snap_flag + slot, a.snap_seq0 + (unsigned long long)slot + 1ULL);
}

// everything of one outer iteration that comes before the H pass.  Returns 0: go on with the pass, 1: leave the loop.
template <class Fn, bool BOUNDED, bool BT, int SH, int KIND>
__device__ __noinline__ int iter_head(const QNIterArgs& a, const Fn& fn, IterSmem& sm, const bool epi_only) {
  constexpr bool SHARDED = SH != 0;
  IterCarry& c = sm.c;
  constexpr int BS = Fn::BS;
  const int tid = threadIdx.x, G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int64_t n = a.n;
  // this CTA's chunk of every O(n) vector: cw elements (even, a multiple of the functor block), one block per thread
  const int64_t nb = (n + BS - 1) / BS;
  int64_t bpc = (nb + G - 1) / G;
  if ((bpc * BS) & 1) bpc += 1;
  const int cw = (int)(bpc * BS);
  const int64_t i0 = (int64_t)cta * cw + (int64_t)tid * BS;  // first element of this thread's block
  const bool own = tid < bpc && i0 + BS <= n;                // (n is a multiple of BS for every functor)
  const bool wact = (int64_t)(tid & ~31) < bpc;              // this warp holds at least one block
  // BackTracking only reads c1 / beta: straight from shared memory.  The other searches mutate their parameters
  // (GLL window, MoreThuenteB t_max): they work on a per-thread copy, written back by one thread below.
  LSParams p_local;
  if (!BT) p_local = c.p;
  LSParams& p = BT ? c.p : p_local;
  // Every thread reads what it needs of the carried state into registers FIRST; the writes below (every thread the same
  // values) come after a barrier, so no thread can see a field another thread has already advanced.
  int gbuf = c.gbuf;
  unsigned int llseq = c.llseq;
  double f0 = c.f0, ys_prev = c.ys_prev, s_norm_in = c.s_norm, y_norm_in = c.y_norm;
  int has_s_in = c.has_s, has_y_in = c.has_y, skip_in = c.skip_prev;
  const int epi_in = c.epi_owed, ls_evals_in = c.ls_evals;
  const long long k_in = c.k;
  const int snap_in = c.snap_it;
  // The step sums of the previous iteration (s.s, y.y, y.s, f+) may still be per-CTA partials: nothing between the end
  // of that head and this point needs them, so they ride along with the epilogue's sums in ONE grid reduction instead
  // of costing one of their own (a pass always leaves an epilogue owed, so ds_in implies epi_in).
  const bool ds_in = c.ds_owed != 0;
  __syncthreads();
  iter_submark(a, sm, 14);  // (resets the sub-phase clock)
  if (!ds_in && !epi_only && is_bad(f0)) {  // ls_solver.rs:37-40
    c.done = 1;
    c.status = OSB_OUT_OF_DOMAIN;
    __syncthreads();
    return 1;
  }
  // ---- the epilogue the previous pass left owed: coefficients of the update that becomes pending, u = H+ g
  double ub_[BS];
  if (epi_in) {
    const int par = epi_in - 1;  // sharded: parity of the exchange buffers holding h, w
    bool skip_prev = skip_in != 0;  // (not known yet while ds_in: it follows from the deferred sums below)
    double hb[BS], wb[BS], sb[BS];
    double e3[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int jq = 0; jq < BS; ++jq) {
      hb[jq] = wb[jq] = sb[jq] = 0.0;
      if (own) {
        const int64_t i = i0 + jq;
        if (SHARDED) {  // rank-ordered sum of the per-rank slots: the same numbers in the same order on every rank
          const double* bh = a.peers[a.rank] + XSLOT_OFF + ((int64_t)(par * a.world) * 2 + 0) * XSLOT_LD + i;
          double vh = 0.0, vw = 0.0;
#pragma unroll 1
          for (int r0 = 0; r0 < a.world; r0 += 8) {  // eight slots' loads in flight, then the adds in rank order
            double th[8], tw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              th[e] = tw[e] = 0.0;
              if (r0 + e < a.world) {
                th[e] = __ldcg(bh + (int64_t)(r0 + e) * 2 * XSLOT_LD);
                tw[e] = __ldcg(bh + (int64_t)(r0 + e) * 2 * XSLOT_LD + XSLOT_LD);
              }
            }
            if (r0 == 0) {
              vh = th[0];
              vw = tw[0];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if ((r0 > 0 || e > 0) && r0 + e < a.world) {
                vh = vh + th[e];
                vw = vw + tw[e];
              }
            }
          }
          hb[jq] = vh;
          wb[jq] = vw;
        } else {
          hb[jq] = __ldcg(a.h + i);
          wb[jq] = __ldcg(a.w + i);
        }
        sb[jq] = __ldcg(a.s + i);
        if (ds_in || !skip_prev) {
          const double yi = __ldcg(a.y + i), gi = __ldcg(a.g + i);
          e3[0] = fma(yi, hb[jq], e3[0]);     // y.h
          e3[1] = fma(sb[jq], gi, e3[1]);     // s.g
          e3[2] = fma(hb[jq], gi, e3[2]);     // h.g
        }
      }
    }
    double ca = 0.0, cb = 0.0;
    double yh = 0.0, sg = 0.0, hg = 0.0;
    if (ds_in || !skip_prev) {
      warp_put(sm, 0, e3[0], wact);
      warp_put(sm, 1, e3[1], wact);
      warp_put(sm, 2, e3[2], wact);
      if (ds_in && tid < 4) ll_put(ll_row(a, (int)((llseq + 1u) & 1u), cta), 3 + tid, sm.ds_part[tid], llseq + 1u);
      iter_submark(a, sm, 4);
      grid_reduce<true>(a, sm, ds_in ? 7 : 3, 3, &gbuf, &llseq, false);
      iter_submark(a, sm, 5);
      yh = sm.res[0];
      sg = sm.res[1];
      hg = sm.res[2];
      if (ds_in) {  // the deferred step sums: what the end of the previous head would have set
        const double ss = sm.res[3], yy = sm.res[4], ys = sm.res[5], fn_ = sm.res[6];
        const double sn = sqrt(ss), yn = sqrt(yy);
        skip_prev = sn < a.tol || yn < a.tol;  // bfgs.rs:106-112
        f0 = fn_;
        ys_prev = ys;
        s_norm_in = sn;
        y_norm_in = yn;
        has_s_in = has_y_in = 1;
        skip_in = skip_prev ? 1 : 0;
        __syncthreads();  // every thread has read sm.res
        c.ss = ss;
        c.yy = yy;
        c.ys_prev = ys;
        c.f0 = fn_;
        c.s_norm = sn;
        c.y_norm = yn;
        c.has_s = c.has_y = 1;
        c.skip_prev = skip_in;
        c.ds_owed = 0;
        c.llseq = llseq;
        c.gbuf = gbuf;
        __syncthreads();
        if (a.snap_st != nullptr) {  // the previous iteration's snapshot (the barrier keeps the other threads' later
          if (cta == 0 && tid == 0) iter_publish(a, c, snap_in - 1);  // writes of c.done away from the publisher's reads)
          __syncthreads();
        }
        if (!epi_only && is_bad(f0)) {  // ls_solver.rs:37-40 (the epilogue stays owed, as on the undeferred path)
          c.done = 1;
          c.status = OSB_OUT_OF_DOMAIN;
          __syncthreads();
          return 1;
        }
      }
    }
    if (!skip_prev) {
      double c0, c1, c2;
      if (KIND == QN_BFGS) {  // bfgs.rs:115-124, no curvature safeguard
        const double rho = 1.0 / ys_prev;
        c0 = rho * rho * yh + rho;
        c1 = -rho;
        c2 = 0.0;
      } else {  // dfp.rs:115-120
        c0 = 1.0 / ys_prev;
        c1 = 0.0;
        c2 = -1.0 / yh;
      }
      ca = c0 * sg + c1 * hg;
      cb = c1 * sg + c2 * hg;
      // (all reads of the carried state happened before the barriers of grid_reduce)
      c.yh = yh;
      c.cc0 = c.pc0 = c0;
      c.cc1 = c.pc1 = c1;
      c.cc2 = c.pc2 = c2;
      c.pending = 1;
    } else {  // bfgs.rs:106-112: no new update; the stored matrix is exact once the pending one is applied
      c.pending = 0;
      c.pc0 = c.pc1 = c.pc2 = 0.0;
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
    c.epi_owed = 0;
  } else {
#pragma unroll
    for (int jq = 0; jq < BS; ++jq) ub_[jq] = own ? __ldcg(a.u + i0 + jq) : 0.0;
  }
  c.gbuf = gbuf;
    c.llseq = llseq;
  if (epi_only) return 1;
  // ---- has_converged (bfgs.rs:64-76), first two tests
  if (has_s_in && s_norm_in < a.tol) {
    c.done = 1;
    c.status = OSB_OK;
    c.reason = OSB_REASON_S_NORM;
    __syncthreads();
    return 1;
  }
  if (has_y_in && y_norm_in < a.tol) {
    c.done = 1;
    c.status = OSB_OK;
    c.reason = OSB_REASON_Y_NORM;
    __syncthreads();
    return 1;
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
  const int64_t j0 = (int64_t)cta * cw;
  int nblk = 0;
  if (BT) {
    const int64_t left = (n - j0) / BS;
    nblk = (int)(left < 0 ? 0 : (left < bpc ? left : bpc));
    if (own) {
#pragma unroll
      for (int jq = 0; jq < BS; ++jq) {
        sm.xs[tid * BS + jq] = xb[jq];
        sm.ds[tid * BS + jq] = db[jq];
      }
    }
    __syncthreads();
  }
  iter_submark(a, sm, 6);
  LSMachine m;
  int evals = 0;
  double gd0 = 0.0;
  if (BT) {
    // backtracking visits 1, beta, beta^2, ... whatever the outcome of a trial (backtracking.rs:37-55): the first
    // IT_NSPEC steps are evaluated before g.d is even known, and fed to the automaton in order afterwards
    double t0 = 1.0;
    bool first_round = true;
    for (;;) {
      warp_put(sm, 0, first_round ? gg : 0.0, wact);
      warp_put(sm, 1, first_round ? gd : 0.0, wact);
      // One warp per trial step: warp q evaluates t0 beta^q over the whole chunk (x and d staged in shared memory), its
      // lanes walking the blocks l, l + 32, ...; one interleaved shuffle tree for the three sums, and lane 0 writes the
      // CTA's partial directly.  (16 steps x 112 elements on 2 warps, with 48 serial shuffle trees, took 8 us.)
      {
        const int q = tid >> 5, lane = tid & 31;
        double tq = t0;
        for (int e = 0; e < q; ++e) tq = tq * p.beta;  // the same products, in the same order, as the automaton's t *= beta
        double f3[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
        for (int b = lane; b < nblk; b += 32) {
          double xt[BS], gt[BS], dd[BS];
#pragma unroll
          for (int jq = 0; jq < BS; ++jq) {
            const double xq = sm.xs[b * BS + jq];
            dd[jq] = sm.ds[b * BS + jq];
            const double td = tq * dd[jq];
            xt[jq] = xq + td;
            const double df = xt[jq] - xq;
            f3[2] = f3[2] + df * df;
          }
          f3[0] = f3[0] + fn.block(j0 + (int64_t)b * BS, xt, gt);
#pragma unroll
          for (int jq = 0; jq < BS; ++jq) f3[1] = f3[1] + gt[jq] * dd[jq];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          f3[0] = f3[0] + __shfl_xor_sync(0xffffffffu, f3[0], o);
          f3[1] = f3[1] + __shfl_xor_sync(0xffffffffu, f3[1], o);
          f3[2] = f3[2] + __shfl_xor_sync(0xffffffffu, f3[2], o);
        }
        if (lane < 3) ll_put(ll_row(a, (int)((llseq + 1u) & 1u), cta), 2 + 3 * q + lane, lane == 0 ? f3[0] : lane == 1 ? f3[1] : f3[2], llseq + 1u);
      }
      iter_submark(a, sm, 7);
      grid_reduce<true>(a, sm, 2 + 3 * IT_NSPEC, 2, &gbuf, &llseq, false);
      iter_submark(a, sm, 8);
      if (first_round) {
        if (sqrt(sm.res[0]) < a.tol) {  // bfgs.rs:74
          c.done = 1;
          c.status = OSB_OK;
          c.reason = OSB_REASON_GRAD_TOL;
          c.gbuf = gbuf;
    c.llseq = llseq;
          __syncthreads();
          return 1;
        }
        gd0 = sm.res[1];
        m.template begin<LS_BACKTRACKING>(p, f0, gd0, a.max_ls, INFINITY);
        first_round = false;
      }
      double tq = t0;
#pragma unroll 1
      for (int q = 0; q < IT_NSPEC; ++q) {
        if (m.done) break;
        if (m.request(p) == tq) {
          m.template feed<LS_BACKTRACKING>(p, sm.res[2 + 3 * q], sm.res[3 + 3 * q], sm.res[4 + 3 * q]);
          ++evals;
        }
        tq = tq * p.beta;
      }
      if (m.done) break;
      t0 = m.request(p);
      __syncthreads();  // sm.res is rewritten by the next round
    }
  } else {
    warp_put(sm, 0, gg, wact);
    warp_put(sm, 1, gd, wact);
    grid_reduce<true>(a, sm, 2, 2, &gbuf, &llseq, false);
    if (sqrt(sm.res[0]) < a.tol) {  // bfgs.rs:74
      c.done = 1;
      c.status = OSB_OK;
      c.reason = OSB_REASON_GRAD_TOL;
      c.gbuf = gbuf;
    c.llseq = llseq;
      __syncthreads();
      return 1;
    }
    gd0 = sm.res[1];
    double tmaxc = INFINITY;
    if (need_tmax) {
      __syncthreads();
      const double wm = warp_min(tm);
      if ((tid & 31) == 0) sm.w[tid >> 5] = wm;
      grid_reduce<true>(a, sm, 1, 1, &gbuf, &llseq, true);
      tmaxc = sm.res[0];
    }
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
      __syncthreads();  // sm.res of the previous round has been consumed by every thread
      warp_put(sm, 0, a3[0], wact);
      warp_put(sm, 1, a3[1], wact);
      warp_put(sm, 2, a3[2], wact);
      grid_reduce<true>(a, sm, 3, 3, &gbuf, &llseq, false);
      m.feed(p, sm.res[0], sm.res[1], sm.res[2]);
      ++evals;
    }
  }
  iter_submark(a, sm, 9);
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
      if (a.snap_x != nullptr) {  // this iteration's snapshot for the host callback (device ring, copied out by a DMA)
        double* sx = a.snap_x + (int64_t)snap_in * 2 * a.ld;
        sx[i] = xn[jq];
        sx[a.ld + i] = gn[jq];
      }
      a4[0] = a4[0] + si * si;
      a4[1] = a4[1] + yi * yi;
      a4[2] = a4[2] + yi * si;
    }
  }
  iter_submark(a, sm, 10);
  __syncthreads();  // sm.res of the line search has been consumed by every thread
  warp_put(sm, 0, a4[0], wact);
  warp_put(sm, 1, a4[1], wact);
  warp_put(sm, 2, a4[2], wact);
  warp_put(sm, 3, a4[3], wact);
  // The four sums stay per-CTA partials (the next head, or the end of the launch, reduces them over the grid); what the
  // pass needs now is only a grid barrier, which publishes s, y, x, g, ps, ph and the snapshot slot.
  __syncthreads();
  if (tid < 4) {
    double v = 0.0;
#pragma unroll 1
    for (int wq = 0; wq < IT_NW; ++wq) v = v + sm.w[tid * IT_NW + wq];
    sm.ds_part[tid] = v;
  }
  cg::this_grid().sync();
  iter_submark(a, sm, 11);
  {
    c.ds_owed = 1;
    c.t_last = t;
    c.gd0_last = gd0;
    c.k = k_in + 1;
    c.snap_it = snap_in + 1;
    c.ls_evals = ls_evals_in + evals + 1;
    c.gbuf = gbuf;
    c.llseq = llseq;
  }
  if (!BT && (p.kind == LS_GLL || p.kind == LS_MORETHUENTE_B)) {  // only f_previous / t_max persist across iterations
    __syncthreads();
    if (tid == 0) c.p = p_local;
  }
  __syncthreads();
  return 0;
}

// the deferred step sums at the end of a launch (no next head to carry them): one grid reduction of their own
__device__ __noinline__ void iter_finalize_ds(const QNIterArgs& a, IterSmem& sm) {
  IterCarry& c = sm.c;
  const int tid = threadIdx.x, cta = (int)blockIdx.x;
  int gbuf = c.gbuf;
  unsigned int llseq = c.llseq;
  __syncthreads();
  if (tid < 4) ll_put(ll_row(a, (int)((llseq + 1u) & 1u), cta), tid, sm.ds_part[tid], llseq + 1u);
  grid_reduce<true>(a, sm, 4, 0, &gbuf, &llseq, false);
  const double ss = sm.res[0], yy = sm.res[1], ys = sm.res[2], fn_ = sm.res[3];
  const double sn = sqrt(ss), yn = sqrt(yy);
  __syncthreads();
  c.ss = ss;
  c.yy = yy;
  c.ys_prev = ys;
  c.f0 = fn_;
  c.s_norm = sn;
  c.y_norm = yn;
  c.has_s = c.has_y = 1;
  c.skip_prev = (sn < a.tol || yn < a.tol) ? 1 : 0;  // bfgs.rs:106-112
  c.ds_owed = 0;
  c.llseq = llseq;
  c.gbuf = gbuf;
  __syncthreads();
  if (a.snap_st != nullptr && cta == 0 && tid == 0) iter_publish(a, c, c.snap_it - 1);
}

// fold of the column partials (+ exchange) for this CTA's chunk, out of line like the head
template <class Fn, int SH>
__device__ __noinline__ void iter_tail(const QNIterArgs& a, IterSmem& sm) {
  constexpr bool SHARDED = SH != 0;
  IterCarry& c = sm.c;
  constexpr int BS = Fn::BS;
  const int G = (int)gridDim.x;
  const int64_t nb = (a.n + BS - 1) / BS;
  int64_t bpc = (nb + G - 1) / G;
  if ((bpc * BS) & 1) bpc += 1;
  const int cw = (int)(bpc * BS);
  const unsigned long long seq = c.seq + 1ULL;
  iter_fold<SH>(a, sm, (int64_t)blockIdx.x * cw, cw, seq, a.st);
  __syncthreads();
  c.seq = seq;
  c.epi_owed = SHARDED ? 1 + (int)(seq & 1ULL) : 1;
  __syncthreads();
}

__device__ __noinline__ void iter_load_state(const QNIterArgs& a, IterCarry& c, bool sharded) {
  const DevState* st = a.st;
  c.f0 = st->f;
  c.has_s = st->has_s;
  c.has_y = st->has_y;
  c.s_norm = st->s_norm;
  c.y_norm = st->y_norm;
  c.ys_prev = st->ys;
  c.yh = st->yh;
  c.skip_prev = st->skip;
  c.pending = st->pending;
  c.epi_owed = st->epi;
  c.pc0 = st->pc0;
  c.pc1 = st->pc1;
  c.pc2 = st->pc2;
  c.cc0 = st->c0;
  c.cc1 = st->c1;
  c.cc2 = st->c2;
  c.k = st->k;
  c.ls_evals = st->ls_evals;
  c.t_last = st->t_last;
  c.gd0_last = st->gd0;
  c.ss = st->ss;
  c.yy = st->yy;
  c.seq = sharded ? *a.seq : 0ULL;
  c.status = st->status;
  c.reason = st->reason;
  c.done = 0;
  c.gbuf = 0;
  c.llseq = (unsigned int)st->ll_seq;
  c.snap_it = 0;
  c.ds_owed = 0;
  c.p = *a.lsp;
}

// persist the carried state (one thread; every CTA holds the same values)
__device__ __noinline__ void iter_store_state(const QNIterArgs& a, const IterCarry& c, bool sharded) {
  DevState* st = a.st;
  st->f = c.f0;
  st->ft = c.f0;
  st->gd0 = c.gd0_last;
  st->ss = c.ss;
  st->yy = c.yy;
  st->ys = c.ys_prev;
  st->yh = c.yh;
  st->s_norm = c.s_norm;
  st->y_norm = c.y_norm;
  st->has_s = c.has_s;
  st->has_y = c.has_y;
  st->skip = c.skip_prev;
  st->c0 = c.cc0;
  st->c1 = c.cc1;
  st->c2 = c.cc2;
  st->pc0 = c.pc0;
  st->pc1 = c.pc1;
  st->pc2 = c.pc2;
  st->pending = c.pending;
  st->epi = c.epi_owed;
  st->t_last = c.t_last;
  st->k = c.k;
  st->ls_evals = c.ls_evals;
  st->ll_seq = (int)c.llseq;
  if (c.done) {
    st->done = 1;
    st->status = c.status;
    st->reason = c.reason;
  }
  if (sharded) *a.seq = c.seq;
  if (c.p.kind == LS_GLL || c.p.kind == LS_MORETHUENTE_B) *a.lsp = c.p;  // only f_previous / t_max persist across iterations
}

__device__ __forceinline__ long long iter_stamp() {
  long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

__device__ __noinline__ void iter_mark(const QNIterArgs& a, IterSmem& sm, int slot) {
  if (a.prof == nullptr || blockIdx.x != 0 || threadIdx.x != 0) return;
  const long long now = iter_stamp();
  sm.prof[slot] += now - sm.prof[3];
  sm.prof[3] = now;
}

template <class Fn, bool BOUNDED, bool BT, int SH, int KIND>
__global__ void __launch_bounds__(IT_NT, 1) qn_iter_kernel(const __grid_constant__ QNIterArgs a, const __grid_constant__ Fn fn) {
  constexpr bool SHARDED = SH != 0;
  cg::grid_group grid = cg::this_grid();
  __shared__ IterSmem sm;
  if (a.st->done) return;
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  iter_load_state(a, sm.c, SHARDED);
  __syncthreads();
  {
    const int64_t T = (a.n + QN_R - 1) / QN_R;
    // partial vector q is valid on the columns below the first row of CTA q's first tile
    for (int q = threadIdx.x; q < (int)gridDim.x; q += IT_NT)
      sm.ext[q] = (int)sym_first_row<SHARDED>(T, a.world, a.rank, (int)gridDim.x, q);
  }
  grid.sync();  // every CTA has read the entry state before anybody can write it
  // (time stamps live in shared memory, not in registers: nothing but &c and the loop counter is live across the pass)
  if (a.prof != nullptr && leader) {
    for (int q = 0; q < 16; ++q) sm.prof[q] = 0;
    sm.prof_skip = 0;
    sm.prof[3] = iter_stamp();
    sm.prof[15] = sm.prof[3];
  }
  int it = 0;
  if (a.epi_only) {
    iter_head<Fn, BOUNDED, BT, SH, KIND>(a, fn, sm, true);
  } else {
    for (; it < a.iters; ++it) {
      if (iter_head<Fn, BOUNDED, BT, SH, KIND>(a, fn, sm, false)) {
        if (a.snap_st != nullptr && leader) iter_publish(a, sm.c, sm.c.snap_it);  // done: tells the host to stop waiting
        break;
      }
      iter_mark(a, sm, 0);
      // ---- the H pass: pending update + h = H y + w = H g over (this rank's share of) the packed triangle.  Inlined,
      // every pointer straight from the kernel's constant bank; only &c is live across it.
      {
        QNLazyArgs la{};
        la.ps = a.ps;
        la.ph = a.ph;
        la.y = a.y;
        la.g = a.g;
        la.h = a.h;
        la.w = a.w;
        const QNSymArgs sa{a.P, a.P, a.colpart, a.n, a.ld, SHARDED ? a.world : 1, SHARDED ? a.rank : 0, a.peers, a.seq, (int)gridDim.x, 0};
        sym_pass_body<KIND, SHARDED, IT_NT, false, false>(la, sa, sm.c.pc0, sm.c.pc1, sm.c.pc2, 0, (int)gridDim.x, (int)blockIdx.x);
      }
      grid.sync();
      iter_mark(a, sm, 1);
      iter_tail<Fn, SH>(a, sm);
      iter_mark(a, sm, 2);
      if (a.prof != nullptr && leader && it == 0 && a.iters > 1) {
        // the first iteration of a launch waits for the slowest rank's kernel to START: not part of the steady state
        for (int q = 0; q < 15; ++q)
          if (q != 3) sm.prof[q] = 0;
        sm.prof_skip = 1;
      }
    }
  }
  if (sm.c.ds_owed) iter_finalize_ds(a, sm);  // (the same value in every thread of every CTA)
  if (leader) {
    iter_store_state(a, sm.c, SHARDED);
    if (a.prof != nullptr) {
      a.prof[0] += sm.prof[0];
      a.prof[1] += sm.prof[1];
      a.prof[2] += sm.prof[2];
      a.prof[3] += it - sm.prof_skip;
      for (int q = 4; q < 15; ++q) a.prof[q] += sm.prof[q];
    }
  }
}

// calibration: the fixed cost of one grid barrier of this kernel's shape (148 CTAs x 512 threads, cooperative launch)
__global__ void __launch_bounds__(IT_NT, 1) grid_sync_bench_kernel(int reps, long long* out) {
  cg::grid_group grid = cg::this_grid();
  grid.sync();
  const long long t0 = iter_stamp();
  for (int r = 0; r < reps; ++r) grid.sync();
  const long long t1 = iter_stamp();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}
double bench_grid_sync(Ctx* ctx, int reps) {
  long long* d_out = nullptr;
  OSB_CUDA(cudaMalloc(&d_out, sizeof(long long)));
  void* params[] = {(void*)&reps, (void*)&d_out};
  for (int rep = 0; rep < 2; ++rep)
    OSB_CUDA(cudaLaunchCooperativeKernel((const void*)grid_sync_bench_kernel, dim3(qn_iter_grid(ctx)), dim3(IT_NT), params, 0, ctx->stream));
  long long ns = 0;
  OSB_CUDA(cudaMemcpyAsync(&ns, d_out, sizeof(ns), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  cudaFree(d_out);
  return (double)ns * 1e-3 / reps;  // us per barrier
}

// ---- host side ------------------------------------------------------------------------------
int qn_iter_grid(Ctx* ctx) { return ctx->num_sms < IT_MAXG ? ctx->num_sms : IT_MAXG; }
int64_t qn_iter_gpart_doubles(Ctx* ctx) {  // 2 plain buffers | spare | 2 flagged buffers of 2 words per value
  return 2 * (int64_t)qn_iter_grid(ctx) * IT_GPK + IT_MAXG + 2 * (int64_t)qn_iter_grid(ctx) * 2 * IT_GPK;
}

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

template <class Fn, bool BOUNDED, bool BT, int SH, int KIND>
static void launch_iter_kk(Ctx* ctx, const QNIterArgs& a, const Fn& fn) {
  auto kern = qn_iter_kernel<Fn, BOUNDED, BT, SH, KIND>;
  QNIterArgs aa = a;
  Fn f = fn;
  void* params[] = {(void*)&aa, (void*)&f};
  OSB_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(qn_iter_grid(ctx)), dim3(IT_NT), params, 0, ctx->stream));
  ctx->counters[0]++;
}

template <class Fn, bool BOUNDED, bool BT, int SH>
static void launch_iter_k(Ctx* ctx, const QNIterArgs& a, const Fn& fn) {
  if (a.kind == QN_BFGS) launch_iter_kk<Fn, BOUNDED, BT, SH, QN_BFGS>(ctx, a, fn);
  else launch_iter_kk<Fn, BOUNDED, BT, SH, QN_DFP>(ctx, a, fn);
}

template <class Fn>
static void launch_iter_fn(Ctx* ctx, const QNIterArgs& a, const Fn& fn, bool bounded, bool bt, int sh) {
#define OSB_IT(B, T)                                   \
  do {                                                 \
    if (sh == 0) launch_iter_k<Fn, B, T, 0>(ctx, a, fn);      \
    else launch_iter_k<Fn, B, T, 1>(ctx, a, fn);              \
  } while (0)
  if (bounded) {
    if (bt) OSB_IT(true, true);
    else OSB_IT(true, false);
  } else {
    if (bt) OSB_IT(false, true);
    else OSB_IT(false, false);
  }
#undef OSB_IT
}

void qn_launch_iter(Ctx* ctx, int functor_kind, const double* fn_a, const double* fn_b, bool bounded, int ls_kind, const QNIterArgs& a) {
  const bool bt = ls_kind == LS_BACKTRACKING;
  const int sh = a.world > 1 ? 1 : 0;
  if (functor_kind == FN_ROSENBROCK) {
    launch_iter_fn(ctx, a, RosenbrockFn{}, bounded, bt, sh);
  } else if (functor_kind == FN_SEPQUAD) {
    SepQuadFn fn;
    fn.c = fn_a;
    fn.a = fn_b;
    fn.index0 = 0;
    launch_iter_fn(ctx, a, fn, bounded, bt, sh);
  } else {
    throw Error(OSB_ERR_UNSUPPORTED, "objective has no block functor for the device-resident engine");
  }
}

}  // namespace osb
