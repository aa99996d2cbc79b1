// PSO / competitive PSO: one fused kernel per synchronous generation.
// Reference: stochopy/optimize/cpso/_cpso.py:324-361 (mutation, pso_sync),
// cpso/_constraints.py (NoConstraint, Shrink), :405-426 (restart) and
// selection_sync (_common.py:123-160) with cand = X, x = pbest.
//
// Per particle (row-local, in place): read X, V, pbest; write X, V and, on
// improvement only, pbest.  Algorithmic HBM bytes per particle: 5 * N * s + 3 * s.
#include <cstdlib>
#include <type_traits>

#include "peer.cuh"
#include "rows.cuh"

namespace sp {

template <typename T>
struct PsoArgs {
  int objective, constraint, it, maxiter, N, propose_only;
  int64_t P, ld;
  T w, c1, c2;
  double xtol, ftol;
  uint64_t seed;
  T* X;
  T* V;
  T* pbest;
  T* pbestfit;
  T* pfit;
  T* gbest;
  const T* lower;
  const T* upper;
  sp_ctrl* ctrl;
  Best* scratch;
  const T* r1;
  const T* r2;
  int64_t row0, P_total;  // sharded swarm: global index of local row 0, whole swarm size
  T* xch;                  // shard mode: [fit, x_0..x_{N-1}] of the local best for the exchange
  int shard;
  PeerArgs peer;           // shard == 2: exchange inside the kernel over the peer mailboxes
  int chain;               // SP_CHAIN_IN / SP_CHAIN_OUT (plain PSO, whole swarm)
  T* chain_rows;           // [2 parities][kChainRegion][ld]: the best row of every CTA
  int rbound;              // CPSO, whole swarm, lazy run: the restart decision is taken in this kernel's epilogue (below)
  double gamma, delta;
};

// CPSO restart decision without a second pass over the swarm (rbound).  The reference tests
//   radius(g) = max_i |X_i - gbest(g)| / sqrt(4 N) < delta            (cpso/_cpso.py:405-412)
// with the gbest of THIS generation, which only exists once every CTA has finished.  But every CTA knows
// gbest(g-1) while it moves its rows, and for every i  | |X_i - gbest(g)| - |X_i - gbest(g-1)| | <= dist with
// dist = |gbest(g) - gbest(g-1)| -- which the epilogue computes anyway (it is selection_sync's x-tolerance test).
// So with M = max_i |X_i - gbest(g-1)| (accumulated in the row loop, 4 subtractions + 4 FMAs + a group reduce per
// row) the radius lies in [M - dist, M + dist] / sqrt(4 N): if the whole interval is on one side of delta the
// decision is the reference's exactly; dist = 0 (gbest did not move) makes it exact always.  Otherwise the run parks
// with ctrl->flag = -1 and sp_cpso_restart_resume takes the exact decision with the radius kernel.  eps covers the
// rounding of the working-precision sums.
template <typename T>
__device__ __forceinline__ void restart_decide_bound(sp_ctrl* ctrl, double m2, int N, int64_t Ptot, int it, int maxiter,
                                                     double gamma, double delta, bool force_undecided) {
  if (ctrl->status != SP_RUNNING) {  // the generation terminated the run: no restart (_cpso.py:304)
    ctrl->flag = 0;
    return;
  }
  const double eps = sizeof(T) == 4 ? 1.0e-4 : 1.0e-10;
  const double dist = ctrl->dist, M = sqrt(m2), s4n = sqrt(4.0 * (double)N);
  const double hi = (M * (1.0 + eps) + dist) / s4n, lo = (M * (1.0 - eps) - dist) / s4n;
  ctrl->aux[1] = M / s4n;
  if (force_undecided) {  // test hook (SP_CPSO_FORCE_AMBIGUOUS): every generation goes through the exact path
    ctrl->flag = -1;
    ctrl->status = SP_STATUS_RESTART_PENDING;
    return;
  }
  if (lo >= delta) {  // certainly no restart
    ctrl->flag = 0;
    return;
  }
  int nw = -1;        // -1: cannot tell from the bound
  if (hi < delta) {   // certainly a restart: nw of _cpso.py:415-416
    const double inorm = (double)it / (double)maxiter;
    nw = (int)(((double)Ptot - 1.0) / (1.0 + exp(1.0 / 0.09 * (inorm - gamma + 0.5))));
    if (nw <= 0) {
      ctrl->flag = 0;
      return;
    }
  }
  ctrl->flag = nw;
  ctrl->status = SP_STATUS_RESTART_PENDING;
}

// shard == 2, last CTA of the generation kernel: local best -> every peer's mailbox (NVLink
// stores), flags, wait, then selection_sync's reduction over the `world` records on every rank.
template <typename T>
__device__ __forceinline__ void peer_best_exchange(const PsoArgs<T>& a, Best top) {
  const PeerArgs& p = a.peer;
  const int par = a.it & 1;
  const int off = 16 / (int)sizeof(T);  // x starts 16-byte aligned behind the fitness
  const T* src = a.pbest + top.row * a.ld;
  for (int r = 0; r < p.world; ++r) {
    T* dst = peer_rec<T>(p, r, par, p.rank);
    for (int j = threadIdx.x; j < a.N; j += blockDim.x) dst[off + j] = src[j];
    if (threadIdx.x == 0) dst[0] = (T)top.f;
  }
  if (!peer_exchange_flags(p, kPeerBest, par, (uint32_t)a.it)) {
    if (threadIdx.x == 0) a.ctrl->status = SP_STATUS_PEER_TIMEOUT;
    return;
  }
  int best = 0;
  T bf = ld_volatile(peer_rec<T>(p, p.rank, par, 0));
  for (int r = 1; r < p.world; ++r) {  // first minimum: rank order is row order (np.argmin's tie rule)
    const T f = ld_volatile(peer_rec<T>(p, p.rank, par, r));
    if (f < bf) {
      bf = f;
      best = r;
    }
  }
  finalize_generation<T, true>(Best{(double)bf, (long long)best}, peer_rec<T>(p, p.rank, par, 0) + off, p.L.rec_ld, a.N,
                               a.gbest, a.ctrl, a.it, a.maxiter, a.xtol, a.ftol);
}

// CHAIN: compiled with the chained-generation prologue / epilogue (fp32 whole-swarm PSO); the
// plain instantiation keeps the register footprint of the unchained kernel.
// PLAIN (host-checked): no constraint, not propose-only, ndim == CH * LPR * VEC == ld and popsize a multiple
// of the rows per warp -- no bounds predicates, no dead-row tests, compile-time row stride.
// OBJ >= 0: the objective is a compile-time constant (no jump table in the row loop); -1: a.objective.
template <typename T, int CH, int LPR, bool PHILOX, bool CHAIN, int OBJ = -1, bool PLAIN = false>
__global__ void __launch_bounds__(kThreads)
pso_generation_kernel(const PsoArgs<T> a, const PhiloxKeys keys) {
  using TL = Tile<T, CH, LPR>;
  constexpr int VEC = Num<T>::VEC;
  pdl_wait();  // chained launches overlap their launch with the previous generation's drain
  if (!running(a.ctrl)) return;
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, l = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t groups = (a.P + TL::RPW - 1) / TL::RPW;
  const int ld = PLAIN ? TL::COLS : (int)a.ld;
  const int64_t ldr = PLAIN ? (int64_t)TL::COLS : a.ld;  // row stride
  const int N = PLAIN ? TL::COLS : a.N;

  // software pipeline: the next row group's X, V, pbest are in flight while this one computes
  TL nx, nv, npb;
  auto fetch = [&](int64_t g) {
    int64_t r = g * TL::RPW + sub;
    if (!PLAIN && r >= a.P) r = a.P - 1;
    nx.load(a.X + r * ldr, l, ld);
    nv.load(a.V + r * ldr, l, ld);
    npb.load(a.pbest + r * ldr, l, ld);
  };
  bool have_first = false;
  TL gb;
  if (CHAIN && (a.chain & SP_CHAIN_IN)) {
    if (warp < groups) {  // the first rows do not depend on gbest: their loads overlap the reduction below
      fetch(warp);
      have_first = true;
    }
    // chained generations (see common.cuh): reduce the previous launch's per-CTA minima here; the
    // winning row was left by its CTA in chain_rows (pbest itself is updated in place below)
    int idx;
    const Best top = chain_best_block(chain_region(a.scratch, a.it - 1), (int)gridDim.x, &idx);
    const T* win = a.chain_rows + ((size_t)((a.it - 1) & 1) * kChainRegion + idx) * a.ld;
    if (blockIdx.x == 0 && (threadIdx.x >> 5) == (kThreads / 32) - 1) {  // one warp: gbest / dist / nit / status
      Best rec{top.f, 0};
      finalize_generation_warp<T>(rec, win, a.ld, a.N, a.gbest, a.ctrl, a.it - 1, a.maxiter, a.xtol, a.ftol);
      if (lane == 0) a.ctrl->gbest_row = top.row;
    }
    if (chain_stops(top.f, a.it - 1, a.maxiter, a.ftol)) return;
    gb.load(win, l, ld);
  } else {
    gb.load(a.gbest, l, ld);
  }

  Best mine{1.0 / 0.0, 0x7fffffffffffffffLL};
  T mine_f = Num<T>::inf();  // PLAIN: the lane's minimum in T (rows ascend, strict < keeps the first = np.argmin's tie rule)
  int64_t mine_row = 0x7fffffffffffffffLL;
  T dmax = 0;  // rbound: max over this thread's rows of |X_i - gbest(g-1)|^2
  // measured on B200 (C3, fp32 N=64): prefetching one group ahead costs 26 registers and a resident
  // CTA per SM and is slower (21.3 vs 17.0 us per generation) -- the state is L2 resident; keep it off
  // (round 2, PLAIN variant, 48 / 63 registers with the prefetch: 11.2 vs 11.1 us per generation, no gain)
  constexpr bool kPrefetch = false;
  if (kPrefetch && warp < groups && !(CHAIN && have_first)) fetch(warp);
  for (int64_t g = warp; g < groups; g += nwarps) {
    int64_t row = g * TL::RPW + sub;
    const bool live = PLAIN || row < a.P;
    if (!live) row = a.P - 1;

    if (!kPrefetch && !(CHAIN && have_first)) fetch(g);
    have_first = false;
    TL x = nx, v = nv;
    {
      TL pb = npb;
      if (kPrefetch && g + nwarps < groups) fetch(g + nwarps);
      // V = w V + c1 r1 (pbest - X) + c2 r2 (gbest - X), _cpso.py:326 (numpy's order)
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int j0 = TL::col(c, l, 0);
        T r1[VEC], r2[VEC];
        if (PHILOX) {  // r1 and r2 of four columns from one call (philox.cuh pso_r12)
          pso_r12(philox4x32_keyed((uint32_t)(j0 >> 2), (uint32_t)(a.row0 + row), (uint32_t)a.it, kPsoR1, keys), j0, r1, r2);
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const bool in = j0 + e < a.N;
            r1[e] = in ? a.r1[row * a.ld + j0 + e] : T(0);
            r2[e] = in ? a.r2[row * a.ld + j0 + e] : T(0);
          }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const T xe = x.v[c][e];
          T t = mul_rn(a.w, v.v[c][e]);
          t = add_rn(t, mul_rn(mul_rn(a.c1, r1[e]), sub_rn(pb.v[c][e], xe)));
          t = add_rn(t, mul_rn(mul_rn(a.c2, r2[e]), sub_rn(gb.v[c][e], xe)));
          v.v[c][e] = (PLAIN || j0 + e < a.N) ? t : T(0);
        }
      }
    }

    if (!PLAIN && a.constraint == SP_CONS_SHRINK) {  // cpso/_constraints.py:22-55
      T beta = Num<T>::inf();
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int j = TL::col(c, l, e);
          if (j < a.N) {
            const T trial = add_rn(x.v[c][e], v.v[c][e]);
            const T lo = a.lower[j], hi = a.upper[j];
            if (trial < lo) beta = fmin(beta, div_rn(sub_rn(lo, x.v[c][e]), v.v[c][e]));
            if (trial > hi) beta = fmin(beta, div_rn(sub_rn(hi, x.v[c][e]), v.v[c][e]));
          }
        }
      beta = group_min<LPR>(beta);
      if (beta != Num<T>::inf()) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int e = 0; e < VEC; ++e) v.v[c][e] = mul_rn(v.v[c][e], beta);
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < VEC; ++e) x.v[c][e] = add_rn(x.v[c][e], v.v[c][e]);

    if (live) {
      x.store(a.X + row * ldr, l, ld);
      v.store(a.V + row * ldr, l, ld);
    }
    if (!CHAIN && a.rbound) {  // (padding columns are zero in both; the chained kernels never take this decision)
      T d2 = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const T t = x.v[c][e] - gb.v[c][e];
          d2 += t * t;
        }
      d2 = group_sum<LPR>(d2);
      if (live && d2 > dmax) dmax = d2;
    }
    if (!PLAIN && a.propose_only) continue;  // SP_OBJ_HOST: caller evaluates X, then sp_select_sync(copy_when=1)

    const T f = evaluate_tile<T, CH, LPR>(OBJ >= 0 ? OBJ : a.objective, x, l, N);
    T best = a.pbestfit[row];
    const bool win = f < best;
    if (win) best = f;
    if (live) {
      if (win) x.store(a.pbest + row * ldr, l, ld);
      if (l == 0) {
        if (!PLAIN || win) a.pbestfit[row] = best;
        a.pfit[row] = f;
        if (PLAIN) {
          if (best < mine_f) {
            mine_f = best;
            mine_row = row;
          }
        } else if (better((double)best, row, mine.f, mine.row)) {
          mine = Best{(double)best, row};
        }
      }
    }
  }
  if (!PLAIN && a.propose_only) return;
  if (PLAIN) mine = Best{(double)mine_f, mine_row};
  if (CHAIN && (a.chain & SP_CHAIN_OUT)) {  // leave the CTA's minimum and its row for the next launch's prologue
    __shared__ Best s_red[32];
    __shared__ long long s_row;
    const Best b = block_best(mine, s_red);
    if (threadIdx.x == 0) {
      chain_region(a.scratch, a.it)[blockIdx.x] = b;
      s_row = b.row;
    }
    __syncthreads();
    if (s_row < a.P) {  // a CTA without rows keeps the +inf record
      const T* src = a.pbest + s_row * a.ld;
      T* dst = a.chain_rows + ((size_t)(a.it & 1) * kChainRegion + blockIdx.x) * a.ld;
      for (int j = threadIdx.x; j < ld; j += blockDim.x) dst[j] = src[j];
    }
    return;
  }
  __shared__ double s_dmax[kThreads / 32];
  if (!CHAIN && a.rbound) {  // the CTA's maximum goes to scratch region 1 (the chained kernels' region; this one is not chained)
    double d = (double)dmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) s_dmax[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kThreads / 32; ++w) d = fmax(d, s_dmax[w]);
      chain_region(a.scratch, 1)[blockIdx.x].f = d;  // published by grid_best's fence + arrival counter
    }
  }
  Best top;
  if (grid_best(mine, a.scratch, a.ctrl, &top)) {
    if (a.shard == 2) {
      peer_best_exchange<T>(a, top);
    } else if (a.shard) {  // local best -> exchange record; gbest/status come from sp_gbest_reduce on every rank
      const T* src = a.pbest + top.row * a.ld;
      for (int j = threadIdx.x; j < a.N; j += blockDim.x) a.xch[1 + j] = src[j];
      if (threadIdx.x == 0) a.xch[0] = (T)top.f;
    } else {
      finalize_generation<T>(top, a.pbest, a.ld, a.N, a.gbest, a.ctrl, a.it, a.maxiter, a.xtol, a.ftol);
      if (!CHAIN && a.rbound) {  // (thread 0 wrote ctrl->dist / status in finalize_generation; it also decides)
        double d = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) d = fmax(d, __ldcg(&chain_region(a.scratch, 1)[i].f));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_dmax[threadIdx.x >> 5] = d;
        __syncthreads();
        if (threadIdx.x == 0) {
          for (int w = 1; w < kThreads / 32; ++w) d = fmax(d, s_dmax[w]);
          restart_decide_bound<T>(a.ctrl, d, a.N, a.P, a.it, a.maxiter, a.gamma, a.delta, a.rbound == 2);
        }
      }
    }
  }
}

// ---- competitive restart, _cpso.py:405-426 ------------------------------------
// (1) max_i |X_i - gbest|^2 -> ctrl->aux[0] (bit pattern, atomicMax on non-negative doubles)
// parked != 0: runs only for a run parked by the bound decision with ctrl->flag == -1 (radius undecided)
template <typename T>
__global__ void __launch_bounds__(kThreads)
radius_kernel(const T* __restrict__ X, const T* __restrict__ gbest, int64_t P, int N, int64_t ld, sp_ctrl* ctrl,
              int parked = 0) {
  if (parked) {
    if (ctrl->status != SP_STATUS_RESTART_PENDING || ctrl->flag != -1) return;
  } else if (!running(ctrl)) {
    return;
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  double worst = 0.0;
  for (int64_t row = warp; row < P; row += nwarps) {
    double acc = 0.0;
    for (int j = lane; j < N; j += 32) {
      double d = (double)(T)(X[row * ld + j] - gbest[j]);
      acc += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    worst = fmax(worst, acc);
  }
  __shared__ double s_w[kThreads / 32];
  if (lane == 0) s_w[threadIdx.x >> 5] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) worst = fmax(worst, s_w[w]);
    atomicMax(reinterpret_cast<unsigned long long*>(&ctrl->aux[0]), (unsigned long long)__double_as_longlong(worst));
  }
}

// (1)+(2) fused for the whole-swarm fast path: 16-byte loads, a lane group per row, the last CTA
// to finish takes the decision (no second launch).  pause != 0: a restart that fires parks the
// optimiser (status = SP_STATUS_RESTART_PENDING, every later kernel of the chunk returns at
// once) so that the three gated restart kernels need not be enqueued after every generation;
// the host then runs sp_cpso_restart_resume and continues with the next generation.
template <typename T>
__global__ void __launch_bounds__(kThreads)
radius_plan_kernel(const T* __restrict__ X, const T* __restrict__ gbest, int64_t P, int N, int64_t ld, sp_ctrl* ctrl,
                   int it, int maxiter, double gamma, double delta, int lpr, int pause, int64_t Ptot,
                   const PeerArgs peer) {
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  pdl_wait();
  if (!running(ctrl)) {  // terminated: no restart; parked: ctrl->flag still holds the nw of the pending restart
    if (blockIdx.x == 0 && threadIdx.x == 0 && ctrl->status != SP_STATUS_RESTART_PENDING) ctrl->flag = 0;
    return;
  }
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, l = lane % lpr, sub = lane / lpr, rpw = 32 / lpr;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int nvec = (int)(ld / VEC);
  const int64_t groups = (P + rpw - 1) / rpw;
  double worst = 0.0;
  for (int64_t g = warp; g < groups; g += nwarps) {
    const int64_t row = g * rpw + sub;
    double acc = 0.0;
    if (row < P) {
      const V* xr = reinterpret_cast<const V*>(X + row * ld);
      for (int j = l; j < nvec; j += lpr) {
        const V a = xr[j], b = __ldg(reinterpret_cast<const V*>(gbest) + j);
        const T* pa = reinterpret_cast<const T*>(&a);
        const T* pb = reinterpret_cast<const T*>(&b);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const double d = (double)(T)(pa[e] - pb[e]);  // padding columns are zero in both
          acc += d * d;
        }
      }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    worst = fmax(worst, acc);
  }
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  __shared__ double s_w[kThreads / 32];
  __shared__ bool s_last;
  if (lane == 0) s_w[threadIdx.x >> 5] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) worst = fmax(worst, s_w[w]);
    atomicMax(reinterpret_cast<unsigned long long*>(&ctrl->aux[0]), (unsigned long long)__double_as_longlong(worst));
    __threadfence();
    s_last = atomicAdd(&ctrl->done_blocks, 1u) == gridDim.x - 1;
    if (s_last) __threadfence();
  }
  __syncthreads();
  if (!s_last) return;
  // last CTA: this rank's maximum is complete
  double m2 = __longlong_as_double((long long)__ldcg(reinterpret_cast<const unsigned long long*>(&ctrl->aux[0])));
  if (peer.world > 1) {  // sharded swarm: max-reduce over the ranks through the peer mailboxes
    const int par = it & 1;
    for (int r = threadIdx.x; r < peer.world; r += blockDim.x) *peer_rad(peer, r, par, peer.rank) = m2;
    const bool ok = peer_exchange_flags(peer, kPeerRadius, par, (uint32_t)it);
    if (threadIdx.x != 0) return;
    if (!ok) {
      ctrl->status = SP_STATUS_PEER_TIMEOUT;
      ctrl->flag = 0;
      ctrl->done_blocks = 0;
      return;
    }
    m2 = 0.0;
    for (int r = 0; r < peer.world; ++r) m2 = fmax(m2, ld_volatile(peer_rad(peer, peer.rank, par, r)));
  }
  if (threadIdx.x != 0) return;
  const double radius = sqrt(m2) / sqrt(4.0 * (double)N);
  int nw = 0;
  if (radius < delta) {
    const double inorm = (double)it / (double)maxiter;
    nw = (int)(((double)Ptot - 1.0) / (1.0 + exp(1.0 / 0.09 * (inorm - gamma + 0.5))));
    if (nw < 0) nw = 0;
  }
  ctrl->flag = nw;
  ctrl->aux[1] = radius;
  ctrl->aux[0] = 0.0;
  ctrl->done_blocks = 0;
  if (pause && nw > 0) ctrl->status = SP_STATUS_RESTART_PENDING;
}

__global__ void restart_resume_kernel(sp_ctrl* ctrl) {
  if (ctrl->status == SP_STATUS_RESTART_PENDING) ctrl->status = SP_RUNNING;
  ctrl->flag = 0;
}

// (2) decision: radius < delta -> nw rows to reset (ctrl->flag), ctrl->aux[1] = radius
__global__ void restart_plan_kernel(sp_ctrl* ctrl, int64_t P, int N, int it, int maxiter, double gamma, double delta,
                                    int parked = 0) {  // P = whole swarm
  if (parked) {  // exact decision for a run the bound could not decide (flag == -1); a decided one keeps its nw
    if (ctrl->status != SP_STATUS_RESTART_PENDING || ctrl->flag != -1) return;
  } else if (!running(ctrl)) {
    ctrl->flag = 0;
    return;
  }
  const double radius = sqrt(ctrl->aux[0]) / sqrt(4.0 * (double)N);
  int nw = 0;
  if (radius < delta) {
    const double inorm = (double)it / (double)maxiter;
    nw = (int)(((double)P - 1.0) / (1.0 + exp(1.0 / 0.09 * (inorm - gamma + 0.5))));
    if (nw < 0) nw = 0;
  }
  ctrl->flag = nw;
  ctrl->aux[1] = radius;
  ctrl->aux[0] = 0.0;
}

// (2') sharded swarm: the local maxima travel through the peer mailboxes first (max-reduce)
__global__ void restart_plan_peer_kernel(sp_ctrl* ctrl, int64_t P, int N, int it, int maxiter, double gamma,
                                         double delta, const PeerArgs p) {
  if (!running(ctrl)) {
    if (threadIdx.x == 0) ctrl->flag = 0;
    return;
  }
  const int par = it & 1;
  const double mine = ctrl->aux[0];
  for (int r = threadIdx.x; r < p.world; r += blockDim.x) *peer_rad(p, r, par, p.rank) = mine;
  const bool ok = peer_exchange_flags(p, kPeerRadius, par, (uint32_t)it);
  if (threadIdx.x != 0) return;
  if (!ok) {
    ctrl->status = SP_STATUS_PEER_TIMEOUT;
    ctrl->flag = 0;
    return;
  }
  double m = 0.0;
  for (int r = 0; r < p.world; ++r) m = fmax(m, ld_volatile(peer_rad(p, p.rank, par, r)));
  const double radius = sqrt(m) / sqrt(4.0 * (double)N);
  int nw = 0;
  if (radius < delta) {
    const double inorm = (double)it / (double)maxiter;
    nw = (int)(((double)P - 1.0) / (1.0 + exp(1.0 / 0.09 * (inorm - gamma + 0.5))));
    if (nw < 0) nw = 0;
  }
  ctrl->flag = nw;
  ctrl->aux[1] = radius;
  ctrl->aux[0] = 0.0;
}

// (3') restart fired (ctrl->flag > 0 on every rank alike): all-gather the pbestfit shards by
// storing this rank's shard into every peer's mailbox, so each rank can rank the whole swarm
template <typename T>
__global__ void __launch_bounds__(1024)
peer_gather_fit_kernel(const T* __restrict__ pbestfit, int64_t P, int64_t row0, sp_ctrl* ctrl, int it,
                       const PeerArgs p) {
  if ((!running(ctrl) && ctrl->status != SP_STATUS_RESTART_PENDING) || ctrl->flag <= 0) return;
  for (int r = 0; r < p.world; ++r) {
    T* dst = peer_fit<T>(p, r) + row0;
    for (int64_t i = threadIdx.x; i < P; i += blockDim.x) dst[i] = pbestfit[i];
  }
  if (!peer_exchange_flags(p, kPeerFit, it & 1, (uint32_t)it) && threadIdx.x == 0) {
    ctrl->status = SP_STATUS_PEER_TIMEOUT;
    ctrl->flag = 0;
  }
}

// (4) reset the nw worst: V = 0, X = U(lower, upper), pbest = X, pbestfit = 1e30
// A warp per row: one rank load decides for the whole row (most rows are left alone late in a run), the draws come
// one Philox call per 16-byte vector (the element-per-thread version recomputed the call for each of its VEC
// elements and divided every index by N: 10 us at P = 32768 x 64, 32 us per 131072 x 64 shard).
template <typename T>
__global__ void __launch_bounds__(kThreads)
restart_apply_kernel(T* X, T* V, T* pbest, T* pbestfit, const int32_t* __restrict__ rank, const T* __restrict__ lower,
                     const T* __restrict__ upper, int64_t P, int N, int64_t ld, int it, uint64_t seed,
                     const T* __restrict__ fresh, const sp_ctrl* ctrl, int64_t Ptot, int64_t row0) {
  constexpr int VEC = Num<T>::VEC;
  using Vt = typename Num<T>::vec_t;
  const int nw = ctrl->flag;
  if (nw <= 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  for (int64_t i = warp; i < P; i += nwarps) {
    const int64_t r = rank[i];
    if (r < Ptot - nw) continue;
    for (int j0 = lane * VEC; j0 < (int)ld; j0 += 32 * VEC) {  // ld is a multiple of VEC: whole vectors, padding stays 0
      T val[VEC];
      if (fresh != nullptr) {
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          val[e] = j0 + e < N ? fresh[(Ptot - 1 - r) * ld + j0 + e] : T(0);  // reset order: worst first (argsort()[:-nw-1:-1])
      } else {
        T blk[VEC];
        uniform_block(philox4x32((uint32_t)(j0 / VEC), (uint32_t)(row0 + i), (uint32_t)it, kPsoRestart, seed), blk);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          val[e] = j0 + e < N ? add_rn(lower[j0 + e], mul_rn(sub_rn(upper[j0 + e], lower[j0 + e]), blk[e])) : T(0);
      }
      Vt v, z;
      T* pv = reinterpret_cast<T*>(&v);
      T* pz = reinterpret_cast<T*>(&z);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        pv[e] = val[e];
        pz[e] = T(0);
      }
      *reinterpret_cast<Vt*>(X + i * ld + j0) = v;
      *reinterpret_cast<Vt*>(pbest + i * ld + j0) = v;
      *reinterpret_cast<Vt*>(V + i * ld + j0) = z;
    }
    if (lane == 0) pbestfit[i] = T(1.0e30);
  }
}

template <typename T>
static PeerArgs peer_args(const sp_pso_state* st) {
  PeerArgs p = {};
  if (st->shard == 2) {
    p.peers = reinterpret_cast<unsigned char* const*>(st->peers);
    p.world = st->world;
    p.rank = st->rank;
    p.L = peer_layout(st->world, st->ld, st->P_total, sizeof(T));
  }
  return p;
}

static bool pso_chainable(const sp_pso_state* st) {
  // fp32 only: the fp64 kernel (76-80 registers, 3 CTAs per SM, two waves) measured slower chained
  // (C3 fp64: 1.28e9 -> 1.15e9 evals/s), the fp32 one faster (1.83e9 -> 2.04e9)
  return st != nullptr && st->dtype == SP_F32 && st->chain_rows != nullptr && st->shard == 0 && st->gamma < 0.0 &&
         st->r1 == nullptr &&
         st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG;
}

// the chained instantiation exists for fp32 only (pso_chainable)
template <int C, int L>
static void launch_chained(const PsoArgs<float>& a, int grid, cudaStream_t s, bool pdl, const PhiloxKeys& keys) {
  launch_pdl(pso_generation_kernel<float, C, L, true, true>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);
}
template <int C, int L>
static void launch_chained(const PsoArgs<double>&, int, cudaStream_t, bool, const PhiloxKeys&) {}

// PLAIN instantiations with a compile-time objective: the shapes and objectives of the BASELINE configurations
// (C3: Styblinski-Tang ndim 64; also Rastrigin / Rosenbrock, ndim 64 / 128 in fp32 and 64 in fp64).
// SP_PSO_GENERIC=1 keeps the generic kernel (profiling / parity switch).  Returns false if not applicable.
template <typename T, int C, int L, int OBJ>
static void launch_plain(const PsoArgs<T>& a, int grid, cudaStream_t s, bool pdl, const PhiloxKeys& keys) {
  if constexpr (std::is_same<T, float>::value) {
    if (a.chain != 0) {
      launch_pdl(pso_generation_kernel<T, C, L, true, true, OBJ, true>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);
      return;
    }
  }
  launch_pdl(pso_generation_kernel<T, C, L, true, false, OBJ, true>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);
}
template <typename T, int C, int L>
static bool pso_launch_plain(const PsoArgs<T>& a, int grid, cudaStream_t s, bool pdl, const PhiloxKeys& keys) {
  constexpr bool shape_ok = C == 1 && (L == 32 || (L == 16 && std::is_same<T, float>::value));
  if constexpr (shape_ok) {
    static const bool generic = getenv("SP_PSO_GENERIC") != nullptr;
    using TL = Tile<T, C, L>;
    if (generic || a.constraint != SP_CONS_NONE || a.propose_only || a.N != TL::COLS || a.ld != a.N ||
        a.P % TL::RPW != 0 || a.r1 != nullptr)
      return false;
    switch (a.objective) {
      case SP_OBJ_STYBLINSKI_TANG: launch_plain<T, C, L, SP_OBJ_STYBLINSKI_TANG>(a, grid, s, pdl, keys); return true;
      case SP_OBJ_RASTRIGIN: launch_plain<T, C, L, SP_OBJ_RASTRIGIN>(a, grid, s, pdl, keys); return true;
      case SP_OBJ_ROSENBROCK: launch_plain<T, C, L, SP_OBJ_ROSENBROCK>(a, grid, s, pdl, keys); return true;
      default: return false;
    }
  }
  return false;
}

template <typename T>
static int pso_launch(const sp_pso_state* st, int it, int propose_only, cudaStream_t s, int chain = 0,
                      bool after_kernel = false, bool rbound = false) {
  Shape sh;
  if (!pick_shape(st->N, Num<T>::VEC, &sh)) {
    set_error("sp_pso_generation: ndim %d exceeds the compiled row shapes", st->N);
    return SP_ERR_SHAPE;
  }
  PsoArgs<T> a;
  a.objective = st->objective;
  a.constraint = st->constraint;
  a.it = it;
  a.maxiter = st->maxiter;
  a.N = st->N;
  a.propose_only = propose_only;
  a.P = st->P;
  a.ld = st->ld;
  a.w = (T)st->w;
  a.c1 = (T)st->c1;
  a.c2 = (T)st->c2;
  a.xtol = st->xtol;
  a.ftol = st->ftol;
  a.seed = st->seed;
  a.X = (T*)st->X;
  a.V = (T*)st->V;
  a.pbest = (T*)st->pbest;
  a.pbestfit = (T*)st->pbestfit;
  a.pfit = (T*)st->pfit;
  a.gbest = (T*)st->gbest;
  a.lower = (const T*)st->lower;
  a.upper = (const T*)st->upper;
  a.ctrl = st->ctrl;
  a.scratch = (Best*)st->scratch;
  a.r1 = (const T*)st->r1;
  a.r2 = (const T*)st->r2;
  a.row0 = st->shard ? st->row0 : 0;
  a.P_total = st->shard ? st->P_total : st->P;
  a.xch = (T*)st->xch;
  a.shard = st->shard;
  a.peer = peer_args<T>(st);
  a.chain = propose_only ? 0 : chain;
  a.chain_rows = (T*)st->chain_rows;
  a.rbound = rbound ? (getenv("SP_CPSO_FORCE_AMBIGUOUS") != nullptr ? 2 : 1) : 0;  // 2: test hook, see restart_decide_bound
  a.gamma = st->gamma;
  a.delta = st->delta;
  const bool philox = st->r1 == nullptr;
  const PhiloxKeys keys = philox_keys(st->seed);
  int grid = grid_for_rows(st->P, sh.lpr, sh.ch >= 4 ? 2 : 4);
  if (a.chain != 0 && grid > kChainRegion) grid = kChainRegion;  // one record + row slot per CTA
  const bool pdl = after_kernel || (a.chain & SP_CHAIN_IN) != 0;  // follows another kernel of the chain directly
#define SP_CALL(TT, C, L)                                                                                    \
  do {                                                                                                       \
    if (pso_launch_plain<TT, C, L>(a, grid, s, pdl, keys)) break;                                            \
    if (a.chain != 0)  /* pso_chainable(): fp32 only */                                                      \
      launch_chained<C, L>(a, grid, s, pdl, keys);                                                           \
    else if (philox)                                                                                         \
      launch_pdl(pso_generation_kernel<TT, C, L, true, false>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);  \
    else                                                                                                     \
      launch_pdl(pso_generation_kernel<TT, C, L, false, false>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys); \
  } while (0)
  SP_DISPATCH_SHAPE(T, sh, SP_CALL);
#undef SP_CALL
  SP_CHECK_LAUNCH();
  return SP_OK;
}

static int pso_check(const sp_pso_state* st, int it) {
  SP_CHECK_ARG(st != nullptr, "null state");
  SP_CHECK_ARG(st->dtype == SP_F32 || st->dtype == SP_F64, "dtype");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || st->constraint == SP_CONS_SHRINK, "constraint");
  SP_CHECK_ARG(st->N >= 1 && st->P >= 2 && st->P < (1LL << 31), "popsize / ndim");
  const int vec = st->dtype == SP_F32 ? 4 : 2;
  SP_CHECK_ARG(st->ld >= st->N && st->ld % vec == 0, "ld must be a multiple of 16/sizeof(T)");
  SP_CHECK_ARG(st->X && st->V && st->pbest && st->pbestfit && st->pfit && st->gbest && st->ctrl && st->scratch,
               "null buffer");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || (st->lower && st->upper), "bounds needed for Shrink");
  SP_CHECK_ARG((st->r1 == nullptr) == (st->r2 == nullptr), "r1 and r2 must come together");
  SP_CHECK_ARG(st->shard >= 0 && st->shard <= 2, "shard");
  SP_CHECK_ARG(!st->shard || (st->r1 == nullptr && st->row0 >= 0 && st->row0 + st->P <= st->P_total),
               "shard mode needs in-kernel draws and a row range inside the swarm");
  SP_CHECK_ARG(st->shard != 1 || st->xch, "shard == 1 needs the exchange record");
  SP_CHECK_ARG(st->shard != 2 || (st->mailbox && st->peers && st->world >= 1 && st->world <= 64 && st->rank >= 0 &&
                                  st->rank < st->world),
               "shard == 2 needs the peer mailboxes (sp_peer_alloc / sp_peer_open), world <= 64");
  SP_CHECK_ARG(it >= 2, "generation index starts at 2 (_cpso.py:256-258)");
  return SP_OK;
}

template <typename T>
static int restart_plan_launch(const sp_pso_state* st, int it, int32_t* rank, cudaStream_t s) {
  {  // radius + decision fused (the last CTA decides): one launch instead of two
    const int vec = Num<T>::VEC;
    int lpr = 1;
    while (lpr < 32 && lpr * vec < st->ld) lpr <<= 1;
    const int grid = grid_for_rows(st->P, lpr, 4);
    radius_plan_kernel<T><<<grid, kThreads, 0, s>>>((const T*)st->X, (const T*)st->gbest, st->P, st->N, st->ld, st->ctrl, it,
                                                  st->maxiter, st->gamma, st->delta, lpr, 0, st->P, PeerArgs{});
    SP_CHECK_LAUNCH();
  }
  if (rank_launch<T>((const T*)st->pbestfit, st->P, rank, &st->ctrl->flag, s) != cudaSuccess) return SP_ERR_CUDA;
  return SP_OK;
}

template <typename T>
static int restart_apply_launch(const sp_pso_state* st, int it, const int32_t* rank, const void* fresh,
                                cudaStream_t s) {
  int64_t need = (st->P + kThreads / 32 - 1) / (kThreads / 32), cap = (int64_t)sm_count() * 8;
  restart_apply_kernel<T><<<(int)(need < cap ? need : cap), kThreads, 0, s>>>(
      (T*)st->X, (T*)st->V, (T*)st->pbest, (T*)st->pbestfit, rank, (const T*)st->lower, (const T*)st->upper, st->P,
      st->N, st->ld, it, st->seed, (const T*)fresh, st->ctrl, st->shard ? st->P_total : st->P, st->shard ? st->row0 : 0);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

template <typename T>
static int pso_run_sharded(const sp_pso_state* st, int it_first, int n, int32_t* rank_all, cudaStream_t s) {
  const PeerArgs p = peer_args<T>(st);
  const T* fit_all = reinterpret_cast<const T*>(static_cast<const unsigned char*>(st->mailbox) + p.L.fit);
  // profiling switch (results are then wrong, timing only): bit 0 skips radius + decision, 1 the pbestfit
  // all-gather, 2 the ranking, 3 the reset
  static const int skip = getenv("SP_SHARD_SKIP") != nullptr ? atoi(getenv("SP_SHARD_SKIP")) : 0;
  for (int g = 0; g < n; ++g) {
    const int it = it_first + g;
    int rc = pso_launch<T>(st, it, 0, s);
    if (rc) return rc;
    if (st->gamma < 0.0) continue;
    if (!(skip & 1)) {  // radius + max-reduce over the ranks + decision in ONE kernel (the last CTA exchanges), launched programmatically
      const int vec = Num<T>::VEC;
      int lpr = 1;
      while (lpr < 32 && lpr * vec < st->ld) lpr <<= 1;
      const int grid = grid_for_rows(st->P, lpr, 4);
      cudaError_t e = launch_pdl(radius_plan_kernel<T>, dim3(grid), dim3(kThreads), 0, s, true, (const T*)st->X,
                                 (const T*)st->gbest, st->P, st->N, st->ld, st->ctrl, it, st->maxiter, st->gamma,
                                 st->delta, lpr, 0, st->P_total, p);
      if (e != cudaSuccess) {
        set_error("sp_pso_run_sharded: %s", cudaGetErrorString(e));
        return SP_ERR_CUDA;
      }
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (!(skip & 2)) {
      peer_gather_fit_kernel<T><<<1, 1024, 0, s>>>((const T*)st->pbestfit, st->P, st->row0, st->ctrl, it, p);
      SP_CHECK_LAUNCH();
    }
    // every rank sorts all chunks of the all-gathered fitness but merges only the chunks of its own rows
    if (!(skip & 4) &&
        rank_launch<T>(fit_all, st->P_total, rank_all, &st->ctrl->flag, s, nullptr, st->row0, st->P) != cudaSuccess)
      return SP_ERR_CUDA;
    if (!(skip & 8)) {
      rc = restart_apply_launch<T>(st, it, rank_all + st->row0, nullptr, s);
      if (rc) return rc;
    }
  }
  return SP_OK;
}

}  // namespace sp

extern "C" int sp_pso_run_sharded(const sp_pso_state* st, int it_first, int n, int32_t* rank_all, void* stream) {
  using namespace sp;
  int rc = pso_check(st, it_first);
  if (rc) return rc;
  SP_CHECK_ARG(st->shard == 2, "sp_pso_run_sharded needs shard == 2");
  SP_CHECK_ARG(st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG, "device objective required");
  SP_CHECK_ARG(st->gamma < 0.0 || (rank_all != nullptr && st->lower && st->upper), "rank scratch and bounds");
  return st->dtype == SP_F32 ? pso_run_sharded<float>(st, it_first, n, rank_all, (cudaStream_t)stream)
                             : pso_run_sharded<double>(st, it_first, n, rank_all, (cudaStream_t)stream);
}

namespace sp {
// ranking + reset of a parked restart; sharded: all-gather the pbestfit shards over the mailboxes first
template <typename T>
static int restart_resume(const sp_pso_state* st, int it, int32_t* rank, cudaStream_t s) {
  const T* fit = (const T*)st->pbestfit;
  int64_t n = st->P;
  const int32_t* mine = rank;
  if (st->shard == 0) {  // a run parked by the bound decision with flag == -1: the exact radius decides (no-ops otherwise)
    const int grid = grid_for_rows(st->P, 32, 8);
    radius_kernel<T><<<grid, kThreads, 0, s>>>((const T*)st->X, (const T*)st->gbest, st->P, st->N, st->ld, st->ctrl, 1);
    SP_CHECK_LAUNCH();
    restart_plan_kernel<<<1, 1, 0, s>>>(st->ctrl, st->P, st->N, it, st->maxiter, st->gamma, st->delta, 1);
    SP_CHECK_LAUNCH();
  }
  if (st->shard == 2) {
    const PeerArgs p = peer_args<T>(st);
    peer_gather_fit_kernel<T><<<1, 1024, 0, s>>>((const T*)st->pbestfit, st->P, st->row0, st->ctrl, it, p);
    SP_CHECK_LAUNCH();
    fit = reinterpret_cast<const T*>(static_cast<const unsigned char*>(st->mailbox) + p.L.fit);
    n = st->P_total;
    mine = rank + st->row0;
  }
  if (rank_launch<T>(fit, n, rank, &st->ctrl->flag, s, nullptr, st->shard == 2 ? st->row0 : 0, st->shard == 2 ? st->P : -1) !=
      cudaSuccess) {
    set_error("sp_cpso_restart_resume: ranking failed");
    return SP_ERR_CUDA;
  }
  return restart_apply_launch<T>(st, it, mine, nullptr, s);
}

template <typename T>
static int pso_run_lazy(const sp_pso_state* st, int it_first, int n, cudaStream_t s) {
  const int vec = Num<T>::VEC;
  int lpr = 1;
  while (lpr < 32 && lpr * vec < st->ld) lpr <<= 1;
  const int grid = grid_for_rows(st->P, lpr, 4);
  // whole swarm: the restart decision rides the generation kernel's epilogue (restart_decide_bound): ONE launch per
  // generation; SP_CPSO_EXACT_RADIUS=1 keeps the separate radius + decision kernel (profiling / parity switch)
  static const bool exact = getenv("SP_CPSO_EXACT_RADIUS") != nullptr;
  const bool bound = st->shard == 0 && !exact;
  for (int g = 0; g < n; ++g) {
    const int it = it_first + g;
    int rc = pso_launch<T>(st, it, 0, s, 0, g > 0, bound);
    if (rc) return rc;
    if (bound) continue;
    cudaError_t e = launch_pdl(radius_plan_kernel<T>, dim3(grid), dim3(kThreads), 0, s, true, (const T*)st->X,
                               (const T*)st->gbest, st->P, st->N, st->ld, st->ctrl, it, st->maxiter, st->gamma,
                               st->delta, lpr, 1, st->shard ? st->P_total : st->P, peer_args<T>(st));
    if (e != cudaSuccess) {
      set_error("sp_pso_run_lazy: %s", cudaGetErrorString(e));
      return SP_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  return SP_OK;
}

}  // namespace sp

using namespace sp;

extern "C" {

int sp_pso_generation(const sp_pso_state* st, int it, void* stream) {
  int rc = pso_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG,
               "device objective required (use sp_pso_propose + sp_select_sync for host objectives)");
  return st->dtype == SP_F32 ? pso_launch<float>(st, it, 0, (cudaStream_t)stream)
                             : pso_launch<double>(st, it, 0, (cudaStream_t)stream);
}

int sp_pso_propose(const sp_pso_state* st, int it, void* stream) {
  int rc = pso_check(st, it);
  if (rc) return rc;
  return st->dtype == SP_F32 ? pso_launch<float>(st, it, 1, (cudaStream_t)stream)
                             : pso_launch<double>(st, it, 1, (cudaStream_t)stream);
}

int sp_cpso_restart_plan(const sp_pso_state* st, int it, int32_t* rank, void* stream) {
  int rc = pso_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(rank != nullptr && st->lower && st->upper && st->gamma >= 0.0, "rank scratch, bounds, competitivity");
  return st->dtype == SP_F32 ? restart_plan_launch<float>(st, it, rank, (cudaStream_t)stream)
                             : restart_plan_launch<double>(st, it, rank, (cudaStream_t)stream);
}

int sp_cpso_radius(const sp_pso_state* st, int it, void* stream) {
  int rc = pso_check(st, it);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for_rows(st->P, 32, 8);
  if (st->dtype == SP_F32)
    radius_kernel<float><<<grid, kThreads, 0, s>>>((const float*)st->X, (const float*)st->gbest, st->P, st->N, st->ld, st->ctrl);
  else
    radius_kernel<double><<<grid, kThreads, 0, s>>>((const double*)st->X, (const double*)st->gbest, st->P, st->N, st->ld, st->ctrl);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_cpso_decide(const sp_pso_state* st, int it, void* stream) {
  int rc = pso_check(st, it);
  if (rc) return rc;
  restart_plan_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(st->ctrl, st->shard ? st->P_total : st->P, st->N, it,
                                                         st->maxiter, st->gamma, st->delta);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_cpso_restart_apply(const sp_pso_state* st, int it, const int32_t* rank, const void* fresh, void* stream) {
  int rc = pso_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(rank != nullptr && st->lower && st->upper, "rank scratch and bounds");
  return st->dtype == SP_F32 ? restart_apply_launch<float>(st, it, rank, fresh, (cudaStream_t)stream)
                             : restart_apply_launch<double>(st, it, rank, fresh, (cudaStream_t)stream);
}

int sp_cpso_restart(const sp_pso_state* st, int it, int32_t* rank, void* stream) {
  int rc = sp_cpso_restart_plan(st, it, rank, stream);
  if (rc) return rc;
  return sp_cpso_restart_apply(st, it, rank, nullptr, stream);
}

int64_t sp_pso_chain_scalars(int64_t ld) { return 2 * (int64_t)kChainRegion * ld; }

int sp_pso_run_lazy(const sp_pso_state* st, int it_first, int n, void* stream) {
  int rc = pso_check(st, it_first);
  if (rc) return rc;
  SP_CHECK_ARG(st->r1 == nullptr && (st->shard == 0 || st->shard == 2) && st->gamma >= 0.0 && st->lower && st->upper,
               "whole swarm or peer-sharded swarm, in-kernel draws, competitivity and bounds");
  SP_CHECK_ARG(st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG, "device objective required");
  return st->dtype == SP_F32 ? pso_run_lazy<float>(st, it_first, n, (cudaStream_t)stream)
                             : pso_run_lazy<double>(st, it_first, n, (cudaStream_t)stream);
}

int sp_cpso_restart_resume(const sp_pso_state* st, int it, int32_t* rank, void* stream) {
  int rc = pso_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(rank != nullptr && st->lower && st->upper && (st->shard == 0 || st->shard == 2),
               "rank scratch (P_total entries when sharded), bounds, whole or peer-sharded swarm");
  cudaStream_t s = (cudaStream_t)stream;
  rc = st->dtype == SP_F32 ? restart_resume<float>(st, it, rank, s) : restart_resume<double>(st, it, rank, s);
  if (rc) return rc;
  restart_resume_kernel<<<1, 1, 0, s>>>(st->ctrl);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_pso_run(const sp_pso_state* st, int it_first, int n, int32_t* rank, void* stream) {
  SP_CHECK_ARG(st != nullptr && st->r1 == nullptr, "sp_pso_run needs in-kernel draws");
  if (pso_chainable(st)) {  // plain PSO: only the last generation of the chunk runs the last-CTA epilogue
    int rc = pso_check(st, it_first);
    if (rc) return rc;
    for (int g = 0; g < n; ++g) {
      const int flags = (g > 0 ? SP_CHAIN_IN : 0) | (g < n - 1 ? SP_CHAIN_OUT : 0);
      rc = st->dtype == SP_F32 ? pso_launch<float>(st, it_first + g, 0, (cudaStream_t)stream, flags)
                               : pso_launch<double>(st, it_first + g, 0, (cudaStream_t)stream, flags);
      if (rc) return rc;
    }
    return SP_OK;
  }
  for (int g = 0; g < n; ++g) {
    int rc = sp_pso_generation(st, it_first + g, stream);
    if (rc) return rc;
    if (st->gamma >= 0.0) {
      rc = sp_cpso_restart(st, it_first + g, rank, stream);
      if (rc) return rc;
    }
  }
  return SP_OK;
}

}  // extern "C"
