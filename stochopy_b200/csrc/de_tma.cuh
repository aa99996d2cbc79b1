// DE generation kernel for full-warp rows (32 lanes per row): a per-SM row pool.
//
// One CTA per SM owns a contiguous slice of the population (about P/148 rows):
//   phase 0  every thread draws the donors + forced crossover column of one row of
//            the slice (one Philox call per row) into a per-row record table in shared
//            memory (fixed offset: every access in the row loop is [row * stride + imm])
//            and stages pbestfit for the slice;
//   phase 1  warps claim rows from a shared counter -- fast warps take more rows, so the
//            whole SM stays busy until the slice is empty.  Two ways to move the rows:
//            * register pipeline (CH == 1, K <= 3; the headline shape): rows are claimed in
//              PAIRS; while row A is processed, the own row + K donor rows of row B travel
//              into a second register set with plain 16-byte loads;
//            * TMA ring (wider rows, 4 / 5 donors): S claimed rows in flight in a
//              warp-private ring, a stage = the K donor rows fetched by cp.async.bulk
//              (1-D TMA, SASS UBLKCP) onto the stage's mbarrier, the own row one row ahead
//              through registers;
//            mutant -> crossover (16-bit pieces of one Philox2x32-10 call per 4 columns,
//            philox.cuh) -> repair -> objective -> strict-< selection -> 16-byte row store;
//   phase 2  pbestfit / pfit of the slice are written back coalesced, the CTA's
//            (fitness, row) minimum goes to scratch and the last CTA finalises (or, chained,
//            the next launch's prologue does).
// Strategy is a template parameter (donor count, mutant formula static); for the register
// pipeline's FULL + PLAIN variant the objective can be one too (Rosenbrock, Rastrigin).
// FULL: ndim == 32 * VEC * CH == ld, so no padding masks are needed.
#pragma once
#include "de_common.cuh"
#include "tma.cuh"

namespace sp {

template <int STRAT>
struct Strat {
  static constexpr int K = STRAT == SP_DE_RAND1BIN ? 3 : STRAT == SP_DE_RAND2BIN ? 5 : STRAT == SP_DE_BEST1BIN ? 2 : 4;
  static constexpr bool kBest = STRAT >= SP_DE_BEST1BIN;
};



// own row + K donor rows of one individual, in registers
template <typename T, int CH, int K>
struct RowSet {
  Tile<T, CH, 32> x, d[K];
};

// mutant, de/_strategy.py:1-38 (numpy's operation order, no FMA contraction)
template <typename T, int CH, int STRAT>
__device__ __forceinline__ void mutant(const RowSet<T, CH, Strat<STRAT>::K>& r, const Tile<T, CH, 32>& gb, T F,
                                       Tile<T, CH, 32>& u) {
  constexpr int VEC = Num<T>::VEC;
  constexpr int K = Strat<STRAT>::K;
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (STRAT == SP_DE_BEST1BIN)
        u.v[c][e] = add_rn(gb.v[c][e], mul_rn(F, sub_rn(r.d[0].v[c][e], r.d[1].v[c][e])));
      else if (STRAT == SP_DE_RAND1BIN)
        u.v[c][e] = add_rn(r.d[0].v[c][e], mul_rn(F, sub_rn(r.d[1].v[c][e], r.d[2 % K].v[c][e])));
      else if (STRAT == SP_DE_BEST2BIN)
        u.v[c][e] = add_rn(
            gb.v[c][e],
            mul_rn(F, sub_rn(sub_rn(add_rn(r.d[0].v[c][e], r.d[1].v[c][e]), r.d[2 % K].v[c][e]), r.d[3 % K].v[c][e])));
      else
        u.v[c][e] = add_rn(r.d[0].v[c][e],
                           mul_rn(F, sub_rn(sub_rn(add_rn(r.d[1].v[c][e], r.d[2 % K].v[c][e]), r.d[3 % K].v[c][e]),
                                            r.d[4 % K].v[c][e])));
    }
}

constexpr int kPoolStages = 3;

// Per-row record of the slice tables, at a FIXED shared-memory offset so every access in the row loop is
// `[row * stride + immediate]`: word 0 = forced crossover column, words 1..K = donor rows, then (aligned to
// sizeof(T)) the personal best fitness and the candidate fitness of this generation.
template <typename T, int K>
struct PoolRec {
  static constexpr int kT = (int)(sizeof(T) / 4);
  static constexpr int kBest = ((1 + K + kT - 1) / kT) * kT;
  static constexpr int kFnew = kBest + kT;
  static constexpr int kWords = ((kFnew + kT + 3) / 4) * 4;
  static constexpr int kBytes = kWords * 4;
};
constexpr uint32_t kPoolRecBase = 16;  // [0,16): claim counter

// shared-memory carve-up (bytes), identical on host and device
struct PoolLayout {
  uint32_t bars, queue, ring, total;
};
__host__ __device__ inline PoolLayout pool_layout(int warps, int nb, int rec_bytes, int K, int64_t ld, size_t elem,
                                                  bool with_ring = true) {
  PoolLayout L;
  auto up16 = [](uint32_t v) { return (v + 15u) & ~15u; };
  uint32_t o = up16(kPoolRecBase + (uint32_t)nb * (uint32_t)rec_bytes);
  L.bars = o;
  if (with_ring) o = up16(o + (uint32_t)warps * kPoolStages * 8);
  L.queue = o;
  if (with_ring) o = up16(o + (uint32_t)warps * kPoolStages * 4);
  o = (o + 127u) & ~127u;
  L.ring = o;
  if (with_ring) o += (uint32_t)warps * kPoolStages * (uint32_t)K * (uint32_t)ld * (uint32_t)elem;
  L.total = o;
  return L;
}

// PLAIN: no bound repair and not propose-only (the common case) -- those branches vanish.
// RING: donor rows through the TMA ring (any shape); !RING: own row and donors prefetched one
// row ahead with plain 16-byte loads into a second register set (CH == 1, K <= 3 only) -- the
// ring costs ~45 warp instructions per row (mbarrier wait, elected-lane issue, uniform-register
// traffic, queue) against ~10 for two address computations and loads, and with 32 warps per SM
// one row of lookahead (~2000 cycles) already covers the L2 / HBM latency.
// OBJ >= 0: the objective is a compile-time constant (no jump table in the row loop); -1: a.objective.
template <typename T, int CH, int STRAT, bool FULL, bool PLAIN, bool RING, int OBJ>
__global__ void __launch_bounds__(1024 / CH, 1)
de_pool_kernel(const DeArgs<T> a, int nb, const CrossKeys keys) {
  using TL = Tile<T, CH, 32>;
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  constexpr int K = Strat<STRAT>::K;
  constexpr int S = kPoolStages;
  using RS = RowSet<T, CH, K>;
  using Rec = PoolRec<T, K>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, warps = blockDim.x >> 5;
  const int ld = FULL ? TL::COLS : (int)a.ld;
  const int N = FULL ? TL::COLS : a.N;
  int* s_next = reinterpret_cast<int*>(smem_raw);
  unsigned char* recs = smem_raw + kPoolRecBase;  // compile-time offset: accesses are [row * kBytes + imm]
  auto rec_u32 = [&](int r, int w) -> uint32_t& { return *reinterpret_cast<uint32_t*>(recs + r * Rec::kBytes + 4 * w); };
  auto rec_best = [&](int r) -> T& { return *reinterpret_cast<T*>(recs + r * Rec::kBytes + 4 * Rec::kBest); };
  auto rec_fnew = [&](int r) -> T& { return *reinterpret_cast<T*>(recs + r * Rec::kBytes + 4 * Rec::kFnew); };
  const uint32_t row_bytes = (uint32_t)(ld * sizeof(T));
  const uint32_t stage_elems = (uint32_t)K * (uint32_t)ld;  // a stage = the K donor rows

  // slice of this CTA
  const int64_t q = a.P / gridDim.x, rem = a.P % gridDim.x;
  const int64_t b0 = blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
  const int rows = (int)(q + (blockIdx.x < rem ? 1 : 0));

  // ---- phase 0: tables ------------------------------------------------------------------
  // (an explicit cp.async.bulk.prefetch.L2 of the CTA's slice was measured slower when the L2 is
  // full of dirty lines: 38.3 vs 36.1 us per generation -- demand fetches are left alone)
  if (tid == 0) *s_next = 0;
  for (int t = tid; t < rows; t += blockDim.x) {
    uint32_t dd[5];
    int ir;
    draw_donors((uint32_t)(b0 + t), (uint32_t)a.P, K, a.it, a.seed, a.N, dd, &ir);
    rec_u32(t, 0) = (uint32_t)ir;
#pragma unroll
    for (int k = 0; k < K; ++k) rec_u32(t, 1 + k) = dd[k];
  }
  // Programmatic dependent launch: everything above depends only on (seed, generation), so
  // it overlaps the tail of the previous generation's kernel; from here on we read what that
  // kernel wrote (status, pbestfit, gbest, the population).
  pdl_wait();
  if (!running(a.ctrl)) return;
  const bool chain_in = (a.chain & SP_CHAIN_IN) != 0;
  __shared__ Best s_top;
  if (chain_in && wib == 0) {  // best of the generation before, from its per-CTA minima
    const Best b = chain_best(chain_region(a.scratch, a.it - 1), (int)gridDim.x);
    if (lane == 0) s_top = b;
  }
  pdl_launch_dependents();  // the next generation may start its own prologue as SMs free up
  for (int t = tid; t < rows; t += blockDim.x) rec_best(t) = a.pbestfit[b0 + t];
  TL gb;
  if (Strat<STRAT>::kBest && !chain_in) gb.load(a.gbest, lane, ld);
  const T F = a.F;
  const uint32_t it = (uint32_t)a.it;
  __syncthreads();
  if (chain_in) {
    const Best top = s_top;
    // one warp of CTA 0 writes gbest / dist / nit / status of generation it-1 (off the critical path)
    if (blockIdx.x == 0 && wib == warps - 1)
      finalize_generation_warp<T>(top, a.Xold, a.ld, a.N, a.gbest, a.ctrl, a.it - 1, a.maxiter, a.xtol, a.ftol);
    if (chain_stops(top.f, a.it - 1, a.maxiter, a.ftol)) return;
    if (Strat<STRAT>::kBest) gb.load(a.Xold + top.row * (int64_t)ld, lane, ld);
  }

  // ---- phase 1: claim / fetch / process ----------------------------------------------------
  const uint32_t s_next_addr = smem_u32(s_next);
  auto claim = [&](uint32_t n) {  // one elected lane claims n rows (elect.sync: no warp-aggregation preamble)
    int r = 0;
    uint32_t leader = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync %1|p, 0xffffffff;\n"
        "@p atom.shared.add.u32 %0, [%2], %3;\n"
        "}\n"
        : "+r"(r), "+r"(leader)
        : "r"(s_next_addr), "r"(n)
        : "memory");
    return __shfl_sync(0xffffffffu, r, leader);
  };
  // crossover, repair, objective, selection and store of one individual (u: mutant, xi: own row)
  auto finish = [&](TL& u, const TL& xi, int r) {
    const int irand = (int)rec_u32(r, 0);
    const uint32_t row = (uint32_t)(b0 + r);
    // binomial crossover (_de.py:339-344) and Random repair (de/_constraints.py:22-26)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int j0 = TL::col(c, lane, 0);
      if (FULL || j0 < N) {
        bool t4[4];
        de_cross_take(row, (uint32_t)(j0 >> 2), keys, t4);
        const int d = irand - j0;  // the forced column, relative to this vector
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const bool tk = VEC == 4 ? t4[e] : ((j0 & 2) ? t4[2 + (e & 1)] : t4[e & 1]);
          const bool t = (tk || d == e) && (FULL || j0 + e < N);
          u.v[c][e] = t ? u.v[c][e] : xi.v[c][e];
        }
        if (!PLAIN && a.constraint == SP_CONS_RANDOM) {
          bool any = false;
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (FULL || j0 + e < N) any |= (u.v[c][e] < a.lower[j0 + e]) || (u.v[c][e] > a.upper[j0 + e]);
          if (any) {
            T qv[VEC];
            uniform_block(philox4x32((uint32_t)(j0 / VEC), row, it, kDeRepair, a.seed), qv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const int jj = j0 + e;
              if (FULL || jj < N) {
                const T lo = a.lower[jj], hi = a.upper[jj];
                if (u.v[c][e] < lo || u.v[c][e] > hi) u.v[c][e] = add_rn(lo, mul_rn(sub_rn(hi, lo), qv[e]));
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) u.v[c][e] = xi.v[c][e];
      }
    }

    T* out_row = a.Xnew + (int64_t)row * ld;
    if (!PLAIN && a.propose_only) {
      u.store(out_row, lane, ld);
      return;
    }
    const T f = evaluate_tile<T, CH, 32>(OBJ >= 0 ? OBJ : a.objective, u, lane, N);
    const bool win = f < rec_best(r);  // strict, _common.py:127
    if (win) u.store(out_row, lane, ld);  // two predicated 16-byte stores instead of a select per scalar
    else xi.store(out_row, lane, ld);
    // every lane holds the same f: same-address, same-value stores need no lane-0 branch; the warp barrier
    // orders them after every lane's read of the record (racecheck: intra-warp write-after-read)
    __syncwarp();
    rec_fnew(r) = f;
    if (win) rec_best(r) = f;
  };

  if (RING) {
    const PoolLayout L = pool_layout(warps, nb, Rec::kBytes, K, ld, sizeof(T), true);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars) + wib * S;
    int* queue = reinterpret_cast<int*>(smem_raw + L.queue) + wib * S;
    T* ring = reinterpret_cast<T*>(smem_raw + L.ring) + (size_t)wib * S * stage_elems;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
    auto issue = [&](int st, int r) {  // one elected lane fills stage `st` with the donors of row r
      if (lane == 0) {
        T* dst = ring + (size_t)st * stage_elems;
        mbar_expect_tx(&bars[st], (uint32_t)K * row_bytes);
#pragma unroll
        for (int k = 0; k < K; ++k)
          tma_load_row(dst + (size_t)k * ld, a.Xold + (int64_t)rec_u32(r, 1 + k) * ld, row_bytes, &bars[st]);
        queue[st] = r;
      }
    };
#pragma unroll
    for (int st = 0; st < S; ++st) {
      const int r = claim(1);
      if (r < rows) issue(st, r);
      else if (lane == 0) queue[st] = r;
    }
    __syncwarp();

    // the individual's own row travels through registers, one row ahead (a row takes a warp
    // a few thousand cycles on a full SM, which covers the HBM latency)
    TL own;
    if (queue[0] < rows) own.load(a.Xold + (b0 + queue[0]) * ld, lane, ld);

    for (uint32_t n = 0;; ++n) {
      const uint32_t st = n % S;
      const int r = queue[st];
      if (r >= rows) break;  // claims are monotonic: nothing further is in flight
      TL xi = own;
      {
        const int rn = queue[(n + 1) % S];
        if (rn < rows) own.load(a.Xold + (b0 + rn) * ld, lane, ld);
      }
      mbar_wait(&bars[st], (n / S) & 1u);
      const T* sx = ring + (size_t)st * stage_elems;
      auto lds = [&](TL& t, int slot) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int j0 = TL::col(c, lane, 0);
          if (FULL || j0 < ld) {
            V v = *reinterpret_cast<const V*>(sx + (size_t)slot * ld + j0);
            const T* p = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) t.v[c][e] = p[e];
          } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) t.v[c][e] = T(0);
          }
        }
      };
      TL u;
      {
        RS cur;
#pragma unroll
        for (int k = 0; k < K; ++k) lds(cur.d[k], k);
        mutant<T, CH, STRAT>(cur, gb, F, u);
      }
      __syncwarp();  // every lane holds its part of the stage: hand it back and refill it
      {
        const int r2 = claim(1);
        if (r2 < rows) issue(st, r2);
        else if (lane == 0) queue[st] = r2;
      }

      finish(u, xi, r);
    }
  } else {
    // rows travel through registers: while row A is processed, row B (own + K donor rows) is in flight.
    // Rows are claimed in PAIRS (2c, 2c+1): one shared atomic per two rows.
    // (measured: claiming two rows ahead and requesting the third row's lines with prefetch.global.L2
    // is slower, HBM-cold 30.1 vs 28.4 us per generation -- one row of lookahead is left alone)
    RS A, B;
    auto fetch = [&](RS& t, int r) {
      t.x.load(a.Xold + (b0 + r) * ld, lane, ld);
      if (K <= 3) {
        const uint4 rec = *reinterpret_cast<const uint4*>(recs + r * Rec::kBytes);  // (irand, d0, d1, d2)
        const uint32_t dn[3] = {rec.y, rec.z, rec.w};
#pragma unroll
        for (int k = 0; k < K; ++k) t.d[k].load(a.Xold + (int64_t)dn[k % 3] * ld, lane, ld);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) t.d[k].load(a.Xold + (int64_t)rec_u32(r, 1 + k) * ld, lane, ld);
      }
    };
    // (fetches are unconditional on a clamped row: a predicated fetch makes ptxas load into temporaries and
    // copy 12 registers per row; the price is one wasted row load per warp at the end of the slice)
    const int last = rows - 1;
    int ra = claim(2);
    fetch(A, ra < last ? ra : last);
    while (ra < rows) {
      const int rb = ra + 1;
      fetch(B, rb < last ? rb : last);
      {
        TL u;
        mutant<T, CH, STRAT>(A, gb, F, u);
        finish(u, A.x, ra);
      }
      if (rb >= rows) break;
      ra = claim(2);
      fetch(A, ra < last ? ra : last);
      {
        TL u;
        mutant<T, CH, STRAT>(B, gb, F, u);
        finish(u, B.x, rb);
      }
    }
  }
  if (!PLAIN && a.propose_only) return;

  // ---- phase 2: coalesced write-back + argmin --------------------------------------------------
  __syncthreads();
  Best mine{1.0 / 0.0, 0x7fffffffffffffffLL};
  for (int t = tid; t < rows; t += blockDim.x) {
    const T b = rec_best(t);
    a.pbestfit[b0 + t] = b;
    a.pfit[b0 + t] = rec_fnew(t);
    if (better((double)b, b0 + t, mine.f, mine.row)) mine = Best{(double)b, b0 + t};
  }
  if (a.chain & SP_CHAIN_OUT) {  // leave the CTA minimum for the next launch's prologue
    __shared__ Best s_red[32];
    const Best b = block_best(mine, s_red);
    if (tid == 0) chain_region(a.scratch, a.it)[blockIdx.x] = b;
    return;
  }
  Best top;
  if (grid_best(mine, a.scratch, a.ctrl, &top))
    finalize_generation<T>(top, a.Xnew, a.ld, a.N, a.gbest, a.ctrl, a.it, a.maxiter, a.xtol, a.ftol);
}

// launch shape: one CTA per SM, as many warps as the ring leaves room for
struct PoolShape {
  int grid, threads, nb;
  size_t smem;
};
template <int CH>
inline bool pool_shape(int64_t P, int rec_bytes, int K, int64_t ld, size_t elem, PoolShape* ps, bool with_ring = true) {
  const int sms = sm_count();
  int64_t grid = P < sms ? P : sms;
  const int nb = (int)((P + grid - 1) / grid);
  static const char* env_w = getenv("SP_DE_WARPS");  // profiling switch: warps per CTA (<= 32 / CH)
  int wmax = 32 / CH;
  if (env_w != nullptr && atoi(env_w) >= 1 && atoi(env_w) < wmax) wmax = atoi(env_w);
  for (int warps = wmax; warps >= 1; warps = warps > 1 && (warps & (warps - 1)) ? warps - 1 : warps >> 1) {
    const PoolLayout L = pool_layout(warps, nb, rec_bytes, K, ld, elem, with_ring);
    if (L.total <= 220 * 1024) {
      *ps = {(int)grid, warps * 32, nb, L.total};
      return true;
    }
  }
  return false;
}
inline int pool_rec_bytes(int K, size_t elem) {  // PoolRec<T, K>::kBytes without the types
  const int kt = (int)(elem / 4);
  const int best = ((1 + K + kt - 1) / kt) * kt;
  return ((best + 2 * kt + 3) / 4) * 16;
}

template <typename T, int CH, int STRAT, bool FULL, bool PLAIN, bool RING, int OBJ>
static cudaError_t de_pool_launch_v(const DeArgs<T>& a, cudaStream_t s) {
  auto kern = de_pool_kernel<T, CH, STRAT, FULL, PLAIN, RING, OBJ>;
  constexpr int K = Strat<STRAT>::K;
  PoolShape ps;
  if (!pool_shape<CH>(a.P, PoolRec<T, K>::kBytes, K, a.ld, sizeof(T), &ps, RING)) return cudaErrorInvalidConfiguration;
  static thread_local size_t configured[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (configured[dev] < ps.smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    configured[dev] = 220 * 1024;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ps.grid);
  cfg.blockDim = dim3(ps.threads);
  cfg.dynamicSmemBytes = ps.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a, ps.nb, de_cross_keys(a.seed, a.it, (double)a.CR));
}

// register-pipelined variant where two row sets fit the register budget, TMA ring otherwise
// (SP_DE_RING=1 forces the ring: profiling switch).  The register-pipelined FULL + PLAIN kernel is also
// instantiated with the objective as a compile-time constant for the two objectives of the BASELINE
// configurations (SP_DE_GENERIC_OBJ=1 forces the run-time switch: profiling / parity switch).
template <typename T, int CH, int STRAT, bool FULL, bool PLAIN>
static cudaError_t de_pool_launch(const DeArgs<T>& a, cudaStream_t s) {
  if constexpr (CH == 1 && Strat<STRAT>::K <= 3) {
    static const bool force_ring = getenv("SP_DE_RING") != nullptr;
    if (!force_ring) {
      if constexpr (FULL && PLAIN) {
        static const bool generic = getenv("SP_DE_GENERIC_OBJ") != nullptr;
        if (!generic && a.objective == SP_OBJ_ROSENBROCK)
          return de_pool_launch_v<T, CH, STRAT, FULL, PLAIN, false, SP_OBJ_ROSENBROCK>(a, s);
        if (!generic && a.objective == SP_OBJ_RASTRIGIN)
          return de_pool_launch_v<T, CH, STRAT, FULL, PLAIN, false, SP_OBJ_RASTRIGIN>(a, s);
      }
      return de_pool_launch_v<T, CH, STRAT, FULL, PLAIN, false, -1>(a, s);
    }
  }
  return de_pool_launch_v<T, CH, STRAT, FULL, PLAIN, true, -1>(a, s);
}

template <typename T, int CH, int STRAT>
static cudaError_t de_rows_pick(const DeArgs<T>& a, cudaStream_t s) {
  const bool full = a.N == Tile<T, CH, 32>::COLS && a.ld == a.N;
  const bool plain = a.constraint == SP_CONS_NONE && !a.propose_only;
  if (plain) return full ? de_pool_launch<T, CH, STRAT, true, true>(a, s) : de_pool_launch<T, CH, STRAT, false, true>(a, s);
  return full ? de_pool_launch<T, CH, STRAT, true, false>(a, s) : de_pool_launch<T, CH, STRAT, false, false>(a, s);
}

template <typename T, int STRAT>
static cudaError_t de_tma_by_ch(const DeArgs<T>& a, int ch, cudaStream_t s) {
  switch (ch) {
    case 1: return de_rows_pick<T, 1, STRAT>(a, s);
    case 2: return de_rows_pick<T, 2, STRAT>(a, s);
    case 4: return de_rows_pick<T, 4, STRAT>(a, s);
    case 8: return de_rows_pick<T, 8, STRAT>(a, s);
    default: return de_rows_pick<T, 16, STRAT>(a, s);
  }
}

// does the pool kernel have a launch shape for this problem? (host-side test used by the dispatcher)
inline bool de_pool_fits(int ch, int64_t P, int K, int64_t ld, size_t elem) {
  PoolShape ps;
  const int rb = pool_rec_bytes(K, elem);
  switch (ch) {
    case 1: return pool_shape<1>(P, rb, K, ld, elem, &ps);
    case 2: return pool_shape<2>(P, rb, K, ld, elem, &ps);
    case 4: return pool_shape<4>(P, rb, K, ld, elem, &ps);
    case 8: return pool_shape<8>(P, rb, K, ld, elem, &ps);
    default: return pool_shape<16>(P, rb, K, ld, elem, &ps);
  }
}

}  // namespace sp
