// DE generation kernels for full-warp rows (32 lanes per row).
//
// Each warp owns a contiguous, balanced range of rows.  Donor indices for up to
// 32 rows are drawn lane-parallel (one Philox call per lane, not per row);
// pbestfit is read and pbestfit/pfit are written lane-parallel (coalesced).  The
// crossover test `u <= CR` is done on the raw Philox words (crossover_cut).  The
// strategy is a template parameter: donor count and mutant formula are static.
//
// FETCH selects how a row and its K donor rows reach the registers:
//   kFetchTma  a warp-private S-stage ring in shared memory; a stage is filled by
//              K+1 cp.async.bulk copies (1-D TMA, SASS UBLKCP) completing on the
//              stage's mbarrier, S rows ahead, issued by the lane that drew the row
//   kFetchLdg  16-byte loads software-pipelined one row ahead in registers (a row
//              takes a warp a few thousand cycles, which covers the HBM latency)
// FULL: ndim == 32 * VEC * CH == ld, so no padding masks are needed.
#pragma once
#include "de_common.cuh"
#include "tma.cuh"

namespace sp {

constexpr int kFetchTma = 0, kFetchLdg = 1;

template <int STRAT>
struct Strat {
  static constexpr int K = STRAT == SP_DE_RAND1BIN ? 3 : STRAT == SP_DE_RAND2BIN ? 5 : STRAT == SP_DE_BEST1BIN ? 2 : 4;
  static constexpr bool kBest = STRAT >= SP_DE_BEST1BIN;
};

__device__ __forceinline__ bool cross_take(uint32_t w, uint64_t cut) { return w <= (uint32_t)cut; }
__device__ __forceinline__ bool cross_take(unsigned long long m53, uint64_t cut) { return m53 <= cut; }

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
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (STRAT == SP_DE_BEST1BIN)
        u.v[c][e] = add_rn(gb.v[c][e], mul_rn(F, sub_rn(r.d[0].v[c][e], r.d[1].v[c][e])));
      else if (STRAT == SP_DE_RAND1BIN)
        u.v[c][e] = add_rn(r.d[0].v[c][e], mul_rn(F, sub_rn(r.d[1].v[c][e], r.d[2 % Strat<STRAT>::K].v[c][e])));
      else if (STRAT == SP_DE_BEST2BIN)
        u.v[c][e] = add_rn(gb.v[c][e],
                           mul_rn(F, sub_rn(sub_rn(add_rn(r.d[0].v[c][e], r.d[1].v[c][e]), r.d[2 % Strat<STRAT>::K].v[c][e]),
                                            r.d[3 % Strat<STRAT>::K].v[c][e])));
      else
        u.v[c][e] = add_rn(r.d[0].v[c][e],
                           mul_rn(F, sub_rn(sub_rn(add_rn(r.d[1].v[c][e], r.d[2 % Strat<STRAT>::K].v[c][e]),
                                                   r.d[3 % Strat<STRAT>::K].v[c][e]),
                                            r.d[4 % Strat<STRAT>::K].v[c][e])));
    }
}

template <typename T, int CH, int STRAT, int S, int WARPS, int FETCH, bool FULL>
__global__ void __launch_bounds__(WARPS * 32, (CH == 1 ? 8 : CH == 2 ? 4 : CH == 4 ? 2 : 1))
de_rows_kernel(const DeArgs<T> a) {
  using TL = Tile<T, CH, 32>;
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  constexpr int K = Strat<STRAT>::K;
  using RS = RowSet<T, CH, K>;
  if (!running(a.ctrl)) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ld = FULL ? TL::COLS : (int)a.ld;
  const int N = FULL ? TL::COLS : a.N;
  const uint32_t row_bytes = (uint32_t)(ld * sizeof(T));
  const uint32_t stage_elems = (uint32_t)(K + 1) * (uint32_t)ld;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + wib * S;  // WARPS*S*8 <= 128 bytes
  T* ring = reinterpret_cast<T*>(smem_raw + 128) + (size_t)wib * S * stage_elems;

  if (FETCH == kFetchTma) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }

  const int64_t gw = (int64_t)blockIdx.x * WARPS + wib, GW = (int64_t)gridDim.x * WARPS;
  const int64_t q = a.P / GW, rem = a.P % GW;
  const int64_t r0 = gw * q + (gw < rem ? gw : rem);
  const int nrows = (int)(q + (gw < rem ? 1 : 0));

  TL gb;
  if (Strat<STRAT>::kBest) gb.load(a.gbest, lane, ld);
  const uint64_t cut = a.cr_cut;
  const T F = a.F;
  const uint32_t it = (uint32_t)a.it;

  Best mine{1.0 / 0.0, 0x7fffffffffffffffLL};
  uint32_t issued = 0, consumed = 0;  // TMA ring counters; parity = (n / S) & 1

  for (int base = 0; base < nrows; base += 32) {
    const int m = (nrows - base) < 32 ? (nrows - base) : 32;
    const int64_t my_row = r0 + base + lane;  // the row this lane draws for, fetches (TMA) and reports
    uint32_t d[K];
    int my_irand = 0;
    T my_best = T(0), my_f = T(0);
#pragma unroll
    for (int t = 0; t < K; ++t) d[t] = 0;
    if (lane < m) {
      uint32_t dd[5];
      draw_donors((uint32_t)my_row, (uint32_t)a.P, K, a.it, a.seed, a.N, dd, &my_irand);
#pragma unroll
      for (int t = 0; t < K; ++t) d[t] = dd[t];
      my_best = a.pbestfit[my_row];
    }

    // ---- fetch machinery ---------------------------------------------------------
    auto tma_issue = [&](int j) {  // lane j fills the next ring stage with its own rows
      const uint32_t st = issued % S;
      if (lane == j) {
        T* dst = ring + (size_t)st * stage_elems;
        mbar_expect_tx(&bars[st], (uint32_t)(K + 1) * row_bytes);
        tma_load_row(dst, a.Xold + my_row * ld, row_bytes, &bars[st]);
#pragma unroll
        for (int t = 0; t < K; ++t) tma_load_row(dst + (size_t)(t + 1) * ld, a.Xold + (int64_t)d[t] * ld, row_bytes, &bars[st]);
      }
      ++issued;
    };
    auto ldg_fetch = [&](RS& r, int j) {  // all lanes: rows of individual base+j straight to registers
      r.x.load(a.Xold + (r0 + base + j) * ld, lane, ld);
#pragma unroll
      for (int t = 0; t < K; ++t) {
        const uint32_t dj = __shfl_sync(0xffffffffu, d[t], j);
        r.d[t].load(a.Xold + (int64_t)dj * ld, lane, ld);
      }
    };
    auto ring_read = [&](RS& r) {  // wait for the oldest stage and copy this lane's part out
      const uint32_t st = consumed % S;
      mbar_wait(&bars[st], (consumed / S) & 1u);
      ++consumed;
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
      lds(r.x, 0);
#pragma unroll
      for (int t = 0; t < K; ++t) lds(r.d[t], t + 1);
    };

    RS nxt;
    if (FETCH == kFetchTma) {
#pragma unroll
      for (int j = 0; j < S; ++j)
        if (j < m) tma_issue(j);
    } else {
      ldg_fetch(nxt, 0);
    }

    T* out_row = a.Xnew + (r0 + base) * ld;
    for (int j = 0; j < m; ++j, out_row += ld) {
      const uint32_t row = (uint32_t)(r0 + base + j);
      const int irand = __shfl_sync(0xffffffffu, my_irand, j);
      TL xi, u;
      if (FETCH == kFetchTma) {
        RS cur;
        ring_read(cur);
        xi = cur.x;
        mutant<T, CH, STRAT>(cur, gb, F, u);
        __syncwarp();  // every lane holds its part of the stage: hand it back
        if (j + S < m) tma_issue(j + S);
      } else {
        xi = nxt.x;
        mutant<T, CH, STRAT>(nxt, gb, F, u);
        if (j + 1 < m) ldg_fetch(nxt, j + 1);  // next individual's rows are in flight during the rest
      }

      // binomial crossover (_de.py:339-344) and Random repair (de/_constraints.py:22-26)
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int j0 = TL::col(c, lane, 0);
        if (FULL || j0 < N) {
          const uint4 o = philox4x32((uint32_t)(j0 / VEC), row, it, kDeCross, a.seed);
          bool take[VEC];
          if (VEC == 4) {
            take[0] = cross_take(o.x, cut);
            take[1] = cross_take(o.y, cut);
            take[2 % VEC] = cross_take(o.z, cut);
            take[3 % VEC] = cross_take(o.w, cut);
          } else {
            take[0] = cross_take(((unsigned long long)o.x << 21) | (o.y >> 11), cut);
            take[1] = cross_take(((unsigned long long)o.z << 21) | (o.w >> 11), cut);
          }
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const int jj = j0 + e;
            const bool t = (take[e] || jj == irand) && (FULL || jj < N);
            u.v[c][e] = t ? u.v[c][e] : xi.v[c][e];
          }
          if (a.constraint == SP_CONS_RANDOM) {
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

      if (a.propose_only) {
        u.store(out_row, lane, ld);
        continue;
      }
      const T f = evaluate_tile<T, CH, 32>(a.objective, u, lane, N);
      const T old = __shfl_sync(0xffffffffu, my_best, j);
      const bool win = f < old;  // strict, _common.py:127
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) u.v[c][e] = win ? u.v[c][e] : xi.v[c][e];
      u.store(out_row, lane, ld);
      if (lane == j) {
        my_f = f;
        my_best = win ? f : old;
      }
    }
    if (!a.propose_only && lane < m) {  // coalesced report for the rows of this pass
      a.pbestfit[my_row] = my_best;
      a.pfit[my_row] = my_f;
      if (better((double)my_best, my_row, mine.f, mine.row)) mine = Best{(double)my_best, my_row};
    }
  }
  if (a.propose_only) return;
  Best top;
  if (grid_best(mine, a.scratch, a.ctrl, &top))
    finalize_generation<T>(top, a.Xnew, a.ld, a.N, a.gbest, a.ctrl, a.it, a.maxiter, a.xtol, a.ftol);
}

template <typename T, int CH, int STRAT, int FETCH, bool FULL>
static cudaError_t de_rows_launch(const DeArgs<T>& a, cudaStream_t s) {
  auto kern = de_rows_kernel<T, CH, STRAT, kTmaStages, kTmaWarps, FETCH, FULL>;
  const size_t smem = FETCH == kFetchTma ? de_tma_smem(Strat<STRAT>::K, a.ld, sizeof(T)) : 0;
  static thread_local size_t configured[64];
  static thread_local int per_sm_cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (per_sm_cache[dev] == 0 || configured[dev] != smem) {
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTmaWarps * 32, smem);
    if (e != cudaSuccess) return e;
    per_sm_cache[dev] = per_sm < 1 ? 1 : per_sm;
    configured[dev] = smem;
  }
  int64_t cap = (int64_t)sm_count() * per_sm_cache[dev];
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  const int64_t need = (a.P + kTmaWarps - 1) / kTmaWarps;  // at least one row per warp
  kern<<<(int)(need < cap ? need : cap), kTmaWarps * 32, smem, s>>>(a);
  return cudaSuccess;
}

// Measured on B200 (headline shape, fp32 N=128 best1bin): TMA ring 37.9 us per generation,
// register-pipelined loads 49.0 us (spills at 64 registers) -> the ring is the default and
// the LDG policy is kept only as a profiling switch (SP_DE_FETCH=ldg, CH == 1 builds).
template <typename T, int CH, int STRAT>
static cudaError_t de_rows_pick(const DeArgs<T>& a, cudaStream_t s) {
  static const bool ldg = [] {
    const char* e = getenv("SP_DE_FETCH");
    return e != nullptr && e[0] == 'l';
  }();
  const bool full = a.N == Tile<T, CH, 32>::COLS && a.ld == a.N;
  if (CH == 1 && ldg)
    return full ? de_rows_launch<T, (CH == 1 ? CH : 1), STRAT, kFetchLdg, true>(a, s)
                : de_rows_launch<T, (CH == 1 ? CH : 1), STRAT, kFetchLdg, false>(a, s);
  return full ? de_rows_launch<T, CH, STRAT, kFetchTma, true>(a, s) : de_rows_launch<T, CH, STRAT, kFetchTma, false>(a, s);
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

}  // namespace sp
