// Differential evolution: one fused kernel per synchronous generation.
// Reference: stochopy/optimize/de/_de.py:314-351 (de_sync), de/_strategy.py,
// de/_constraints.py, and selection_sync (_common.py:123-160).
//
// Per individual i (one lane group per row, rows double-buffered):
//   donors (k distinct rows != i)  -> mutant V
//   binomial crossover with own row (forced column irand)      -> trial U
//   Random repair of out-of-bounds coordinates
//   f(U), strict-< selection against pbestfit[i]
//   write row of the next population, track argmin(pbestfit)
// The last CTA refreshes gbest and evaluates the termination ladder.
// Algorithmic HBM bytes per individual: (k + 2) * N * s + 3 * s.
#include <cstdlib>

#include "de_common.cuh"

namespace sp {

// test hook: SP_DE_DIRECT=1 keeps the direct-load kernel for every shape
static const bool g_force_direct = getenv("SP_DE_DIRECT") != nullptr;

template <typename T, int CH, int LPR, bool PHILOX>
__global__ void __launch_bounds__(kThreads)
de_generation_kernel(const DeArgs<T> a, const CrossKeys keys) {
  using TL = Tile<T, CH, LPR>;
  constexpr int VEC = Num<T>::VEC;
  if (!running(a.ctrl)) return;
  const int lane = threadIdx.x & 31, l = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t groups = (a.P + TL::RPW - 1) / TL::RPW;
  const int k = de_donor_count(a.strategy);
  const bool from_best = a.strategy >= SP_DE_BEST1BIN;
  const int ld = (int)a.ld;

  TL gb;  // gbest stays in registers for the whole CTA lifetime
  if (from_best) gb.load(a.gbest, l, ld);

  Best mine{1.0 / 0.0, 0x7fffffffffffffffLL};
  for (int64_t g = warp; g < groups; g += nwarps) {
    int64_t row = g * TL::RPW + sub;
    const bool live = row < a.P;
    if (!live) row = a.P - 1;

    uint32_t d[5] = {0, 0, 0, 0, 0};
    int irand;
    if (PHILOX) {
      draw_donors((uint32_t)row, (uint32_t)a.P, k, a.it, a.seed, a.N, d, &irand);
    } else {
      irand = (int)a.irand[row];
#pragma unroll
      for (int t = 0; t < 5; ++t)
        if (t < k) d[t] = (uint32_t)a.donors[(int64_t)t * a.P + row];
    }

    TL xi, u;
    xi.load(a.Xold + row * a.ld, l, ld);
    {  // mutant, de/_strategy.py:1-38 (numpy's operation order, no FMA contraction)
      TL d0, d1;
      d0.load(a.Xold + (int64_t)d[0] * a.ld, l, ld);
      d1.load(a.Xold + (int64_t)d[1] * a.ld, l, ld);
      if (a.strategy == SP_DE_BEST1BIN) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int e = 0; e < VEC; ++e) u.v[c][e] = add_rn(gb.v[c][e], mul_rn(a.F, sub_rn(d0.v[c][e], d1.v[c][e])));
      } else {
        TL d2;
        d2.load(a.Xold + (int64_t)d[2] * a.ld, l, ld);
        if (a.strategy == SP_DE_RAND1BIN) {
#pragma unroll
          for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int e = 0; e < VEC; ++e) u.v[c][e] = add_rn(d0.v[c][e], mul_rn(a.F, sub_rn(d1.v[c][e], d2.v[c][e])));
        } else {
          TL d3;
          d3.load(a.Xold + (int64_t)d[3] * a.ld, l, ld);
          if (a.strategy == SP_DE_BEST2BIN) {
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
              for (int e = 0; e < VEC; ++e)
                u.v[c][e] = add_rn(
                    gb.v[c][e],
                    mul_rn(a.F, sub_rn(sub_rn(add_rn(d0.v[c][e], d1.v[c][e]), d2.v[c][e]), d3.v[c][e])));
          } else {  // rand2bin
            TL d4;
            d4.load(a.Xold + (int64_t)d[4] * a.ld, l, ld);
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
              for (int e = 0; e < VEC; ++e)
                u.v[c][e] = add_rn(
                    d0.v[c][e],
                    mul_rn(a.F, sub_rn(sub_rn(add_rn(d1.v[c][e], d2.v[c][e]), d3.v[c][e]), d4.v[c][e])));
          }
        }
      }
    }

    // binomial crossover (_de.py:339-344) and Random repair (de/_constraints.py:22-26)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int j0 = TL::col(c, l, 0);
      if (j0 < a.N) {
        bool tk[VEC];
        if (PHILOX) {  // 16-bit pieces of one Philox2x32-10 call per 4 columns (philox.cuh: DE crossover stream)
          bool t4[4];
          de_cross_take((uint32_t)row, (uint32_t)(j0 >> 2), keys, t4);
#pragma unroll
          for (int e = 0; e < VEC; ++e) tk[e] = VEC == 4 ? t4[e] : ((j0 & 2) ? t4[2 + (e & 1)] : t4[e & 1]);
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) tk[e] = (j0 + e < a.N) ? (a.r1[row * a.ld + j0 + e] <= a.CR) : false;
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int j = j0 + e;
          const bool take = (j == irand) || tk[e];
          u.v[c][e] = (take && j < a.N) ? u.v[c][e] : xi.v[c][e];
        }
        if (a.constraint == SP_CONS_RANDOM) {
          bool any = false;
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (j0 + e < a.N) any |= (u.v[c][e] < a.lower[j0 + e]) || (u.v[c][e] > a.upper[j0 + e]);
          if (any) {
            T q[VEC];
            if (PHILOX) uniform_block(philox4x32((uint32_t)(j0 / VEC), (uint32_t)row, (uint32_t)a.it, kDeRepair, a.seed), q);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const int j = j0 + e;
              if (j < a.N) {
                const T lo = a.lower[j], hi = a.upper[j];
                if (u.v[c][e] < lo || u.v[c][e] > hi)
                  u.v[c][e] = PHILOX ? add_rn(lo, mul_rn(sub_rn(hi, lo), q[e])) : a.repair[row * a.ld + j];
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) u.v[c][e] = xi.v[c][e];
      }
    }

    if (a.propose_only) {  // SP_OBJ_HOST: the caller evaluates U, then sp_select_sync(copy_when=0)
      if (live) u.store(a.Xnew + row * a.ld, l, ld);
      continue;
    }

    const T f = evaluate_tile<T, CH, LPR>(a.objective, u, l, a.N);
    T best = a.pbestfit[row];
    const bool win = f < best;  // strict, _common.py:127
    if (win) best = f;
    if (live) {
      if (win) u.store(a.Xnew + row * a.ld, l, ld);
      else xi.store(a.Xnew + row * a.ld, l, ld);
      if (l == 0) {
        a.pbestfit[row] = best;
        a.pfit[row] = f;
        if (better((double)best, row, mine.f, mine.row)) mine = Best{(double)best, row};
      }
    }
  }
  if (a.propose_only) return;
  Best top;
  if (grid_best(mine, a.scratch, a.ctrl, &top))
    finalize_generation<T>(top, a.Xnew, a.ld, a.N, a.gbest, a.ctrl, a.it, a.maxiter, a.xtol, a.ftol);
}


template <typename T>
static int de_launch(const sp_de_state* st, int it, int propose_only, int chain, cudaStream_t s) {
  Shape sh;
  if (!pick_shape(st->N, Num<T>::VEC, &sh)) {
    set_error("sp_de_generation: ndim %d exceeds the compiled row shapes", st->N);
    return SP_ERR_SHAPE;
  }
  DeArgs<T> a;
  a.objective = st->objective;
  a.strategy = st->strategy;
  a.constraint = st->constraint;
  a.it = it;
  a.maxiter = st->maxiter;
  a.N = st->N;
  a.propose_only = propose_only;
  a.P = st->P;
  a.ld = st->ld;
  a.F = (T)st->F;
  a.CR = (T)st->CR;
  a.xtol = st->xtol;
  a.ftol = st->ftol;
  a.seed = st->seed;
  a.Xold = (const T*)st->X[it & 1];
  a.Xnew = (T*)st->X[(it & 1) ^ 1];
  a.pbestfit = (T*)st->pbestfit;
  a.pfit = (T*)st->pfit;
  a.gbest = (T*)st->gbest;
  a.lower = (const T*)st->lower;
  a.upper = (const T*)st->upper;
  a.ctrl = st->ctrl;
  a.scratch = (Best*)st->scratch;
  a.r1 = (const T*)st->r1;
  a.donors = st->donors;
  a.irand = st->irand;
  a.repair = (const T*)st->repair;
  a.chain = 0;
  const bool philox = st->r1 == nullptr;
  const int k = de_donor_count(st->strategy);
  const bool pool = philox && sh.lpr == 32 && !g_force_direct && de_tma_fits(sh.ch, st->P, k, st->ld, sizeof(T));
  if (!pool && chain != 0) {  // the direct-load kernel always resolves its own generation
    set_error("sp_de_generation_chained: this problem shape does not run on the pool kernel (sp_de_chainable)");
    return SP_ERR_ARG;
  }
  if (pool) {
    a.chain = propose_only ? 0 : chain;
    cudaError_t e = de_tma_dispatch(a, sh.ch, s);
    if (e != cudaSuccess) {
      set_error("sp_de_generation: %s", cudaGetErrorString(e));
      return SP_ERR_CUDA;
    }
    SP_CHECK_LAUNCH();
    return SP_OK;
  }
  const int grid = grid_for_rows(st->P, sh.lpr, sh.ch >= 4 ? 2 : 4);
  const CrossKeys ck = de_cross_keys(a.seed, a.it, (double)a.CR);
#define SP_CALL(TT, C, L)                                                        \
  do {                                                                           \
    if (philox) de_generation_kernel<TT, C, L, true><<<grid, kThreads, 0, s>>>(a, ck); \
    else de_generation_kernel<TT, C, L, false><<<grid, kThreads, 0, s>>>(a, ck);     \
  } while (0)
  SP_DISPATCH_SHAPE(T, sh, SP_CALL);
#undef SP_CALL
  SP_CHECK_LAUNCH();
  return SP_OK;
}

static int de_check(const sp_de_state* st, int it) {
  SP_CHECK_ARG(st != nullptr, "null state");
  SP_CHECK_ARG(st->dtype == SP_F32 || st->dtype == SP_F64, "dtype");
  SP_CHECK_ARG(st->strategy >= SP_DE_RAND1BIN && st->strategy <= SP_DE_BEST2BIN, "strategy");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || st->constraint == SP_CONS_RANDOM, "constraint");
  SP_CHECK_ARG(st->N >= 1 && st->P > de_donor_count(st->strategy) && st->P < (1LL << 31), "popsize / ndim");
  const int vec = st->dtype == SP_F32 ? 4 : 2;
  SP_CHECK_ARG(st->ld >= st->N && st->ld % vec == 0, "ld must be a multiple of 16/sizeof(T)");
  SP_CHECK_ARG(st->X[0] && st->X[1] && st->pbestfit && st->pfit && st->gbest && st->ctrl && st->scratch, "null buffer");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || (st->lower && st->upper), "bounds needed for Random");
  const bool any = st->r1 || st->donors || st->irand;
  const bool all = st->r1 && st->donors && st->irand && (st->constraint == SP_CONS_NONE || st->repair);
  SP_CHECK_ARG(!any || all, "explicit draws must be given together");
  SP_CHECK_ARG(it >= 2, "generation index starts at 2 (_de.py:246-248)");
  return SP_OK;
}

}  // namespace sp

using namespace sp;

extern "C" {

int sp_de_generation(const sp_de_state* st, int it, void* stream) {
  int rc = de_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG,
               "device objective required (use sp_de_propose + sp_select_sync for host objectives)");
  return st->dtype == SP_F32 ? de_launch<float>(st, it, 0, 0, (cudaStream_t)stream)
                             : de_launch<double>(st, it, 0, 0, (cudaStream_t)stream);
}

int sp_de_chainable(const sp_de_state* st) {
  if (st == nullptr || st->r1 != nullptr || g_force_direct) return 0;
  if (st->objective < SP_OBJ_ACKLEY || st->objective > SP_OBJ_STYBLINSKI_TANG) return 0;
  const int vec = st->dtype == SP_F32 ? 4 : 2;
  Shape sh;
  if (!pick_shape(st->N, vec, &sh) || sh.lpr != 32) return 0;
  return de_tma_fits(sh.ch, st->P, de_donor_count(st->strategy), st->ld, st->dtype == SP_F32 ? 4 : 8) ? 1 : 0;
}

int sp_de_generation_chained(const sp_de_state* st, int it, int flags, void* stream) {
  int rc = de_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG((flags & ~(SP_CHAIN_IN | SP_CHAIN_OUT)) == 0, "flags");
  SP_CHECK_ARG(st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG, "device objective required");
  SP_CHECK_ARG(flags == 0 || sp_de_chainable(st), "state is not chainable (sp_de_chainable)");
  return st->dtype == SP_F32 ? de_launch<float>(st, it, 0, flags, (cudaStream_t)stream)
                             : de_launch<double>(st, it, 0, flags, (cudaStream_t)stream);
}

int sp_de_propose(const sp_de_state* st, int it, void* stream) {
  int rc = de_check(st, it);
  if (rc) return rc;
  return st->dtype == SP_F32 ? de_launch<float>(st, it, 1, 0, (cudaStream_t)stream)
                             : de_launch<double>(st, it, 1, 0, (cudaStream_t)stream);
}

int sp_de_run(const sp_de_state* st, int it_first, int n, void* stream) {
  SP_CHECK_ARG(st != nullptr && st->r1 == nullptr, "sp_de_run needs in-kernel draws");
  // inside the chunk the generations are chained: only the last one runs the last-CTA epilogue
  const bool chain = sp_de_chainable(st) != 0;
  for (int g = 0; g < n; ++g) {
    const int flags = chain ? ((g > 0 ? SP_CHAIN_IN : 0) | (g < n - 1 ? SP_CHAIN_OUT : 0)) : 0;
    int rc = sp_de_generation_chained(st, it_first + g, flags, stream);
    if (rc) return rc;
  }
  return SP_OK;
}

}  // extern "C"
