// Row-wise kernels shared by several translation units: batched objective
// evaluation (a1/a2) and the counting rank used for argsort-style selections.
#pragma once
#include "objectives.cuh"
#include "philox.cuh"
#include "rank.cuh"

namespace sp {

// ---- a1/a2 --------------------------------------------------------------------
template <typename T, int CH, int LPR>
__global__ void __launch_bounds__(kThreads)
eval_kernel(int objective, const T* __restrict__ X, int64_t P, int N, int64_t ld, const T* __restrict__ scale,
            const T* __restrict__ shift, T* __restrict__ f, int clip, const int32_t* live) {
  using TL = Tile<T, CH, LPR>;
  // `live` (optional): the control block's status -- generations enqueued behind a stop do nothing
  if (live != nullptr && *reinterpret_cast<const volatile int32_t*>(live) != SP_RUNNING) return;
  constexpr int VEC = Num<T>::VEC;
  const int lane = threadIdx.x & 31, l = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t groups = (P + TL::RPW - 1) / TL::RPW;
  for (int64_t g = warp; g < groups; g += nwarps) {
    int64_t row = g * TL::RPW + sub;
    const bool live = row < P;
    if (!live) row = P - 1;
    TL x;
    x.load(X + row * ld, l, (int)ld);
    if (clip) {  // Penalize evaluates the population clipped to [-1, 1] (cmaes/_constraints.py:30-32)
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) x.v[c][e] = x.v[c][e] < T(-1) ? T(-1) : (x.v[c][e] > T(1) ? T(1) : x.v[c][e]);
    }
    if (scale != nullptr) {  // un-standardise: x * xstd + xm (_cmaes.py:168-173)
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          int j = TL::col(c, l, e);
          if (j < N) x.v[c][e] = add_rn(mul_rn(x.v[c][e], scale[j]), shift[j]);
        }
    }
    T val = evaluate_tile<T, CH, LPR>(objective, x, l, N);
    if (live && l == 0) f[row] = val;
  }
}


template <typename T>
inline int eval_launch(int objective, const void* X, int64_t P, int N, int64_t ld, const void* scale,
                       const void* shift, void* f, int clip, cudaStream_t s, const int32_t* live = nullptr) {
  Shape sh;
  if (!pick_shape(N, Num<T>::VEC, &sh)) {
    set_error("sp_eval: ndim %d exceeds the compiled row shapes", N);
    return SP_ERR_SHAPE;
  }
  const int grid = grid_for_rows(P, sh.lpr, 8);
#define SP_CALL(TT, C, L)                                                                                     \
  eval_kernel<TT, C, L><<<grid, kThreads, 0, s>>>(objective, (const TT*)X, P, N, ld, (const TT*)scale,        \
                                                  (const TT*)shift, (TT*)f, clip, live)
  SP_DISPATCH_SHAPE(T, sh, SP_CALL);
#undef SP_CALL
  SP_CHECK_LAUNCH();
  return SP_OK;
}

// out[row][j] = Philox uniform [0,1) or N(0,1) draw (block = j / VEC, row, it, purpose)
template <typename T>
__global__ void random_fill_kernel(T* __restrict__ out, int64_t P, int N, int64_t ld, int it, uint64_t seed,
                                   uint32_t purpose, int normal) {
  constexpr int VEC = Num<T>::VEC;
  const int nb = (N + VEC - 1) / VEC;
  const int64_t total = P * (int64_t)nb;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / nb;
    const int b = (int)(t - row * nb);
    T z[VEC];
    const uint4 o = philox4x32_for(purpose, (uint32_t)b, (uint32_t)row, (uint32_t)it, seed);
    if (normal) normal_block(o, z);
    else uniform_block(o, z);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (b * VEC + e < N) out[row * ld + b * VEC + e] = z[e];
  }
}

}  // namespace sp
