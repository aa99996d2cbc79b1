// The seven benchmark objectives evaluated on a register-resident row tile.
// Reference formulas: stochopy/factory/benchmark.py (line numbers per case).
#pragma once
#include "common.cuh"

namespace sp {

// fp32 cos(2 pi x): x - rint(x) is exact, so the SFU cosine sees an argument in [-pi, pi]
// (abs error ~2^-21, one multiply + one MUFU instead of the ~18-instruction cospif)
__device__ __forceinline__ float cos2pi(float x) { return __cosf(6.2831855f * (x - rintf(x))); }
__device__ __forceinline__ double cos2pi(double x) { return cospi(2.0 * x); }
__device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float t_exp(float x) { return expf(x); }
__device__ __forceinline__ double t_exp(double x) { return exp(x); }
__device__ __forceinline__ float t_cos(float x) { return cosf(x); }
__device__ __forceinline__ double t_cos(double x) { return cos(x); }

// x[j+1] for element (c, e) of a tile: next element of the vector, the first
// element of the next lane's vector, or (last lane of the row) the first
// element of lane 0's next chunk.
template <typename T, int CH, int LPR>
__device__ __forceinline__ void right_neighbours(const Tile<T, CH, LPR>& x, int l, T (&nb)[CH]) {
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    if (LPR == 1) {
      nb[c] = (c + 1 < CH) ? x.v[(c + 1 < CH) ? c + 1 : c][0] : T(0);
    } else {
      T same = __shfl_down_sync(0xffffffffu, x.v[c][0], 1, LPR);
      T wrap = (c + 1 < CH) ? __shfl_sync(0xffffffffu, x.v[(c + 1 < CH) ? c + 1 : c][0], 0, LPR) : T(0);
      nb[c] = (l == LPR - 1) ? wrap : same;
    }
  }
}

// f(x) for the row held by this lane group; every lane of the group gets f.
// Columns >= N (padding) are masked out.
template <typename T, int CH, int LPR>
__device__ __forceinline__ T evaluate_tile(int objective, const Tile<T, CH, LPR>& x, int l, int N) {
  constexpr int VEC = Num<T>::VEC;
  using TL = Tile<T, CH, LPR>;
  switch (objective) {
    case SP_OBJ_SPHERE: {  // benchmark.py:121-136
      T s = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (TL::col(c, l, e) < N) s += x.v[c][e] * x.v[c][e];
      return group_sum<LPR>(s);
    }
    case SP_OBJ_ROSENBROCK: {  // benchmark.py:100-118: 100*sum((x[1:]-x[:-1]^2)^2) + sum((1-x[:-1])^2)
      T nb[CH];
      right_neighbours(x, l, nb);
      T s1 = 0, s2 = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (TL::col(c, l, e) + 1 < N) {
            T xi = x.v[c][e];
            T xn = (e + 1 < VEC) ? x.v[c][(e + 1 < VEC) ? e + 1 : e] : nb[c];
            T a = xn - xi * xi, b = T(1) - xi;
            s1 += a * a;
            s2 += b * b;
          }
        }
      // one butterfly for both sums: sum_l (100 s1_l + s2_l) == 100 sum s1 + sum s2
      return group_sum<LPR>(T(100) * s1 + s2);
    }
    case SP_OBJ_RASTRIGIN: {  // benchmark.py:79-97
      T s = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (TL::col(c, l, e) < N) s += x.v[c][e] * x.v[c][e] - T(10) * cos2pi(x.v[c][e]);
      return T(10) * T(N) + group_sum<LPR>(s);
    }
    case SP_OBJ_STYBLINSKI_TANG: {  // benchmark.py:139-156
      T s = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (TL::col(c, l, e) < N) {
            T v = x.v[c][e], v2 = v * v;
            s += v2 * v2 - T(16) * v2 + T(5) * v;
          }
      return T(0.5) * group_sum<LPR>(s) + T(39.16599) * T(N);
    }
    case SP_OBJ_QUARTIC: {  // benchmark.py:59-76
      T s = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          int j = TL::col(c, l, e);
          if (j < N) {
            T v2 = x.v[c][e] * x.v[c][e];
            s += T(j + 1) * (v2 * v2);
          }
        }
      return group_sum<LPR>(s);
    }
    case SP_OBJ_ACKLEY: {  // benchmark.py:14-34 (e = 2.7182818284590451 literal)
      T s1 = 0, s2 = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (TL::col(c, l, e) < N) {
            s1 += x.v[c][e] * x.v[c][e];
            s2 += cos2pi(x.v[c][e]);
          }
      s1 = group_sum<LPR>(s1);
      s2 = group_sum<LPR>(s2);
      const T inv = T(1) / T(N);
      return T(20) + T(2.7182818284590451) - T(20) * t_exp(T(-0.2) * t_sqrt(inv * s1)) - t_exp(inv * s2);
    }
    case SP_OBJ_GRIEWANK: {  // benchmark.py:37-56
      T s = 0, p = 1;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          int j = TL::col(c, l, e);
          if (j < N) {
            s += x.v[c][e] * x.v[c][e];
            p *= t_cos(x.v[c][e] / t_sqrt(T(j + 1)));
          }
        }
      s = group_sum<LPR>(s);
      p = group_prod<LPR>(p);
      return T(1) + s / T(4000) - p;
    }
    default:
      return T(0);
  }
}

}  // namespace sp
