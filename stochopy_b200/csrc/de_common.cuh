// Shared pieces of the DE kernels (direct-load and TMA-staged variants).
#pragma once
#include "objectives.cuh"
#include "philox.cuh"

namespace sp {

template <typename T>
struct DeArgs {
  int objective, strategy, constraint, it, maxiter, N, propose_only;
  int64_t P, ld;
  T F, CR;
  double xtol, ftol;
  uint64_t seed;
  const T* Xold;
  T* Xnew;
  T* pbestfit;
  T* pfit;
  T* gbest;
  const T* lower;
  const T* upper;
  sp_ctrl* ctrl;
  Best* scratch;
  const T* r1;
  const int64_t* donors;
  const int64_t* irand;
  const T* repair;
  int chain;  // SP_CHAIN_IN / SP_CHAIN_OUT (pool kernel only)
};

__host__ __device__ constexpr int de_donor_count(int strategy) {
  return strategy == SP_DE_RAND1BIN ? 3 : strategy == SP_DE_RAND2BIN ? 5 : strategy == SP_DE_BEST1BIN ? 2 : 4;
}

// k distinct donors != row: draw from the shrinking range and step over the
// excluded indices in ascending order (uniform over what is left; equals in
// distribution the first k entries of the reference's permutation, _de.py:306).
__device__ __forceinline__ void draw_donors(uint32_t row, uint32_t P, int k, int it, uint64_t seed, int N,
                                            uint32_t (&d)[5], int* irand) {
  uint4 a = philox4x32(0u, row, (uint32_t)it, kDeIndex, seed);
  uint32_t words[5] = {a.y, a.z, a.w, 0u, 0u};
  if (k > 3) {
    uint4 b = philox4x32(1u, row, (uint32_t)it, kDeIndex, seed);
    words[3] = b.x;
    words[4] = b.y;
  }
  *irand = (int)bounded(a.x, (uint32_t)N);
  uint32_t excl[6] = {row, 0u, 0u, 0u, 0u, 0u};  // ascending; t + 1 entries valid at step t
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    if (t < k) {
      uint32_t r = bounded(words[t], P - 1u - (uint32_t)t);
#pragma unroll
      for (int e = 0; e <= t; ++e)
        if (r >= excl[e]) ++r;
      d[t] = r;
      excl[t + 1] = r;  // one backward bubble pass restores the order
#pragma unroll
      for (int e = t + 1; e > 0; --e)
        if (excl[e - 1] > excl[e]) {
          uint32_t tmp = excl[e - 1];
          excl[e - 1] = excl[e];
          excl[e] = tmp;
        }
    }
  }
}


// pool kernel (de_tma.cuh), one translation unit per (dtype, strategy)
bool de_tma_fits(int ch, int64_t P, int K, int64_t ld, size_t elem);
cudaError_t de_tma_dispatch(const DeArgs<float>& a, int ch, cudaStream_t s);
cudaError_t de_tma_dispatch(const DeArgs<double>& a, int ch, cudaStream_t s);

}  // namespace sp
