// Strategy dispatch of the TMA-staged DE kernels (instantiated in de_tma_<dtype>_s<k>.cu).
#include "de_tma.cuh"
namespace sp {
bool de_tma_fits(int ch, int64_t P, int K, int64_t ld, size_t elem) { return de_pool_fits(ch, P, K, ld, elem); }
#define SP_DECL(T)                                                              \
  cudaError_t de_tma_##T##_s0(const DeArgs<T>&, int, cudaStream_t);             \
  cudaError_t de_tma_##T##_s1(const DeArgs<T>&, int, cudaStream_t);             \
  cudaError_t de_tma_##T##_s2(const DeArgs<T>&, int, cudaStream_t);             \
  cudaError_t de_tma_##T##_s3(const DeArgs<T>&, int, cudaStream_t);
SP_DECL(float)
SP_DECL(double)
#undef SP_DECL
cudaError_t de_tma_dispatch(const DeArgs<float>& a, int ch, cudaStream_t s) {
  switch (a.strategy) {
    case SP_DE_RAND1BIN: return de_tma_float_s0(a, ch, s);
    case SP_DE_RAND2BIN: return de_tma_float_s1(a, ch, s);
    case SP_DE_BEST1BIN: return de_tma_float_s2(a, ch, s);
    default: return de_tma_float_s3(a, ch, s);
  }
}
cudaError_t de_tma_dispatch(const DeArgs<double>& a, int ch, cudaStream_t s) {
  switch (a.strategy) {
    case SP_DE_RAND1BIN: return de_tma_double_s0(a, ch, s);
    case SP_DE_RAND2BIN: return de_tma_double_s1(a, ch, s);
    case SP_DE_BEST1BIN: return de_tma_double_s2(a, ch, s);
    default: return de_tma_double_s3(a, ch, s);
  }
}
}  // namespace sp
