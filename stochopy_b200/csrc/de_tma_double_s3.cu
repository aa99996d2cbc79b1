// DE TMA kernel instantiations: dtype double, strategy 3 (one unit per pair so the build parallelises)
#include <cstdlib>

#include "de_tma.cuh"
namespace sp {
cudaError_t de_tma_double_s3(const DeArgs<double>& a, int ch, cudaStream_t s) { return de_tma_by_ch<double, 3>(a, ch, s); }
}  // namespace sp
