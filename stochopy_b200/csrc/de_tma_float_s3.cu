// DE TMA kernel instantiations: dtype float, strategy 3 (one unit per pair so the build parallelises)
#include <cstdlib>

#include "de_tma.cuh"
namespace sp {
cudaError_t de_tma_float_s3(const DeArgs<float>& a, int ch, cudaStream_t s) { return de_tma_by_ch<float, 3>(a, ch, s); }
}  // namespace sp
