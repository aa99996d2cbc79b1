// DE TMA kernel instantiations: dtype float, strategy 0 (one unit per pair so the build parallelises)
#include <cstdlib>

#include "de_tma.cuh"
namespace sp {
cudaError_t de_tma_float_s0(const DeArgs<float>& a, int ch, cudaStream_t s) { return de_tma_by_ch<float, 0>(a, ch, s); }
}  // namespace sp
