// VD-CMA kernels and launchers, dtype float (see vdcma_impl.cuh)
#include "vdcma_impl.cuh"
namespace sp {
int vd_sample_f32(const sp_vd_state* st, int it, int evaluate, cudaStream_t s) { return vd_sample<float>(st, it, evaluate, s); }
int vd_update_f32(const sp_vd_state* st, int it, cudaStream_t s) { return vd_update<float>(st, it, s); }
int vd_refresh_f32(const sp_vd_state* st, cudaStream_t s) { return vd_refresh_t<float>(st, s); }
int vd_clocks_f32(long long* out16) { return vd_clocks_t(out16); }
}  // namespace sp
