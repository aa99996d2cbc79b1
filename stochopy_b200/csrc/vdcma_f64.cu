// VD-CMA kernels and launchers, dtype double (see vdcma_impl.cuh)
#include "vdcma_impl.cuh"
namespace sp {
int vd_sample_f64(const sp_vd_state* st, int it, int evaluate, cudaStream_t s) { return vd_sample<double>(st, it, evaluate, s); }
int vd_update_f64(const sp_vd_state* st, int it, cudaStream_t s) { return vd_update<double>(st, it, s); }
int vd_refresh_f64(const sp_vd_state* st, cudaStream_t s) { return vd_refresh_t<double>(st, s); }
int vd_clocks_f64(long long* out16) { return vd_clocks_t(out16); }
}  // namespace sp
