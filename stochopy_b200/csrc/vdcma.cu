// VD-CMA: C entry points (the kernels and launchers live in vdcma_impl.cuh, compiled once per dtype in
// vdcma_f32.cu / vdcma_f64.cu).  Reference: stochopy/optimize/vdcma/_vdcma.py:235-458.
#include "es_common.cuh"
#include "rank.cuh"

namespace sp {

#define SP_VD_DECL(DT)                                                              \
  int vd_sample_##DT(const sp_vd_state* st, int it, int evaluate, cudaStream_t s);  \
  int vd_update_##DT(const sp_vd_state* st, int it, cudaStream_t s);                \
  int vd_refresh_##DT(const sp_vd_state* st, cudaStream_t s);                       \
  int vd_clocks_##DT(long long* out16);
SP_VD_DECL(f32)
SP_VD_DECL(f64)
#undef SP_VD_DECL

static int vd_check(const sp_vd_state* st, int it) {
  SP_CHECK_ARG(st != nullptr, "null state");
  SP_CHECK_ARG(st->dtype == SP_F32 || st->dtype == SP_F64, "dtype");
  SP_CHECK_ARG(st->N >= 1 && st->P >= 2 && st->P < (1LL << 31) && st->mu >= 1 && st->mu <= st->P, "popsize / mu / ndim");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || st->constraint == SP_CONS_PENALIZE, "constraint");
  const int vec = st->dtype == SP_F32 ? 4 : 2;
  SP_CHECK_ARG(st->ld >= st->N && st->ld % vec == 0, "ld must be a multiple of 16/sizeof(T)");
  SP_CHECK_ARG(st->xmean && st->xold && st->dx && st->pc && st->dvec && st->vvec && st->vn && st->diagC && st->dy &&
                   st->ginj && st->arx && st->ary && st->yvn && st->arfit && st->weights && st->xscale && st->xshift &&
                   st->besthist && st->work && st->rank && st->ctrl,
               "null buffer");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || (st->bnd_weights && st->dfithist && st->hist_cap >= 2), "Penalize buffers");
  SP_CHECK_ARG(it >= 1 && it <= st->maxiter, "generation index in [1, maxiter]");
  return SP_OK;
}

}  // namespace sp

using namespace sp;

extern "C" {

int64_t sp_vd_work_scalars(int N, int64_t P) {
  return (int64_t)kVdChunks * 4 * N + 15LL * N + 8 + P + kVdChunks + 2 + 2 * 3 * kVdUpMax + 4 * (P + 2 * kRankChunk);
}

int sp_vd_refresh(const sp_vd_state* st, void* stream) {
  int rc = vd_check(st, 1);
  if (rc) return rc;
  return st->dtype == SP_F32 ? vd_refresh_f32(st, (cudaStream_t)stream) : vd_refresh_f64(st, (cudaStream_t)stream);
}

int sp_vd_sample(const sp_vd_state* st, int it, int evaluate, void* stream) {
  int rc = vd_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(!evaluate || (st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG),
               "device objective required to evaluate in the sampling kernel");
  return st->dtype == SP_F32 ? vd_sample_f32(st, it, evaluate, (cudaStream_t)stream)
                             : vd_sample_f64(st, it, evaluate, (cudaStream_t)stream);
}

int sp_vd_update(const sp_vd_state* st, int it, void* stream) {
  int rc = vd_check(st, it);
  if (rc) return rc;
  return st->dtype == SP_F32 ? vd_update_f32(st, it, (cudaStream_t)stream) : vd_update_f64(st, it, (cudaStream_t)stream);
}

int sp_vd_generation(const sp_vd_state* st, int it, void* stream) {
  int rc = sp_vd_sample(st, it, 1, stream);
  if (rc) return rc;
  return sp_vd_update(st, it, stream);
}

int sp_vd_run(const sp_vd_state* st, int it_first, int n, void* stream) {
  SP_CHECK_ARG(st != nullptr && !st->host_z, "sp_vd_run needs in-kernel draws");
  for (int g = 0; g < n; ++g) {
    int rc = sp_vd_generation(st, it_first + g, stream);
    if (rc) return rc;
  }
  return SP_OK;
}

}  // extern "C"

// profiling hook (not part of the public header): the update kernel's stage clocks of the last generation of the
// given dtype (SP_F32 / SP_F64: one copy of the stamp array per translation unit)
extern "C" int sp_debug_vd_clocks(long long* out16, int dtype) {
  return dtype == SP_F32 ? sp::vd_clocks_f32(out16) : sp::vd_clocks_f64(out16);
}
