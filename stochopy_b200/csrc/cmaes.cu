// (mu,lambda)-CMA-ES: one generation = a short chain of kernels on one stream,
// everything (sigma, paths, covariance, eigenbasis, termination) device resident.
// Reference: stochopy/optimize/cmaes/_cmaes.py:228-343 (generation), :360-434
// (converge), cmaes/_constraints.py:4-82 (Penalize).
//
//   normal_fill      Z ~ N(0, I) (Philox + Box-Muller)              [skipped with host_z]
//   cma_sample       arx = xmean + sigma * Z (B diag(D))^T          GEMM NT  2 P N^2 flop
//   eval             arfit = fun(clip?(arx) * xscale + xshift)
//   [Penalize]       rank -> percentiles -> weights -> arfit += penalty
//   rank             np.argsort(arfit) as a rank per individual
//   cma_gather       the mu best rows in rank order -> contiguous panel (reuses Z), best row
//   cma_mean         xmean' = sum_r w[r] panel_r                    (deterministic two-stage)
//   cma_paths        ps, hsig, pc, sigma', eigen-update decision    one CTA
//   cma_cov          C' = (1-c1-cmu) C + cmu A^T diag(w) A + ...    GEMM TN  2 mu N^2 flop
//   [jacobi + post]  C = B diag(D^2) B^T, invsqrtC                  when due
//   cma_converge     termination ladder                             one CTA
#include "es_common.cuh"
#include "linalg.cuh"

namespace sp {

constexpr int kMeanChunks = 64;
constexpr int kCovSplits = 8;

template <typename T>
struct CmaPtrs {
  T *xmean, *xold, *pc, *ps, *C, *B, *D, *invsqrtC, *arx, *arfit, *Z, *weights, *xscale, *xshift, *besthist,
      *work, *bnd_weights, *dfithist;
  int32_t* rank;
  sp_es_ctrl* ctrl;
  int N, mu, maxiter, ilim, hist_cap, constraint, it;
  int64_t P, ld;
  double cc, cs, c1, cmu, damps, chind, mueff, xtol, ftol, insigma;
  // workspace slices
  __host__ __device__ T* mean_part() const { return work; }                                  // kMeanChunks * N
  __host__ __device__ T* cov_part() const { return work + (size_t)kMeanChunks * N; }          // kCovSplits * N * N
  __host__ __device__ T* jac() const { return cov_part() + (size_t)kCovSplits * N * N; }      // 2 N^2 + N + 64
  __host__ __device__ T* vec() const { return jac() + 2 * (size_t)N * N + N + 64; }           // 8 * N (diff, coef, ...)
  __host__ __device__ T* inv_d() const { return vec() + 4 * (size_t)N; }                      // N: 1 / D of the last decomposition
  __host__ __device__ T* sorted() const { return vec() + 8 * (size_t)N; }                     // P (Penalize percentiles)
};

// ---- arx = xmean + sigma * (Z diag(D)) B^T  (_cmaes.py:232-237) --------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
cma_sample_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const T sigma = (T)a.ctrl->sigma;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.ctrl->sigma_gen = a.ctrl->sigma;
  const T* Z = a.Z;
  const T* Bm = a.B;
  const T* D = a.D;
  const int64_t ld = a.ld;
  const int N = a.N;
  // y_i = B (D o z_i): the scaling rides on the A operand so products associate like the reference's
  gemm_nt_tile<T>((int)a.P, N, N, blockIdx.y * kGemmTile, blockIdx.x * kGemmTile,
                  [=] __device__(int m, int k) { return mul_rn(D[k], Z[(int64_t)m * ld + k]); },
                  [=] __device__(int n, int k) { return Bm[(size_t)n * N + k]; },
                  [=] __device__(int m, int n, T acc) { a.arx[(int64_t)m * ld + n] = add_rn(a.xmean[n], mul_rn(sigma, acc)); });
}

// ---- the mu best rows, in rank order, as a contiguous panel (_cmaes.py:272-274) ------------------
// panel[r][:] = arx[i][:] for r = rank[i] < mu.  Z is dead once the sampling GEMM has run, so it holds
// the panel; the mean and the rank-mu GEMM then stream mu contiguous rows instead of testing the rank
// of all P.  The warp that meets rank 0 records the best row of the generation.
template <typename T>
__global__ void __launch_bounds__(256)
cma_gather_kernel(const CmaPtrs<T> a) {
  using V = typename Num<T>::vec_t;
  if (!es_running(a.ctrl)) return;
  const int lane = threadIdx.x & 31, nv = (int)(a.ld / Num<T>::VEC);
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = gw; i < a.P; i += nw) {
    const int r = a.rank[i];
    if (r == 0 && lane == 0) a.ctrl->base.gbest_row = i;
    if (r >= a.mu) continue;
    const V* __restrict__ src = reinterpret_cast<const V*>(a.arx + i * a.ld);
    V* __restrict__ dst = reinterpret_cast<V*>(a.Z + (int64_t)r * a.ld);
    for (int k = lane; k < nv; k += 32) dst[k] = src[k];
  }
}

// ---- weighted mean of the mu best, stage 1 (_cmaes.py:274): chunk c sums its panel rows in rank order
template <typename T>
__global__ void __launch_bounds__(256)
cma_mean_partial_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const int per = (a.mu + kMeanChunks - 1) / kMeanChunks;
  const int r0 = blockIdx.x * per, r1 = (r0 + per < a.mu) ? r0 + per : a.mu;
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
    T acc = 0;
#pragma unroll 8
    for (int r = r0; r < r1; ++r) acc += a.weights[r] * a.Z[(int64_t)r * a.ld + n];
    a.mean_part()[(size_t)blockIdx.x * a.N + n] = acc;
  }
}

// ---- paths, step size, eigen decision: one CTA (_cmaes.py:272-301) ---------------------------
template <typename T>
__global__ void __launch_bounds__(256)
cma_paths_kernel(const CmaPtrs<T> a) {
  __shared__ double s_red[kRedDoubles];
  sp_es_ctrl* c = a.ctrl;
  if (!es_running(c)) return;
  const int N = a.N, tid = threadIdx.x;
  T* diff = a.vec();
  const T sigma = (T)c->sigma_gen;
  // xold = xmean; xmean = sum of the chunk partials (fixed order)
  for (int n = tid; n < N; n += blockDim.x) {
    T acc = 0;
    for (int g = 0; g < kMeanChunks; ++g) acc += a.mean_part()[(size_t)g * N + n];
    const T old = a.xmean[n];
    a.xold[n] = old;
    a.xmean[n] = acc;
    diff[n] = sub_rn(acc, old);
  }
  __syncthreads();
  // ps = (1-cs) ps + sqrt(cs (2-cs) mueff) invsqrtC (xmean - xold) / sigma
  const T kps = (T)sqrt(a.cs * (2.0 - a.cs) * a.mueff);
  double sq = 0.0;
  for (int n = tid; n < N; n += blockDim.x) {
    // invsqrtC = B diag(1/D) B^T is symmetric: column n read with unit stride across the threads
    T dot = 0;
    const T* col = a.invsqrtC + n;
#pragma unroll 8
    for (int k = 0; k < N; ++k) dot += col[(size_t)k * N] * diff[k];
    const T v = add_rn(mul_rn((T)(1.0 - a.cs), a.ps[n]), div_rn(mul_rn(kps, dot), sigma));
    a.ps[n] = v;
    sq += (double)v * (double)v;
  }
  const double psn = sqrt(block_sum(sq, s_red));
  const int64_t nfev = c->nfev + a.P;
  const bool hsig = psn / sqrt(1.0 - pow(1.0 - a.cs, 2.0 * (double)nfev / (double)a.P)) / a.chind < 1.4 + 2.0 / (N + 1.0);
  const T kpc = (T)sqrt(a.cc * (2.0 - a.cc) * a.mueff);
  for (int n = tid; n < N; n += blockDim.x) {
    T v = mul_rn(a.pc[n], (T)(1.0 - a.cc));
    if (hsig) v = add_rn(v, div_rn(mul_rn(kpc, diff[n]), sigma));
    a.pc[n] = v;
  }
  if (tid == 0) {
    const double best = (double)a.arfit[c->base.gbest_row];  // row of rank 0 (cma_gather_kernel)
    c->base.gfit = best;
    a.besthist[a.it - 1] = (T)best;
    c->ps_norm = psn;
    c->hsig = hsig ? 1 : 0;
    c->nfev = nfev;
    c->sigma = (double)mul_rn((T)c->sigma_gen, (T)exp((a.cs / a.damps) * (psn / a.chind - 1.0)));
    const bool due = (double)(nfev - c->eigeneval) > (double)a.P / (a.c1 + a.cmu) / (double)N / 10.0;
    c->do_eig = due ? 1 : 0;
    if (due) c->eigeneval = nfev;
  }
}

// ---- rank-mu partial products (_cmaes.py:290-293) over the panel of the mu best rows ------------
// artmp_r = (x_r - xold) / sigma (as a multiplication by 1 / sigma: <= 1 ulp from the reference's quotient)
template <typename T>
__global__ void __launch_bounds__(256)
cma_cov_partial_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const T inv_sigma = div_rn(T(1), (T)a.ctrl->sigma_gen);
  const int N = a.N;
  const int per = (a.mu + kCovSplits - 1) / kCovSplits;
  const int i0 = blockIdx.z * per, i1 = (i0 + per < a.mu) ? i0 + per : a.mu;
  T* out = a.cov_part() + (size_t)blockIdx.z * N * N;
  const int64_t ld = a.ld;
  const T* __restrict__ panel = a.Z;
  gemm_tn_tile<T>(N, N, i0, i1, blockIdx.y * kGemmTile, blockIdx.x * kGemmTile,
                  [=] __device__(int i, int r) {
                    return mul_rn(mul_rn(sub_rn(panel[(int64_t)i * ld + r], a.xold[r]), inv_sigma), a.weights[i]);
                  },
                  [=] __device__(int i, int cc) { return mul_rn(sub_rn(panel[(int64_t)i * ld + cc], a.xold[cc]), inv_sigma); },
                  [=] __device__(int r, int cc, T acc) { out[(size_t)r * N + cc] = acc; });
}

// ---- C = (1-c1-cmu) C + cmu M + c1 pc pc^T + [!hsig] c1 cc (2-cc) C_old  (_cmaes.py:291-295) -----
template <typename T>
__global__ void cma_cov_combine_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const int N = a.N;
  const bool hsig = a.ctrl->hsig != 0;
  const T keep = (T)(1.0 - a.c1 - a.cmu), cmu = (T)a.cmu, c1 = (T)a.c1, extra = (T)(a.c1 * a.cc * (2.0 - a.cc));
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N * N; e += gridDim.x * blockDim.x) {
    const int r = e / N, cc = e - r * N;
    T m = 0;
    for (int s = 0; s < kCovSplits; ++s) m += a.cov_part()[(size_t)s * N * N + e];
    const T old = a.C[e];
    T v = mul_rn(old, keep);
    v = add_rn(v, mul_rn(cmu, m));
    v = add_rn(v, mul_rn(c1, mul_rn(a.pc[r], a.pc[cc])));
    if (!hsig) v = add_rn(v, mul_rn(extra, old));
    a.C[e] = v;
  }
}

// ---- after eigh: D = sqrt(eigenvalues)  (_cmaes.py:308) ----------------------------------------
template <typename T>
__global__ void cma_sqrt_d_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl) || a.ctrl->do_eig == 0) return;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.N; k += gridDim.x * blockDim.x) {
    const T d = (T)sqrt((double)a.D[k]);  // no negative guard, like the reference
    a.D[k] = d;
    a.inv_d()[k] = div_rn(T(1), d);
  }
}
// invsqrtC = (B diag(1/D)) B^T  (_cmaes.py:309)
template <typename T>
__global__ void __launch_bounds__(256)
cma_invsqrt_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl) || a.ctrl->do_eig == 0) return;
  const int N = a.N;
  gemm_nt_tile<T>(N, N, N, blockIdx.y * kGemmTile, blockIdx.x * kGemmTile,
                  [=] __device__(int m, int k) { return mul_rn(a.B[(size_t)m * N + k], a.inv_d()[k]); },
                  [=] __device__(int n, int k) { return a.B[(size_t)n * N + k]; },
                  [=] __device__(int m, int n, T acc) { a.invsqrtC[(size_t)m * N + n] = acc; });
}

template <typename T>
__global__ void __launch_bounds__(256)
cma_converge_kernel(const CmaPtrs<T> a) {
  __shared__ double s_red[kRedDoubles];
  if (!es_running(a.ctrl)) return;
  converge_ladder<T>(a.ctrl, a.it, a.N, a.maxiter, a.ilim, a.P, a.xmean, a.xold, a.besthist, a.arfit, a.pc, a.C,
                     a.N + 1, a.B, a.D, a.xtol, a.ftol, a.insigma, s_red);
}

template <typename T>
__global__ void __launch_bounds__(256)
cma_penalty_state_kernel(const CmaPtrs<T> a) {
  __shared__ double s_red[kRedDoubles];
  if (!es_running(a.ctrl)) return;
  penalize_state<T>(a.ctrl, a.it, a.N, a.P, a.hist_cap, a.mueff, a.sorted(), a.xmean, a.xold, a.C, a.N + 1,
                    a.bnd_weights, a.dfithist, a.vec() + 2 * a.N, s_red);
}

template <typename T>
static CmaPtrs<T> cma_ptrs(const sp_cma_state* st, int it) {
  CmaPtrs<T> a;
  a.xmean = (T*)st->xmean;
  a.xold = (T*)st->xold;
  a.pc = (T*)st->pc;
  a.ps = (T*)st->ps;
  a.C = (T*)st->C;
  a.B = (T*)st->B;
  a.D = (T*)st->D;
  a.invsqrtC = (T*)st->invsqrtC;
  a.arx = (T*)st->arx;
  a.arfit = (T*)st->arfit;
  a.Z = (T*)st->Z;
  a.weights = (T*)st->weights;
  a.xscale = (T*)st->xscale;
  a.xshift = (T*)st->xshift;
  a.besthist = (T*)st->besthist;
  a.work = (T*)st->work;
  a.bnd_weights = (T*)st->bnd_weights;
  a.dfithist = (T*)st->dfithist;
  a.rank = st->rank;
  a.ctrl = st->ctrl;
  a.N = st->N;
  a.mu = st->mu;
  a.maxiter = st->maxiter;
  a.ilim = st->ilim;
  a.hist_cap = st->hist_cap;
  a.constraint = st->constraint;
  a.it = it;
  a.P = st->P;
  a.ld = st->ld;
  a.cc = st->cc;
  a.cs = st->cs;
  a.c1 = st->c1;
  a.cmu = st->cmu;
  a.damps = st->damps;
  a.chind = st->chind;
  a.mueff = st->mueff;
  a.xtol = st->xtol;
  a.ftol = st->ftol;
  a.insigma = st->insigma;
  return a;
}

template <typename T>
static int cma_tail(const sp_cma_state* st, const CmaPtrs<T>& a, cudaStream_t s) {
  const int N = st->N, tiles = cdiv(N, kGemmTile);
  cma_sqrt_d_kernel<T><<<cdiv(N, 256), 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  cma_invsqrt_kernel<T><<<dim3(tiles, tiles), 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  cma_converge_kernel<T><<<1, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

// phase A: draws + sampling GEMM
template <typename T>
static int cma_sample(const sp_cma_state* st, int it, cudaStream_t s) {
  const CmaPtrs<T> a = cma_ptrs<T>(st, it);
  const int N = st->N, tiles = cdiv(N, kGemmTile);
  const int64_t P = st->P;
  if (!st->host_z) {
    const int64_t need = cdiv(P * (int64_t)cdiv(N, Num<T>::VEC), 256), cap = (int64_t)sm_count() * 8;
    normal_fill_kernel<T><<<(int)(need < cap ? need : cap), 256, 0, s>>>(a.Z, P, N, st->ld, it, st->seed, kEsZ, st->ctrl);
    SP_CHECK_LAUNCH();
  }
  cma_sample_kernel<T><<<dim3(tiles, cdiv(P, kGemmTile)), 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

// phase B: objective on the (clipped, un-standardised) population
template <typename T>
static int cma_eval(const sp_cma_state* st, int it, cudaStream_t s) {
  const int clip = st->constraint == SP_CONS_PENALIZE ? 1 : 0;
  return eval_launch<T>(st->objective, st->arx, st->P, st->N, st->ld, st->xscale, st->xshift, st->arfit, clip, s,
                        &st->ctrl->base.status);
}

// phase C: penalty, selection, paths, covariance, eigenbasis, termination
template <typename T>
static int cma_update(const sp_cma_state* st, int it, cudaStream_t s) {
  const CmaPtrs<T> a = cma_ptrs<T>(st, it);
  const int N = st->N, tiles = cdiv(N, kGemmTile);
  const int64_t P = st->P;
  if (st->constraint == SP_CONS_PENALIZE) {
    if (rank_launch<T>(a.arfit, P, a.rank, nullptr, s, &st->ctrl->base.status) != cudaSuccess) return SP_ERR_CUDA;
    scatter_sorted_kernel<T><<<cdiv(P, 256) < 1024 ? cdiv(P, 256) : 1024, 256, 0, s>>>(a.arfit, a.rank, a.sorted(), P, st->ctrl);
    SP_CHECK_LAUNCH();
    cma_penalty_state_kernel<T><<<1, 256, 0, s>>>(a);
    SP_CHECK_LAUNCH();
    penalty_add_kernel<T><<<cdiv(P, 8) < sm_count() * 8 ? cdiv(P, 8) : sm_count() * 8, 256, 0, s>>>(
        a.arx, a.vec() + 2 * N, a.arfit, P, N, st->ld, st->ctrl);
    SP_CHECK_LAUNCH();
  }
  if (rank_launch<T>(a.arfit, P, a.rank, nullptr, s, &st->ctrl->base.status) != cudaSuccess) return SP_ERR_CUDA;
  cma_gather_kernel<T><<<cdiv(P, 8) < sm_count() * 4 ? cdiv(P, 8) : sm_count() * 4, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  cma_mean_partial_kernel<T><<<kMeanChunks, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  cma_paths_kernel<T><<<1, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  cma_cov_partial_kernel<T><<<dim3(tiles, tiles, kCovSplits), 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  cma_cov_combine_kernel<T><<<cdiv((int64_t)N * N, 256) < 2048 ? cdiv((int64_t)N * N, 256) : 2048, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  if (st->host_eigh) return SP_OK;  // caller inspects ctrl->do_eig, then sp_cma_finish_generation
  cudaError_t e = jacobi_launch<T>(a.C, N, a.D, a.B, a.jac(), 1, &st->ctrl->do_eig, &st->ctrl->base.status,
                                   &st->ctrl->sweeps, s);
  if (e != cudaSuccess) {
    set_error("sp_cma_update: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  SP_CHECK_LAUNCH();
  return cma_tail<T>(st, a, s);
}

static int cma_check(const sp_cma_state* st, int it) {
  SP_CHECK_ARG(st != nullptr, "null state");
  SP_CHECK_ARG(st->dtype == SP_F32 || st->dtype == SP_F64, "dtype");
  SP_CHECK_ARG(st->N >= 1 && st->N <= 1024, "ndim must be in [1, 1024] for the dense covariance path");
  SP_CHECK_ARG(st->P >= 1 && st->P < (1LL << 31) && st->mu >= 1 && st->mu <= st->P, "popsize / mu");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || st->constraint == SP_CONS_PENALIZE, "constraint");
  const int vec = st->dtype == SP_F32 ? 4 : 2;
  SP_CHECK_ARG(st->ld >= st->N && st->ld % vec == 0, "ld must be a multiple of 16/sizeof(T)");
  SP_CHECK_ARG(st->xmean && st->xold && st->pc && st->ps && st->C && st->B && st->D && st->invsqrtC &&
                   st->arx && st->arfit && st->Z && st->weights && st->xscale && st->xshift && st->besthist &&
                   st->work && st->rank && st->ctrl,
               "null buffer");
  SP_CHECK_ARG(st->constraint == SP_CONS_NONE || (st->bnd_weights && st->dfithist && st->hist_cap >= 2), "Penalize buffers");
  SP_CHECK_ARG(it >= 1 && it <= st->maxiter, "generation index in [1, maxiter]");
  return SP_OK;
}

}  // namespace sp

using namespace sp;

extern "C" {

int64_t sp_cma_work_scalars(int N, int64_t P) {
  return (int64_t)kMeanChunks * N + (int64_t)kCovSplits * N * N + 2LL * N * N + 9LL * N + 64 + P;
}

#define SP_CMA_PHASE(NAME, FN)                                                     \
  int NAME(const sp_cma_state* st, int it, void* stream) {                         \
    int rc = cma_check(st, it);                                                    \
    if (rc) return rc;                                                             \
    return st->dtype == SP_F32 ? FN<float>(st, it, (cudaStream_t)stream) : FN<double>(st, it, (cudaStream_t)stream); \
  }
SP_CMA_PHASE(sp_cma_sample, cma_sample)
SP_CMA_PHASE(sp_cma_update, cma_update)
#undef SP_CMA_PHASE

int sp_cma_generation(const sp_cma_state* st, int it, void* stream) {
  int rc = cma_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG,
               "device objective required (use sp_cma_sample / sp_cma_update around a host evaluation)");
  cudaStream_t s = (cudaStream_t)stream;
  if (st->dtype == SP_F32) {
    if ((rc = cma_sample<float>(st, it, s)) || (rc = cma_eval<float>(st, it, s))) return rc;
    return cma_update<float>(st, it, s);
  }
  if ((rc = cma_sample<double>(st, it, s)) || (rc = cma_eval<double>(st, it, s))) return rc;
  return cma_update<double>(st, it, s);
}

int sp_cma_finish_generation(const sp_cma_state* st, int it, void* stream) {
  int rc = cma_check(st, it);
  if (rc) return rc;
  if (st->dtype == SP_F32) return cma_tail<float>(st, cma_ptrs<float>(st, it), (cudaStream_t)stream);
  return cma_tail<double>(st, cma_ptrs<double>(st, it), (cudaStream_t)stream);
}

int sp_cma_run(const sp_cma_state* st, int it_first, int n, void* stream) {
  SP_CHECK_ARG(st != nullptr && !st->host_z && !st->host_eigh, "sp_cma_run needs in-kernel draws and the device eigensolver");
  for (int g = 0; g < n; ++g) {
    int rc = sp_cma_generation(st, it_first + g, stream);
    if (rc) return rc;
  }
  return SP_OK;
}

}  // extern "C"
