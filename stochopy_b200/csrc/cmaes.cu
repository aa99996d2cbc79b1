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
//   cma_mean         xmean' = sum_i w[rank_i] arx_i                 (deterministic two-stage)
//   cma_paths        ps, hsig, pc, sigma', eigen-update decision    one CTA
//   cma_cov          C' = (1-c1-cmu) C + cmu A^T diag(w) A + ...    GEMM TN  2 mu N^2 flop
//   [jacobi + post]  C = B diag(D^2) B^T, invsqrtC                  when due
//   cma_converge     termination ladder                             one CTA
#include "linalg.cuh"
#include "rows.cuh"

namespace sp {

constexpr int kMeanChunks = 64;
constexpr int kCovSplits = 8;

template <typename T>
struct CmaPtrs {
  T *xmean, *xold, *pc, *ps, *C, *B, *D, *BD, *invsqrtC, *arx, *arfit, *Z, *weights, *xscale, *xshift, *besthist,
      *work, *bnd_weights, *dfithist;
  int32_t* rank;
  sp_es_ctrl* ctrl;
  int N, mu, maxiter, ilim, hist_cap, constraint, it;
  int64_t P, ld;
  double cc, cs, c1, cmu, damps, chind, mueff, xtol, ftol, insigma;
  // workspace slices
  __host__ __device__ T* mean_part() const { return work; }                                  // kMeanChunks * N
  __host__ __device__ T* cov_part() const { return work + (size_t)kMeanChunks * N; }          // kCovSplits * N * N
  __host__ __device__ T* jac() const { return cov_part() + (size_t)kCovSplits * N * N; }      // 2 * N * N
  __host__ __device__ T* vec() const { return jac() + 2 * (size_t)N * N; }                    // 8 * N (diff, coef, ...)
  __host__ __device__ T* sorted() const { return vec() + 8 * (size_t)N; }                     // P (Penalize percentiles)
};

__device__ __forceinline__ bool es_running(const sp_es_ctrl* c) {
  return *reinterpret_cast<const volatile int32_t*>(&c->base.status) == SP_RUNNING;
}

// ---- Z ~ N(0, I) --------------------------------------------------------------------
template <typename T>
__global__ void normal_fill_kernel(T* __restrict__ Z, int64_t P, int N, int64_t ld, int it, uint64_t seed,
                                   uint32_t purpose, const sp_es_ctrl* ctrl) {
  constexpr int VEC = Num<T>::VEC;
  if (ctrl != nullptr && !es_running(ctrl)) return;
  const int nb = (N + VEC - 1) / VEC;
  const int64_t total = P * (int64_t)nb;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / nb;
    const int b = (int)(t - row * nb);
    T z[VEC];
    normal_block(philox4x32((uint32_t)b, (uint32_t)row, (uint32_t)it, purpose, seed), z);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (b * VEC + e < N) Z[row * ld + b * VEC + e] = z[e];
  }
}

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

// ---- weighted mean of the mu best, stage 1 (_cmaes.py:274) ---------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
cma_mean_partial_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const int64_t per = (a.P + kMeanChunks - 1) / kMeanChunks;
  const int64_t i0 = blockIdx.x * per, i1 = (i0 + per < a.P) ? i0 + per : a.P;
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
    T acc = 0;
    for (int64_t i = i0; i < i1; ++i) {
      const int r = a.rank[i];
      if (r < a.mu) acc += a.weights[r] * a.arx[i * a.ld + n];
    }
    a.mean_part()[(size_t)blockIdx.x * a.N + n] = acc;
  }
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
  return t;
}

// ---- paths, step size, eigen decision: one CTA (_cmaes.py:272-301) ---------------------------
template <typename T>
__global__ void __launch_bounds__(256)
cma_paths_kernel(const CmaPtrs<T> a) {
  __shared__ double s_red[8];
  __shared__ int s_best;
  sp_es_ctrl* c = a.ctrl;
  if (!es_running(c)) return;
  const int N = a.N, tid = threadIdx.x;
  T* diff = a.vec();
  const T sigma = (T)c->sigma_gen;
  if (tid == 0) s_best = 0x7fffffff;
  __syncthreads();
  for (int64_t i = tid; i < a.P; i += blockDim.x)
    if (a.rank[i] == 0) s_best = (int)i;
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
    T dot = 0;
    const T* row = a.invsqrtC + (size_t)n * N;
    for (int k = 0; k < N; ++k) dot += row[k] * diff[k];
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
    const int b = s_best;
    const double best = (double)a.arfit[b];
    c->base.gbest_row = b;
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

// ---- rank-mu partial products (_cmaes.py:290-293) ---------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
cma_cov_partial_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const T sigma = (T)a.ctrl->sigma_gen;
  const int N = a.N;
  const int64_t per = (a.P + kCovSplits - 1) / kCovSplits;
  const int64_t i0 = blockIdx.z * per, i1 = (i0 + per < a.P) ? i0 + per : a.P;
  T* out = a.cov_part() + (size_t)blockIdx.z * N * N;
  const int64_t ld = a.ld;
  gemm_tn_tile<T>(N, N, (int)i0, (int)i1, blockIdx.y * kGemmTile, blockIdx.x * kGemmTile,
                  [=] __device__(int i, int r) {
                    const int rk = a.rank[i];
                    if (rk >= a.mu) return T(0);
                    return mul_rn(div_rn(sub_rn(a.arx[(int64_t)i * ld + r], a.xold[r]), sigma), a.weights[rk]);
                  },
                  [=] __device__(int i, int cc) {
                    if (a.rank[i] >= a.mu) return T(0);
                    return div_rn(sub_rn(a.arx[(int64_t)i * ld + cc], a.xold[cc]), sigma);
                  },
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
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.N; k += gridDim.x * blockDim.x)
    a.D[k] = (T)sqrt((double)a.D[k]);  // no negative guard, like the reference
}
// invsqrtC = (B diag(1/D)) B^T  (_cmaes.py:309)
template <typename T>
__global__ void __launch_bounds__(256)
cma_invsqrt_kernel(const CmaPtrs<T> a) {
  if (!es_running(a.ctrl) || a.ctrl->do_eig == 0) return;
  const int N = a.N;
  gemm_nt_tile<T>(N, N, N, blockIdx.y * kGemmTile, blockIdx.x * kGemmTile,
                  [=] __device__(int m, int k) { return mul_rn(a.B[(size_t)m * N + k], div_rn(T(1), a.D[k])); },
                  [=] __device__(int n, int k) { return a.B[(size_t)n * N + k]; },
                  [=] __device__(int m, int n, T acc) { a.invsqrtC[(size_t)m * N + n] = acc; });
}

// ---- termination ladder (_cmaes.py:360-434), one CTA ----------------------------------------------
// `with_basis`: CMA-ES passes B and D (rungs -2, -4); VD-CMA does not (_vdcma.py:380-396).
template <typename T>
__device__ void converge_ladder(sp_es_ctrl* c, int it, int N, int maxiter, int ilim, int64_t P, const T* xmean,
                                const T* xold, const T* besthist, const T* arfit, const T* pc, const T* diagC,
                                int diag_stride, const T* B, const T* D, double xtol, double ftol, double insigma,
                                double* s_red) {
  const int tid = threadIdx.x;
  const double sigma = c->sigma;
  const double best = c->base.gfit;
  double dsq = 0.0, fmin_ = 1.0 / 0.0, fmax_ = -1.0 / 0.0, hmin = 1.0 / 0.0, hmax = -1.0 / 0.0;
  double wmin = 1.0 / 0.0, wmax = -1.0 / 0.0, dmin = 1.0 / 0.0, dmax = -1.0 / 0.0, sdmax = 0.0;
  int axis_all = 1, coord_any = 0, tolxup_any = 0, tolx_all = 1;
  const int ax = it % N;
  for (int n = tid; n < N; n += blockDim.x) {
    const double d = (double)xold[n] - (double)xmean[n];
    dsq += d * d;
    const double sd = sqrt((double)diagC[(size_t)n * diag_stride]);
    sdmax = fmax(sdmax, sd);
    if (0.2 * sigma * sd < 1.0e-10) coord_any = 1;
    if (sigma * sd > 1.0e3 * insigma) tolxup_any = 1;
    if (!(sigma * fabs((double)pc[n]) < 1.0e-11 * insigma)) tolx_all = 0;
    if (B != nullptr) {
      if (!(fabs(0.1 * sigma * (double)B[(size_t)n * N + ax] * (double)D[ax]) < 1.0e-10)) axis_all = 0;
      dmin = fmin(dmin, (double)D[n]);
      dmax = fmax(dmax, (double)D[n]);
    }
  }
  for (int64_t i = tid; i < P; i += blockDim.x) {
    fmin_ = fmin(fmin_, (double)arfit[i]);
    fmax_ = fmax(fmax_, (double)arfit[i]);
  }
  for (int i = tid; i < maxiter; i += blockDim.x) {  // zero padded history, all of it (_cmaes.py:424-427)
    hmin = fmin(hmin, (double)besthist[i]);
    hmax = fmax(hmax, (double)besthist[i]);
    if (i >= it - ilim && i <= it) {  // window incl. one not-yet-written zero (_cmaes.py:412-414)
      wmin = fmin(wmin, (double)besthist[i]);
      wmax = fmax(wmax, (double)besthist[i]);
    }
  }
  // block reductions (sum / min / max / and / or) through shared memory
  auto red = [&](double v, int op) {
    for (int o = 16; o > 0; o >>= 1) {
      const double u = __shfl_xor_sync(0xffffffffu, v, o);
      v = op == 0 ? v + u : (op == 1 ? fmin(v, u) : fmax(v, u));
    }
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    double t = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = op == 0 ? t + s_red[w] : (op == 1 ? fmin(t, s_red[w]) : fmax(t, s_red[w]));
    return t;
  };
  dsq = red(dsq, 0);
  fmin_ = red(fmin_, 1);
  fmax_ = red(fmax_, 2);
  hmin = red(hmin, 1);
  hmax = red(hmax, 2);
  wmin = red(wmin, 1);
  wmax = red(wmax, 2);
  dmin = red(dmin, 1);
  dmax = red(dmax, 2);
  sdmax = red(sdmax, 2);
  axis_all = red((double)axis_all, 1) > 0.5;
  coord_any = red((double)coord_any, 2) > 0.5;
  tolxup_any = red((double)tolxup_any, 2) > 0.5;
  tolx_all = red((double)tolx_all, 1) > 0.5;
  if (tid == 0) {
    int st = SP_RUNNING;
    if (it >= maxiter) st = -1;
    else if (sqrt(dsq) <= xtol && best < ftol) st = 0;
    else if (best <= ftol) st = 1;
    else if (B != nullptr && axis_all) st = -2;
    else if (coord_any) st = -3;
    else if (B != nullptr && dmax > 1.0e7 * dmin) st = -4;
    else if (it >= ilim && wmax - wmin < 1.0e-10) st = -5;
    else if (tolxup_any) st = -6;
    else if (it > 2 && fmax(fmax_, hmax) - fmin(fmin_, hmin) < 1.0e-12) st = -7;
    else if (tolx_all && sigma * sdmax < 1.0e-11 * insigma) st = -8;
    c->base.nit = it;
    c->base.status = st;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
cma_converge_kernel(const CmaPtrs<T> a) {
  __shared__ double s_red[8];
  if (!es_running(a.ctrl)) return;
  converge_ladder<T>(a.ctrl, a.it, a.N, a.maxiter, a.ilim, a.P, a.xmean, a.xold, a.besthist, a.arfit, a.pc, a.C,
                     a.N + 1, a.B, a.D, a.xtol, a.ftol, a.insigma, s_red);
}

// ---- Penalize (cmaes/_constraints.py:4-82) ------------------------------------------------------------
// sorted[rank[i]] = arfit[i] (raw fitness of the clipped population)
template <typename T>
__global__ void scatter_sorted_kernel(const T* __restrict__ fit, const int32_t* __restrict__ rank, T* __restrict__ sorted,
                                      int64_t P, const sp_es_ctrl* ctrl) {
  if (!es_running(ctrl)) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x)
    sorted[rank[i]] = fit[i];
}

// np.percentile(..., method="linear"): lerp as numpy does it (a + (b-a) t, from b when t >= 0.5)
template <typename F>
__device__ __forceinline__ double np_percentile(const double q, int64_t P, const F& at) {
  const double pos = q / 100.0 * (double)(P - 1);
  const int64_t lo = (int64_t)floor(pos);
  const int64_t hi = lo + 1 < P ? lo + 1 : P - 1;
  const double t = pos - (double)lo;
  const double a = at(lo), b = at(hi);
  const double d = b - a;
  return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

// state update: delta from the inter-quartile range, dfithist ring, boundary weights,
// coef[j] = bnd_weights[j] / bnd_scale[j].  One CTA; diagC(n) = diag[n * diag_stride].
template <typename T>
__device__ void penalize_state(sp_es_ctrl* c, int it, int N, int64_t P, int hist_cap, double mueff, const T* sorted,
                               const T* xmean, const T* xold, const T* diag, int diag_stride, T* bnd_weights,
                               T* dfithist, T* coef, double* s_red) {
  const int tid = threadIdx.x;
  const double sigma = c->sigma;
  double dsum = 0.0, lsum = 0.0;
  int out_any = 0;
  for (int n = tid; n < N; n += blockDim.x) {
    const double dc = (double)diag[(size_t)n * diag_stride];
    dsum += dc;
    lsum += log(dc);
    const double xm = (double)xmean[n];
    if (xm < -1.0 || xm > 1.0) out_any = 1;
  }
  dsum = block_sum(dsum, s_red);
  lsum = block_sum(lsum, s_red);
  out_any = block_sum((double)out_any, s_red) > 0.5;
  __shared__ double s_w0;
  __shared__ int s_set;
  if (tid == 0) {
    auto at = [&](int64_t k) { return (double)sorted[k]; };
    const double q25 = np_percentile(25.0, P, at), q75 = np_percentile(75.0, P, at);
    double delta = (q75 - q25) / (double)N / (dsum / (double)N) / (sigma * sigma);
    int len = c->hist_len;
    if (delta == 0.0) {  // smallest positive delta seen so far
      double m = 1.0 / 0.0;
      for (int k = 0; k < len; ++k)
        if ((double)dfithist[k] > 0.0) m = fmin(m, (double)dfithist[k]);
      delta = m;
    } else if (!c->validfitval) {
      len = 0;
      c->validfitval = 1;
    }
    if ((double)len < 20.0 + (3.0 * N) / (double)P && len < hist_cap) {
      dfithist[len++] = (T)delta;
    } else {
      for (int k = 1; k < len; ++k) dfithist[k - 1] = dfithist[k];
      dfithist[len - 1] = (T)delta;
    }
    c->hist_len = len;
    s_set = 0;
    if (c->iniphase && out_any) {  // bnd_weights = 2.0002 * median(dfithist)
      // selection sort on a copy in aux space is overkill: len <= hist_cap is tiny
      double med;
      {
        // median by counting ranks
        int lo_i = (len - 1) / 2, hi_i = len / 2;
        double vlo = 0.0, vhi = 0.0;
        for (int k = 0; k < len; ++k) {
          int rk = 0;
          for (int j = 0; j < len; ++j) rk += ((double)dfithist[j] < (double)dfithist[k]) || (dfithist[j] == dfithist[k] && j < k);
          if (rk == lo_i) vlo = (double)dfithist[k];
          if (rk == hi_i) vhi = (double)dfithist[k];
        }
        med = 0.5 * (vlo + vhi);
      }
      s_w0 = 2.0002 * med;
      s_set = 1;
      if (c->validfitval && it > 2) c->iniphase = 0;
    }
  }
  __syncthreads();
  const double lmean = lsum / (double)N;
  const double thr = 3.0 * fmax(1.0, sqrt((double)N / mueff)) * sigma;
  const double grow = pow(1.2, fmin(1.0, mueff / 10.0 / (double)N));
  for (int n = tid; n < N; n += blockDim.x) {
    double w = s_set ? s_w0 : (double)bnd_weights[n];
    const double xm = (double)xmean[n], dc = (double)diag[(size_t)n * diag_stride];
    if (out_any) {
      const bool ti = xm < -1.0 || xm > 1.0;
      const double tx = xm - (xm > 1.0 ? 1.0 : xm);  // lower clip lost, as in the reference (:53-54)
      const double dm = xm - (double)xold[n];
      const int s1 = (tx > 0.0) - (tx < 0.0), s2 = (dm > 0.0) - (dm < 0.0);
      if (ti && fabs(tx) > thr * sqrt(dc) && s1 == s2) w *= grow;
    }
    bnd_weights[n] = (T)w;
    coef[n] = (T)(w / exp(0.9 * (log(dc) - lmean)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
cma_penalty_state_kernel(const CmaPtrs<T> a) {
  __shared__ double s_red[8];
  if (!es_running(a.ctrl)) return;
  penalize_state<T>(a.ctrl, a.it, a.N, a.P, a.hist_cap, a.mueff, a.sorted(), a.xmean, a.xold, a.C, a.N + 1,
                    a.bnd_weights, a.dfithist, a.vec() + 2 * a.N, s_red);
}

// arfit[i] += sum_j (clip(x_ij) - x_ij)^2 coef_j
template <typename T>
__global__ void penalty_add_kernel(const T* __restrict__ arx, const T* __restrict__ coef, T* __restrict__ arfit,
                                   int64_t P, int N, int64_t ld, const sp_es_ctrl* ctrl) {
  if (!es_running(ctrl)) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < P; i += nw) {
    T acc = 0;
    for (int j = lane; j < N; j += 32) {
      const T x = arx[i * ld + j];
      const T v = x < T(-1) ? T(-1) : (x > T(1) ? T(1) : x);
      const T d = v - x;
      acc += d * d * coef[j];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) arfit[i] = add_rn(arfit[i], acc);
  }
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
  a.BD = (T*)st->BD;
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

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

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
  return eval_launch<T>(st->objective, st->arx, st->P, st->N, st->ld, st->xscale, st->xshift, st->arfit, clip, s);
}

// phase C: penalty, selection, paths, covariance, eigenbasis, termination
template <typename T>
static int cma_update(const sp_cma_state* st, int it, cudaStream_t s) {
  const CmaPtrs<T> a = cma_ptrs<T>(st, it);
  const int N = st->N, tiles = cdiv(N, kGemmTile);
  const int64_t P = st->P;
  const int rank_grid = cdiv(P, kThreads);
  if (st->constraint == SP_CONS_PENALIZE) {
    rank_kernel<T><<<rank_grid, kThreads, 0, s>>>(a.arfit, P, a.rank, nullptr);
    SP_CHECK_LAUNCH();
    scatter_sorted_kernel<T><<<cdiv(P, 256) < 1024 ? cdiv(P, 256) : 1024, 256, 0, s>>>(a.arfit, a.rank, a.sorted(), P, st->ctrl);
    SP_CHECK_LAUNCH();
    cma_penalty_state_kernel<T><<<1, 256, 0, s>>>(a);
    SP_CHECK_LAUNCH();
    penalty_add_kernel<T><<<cdiv(P, 8) < sm_count() * 8 ? cdiv(P, 8) : sm_count() * 8, 256, 0, s>>>(
        a.arx, a.vec() + 2 * N, a.arfit, P, N, st->ld, st->ctrl);
    SP_CHECK_LAUNCH();
  }
  rank_kernel<T><<<rank_grid, kThreads, 0, s>>>(a.arfit, P, a.rank, nullptr);
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
  return (int64_t)kMeanChunks * N + (int64_t)kCovSplits * N * N + 2LL * N * N + 8LL * N + P;
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
