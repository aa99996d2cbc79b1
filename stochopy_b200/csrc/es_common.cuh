// Pieces shared by the evolution-strategy kernels (CMA-ES, VD-CMA): control-block
// gate, N(0,I) fill, block reductions, the termination ladder of converge()
// (_cmaes.py:360-434) and the Penalize box handling (cmaes/_constraints.py:4-82).
#pragma once
#include "rows.cuh"

namespace sp {

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
    normal_block(philox4x32_for(purpose, (uint32_t)b, (uint32_t)row, (uint32_t)it, seed), z);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (b * VEC + e < N) Z[row * ld + b * VEC + e] = z[e];
  }
}

// VD-CMA work-buffer layout (vdcma_impl.cuh; sp_vd_work_scalars in vdcma.cu)
constexpr int kVdChunks = 256;  // row chunks of the weighted sums
constexpr int kVdUpMax = 512;   // CTAs of vd_update_kernel's chunk reduction: 3 N / 32, N <= 2048 -> <= 192

// Block reductions through shared memory.  s_red must hold kRedDoubles doubles.
constexpr int kRedMax = 16;                       // values per combined reduction
constexpr int kRedDoubles = kRedMax * 32 + kRedMax;
enum RedOp { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2 };

__device__ __forceinline__ double red_apply(double a, double b, int op) {
  return op == RED_SUM ? a + b : (op == RED_MIN ? fmin(a, b) : fmax(a, b));
}
__device__ __forceinline__ double red_identity(int op) {
  return op == RED_SUM ? 0.0 : (op == RED_MIN ? 1.0 / 0.0 : -1.0 / 0.0);
}
__device__ __forceinline__ double red_warp(double v, int op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_apply(v, __shfl_xor_sync(0xffffffffu, v, o), op);
  return v;
}

// K values at once (K <= kRedMax), each with its own operator: warp butterflies, one
// shared-memory row per value, warp k folds row k; 3 barriers whatever K is.  Every
// thread of the CTA must call it; every thread receives all K results.  The order of
// the additions is fixed by the launch shape, so results are reproducible.
template <int K>
__device__ __forceinline__ void block_reduce(double (&v)[K], const int (&op)[K], double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = red_warp(v[k], op[k]);
  __syncthreads();  // s_red may still be read by the previous reduction
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) s_red[k * 32 + warp] = v[k];
  }
  __syncthreads();
  for (int k = warp; k < K; k += nw) {
    double x = lane < nw ? s_red[k * 32 + lane] : red_identity(op[k]);
    x = red_warp(x, op[k]);
    if (lane == 0) s_red[kRedMax * 32 + k] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = s_red[kRedMax * 32 + k];
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  double x[1] = {v};
  const int op[1] = {RED_SUM};
  block_reduce<1>(x, op, s_red);
  return x[0];
}

// ---- termination ladder (_cmaes.py:360-434), one CTA ----------------------------------------------
// `with_basis`: CMA-ES passes B and D (rungs -2, -4); VD-CMA does not (_vdcma.py:380-396).
template <typename T>
__device__ void converge_ladder(sp_es_ctrl* c, int it, int N, int maxiter, int ilim, int64_t P, const T* xmean,
                                const T* xold, const T* besthist, const T* arfit, const T* pc, const T* diagC,
                                int diag_stride, const T* B, const T* D, double xtol, double ftol, double insigma,
                                double* s_red, const double* fit_range = nullptr) {
  const int tid = threadIdx.x;
  const double sigma = c->sigma;
  const double best = c->base.gfit;
  const double inf = 1.0 / 0.0;
  double dsq = 0.0, fmin_ = inf, fmax_ = -inf, hmin = inf, hmax = -inf;
  double wmin = inf, wmax = -inf, dmin = inf, dmax = -inf, sdmax = 0.0;
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
  if (fit_range != nullptr) {  // this thread's share of min / max arfit was scanned by the caller
    fmin_ = fit_range[0];
    fmax_ = fit_range[1];
  } else {
    for (int64_t i = tid; i < P; i += blockDim.x) {
      fmin_ = fmin(fmin_, (double)arfit[i]);
      fmax_ = fmax(fmax_, (double)arfit[i]);
    }
  }
  for (int i = tid; i < maxiter; i += blockDim.x) {  // zero padded history, all of it (_cmaes.py:424-427)
    hmin = fmin(hmin, (double)besthist[i]);
    hmax = fmax(hmax, (double)besthist[i]);
    if (i >= it - ilim && i <= it) {  // window incl. one not-yet-written zero (_cmaes.py:412-414)
      wmin = fmin(wmin, (double)besthist[i]);
      wmax = fmax(wmax, (double)besthist[i]);
    }
  }
  double v[14] = {dsq, fmin_, fmax_, hmin, hmax, wmin, wmax, dmin, dmax, sdmax,
                  (double)axis_all, (double)coord_any, (double)tolxup_any, (double)tolx_all};
  const int op[14] = {RED_SUM, RED_MIN, RED_MAX, RED_MIN, RED_MAX, RED_MIN, RED_MAX, RED_MIN, RED_MAX, RED_MAX,
                      RED_MIN, RED_MAX, RED_MAX, RED_MIN};
  block_reduce<14>(v, op, s_red);
  dsq = v[0], fmin_ = v[1], fmax_ = v[2], hmin = v[3], hmax = v[4], wmin = v[5], wmax = v[6];
  dmin = v[7], dmax = v[8], sdmax = v[9];
  axis_all = v[10] > 0.5, coord_any = v[11] > 0.5, tolxup_any = v[12] > 0.5, tolx_all = v[13] > 0.5;
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


static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace sp
