// VD-CMA: restricted covariance C = D (I + v v^T) D, everything O(N) per individual.
// Reference: stochopy/optimize/vdcma/_vdcma.py:235-409 (generation), :426-458
// (pvec_and_qvec, ngv_ngd), converge from cmaes/_cmaes.py:360-434 without B, D.
//
//   vd_inject        dy = |g| / sqrt(mnorm) dx                       one CTA      [from gen 2]
//   vd_sample_eval   z -> y -> x, (y/d).vn, objective               row tiles, HBM: write 2 rows
//   [Penalize]       rank -> percentiles -> weights -> arfit += penalty
//   rank
//   vd_wsum/wreduce  S_x, S_y, P_mu, Q_mu over the mu best           column-parallel, read 2 mu rows
//   vd_update        mean, sigma (rank gap of rows 0/1), pc, natural gradient on (v, D), ladder
//   vd_refresh       |v|^2, vn, diagC for the next generation
#include "es_common.cuh"

namespace sp {

constexpr int kVdChunks = 128;

template <typename T>
struct VdPtrs {
  T *xmean, *xold, *dx, *pc, *dvec, *vvec, *vn, *diagC, *dy, *ginj, *arx, *ary, *yvn, *arfit, *weights, *xscale,
      *xshift, *besthist, *work, *bnd_weights, *dfithist;
  int32_t* rank;
  sp_es_ctrl* ctrl;
  int N, mu, maxiter, ilim, hist_cap, constraint, objective, it, host_z, evaluate;
  int64_t P, ld;
  double cc, c1, cmu, mueff, wsum, xtol, ftol, insigma;
  uint64_t seed;
  __host__ __device__ T* part() const { return work; }                               // kVdChunks * 4 * N
  __host__ __device__ T* coef() const { return work + (size_t)kVdChunks * 4 * N; }    // N
  __host__ __device__ T* tmp() const { return coef() + N; }                           // 8 * N
  __host__ __device__ T* sorted() const { return tmp() + 8 * (size_t)N; }             // P
  __host__ __device__ T* sums() const { return sorted() + P; }                        // 4 * N reduced partials
};

// ctrl->aux: [0] |v|^2, [1] |v|
template <typename T>
__global__ void __launch_bounds__(256)
vd_refresh_kernel(const VdPtrs<T> a) {
  __shared__ double s_red[8];
  sp_es_ctrl* c = a.ctrl;
  double sq = 0.0;
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) sq += (double)a.vvec[n] * (double)a.vvec[n];
  const double nv2 = block_sum(sq, s_red), nv = sqrt(nv2);
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
    const T v = a.vvec[n], d = a.dvec[n];
    a.vn[n] = div_rn(v, (T)nv);
    a.diagC[n] = mul_rn(mul_rn(d, add_rn(T(1), mul_rn(v, v))), d);  // _vdcma.py:251-256
  }
  if (threadIdx.x == 0) {
    c->aux[0] = nv2;
    c->aux[1] = nv;
  }
}

// injection, _vdcma.py:243-246
template <typename T>
__global__ void __launch_bounds__(256)
vd_inject_kernel(const VdPtrs<T> a) {
  constexpr int VEC = Num<T>::VEC;
  __shared__ double s_red[8];
  sp_es_ctrl* c = a.ctrl;
  if (!es_running(c) || !c->inject) return;
  const double nv2 = c->aux[0];
  double g2 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
    T g;
    if (a.host_z) {
      g = a.ginj[n];
    } else {
      T z[VEC];
      normal_block(philox4x32((uint32_t)(n / VEC), 0u, (uint32_t)a.it, kVdInject, a.seed), z);
      g = z[n % VEC];
    }
    g2 += (double)g * (double)g;
    const double ddx = (double)div_rn(a.dx[n], a.dvec[n]);
    s1 += ddx * ddx;
    s2 += ddx * (double)a.vvec[n];
  }
  g2 = block_sum(g2, s_red);
  s1 = block_sum(s1, s_red);
  s2 = block_sum(s2, s_red);
  const double mnorm = s1 - s2 * s2 / (1.0 + nv2);
  const T k = (T)(sqrt(g2) / sqrt(mnorm));
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) a.dy[n] = mul_rn(k, a.dx[n]);
}

// row-local sampling + objective, _vdcma.py:239-277
template <typename T, int CH, int LPR>
__global__ void __launch_bounds__(kThreads)
vd_sample_eval_kernel(const VdPtrs<T> a) {
  using TL = Tile<T, CH, LPR>;
  constexpr int VEC = Num<T>::VEC;
  const sp_es_ctrl* c = a.ctrl;
  if (!es_running(c)) return;
  const int lane = threadIdx.x & 31, l = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t groups = (a.P + TL::RPW - 1) / TL::RPW;
  const int ld = (int)a.ld, N = a.N;
  const T sigma = (T)c->sigma;
  const T fac = (T)(sqrt(1.0 + c->aux[0]) - 1.0);
  const bool inject = c->inject != 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) const_cast<sp_es_ctrl*>(c)->sigma_gen = c->sigma;
  const bool clip = a.constraint == SP_CONS_PENALIZE;

  TL vn, dv;
  vn.load(a.vn, l, ld);
  dv.load(a.dvec, l, ld);
  for (int64_t g = warp; g < groups; g += nwarps) {
    int64_t row = g * TL::RPW + sub;
    const bool live = row < a.P;
    if (!live) row = a.P - 1;
    TL y;
    if (a.host_z) {
      y.load(a.ary + row * a.ld, l, ld);
    } else {
#pragma unroll
      for (int cc = 0; cc < CH; ++cc) {
        const int j0 = TL::col(cc, l, 0);
        T z[VEC];
        if (j0 < N) normal_block(philox4x32((uint32_t)(j0 / VEC), (uint32_t)row, (uint32_t)a.it, kEsZ, a.seed), z);
#pragma unroll
        for (int e = 0; e < VEC; ++e) y.v[cc][e] = (j0 + e < N) ? z[e] : T(0);
      }
    }
    T zv = 0;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc)
#pragma unroll
      for (int e = 0; e < VEC; ++e) zv += y.v[cc][e] * vn.v[cc][e];
    zv = group_sum<LPR>(zv);
    const bool inj_row = inject && row < 2;
    if (inj_row) {
      TL dyv;
      dyv.load(a.dy, l, ld);
#pragma unroll
      for (int cc = 0; cc < CH; ++cc)
#pragma unroll
        for (int e = 0; e < VEC; ++e) y.v[cc][e] = row == 0 ? dyv.v[cc][e] : -dyv.v[cc][e];
    } else {
#pragma unroll
      for (int cc = 0; cc < CH; ++cc)
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          y.v[cc][e] = mul_rn(dv.v[cc][e], add_rn(y.v[cc][e], mul_rn(fac, mul_rn(zv, vn.v[cc][e]))));
    }
    // (y / dvec) . vn for the rank-mu update (_vdcma.py:428 with y = ary / dvec)
    T yv = 0;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc)
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (TL::col(cc, l, e) < N) yv += div_rn(y.v[cc][e], dv.v[cc][e]) * vn.v[cc][e];
    yv = group_sum<LPR>(yv);
    TL x;
    {
      TL xm;
      xm.load(a.xmean, l, ld);
#pragma unroll
      for (int cc = 0; cc < CH; ++cc)
#pragma unroll
        for (int e = 0; e < VEC; ++e) x.v[cc][e] = add_rn(xm.v[cc][e], mul_rn(sigma, y.v[cc][e]));
    }
    if (live) {
      y.store(a.ary + row * a.ld, l, ld);
      x.store(a.arx + row * a.ld, l, ld);
      if (l == 0) a.yvn[row] = yv;
    }
    if (!a.evaluate) continue;
    {
      TL sc, sh;
      sc.load(a.xscale, l, ld);
      sh.load(a.xshift, l, ld);
#pragma unroll
      for (int cc = 0; cc < CH; ++cc)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          T v = x.v[cc][e];
          if (clip) v = v < T(-1) ? T(-1) : (v > T(1) ? T(1) : v);
          x.v[cc][e] = add_rn(mul_rn(v, sc.v[cc][e]), sh.v[cc][e]);
        }
    }
    const T f = evaluate_tile<T, CH, LPR>(a.objective, x, l, N);
    if (live && l == 0) a.arfit[row] = f;
  }
}

// weighted sums over the mu best, column-parallel; chunk partials in a fixed order.
// part[chunk][0..3][n] = S_x, S_y, P_mu, Q_mu   (_vdcma.py:291, 313, 426-441)
template <typename T>
__global__ void __launch_bounds__(256)
vd_wsum_kernel(const VdPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (a.P + kVdChunks - 1) / kVdChunks;
  const int64_t i0 = blockIdx.y * per, i1 = (i0 + per < a.P) ? i0 + per : a.P;
  if (n >= a.N) return;
  const double nv2 = a.ctrl->aux[0];
  const T k1 = (T)(nv2 / (1.0 + nv2));
  const T vn = a.vn[n], dv = a.dvec[n];
  const bool with_mu = a.cmu != 0.0;
  T sx = 0, sy = 0, pm = 0, qm = 0;
  for (int64_t i = i0; i < i1; ++i) {
    const int r = a.rank[i];
    if (r >= a.mu) continue;
    const T w = a.weights[r];
    const T y = a.ary[i * a.ld + n];
    sx += w * a.arx[i * a.ld + n];
    sy += w * y;
    if (with_mu) {
      const T yd = div_rn(y, dv), yv = a.yvn[i];
      pm += w * (yd * yd - k1 * (yv * (yd * vn)) - T(1));
      qm += w * (yv * yd - (T(0.5) * (yv * yv + T(1) + (T)nv2)) * vn);
    }
  }
  T* out = a.part() + (size_t)blockIdx.y * 4 * a.N;
  out[n] = sx;
  out[a.N + n] = sy;
  out[2 * a.N + n] = pm;
  out[3 * a.N + n] = qm;
}

// chunk partials -> sums[q][n] in a fixed order (deterministic), one thread per (q, n)
template <typename T>
__global__ void __launch_bounds__(256)
vd_wreduce_kernel(const VdPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * a.N) return;
  const T* p = a.part() + e;
  T acc = 0;
#pragma unroll 8
  for (int g = 0; g < kVdChunks; ++g) acc += p[(size_t)g * 4 * a.N];
  a.sums()[e] = acc;
}

// mean, step size, paths, natural gradient, termination: one CTA (_vdcma.py:290-396)
template <typename T>
__global__ void __launch_bounds__(256)
vd_update_kernel(const VdPtrs<T> a) {
  __shared__ double s_red[8];
  __shared__ int s_best, s_r0, s_r1;
  sp_es_ctrl* c = a.ctrl;
  if (!es_running(c)) return;
  const int N = a.N, tid = threadIdx.x;
  T* Sy = a.tmp();
  T* pv = a.tmp() + N;
  T* qv = a.tmp() + 2 * N;
  T* sv = a.tmp() + 3 * N;
  T* ngv = a.tmp() + 4 * N;
  T* ngd = a.tmp() + 5 * N;
  const double nv2 = c->aux[0], nv = c->aux[1];
  if (tid == 0) s_best = 0;
  __syncthreads();
  for (int64_t i = tid; i < a.P; i += blockDim.x)
    if (a.rank[i] == 0) s_best = (int)i;
  if (tid == 0) {
    s_r0 = a.rank[0];
    s_r1 = a.P > 1 ? a.rank[1] : 0;
  }
  // reduced sums -> dx, xmean, S_y, P_mu, Q_mu
  for (int n = tid; n < N; n += blockDim.x) {
    const T* p = a.sums();
    const T sx = p[n], sy = p[N + n], pm = p[2 * N + n], qm = p[3 * N + n];
    const T xm = a.xmean[n];
    const T dx = sub_rn(sx, mul_rn((T)a.wsum, xm));  // _vdcma.py:291
    a.dx[n] = dx;
    a.xold[n] = xm;
    a.xmean[n] = add_rn(xm, dx);
    Sy[n] = sy;
    pv[n] = pm;
    qv[n] = qm;
  }
  __syncthreads();
  // sigma from the rank gap of the injected pair, _vdcma.py:299-307
  bool hsig = true;
  double sigma = c->sigma_gen;
  if (c->inject) {
    const double alpha_act = (double)(s_r1 - s_r0) / ((double)a.P - 1.0);
    const double ps = c->vd_ps + 0.3 * (alpha_act - c->vd_ps);
    sigma *= exp(ps / sqrt((double)N));
    hsig = ps < 0.5;
    if (tid == 0) c->vd_ps = ps;
  }
  __syncthreads();
  const T kpc = (T)sqrt(a.cc * (2.0 - a.cc) * a.mueff);
  double vmax = 0.0;
  for (int n = tid; n < N; n += blockDim.x) {
    T v = mul_rn(a.pc[n], (T)(1.0 - a.cc));
    if (hsig) v = add_rn(v, mul_rn(kpc, Sy[n]));
    a.pc[n] = v;
    const double vn = (double)a.vn[n];
    vmax = fmax(vmax, vn * vn);
  }
  // block max of vnn
  for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = vmax;
  __syncthreads();
  vmax = s_red[0];
  for (int w = 1; w < 8; ++w) vmax = fmax(vmax, s_red[w]);
  __syncthreads();
  // alpha and friends, _vdcma.py:318-329
  const double gamma = 1.0 / sqrt(1.0 + nv2);
  double alpha = sqrt(nv2 * nv2 + (1.0 + nv2) / vmax * (2.0 - gamma)) / (2.0 + nv2), beta = 0.0;
  if (alpha < 1.0) beta = (4.0 - (2.0 - gamma) / vmax) / ((1.0 + 2.0 / nv2) * (1.0 + 2.0 / nv2));
  else alpha = 1.0;
  const double bsca = 2.0 * alpha * alpha - beta;
  // rank-one vectors from pc / dvec, then p = cmu p_mu (+ c1 p_1), q likewise
  double yv1 = 0.0;
  for (int n = tid; n < N; n += blockDim.x) yv1 += (double)div_rn(a.pc[n], a.dvec[n]) * (double)a.vn[n];
  yv1 = block_sum(yv1, s_red);
  const double k1 = nv2 / (1.0 + nv2);
  double vq = 0.0;
  for (int n = tid; n < N; n += blockDim.x) {
    const double vn = (double)a.vn[n];
    double p = a.cmu == 0.0 ? 0.0 : a.cmu * (double)pv[n];
    double q = a.cmu == 0.0 ? 0.0 : a.cmu * (double)qv[n];
    if (hsig && a.c1 != 0.0) {
      const double y1 = (double)div_rn(a.pc[n], a.dvec[n]);
      p += a.c1 * (y1 * y1 - k1 * (yv1 * y1 * vn) - 1.0);
      q += a.c1 * (yv1 * y1 - (0.5 * (yv1 * yv1 + 1.0 + nv2)) * vn);
    }
    pv[n] = (T)p;
    qv[n] = (T)q;
    vq += vn * q;
  }
  vq = block_sum(vq, s_red);
  double up = 1.0;
  if (a.cmu + a.c1 > 0.0) {  // natural gradient, _vdcma.py:444-458
    double ria = 0.0, via = 0.0;
    for (int n = tid; n < N; n += blockDim.x) {
      const double vn = (double)a.vn[n], vnn = vn * vn, avec = 2.0 - (bsca + 2.0 * alpha * alpha) * vnn;
      const double r = (double)pv[n] - alpha / (1.0 + nv2) * ((2.0 + nv2) * (double)qv[n] * vn - nv2 * vq * vnn);
      sv[n] = (T)r;  // r for now
      ria += r * (vnn / avec);
      via += vnn * (vnn / avec);
    }
    ria = block_sum(ria, s_red);
    via = block_sum(via, s_red);
    double svnn = 0.0;
    for (int n = tid; n < N; n += blockDim.x) {
      const double vn = (double)a.vn[n], vnn = vn * vn, avec = 2.0 - (bsca + 2.0 * alpha * alpha) * vnn;
      const double s = (double)sv[n] / avec - bsca * ria / (1.0 + bsca * via) * (vnn / avec);
      sv[n] = (T)s;
      svnn += s * vnn;
    }
    svnn = block_sum(svnn, s_red);
    double g2 = 0.0, dmin = 1.0 / 0.0;
    for (int n = tid; n < N; n += blockDim.x) {
      const double vn = (double)a.vn[n], s = (double)sv[n];
      const double gv = (double)qv[n] / nv - alpha / nv * ((2.0 + nv2) * (vn * s) - svnn * vn);
      const double gd = (double)a.dvec[n] * s;
      ngv[n] = (T)gv;
      ngd[n] = (T)gd;
      g2 += gv * gv;
      dmin = fmin(dmin, (double)a.dvec[n] / fabs(gd));
    }
    g2 = block_sum(g2, s_red);
    for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = dmin;
    __syncthreads();
    dmin = s_red[0];
    for (int w = 1; w < 8; ++w) dmin = fmin(dmin, s_red[w]);
    __syncthreads();
    up = fmin(1.0, 0.7 * nv / sqrt(g2));  // at most 70 % change, _vdcma.py:361-363
    up = fmin(up, 0.7 * dmin);
    for (int n = tid; n < N; n += blockDim.x) {
      a.vvec[n] = add_rn(a.vvec[n], mul_rn((T)up, ngv[n]));
      a.dvec[n] = add_rn(a.dvec[n], mul_rn((T)up, ngd[n]));
    }
  }
  if (tid == 0) {
    const double best = (double)a.arfit[s_best];
    c->base.gbest_row = s_best;
    c->base.gfit = best;
    a.besthist[a.it - 1] = (T)best;
    c->hsig = hsig ? 1 : 0;
    c->nfev += a.P;
    c->sigma = sigma;
    c->inject = 1;
    c->aux[2] = up;
  }
  __syncthreads();
  // diagC still describes the population just evaluated (_vdcma.py:380-396: no B, D)
  converge_ladder<T>(c, a.it, N, a.maxiter, a.ilim, a.P, a.xmean, a.xold, a.besthist, a.arfit, a.pc, a.diagC, 1,
                     (const T*)nullptr, (const T*)nullptr, a.xtol, a.ftol, a.insigma, s_red);
}

template <typename T>
__global__ void __launch_bounds__(256)
vd_penalty_state_kernel(const VdPtrs<T> a) {
  __shared__ double s_red[8];
  if (!es_running(a.ctrl)) return;
  penalize_state<T>(a.ctrl, a.it, a.N, a.P, a.hist_cap, a.mueff, a.sorted(), a.xmean, a.xold, a.diagC, 1,
                    a.bnd_weights, a.dfithist, a.coef(), s_red);
}

template <typename T>
static VdPtrs<T> vd_ptrs(const sp_vd_state* st, int it, int evaluate) {
  VdPtrs<T> a;
  a.xmean = (T*)st->xmean;
  a.xold = (T*)st->xold;
  a.dx = (T*)st->dx;
  a.pc = (T*)st->pc;
  a.dvec = (T*)st->dvec;
  a.vvec = (T*)st->vvec;
  a.vn = (T*)st->vn;
  a.diagC = (T*)st->diagC;
  a.dy = (T*)st->dy;
  a.ginj = (T*)st->ginj;
  a.arx = (T*)st->arx;
  a.ary = (T*)st->ary;
  a.yvn = (T*)st->yvn;
  a.arfit = (T*)st->arfit;
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
  a.objective = st->objective;
  a.it = it;
  a.host_z = st->host_z;
  a.evaluate = evaluate;
  a.P = st->P;
  a.ld = st->ld;
  a.cc = st->cc;
  a.c1 = st->c1;
  a.cmu = st->cmu;
  a.mueff = st->mueff;
  a.wsum = st->wsum;
  a.xtol = st->xtol;
  a.ftol = st->ftol;
  a.insigma = st->insigma;
  a.seed = st->seed;
  return a;
}

template <typename T>
static int vd_sample(const sp_vd_state* st, int it, int evaluate, cudaStream_t s) {
  const VdPtrs<T> a = vd_ptrs<T>(st, it, evaluate);
  Shape sh;
  if (!pick_shape(st->N, Num<T>::VEC, &sh)) {
    set_error("sp_vd_sample: ndim %d exceeds the compiled row shapes", st->N);
    return SP_ERR_SHAPE;
  }
  vd_inject_kernel<T><<<1, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  const int grid = grid_for_rows(st->P, sh.lpr, sh.ch >= 4 ? 2 : 4);
#define SP_CALL(TT, C, L) vd_sample_eval_kernel<TT, C, L><<<grid, kThreads, 0, s>>>(a)
  SP_DISPATCH_SHAPE(T, sh, SP_CALL);
#undef SP_CALL
  SP_CHECK_LAUNCH();
  return SP_OK;
}

template <typename T>
static int vd_update(const sp_vd_state* st, int it, cudaStream_t s) {
  const VdPtrs<T> a = vd_ptrs<T>(st, it, 1);
  const int64_t P = st->P;
  const int N = st->N;
  if (st->constraint == SP_CONS_PENALIZE) {
    if (rank_launch<T>(a.arfit, P, a.rank, nullptr, s) != cudaSuccess) return SP_ERR_CUDA;
    g_launches.fetch_add(2);
    scatter_sorted_kernel<T><<<cdiv(P, 256) < 1024 ? cdiv(P, 256) : 1024, 256, 0, s>>>(a.arfit, a.rank, a.sorted(), P, st->ctrl);
    SP_CHECK_LAUNCH();
    vd_penalty_state_kernel<T><<<1, 256, 0, s>>>(a);
    SP_CHECK_LAUNCH();
    penalty_add_kernel<T><<<cdiv(P, 8) < sm_count() * 8 ? cdiv(P, 8) : sm_count() * 8, 256, 0, s>>>(
        a.arx, a.coef(), a.arfit, P, N, st->ld, st->ctrl);
    SP_CHECK_LAUNCH();
  }
  if (rank_launch<T>(a.arfit, P, a.rank, nullptr, s) != cudaSuccess) return SP_ERR_CUDA;
  g_launches.fetch_add(2);
  vd_wsum_kernel<T><<<dim3(cdiv(N, 256), kVdChunks), 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  vd_wreduce_kernel<T><<<cdiv(4 * N, 256), 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  vd_update_kernel<T><<<1, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  vd_refresh_kernel<T><<<1, 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

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

int64_t sp_vd_work_scalars(int N, int64_t P) { return (int64_t)kVdChunks * 4 * N + 13LL * N + P; }

int sp_vd_refresh(const sp_vd_state* st, void* stream) {
  int rc = vd_check(st, 1);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (st->dtype == SP_F32) vd_refresh_kernel<float><<<1, 256, 0, s>>>(vd_ptrs<float>(st, 1, 0));
  else vd_refresh_kernel<double><<<1, 256, 0, s>>>(vd_ptrs<double>(st, 1, 0));
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_vd_sample(const sp_vd_state* st, int it, int evaluate, void* stream) {
  int rc = vd_check(st, it);
  if (rc) return rc;
  SP_CHECK_ARG(!evaluate || (st->objective >= SP_OBJ_ACKLEY && st->objective <= SP_OBJ_STYBLINSKI_TANG),
               "device objective required to evaluate in the sampling kernel");
  return st->dtype == SP_F32 ? vd_sample<float>(st, it, evaluate, (cudaStream_t)stream)
                             : vd_sample<double>(st, it, evaluate, (cudaStream_t)stream);
}

int sp_vd_update(const sp_vd_state* st, int it, void* stream) {
  int rc = vd_check(st, it);
  if (rc) return rc;
  return st->dtype == SP_F32 ? vd_update<float>(st, it, (cudaStream_t)stream)
                             : vd_update<double>(st, it, (cudaStream_t)stream);
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
