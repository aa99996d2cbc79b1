// VD-CMA: restricted covariance C = D (I + v v^T) D, everything O(N) per individual.
// Reference: stochopy/optimize/vdcma/_vdcma.py:235-409 (generation), :426-458
// (pvec_and_qvec, ngv_ngd), converge from cmaes/_cmaes.py:360-434 without B, D.
//
//   vd_sample_eval   z -> y -> x, (y/d).vn, objective               row tiles, HBM: write 2 rows
//   [Penalize]       rank -> percentiles -> weights -> arfit += penalty
//   rank             chunk sort + merge (rank.cuh)
//   vd_wsum/wreduce  S_x, S_y, P_mu, Q_mu over the mu best           column-parallel, read mu rows of y
//   vd_update        mean, sigma (rank gap of rows 0/1), pc, natural gradient on (v, D), ladder,
//                    then |v|^2, vn, diagC and the injection dy of the next generation   one CTA
//   (vd_inject / vd_refresh stand alone only for host-provided draws and the first generation)
#include <cstdlib>

#include "es_common.cuh"

namespace sp {

constexpr int kVdChunks = 256;

template <typename T>
struct VdPtrs {
  T *xmean, *xold, *dx, *pc, *dvec, *vvec, *vn, *diagC, *dy, *ginj, *arx, *ary, *yvn, *arfit, *weights, *xscale,
      *xshift, *besthist, *work, *bnd_weights, *dfithist;
  int32_t* rank;
  sp_es_ctrl* ctrl;
  int N, mu, maxiter, ilim, hist_cap, constraint, objective, it, host_z, evaluate, chunks, stream_stores;
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
__device__ void vd_refresh_body(const VdPtrs<T>& a, double* s_red) {
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
template <typename T>
__global__ void __launch_bounds__(256)
vd_refresh_kernel(const VdPtrs<T> a) {
  __shared__ double s_red[kRedDoubles];
  vd_refresh_body<T>(a, s_red);
}

// injection of generation `it`, _vdcma.py:243-246 (nv2 = |v|^2 of the current v)
template <typename T>
__device__ void vd_inject_body(const VdPtrs<T>& a, int it, double nv2, double* s_red) {
  constexpr int VEC = Num<T>::VEC;
  double g2 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
    T g;
    if (a.host_z) {
      g = a.ginj[n];
    } else {
      T z[VEC];
      normal_block(philox4x32((uint32_t)(n / VEC), 0u, (uint32_t)it, kVdInject, a.seed), z);
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
template <typename T>
__global__ void __launch_bounds__(256)
vd_inject_kernel(const VdPtrs<T> a) {
  __shared__ double s_red[kRedDoubles];
  sp_es_ctrl* c = a.ctrl;
  if (!es_running(c) || !c->inject) return;
  vd_inject_body<T>(a, a.it, c->aux[0], s_red);
}

// row-local sampling + objective, _vdcma.py:239-277.  One register tile per row (z, then y,
// then x in place); the N-vectors (vn, D, mean, scale, shift) are re-read from L1 as 16-byte
// read-only loads per chunk, so a 1024-wide row costs ~40 registers and 3-4 CTAs fit an SM.
// (y / D) . vn is taken from t = z + fac (z.vn) vn before the multiplication by D (y = D t)
// instead of dividing y by D again (the reference divides, _vdcma.py:428: <= 1 ulp apart).
// FULL: ndim == CH * LPR * VEC == ld (no padding, no bounds predicates); CLIP: Penalize is on.
template <typename T, int CH, int LPR, bool FULL, bool CLIP>
__global__ void __launch_bounds__(kThreads, (CH * (int)sizeof(T) <= 32 ? 3 : 2))
vd_sample_eval_kernel(const VdPtrs<T> a) {
  using TL = Tile<T, CH, LPR>;
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  const sp_es_ctrl* c = a.ctrl;
  if (!es_running(c)) return;
  const int lane = threadIdx.x & 31, l = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t groups = (a.P + TL::RPW - 1) / TL::RPW;
  const int ld = FULL ? TL::COLS : (int)a.ld, N = FULL ? TL::COLS : a.N;
  const T sigma = (T)c->sigma;
  const T fac = (T)(sqrt(1.0 + c->aux[0]) - 1.0);
  const bool inject = c->inject != 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) const_cast<sp_es_ctrl*>(c)->sigma_gen = c->sigma;
  constexpr bool clip = CLIP;
  auto vec = [&](const T* __restrict__ p, int cc, T (&o)[VEC]) {
    const int j0 = TL::col(cc, l, 0);
    if (FULL || j0 < ld) {
      const V t = __ldg(reinterpret_cast<const V*>(p + j0));
      const T* q = reinterpret_cast<const T*>(&t);
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = q[e];
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = T(0);
    }
  };

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
        if (FULL || j0 < N) normal_block(philox4x32((uint32_t)(j0 / VEC), (uint32_t)row, (uint32_t)a.it, kEsZ, a.seed), z);
#pragma unroll
        for (int e = 0; e < VEC; ++e) y.v[cc][e] = (FULL || j0 + e < N) ? z[e] : T(0);
      }
    }
    T zv = 0;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      T vn[VEC];
      vec(a.vn, cc, vn);
#pragma unroll
      for (int e = 0; e < VEC; ++e) zv += y.v[cc][e] * vn[e];
    }
    zv = group_sum<LPR>(zv);
    const bool inj_row = inject && row < 2;
    T yv = 0;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      T vn[VEC], dv[VEC];
      vec(a.vn, cc, vn);
      vec(a.dvec, cc, dv);
      if (inj_row) {  // rows 0 / 1 carry +-dy (_vdcma.py:247-248)
        T dy[VEC];
        vec(a.dy, cc, dy);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          y.v[cc][e] = row == 0 ? dy[e] : -dy[e];
          if (FULL || TL::col(cc, l, e) < N) yv += div_rn(y.v[cc][e], dv[e]) * vn[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const T t = add_rn(y.v[cc][e], mul_rn(fac, mul_rn(zv, vn[e])));
          yv += t * vn[e];
          y.v[cc][e] = mul_rn(dv[e], t);
        }
      }
    }
    yv = group_sum<LPR>(yv);
    if (live) {
      if (a.stream_stores) y.store_cs(a.ary + row * a.ld, l, ld);
      else y.store(a.ary + row * a.ld, l, ld);
      if (l == 0) a.yvn[row] = yv;
    }
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      T xm[VEC];
      vec(a.xmean, cc, xm);
#pragma unroll
      for (int e = 0; e < VEC; ++e) y.v[cc][e] = add_rn(xm[e], mul_rn(sigma, y.v[cc][e]));
    }
    if (live) {
      if (a.stream_stores) y.store_cs(a.arx + row * a.ld, l, ld);
      else y.store(a.arx + row * a.ld, l, ld);
    }
    if (!a.evaluate) continue;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      T sc[VEC], sh[VEC];
      vec(a.xscale, cc, sc);
      vec(a.xshift, cc, sh);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        T v = y.v[cc][e];
        if (clip) v = v < T(-1) ? T(-1) : (v > T(1) ? T(1) : v);
        y.v[cc][e] = add_rn(mul_rn(v, sc[e]), sh[e]);
      }
    }
    const T f = evaluate_tile<T, CH, LPR>(a.objective, y, l, N);
    if (live && l == 0) a.arfit[row] = f;
  }
}

// weighted sums over the mu best; part[chunk][0..3][n] = S_x, S_y, P_mu, Q_mu
// (_vdcma.py:291, 313, 426-441).  A CTA owns 256 x VEC columns and one chunk of rows: the
// selected rows of the chunk are compacted (in row order, so the sums are deterministic) into
// shared memory, then streamed 4 rows at a time with 16-byte loads by two thread groups that
// take alternate batches and are folded in a fixed order.  x is rebuilt from y (x = mean +
// sigma y, the very operations of the sampling kernel), so only y is read.
constexpr int kWsTile = 512, kWsUnroll = 4, kWsThreads = 512;
template <typename T>
__global__ void __launch_bounds__(kWsThreads, 2)
vd_wsum_kernel(const VdPtrs<T> a) {
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  if (!es_running(a.ctrl)) return;
  __shared__ int s_row[kWsTile];
  __shared__ T s_w[kWsTile], s_yv[kWsTile];
  __shared__ int s_cnt[kWsThreads / 32];
  __shared__ T s_acc[4 * VEC][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t = tid & 255, grp = tid >> 8;
  const int j0 = (blockIdx.x * 256 + t) * VEC;
  const bool col_ok = j0 < (int)a.ld;
  const int64_t per = (a.P + gridDim.y - 1) / gridDim.y;
  const int64_t i0 = blockIdx.y * per, i1 = (i0 + per < a.P) ? i0 + per : a.P;
  const double nv2 = a.ctrl->aux[0];
  const T k1 = (T)(nv2 / (1.0 + nv2)), nv2t = (T)nv2;
  const T sigma = (T)a.ctrl->sigma_gen;
  const bool with_mu = a.cmu != 0.0;
  T vn[VEC], inv[VEC], xm[VEC], sx[VEC], sy[VEC], pm[VEC], qm[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const bool ok = col_ok && j0 + e < a.N;
    vn[e] = ok ? a.vn[j0 + e] : T(0);
    inv[e] = ok ? div_rn(T(1), a.dvec[j0 + e]) : T(0);
    xm[e] = ok ? a.xmean[j0 + e] : T(0);
    sx[e] = sy[e] = pm[e] = qm[e] = T(0);
  }
  for (int64_t t0 = i0; t0 < i1; t0 += kWsTile) {
    __syncthreads();
    const int64_t i = t0 + tid;
    const int r = i < i1 ? a.rank[i] : a.mu;
    const bool sel = r < a.mu;
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_cnt[warp] = __popc(m);
    __syncthreads();
    int off = 0, cnt = 0;
#pragma unroll
    for (int w = 0; w < kWsThreads / 32; ++w) {
      off += w < warp ? s_cnt[w] : 0;
      cnt += s_cnt[w];
    }
    if (sel) {
      off += __popc(m & ((1u << lane) - 1u));
      s_row[off] = (int)(i - i0);
      s_w[off] = a.weights[r];
      s_yv[off] = a.yvn[i];
    }
    __syncthreads();
    if (!col_ok) continue;
    const T* __restrict__ ybase = a.ary + i0 * a.ld + j0;
    for (int k = grp * kWsUnroll; k < cnt; k += 2 * kWsUnroll) {
      V yv[kWsUnroll];
#pragma unroll
      for (int u = 0; u < kWsUnroll; ++u)
        if (k + u < cnt) yv[u] = *reinterpret_cast<const V*>(ybase + (int64_t)s_row[k + u] * a.ld);
#pragma unroll
      for (int u = 0; u < kWsUnroll; ++u) {
        if (k + u < cnt) {
          const T w = s_w[k + u], yn = s_yv[k + u];
          const T* yy = reinterpret_cast<const T*>(&yv[u]);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const T y = yy[e];
            sx[e] += w * add_rn(xm[e], mul_rn(sigma, y));
            sy[e] += w * y;
            if (with_mu) {
              const T yd = y * inv[e];
              pm[e] += w * (yd * yd - k1 * (yn * (yd * vn[e])) - T(1));
              qm[e] += w * (yn * yd - (T(0.5) * (yn * yn + T(1) + nv2t)) * vn[e]);
            }
          }
        }
      }
    }
  }
  // fold group 1 into group 0, write the chunk's partial
  __syncthreads();
  if (grp == 1) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      s_acc[e][t] = sx[e];
      s_acc[VEC + e][t] = sy[e];
      s_acc[2 * VEC + e][t] = pm[e];
      s_acc[3 * VEC + e][t] = qm[e];
    }
  }
  __syncthreads();
  if (grp != 0) return;
  T* out = a.part() + (size_t)blockIdx.y * 4 * a.N;
#pragma unroll
  for (int e = 0; e < VEC; ++e)
    if (col_ok && j0 + e < a.N) {
      out[j0 + e] = sx[e] + s_acc[e][t];
      out[a.N + j0 + e] = sy[e] + s_acc[VEC + e][t];
      out[2 * a.N + j0 + e] = pm[e] + s_acc[2 * VEC + e][t];
      out[3 * a.N + j0 + e] = qm[e] + s_acc[3 * VEC + e][t];
    }
}

// chunk partials -> sums[q][n], fixed order: a CTA owns 32 outputs, 8 thread groups take
// every 8th chunk, then the 8 group sums are added in order
template <typename T>
__global__ void __launch_bounds__(256)
vd_wreduce_kernel(const VdPtrs<T> a) {
  if (!es_running(a.ctrl)) return;
  __shared__ T s_p[8][32];
  const int o = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + o;
  T acc = 0;
  if (e < 4 * a.N) {
    const T* p = a.part() + e;
#pragma unroll 8
    for (int c = g; c < a.chunks; c += 8) acc += p[(size_t)c * 4 * a.N];
  }
  s_p[g][o] = acc;
  __syncthreads();
  if (g == 0 && e < 4 * a.N) {
    T tot = s_p[0][o];
#pragma unroll
    for (int k = 1; k < 8; ++k) tot += s_p[k][o];
    a.sums()[e] = tot;
  }
}

// mean, step size, paths, natural gradient, termination: one CTA (_vdcma.py:290-396).
// The N-vectors live in registers (kVdNpt elements per thread, 256 threads: the scalar fp64
// algebra between the reductions is replicated per warp, so few warps) and every dependent step
// is one combined block reduction: ~10 barrier rounds instead of ~30 global round trips.
template <typename T, int kVdNpt>
__global__ void __launch_bounds__(256)
vd_update_kernel(const VdPtrs<T> a) {
  __shared__ double s_red[kRedDoubles];
  __shared__ int s_best;
  sp_es_ctrl* c = a.ctrl;
  if (!es_running(c)) return;
  const int N = a.N, tid = threadIdx.x, nt = blockDim.x;
  const double nv2 = c->aux[0], nv = c->aux[1];
  // ---- everything this generation reads, issued up front ------------------------------------------
  bool ok[kVdNpt];
  T xm[kVdNpt], pc[kVdNpt], vnT[kVdNpt], dv[kVdNpt], vv[kVdNpt], Sx[kVdNpt], Sy[kVdNpt], Pm[kVdNpt], Qm[kVdNpt];
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    const int n = tid + k * nt;
    ok[k] = n < N;
    const int m = ok[k] ? n : 0;
    const T* p = a.sums();
    Sx[k] = p[m], Sy[k] = p[N + m], Pm[k] = p[2 * N + m], Qm[k] = p[3 * N + m];
    xm[k] = a.xmean[m], pc[k] = a.pc[m], vnT[k] = a.vn[m], dv[k] = a.dvec[m], vv[k] = a.vvec[m];
  }
  const int r0 = a.rank[0], r1 = a.P > 1 ? a.rank[1] : 0;
  for (int64_t i = tid; i < a.P; i += nt)
    if (a.rank[i] == 0) s_best = (int)i;
  // sigma from the rank gap of the injected pair, _vdcma.py:299-307
  bool hsig = true;
  double sigma = c->sigma_gen, ps_new = c->vd_ps;
  const bool injected = c->inject != 0;
  if (injected) {
    const double alpha_act = (double)(r1 - r0) / ((double)a.P - 1.0);
    ps_new = c->vd_ps + 0.3 * (alpha_act - c->vd_ps);
    sigma *= exp(ps_new / sqrt((double)N));
    hsig = ps_new < 0.5;
  }
  // mean, _vdcma.py:291; evolution path, :310-315
  const T kpc = (T)sqrt(a.cc * (2.0 - a.cc) * a.mueff);
  T dx[kVdNpt];
  double red2[2] = {0.0, 0.0};  // max vn^2, (pc / D) . vn
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    dx[k] = sub_rn(Sx[k], mul_rn((T)a.wsum, xm[k]));
    T v = mul_rn(pc[k], (T)(1.0 - a.cc));
    if (hsig) v = add_rn(v, mul_rn(kpc, Sy[k]));
    pc[k] = v;
    if (ok[k]) {
      const int n = tid + k * nt;
      a.dx[n] = dx[k];
      a.xold[n] = xm[k];
      a.xmean[n] = add_rn(xm[k], dx[k]);
      a.pc[n] = v;
      const double vn = (double)vnT[k];
      red2[0] = fmax(red2[0], vn * vn);
      red2[1] += (double)div_rn(v, dv[k]) * vn;
    }
  }
  {
    const int op[2] = {RED_MAX, RED_SUM};
    block_reduce<2>(red2, op, s_red);
  }
  const double vmax = red2[0], yv1 = red2[1];
  // alpha and friends, _vdcma.py:318-329
  const double gamma = 1.0 / sqrt(1.0 + nv2);
  double alpha = sqrt(nv2 * nv2 + (1.0 + nv2) / vmax * (2.0 - gamma)) / (2.0 + nv2), beta = 0.0;
  if (alpha < 1.0) beta = (4.0 - (2.0 - gamma) / vmax) / ((1.0 + 2.0 / nv2) * (1.0 + 2.0 / nv2));
  else alpha = 1.0;
  const double bsca = 2.0 * alpha * alpha - beta;
  // rank-one vectors from pc / dvec, then p = cmu p_mu (+ c1 p_1), q likewise
  const double k1 = nv2 / (1.0 + nv2);
  T pv[kVdNpt], qv[kVdNpt];
  double vq = 0.0;
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    const double vn = (double)vnT[k];
    double p = a.cmu == 0.0 ? 0.0 : a.cmu * (double)Pm[k];
    double q = a.cmu == 0.0 ? 0.0 : a.cmu * (double)Qm[k];
    if (hsig && a.c1 != 0.0) {
      const double y1 = (double)div_rn(pc[k], dv[k]);
      p += a.c1 * (y1 * y1 - k1 * (yv1 * y1 * vn) - 1.0);
      q += a.c1 * (yv1 * y1 - (0.5 * (yv1 * yv1 + 1.0 + nv2)) * vn);
    }
    pv[k] = (T)p;
    qv[k] = (T)q;
    if (ok[k]) vq += vn * q;
  }
  double up = 1.0;
  if (a.cmu + a.c1 > 0.0) {  // natural gradient, _vdcma.py:444-458
    vq = block_sum(vq, s_red);
    T sv[kVdNpt];
    double red3[2] = {0.0, 0.0};  // ria, via
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k) {
      const double vn = (double)vnT[k], vnn = vn * vn, avec = 2.0 - (bsca + 2.0 * alpha * alpha) * vnn;
      const double r = (double)pv[k] - alpha / (1.0 + nv2) * ((2.0 + nv2) * (double)qv[k] * vn - nv2 * vq * vnn);
      sv[k] = (T)r;
      if (ok[k]) {
        red3[0] += r * (vnn / avec);
        red3[1] += vnn * (vnn / avec);
      }
    }
    {
      const int op[2] = {RED_SUM, RED_SUM};
      block_reduce<2>(red3, op, s_red);
    }
    const double ria = red3[0], via = red3[1];
    double svnn = 0.0;
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k) {
      const double vn = (double)vnT[k], vnn = vn * vn, avec = 2.0 - (bsca + 2.0 * alpha * alpha) * vnn;
      const double sn = (double)sv[k] / avec - bsca * ria / (1.0 + bsca * via) * (vnn / avec);
      sv[k] = (T)sn;
      if (ok[k]) svnn += sn * vnn;
    }
    svnn = block_sum(svnn, s_red);
    T ngv[kVdNpt], ngd[kVdNpt];
    double red4[2] = {0.0, 1.0 / 0.0};  // |ngv|^2, min D / |ngd|
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k) {
      const double vn = (double)vnT[k], sn = (double)sv[k];
      const double gv = (double)qv[k] / nv - alpha / nv * ((2.0 + nv2) * (vn * sn) - svnn * vn);
      const double gd = (double)dv[k] * sn;
      ngv[k] = (T)gv;
      ngd[k] = (T)gd;
      if (ok[k]) {
        red4[0] += gv * gv;
        red4[1] = fmin(red4[1], (double)dv[k] / fabs(gd));
      }
    }
    {
      const int op[2] = {RED_SUM, RED_MIN};
      block_reduce<2>(red4, op, s_red);
    }
    up = fmin(1.0, 0.7 * nv / sqrt(red4[0]));  // at most 70 % change, _vdcma.py:361-363
    up = fmin(up, 0.7 * red4[1]);
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k)
      if (ok[k]) {
        const int n = tid + k * nt;
        a.vvec[n] = add_rn(vv[k], mul_rn((T)up, ngv[k]));
        a.dvec[n] = add_rn(dv[k], mul_rn((T)up, ngd[k]));
      }
  }
  __syncthreads();  // s_best, and the vectors written above, are visible to the whole CTA
  if (tid == 0) {
    const double best = (double)a.arfit[s_best];
    c->base.gbest_row = s_best;
    c->base.gfit = best;
    a.besthist[a.it - 1] = (T)best;
    c->hsig = hsig ? 1 : 0;
    c->nfev += a.P;
    c->sigma = sigma;
    if (injected) c->vd_ps = ps_new;
    c->inject = 1;
    c->aux[2] = up;
  }
  __syncthreads();
  // diagC still describes the population just evaluated (_vdcma.py:380-396: no B, D)
  converge_ladder<T>(c, a.it, N, a.maxiter, a.ilim, a.P, a.xmean, a.xold, a.besthist, a.arfit, a.pc, a.diagC, 1,
                     (const T*)nullptr, (const T*)nullptr, a.xtol, a.ftol, a.insigma, s_red);
  // next generation's |v|^2, vn, diagC and (in-kernel draws) its injected direction dy
  __syncthreads();
  vd_refresh_body<T>(a, s_red);
  __syncthreads();
  if (!a.host_z && es_running(c)) vd_inject_body<T>(a, a.it + 1, c->aux[0], s_red);
}

template <typename T>
__global__ void __launch_bounds__(256)
vd_penalty_state_kernel(const VdPtrs<T> a) {
  __shared__ double s_red[kRedDoubles];
  if (!es_running(a.ctrl)) return;
  penalize_state<T>(a.ctrl, a.it, a.N, a.P, a.hist_cap, a.mueff, a.sorted(), a.xmean, a.xold, a.diagC, 1,
                    a.bnd_weights, a.dfithist, a.coef(), s_red);
}

// row chunks of the weighted sums: >= 64 rows each, at most kVdChunks
static inline int vd_chunks(int64_t P) {
  int64_t c = P / 64;
  return (int)(c < 1 ? 1 : (c > kVdChunks ? kVdChunks : c));
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
  // y and x of a population larger than the L2 are stored evict-first (st.global.cs); SP_VD_PLAIN_STORES=1
  // keeps normal stores (profiling switch)
  static const bool plain_stores = getenv("SP_VD_PLAIN_STORES") != nullptr;
  a.stream_stores = !plain_stores && 2 * (size_t)st->P * st->ld * sizeof(T) > ((size_t)96 << 20) ? 1 : 0;
  a.chunks = vd_chunks(st->P);
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
  if (st->host_z) {  // in-kernel draws: dy was prepared by the previous generation's vd_update_kernel
    vd_inject_kernel<T><<<1, 256, 0, s>>>(a);
    SP_CHECK_LAUNCH();
  }
  const int grid = grid_for_rows(st->P, sh.lpr, sh.ch * (int)sizeof(T) <= 32 ? 3 : 2);
  // fixed-shape instantiations only for full-warp rows (the large-N case that matters)
  const bool full = sh.lpr == 32 && st->N == sh.ch * 32 * Num<T>::VEC && st->ld == st->N;
  const bool clip = st->constraint == SP_CONS_PENALIZE;
#define SP_CALL(TT, C, L)                                                                         \
  do {                                                                                            \
    if (L == 32 && full) {                                                                        \
      if (clip) vd_sample_eval_kernel<TT, C, (L == 32 ? 32 : 32), true, true><<<grid, kThreads, 0, s>>>(a);  \
      else vd_sample_eval_kernel<TT, C, (L == 32 ? 32 : 32), true, false><<<grid, kThreads, 0, s>>>(a);      \
    } else if (clip) {                                                                            \
      vd_sample_eval_kernel<TT, C, L, false, true><<<grid, kThreads, 0, s>>>(a);                  \
    } else {                                                                                      \
      vd_sample_eval_kernel<TT, C, L, false, false><<<grid, kThreads, 0, s>>>(a);                 \
    }                                                                                             \
  } while (0)
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
    scatter_sorted_kernel<T><<<cdiv(P, 256) < 1024 ? cdiv(P, 256) : 1024, 256, 0, s>>>(a.arfit, a.rank, a.sorted(), P, st->ctrl);
    SP_CHECK_LAUNCH();
    vd_penalty_state_kernel<T><<<1, 256, 0, s>>>(a);
    SP_CHECK_LAUNCH();
    penalty_add_kernel<T><<<cdiv(P, 8) < sm_count() * 8 ? cdiv(P, 8) : sm_count() * 8, 256, 0, s>>>(
        a.arx, a.coef(), a.arfit, P, N, st->ld, st->ctrl);
    SP_CHECK_LAUNCH();
  }
  if (rank_launch<T>(a.arfit, P, a.rank, nullptr, s) != cudaSuccess) return SP_ERR_CUDA;
  vd_wsum_kernel<T><<<dim3(cdiv(st->ld, 256 * Num<T>::VEC), a.chunks), kWsThreads, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  vd_wreduce_kernel<T><<<cdiv(4 * N, 32), 256, 0, s>>>(a);
  SP_CHECK_LAUNCH();
  // also refreshes vn / diagC (and dy) for the next generation
  if (N <= 256) vd_update_kernel<T, 1><<<1, 256, 0, s>>>(a);
  else if (N <= 512) vd_update_kernel<T, 2><<<1, 256, 0, s>>>(a);
  else if (N <= 1024) vd_update_kernel<T, 4><<<1, 256, 0, s>>>(a);
  else vd_update_kernel<T, 8><<<1, 256, 0, s>>>(a);
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
