// (implementation header of vdcma_f32.cu / vdcma_f64.cu; the C entry points are in vdcma.cu)
// VD-CMA: restricted covariance C = D (I + v v^T) D, everything O(N) per individual.
// Reference: stochopy/optimize/vdcma/_vdcma.py:235-409 (generation), :426-458
// (pvec_and_qvec, ngv_ngd), converge from cmaes/_cmaes.py:360-434 without B, D.
//
//   vd_sample_eval   z -> y -> x, (y/d).vn, objective               row tiles, HBM: write 2 rows
//   [Penalize]       rank -> percentiles -> weights -> arfit += penalty
//   rank             chunk sort + merge (rank.cuh)
//   vd_wsum          S_x, S_y, P_mu, Q_mu over the mu best           column-parallel, read mu rows of y
//   vd_update        chunk partials -> sums (all CTAs), then in the last CTA: mean, sigma (rank gap of rows
//                    0/1), pc, natural gradient on (v, D), ladder, |v|^2, vn, diagC, fused sampling
//                    constants and the injection dy of the next generation
//   (vd_inject / vd_refresh stand alone only for host-provided draws and the first generation)
#pragma once
#include <cstdlib>
#include <type_traits>

#include "es_common.cuh"

namespace sp {

constexpr int kVdAuxZgen = 15;  // ctrl->aux slot: generation whose z draws sit in arx (vd_zgen_rows), 0 = none
constexpr int kVdUpOut = 32;   // outputs of the chunk reduction per CTA of vd_update_kernel (3 N / 32 CTAs)

// profiling hook, read back with sp_debug_vd_clocks(): [0..11] SM cycle counter at the stages of the update kernel's
// single-CTA phase (thread 0 of the last CTA); [12..15] %globaltimer (ns) when CTA 0 of the sampling / weighted-sum /
// update kernels passes its griddepcontrol.wait and when the update kernel ends -- the timeline of the last generation
__device__ long long g_vd_clk[16];
#define VD_STAMP(i)                                    \
  do {                                                 \
    if (threadIdx.x == 0) g_vd_clk[i] = clock64();     \
  } while (0)
__device__ __forceinline__ void vd_time_stamp(int i) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_vd_clk[i] = (long long)t;
}

template <typename T>
struct VdPtrs {
  T *xmean, *xold, *dx, *pc, *dvec, *vvec, *vn, *diagC, *dy, *ginj, *arx, *ary, *yvn, *arfit, *weights, *xscale,
      *xshift, *besthist, *work, *bnd_weights, *dfithist;
  int32_t* rank;
  sp_es_ctrl* ctrl;
  int N, mu, maxiter, ilim, hist_cap, constraint, objective, it, host_z, evaluate, chunks, stream_stores, lean;
  int64_t P, ld;
  double cc, c1, cmu, mueff, wsum, xtol, ftol, insigma;
  uint64_t seed;
  __host__ __device__ T* part() const { return work; }                               // kVdChunks * 4 * N
  __host__ __device__ T* coef() const { return work + (size_t)kVdChunks * 4 * N; }    // N
  __host__ __device__ T* tmp() const { return coef() + N; }                           // 8 * N
  __host__ __device__ T* sorted() const { return tmp() + 8 * (size_t)N; }             // P
  __host__ __device__ T* sums() const { return sorted() + P; }                        // 4 * N reduced partials
  // fused per-column constants of the lean sampling kernel (vd_refresh_body): the objective sees
  //   (xmean + sigma D t) xscale + xshift = t * fuse_a + fuse_b,  fuse_a = sigma D xscale, fuse_b = xmean xscale + xshift
  __host__ __device__ T* fuse_a() const { return sums() + 4 * (size_t)N; }            // N + 4
  __host__ __device__ T* fuse_b() const { return fuse_a() + N + 4; }                  // N + 4
  __host__ __device__ T* hpart() const { return fuse_b() + N + 4; }                   // kVdChunks (vd_wsum's scalar partials)
  // per-CTA scan results of vd_update_kernel's phase 1: (min f, max f, row of rank 0 or -1) as doubles,
  // kVdUpMax CTAs; the slices before it hold an even number of scalars past kVdChunks * 4 * N + ... only
  // when N and P are even, so the address is rounded up to 8 bytes
  __host__ __device__ double* fpart() const {
    return reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(hpart() + kVdChunks) + 7) & ~(uintptr_t)7);
  }
  __host__ __device__ unsigned char* rank_ws() const {
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fpart() + 3 * kVdUpMax) + 15) & ~(uintptr_t)15);
  }
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
    const T sc = a.xscale[n];
    a.fuse_a()[n] = ((T)c->sigma * d) * sc;
    a.fuse_b()[n] = a.xmean[n] * sc + a.xshift[n];
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

// row-local sampling + objective, _vdcma.py:239-277.  One register tile per row (z, then the
// un-standardised x in place); the four N-vectors of the inner loop -- vn, D and the fused
// constants fuse_a = sigma D xscale, fuse_b = xmean xscale + xshift that vd_refresh_body prepared --
// are staged once per CTA in shared memory, so an element costs three fused multiply-adds and a product
//   t = z + (fac z.vn) vn,   (y/D).vn += t vn,   y = D t,   x_user = t fuse_a + fuse_b
// ((y/D).vn is taken from t before the multiplication by D instead of dividing y by D again as the
// reference does, _vdcma.py:428; all of it within 2 ulp of numpy's operation order).
// a.lean (device-resident loop: in-kernel draws, device objective, nobody reads arx): only y, (y/D).vn
// and the fitness are written -- x_i = xold + sigma_gen y_i rebuilds a row when one is wanted (vd_wsum
// does, and the front-end for the result); otherwise arx = xmean + sigma y is stored too.
// FULL: ndim == CH * LPR * VEC == ld (no padding, no bounds predicates); CLIP: Penalize is on (the
// objective then sees clip(xmean + sigma y, -1, 1) xscale + xshift, cmaes/_constraints.py:30-32).
// FAST (only with !CLIP): the device-resident loop's configuration fixed at compile time -- in-kernel draws,
// lean, objective evaluated, plain (1) or evict-first (2) stores of y -- so the row body has no run-time
// switches between its chunks; 0: those are read from the state.  The two injected rows take their own
// copy of the row body (a row-level branch), which keeps the common one straight-line.
template <typename T, int CH, int LPR, bool FULL, bool CLIP, int FAST>
__global__ void __launch_bounds__(kThreads, (FAST != 0 && CH * (int)sizeof(T) <= 32 ? 3 : 2))
vd_sample_eval_kernel(const VdPtrs<T> a, const PhiloxKeys keys) {
  using TL = Tile<T, CH, LPR>;
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  constexpr int COLS = TL::COLS;
  static_assert(!(CLIP && FAST != 0), "the fast variants never clip");
  __shared__ __align__(16) T s_vn[COLS], s_dv[COLS], s_fa[COLS], s_fb[COLS];
  const sp_es_ctrl* c = a.ctrl;
  pdl_launch_dependents();
  pdl_wait();  // the previous generation's update kernel wrote everything read below
  if (!es_running(c)) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) vd_time_stamp(12);
  const int lane = threadIdx.x & 31, l = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t groups = (a.P + TL::RPW - 1) / TL::RPW;
  const int ld = FULL ? COLS : (int)a.ld, N = FULL ? COLS : a.N;
  for (int j = threadIdx.x; j < COLS; j += kThreads) {
    const bool ok = FULL || j < N;
    s_vn[j] = ok ? a.vn[j] : T(0);
    s_dv[j] = ok ? a.dvec[j] : T(0);
    s_fa[j] = ok ? a.fuse_a()[j] : T(0);
    s_fb[j] = ok ? a.fuse_b()[j] : T(0);
  }
  const T sigma = (T)c->sigma;
  const T fac = (T)(sqrt(1.0 + c->aux[0]) - 1.0);
  const bool inject = c->inject != 0;
  const bool host_z = FAST ? false : a.host_z != 0;
  const bool store_x = FAST ? false : !a.lean;
  const bool stream = FAST == 2 ? true : (FAST == 1 ? false : a.stream_stores != 0);
  const bool evaluate = FAST ? true : a.evaluate != 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) const_cast<sp_es_ctrl*>(c)->sigma_gen = c->sigma;
  __syncthreads();
  auto svec = [&](const T* p, int cc, T (&o)[VEC]) {  // shared memory: always in bounds (COLS wide)
    const V t = *reinterpret_cast<const V*>(p + TL::col(cc, l, 0));
    const T* q = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int e = 0; e < VEC; ++e) o[e] = q[e];
  };
  auto gvec = [&](const T* __restrict__ p, int cc, T (&o)[VEC]) {
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
  auto put = [&](T* __restrict__ rowp, int cc, const T (&val)[VEC]) {
    const int j0 = TL::col(cc, l, 0);
    if (FULL || j0 < ld) {
      V t;
      T* q = reinterpret_cast<T*>(&t);
#pragma unroll
      for (int e = 0; e < VEC; ++e) q[e] = val[e];
      if (stream) __stcs(reinterpret_cast<V*>(rowp + j0), t);
      else *reinterpret_cast<V*>(rowp + j0) = t;
    }
  };
  // one row (group of RPW rows per warp); INJ: rows 0 / 1 of an injecting generation carry +-dy
  auto row_body = [&](int64_t row, bool live, auto inj_tag) {
    constexpr bool INJ = decltype(inj_tag)::value;
    TL y;
    if (host_z) y.load(a.ary + row * a.ld, l, ld);
    T zv = 0;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      const int j0 = TL::col(cc, l, 0);
      if (!host_z) {
        T z[VEC];
        if (FULL || j0 < N)
          normal_block(philox4x32_keyed<kEsZRounds>((uint32_t)(j0 / VEC), (uint32_t)row, (uint32_t)a.it, kEsZ, keys), z);
#pragma unroll
        for (int e = 0; e < VEC; ++e) y.v[cc][e] = (FULL || j0 + e < N) ? z[e] : T(0);
      }
      T vn[VEC];
      svec(s_vn, cc, vn);
#pragma unroll
      for (int e = 0; e < VEC; ++e) zv += y.v[cc][e] * vn[e];
    }
    zv = group_sum<LPR>(zv);
    const T k = fac * zv;
    T yv = 0;
    T* __restrict__ yrow = a.ary + row * a.ld;
    T* __restrict__ xrow = a.arx + row * a.ld;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      const int j0 = TL::col(cc, l, 0);
      T vn[VEC], dv[VEC], yy[VEC], tt[VEC];
      svec(s_vn, cc, vn);
      svec(s_dv, cc, dv);
      if (INJ) {  // _vdcma.py:247-248; t = y / D
        T dy[VEC];
        gvec(a.dy, cc, dy);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (row < 2) {
            yy[e] = row == 0 ? dy[e] : -dy[e];
            tt[e] = (FULL || j0 + e < N) ? div_rn(yy[e], dv[e]) : T(0);
          } else {  // RPW > 1: the other rows of the group keep the sampling formula (per-lane select, no divergence
                    // around the warp shuffles of the row reductions)
            tt[e] = k * vn[e] + y.v[cc][e];
            yy[e] = dv[e] * tt[e];
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          tt[e] = k * vn[e] + y.v[cc][e];
          yy[e] = dv[e] * tt[e];
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) yv += tt[e] * vn[e];
      if (live) put(yrow, cc, yy);
      if (CLIP || store_x) {  // the standardised x itself is needed
        T xm[VEC], xs[VEC];
        gvec(a.xmean, cc, xm);
#pragma unroll
        for (int e = 0; e < VEC; ++e) xs[e] = add_rn(xm[e], mul_rn(sigma, yy[e]));
        if (live && store_x) put(xrow, cc, xs);
        if (CLIP) {
          T sc[VEC], sh[VEC];
          gvec(a.xscale, cc, sc);
          gvec(a.xshift, cc, sh);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const T v = xs[e] < T(-1) ? T(-1) : (xs[e] > T(1) ? T(1) : xs[e]);
            y.v[cc][e] = add_rn(mul_rn(v, sc[e]), sh[e]);
          }
        }
      }
      if (!CLIP) {
        T fa[VEC], fb[VEC];
        svec(s_fa, cc, fa);
        svec(s_fb, cc, fb);
#pragma unroll
        for (int e = 0; e < VEC; ++e) y.v[cc][e] = tt[e] * fa[e] + fb[e];
      }
    }
    yv = group_sum<LPR>(yv);
    if (live && l == 0) a.yvn[row] = yv;
    if (!evaluate) return;
    const T f = evaluate_tile<T, CH, LPR>(a.objective, y, l, N);
    if (live && l == 0) a.arfit[row] = f;
  };

  for (int64_t g = warp; g < groups; g += nwarps) {
    int64_t row = g * TL::RPW + sub;
    const bool live = row < a.P;
    if (!live) row = a.P - 1;
    // rows 0 and 1 share a warp only when RPW > 1; then the whole group takes the injecting copy, whose
    // per-row test keeps the other rows on the sampling formula
    if (inject && g * TL::RPW < 2) {  // the group holds row 0 and / or row 1 (RPW == 1: two groups); warp-uniform
      row_body(row, live, std::true_type{});
    } else {
      row_body(row, live, std::false_type{});
    }
  }
}

// Wide rows (CH >= 4, i.e. more than 16 scalars per lane) of the device-resident loop (in-kernel draws, lean,
// objective on the device, no Penalize): the same row algorithm as above with the row tile kept in a
// WARP-PRIVATE SHARED-MEMORY ROW instead of registers.  Every lane only ever re-reads the 16-byte vectors it
// wrote itself, so no synchronisation is needed -- the shared row is an explicitly managed spill area.  The
// register-tile version holds CH * VEC scalars per lane across both passes: at CH = 8 ptxas spills ~100 MB per
// launch to local memory under the 80-register cap (l1tex local sectors in profiles/r02_vd_sample_l2_metrics.csv)
// and its fully unrolled row body is 20 KB of SASS per variant (13.5 k instructions in the kernel; 18 % of the
// stall samples were instruction fetches).  Here the column loops are real loops (unrolled by 2), the live state
// between passes is a handful of scalars, and only the objective sees a register tile (loaded from the shared row
// at the end).  Rows 0 / 1 of an injecting generation branch warp-uniformly inside pass 2.
template <typename T, int CH, bool FULL>
__global__ void __launch_bounds__(kThreads, (CH * (int)sizeof(T) <= 32 ? 4 : 2))
vd_sample_smem_kernel(const VdPtrs<T> a, const PhiloxKeys keys) {
  using TL = Tile<T, CH, 32>;
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  constexpr int COLS = TL::COLS;
  extern __shared__ __align__(16) unsigned char vd_smem[];
  T* s_vn = reinterpret_cast<T*>(vd_smem);
  T* s_dv = s_vn + COLS;
  T* s_fa = s_dv + COLS;
  T* s_fb = s_fa + COLS;
  T* srow = s_fb + COLS + (size_t)(threadIdx.x >> 5) * COLS;
  const sp_es_ctrl* c = a.ctrl;
  pdl_launch_dependents();
  pdl_wait();  // the previous generation's update kernel wrote everything read below
  if (!es_running(c)) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) vd_time_stamp(12);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int ld = FULL ? COLS : (int)a.ld, N = FULL ? COLS : a.N;
  const int64_t ldr = FULL ? (int64_t)COLS : a.ld;
  for (int j = threadIdx.x; j < COLS; j += kThreads) {
    const bool ok = FULL || j < N;
    s_vn[j] = ok ? a.vn[j] : T(0);
    s_dv[j] = ok ? a.dvec[j] : T(0);
    s_fa[j] = ok ? a.fuse_a()[j] : T(0);
    s_fb[j] = ok ? a.fuse_b()[j] : T(0);
  }
  const T fac = (T)(sqrt(1.0 + c->aux[0]) - 1.0);
  const bool inject = c->inject != 0;
  const bool stream = a.stream_stores != 0;
  const uint32_t it = (uint32_t)a.it;
  // z of this generation may already sit in the (otherwise unused, lean) arx buffer: the previous generation's
  // update kernel draws it on the SMs its single-CTA phase leaves idle (vd_zgen_rows) and says so in aux[15]
  const bool zpre = c->aux[kVdAuxZgen] == (double)a.it;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const_cast<sp_es_ctrl*>(c)->sigma_gen = c->sigma;
    const_cast<sp_es_ctrl*>(c)->pad_ = 0;  // row-claim counter of the next vd_zgen_rows
  }
  __syncthreads();
  auto lds = [&](const T* p, int j0, T (&o)[VEC]) {
    const V t = *reinterpret_cast<const V*>(p + j0);
    const T* q = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int e = 0; e < VEC; ++e) o[e] = q[e];
  };
  auto sts = [&](T* p, int j0, const T (&val)[VEC]) {
    V t;
    T* q = reinterpret_cast<T*>(&t);
#pragma unroll
    for (int e = 0; e < VEC; ++e) q[e] = val[e];
    *reinterpret_cast<V*>(p + j0) = t;
  };
  for (int64_t row = warp; row < a.P; row += nwarps) {
    // pass 1: z -> shared row, z . vn
    T zv = 0;
#pragma unroll 2
    for (int cc = 0; cc < CH; ++cc) {
      const int j0 = TL::col(cc, lane, 0);
      T z[VEC], vn[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) z[e] = T(0);
      if (zpre) {
        if (FULL || j0 < ld) {
          const V t = __ldcs(reinterpret_cast<const V*>(a.arx + row * ldr + j0));  // read once
          const T* q = reinterpret_cast<const T*>(&t);
#pragma unroll
          for (int e = 0; e < VEC; ++e) z[e] = q[e];
        }
      } else if (FULL || j0 < N) {
        normal_block(philox4x32_keyed<kEsZRounds>((uint32_t)(j0 / VEC), (uint32_t)row, it, kEsZ, keys), z);
        if (!FULL) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) z[e] = (j0 + e < N) ? z[e] : T(0);
        }
      }
      lds(s_vn, j0, vn);
#pragma unroll
      for (int e = 0; e < VEC; ++e) zv += z[e] * vn[e];
      sts(srow, j0, z);
    }
    zv = group_sum<32>(zv);
    const T k = fac * zv;
    const bool inj = inject && row < 2;  // warp-uniform: the pair +-dy of _vdcma.py:247-248
    T yv = 0;
    T* __restrict__ yrow = a.ary + row * ldr;
#pragma unroll 2
    for (int cc = 0; cc < CH; ++cc) {
      const int j0 = TL::col(cc, lane, 0);
      T z[VEC], vn[VEC], dv[VEC], yy[VEC], tt[VEC], fa[VEC], fb[VEC];
      lds(srow, j0, z);
      lds(s_vn, j0, vn);
      lds(s_dv, j0, dv);
      if (inj) {  // t = y / D
        if (FULL || j0 < ld) {
          const V t = __ldg(reinterpret_cast<const V*>(a.dy + j0));
          const T* q = reinterpret_cast<const T*>(&t);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            yy[e] = row == 0 ? q[e] : -q[e];
            tt[e] = (FULL || j0 + e < N) ? div_rn(yy[e], dv[e]) : T(0);
          }
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) yy[e] = tt[e] = T(0);
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          tt[e] = k * vn[e] + z[e];
          yy[e] = dv[e] * tt[e];
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) yv += tt[e] * vn[e];
      if (FULL || j0 < ld) {
        V t;
        T* q = reinterpret_cast<T*>(&t);
#pragma unroll
        for (int e = 0; e < VEC; ++e) q[e] = yy[e];
        if (stream) __stcs(reinterpret_cast<V*>(yrow + j0), t);
        else *reinterpret_cast<V*>(yrow + j0) = t;
      }
      lds(s_fa, j0, fa);
      lds(s_fb, j0, fb);
#pragma unroll
      for (int e = 0; e < VEC; ++e) z[e] = tt[e] * fa[e] + fb[e];  // what the objective sees
      sts(srow, j0, z);
    }
    yv = group_sum<32>(yv);
    if (lane == 0) a.yvn[row] = yv;
    // pass 3: the objective on a register tile loaded from the shared row
    TL x;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) lds(srow, TL::col(cc, lane, 0), x.v[cc]);
    const T f = evaluate_tile<T, CH, 32>(a.objective, x, lane, N);
    if (lane == 0) a.arfit[row] = f;
  }
}
template <int CH, typename T>
constexpr size_t vd_smem_bytes() { return (size_t)(4 + kThreads / 32) * Tile<T, CH, 32>::COLS * sizeof(T); }

// weighted sums over the mu best (_vdcma.py:291, 313, 426-441), factored so that a row costs four
// instructions per element.  With yd = y / D, yn = yd . vn and h = (yn^2 + 1 + |v|^2) / 2 the reference needs
//   S_y  = sum w y                                       (evolution path; and dx = sum w x - (sum w) xmean = sigma S_y)
//   P_mu = sum w (yd^2 - k1 yn vn yd - 1) = A / D^2 - k1 vn B / D - sum w
//   Q_mu = sum w (yn yd - h vn)           = B / D - vn H
// where  A = sum w y^2,  B = sum (w yn) y  are column sums and  H = sum w h  is one scalar.
// part[chunk][0..2][n] = S_y, A, B of the chunk's selected rows, hpart[chunk] = its share of H.
// A CTA owns 256 x VEC columns and one chunk of rows: the selected rows of the chunk are compacted
// (in row order, so the sums are deterministic) into shared memory, then streamed kWsUnroll rows at a
// time with 16-byte loads by two thread groups that take alternate batches and are folded in a fixed
// order.  Only y is read: x is never needed.
constexpr int kWsTile = 512, kWsUnroll = 8, kWsThreads = 512;
template <typename T>
__global__ void __launch_bounds__(kWsThreads, 2)
vd_wsum_kernel(const VdPtrs<T> a) {
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  pdl_launch_dependents();
  pdl_wait();
  if (!es_running(a.ctrl)) return;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) vd_time_stamp(13);
  __shared__ int s_row[kWsTile];
  __shared__ T s_w[kWsTile], s_wyn[kWsTile];
  __shared__ int s_cnt[kWsThreads / 32];
  __shared__ T s_h[kWsThreads / 32];
  __shared__ T s_acc[3 * VEC][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t = tid & 255, grp = tid >> 8;
  const int j0 = (blockIdx.x * 256 + t) * VEC;
  const bool col_ok = j0 < (int)a.ld;
  const int64_t per = (a.P + gridDim.y - 1) / gridDim.y;
  const int64_t i0 = blockIdx.y * per, i1 = (i0 + per < a.P) ? i0 + per : a.P;
  const T nv2t = (T)a.ctrl->aux[0];
  T sy[VEC], sa[VEC], sb[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) sy[e] = sa[e] = sb[e] = T(0);
  T hsum = 0;
  for (int64_t t0 = i0; t0 < i1; t0 += kWsTile) {
    __syncthreads();
    const int64_t i = t0 + tid;
    const int r = i < i1 ? a.rank[i] : a.mu;
    const T yn = i < i1 ? a.yvn[i] : T(0);  // in flight together with the rank
    const bool sel = r < a.mu;
    const T w = sel ? a.weights[r] : T(0);
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    T h = w * (T(0.5) * (yn * yn + T(1) + nv2t));  // 0 for the rows not selected
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) {
      s_cnt[warp] = __popc(m);
      s_h[warp] = h;
    }
    __syncthreads();
    int off = 0, cnt = 0;
#pragma unroll
    for (int q = 0; q < kWsThreads / 32; ++q) {
      off += q < warp ? s_cnt[q] : 0;
      cnt += s_cnt[q];
      hsum += s_h[q];  // same order in every thread
    }
    if (sel) {
      off += __popc(m & ((1u << lane) - 1u));
      s_row[off] = (int)(i - i0);
      s_w[off] = w;
      s_wyn[off] = w * yn;
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
          const T w = s_w[k + u], wyn = s_wyn[k + u];
          const T* yy = reinterpret_cast<const T*>(&yv[u]);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const T y = yy[e], wy = w * y;
            sy[e] += wy;
            sa[e] += wy * y;
            sb[e] += wyn * y;
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
      s_acc[e][t] = sy[e];
      s_acc[VEC + e][t] = sa[e];
      s_acc[2 * VEC + e][t] = sb[e];
    }
  }
  __syncthreads();
  if (grp != 0) return;
  if (blockIdx.x == 0 && tid == 0) a.hpart()[blockIdx.y] = hsum;
  T* out = a.part() + (size_t)blockIdx.y * 3 * a.N;
#pragma unroll
  for (int e = 0; e < VEC; ++e)
    if (col_ok && j0 + e < a.N) {
      out[j0 + e] = sy[e] + s_acc[e][t];
      out[a.N + j0 + e] = sa[e] + s_acc[VEC + e][t];
      out[2 * a.N + j0 + e] = sb[e] + s_acc[2 * VEC + e][t];
    }
}

// ---- z of the NEXT generation, drawn while the update kernel's single-CTA phase runs (experiment, off by default:
// measured slower, see vd_update) ---------------------------------------------------------------------------------
// The N(0,1) draws depend only on (seed, row, generation): nothing of the update is needed for them, and they are 45 %
// of the sampling kernel's instructions.  The update kernel therefore runs with one CTA per SM; the CTAs that are not
// (or no longer) needed by the update claim rows from a device counter (ctrl->pad_) and store z of generation it + 1
// into arx -- unused in the lean device-resident loop -- and the sampling kernel of it + 1 reads it (aux[15] holds the
// generation the buffer is valid for).  Same Philox counters as the in-kernel draws: bitwise the same run.
constexpr int kVdZgenBatch = 4;  // rows per claim
template <typename T>
__device__ __forceinline__ void vd_zgen_rows(const VdPtrs<T>& a, const PhiloxKeys& keys) {
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  const int lane = threadIdx.x & 31;
  const uint32_t it = (uint32_t)(a.it + 1);
  int32_t* counter = &a.ctrl->pad_;
  for (;;) {
    int r0 = 0;
    if (lane == 0) r0 = atomicAdd(counter, kVdZgenBatch);
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    if (r0 >= a.P) break;
#pragma unroll 1
    for (int r = r0; r < r0 + kVdZgenBatch && r < a.P; ++r) {
      T* __restrict__ zrow = a.arx + (int64_t)r * a.ld;
#pragma unroll 2
      for (int j0 = lane * VEC; j0 < (int)a.ld; j0 += 32 * VEC) {
        T z[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) z[e] = T(0);
        if (j0 < a.N) {
          normal_block(philox4x32_keyed<kEsZRounds>((uint32_t)(j0 / VEC), (uint32_t)r, it, kEsZ, keys), z);
#pragma unroll
          for (int e = 0; e < VEC; ++e) z[e] = (j0 + e < a.N) ? z[e] : T(0);
        }
        V t;
        T* q = reinterpret_cast<T*>(&t);
#pragma unroll
        for (int e = 0; e < VEC; ++e) q[e] = z[e];
        *reinterpret_cast<V*>(zrow + j0) = t;
      }
    }
  }
}

// Block reduction with ONE barrier: warp butterflies, one shared slot per (value, warp), then every thread
// folds the NW warp results itself, in warp order (deterministic).  Two slot sets alternate (`ph`), so a
// thread may start the next reduction while others still read this one's slots.
template <int K, int NW>
__device__ __forceinline__ void reduce1(double (&v)[K], const int (&op)[K], double (*buf)[kRedMax][NW], int& ph) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = red_warp(v[k], op[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) buf[ph][k][warp] = v[k];
  }
  __syncthreads();
  if constexpr (NW <= 8) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = buf[ph][k][0];
#pragma unroll
      for (int w = 1; w < NW; ++w) x = red_apply(x, buf[ph][k][w], op[k]);
      v[k] = x;
    }
  } else {
    // many warps: warp k folds value k with one more butterfly and leaves it in slot 0 (second barrier); a full
    // second butterfly of all K values in every warp costs 5 x K fp64 shuffle steps per warp again (measured:
    // 5.9 us for K = 15 with 32 warps)
    static_assert(K <= NW, "one warp per value");
    if (warp < K) {
      double x = red_identity(RED_SUM);
      int myop = RED_SUM;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k == warp) {
          myop = op[k];
          x = lane < NW ? buf[ph][k][lane < NW ? lane : 0] : red_identity(op[k]);
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x = red_apply(x, __shfl_xor_sync(0xffffffffu, x, o), myop);
      if (lane == 0) buf[ph][warp][0] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = buf[ph][k][0];
  }
  ph ^= 1;
}

// mean, step size, paths, natural gradient, termination (_vdcma.py:290-396), fused with the
// reduction of the chunk partials of vd_wsum.
//   phase 1 (every CTA of the grid): sums[q][n] = sum over the chunks of part[chunk][q][n] in a fixed
//     order -- a CTA owns kUpOut outputs, 8 thread groups take every 8th chunk with all their loads in
//     flight at once, then the 8 group sums are added in order.  Every CTA also scans a slice of the
//     population for the row of rank 0 and the fitness range the ladder needs.  The LAST CTA to finish
//     goes on;
//   phase 2 (that one CTA): the N-vectors live in registers (kVdNpt columns per thread; the scalar
//     fp64 algebra between the reductions is replicated per warp, so few warps) and every dependent
//     step is one combined block reduction.
// UT threads per CTA (template parameter): one column per thread up to N = 1024 -- the single-CTA phase is a chain
// of dependent fp64 operations (measured: ~14 cycles per instruction per warp with 8 warps; 26 us at N = 1024 with
// four columns per thread), so it wants as many warps as there are columns, not registers per thread.
// n1: CTAs of phase 1 (the first n1 of the grid); zgen != 0: the grid has one CTA per SM and every CTA that is not
// (or no longer) busy with the update draws z of the next generation (vd_zgen_rows).
template <typename T, int kVdNpt, int kUpThreads>
__global__ void __launch_bounds__(kUpThreads)
vd_update_kernel(const VdPtrs<T> a, const int n1, const int zgen, const PhiloxKeys keys) {
  constexpr int kUpOut = kVdUpOut, kGroups = kUpThreads / kUpOut;
  __shared__ double s_red[kRedDoubles];
  __shared__ T s_p[kGroups][kUpOut];
  __shared__ bool s_last;
  pdl_launch_dependents();
  pdl_wait();
  sp_es_ctrl* c = a.ctrl;
  if (!es_running(c)) return;
  const int N = a.N, tid = threadIdx.x, nt = blockDim.x;
  const bool draw_next = zgen != 0 && a.it < a.maxiter;
  if ((int)blockIdx.x >= n1) {  // not part of the update: straight to the draws
    if (draw_next) vd_zgen_rows<T>(a, keys);
    return;
  }
  const long long clk0 = clock64();
  if (blockIdx.x == 0 && tid == 0) vd_time_stamp(14);
  // the N-vectors this generation did not touch yet: loaded by every CTA before the grid-wide hand-over, so the
  // single-CTA phase does not start with a round of dependent L2 misses (only `sums` has to wait)
  bool ok[kVdNpt];
  T xm[kVdNpt], pc[kVdNpt], vnT[kVdNpt], dv[kVdNpt], vv[kVdNpt], dC[kVdNpt];
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    const int n = tid + k * nt;
    ok[k] = n < N;
    const int m = ok[k] ? n : 0;
    xm[k] = a.xmean[m], pc[k] = a.pc[m], vnT[k] = a.vn[m], dv[k] = a.dvec[m], vv[k] = a.vvec[m], dC[k] = a.diagC[m];
  }
  {
    const int o = tid % kUpOut, g = tid / kUpOut;
    const int e = blockIdx.x * kUpOut + o;
    T acc = 0;
    if (e < 3 * N) {
      const T* p = a.part() + e + (size_t)g * 3 * N;
      const size_t step = (size_t)kGroups * 3 * N;
      T v[kVdChunks / kGroups];
#pragma unroll
      for (int k = 0; k < kVdChunks / kGroups; ++k) v[k] = (g + kGroups * k < a.chunks) ? __ldcg(p + k * step) : T(0);
#pragma unroll
      for (int k = 0; k < kVdChunks / kGroups; ++k) acc += v[k];
    }
    // this CTA's slice of the population: row of rank 0 (ties by index: the stable rank's first minimum)
    // and min / max fitness
    const int64_t per = (a.P + n1 - 1) / n1;
    const int64_t i0 = blockIdx.x * per, i1 = (i0 + per < a.P) ? i0 + per : a.P;
    double ext[3] = {1.0 / 0.0, -1.0 / 0.0, -1.0};  // min f, max f, best row (or -1)
    for (int64_t i = i0 + tid; i < i1; i += nt) {
      const int rk = a.rank[i];
      const double f = (double)a.arfit[i];
      ext[0] = fmin(ext[0], f);
      ext[1] = fmax(ext[1], f);
      if (rk == 0) ext[2] = (double)i;
    }
    s_p[g][o] = acc;
    {
      const int op[3] = {RED_MIN, RED_MAX, RED_MAX};
      block_reduce<3>(ext, op, s_red);  // (barriers inside also publish s_p)
    }
    if (g == 0 && e < 3 * N) {
      T tot = s_p[0][o];
#pragma unroll
      for (int k = 1; k < kGroups; ++k) tot += s_p[k][o];
      a.sums()[e] = tot;
    }
    if (tid < 3) a.fpart()[3 * blockIdx.x + tid] = ext[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(&c->base.done_blocks, 1u);
      s_last = prev == (unsigned)n1 - 1u;
      if (s_last) c->base.done_blocks = 0;  // ready for the next launch
    }
    __syncthreads();
    if (!s_last) {
      if (draw_next) vd_zgen_rows<T>(a, keys);
      return;
    }
    __threadfence();
  }
  if (tid == 0) g_vd_clk[0] = clk0;
  VD_STAMP(1);
  // ---- phase 2 -----------------------------------------------------------------------------------
  // Six dependent block reductions, one barrier each (reduce1), instead of a dozen three-barrier ones:
  //   L1 max vn^2, (pc/D).vn, H and everything the termination ladder needs   L2 vn.q   L3 ria, via
  //   L4 s.vn^2   L5 |ngv|^2, min D/|ngd|   L6 |v'|^2 and the three sums of the next injection
  __shared__ double s_r1[2][kRedMax][kUpThreads / 32];
  int ph = 0;
  const double nv2 = c->aux[0], nv = c->aux[1];
  T Sy[kVdNpt], Sa[kVdNpt], Sb[kVdNpt];
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    const int m = ok[k] ? tid + k * nt : 0;
    const T* p = a.sums();
    Sy[k] = __ldcg(p + m), Sa[k] = __ldcg(p + N + m), Sb[k] = __ldcg(p + 2 * N + m);
  }
  const int r0 = a.rank[0], r1 = a.P > 1 ? a.rank[1] : 0;
  const double inf = 1.0 / 0.0;
  // [0] max vn^2  [1] (pc/D).vn  [2] H  [3] |xold - xmean|^2  [4] max sd  [5] any 0.2 sigma sd < 1e-10
  // [6] any sigma sd > 1e3 sigma0  [7] all sigma |pc| < 1e-11 sigma0  [8..11] min / max of the zero-padded
  // best-fitness history (all of it, and the window it-ilim..it) WITHOUT this generation's entry
  // [12] min f  [13] max f  [14] row of rank 0
  double L1[15] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, inf, -inf, inf, -inf, inf, -inf, -1.0};
  if (tid < a.chunks) L1[2] = (double)__ldcg(a.hpart() + tid);
  for (int b = tid; b < n1; b += nt) {  // the per-CTA scan results of phase 1
    const double* fp = a.fpart() + 3 * b;
    L1[12] = fmin(L1[12], __ldcg(fp));
    L1[13] = fmax(L1[13], __ldcg(fp + 1));
    L1[14] = fmax(L1[14], __ldcg(fp + 2));
  }
  for (int i = tid; i < a.maxiter; i += nt) {  // _cmaes.py:412-414 (window incl. one not-yet-written zero), :424-427
    if (i == a.it - 1) continue;
    const double h = (double)a.besthist[i];
    L1[8] = fmin(L1[8], h);
    L1[9] = fmax(L1[9], h);
    if (i >= a.it - a.ilim && i <= a.it) {
      L1[10] = fmin(L1[10], h);
      L1[11] = fmax(L1[11], h);
    }
  }
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
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    dx[k] = mul_rn((T)c->sigma_gen, Sy[k]);  // = w . x[top mu] - (sum w) xmean (_vdcma.py:291), without the cancellation
    T v = mul_rn(pc[k], (T)(1.0 - a.cc));
    if (hsig) v = add_rn(v, mul_rn(kpc, Sy[k]));
    pc[k] = v;
    if (ok[k]) {
      const int n = tid + k * nt;
      const T xnew = add_rn(xm[k], dx[k]);
      a.dx[n] = dx[k];
      a.xold[n] = xm[k];
      a.xmean[n] = xnew;
      a.pc[n] = v;
      const double vn = (double)vnT[k];
      L1[0] = fmax(L1[0], vn * vn);
      L1[1] += (double)div_rn(v, dv[k]) * vn;
      const double d = (double)xm[k] - (double)xnew, sd = sqrt((double)dC[k]);
      L1[3] += d * d;
      L1[4] = fmax(L1[4], sd);
      if (0.2 * sigma * sd < 1.0e-10) L1[5] = 1.0;
      if (sigma * sd > 1.0e3 * a.insigma) L1[6] = 1.0;
      if (!(sigma * fabs((double)v) < 1.0e-11 * a.insigma)) L1[7] = 0.0;
      xm[k] = xnew;
    }
  }
  {
    const int op[15] = {RED_MAX, RED_SUM, RED_SUM, RED_SUM, RED_MAX, RED_MAX, RED_MAX, RED_MIN, RED_MIN, RED_MAX,
                        RED_MIN, RED_MAX, RED_MIN, RED_MAX, RED_MAX};
    VD_STAMP(2);
    reduce1<15, kUpThreads / 32>(L1, op, s_r1, ph);
    VD_STAMP(3);
  }
  const double vmax = L1[0], yv1 = L1[1], hmu = L1[2];
  const double best = L1[12];  // the row of rank 0 carries the minimum
  const int best_row = (int)L1[14];  // -1 would mean the ranking holds no rank 0: reported as SP_STATUS_INTERNAL below
  // termination ladder (_cmaes.py:360-434 without B, D: _vdcma.py:380-396); diagC still describes the
  // population just evaluated
  int status = SP_RUNNING;
  {
    const double hmin = fmin(L1[8], best), hmax = fmax(L1[9], best), wmin = fmin(L1[10], best), wmax = fmax(L1[11], best);
    if (a.it >= a.maxiter) status = -1;
    else if (sqrt(L1[3]) <= a.xtol && best < a.ftol) status = 0;
    else if (best <= a.ftol) status = 1;
    else if (L1[5] > 0.5) status = -3;
    else if (a.it >= a.ilim && wmax - wmin < 1.0e-10) status = -5;
    else if (L1[6] > 0.5) status = -6;
    else if (a.it > 2 && fmax(L1[13], hmax) - fmin(L1[12], hmin) < 1.0e-12) status = -7;
    else if (L1[7] > 0.5 && sigma * L1[4] < 1.0e-11 * a.insigma) status = -8;
    if (best_row < 0) status = SP_STATUS_INTERNAL;
  }
  // alpha and friends, _vdcma.py:318-329
  const double gamma = 1.0 / sqrt(1.0 + nv2);
  double alpha = sqrt(nv2 * nv2 + (1.0 + nv2) / vmax * (2.0 - gamma)) / (2.0 + nv2), beta = 0.0;
  if (alpha < 1.0) beta = (4.0 - (2.0 - gamma) / vmax) / ((1.0 + 2.0 / nv2) * (1.0 + 2.0 / nv2));
  else alpha = 1.0;
  const double bsca = 2.0 * alpha * alpha - beta;
  // rank-one vectors from pc / dvec, then p = cmu p_mu (+ c1 p_1), q likewise
  const double k1 = nv2 / (1.0 + nv2);
  T pv[kVdNpt], qv[kVdNpt];
  double vq[1] = {0.0};
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    const double vn = (double)vnT[k];
    double p = 0.0, q = 0.0;
    if (a.cmu != 0.0) {  // rank-mu vectors from the factored sums (see vd_wsum_kernel)
      const double inv = 1.0 / (double)dv[k], bq = (double)Sb[k] * inv;
      p = a.cmu * ((double)Sa[k] * inv * inv - k1 * (vn * bq) - a.wsum);
      q = a.cmu * (bq - vn * hmu);
    }
    if (hsig && a.c1 != 0.0) {
      const double y1 = (double)div_rn(pc[k], dv[k]);
      p += a.c1 * (y1 * y1 - k1 * (yv1 * y1 * vn) - 1.0);
      q += a.c1 * (yv1 * y1 - (0.5 * (yv1 * yv1 + 1.0 + nv2)) * vn);
    }
    pv[k] = (T)p;
    qv[k] = (T)q;
    if (ok[k]) vq[0] += vn * q;
  }
  double up = 1.0;
  T vv2[kVdNpt], dv2[kVdNpt];
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) vv2[k] = vv[k], dv2[k] = dv[k];
  if (a.cmu + a.c1 > 0.0) {  // natural gradient, _vdcma.py:444-458
    const int sum1[1] = {RED_SUM};
    VD_STAMP(4);
    reduce1<1, kUpThreads / 32>(vq, sum1, s_r1, ph);
    VD_STAMP(5);
    T sv[kVdNpt];
    double red3[2] = {0.0, 0.0};  // ria, via
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k) {
      const double vn = (double)vnT[k], vnn = vn * vn, avec = 2.0 - (bsca + 2.0 * alpha * alpha) * vnn;
      const double r = (double)pv[k] - alpha / (1.0 + nv2) * ((2.0 + nv2) * (double)qv[k] * vn - nv2 * vq[0] * vnn);
      sv[k] = (T)r;
      if (ok[k]) {
        red3[0] += r * (vnn / avec);
        red3[1] += vnn * (vnn / avec);
      }
    }
    {
      const int op[2] = {RED_SUM, RED_SUM};
      reduce1<2, kUpThreads / 32>(red3, op, s_r1, ph);
      VD_STAMP(6);
    }
    const double ria = red3[0], via = red3[1];
    double svnn[1] = {0.0};
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k) {
      const double vn = (double)vnT[k], vnn = vn * vn, avec = 2.0 - (bsca + 2.0 * alpha * alpha) * vnn;
      const double sn = (double)sv[k] / avec - bsca * ria / (1.0 + bsca * via) * (vnn / avec);
      sv[k] = (T)sn;
      if (ok[k]) svnn[0] += sn * vnn;
    }
    reduce1<1, kUpThreads / 32>(svnn, sum1, s_r1, ph);
    VD_STAMP(7);
    T ngv[kVdNpt], ngd[kVdNpt];
    double red4[2] = {0.0, inf};  // |ngv|^2, min D / |ngd|
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k) {
      const double vn = (double)vnT[k], sn = (double)sv[k];
      const double gv = (double)qv[k] / nv - alpha / nv * ((2.0 + nv2) * (vn * sn) - svnn[0] * vn);
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
      reduce1<2, kUpThreads / 32>(red4, op, s_r1, ph);
      VD_STAMP(8);
    }
    up = fmin(1.0, 0.7 * nv / sqrt(red4[0]));  // at most 70 % change, _vdcma.py:361-363
    up = fmin(up, 0.7 * red4[1]);
#pragma unroll
    for (int k = 0; k < kVdNpt; ++k) {
      vv2[k] = add_rn(vv[k], mul_rn((T)up, ngv[k]));
      dv2[k] = add_rn(dv[k], mul_rn((T)up, ngd[k]));
      if (ok[k]) {
        const int n = tid + k * nt;
        a.vvec[n] = vv2[k];
        a.dvec[n] = dv2[k];
      }
    }
  }
  // next generation: |v|^2, vn, diagC, the fused sampling constants and (in-kernel draws, still running)
  // the injected direction dy = |g| / |dx|_C dx of _vdcma.py:243-246 with a fresh g ~ N(0, I)
  const bool inject_next = !a.host_z && status == SP_RUNNING;
  double L6[4] = {0.0, 0.0, 0.0, 0.0};  // |v'|^2, |g|^2, |dx/D'|^2, (dx/D').v'
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    if (!ok[k]) continue;
    const int n = tid + k * nt;
    L6[0] += (double)vv2[k] * (double)vv2[k];
    if (inject_next) {
      T z[Num<T>::VEC];
      normal_block(philox4x32((uint32_t)(n / Num<T>::VEC), 0u, (uint32_t)(a.it + 1), kVdInject, a.seed), z);
      const T g = z[n % Num<T>::VEC];
      const double ddx = (double)div_rn(dx[k], dv2[k]);
      L6[1] += (double)g * (double)g;
      L6[2] += ddx * ddx;
      L6[3] += ddx * (double)vv2[k];
    }
  }
  {
    const int op[4] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM};
    VD_STAMP(9);
    reduce1<4, kUpThreads / 32>(L6, op, s_r1, ph);
    VD_STAMP(10);
  }
  const double nv2n = L6[0], nvn = sqrt(nv2n);
  const T kinj = inject_next ? (T)(sqrt(L6[1]) / sqrt(L6[2] - L6[3] * L6[3] / (1.0 + nv2n))) : T(0);
#pragma unroll
  for (int k = 0; k < kVdNpt; ++k) {
    if (!ok[k]) continue;
    const int n = tid + k * nt;
    const T v = vv2[k], d = dv2[k], sc = a.xscale[n];
    a.vn[n] = div_rn(v, (T)nvn);
    a.diagC[n] = mul_rn(mul_rn(d, add_rn(T(1), mul_rn(v, v))), d);  // _vdcma.py:251-256
    a.fuse_a()[n] = ((T)sigma * d) * sc;
    a.fuse_b()[n] = xm[k] * sc + a.xshift[n];
    if (inject_next) a.dy[n] = mul_rn(kinj, dx[k]);
  }
  if (tid == 0) {
    c->base.gbest_row = best_row;
    c->base.gfit = best;
    a.besthist[a.it - 1] = (T)best;
    c->hsig = hsig ? 1 : 0;
    c->nfev += a.P;
    c->sigma = sigma;
    if (injected) c->vd_ps = ps_new;
    c->inject = 1;
    c->aux[0] = nv2n;
    c->aux[1] = nvn;
    c->aux[2] = up;
    c->base.nit = a.it;
    c->base.status = status;
    c->aux[kVdAuxZgen] = (draw_next && status == SP_RUNNING) ? (double)(a.it + 1) : 0.0;
  }
  VD_STAMP(11);
  if (tid == 0) vd_time_stamp(15);
  if (draw_next) vd_zgen_rows<T>(a, keys);  // whatever rows are left
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
  a.lean = st->lean;
  a.evaluate = evaluate;
  // y and x of a population larger than the L2 are stored evict-first (st.global.cs); SP_VD_PLAIN_STORES=1
  // keeps normal stores (profiling switch)
  static const bool plain_stores = getenv("SP_VD_PLAIN_STORES") != nullptr;
  // (lean: only y is stored, and vd_wsum re-reads half of it right away -- keep it in the L2 when it fits)
  a.stream_stores = !plain_stores && (st->lean ? 1 : 2) * (size_t)st->P * st->ld * sizeof(T) > ((size_t)96 << 20) ? 1 : 0;
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

// the device-resident loop on full-warp rows: register tiles up to 16 scalars per lane (compile-time configured
// variants), the shared-memory row kernel above for wider rows
template <typename T, int C>
static void vd_sample_fast(const VdPtrs<T>& a, const PhiloxKeys& keys, int64_t P, bool full, int fast, int grid, bool pdl,
                           cudaStream_t s) {
  if constexpr (C >= 4) {
    constexpr size_t smem = vd_smem_bytes<C, T>();
    constexpr int per_sm = C * (int)sizeof(T) <= 32 ? 4 : 2;
    auto kf = vd_sample_smem_kernel<T, C, true>;
    auto kp = vd_sample_smem_kernel<T, C, false>;
    static thread_local bool configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {
      cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      configured[dev] = true;
    }
    int64_t g = (P + kThreads / 32 - 1) / (kThreads / 32), cap = (int64_t)sm_count() * per_sm;
    if (g > cap) g = cap;
    if (full) launch_pdl(kf, dim3((unsigned)g), dim3(kThreads), smem, s, pdl, a, keys);
    else launch_pdl(kp, dim3((unsigned)g), dim3(kThreads), smem, s, pdl, a, keys);
  } else {
    if (fast == 1) {
      if (full) launch_pdl(vd_sample_eval_kernel<T, C, 32, true, false, 1>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);
      else launch_pdl(vd_sample_eval_kernel<T, C, 32, false, false, 1>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);
    } else {
      if (full) launch_pdl(vd_sample_eval_kernel<T, C, 32, true, false, 2>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);
      else launch_pdl(vd_sample_eval_kernel<T, C, 32, false, false, 2>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys);
    }
  }
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
  int grid = grid_for_rows(st->P, sh.lpr, sh.ch * (int)sizeof(T) <= 32 ? 3 : 2);
  // fixed-shape instantiations only for full-warp rows (the large-N case that matters)
  const bool full = sh.lpr == 32 && st->N == sh.ch * 32 * Num<T>::VEC && st->ld == st->N;
  const bool clip = st->constraint == SP_CONS_PENALIZE;
  const PhiloxKeys keys = philox_keys(st->seed);
  const bool pdl = !st->host_z;  // with host draws the caller's copies sit between the generations anyway
  // the device-resident loop (in-kernel draws, lean, objective on the device, no Penalize) on full-warp rows
  // runs the compile-time configured variants
  const int fast = (st->lean && !st->host_z && evaluate && !clip && sh.lpr == 32) ? (a.stream_stores ? 2 : 1) : 0;
  if (!fast) grid = grid_for_rows(st->P, sh.lpr, 2);  // the run-time configured variant keeps 128 registers
#define SP_GO(TT, C, L, F, CL, FA) launch_pdl(vd_sample_eval_kernel<TT, C, L, F, CL, FA>, dim3(grid), dim3(kThreads), 0, s, pdl, a, keys)
#define SP_CALL(TT, C, L)                                                  \
  do {                                                                     \
    if (L == 32 && fast != 0) {                                            \
      vd_sample_fast<TT, C>(a, keys, st->P, full, fast, grid, pdl, s);     \
    } else if (L == 32 && full) {                                          \
      if (clip) SP_GO(TT, C, 32, true, true, 0);                           \
      else SP_GO(TT, C, 32, true, false, 0);                               \
    } else if (clip) {                                                     \
      SP_GO(TT, C, L, false, true, 0);                                     \
    } else {                                                               \
      SP_GO(TT, C, L, false, false, 0);                                    \
    }                                                                      \
  } while (0)
  SP_DISPATCH_SHAPE(T, sh, SP_CALL);
#undef SP_CALL
#undef SP_GO
  SP_CHECK_LAUNCH();
  return SP_OK;
}

template <typename T>
static int vd_update(const sp_vd_state* st, int it, cudaStream_t s) {
  const VdPtrs<T> a = vd_ptrs<T>(st, it, 1);
  const int64_t P = st->P;
  const int N = st->N;
  if (st->constraint == SP_CONS_PENALIZE) {
    if (rank_launch<T>(a.arfit, P, a.rank, nullptr, s, &st->ctrl->base.status) != cudaSuccess) return SP_ERR_CUDA;
    scatter_sorted_kernel<T><<<cdiv(P, 256) < 1024 ? cdiv(P, 256) : 1024, 256, 0, s>>>(a.arfit, a.rank, a.sorted(), P, st->ctrl);
    SP_CHECK_LAUNCH();
    vd_penalty_state_kernel<T><<<1, 256, 0, s>>>(a);
    SP_CHECK_LAUNCH();
    penalty_add_kernel<T><<<cdiv(P, 8) < sm_count() * 8 ? cdiv(P, 8) : sm_count() * 8, 256, 0, s>>>(
        a.arx, a.coef(), a.arfit, P, N, st->ld, st->ctrl);
    SP_CHECK_LAUNCH();
  }
  if (rank_launch_ws<T>(a.arfit, P, a.rank, a.rank_ws(), s, &st->ctrl->base.status) != cudaSuccess) return SP_ERR_CUDA;
  launch_pdl(vd_wsum_kernel<T>, dim3(cdiv(st->ld, 256 * Num<T>::VEC), a.chunks), dim3(kWsThreads), 0, s, true, a);
  SP_CHECK_LAUNCH();
  // chunk partials -> sums, then (last CTA) the update itself; also refreshes vn / diagC / fuse_a / fuse_b
  // (and dy) for the next generation
  cudaError_t le;
  const int n1 = (int)cdiv(3 * (int64_t)N, kVdUpOut);
  // z of the next generation CAN be drawn inside this kernel (on the SMs its single-CTA phase leaves idle) when the
  // next sampling launch is the shared-memory-row kernel of the lean device-resident loop (vd_sample_fast with
  // CH >= 4).  Measured on B200 (C5 fp32, profiles/r02_vd_zgen_experiment.txt): SLOWER, 101.1 vs 90.3 us per
  // generation -- the draws take longer than the single-CTA phase they hide behind (update 20 -> 25.5 us) and the
  // sampling kernel gains nothing from reading z back (64 MB of z + 64 MB of y no longer fit the L2 together;
  // sample + rank 50.7 -> 55.1 us).  Off unless SP_VD_ZGEN=1.
  static const bool no_zgen = getenv("SP_VD_ZGEN") == nullptr;
  Shape sh;
  const bool wide = pick_shape(N, Num<T>::VEC, &sh) && sh.lpr == 32 && sh.ch >= 4;
  const int zgen = (!no_zgen && st->lean && !st->host_z && st->constraint != SP_CONS_PENALIZE && wide) ? 1 : 0;
  const int grid = zgen && sm_count() > n1 ? sm_count() : n1;
  const PhiloxKeys keys = philox_keys(st->seed);
  if (N <= 256) le = launch_pdl(vd_update_kernel<T, 1, 256>, dim3(grid), dim3(256), 0, s, true, a, n1, zgen, keys);
  else if (N <= 512) le = launch_pdl(vd_update_kernel<T, 1, 512>, dim3(grid), dim3(512), 0, s, true, a, n1, zgen, keys);
  else if (N <= 1024) le = launch_pdl(vd_update_kernel<T, 1, 1024>, dim3(grid), dim3(1024), 0, s, true, a, n1, zgen, keys);
  else le = launch_pdl(vd_update_kernel<T, 2, 1024>, dim3(grid), dim3(1024), 0, s, true, a, n1, zgen, keys);
  (void)le;
  SP_CHECK_LAUNCH();
  return SP_OK;
}

// ---- per-dtype entry points (one translation unit per dtype: vdcma_f32.cu / vdcma_f64.cu, so the build of this --
// ---- file's many sampling-kernel instantiations runs in parallel) ------------------------------------------------
template <typename T>
static int vd_refresh_t(const sp_vd_state* st, cudaStream_t s) {
  vd_refresh_kernel<T><<<1, 256, 0, s>>>(vd_ptrs<T>(st, 1, 0));
  SP_CHECK_LAUNCH();
  return SP_OK;
}
static inline int vd_clocks_t(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_vd_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}

}  // namespace sp
