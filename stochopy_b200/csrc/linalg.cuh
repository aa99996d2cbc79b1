// Small dense linear algebra for the CMA-ES path: tiled GEMMs with fused
// prologues/epilogues and a one-sided Jacobi symmetric eigensolver.
// CMA-ES matrices are N x N with N <= ~1024 and the GEMMs are <= 1 GFLOP, fp64 by
// default.  tcgen05.mma has no f64 kind, so the fp64 GEMMs run on the tensor cores through
// mma.sync.m8n8k4.f64 (DMMA); the fp32 ones are CUDA-core FFMA kernels.
#pragma once
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace sp {

constexpr int kGemmTile = 64;  // output tile edge
constexpr int kGemmK = 16;     // k-chunk
constexpr int kGemmLd = kGemmTile + 8;  // shared-memory row stride: 72 = 8 mod 32 keeps the mma fragment loads conflict-free

// Inner product of one staged k-chunk, 256 threads, As[k][m] / Bs[k][n] in shared memory.
//  * double: the fp64 tensor-core instruction mma.sync.m8n8k4.f64 (SASS DMMA; tcgen05.mma has
//    no f64 kind).  8 warps as 2 (m) x 4 (n); a warp owns a 32 x 16 block = 4 x 2 mma tiles;
//    fragments (PTX ISA, m8n8k4): A row = lane / 4, k = lane % 4; B k = lane % 4, col = lane / 4;
//    C row = lane / 4, cols = 2 (lane % 4) + {0, 1}.  6 shared loads feed 8 DMMA = 2048 FMA.
//  * float: CUDA-core FFMA, each thread a 4 x 4 block.
template <typename T>
struct GemmAcc;
template <>
struct GemmAcc<double> {
  double c[4][2][2];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  }
  __device__ __forceinline__ void chunk(const double (*As)[kGemmLd], const double (*Bs)[kGemmLd]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
#pragma unroll
    for (int kk = 0; kk < kGemmK; kk += 4) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk + t4][wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Bs[kk + t4][wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1])
                       : "d"(a[i]), "d"(b[j]));
    }
  }
  template <typename F>
  __device__ __forceinline__ void each(F f) const {  // f(m, n, value) with tile-local indices
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) f(wm + i * 8 + g, wn + j * 8 + t4 * 2 + e, c[i][j][e]);
  }
};
template <>
struct GemmAcc<float> {
  float c[4][4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] = 0.0f;
  }
  __device__ __forceinline__ void chunk(const float (*As)[kGemmLd], const float (*Bs)[kGemmLd]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int kk = 0; kk < kGemmK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] += a[i] * b[j];
    }
  }
  template <typename F>
  __device__ __forceinline__ void each(F f) const {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) f(ty * 4 + i, tx * 4 + j, c[i][j]);
  }
};

// C[m][n] = epi(sum_k A(m,k) * B(n,k))        ("NT": both operands k-contiguous)
// LoadA / LoadB: functors (row, k) -> T (0 outside bounds handled here); Epi: (m, n, acc).
// The k-chunks are software pipelined: the global loads of chunk k+1 are issued (into registers) before
// the tensor-core work on chunk k, so their latency hides behind the mma sequence.
template <typename T, typename LoadA, typename LoadB, typename Epi>
__device__ __forceinline__ void gemm_nt_tile(int M, int Nn, int K, int m0, int n0, LoadA la, LoadB lb, Epi epi) {
  __shared__ T As[kGemmK][kGemmLd];
  __shared__ T Bs[kGemmK][kGemmLd];
  const int tid = threadIdx.x;
  GemmAcc<T> acc;
  acc.clear();
  T ra[4], rb[4];
  // 64 x 16 elements per operand, 4 per thread; k fastest so global reads are contiguous
  auto fetch = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + t * 256, r = e >> 4, kk = e & 15;
      const int m = m0 + r, n = n0 + r, k = k0 + kk;
      ra[t] = (m < M && k < K) ? la(m, k) : T(0);
      rb[t] = (n < Nn && k < K) ? lb(n, k) : T(0);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kGemmK) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + t * 256, r = e >> 4, kk = e & 15;
      As[kk][r] = ra[t];
      Bs[kk][r] = rb[t];
    }
    __syncthreads();
    if (k0 + kGemmK < K) fetch(k0 + kGemmK);
    acc.chunk(As, Bs);
    __syncthreads();
  }
  acc.each([&](int i, int j, T v) {
    const int m = m0 + i, n = n0 + j;
    if (m < M && n < Nn) epi(m, n, v);
  });
}

// C[r][c] = sum_i A(i, r) * B(i, c) over i in [i0, i1)   ("TN": reduction over rows)
template <typename T, typename LoadA, typename LoadB, typename Epi>
__device__ __forceinline__ void gemm_tn_tile(int R, int Cc, int i0, int i1, int r0, int c0, LoadA la, LoadB lb, Epi epi) {
  __shared__ T As[kGemmK][kGemmLd];
  __shared__ T Bs[kGemmK][kGemmLd];
  const int tid = threadIdx.x;
  GemmAcc<T> acc;
  acc.clear();
  T ra[4], rb[4];
  auto fetch = [&](int ib) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + t * 256, kk = e >> 6, cc = e & 63;  // column fastest: contiguous row reads
      const int i = ib + kk;
      ra[t] = (i < i1 && r0 + cc < R) ? la(i, r0 + cc) : T(0);
      rb[t] = (i < i1 && c0 + cc < Cc) ? lb(i, c0 + cc) : T(0);
    }
  };
  fetch(i0);
  for (int ib = i0; ib < i1; ib += kGemmK) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + t * 256, kk = e >> 6, cc = e & 63;
      As[kk][cc] = ra[t];
      Bs[kk][cc] = rb[t];
    }
    __syncthreads();
    if (ib + kGemmK < i1) fetch(ib + kGemmK);
    acc.chunk(As, Bs);
    __syncthreads();
  }
  acc.each([&](int i, int j, T v) {
    const int r = r0 + i, c = c0 + j;
    if (r < R && c < Cc) epi(r, c, v);
  });
}

// ---- symmetric eigendecomposition: cyclic one-sided Jacobi on rows -----------------
// W = Q0 C (Q0 = I, or the previous eigenvectors as rows when warm) is driven to mutually
// orthogonal rows by plane rotations of row pairs (Hestenes).  At convergence
// W = Lambda Q with Q orthogonal, so for the positive (semi)definite C of CMA-ES
//   lambda_j = |w_j|,  q_j = w_j / |w_j|
// and the accumulated rotations never have to be stored: only W is rotated (half the
// shared-memory traffic and flops of carrying Q along).  The stopping test is relative
// (|w_p.w_q| <= tol |w_p||w_q|), so the normalised rows are orthogonal to tol.
// warm != 0: the covariance moves little per generation, so 3-4 sweeps instead of ~9.
template <typename T>
struct JacobiEps;
template <>
struct JacobiEps<double> {
  static __device__ double eps() { return 2.220446049250313e-16; }
};
template <>
struct JacobiEps<float> {
  static __device__ float eps() { return 1.1920929e-7f; }
};

constexpr int kJacobiThreads = 1024;
constexpr int kJacobiRegs = 8;  // row elements per lane staged in registers (N <= 256)

// A sweep whose largest |cos(w_p, w_q)| (before its rotations) stayed below `quiet` needs no follow-up sweep:
// the rotations of that sweep leave a residual of about quiet^2 (quadratic convergence; measured per-sweep maxima
// of a warm-started N = 256 decomposition: 1.8e-2, 4.9e-3, 2.5e-5, 6.1e-10 -- profiles/r02_eigh_sweeps.txt).
// Default: the rounding-noise bound 0.25 sqrt(tol) (3e-8 in fp64, 7e-4 in fp32).  A looser fp64 threshold
// (SP_EIGH_QUIET=5e-5: residual 2.5e-9, 3 sweeps instead of 4, C4 2.49 -> 2.29 ms per generation) was measured and
// REJECTED: B^T B leaves the identity by 3e-8 and a 25-generation CMA-ES trajectory leaves the oracle's by more
// than 1e-6 relative (tests/test_gpu_es.py, tests/test_gpu_sizes.py) -- the switch stays for experiments only.
template <typename T>
__device__ __forceinline__ float jacobi_quiet(T tol, float q64) {
  const float q = 0.25f * sqrtf((float)tol);
  return sizeof(T) == 8 ? fmaxf(q, q64) : q;
}
inline float jacobi_quiet64() {
  static const char* env = getenv("SP_EIGH_QUIET");
  return env != nullptr ? (float)atof(env) : 0.0f;
}

template <typename T>
__device__ __forceinline__ T jacobi_tol(int N) {
  // rotate while |w_p.w_q| exceeds the rounding noise of a length-N dot product
  return T(4) * JacobiEps<T>::eps() * sqrt((T)(N < 16 ? 16 : N));
}

// Rotation (cs, sn) that orthogonalises a pair with |w_p|^2 = al, |w_q|^2 = be, w_p.w_q = ga.
// Every lane of the warp executes this, and fp64 division / sqrt are ~25-instruction
// sequences on a half-rate pipe: done naively (2 divisions, 3 square roots) the scalar
// work outweighs the rotation itself.  Only cs^2 + sn^2 = 1 has to hold to working
// precision (orthogonality of the accumulated transformation); the ANGLE may be
// approximate -- an error of 1e-7 leaves a residual of 1e-7 |ga|, which the quadratically
// convergent sweeps absorb.  So: tan in fp32 from operands scaled by 2^-exponent(nrm)
// (range-safe for any eigenvalue scale), then cs = rsqrt(1 + t^2) by two fp64 Newton
// steps from the fp32 seed (relative error ~1e-14 -> ~1e-28), sn = cs t.
__device__ __forceinline__ float jacobi_tan(float df, float gf) {
  const float zf = df / (2.0f * gf);  // zeta = (be - al) / (2 ga)
  const float az = fabsf(zf);
  // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)); for large |zeta| (zeta^2 would overflow) t = 1 / (2 zeta)
  return az > 1.0e4f ? 0.5f / zf : copysignf(1.0f, zf) / (az + sqrtf(1.0f + zf * zf));
}
// `prod` = al * be (> 0).  Returns |cos(w_p, w_q)| (fp32, for the quiet-sweep test).
__device__ __forceinline__ float jacobi_angle(double al, double be, double ga, double prod, double* cs, double* sn) {
  // 2^-ex with ex = exponent(sqrt(prod)) = exponent(prod) / 2: scales ga and be - al into fp32 range
  const int ex = ((((__double2hiint(prod) >> 20) & 0x7ff) - 1023) >> 1);
  int bexp = 1023 - ex;
  bexp = bexp < 1 ? 1 : (bexp > 2046 ? 2046 : bexp);
  const double sc = __hiloint2double(bexp << 20, 0);
  const float gf = (float)(ga * sc);
  const float tf = jacobi_tan((float)((be - al) * sc), gf);
  const float ratio = fabsf(gf) * rsqrtf((float)(prod * sc * sc));
  const double t = (double)tf;
  const double x = fma(t, t, 1.0);
  double y = (double)rsqrtf(fmaf(tf, tf, 1.0f));
  y = y * fma(-0.5 * x, y * y, 1.5);
  y = y * fma(-0.5 * x, y * y, 1.5);
  *cs = y;
  *sn = y * t;
  return ratio;
}
__device__ __forceinline__ float jacobi_angle(float al, float be, float ga, float prod, float* cs, float* sn) {
  const float sc = rsqrtf(prod);
  const float gf = ga * sc;
  const float t = jacobi_tan((be - al) * sc, gf);
  const float x = fmaf(t, t, 1.0f);
  float y = rsqrtf(x);
  y = y * fmaf(-0.5f * x, y * y, 1.5f);
  *cs = y;
  *sn = y * t;
  return fabsf(gf);
}

// one rotation of the row pair (wp, wq) by a full warp; returns |cos(w_p, w_q)| if the
// pair was rotated, 0 if it already was orthogonal to tol.  REGS: rows of N <= 32 * kJacobiRegs
// scalars travel through registers once (dot products and rotation share the loads).
template <typename T, bool REGS>
__device__ __forceinline__ float jacobi_rotate(T* __restrict__ wp, T* __restrict__ wq, int N, int lane, T tol) {
  T ra[kJacobiRegs], rb[kJacobiRegs];
  T al = 0, be = 0, ga = 0;
  if (REGS) {
#pragma unroll
    for (int u = 0; u < kJacobiRegs; ++u) {
      const int k = lane + 32 * u;
      ra[u] = k < N ? wp[k] : T(0);
      rb[u] = k < N ? wq[k] : T(0);
    }
#pragma unroll
    for (int u = 0; u < kJacobiRegs; ++u) {
      al += ra[u] * ra[u];
      be += rb[u] * rb[u];
      ga += ra[u] * rb[u];
    }
  } else {
    for (int k = lane; k < N; k += 32) {
      const T a = wp[k], b = wq[k];
      al += a * a;
      be += b * b;
      ga += a * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    al += __shfl_xor_sync(0xffffffffu, al, o);
    be += __shfl_xor_sync(0xffffffffu, be, o);
    ga += __shfl_xor_sync(0xffffffffu, ga, o);
  }
  const T prod = al * be;  // |ga| > tol sqrt(al be), without the square root
  if (!(ga * ga > tol * tol * prod) || !(prod > T(0))) return 0.0f;
  T cs, sn;
  const float ratio = jacobi_angle(al, be, ga, prod, &cs, &sn);
  if (REGS) {
#pragma unroll
    for (int u = 0; u < kJacobiRegs; ++u) {
      const int k = lane + 32 * u;
      if (k < N) {
        wp[k] = cs * ra[u] - sn * rb[u];
        wq[k] = sn * ra[u] + cs * rb[u];
      }
    }
  } else {
    for (int k = lane; k < N; k += 32) {
      const T a = wp[k], b = wq[k];
      wp[k] = cs * a - sn * b;
      wq[k] = sn * a + cs * b;
    }
  }
  return ratio;
}

// circle-method tournament: pair i of round r among n (even) players
__device__ __forceinline__ void circle_pair(int n, int r, int i, int* p, int* q) {
  if (i == 0) {
    *p = n - 1;
    *q = r;
  } else {
    *p = (r + i) % (n - 1);
    *q = (r - i + (n - 1)) % (n - 1);
  }
}

// W0 = Q0 C without materialising Q0: W0[j][c] = sum_k B[k][j] C[k][c] (warm) or C (cold).
// Threads e -> (j, c) with c fastest: B[k][j] is a broadcast, C[k][c] is coalesced.
template <typename T>
__device__ __forceinline__ void jacobi_start(const T* __restrict__ C, const T* __restrict__ B, T* __restrict__ W, int N,
                                             int warm, int first, int stride) {
  for (int e = first; e < N * N; e += stride) {
    const int j = e / N, c = e - j * N;
    T acc;
    if (warm) {
      T a0 = 0, a1 = 0;
      int k = 0;
      for (; k + 1 < N; k += 2) {
        a0 += B[k * N + j] * C[k * N + c];
        a1 += B[(k + 1) * N + j] * C[(k + 1) * N + c];
      }
      if (k < N) a0 += B[k * N + j] * C[k * N + c];
      acc = a0 + a1;
    } else {
      acc = C[e];
    }
    W[e] = acc;
  }
}

// rows of W (orthogonal) -> eigenvalues (ascending, stable) and sign-normalised eigenvectors
// in the columns of B.  Phase 1 (norms) and phase 2 (rank, sign, scatter) are separated by
// the caller's barrier; `lam` is N scalars of scratch visible to all participating warps.
template <typename T>
__device__ __forceinline__ void jacobi_norms(const T* __restrict__ W, int N, T* lam, int warp0, int nwarps, int lane) {
  for (int j = warp0; j < N; j += nwarps) {
    T acc = 0;
    for (int k = lane; k < N; k += 32) {
      const T a = W[(size_t)j * N + k];
      acc += a * a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) lam[j] = sqrt(acc);
  }
}
template <typename T>
__device__ __forceinline__ void jacobi_scatter(const T* __restrict__ W, int N, const T* lam, T* __restrict__ w_out,
                                               T* __restrict__ B, int warp0, int nwarps, int lane) {
  for (int j = warp0; j < N; j += nwarps) {
    const T mine = lam[j];
    int rk = 0;
    for (int k = lane; k < N; k += 32) {
      const T o = lam[k];
      rk += (o < mine) || (o == mine && k < j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    // sign: largest |component| positive (first one on ties)
    T best = T(-1);
    int bidx = 0;
    for (int k = lane; k < N; k += 32) {
      const T a = fabs(W[(size_t)j * N + k]);
      if (a > best) {
        best = a;
        bidx = k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) {
        best = ob;
        bidx = oi;
      }
    }
    const T inv = mine > T(0) ? T(1) / mine : T(0);
    const T sgn = (W[(size_t)j * N + bidx] < T(0) ? T(-1) : T(1)) * inv;
    if (lane == 0) w_out[rk] = mine;
    for (int k = lane; k < N; k += 32) B[(size_t)k * N + rk] = sgn * W[(size_t)j * N + k];
  }
}

// One CTA of 1024 threads, W in shared memory (N^2 scalars fit); a warp owns one (p, q)
// pair of the round-robin round.
template <typename T>
__global__ void __launch_bounds__(kJacobiThreads, 1)
jacobi_eigh_kernel(T* __restrict__ C, int N, T* __restrict__ w_out, T* __restrict__ B, T* __restrict__ lam_g,
                   int warm, const int* __restrict__ gate, const int* __restrict__ status_gate,
                   int* __restrict__ sweeps_out, float q64) {
  extern __shared__ __align__(16) unsigned char jsm[];
  __shared__ unsigned int s_off;
  if (gate != nullptr && *gate == 0) return;
  if (status_gate != nullptr && *status_gate != SP_RUNNING) return;
  T* W = reinterpret_cast<T*>(jsm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kJacobiThreads / 32;

  // symmetrise from the upper triangle (_cmaes.py:303): C = triu(C) + triu(C,1)^T
  for (int e = tid; e < N * N; e += kJacobiThreads) {
    const int r = e / N, c = e - r * N;
    if (r > c) C[e] = C[c * N + r];
  }
  __syncthreads();
  jacobi_start<T>(C, B, W, N, warm, tid, kJacobiThreads);
  __syncthreads();

  const T tol = jacobi_tol<T>(N);
  const float quiet = jacobi_quiet<T>(tol, q64);  // a sweep below this needs no follow-up sweep
  const int n = N + (N & 1);  // even player count; index N (if any) is a bye
  const int half = n / 2;
  const bool regs = N <= 32 * kJacobiRegs;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    if (tid == 0) s_off = 0u;
    __syncthreads();
    float worst = 0.0f;
    for (int r = 0; r < n - 1; ++r) {
      for (int i = warp; i < half; i += nwarps) {
        int p, q;
        circle_pair(n, r, i, &p, &q);
        if (p >= N || q >= N) continue;
        const float c = regs ? jacobi_rotate<T, true>(W + (size_t)p * N, W + (size_t)q * N, N, lane, tol)
                             : jacobi_rotate<T, false>(W + (size_t)p * N, W + (size_t)q * N, N, lane, tol);
        worst = fmaxf(worst, c);
      }
      __syncthreads();
    }
    if (lane == 0 && worst > 0.0f) atomicMax(&s_off, __float_as_uint(worst));
    __syncthreads();
    const unsigned int off = s_off;
    __syncthreads();
    if (off == 0u || __uint_as_float(off) < quiet) break;
  }
  if (tid == 0 && sweeps_out != nullptr) *sweeps_out = sweep + 1;
  jacobi_norms<T>(W, N, lam_g, warp, nwarps, lane);
  __syncthreads();
  jacobi_scatter<T>(W, N, lam_g, w_out, B, warp, nwarps, lane);
}

// Fixed-shape fast path of the block kernel: rows of exactly CH * 32 sixteen-byte vectors
// (N = CH * 32 * VEC), every row valid.  Lane l owns vectors l, l + 32, ... of a row, so the
// loads/stores are 128-bit, there are no bounds predicates and all addresses are constants
// off the row base (the generic path spends ~640 warp instructions per pair, mostly
// predicates and address arithmetic; this one ~190).
template <typename T, int CH>
__device__ __forceinline__ float jacobi_rotate_v(T* __restrict__ wp, T* __restrict__ wq, int lane, T tol) {
  using V = typename Num<T>::vec_t;
  constexpr int VEC = Num<T>::VEC;
  T ra[CH][VEC], rb[CH][VEC];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const V a = reinterpret_cast<const V*>(wp)[lane + 32 * c];
    const V b = reinterpret_cast<const V*>(wq)[lane + 32 * c];
    const T* pa = reinterpret_cast<const T*>(&a);
    const T* pb = reinterpret_cast<const T*>(&b);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      ra[c][e] = pa[e];
      rb[c][e] = pb[e];
    }
  }
  T al = 0, be = 0, ga = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      al += ra[c][e] * ra[c][e];
      be += rb[c][e] * rb[c][e];
      ga += ra[c][e] * rb[c][e];
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    al += __shfl_xor_sync(0xffffffffu, al, o);
    be += __shfl_xor_sync(0xffffffffu, be, o);
    ga += __shfl_xor_sync(0xffffffffu, ga, o);
  }
  const T prod = al * be;  // |ga| > tol sqrt(al be), without the square root
  if (!(ga * ga > tol * tol * prod) || !(prod > T(0))) return 0.0f;
  T cs, sn;
  const float ratio = jacobi_angle(al, be, ga, prod, &cs, &sn);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    V a, b;
    T* pa = reinterpret_cast<T*>(&a);
    T* pb = reinterpret_cast<T*>(&b);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      pa[e] = cs * ra[c][e] - sn * rb[c][e];
      pb[e] = sn * ra[c][e] + cs * rb[c][e];
    }
    reinterpret_cast<V*>(wp)[lane + 32 * c] = a;
    reinterpret_cast<V*>(wq)[lane + 32 * c] = b;
  }
  return ratio;
}

// one outer round of the block kernel on the fast path: load the 2b rows of blocks (I, J)
// (warp w owns rows w and w + b: 2 CH independent 128-bit loads per lane, all in flight), a
// complete inner round-robin sweep (pair indices advance by one modulo 2b - 1 per round, so
// no division), write the rows back.
template <typename T, int CH>
__device__ __forceinline__ float jacobi_block_round_v(T* __restrict__ W, T* __restrict__ sW, int b, int I, int J,
                                                      T tol) {
  using V = typename Num<T>::vec_t;
  constexpr int N = CH * 32 * Num<T>::VEC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n2 = 2 * b;
  V* g0 = reinterpret_cast<V*>(W + (size_t)(I * b + warp) * N);
  V* g1 = reinterpret_cast<V*>(W + (size_t)(J * b + warp) * N);
  V* s0 = reinterpret_cast<V*>(sW + (size_t)warp * N);
  V* s1 = reinterpret_cast<V*>(sW + (size_t)(warp + b) * N);
  {
    V t0[CH], t1[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      t0[c] = __ldcg(g0 + lane + 32 * c);
      t1[c] = __ldcg(g1 + lane + 32 * c);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      s0[lane + 32 * c] = t0[c];
      s1[lane + 32 * c] = t1[c];
    }
  }
  __syncthreads();
  float worst = 0.0f;
  int p = warp == 0 ? n2 - 1 : warp, q = warp == 0 ? 0 : n2 - 1 - warp;
  for (int r = 0; r < n2 - 1; ++r) {
    worst = fmaxf(worst, jacobi_rotate_v<T, CH>(sW + (size_t)p * N, sW + (size_t)q * N, lane, tol));
    __syncthreads();
    if (warp != 0) p = p + 1 == n2 - 1 ? 0 : p + 1;
    q = q + 1 == n2 - 1 ? 0 : q + 1;
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    __stcg(g0 + lane + 32 * c, s0[lane + 32 * c]);
    __stcg(g1 + lane + 32 * c, s1[lane + 32 * c]);
  }
  __syncthreads();
  return worst;
}

// Block Jacobi for matrices beyond one SM: the rows are cut into blocks of b rows; in an
// outer round every CTA of the (single) cluster loads one PAIR of blocks (2b rows of W)
// into shared memory, runs a complete inner round-robin sweep over those 2b rows there
// (b warps, one pair per warp, __syncthreads between inner rounds), and writes the rows
// back; outer rounds follow the same circle-method tournament over the blocks and are
// separated by the hardware cluster barrier.  Compared with one global barrier per scalar
// round this needs (2 N / b - 1) barriers per sweep instead of (N - 1) and keeps the
// rotations in shared memory.  A sweep whose largest |cos| was tiny ends the iteration
// without a separate all-quiet sweep (quadratic convergence).
template <typename T>
__global__ void __launch_bounds__(512, 1)
jacobi_block_kernel(T* __restrict__ C, int N, T* __restrict__ w_out, T* __restrict__ B, T* __restrict__ W,
                    T* __restrict__ lam, unsigned int* __restrict__ off, int warm, int b, int nblk,
                    const int* __restrict__ gate, const int* __restrict__ status_gate, int* __restrict__ sweeps_out,
                    float q64) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char jbs[];
  if (gate != nullptr && *gate == 0) return;
  if (status_gate != nullptr && *status_gate != SP_RUNNING) return;
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + tid;
  const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
  T* sW = reinterpret_cast<T*>(jbs);  // [2b][N]

  for (int e = gtid; e < N * N; e += gthreads) {  // C = triu(C) + triu(C,1)^T (_cmaes.py:303)
    const int r = e / N, c = e - r * N;
    if (r > c) C[e] = C[c * N + r];
  }
  if (gtid < 64) off[gtid] = 0u;
  cl.sync();
  jacobi_start<T>(C, B, W, N, warm, gtid, gthreads);
  cl.sync();

  const T tol = jacobi_tol<T>(N);
  const float quiet = jacobi_quiet<T>(tol, q64);
  const int m = nblk / 2, n2 = 2 * b;
  const bool regs = N <= 32 * kJacobiRegs;
  const int nv = N / Num<T>::VEC;  // rows are copied in 16-byte vectors when N allows it
  const bool vec_ok = (N % Num<T>::VEC) == 0;
  using V = typename Num<T>::vec_t;
  const int row_vecs = N / (32 * Num<T>::VEC);
  const int fast_ch = (N % (32 * Num<T>::VEC) == 0 && row_vecs <= 4 && N % b == 0 && nblk * b == N &&
                       (int)blockDim.x == 32 * b) ? row_vecs : 0;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    float worst = 0.0f;
    for (int R = 0; R < nblk - 1; ++R) {
      for (int bp = blockIdx.x; bp < m; bp += gridDim.x) {
        int I, J;
        circle_pair(nblk, R, bp, &I, &J);
        if (fast_ch != 0) {  // N = fast_ch * 32 * VEC, b divides N, one warp per row pair
          float c = 0.0f;
          switch (fast_ch) {
            case 1: c = jacobi_block_round_v<T, 1>(W, sW, b, I, J, tol); break;
            case 2: c = jacobi_block_round_v<T, 2>(W, sW, b, I, J, tol); break;
            case 3: c = jacobi_block_round_v<T, 3>(W, sW, b, I, J, tol); break;
            default: c = jacobi_block_round_v<T, 4>(W, sW, b, I, J, tol); break;
          }
          worst = fmaxf(worst, c);
          continue;
        }
        // load the 2b rows (rows past N are zero)
        if (vec_ok) {
          for (int e = tid; e < n2 * nv; e += blockDim.x) {
            const int lr = e / nv, k = e - lr * nv;
            const int gr = (lr < b ? I * b + lr : J * b + (lr - b));
            V v;
            if (gr < N) v = __ldcg(reinterpret_cast<const V*>(W + (size_t)gr * N) + k);
            else memset(&v, 0, sizeof(V));
            reinterpret_cast<V*>(sW + (size_t)lr * N)[k] = v;
          }
        } else {
          for (int e = tid; e < n2 * N; e += blockDim.x) {
            const int lr = e / N, k = e - lr * N;
            const int gr = (lr < b ? I * b + lr : J * b + (lr - b));
            sW[e] = gr < N ? __ldcg(W + (size_t)gr * N + k) : T(0);
          }
        }
        __syncthreads();
        for (int r = 0; r < n2 - 1; ++r) {
          if (warp < b) {
            int p, q;
            circle_pair(n2, r, warp, &p, &q);
            const int gp = (p < b ? I * b + p : J * b + (p - b)), gq = (q < b ? I * b + q : J * b + (q - b));
            if (gp < N && gq < N) {
              const float c = regs ? jacobi_rotate<T, true>(sW + (size_t)p * N, sW + (size_t)q * N, N, lane, tol)
                                   : jacobi_rotate<T, false>(sW + (size_t)p * N, sW + (size_t)q * N, N, lane, tol);
              worst = fmaxf(worst, c);
            }
          }
          __syncthreads();
        }
        if (vec_ok) {
          for (int e = tid; e < n2 * nv; e += blockDim.x) {  // write the rows back
            const int lr = e / nv, k = e - lr * nv;
            const int gr = (lr < b ? I * b + lr : J * b + (lr - b));
            if (gr < N) __stcg(reinterpret_cast<V*>(W + (size_t)gr * N) + k, reinterpret_cast<const V*>(sW + (size_t)lr * N)[k]);
          }
        } else {
          for (int e = tid; e < n2 * N; e += blockDim.x) {
            const int lr = e / N, k = e - lr * N;
            const int gr = (lr < b ? I * b + lr : J * b + (lr - b));
            if (gr < N) __stcg(W + (size_t)gr * N + k, sW[e]);
          }
        }
        __syncthreads();
      }
      cl.sync();
    }
    if (lane == 0 && worst > 0.0f) atomicMax(&off[sweep], __float_as_uint(worst));
    cl.sync();
    const unsigned int w = __ldcg(&off[sweep]);
    if (w == 0u || __uint_as_float(w) < quiet) break;
  }
  if (gtid == 0 && sweeps_out != nullptr) *sweeps_out = sweep + 1;
  jacobi_norms<T>(W, N, lam, gwarp, gwarps, lane);
  cl.sync();
  jacobi_scatter<T>(W, N, lam, w_out, B, gwarp, gwarps, lane);
}

// scratch scalars needed behind `work` for an N x N decomposition
inline size_t jacobi_work_scalars(int N) { return 2 * (size_t)N * N + (size_t)N + 64; }

template <typename T>
inline cudaError_t jacobi_launch(T* C, int N, T* w, T* B, T* work, int warm, const int* gate,
                                 const int* status_gate, int* sweeps, cudaStream_t s) {
  T* W = work;
  T* lam = work + 2 * (size_t)N * N;
  unsigned int* off = reinterpret_cast<unsigned int*>(lam + N);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  const size_t need = (size_t)N * N * sizeof(T);
  static const char* env_one = getenv("SP_EIGH_ONE_CTA");  // profiling switch: largest N on the one-CTA kernel
  const int one_cta_max = env_one != nullptr ? atoi(env_one) : 64;
  // up to 64 rows one CTA holds a whole round (32 pairs, one per warp); beyond that several SMs
  // working on block pairs beat one SM making several passes per round
  if (need <= 200 * 1024 && N <= one_cta_max) {
    auto kern = jacobi_eigh_kernel<T>;
    static thread_local bool configured[64];
    if (!configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
      configured[dev] = true;
    }
    kern<<<1, kJacobiThreads, need, s>>>(C, N, w, B, lam, warm, gate, status_gate, sweeps, jacobi_quiet64());
    return cudaGetLastError();
  }
  static const char* env_b = getenv("SP_EIGH_BLOCK");  // profiling switch: rows per block
  int b = 16;
  // aim at 16 block pairs (one per CTA of a 16-CTA cluster) per outer round: fewer warps per SM
  // shorten the latency-bound inner rounds (measured, N = 256 fp64 warm: b = 16 / 8 CTAs 2.23 ms,
  // b = 8 / 16 CTAs 1.58 ms, b = 4 / 16 CTAs x 2 pairs 2.46 ms)
  while (b > 2 && N <= 16 * b) b >>= 1;
  if (env_b != nullptr) b = atoi(env_b);
  if (b < 2) b = 2;
  if (b > 16) b = 16;
  while (b > 2 && 2 * (size_t)b * N * sizeof(T) > 200 * 1024) b >>= 1;
  int nblk = (N + b - 1) / b;
  nblk += nblk & 1;
  const int m = nblk / 2;
  static const char* env_c = getenv("SP_EIGH_CTAS");  // profiling switch: 8 = portable cluster size only
  const int cmax = env_c != nullptr && atoi(env_c) == 8 ? 8 : 16;
  const int ctas = m < cmax ? m : cmax;
  const size_t smem = 2 * (size_t)b * N * sizeof(T);
  auto kern = jacobi_block_kernel<T>;
  static thread_local bool configured[64];
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(b * 32 < 128 ? 128 : b * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, C, N, w, B, W, lam, off, warm, b, nblk, gate, status_gate, sweeps, jacobi_quiet64());
}

}  // namespace sp
