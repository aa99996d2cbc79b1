// Small dense linear algebra for the CMA-ES path: tiled GEMMs with fused
// prologues/epilogues and a one-sided Jacobi symmetric eigensolver.
// CMA-ES matrices are N x N with N <= ~1024 and the GEMMs are <= 1 GFLOP, fp64 by
// default (tcgen05 has no f64 kind), so these are CUDA-core DFMA/FFMA kernels.
#pragma once
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace sp {

constexpr int kGemmTile = 64;  // output tile edge
constexpr int kGemmK = 16;     // k-chunk
// 256 threads, each owns a 4 x 4 block of the 64 x 64 tile.

// C[m][n] = epi(sum_k A(m,k) * B(n,k))        ("NT": both operands k-contiguous)
// LoadA / LoadB: functors (row, k) -> T (0 outside bounds handled here); Epi: (m, n, acc).
template <typename T, typename LoadA, typename LoadB, typename Epi>
__device__ __forceinline__ void gemm_nt_tile(int M, int Nn, int K, int m0, int n0, LoadA la, LoadB lb, Epi epi) {
  __shared__ T As[kGemmK][kGemmTile + 4];
  __shared__ T Bs[kGemmK][kGemmTile + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int k0 = 0; k0 < K; k0 += kGemmK) {
    // 64 x 16 elements per operand, 4 per thread; k fastest so global reads are contiguous
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + t * 256, r = e >> 4, kk = e & 15;
      const int m = m0 + r, n = n0 + r, k = k0 + kk;
      As[kk][r] = (m < M && k < K) ? la(m, k) : T(0);
      Bs[kk][r] = (n < Nn && k < K) ? lb(n, k) : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGemmK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < Nn) epi(m, n, acc[i][j]);
    }
}

// C[r][c] = sum_i A(i, r) * B(i, c) over i in [i0, i1)   ("TN": reduction over rows)
template <typename T, typename LoadA, typename LoadB, typename Epi>
__device__ __forceinline__ void gemm_tn_tile(int R, int Cc, int i0, int i1, int r0, int c0, LoadA la, LoadB lb, Epi epi) {
  __shared__ T As[kGemmK][kGemmTile + 4];
  __shared__ T Bs[kGemmK][kGemmTile + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int ib = i0; ib < i1; ib += kGemmK) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + t * 256, kk = e >> 6, cc = e & 63;  // column fastest: contiguous row reads
      const int i = ib + kk;
      As[kk][cc] = (i < i1 && r0 + cc < R) ? la(i, r0 + cc) : T(0);
      Bs[kk][cc] = (i < i1 && c0 + cc < Cc) ? lb(i, c0 + cc) : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGemmK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (r < R && c < Cc) epi(r, c, acc[i][j]);
    }
}

// ---- symmetric eigendecomposition: cyclic one-sided Jacobi on rows -----------------
// W = Q C is driven to mutually orthogonal rows by plane rotations applied to the
// rows of W and Q alike (Hestenes).  At convergence row j of Q is an eigenvector
// q_j and row j of W equals lambda_j q_j, so lambda_j = w_j . q_j (signed).
// One CTA of 1024 threads; a warp owns one (p, q) pair of the round-robin round.
// W and Q live in shared memory when 2 N^2 scalars fit, otherwise in global (L2).
// warm != 0: start from Q = previous eigenvectors (rows), W = Q C -- the covariance
// moves little per generation, so 2-3 sweeps instead of ~8 from the identity.
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

template <typename T>
__global__ void __launch_bounds__(kJacobiThreads, 1)
jacobi_eigh_kernel(T* __restrict__ C, int N, T* __restrict__ w_out, T* __restrict__ B, T* __restrict__ Wg,
                   T* __restrict__ Qg, int use_smem, int warm, const int* __restrict__ gate,
                   const int* __restrict__ status_gate, int* __restrict__ sweeps_out) {
  extern __shared__ __align__(16) unsigned char jsm[];
  __shared__ unsigned int s_off;
  __shared__ int s_rank_tmp;
  if (gate != nullptr && *gate == 0) return;
  if (status_gate != nullptr && *status_gate != SP_RUNNING) return;
  T* W = use_smem ? reinterpret_cast<T*>(jsm) : Wg;
  T* Q = use_smem ? reinterpret_cast<T*>(jsm) + (size_t)N * N : Qg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kJacobiThreads / 32;

  // symmetrise from the upper triangle (_cmaes.py:303): C = triu(C) + triu(C,1)^T
  for (int e = tid; e < N * N; e += kJacobiThreads) {
    const int r = e / N, c = e - r * N;
    if (r > c) C[e] = C[c * N + r];
  }
  __syncthreads();
  if (warm) {  // Q rows = previous eigenvectors (columns of B); W = Q C
    for (int e = tid; e < N * N; e += kJacobiThreads) {
      const int j = e / N, r = e - j * N;
      Q[e] = B[r * N + j];
    }
    __syncthreads();
    for (int e = tid; e < N * N; e += kJacobiThreads) {
      const int j = e / N, c = e - j * N;
      T acc = 0;
      for (int k = 0; k < N; ++k) acc += Q[j * N + k] * C[k * N + c];
      W[e] = acc;
    }
  } else {
    for (int e = tid; e < N * N; e += kJacobiThreads) {
      const int r = e / N, c = e - r * N;
      W[e] = C[e];
      Q[e] = r == c ? T(1) : T(0);
    }
  }
  __syncthreads();

  // rotate while |w_p.w_q| exceeds the rounding noise of a length-N dot product
  const T tol = T(4) * JacobiEps<T>::eps() * sqrt((T)(N < 16 ? 16 : N));
  const int n = N + (N & 1);  // even player count; index N (if any) is a bye
  const int half = n / 2;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    if (tid == 0) s_off = 0u;
    __syncthreads();
    for (int r = 0; r < n - 1; ++r) {
      for (int i = warp; i < half; i += nwarps) {
        int p, q;
        if (i == 0) {
          p = n - 1;
          q = r;
        } else {
          p = (r + i) % (n - 1);
          q = (r - i + (n - 1)) % (n - 1);
        }
        if (p >= N || q >= N) continue;
        T* wp = W + (size_t)p * N;
        T* wq = W + (size_t)q * N;
        T al = 0, be = 0, ga = 0;
        for (int k = lane; k < N; k += 32) {
          const T a = wp[k], b = wq[k];
          al += a * a;
          be += b * b;
          ga += a * b;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        const T lim = tol * sqrt(al * be);
        if (fabs(ga) > lim && al > T(0) && be > T(0)) {
          if (lane == 0) s_off = 1u;
          const T zeta = (be - al) / (T(2) * ga);
          const T t = (zeta >= T(0) ? T(1) : T(-1)) / (fabs(zeta) + sqrt(T(1) + zeta * zeta));
          const T cs = T(1) / sqrt(T(1) + t * t), sn = cs * t;
          T* qp = Q + (size_t)p * N;
          T* qq = Q + (size_t)q * N;
          for (int k = lane; k < N; k += 32) {
            const T a = wp[k], b = wq[k];
            wp[k] = cs * a - sn * b;
            wq[k] = sn * a + cs * b;
            const T c = qp[k], d = qq[k];
            qp[k] = cs * c - sn * d;
            qq[k] = sn * c + cs * d;
          }
        }
      }
      __syncthreads();
    }
    const unsigned int off = s_off;
    __syncthreads();
    if (off == 0u) break;
  }
  if (tid == 0 && sweeps_out != nullptr) *sweeps_out = sweep + 1;

  // eigenvalues (Rayleigh product), ascending stable rank, canonical sign, scatter
  T* lam = w_out;  // temporarily unsorted in global scratch: reuse Wg tail? keep simple: two passes
  __shared__ T s_lam[1024];
  for (int j = warp; j < N; j += nwarps) {
    T acc = 0;
    for (int k = lane; k < N; k += 32) acc += W[(size_t)j * N + k] * Q[(size_t)j * N + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_lam[j] = acc;
  }
  __syncthreads();
  (void)s_rank_tmp;
  for (int j = warp; j < N; j += nwarps) {
    const T mine = s_lam[j];
    int rk = 0;
    for (int k = lane; k < N; k += 32) {
      const T o = s_lam[k];
      rk += (o < mine) || (o == mine && k < j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    // sign: largest |component| positive (first one on ties)
    T best = T(-1);
    int bidx = 0;
    for (int k = lane; k < N; k += 32) {
      const T a = fabs(Q[(size_t)j * N + k]);
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
    const T sgn = Q[(size_t)j * N + bidx] < T(0) ? T(-1) : T(1);
    if (lane == 0) lam[rk] = mine;
    for (int k = lane; k < N; k += 32) B[(size_t)k * N + rk] = sgn * Q[(size_t)j * N + k];
  }
}

// Multi-CTA variant for matrices that do not fit one SM's shared memory: the N/2 pairs of
// a round are spread over the warps of a cooperative grid (one pair per warp), W and Q
// live in global memory (L2 resident, 2 N^2 scalars), rounds are separated by grid-wide
// barriers.  Same rotations, ordering, warm start, sorting and sign rule as above.
// scratch: W, Q (2 N^2), lam (N), sweep flags (64 x 4 bytes).
// CLUSTER: the whole grid is ONE thread-block cluster (<= 16 CTAs) and rounds are separated
// by the hardware cluster barrier (barrier.cluster, a few hundred cycles, release/acquire at
// cluster scope incl. the L1 invalidate) instead of a cooperative grid barrier.
template <typename T, bool CLUSTER, int THREADS>
__global__ void __launch_bounds__(THREADS)
jacobi_grid_kernel(T* __restrict__ C, int N, T* __restrict__ w_out, T* __restrict__ B, T* __restrict__ W,
                   T* __restrict__ Q, T* __restrict__ lam, unsigned int* __restrict__ off, int warm,
                   const int* __restrict__ gate, const int* __restrict__ status_gate, int* __restrict__ sweeps_out) {
  namespace cg = cooperative_groups;
  struct Barrier {
    __device__ void sync() {
      if (CLUSTER) cg::this_cluster().sync();
      else cg::this_grid().sync();
    }
  } grid;
  if (gate != nullptr && *gate == 0) return;  // uniform over the grid: nobody reaches a barrier
  if (status_gate != nullptr && *status_gate != SP_RUNNING) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int gthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + tid;
  const int gwarp = gtid >> 5, gwarps = gthreads >> 5;

  for (int e = gtid; e < N * N; e += gthreads) {  // C = triu(C) + triu(C,1)^T (_cmaes.py:303)
    const int r = e / N, c = e - r * N;
    if (r > c) C[e] = C[c * N + r];
  }
  if (gtid < 64) off[gtid] = 0u;
  grid.sync();
  if (warm) {
    for (int e = gtid; e < N * N; e += gthreads) {
      const int j = e / N, r = e - j * N;
      Q[e] = B[r * N + j];
    }
    grid.sync();
    for (int e = gtid; e < N * N; e += gthreads) {  // W = Q C
      const int j = e / N, c = e - j * N;
      T acc = 0;
      for (int k = 0; k < N; ++k) acc += Q[j * N + k] * C[k * N + c];
      W[e] = acc;
    }
  } else {
    for (int e = gtid; e < N * N; e += gthreads) {
      const int r = e / N, c = e - r * N;
      W[e] = C[e];
      Q[e] = r == c ? T(1) : T(0);
    }
  }
  grid.sync();

  const T tol = T(4) * JacobiEps<T>::eps() * sqrt((T)(N < 16 ? 16 : N));
  const int n = N + (N & 1), half = n / 2;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    for (int r = 0; r < n - 1; ++r) {
      for (int i = gwarp; i < half; i += gwarps) {
        int p, q;
        if (i == 0) {
          p = n - 1;
          q = r;
        } else {
          p = (r + i) % (n - 1);
          q = (r - i + (n - 1)) % (n - 1);
        }
        if (p >= N || q >= N) continue;
        T* wp = W + (size_t)p * N;
        T* wq = W + (size_t)q * N;
        T* qp = Q + (size_t)p * N;
        T* qq = Q + (size_t)q * N;
        if (THREADS <= 512 && N <= 32 * kJacobiRegs) {
          // rows staged in registers: all loads of a phase are in flight together (the four
          // row pointers alias as far as the compiler knows, so a load/store loop serialises)
          T ra[kJacobiRegs], rb[kJacobiRegs];
          T al = 0, be = 0, ga = 0;
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
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            al += __shfl_xor_sync(0xffffffffu, al, o);
            be += __shfl_xor_sync(0xffffffffu, be, o);
            ga += __shfl_xor_sync(0xffffffffu, ga, o);
          }
          if (fabs(ga) > tol * sqrt(al * be) && al > T(0) && be > T(0)) {
            if (lane == 0) off[sweep] = 1u;
            T rc[kJacobiRegs], rd[kJacobiRegs];
#pragma unroll
            for (int u = 0; u < kJacobiRegs; ++u) {
              const int k = lane + 32 * u;
              rc[u] = k < N ? qp[k] : T(0);
              rd[u] = k < N ? qq[k] : T(0);
            }
            const T zeta = (be - al) / (T(2) * ga);
            const T t = (zeta >= T(0) ? T(1) : T(-1)) / (fabs(zeta) + sqrt(T(1) + zeta * zeta));
            const T cs = T(1) / sqrt(T(1) + t * t), sn = cs * t;
#pragma unroll
            for (int u = 0; u < kJacobiRegs; ++u) {
              const int k = lane + 32 * u;
              if (k < N) {
                wp[k] = cs * ra[u] - sn * rb[u];
                wq[k] = sn * ra[u] + cs * rb[u];
                qp[k] = cs * rc[u] - sn * rd[u];
                qq[k] = sn * rc[u] + cs * rd[u];
              }
            }
          }
          continue;
        }
        T al = 0, be = 0, ga = 0;
        for (int k = lane; k < N; k += 32) {
          const T a = wp[k], b = wq[k];
          al += a * a;
          be += b * b;
          ga += a * b;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        if (fabs(ga) > tol * sqrt(al * be) && al > T(0) && be > T(0)) {
          if (lane == 0) off[sweep] = 1u;
          const T zeta = (be - al) / (T(2) * ga);
          const T t = (zeta >= T(0) ? T(1) : T(-1)) / (fabs(zeta) + sqrt(T(1) + zeta * zeta));
          const T cs = T(1) / sqrt(T(1) + t * t), sn = cs * t;
          for (int k = lane; k < N; k += 32) {
            const T a = wp[k], b = wq[k];
            wp[k] = cs * a - sn * b;
            wq[k] = sn * a + cs * b;
            const T c = qp[k], d = qq[k];
            qp[k] = cs * c - sn * d;
            qq[k] = sn * c + cs * d;
          }
        }
      }
      grid.sync();
    }
    if (*reinterpret_cast<volatile unsigned int*>(&off[sweep]) == 0u) break;
  }
  if (gtid == 0 && sweeps_out != nullptr) *sweeps_out = sweep + 1;

  for (int j = gwarp; j < N; j += gwarps) {  // signed eigenvalues
    T acc = 0;
    for (int k = lane; k < N; k += 32) acc += W[(size_t)j * N + k] * Q[(size_t)j * N + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) lam[j] = acc;
  }
  grid.sync();
  for (int j = gwarp; j < N; j += gwarps) {  // ascending stable rank, sign rule, scatter
    const T mine = lam[j];
    int rk = 0;
    for (int k = lane; k < N; k += 32) {
      const T o = lam[k];
      rk += (o < mine) || (o == mine && k < j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    T best = T(-1);
    int bidx = 0;
    for (int k = lane; k < N; k += 32) {
      const T a = fabs(Q[(size_t)j * N + k]);
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
    const T sgn = Q[(size_t)j * N + bidx] < T(0) ? T(-1) : T(1);
    if (lane == 0) w_out[rk] = mine;
    for (int k = lane; k < N; k += 32) B[(size_t)k * N + rk] = sgn * Q[(size_t)j * N + k];
  }
}

// Block Jacobi for matrices beyond one SM: the rows are cut into blocks of b rows; in an
// outer round every CTA of the (single) cluster loads one PAIR of blocks (2b rows of W and
// of Q) into shared memory, runs a complete inner round-robin sweep over those 2b rows
// there (b warps, one pair per warp, __syncthreads between inner rounds), and writes the
// rows back; outer rounds follow the same circle-method tournament over the blocks and are
// separated by the hardware cluster barrier.  Compared with one global barrier per scalar
// round this needs (2 N / b - 1) barriers per sweep instead of (N - 1) and keeps the
// rotations in shared memory.  A sweep whose largest |cos| was tiny ends the iteration
// without a separate all-quiet sweep (quadratic convergence).
__device__ __forceinline__ void circle_pair(int n, int r, int i, int* p, int* q) {
  if (i == 0) {
    *p = n - 1;
    *q = r;
  } else {
    *p = (r + i) % (n - 1);
    *q = (r - i + (n - 1)) % (n - 1);
  }
}

template <typename T>
__global__ void __launch_bounds__(512)
jacobi_block_kernel(T* __restrict__ C, int N, T* __restrict__ w_out, T* __restrict__ B, T* __restrict__ W,
                    T* __restrict__ Q, T* __restrict__ lam, unsigned int* __restrict__ off, int warm, int b, int nblk,
                    const int* __restrict__ gate, const int* __restrict__ status_gate, int* __restrict__ sweeps_out) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char jbs[];
  if (gate != nullptr && *gate == 0) return;
  if (status_gate != nullptr && *status_gate != SP_RUNNING) return;
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + tid;
  const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
  T* sW = reinterpret_cast<T*>(jbs);     // [2b][N]
  T* sQ = sW + (size_t)2 * b * N;        // [2b][N]

  for (int e = gtid; e < N * N; e += gthreads) {  // C = triu(C) + triu(C,1)^T (_cmaes.py:303)
    const int r = e / N, c = e - r * N;
    if (r > c) C[e] = C[c * N + r];
  }
  if (gtid < 64) off[gtid] = 0u;
  cl.sync();
  if (warm) {
    for (int e = gtid; e < N * N; e += gthreads) {
      const int j = e / N, r = e - j * N;
      Q[e] = B[r * N + j];
    }
    cl.sync();
    for (int e = gtid; e < N * N; e += gthreads) {  // W = Q C
      const int j = e / N, c = e - j * N;
      T acc = 0;
      for (int k = 0; k < N; ++k) acc += Q[j * N + k] * C[k * N + c];
      W[e] = acc;
    }
  } else {
    for (int e = gtid; e < N * N; e += gthreads) {
      const int r = e / N, c = e - r * N;
      W[e] = C[e];
      Q[e] = r == c ? T(1) : T(0);
    }
  }
  cl.sync();

  const T tol = T(4) * JacobiEps<T>::eps() * sqrt((T)(N < 16 ? 16 : N));
  const float quiet = 0.25f * sqrtf((float)tol);  // a sweep below this needs no follow-up sweep
  const int m = nblk / 2, n2 = 2 * b;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    for (int R = 0; R < nblk - 1; ++R) {
      for (int bp = blockIdx.x; bp < m; bp += gridDim.x) {
        int I, J;
        circle_pair(nblk, R, bp, &I, &J);
        // load the 2b rows (rows past N are zero)
        for (int e = tid; e < n2 * N; e += blockDim.x) {
          const int lr = e / N, k = e - lr * N;
          const int gr = (lr < b ? I * b + lr : J * b + (lr - b));
          sW[e] = gr < N ? W[(size_t)gr * N + k] : T(0);
          sQ[e] = gr < N ? Q[(size_t)gr * N + k] : T(0);
        }
        __syncthreads();
        for (int r = 0; r < n2 - 1; ++r) {
          if (warp < b) {
            int p, q;
            circle_pair(n2, r, warp, &p, &q);
            const int gp = (p < b ? I * b + p : J * b + (p - b)), gq = (q < b ? I * b + q : J * b + (q - b));
            if (gp < N && gq < N) {
              T* wp = sW + (size_t)p * N;
              T* wq = sW + (size_t)q * N;
              T al = 0, be = 0, ga = 0;
              for (int k = lane; k < N; k += 32) {
                const T a = wp[k], c = wq[k];
                al += a * a;
                be += c * c;
                ga += a * c;
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                al += __shfl_xor_sync(0xffffffffu, al, o);
                be += __shfl_xor_sync(0xffffffffu, be, o);
                ga += __shfl_xor_sync(0xffffffffu, ga, o);
              }
              const T nrm = sqrt(al * be);
              if (fabs(ga) > tol * nrm && al > T(0) && be > T(0)) {
                if (lane == 0) atomicMax(&off[sweep], __float_as_uint((float)(fabs(ga) / nrm)));
                const T zeta = (be - al) / (T(2) * ga);
                const T t = (zeta >= T(0) ? T(1) : T(-1)) / (fabs(zeta) + sqrt(T(1) + zeta * zeta));
                const T cs = T(1) / sqrt(T(1) + t * t), sn = cs * t;
                T* qp = sQ + (size_t)p * N;
                T* qq = sQ + (size_t)q * N;
                for (int k = lane; k < N; k += 32) {
                  const T a = wp[k], c = wq[k];
                  wp[k] = cs * a - sn * c;
                  wq[k] = sn * a + cs * c;
                  const T d = qp[k], f = qq[k];
                  qp[k] = cs * d - sn * f;
                  qq[k] = sn * d + cs * f;
                }
              }
            }
          }
          __syncthreads();
        }
        for (int e = tid; e < n2 * N; e += blockDim.x) {  // write the rows back
          const int lr = e / N, k = e - lr * N;
          const int gr = (lr < b ? I * b + lr : J * b + (lr - b));
          if (gr < N) {
            W[(size_t)gr * N + k] = sW[e];
            Q[(size_t)gr * N + k] = sQ[e];
          }
        }
        __syncthreads();
      }
      cl.sync();
    }
    const unsigned int worst = *reinterpret_cast<volatile unsigned int*>(&off[sweep]);
    if (worst == 0u || __uint_as_float(worst) < quiet) break;
  }
  if (gtid == 0 && sweeps_out != nullptr) *sweeps_out = sweep + 1;

  for (int j = gwarp; j < N; j += gwarps) {  // signed eigenvalues
    T acc = 0;
    for (int k = lane; k < N; k += 32) acc += W[(size_t)j * N + k] * Q[(size_t)j * N + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) lam[j] = acc;
  }
  cl.sync();
  for (int j = gwarp; j < N; j += gwarps) {  // ascending stable rank, sign rule, scatter
    const T mine = lam[j];
    int rk = 0;
    for (int k = lane; k < N; k += 32) {
      const T o = lam[k];
      rk += (o < mine) || (o == mine && k < j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    T best = T(-1);
    int bidx = 0;
    for (int k = lane; k < N; k += 32) {
      const T a = fabs(Q[(size_t)j * N + k]);
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
    const T sgn = Q[(size_t)j * N + bidx] < T(0) ? T(-1) : T(1);
    if (lane == 0) w_out[rk] = mine;
    for (int k = lane; k < N; k += 32) B[(size_t)k * N + rk] = sgn * Q[(size_t)j * N + k];
  }
}

// scratch scalars needed behind `work` for an N x N decomposition
inline size_t jacobi_work_scalars(int N) { return 2 * (size_t)N * N + (size_t)N + 64; }

template <typename T>
inline cudaError_t jacobi_launch(T* C, int N, T* w, T* B, T* work, int warm, const int* gate,
                                 const int* status_gate, int* sweeps, cudaStream_t s) {
  const size_t need = 2 * (size_t)N * N * sizeof(T);
  if (need <= 200 * 1024) {  // whole problem in one SM's shared memory
    auto kern = jacobi_eigh_kernel<T>;
    static thread_local bool configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
      configured[dev] = true;
    }
    kern<<<1, kJacobiThreads, need, s>>>(C, N, w, B, work, work + (size_t)N * N, 1, warm, gate, status_gate, sweeps);
    return cudaGetLastError();
  }
  T* W = work;
  T* Q = work + (size_t)N * N;
  T* lam = Q + (size_t)N * N;
  unsigned int* off = reinterpret_cast<unsigned int*>(lam + N);
  static const bool scalar_rounds = getenv("SP_EIGH_SCALAR") != nullptr;  // profiling switch
  if (!scalar_rounds) {
    int b = 16;
    while (b > 2 && 4 * (size_t)b * N * sizeof(T) > 200 * 1024) b >>= 1;
    int nblk = (N + b - 1) / b;
    nblk += nblk & 1;
    const int m = nblk / 2;
    const int ctas = m < 8 ? m : 8;
    const size_t smem = 4 * (size_t)b * N * sizeof(T);
    auto kern = jacobi_block_kernel<T>;
    static thread_local bool configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
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
    return cudaLaunchKernelEx(&cfg, kern, C, N, w, B, W, Q, lam, off, warm, b, nblk, gate, status_gate, sweeps);
  }
  const int pairs = (N + 1) / 2;
  // one cluster: 8 CTAs (portable) up to 512 pairs' worth of warps, 16 CTAs beyond
  int ctas = pairs <= 8 * 32 ? 8 : 16;
  int warps = (pairs + ctas - 1) / ctas;
  if (warps > 32) warps = 32;
  const int threads = warps * 32;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (threads <= 512) {
    auto kern = jacobi_grid_kernel<T, true, 512>;
    e = cudaLaunchKernelEx(&cfg, kern, C, N, w, B, W, Q, lam, off, warm, gate, status_gate, sweeps);
  } else {
    auto kern = jacobi_grid_kernel<T, true, 1024>;
    if (ctas > 8) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
    }
    e = cudaLaunchKernelEx(&cfg, kern, C, N, w, B, W, Q, lam, off, warm, gate, status_gate, sweeps);
  }
  return e;
}

}  // namespace sp
