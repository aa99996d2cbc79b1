// Shared device/host helpers of the stochopy_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <utility>

#include "../../include/stochopy_b200.h"

namespace sp {

// ---- host side -------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
int sm_count();

#define SP_CHECK_ARG(cond, msg)                      \
  do {                                               \
    if (!(cond)) {                                   \
      sp::set_error("%s: %s", __func__, msg);        \
      return SP_ERR_ARG;                             \
    }                                                \
  } while (0)

#define SP_CHECK_LAUNCH()                                                     \
  do {                                                                        \
    cudaError_t e_ = cudaGetLastError();                                      \
    if (e_ != cudaSuccess) {                                                  \
      sp::set_error("%s: CUDA launch failed: %s", __func__, cudaGetErrorString(e_)); \
      return SP_ERR_CUDA;                                                     \
    }                                                                         \
    sp::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
  } while (0)

constexpr int kThreads = 256;            // 8 warps per CTA
constexpr int kMaxBlocks = 148 * 16;     // scratch is sized for this many per-CTA minima
constexpr int kScratchBytes = kMaxBlocks * 16 + 256;

// ---- programmatic dependent launch ---------------------------------------------------
// A kernel launched with the programmatic-stream-serialization attribute may become resident
// while its predecessor drains; it must not read what the predecessor wrote before pdl_wait().
// Both are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- numeric traits ----------------------------------------------------------
template <typename T> struct Num;
template <> struct Num<float> {
  static constexpr int VEC = 4;
  using vec_t = float4;
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <> struct Num<double> {
  static constexpr int VEC = 2;
  using vec_t = double2;
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

// Round-to-nearest ops that ptxas may not contract into FMAs: the update
// formulas mirror numpy's separate multiply and add so trajectories stay on the
// reference's path; objective evaluation is free to use FMA.
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// ---- row tiles -----------------------------------------------------------------
// A population row of N scalars is held by LPR lanes, CH chunks of VEC scalars
// per lane (16-byte vector loads).  Column of element (c, e) on lane l:
//   j = (c * LPR + l) * VEC + e
// A warp carries 32 / LPR rows at once.
template <typename T, int CH, int LPR>
struct Tile {
  static constexpr int VEC = Num<T>::VEC;
  static constexpr int RPW = 32 / LPR;           // rows per warp
  static constexpr int COLS = CH * LPR * VEC;    // widest row this shape holds
  T v[CH][VEC];

  static __device__ __forceinline__ int col(int c, int l, int e) { return (c * LPR + l) * VEC + e; }

  // rows are padded to ld (multiple of VEC) so whole vectors are always in bounds
  __device__ __forceinline__ void load(const T* __restrict__ row, int l, int ld) {
    using V = typename Num<T>::vec_t;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int j0 = col(c, l, 0);
      if (j0 < ld) {
        V t = *reinterpret_cast<const V*>(row + j0);
        const T* p = reinterpret_cast<const T*>(&t);
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[c][e] = p[e];
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[c][e] = T(0);
      }
    }
  }
  // streaming variant (st.global.cs, evict-first): for arrays larger than the L2 that are not re-read soon
  __device__ __forceinline__ void store_cs(T* __restrict__ row, int l, int ld) const {
    using V = typename Num<T>::vec_t;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int j0 = col(c, l, 0);
      if (j0 < ld) {
        V t;
        T* p = reinterpret_cast<T*>(&t);
#pragma unroll
        for (int e = 0; e < VEC; ++e) p[e] = v[c][e];
        __stcs(reinterpret_cast<V*>(row + j0), t);
      }
    }
  }
  __device__ __forceinline__ void store(T* __restrict__ row, int l, int ld) const {
    using V = typename Num<T>::vec_t;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int j0 = col(c, l, 0);
      if (j0 < ld) {
        V t;
        T* p = reinterpret_cast<T*>(&t);
#pragma unroll
        for (int e = 0; e < VEC; ++e) p[e] = v[c][e];
        *reinterpret_cast<V*>(row + j0) = t;
      }
    }
  }
};

// reductions over the LPR lanes that share a row (xor butterflies stay inside
// an aligned group of LPR lanes)
template <int LPR, typename T>
__device__ __forceinline__ T group_sum(T x) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
template <int LPR, typename T>
__device__ __forceinline__ T group_prod(T x) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) x *= __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
template <int LPR, typename T>
__device__ __forceinline__ T group_min(T x) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    T y = __shfl_xor_sync(0xffffffffu, x, o);
    x = y < x ? y : x;
  }
  return x;
}
template <int LPR, typename T>
__device__ __forceinline__ T group_max(T x) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    T y = __shfl_xor_sync(0xffffffffu, x, o);
    x = y > x ? y : x;
  }
  return x;
}

// (fitness, row) pairs ordered like np.argmin: smaller value, then smaller row
struct Best {
  double f;
  long long row;
};
__device__ __forceinline__ bool better(double f, long long r, double g, long long s) {
  return f < g || (f == g && r < s);
}
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double f = __shfl_xor_sync(0xffffffffu, b.f, o);
    long long r = __shfl_xor_sync(0xffffffffu, b.row, o);
    if (better(f, r, b.f, b.row)) {
      b.f = f;
      b.row = r;
    }
  }
  return b;
}

// minimum over the CTA (valid in thread 0 and, after the barrier, in s_best[0..warps))
__device__ __forceinline__ Best block_best(Best mine, Best* s_best /* [32] shared */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  mine = warp_best(mine);
  if (lane == 0) s_best[warp] = mine;
  __syncthreads();
  Best b{1.0 / 0.0, 0x7fffffffffffffffLL};
  if (warp == 0) {
    if (lane < (int)(blockDim.x >> 5)) b = s_best[lane];
    b = warp_best(b);
  }
  return b;
}

// Per-CTA minimum -> scratch[blockIdx]; the last CTA to arrive reduces all of
// them.  Returns true (for every thread of that last CTA) with the global best.
__device__ __forceinline__ bool grid_best(Best mine, Best* scratch, sp_ctrl* ctrl, Best* out) {
  __shared__ Best s_best[32];  // up to 1024 threads
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Best b0 = block_best(mine, s_best);
  if (warp == 0 && lane == 0) {
    scratch[blockIdx.x] = b0;
    __threadfence();
    unsigned prev = atomicAdd(&ctrl->done_blocks, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  Best b{1.0 / 0.0, 0x7fffffffffffffffLL};
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    Best o{__ldcg(&scratch[i].f), __ldcg(&scratch[i].row)};
    if (better(o.f, o.row, b.f, b.row)) b = o;
  }
  b = warp_best(b);
  __syncthreads();
  if (lane == 0) s_best[warp] = b;
  __syncthreads();
  b = s_best[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (better(s_best[w].f, s_best[w].row, b.f, b.row)) b = s_best[w];
  *out = b;
  if (threadIdx.x == 0) ctrl->done_blocks = 0;  // ready for the next launch
  return true;
}

// ---- chained generations ---------------------------------------------------------
// A generation launched with SP_CHAIN_OUT skips the last-CTA epilogue: every CTA only
// leaves its (fitness, row) minimum in a scratch region picked by the generation's
// parity.  The next launch (SP_CHAIN_IN) reduces those <= 148 records in every CTA's
// prologue -- the stop decision needs only the best fitness and the generation count --
// and one warp of CTA 0 writes gbest / dist / nit / status for the generation before.
// The serial tail (fence, atomic, scratch read, row read: ~9k cycles on one SM while
// 147 idle) leaves the critical path; the scratch regions are disjoint from region 0,
// which grid_best() uses.
constexpr int kChainRegion = 768;  // records per region; regions 1 and 2 (parity): 3 x 768 <= kMaxBlocks
__device__ __forceinline__ Best* chain_region(Best* scratch, int it) { return scratch + (1 + (it & 1)) * kChainRegion; }

// block-level variant for grids of up to kChainRegion CTAs: best record and ITS INDEX (the CTA
// that wrote it -- where that CTA left the winning row); valid in every thread after the call
__device__ __forceinline__ Best chain_best_block(const Best* region, int n, int* index) {
  __shared__ Best s_b[32];
  __shared__ int s_i[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  Best b{1.0 / 0.0, 0x7fffffffffffffffLL};
  int idx = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    Best o{__ldcg(&region[i].f), __ldcg(&region[i].row)};
    if (better(o.f, o.row, b.f, b.row)) {
      b = o;
      idx = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double f = __shfl_xor_sync(0xffffffffu, b.f, o);
    const long long r = __shfl_xor_sync(0xffffffffu, b.row, o);
    const int i = __shfl_xor_sync(0xffffffffu, idx, o);
    if (better(f, r, b.f, b.row)) {
      b.f = f;
      b.row = r;
      idx = i;
    }
  }
  if (lane == 0) {
    s_b[warp] = b;
    s_i[warp] = idx;
  }
  __syncthreads();
  b = s_b[0];
  idx = s_i[0];
  for (int w = 1; w < nw; ++w)
    if (better(s_b[w].f, s_b[w].row, b.f, b.row)) {
      b = s_b[w];
      idx = s_i[w];
    }
  *index = idx;
  return b;
}

// warp-level read of the previous generation's per-CTA minima (valid in every lane)
__device__ __forceinline__ Best chain_best(const Best* region, int n) {
  const int lane = threadIdx.x & 31;
  Best b{1.0 / 0.0, 0x7fffffffffffffffLL};
  for (int i = lane; i < n; i += 32) {
    Best o{__ldcg(&region[i].f), __ldcg(&region[i].row)};
    if (better(o.f, o.row, b.f, b.row)) b = o;
  }
  return warp_best(b);
}

// Last-CTA epilogue of every generation: new best row -> gbest, distance to
// the previous best, status ladder of selection_sync (_common.py:135-158).
// `it` < 0: initial population (no status test).
// VOL: the winning row was stored by a peer GPU into this rank's mailbox -> volatile loads.
template <typename T, bool VOL = false>
__device__ __forceinline__ void finalize_generation(Best b, const T* xrows, int64_t ld, int N,
                                                    T* gbest, sp_ctrl* ctrl, int it, int maxiter, double xtol,
                                                    double ftol) {
  __shared__ double s_part[32];
  const T* src = xrows + b.row * ld;
  double acc = 0.0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    T nv = VOL ? *reinterpret_cast<const volatile T*>(src + j) : src[j];
    double d = (double)(T)(gbest[j] - nv);
    acc += d * d;
    gbest[j] = nv;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
    double dist = sqrt(tot);
    ctrl->gbest_row = b.row;
    ctrl->gfit = b.f;
    ctrl->dist = dist;
    if (it >= 0) {
      ctrl->nit = it;
      int st = SP_RUNNING;
      if (dist <= xtol && b.f <= ftol) st = 0;
      else if (b.f <= ftol) st = 1;
      else if (it >= maxiter) st = -1;
      ctrl->status = st;
    }
  }
}

// the same epilogue executed by ONE warp (chained generations: a warp of CTA 0 resolves the
// generation before while the other warps already work on rows)
template <typename T>
__device__ __forceinline__ void finalize_generation_warp(Best b, const T* __restrict__ xrows, int64_t ld, int N,
                                                         T* gbest, sp_ctrl* ctrl, int it, int maxiter, double xtol,
                                                         double ftol) {
  const int lane = threadIdx.x & 31;
  const T* src = xrows + b.row * ld;
  double acc = 0.0;
  for (int j = lane; j < N; j += 32) {
    T nv = src[j];
    double d = (double)(T)(gbest[j] - nv);
    acc += d * d;
    gbest[j] = nv;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    double dist = sqrt(acc);
    ctrl->gbest_row = b.row;
    ctrl->gfit = b.f;
    ctrl->dist = dist;
    ctrl->nit = it;
    int st = SP_RUNNING;
    if (dist <= xtol && b.f <= ftol) st = 0;
    else if (b.f <= ftol) st = 1;
    else if (it >= maxiter) st = -1;
    ctrl->status = st;
  }
}
__device__ __forceinline__ bool chain_stops(double f, int it, int maxiter, double ftol) {
  return f <= ftol || it >= maxiter;  // status 0 / 1 / -1 of _common.py:135-158: all of them stop
}

__device__ __forceinline__ bool running(const sp_ctrl* ctrl) {
  return *reinterpret_cast<const volatile int32_t*>(&ctrl->status) == SP_RUNNING;
}

// ---- launch shape ------------------------------------------------------------
// Host-side choice of (CH, LPR) for a row of N scalars of type T; returns false
// if N exceeds the compiled shapes.
struct Shape {
  int ch, lpr;
};
inline bool pick_shape(int N, int vec, Shape* s) {
  const int lprs[] = {1, 4, 16, 32};
  for (int l : lprs)
    if (l * vec >= N) {
      *s = {1, l};
      return true;
    }
  const int chs[] = {2, 4, 8, 16};
  for (int c : chs)
    if (c * 32 * vec >= N) {
      *s = {c, 32};
      return true;
    }
  return false;
}
inline int grid_for_rows(int64_t P, int lpr, int blocks_per_sm) {
  int64_t rows_per_block = (int64_t)(kThreads / 32) * (32 / lpr);
  int64_t need = (P + rows_per_block - 1) / rows_per_block;
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  int64_t g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}

// dispatch a functor templated on <T, CH, LPR>
#define SP_DISPATCH_SHAPE(T, shape, CALL)                                   \
  do {                                                                      \
    if ((shape).lpr == 1) { CALL(T, 1, 1); }                                \
    else if ((shape).lpr == 4) { CALL(T, 1, 4); }                           \
    else if ((shape).lpr == 16) { CALL(T, 1, 16); }                         \
    else if ((shape).ch == 1) { CALL(T, 1, 32); }                           \
    else if ((shape).ch == 2) { CALL(T, 2, 32); }                           \
    else if ((shape).ch == 4) { CALL(T, 4, 32); }                           \
    else if ((shape).ch == 8) { CALL(T, 8, 32); }                           \
    else { CALL(T, 16, 32); }                                               \
  } while (0)

}  // namespace sp
