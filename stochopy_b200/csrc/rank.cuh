// Ascending stable rank of a fitness vector: rank[i] = position of i in
// np.argsort(f, kind="stable") (ties by index; NaN last like numpy; -0 == +0).
//
// Items are (order-preserving key bits, index) pairs, unique by construction, so the
// rank of an item is the number of items smaller than it:
//   rank_sort_kernel   each CTA bitonic-sorts a chunk of up to 512 items (registers,
//                      warp shuffles, shared memory only for the widest exchanges); a
//                      population that fits one chunk gets its ranks here;
//   rank_merge_kernel  an item's rank = its position in its own chunk + its lower
//                      bound in every other sorted chunk (the searched chunk is staged
//                      in shared memory; integer atomics on distinct addresses).
// O(P log^2 c + P (P/c) log c) work instead of the P^2 of a counting rank: 32768 keys
// take microseconds, which is what the CPSO restart (cpso/_cpso.py:420), the (mu, lambda)
// weights (cmaes/_cmaes.py:272, vdcma/_vdcma.py:290) and NA's best cells (na/_na.py) need.
// `gate` (optional): every kernel returns immediately unless *gate > 0; `live` (optional): ... unless
// *live == SP_RUNNING (the control block's status: generations enqueued behind a stop do nothing).
#pragma once
#include "common.cuh"

namespace sp {

constexpr int kRankChunk = 512;  // one chunk = one CTA of 256 threads; short stage chains, many SMs busy

template <typename T>
struct RankItem;
template <>
struct RankItem<float> {
  using type = unsigned long long;  // key bits high, index low: one integer compare
  static __device__ __forceinline__ type make(float f, uint32_t i) {
    uint32_t b = __float_as_uint(f + 0.0f);  // -0 -> +0
    b ^= (b >> 31) ? 0xFFFFFFFFu : 0x80000000u;
    return ((unsigned long long)b << 32) | i;
  }
  static __device__ __forceinline__ type pad() { return ~0ull; }
  static __device__ __forceinline__ bool less(type a, type b) { return a < b; }
  static __device__ __forceinline__ uint32_t index(type a) { return (uint32_t)a; }
};
struct __align__(16) RankItem64 {
  unsigned long long k;
  uint32_t i, unused;
};
template <>
struct RankItem<double> {
  using type = RankItem64;
  static __device__ __forceinline__ type make(double f, uint32_t i) {
    unsigned long long b = (unsigned long long)__double_as_longlong(f + 0.0);
    b ^= (b >> 63) ? ~0ull : 0x8000000000000000ull;
    return type{b, i, 0u};
  }
  static __device__ __forceinline__ type pad() { return type{~0ull, 0xFFFFFFFFu, 0u}; }
  static __device__ __forceinline__ bool less(const type& a, const type& b) {
    return a.k < b.k || (a.k == b.k && a.i < b.i);
  }
  static __device__ __forceinline__ uint32_t index(const type& a) { return a.i; }
};

// butterfly exchange of an item between lanes
__device__ __forceinline__ unsigned long long rank_shfl(unsigned long long v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ RankItem64 rank_shfl(const RankItem64& v, int m) {
  RankItem64 o;
  o.k = __shfl_xor_sync(0xffffffffu, v.k, m);
  o.i = __shfl_xor_sync(0xffffffffu, v.i, m);
  o.unused = 0u;
  return o;
}

// Bitonic sort of one chunk; thread t owns items 2t and 2t+1.  Compare distance 1 stays in
// the thread, distances 2..32 are warp shuffles, only distances >= 64 go through shared memory
// (6 of the 45 stages of a 512-item chunk).
template <typename T>
__global__ void __launch_bounds__(kRankChunk / 2)
rank_sort_kernel(const T* __restrict__ fit, int64_t P, int n, typename RankItem<T>::type* __restrict__ sorted,
                 int32_t* __restrict__ rank, const int32_t* gate, int single, const int32_t* live = nullptr) {
  using R = RankItem<T>;
  using I = typename R::type;
  pdl_launch_dependents();  // no-ops unless launched with the programmatic-serialization attribute
  pdl_wait();
  if (gate != nullptr && *gate <= 0) return;
  if (live != nullptr && *reinterpret_cast<const volatile int32_t*>(live) != SP_RUNNING) return;
  __shared__ I s[kRankChunk];
  const int t = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * n;
  const int i0 = 2 * t, i1 = 2 * t + 1;
  I e0 = (i0 < n && base + i0 < P) ? R::make(fit[base + i0], (uint32_t)(base + i0)) : R::pad();
  I e1 = (i1 < n && base + i1 < P) ? R::make(fit[base + i1], (uint32_t)(base + i1)) : R::pad();
  auto take = [](I& mine, const I& other, bool keep_min) {
    if (keep_min ? R::less(other, mine) : R::less(mine, other)) mine = other;
  };
  for (int k = 2; k <= n; k <<= 1) {
    const bool asc = (i0 & k) == 0;  // same for both items of the thread (k >= 2)
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 1) {
        const bool sw = R::less(e1, e0) == asc;
        if (sw) {
          const I tmp = e0;
          e0 = e1;
          e1 = tmp;
        }
      } else if (j <= 32) {
        const bool keep_min = ((i0 & j) == 0) == asc;
        const I o0 = rank_shfl(e0, j >> 1), o1 = rank_shfl(e1, j >> 1);
        take(e0, o0, keep_min);
        take(e1, o1, keep_min);
      } else {
        const bool keep_min = ((i0 & j) == 0) == asc;
        __syncthreads();
        if (i1 < n) {
          s[i0] = e0;
          s[i1] = e1;
        }
        __syncthreads();
        if (i1 < n) {
          const I o0 = s[i0 ^ j], o1 = s[i1 ^ j];
          take(e0, o0, keep_min);
          take(e1, o1, keep_min);
        }
      }
    }
  }
  if (i1 < n) {
    if (single) {
      const uint32_t a = R::index(e0), b = R::index(e1);
      if (a < (uint64_t)P) rank[a] = i0;
      if (b < (uint64_t)P) rank[b] = i1;
    } else {  // own-chunk position now, the other chunks are added by rank_merge_kernel
      sorted[base + i0] = e0;
      sorted[base + i1] = e1;
      const uint32_t a = R::index(e0), b = R::index(e1);
      if (a < (uint64_t)P) rank[a] = i0;
      if (b < (uint64_t)P) rank[b] = i1;
    }
  }
}

// rank[i] += sum over the chunks c of this CTA's group (c = g, g + G, ..., own chunk
// excluded) of the lower bound of item i in chunk c.  The CTA holds one sorted chunk of
// items in registers (one per thread) and walks its group of chunks through a
// double-buffered shared-memory stage: the next chunk is in flight while the current one
// is searched, and each item costs one atomic per CTA instead of one per chunk.
template <typename T>
__global__ void __launch_bounds__(kRankChunk)
rank_merge_kernel(const typename RankItem<T>::type* __restrict__ sorted, int64_t P, int C,
                  int32_t* __restrict__ rank, const int32_t* gate, const int32_t* live = nullptr, int ab0 = 0) {
  using R = RankItem<T>;
  using I = typename R::type;
  constexpr int n = kRankChunk;
  pdl_launch_dependents();
  pdl_wait();
  if (gate != nullptr && *gate <= 0) return;
  if (live != nullptr && *reinterpret_cast<const volatile int32_t*>(live) != SP_RUNNING) return;
  __shared__ I s[2][n];
  const int ab = ab0 + blockIdx.x, G = gridDim.y, t = threadIdx.x;  // ab0: first chunk whose items are ranked
  const I me = sorted[(int64_t)ab * n + t];
  const uint32_t idx = R::index(me);
  const bool real = idx < (uint64_t)P;
  int c = blockIdx.y;
  if (c == ab) c += G;
  I reg = R::pad();
  if (c < C) reg = sorted[(int64_t)c * n + t];
  int buf = 0, cnt = 0;
  while (c < C) {
    s[buf][t] = reg;
    __syncthreads();
    int cn = c + G;
    if (cn == ab) cn += G;
    if (cn < C) reg = sorted[(int64_t)cn * n + t];
    if (real) {
      const I* __restrict__ ch = s[buf];
      int lo = 0;
#pragma unroll
      for (int h = n >> 1; h > 0; h >>= 1) lo += R::less(ch[lo + h - 1], me) ? h : 0;
      lo += R::less(ch[lo], me) ? 1 : 0;
      cnt += lo;
    }
    buf ^= 1;
    c = cn;
  }
  if (real && cnt) atomicAdd(&rank[idx], cnt);
}

// stream-ordered scratch from the device's default pool (kept warm: no trim at syncs)
inline cudaError_t rank_scratch(void** p, size_t bytes, cudaStream_t s) {
  static thread_local bool tuned[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !tuned[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    tuned[dev] = true;
  }
  return cudaMallocAsync(p, bytes, s);
}

// bytes of caller-provided scratch for rank_launch_ws (16-byte aligned)
inline size_t rank_ws_bytes(int64_t P) { return (size_t)((P + kRankChunk - 1) / kRankChunk) * kRankChunk * 16; }

// the same with caller-provided scratch and programmatic dependent launches: no pool traffic between
// the kernels of a generation chain, and each kernel is scheduled while its predecessor drains
template <typename T>
inline cudaError_t rank_launch_ws(const T* fit, int64_t P, int32_t* rank, void* ws, cudaStream_t s, const int32_t* live) {
  using I = typename RankItem<T>::type;
  const int32_t* gate = nullptr;
  if (P <= kRankChunk) {
    int n = 2;
    while (n < P) n <<= 1;
    int threads = n >> 1;
    threads = threads < 32 ? 32 : threads;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_pdl(rank_sort_kernel<T>, dim3(1), dim3(threads), 0, s, true, fit, P, n, (I*)nullptr, rank, gate, 1, live);
  }
  const int n = kRankChunk;
  const int C = (int)((P + n - 1) / n);
  cudaError_t e = launch_pdl(rank_sort_kernel<T>, dim3(C), dim3(n / 2), 0, s, true, fit, P, n, (I*)ws, rank, gate, 0, live);
  if (e != cudaSuccess) return e;
  int G = (3 * sm_count() + C - 1) / C;  // about three CTAs per SM in all
  G = G < 1 ? 1 : (G > C ? C : G);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return launch_pdl(rank_merge_kernel<T>, dim3((unsigned)C, (unsigned)G), dim3(n), 0, s, true, (const I*)ws, P, C, rank, gate, live, 0);
}

// first / count (optional): only the items first .. first + count - 1 need their rank (a rank of a row-sharded
// swarm ranks its own rows against the whole all-gathered vector): every chunk is sorted, but only the chunks
// that overlap the range are merged against the others -- 1 / world of the merge work.  rank[] of the other
// items is left incomplete.
template <typename T>
inline cudaError_t rank_launch(const T* fit, int64_t P, int32_t* rank, const int32_t* gate, cudaStream_t s,
                               const int32_t* live = nullptr, int64_t first = 0, int64_t count = -1) {
  using I = typename RankItem<T>::type;
  if (P <= kRankChunk) {
    int n = 2;
    while (n < P) n <<= 1;
    int threads = n >> 1;
    threads = threads < 32 ? 32 : threads;
    rank_sort_kernel<T><<<1, threads, 0, s>>>(fit, P, n, nullptr, rank, gate, 1, live);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
  }
  const int n = kRankChunk;
  const int C = (int)((P + n - 1) / n);
  void* ws = nullptr;
  cudaError_t e = rank_scratch(&ws, (size_t)C * n * sizeof(I), s);
  if (e != cudaSuccess) return e;
  rank_sort_kernel<T><<<C, n / 2, 0, s>>>(fit, P, n, (I*)ws, rank, gate, 0, live);
  int ab0 = 0, Cm = C;
  if (count >= 0) {
    ab0 = (int)(first / n);
    Cm = (int)((first + (count > 0 ? count : 1) - 1) / n) - ab0 + 1;
    if (ab0 + Cm > C) Cm = C - ab0;
  }
  int G = (3 * sm_count() + Cm - 1) / Cm;  // about three CTAs per SM in all
  G = G < 1 ? 1 : (G > C ? C : G);
  rank_merge_kernel<T><<<dim3((unsigned)Cm, (unsigned)G), n, 0, s>>>((const I*)ws, P, C, rank, gate, live, ab0);
  e = cudaGetLastError();
  g_launches.fetch_add(2, std::memory_order_relaxed);
  cudaError_t f = cudaFreeAsync(ws, s);
  return e != cudaSuccess ? e : f;
}

}  // namespace sp
