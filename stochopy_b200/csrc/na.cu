// Neighbourhood Algorithm: Gibbs walk inside Voronoi cells of the best archived models.
// Reference: stochopy/optimize/na/_na.py:265-305 (mutation), after
// github.com/keithfma/neighborhood.  Cost per generation O(P * M * N), M = archive size.
#include "rows.cuh"

namespace sp {

template <typename T>
__global__ void na_append_kernel(T* __restrict__ archT, int64_t cap, int64_t M, const T* __restrict__ X, int64_t P,
                                 int N, int64_t ld) {
  const int64_t total = P * (int64_t)N;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(t / P);
    const int64_t i = t - (int64_t)j * P;
    archT[(int64_t)j * cap + M + i] = X[i * ld + j];
  }
}

__global__ void na_cells_kernel(const int32_t* __restrict__ rank, int64_t M, int32_t nr, int32_t* __restrict__ cells) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x)
    if (rank[m] < nr) cells[rank[m]] = (int32_t)m;
}

template <typename T>
__global__ void __launch_bounds__(256)
na_resample_kernel(const T* __restrict__ archT, int64_t cap, int64_t M, const int32_t* __restrict__ cells, int32_t nr,
                   T* __restrict__ X, int N, int64_t ld, const int32_t* __restrict__ mask, const T* __restrict__ u,
                   uint64_t seed, int it, T* __restrict__ work) {
  constexpr int VEC = Num<T>::VEC;
  extern __shared__ __align__(16) unsigned char na_sm[];
  T* x = reinterpret_cast<T*>(na_sm);  // walker (N)
  T* ck = x + N;                       // cell centre (N)
  __shared__ T s_lo[8], s_hi[8];
  __shared__ T s_new;
  const int64_t i = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t k = cells[i % nr];
  T* d2 = work + i * M;
  for (int j = tid; j < N; j += blockDim.x) {
    const T c = archT[(int64_t)j * cap + k];
    ck[j] = c;
    x[j] = c;
  }
  __syncthreads();
  // d2[m] = sum_{j >= 1} (U[m, j] - x[j])^2   (_na.py:283)
  for (int64_t m = tid; m < M; m += blockDim.x) {
    T acc = 0;
    for (int j = 1; j < N; ++j) {
      const T d = archT[(int64_t)j * cap + m] - x[j];
      acc += d * d;
    }
    d2[m] = acc;
  }
  T d1 = 0;
  for (int j = 0; j < N; ++j) {
    if (!mask[j]) {  // zero-span axis: value fixed by unnormalize, no distance update (_na.py:286-289)
      if (tid == 0) x[j] = T(0);
      __syncthreads();
      continue;
    }
    const T xj = x[j], cj = ck[j];
    T lo = -Num<T>::inf(), hi = Num<T>::inf();
    const T* col = archT + (int64_t)j * cap;
    for (int64_t m = tid; m < M; m += blockDim.x) {
      if (m == k) continue;
      const T um = col[m];
      const T lim = T(0.5) * (cj + um + (d1 - d2[m]) / (cj - um));  // _na.py:291
      if (lim <= xj) lo = lim > lo ? lim : lo;
      if (lim >= xj) hi = lim < hi ? lim : hi;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const T a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = a > lo ? a : lo;
      hi = b < hi ? b : hi;
    }
    if ((tid & 31) == 0) {
      s_lo[tid >> 5] = lo;
      s_hi[tid >> 5] = hi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w) {
        lo = s_lo[w] > lo ? s_lo[w] : lo;
        hi = s_hi[w] < hi ? s_hi[w] : hi;
      }
      const T low = lo == -Num<T>::inf() ? T(0) : (lo > T(0) ? lo : T(0));   // _na.py:293-297
      const T high = hi == Num<T>::inf() ? T(1) : (hi < T(1) ? hi : T(1));
      T r;
      if (u != nullptr) {
        r = u[i * ld + j];
      } else {
        T blk[VEC];
        uniform_block(philox4x32((uint32_t)(j / VEC), (uint32_t)i, (uint32_t)it, kNaWalk, seed), blk);
        r = blk[j % VEC];
      }
      s_new = add_rn(low, mul_rn(sub_rn(high, low), r));  // uniform(low, high), _na.py:299
    }
    __syncthreads();
    const T xn = s_new;
    if (j < N - 1) {  // _na.py:301-303 (x[j+1] is still the cell centre's coordinate)
      const T xn1 = x[j + 1];
      const T a = cj - xn, b = ck[j + 1] - xn1;
      d1 += a * a - b * b;
      const T* col1 = archT + (int64_t)(j + 1) * cap;
      for (int64_t m = tid; m < M; m += blockDim.x) {
        const T p = col[m] - xn, q = col1[m] - xn1;
        d2[m] += p * p - q * q;
      }
    }
    __syncthreads();
    if (tid == 0) x[j] = xn;
    __syncthreads();
  }
  for (int j = tid; j < N; j += blockDim.x) X[i * ld + j] = x[j];
}

}  // namespace sp

using namespace sp;

extern "C" {

int sp_na_append(int dtype, void* archT, int64_t cap, int64_t M, const void* X, int64_t P, int N, int64_t ld,
                 void* stream) {
  SP_CHECK_ARG(archT && X && P >= 1 && N >= 1 && M >= 0 && M + P <= cap && ld >= N, "null pointer or archive overflow");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t need = (P * (int64_t)N + 255) / 256, capb = (int64_t)sm_count() * 8;
  const int grid = (int)(need < capb ? need : capb);
  if (dtype == SP_F32) na_append_kernel<float><<<grid, 256, 0, s>>>((float*)archT, cap, M, (const float*)X, P, N, ld);
  else na_append_kernel<double><<<grid, 256, 0, s>>>((double*)archT, cap, M, (const double*)X, P, N, ld);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_na_cells(const int32_t* rank, int64_t M, int32_t nr, int32_t* cells, void* stream) {
  SP_CHECK_ARG(rank && cells && M >= 1 && nr >= 1 && nr <= M, "null pointer or nr outside [1, M]");
  const int64_t need = (M + 255) / 256;
  na_cells_kernel<<<(int)(need < 1024 ? need : 1024), 256, 0, (cudaStream_t)stream>>>(rank, M, nr, cells);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_na_resample(int dtype, const void* archT, int64_t cap, int64_t M, const int32_t* cells, int32_t nr, void* X,
                   int64_t P, int N, int64_t ld, const int32_t* mask, const void* u, uint64_t seed, int it, void* work,
                   void* stream) {
  SP_CHECK_ARG(archT && cells && X && mask && work && P >= 1 && N >= 1 && M >= 2 && M <= cap && nr >= 1 && ld >= N,
               "null pointer or bad shape");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = 2 * (size_t)N * (dtype == SP_F32 ? 4 : 8);
  SP_CHECK_ARG(smem <= 48 * 1024, "ndim too large for the walker's shared memory");
  if (dtype == SP_F32)
    na_resample_kernel<float><<<(int)P, 256, smem, s>>>((const float*)archT, cap, M, cells, nr, (float*)X, N, ld, mask,
                                                        (const float*)u, seed, it, (float*)work);
  else
    na_resample_kernel<double><<<(int)P, 256, smem, s>>>((const double*)archT, cap, M, cells, nr, (double*)X, N, ld,
                                                         mask, (const double*)u, seed, it, (double*)work);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

}  // extern "C"
