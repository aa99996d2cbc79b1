// Library plumbing + the method-independent kernels:
//   sp_eval        a1/a2  batched objective
//   sp_lhs_init    a3     Latin hypercube
//   sp_select_sync a4     greedy selection + argmin + termination
//   sp_best_init          first argmin of a fresh population
#include <cstdarg>
#include <cstring>

#include "linalg.cuh"
#include "rows.cuh"

namespace sp {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- a3 -----------------------------------------------------------------------
// X[i,j] = (u[p,j]/P + (-1 + 2p/P)) * 0.5(up-lo) + 0.5(up+lo),  p = perm_j(i)
// (_common.py:109-120; jitter 1/P wide on strata 2/P apart -- kept as is)
template <typename T>
__global__ void lhs_kernel(T* __restrict__ X, int64_t P, int N, int64_t ld, const T* __restrict__ lower,
                           const T* __restrict__ upper, uint64_t seed, const T* __restrict__ jitter,
                           const int64_t* __restrict__ perm, int64_t Ptot, int64_t row0) {
  constexpr int VEC = Num<T>::VEC;
  const int64_t total = P * (int64_t)N;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / N;
    const int j = (int)(t - i * N);
    int64_t p;
    T u;
    if (perm != nullptr) {
      p = perm[(int64_t)j * P + i];
      u = jitter[p * ld + j];
    } else {
      p = lhs_permute((uint32_t)(row0 + i), (uint32_t)Ptot, (uint32_t)j, seed);
      T blk[VEC];
      uniform_block(philox4x32((uint32_t)(j / VEC), (uint32_t)p, 0u, kLhsJitter, seed), blk);
      u = blk[j % VEC];
    }
    // linspace(-1, 1, P, endpoint=False)[p] = -1 + p * (2/P)
    T cell = add_rn(div_rn(u, (T)Ptot), (T)__dadd_rn(__dmul_rn((double)p, 2.0 / (double)Ptot), -1.0));
    T half = mul_rn((T)0.5, sub_rn(upper[j], lower[j]));
    T mid = mul_rn((T)0.5, add_rn(upper[j], lower[j]));
    X[i * ld + j] = add_rn(mul_rn(cell, half), mid);
  }
}

// ---- a4 -----------------------------------------------------------------------
template <typename T, int CH, int LPR>
__global__ void __launch_bounds__(kThreads)
select_kernel(int it, int maxiter, double xtol, double ftol, const T* __restrict__ cand,
              const T* __restrict__ candfun, T* __restrict__ x, T* __restrict__ xfun, int64_t P, int N, int64_t ld,
              int copy_when, T* gbest, sp_ctrl* ctrl, Best* scratch) {
  using TL = Tile<T, CH, LPR>;
  if (!running(ctrl)) return;
  const int lane = threadIdx.x & 31, l = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t groups = (P + TL::RPW - 1) / TL::RPW;
  Best mine{1.0 / 0.0, 0x7fffffffffffffffLL};
  for (int64_t g = warp; g < groups; g += nwarps) {
    int64_t row = g * TL::RPW + sub;
    if (row >= P) continue;
    T cf = candfun[row], xf = xfun[row];
    const bool win = cf < xf;  // strict, _common.py:127
    if (win == (copy_when != 0)) {
      TL t;
      t.load(cand + row * ld, l, (int)ld);
      t.store(x + row * ld, l, (int)ld);
    }
    if (win) xf = cf;
    if (l == 0) {
      xfun[row] = xf;
      if (better((double)xf, row, mine.f, mine.row)) mine = Best{(double)xf, row};
    }
  }
  Best top;
  if (grid_best(mine, scratch, ctrl, &top)) finalize_generation<T>(top, x, ld, N, gbest, ctrl, it, maxiter, xtol, ftol);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
best_init_kernel(const T* __restrict__ x, const T* __restrict__ xfun, int64_t P, int N, int64_t ld, T* gbest,
                 sp_ctrl* ctrl, Best* scratch) {
  Best mine{1.0 / 0.0, 0x7fffffffffffffffLL};
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P; r += (int64_t)gridDim.x * blockDim.x) {
    double v = (double)xfun[r];
    if (better(v, r, mine.f, mine.row)) mine = Best{v, r};
  }
  Best top;
  if (grid_best(mine, scratch, ctrl, &top)) finalize_generation<T>(top, x, ld, N, gbest, ctrl, -1, 0, 0.0, 0.0);
}

// Sharded swarm: every rank holds the records [fit, x_0 .. x_{N-1}] of all ranks' local
// bests (all-gathered); pick the first minimum (ranks own ascending row ranges, so this
// is np.argmin's tie rule), then distance to the previous gbest and the status ladder.
template <typename T>
__global__ void __launch_bounds__(kThreads)
gbest_reduce_kernel(const T* __restrict__ recs, int world, int N, int64_t rec_ld, T* gbest, sp_ctrl* ctrl, int it,
                    int maxiter, double xtol, double ftol) {
  if (!running(ctrl)) return;
  int best = 0;
  for (int r = 1; r < world; ++r)
    if (recs[r * rec_ld] < recs[best * rec_ld]) best = r;
  // finalize_generation reads the winning row from `xrows + row * ld`
  Best b{(double)recs[best * rec_ld], (long long)best};
  finalize_generation<T>(b, recs + 1, rec_ld, N, gbest, ctrl, it, maxiter, xtol, ftol);
}

template <typename T>
static int select_launch(int it, int maxiter, double xtol, double ftol, const void* cand, const void* candfun, void* x,
                         void* xfun, int64_t P, int N, int64_t ld, int copy_when, void* gbest, sp_ctrl* ctrl,
                         void* scratch, cudaStream_t s) {
  Shape sh;
  if (!pick_shape(N, Num<T>::VEC, &sh)) {
    set_error("sp_select_sync: ndim %d exceeds the compiled row shapes", N);
    return SP_ERR_SHAPE;
  }
  const int grid = grid_for_rows(P, sh.lpr, 8);
#define SP_CALL(TT, C, L)                                                                                          \
  select_kernel<TT, C, L><<<grid, kThreads, 0, s>>>(it, maxiter, xtol, ftol, (const TT*)cand, (const TT*)candfun,  \
                                                    (TT*)x, (TT*)xfun, P, N, ld, copy_when, (TT*)gbest, ctrl,      \
                                                    (Best*)scratch)
  SP_DISPATCH_SHAPE(T, sh, SP_CALL);
#undef SP_CALL
  SP_CHECK_LAUNCH();
  return SP_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace sp

using namespace sp;

extern "C" {

int sp_abi_version(void) { return SP_ABI_VERSION; }
const char* sp_last_error(void) { return g_err; }
int64_t sp_launch_count(void) { return (int64_t)g_launches.load(); }
int64_t sp_scratch_bytes(void) { return kScratchBytes; }

int sp_device_info(int device, int* sms, int64_t* l2, int64_t* hbm, int* cc) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) {
    set_error("sp_device_info: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  if (sms) *sms = p.multiProcessorCount;
  if (l2) *l2 = p.l2CacheSize;
  if (hbm) *hbm = (int64_t)p.totalGlobalMem;
  if (cc) *cc = p.major * 10 + p.minor;
  return SP_OK;
}

int sp_eval(int objective, int dtype, const void* X, int64_t P, int N, int64_t ld, const void* scale,
            const void* shift, void* f, void* stream) {
  SP_CHECK_ARG(X && f && P >= 1 && N >= 1, "null pointer or empty population");
  SP_CHECK_ARG(objective >= SP_OBJ_ACKLEY && objective <= SP_OBJ_STYBLINSKI_TANG, "unknown objective");
  SP_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift must come together");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  const int vec = dtype == SP_F32 ? 4 : 2;
  SP_CHECK_ARG(ld >= N && ld % vec == 0 && aligned16(X), "rows must be 16-byte aligned (ld multiple of 16/sizeof(T))");
  cudaStream_t s = (cudaStream_t)stream;
  return dtype == SP_F32 ? eval_launch<float>(objective, X, P, N, ld, scale, shift, f, 0, s)
                         : eval_launch<double>(objective, X, P, N, ld, scale, shift, f, 0, s);
}

int sp_lhs_init(int dtype, void* X, int64_t P, int N, int64_t ld, const void* lower, const void* upper,
                uint64_t seed, const void* jitter, const int64_t* perm, void* stream) {
  SP_CHECK_ARG(X && lower && upper && P >= 1 && N >= 1 && ld >= N, "null pointer or bad shape");
  SP_CHECK_ARG((jitter == nullptr) == (perm == nullptr), "jitter and perm must come together");
  SP_CHECK_ARG(P < (1LL << 31), "popsize");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t total = P * (int64_t)N;
  int grid = (int)((total + 255) / 256 < (int64_t)sm_count() * 8 ? (total + 255) / 256 : (int64_t)sm_count() * 8);
  if (dtype == SP_F32)
    lhs_kernel<float><<<grid, 256, 0, s>>>((float*)X, P, N, ld, (const float*)lower, (const float*)upper, seed,
                                           (const float*)jitter, perm, P, 0);
  else
    lhs_kernel<double><<<grid, 256, 0, s>>>((double*)X, P, N, ld, (const double*)lower, (const double*)upper, seed,
                                            (const double*)jitter, perm, P, 0);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_lhs_init_shard(int dtype, void* X, int64_t P_local, int N, int64_t ld, const void* lower, const void* upper,
                      uint64_t seed, int64_t P_total, int64_t row0, void* stream) {
  SP_CHECK_ARG(X && lower && upper && P_local >= 1 && N >= 1 && ld >= N, "null pointer or bad shape");
  SP_CHECK_ARG(row0 >= 0 && row0 + P_local <= P_total && P_total < (1LL << 31), "shard outside the population");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t total = P_local * (int64_t)N;
  int grid = (int)((total + 255) / 256 < (int64_t)sm_count() * 8 ? (total + 255) / 256 : (int64_t)sm_count() * 8);
  if (dtype == SP_F32)
    lhs_kernel<float><<<grid, 256, 0, s>>>((float*)X, P_local, N, ld, (const float*)lower, (const float*)upper, seed,
                                           nullptr, nullptr, P_total, row0);
  else
    lhs_kernel<double><<<grid, 256, 0, s>>>((double*)X, P_local, N, ld, (const double*)lower, (const double*)upper,
                                            seed, nullptr, nullptr, P_total, row0);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_gbest_reduce(int dtype, const void* recs, int world, int N, int64_t rec_ld, void* gbest, sp_ctrl* ctrl, int it,
                    int maxiter, double xtol, double ftol, void* stream) {
  SP_CHECK_ARG(recs && gbest && ctrl && world >= 1 && N >= 1 && rec_ld >= N + 1, "null pointer or bad shape");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SP_F32)
    gbest_reduce_kernel<float><<<1, kThreads, 0, s>>>((const float*)recs, world, N, rec_ld, (float*)gbest, ctrl, it,
                                                      maxiter, xtol, ftol);
  else
    gbest_reduce_kernel<double><<<1, kThreads, 0, s>>>((const double*)recs, world, N, rec_ld, (double*)gbest, ctrl, it,
                                                       maxiter, xtol, ftol);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_select_sync(int dtype, int it, int maxiter, double xtol, double ftol, const void* cand, const void* candfun,
                   void* x, void* xfun, int64_t P, int N, int64_t ld, int copy_when, void* gbest, sp_ctrl* ctrl,
                   void* scratch, void* stream) {
  SP_CHECK_ARG(cand && candfun && x && xfun && gbest && ctrl && scratch, "null pointer");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  const int vec = dtype == SP_F32 ? 4 : 2;
  SP_CHECK_ARG(P >= 1 && N >= 1 && ld >= N && ld % vec == 0 && aligned16(cand) && aligned16(x), "shape/alignment");
  cudaStream_t s = (cudaStream_t)stream;
  return dtype == SP_F32
             ? select_launch<float>(it, maxiter, xtol, ftol, cand, candfun, x, xfun, P, N, ld, copy_when, gbest, ctrl,
                                    scratch, s)
             : select_launch<double>(it, maxiter, xtol, ftol, cand, candfun, x, xfun, P, N, ld, copy_when, gbest, ctrl,
                                     scratch, s);
}

int sp_best_init(int dtype, const void* x, const void* xfun, int64_t P, int N, int64_t ld, void* gbest, sp_ctrl* ctrl,
                 void* scratch, void* stream) {
  SP_CHECK_ARG(x && xfun && gbest && ctrl && scratch && P >= 1 && N >= 1, "null pointer or bad shape");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t need = (P + kThreads - 1) / kThreads;
  int grid = (int)(need < (int64_t)sm_count() * 4 ? need : (int64_t)sm_count() * 4);
  if (dtype == SP_F32)
    best_init_kernel<float><<<grid, kThreads, 0, s>>>((const float*)x, (const float*)xfun, P, N, ld, (float*)gbest,
                                                      ctrl, (Best*)scratch);
  else
    best_init_kernel<double><<<grid, kThreads, 0, s>>>((const double*)x, (const double*)xfun, P, N, ld, (double*)gbest,
                                                       ctrl, (Best*)scratch);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_random_fill(int dtype, void* out, int64_t P, int N, int64_t ld, int it, int purpose, uint64_t seed, int normal,
                   void* stream) {
  SP_CHECK_ARG(out && P >= 1 && N >= 1 && ld >= N, "null pointer or bad shape");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  SP_CHECK_ARG(purpose != (int)kDeCross && purpose != (int)kPsoR1 && purpose != (int)kPsoR2,
               "the DE crossover / PSO coefficient streams are 16-bit pieces drawn inside the kernels (philox.cuh)");
  cudaStream_t s = (cudaStream_t)stream;
  const int vec = dtype == SP_F32 ? 4 : 2;
  int64_t need = (P * (int64_t)((N + vec - 1) / vec) + 255) / 256, cap = (int64_t)sm_count() * 8;
  const int grid = (int)(need < cap ? need : cap);
  if (dtype == SP_F32)
    random_fill_kernel<float><<<grid, 256, 0, s>>>((float*)out, P, N, ld, it, seed, (uint32_t)purpose, normal);
  else
    random_fill_kernel<double><<<grid, 256, 0, s>>>((double*)out, P, N, ld, it, seed, (uint32_t)purpose, normal);
  SP_CHECK_LAUNCH();
  return SP_OK;
}

int sp_fitness_rank(int dtype, const void* fit, int64_t P, int32_t* rank, void* stream) {
  SP_CHECK_ARG(fit && rank && P >= 1 && P < (1LL << 31), "null pointer or bad popsize");
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = dtype == SP_F32 ? rank_launch<float>((const float*)fit, P, rank, nullptr, s)
                                  : rank_launch<double>((const double*)fit, P, rank, nullptr, s);
  if (e != cudaSuccess) {
    set_error("sp_fitness_rank: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  return SP_OK;
}

int64_t sp_sym_eigh_work_scalars(int N) { return (int64_t)jacobi_work_scalars(N); }

int sp_sym_eigh(int dtype, void* C, int N, void* w, void* B, void* work, int warm, int32_t* sweeps, void* stream) {
  SP_CHECK_ARG(C && w && B && work && N >= 1 && N <= 1024, "null pointer or N outside [1, 1024]");  // work: sp_sym_eigh_work_scalars(N)
  SP_CHECK_ARG(dtype == SP_F32 || dtype == SP_F64, "dtype");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = dtype == SP_F32
                      ? jacobi_launch<float>((float*)C, N, (float*)w, (float*)B, (float*)work, warm, nullptr, nullptr, sweeps, s)
                      : jacobi_launch<double>((double*)C, N, (double*)w, (double*)B, (double*)work, warm, nullptr, nullptr, sweeps, s);
  if (e != cudaSuccess) {
    set_error("sp_sym_eigh: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  SP_CHECK_LAUNCH();
  return SP_OK;
}

}  // extern "C"
