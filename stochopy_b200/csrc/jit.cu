// User objectives compiled at run time: the reference's per-individual
// contract `fun(x, *args) -> float` (stochopy/optimize/_common.py:27-106) keeps arbitrary
// Python callables on the host; an objective given as CUDA C source
//     __device__ real objective(const real* x, int n)
// is compiled with NVRTC for sm_100a and evaluated on the device, so the population never
// leaves HBM: propose (sp_*_propose / sp_cma_sample / sp_vd_sample) -> sp_jit_eval ->
// select (sp_select_sync / sp_*_update), all enqueued on one stream.
//
// Evaluation kernel: the contract is one sequential call per individual, so a thread
// evaluates a row -- but rows are staged through shared memory first: the CTA loads
// `rows_per_cta` rows with coalesced reads into a tile with an odd row stride (no bank
// conflicts when 32 threads walk their rows in lock step), applies the optional
// un-standardisation x * scale + shift (_cmaes.py:168-173), and each thread then calls
// objective() on its row in shared memory.
//
// libnvrtc is opened with dlopen on first use, so the library loads without it.
#include <dlfcn.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace sp {

typedef struct _nvrtcProgram* nvrtcProgram_t;
struct Nvrtc {
  void* so = nullptr;
  int (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram_t, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram_t, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram_t, char*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram_t, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram_t, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram_t*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
      n.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.so) break;
    }
    if (!n.so) return;
#define SP_SYM(field, name) *reinterpret_cast<void**>(&n.field) = dlsym(n.so, name)
    SP_SYM(CreateProgram, "nvrtcCreateProgram");
    SP_SYM(CompileProgram, "nvrtcCompileProgram");
    SP_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    SP_SYM(GetProgramLog, "nvrtcGetProgramLog");
    SP_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    SP_SYM(GetCUBIN, "nvrtcGetCUBIN");
    SP_SYM(DestroyProgram, "nvrtcDestroyProgram");
    SP_SYM(GetErrorString, "nvrtcGetErrorString");
#undef SP_SYM
    n.ok = n.CreateProgram && n.CompileProgram && n.GetProgramLogSize && n.GetProgramLog && n.GetCUBINSize &&
           n.GetCUBIN && n.DestroyProgram;
  });
  return n;
}

static const char* kJitEpilogue = R"SRC(
extern "C" __global__ void __launch_bounds__(128)
sp_jit_eval_kernel(const real* __restrict__ X, long long P, int N, long long ld, const real* __restrict__ scale,
                   const real* __restrict__ shift, real* __restrict__ f, int rows_per_cta, int stride) {
  extern __shared__ __align__(16) unsigned char sp_jit_smem[];
  real* tile = reinterpret_cast<real*>(sp_jit_smem);
  for (long long base = (long long)blockIdx.x * rows_per_cta; base < P; base += (long long)gridDim.x * rows_per_cta) {
    const long long left = P - base;
    const int rows = left < rows_per_cta ? (int)left : rows_per_cta;
    for (int e = threadIdx.x; e < rows * N; e += blockDim.x) {
      const int r = e / N, j = e - r * N;
      real v = X[(base + r) * ld + j];
      if (scale != nullptr) v = SP_ADD_RN(SP_MUL_RN(v, scale[j]), shift[j]);
      tile[r * stride + j] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < rows) f[base + threadIdx.x] = objective(tile + threadIdx.x * stride, N);
    __syncthreads();
  }
}
)SRC";

struct JitObjective {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
  int dtype = SP_F64;
  std::vector<char> cubin;
};

// source -> sm_100a cubin (no GPU needed); error text (incl. the compiler log) in *err
static bool jit_build(const char* user_src, int dtype, std::vector<char>* cubin, std::string* err) {
  Nvrtc& n = nvrtc();
  if (!n.ok) {
    *err = "libnvrtc.so.12 could not be loaded (dlopen)";
    return false;
  }
  std::string src;
  if (dtype == SP_F32)
    src += "typedef float real;\n#define SP_ADD_RN(a, b) __fadd_rn(a, b)\n#define SP_MUL_RN(a, b) __fmul_rn(a, b)\n";
  else
    src += "typedef double real;\n#define SP_ADD_RN(a, b) __dadd_rn(a, b)\n#define SP_MUL_RN(a, b) __dmul_rn(a, b)\n";
  src += "#line 1 \"objective.cu\"\n";
  src += user_src;
  src += "\n";
  src += kJitEpilogue;
  nvrtcProgram_t prog = nullptr;
  int rc = n.CreateProgram(&prog, src.c_str(), "sp_jit_objective.cu", 0, nullptr, nullptr);
  if (rc != 0) {
    *err = std::string("nvrtcCreateProgram: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error");
    return false;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "--fmad=true"};
  rc = n.CompileProgram(prog, 4, opts);
  if (rc != 0) {
    size_t len = 0;
    n.GetProgramLogSize(prog, &len);
    std::string log(len, '\0');
    if (len > 0) n.GetProgramLog(prog, &log[0]);
    *err = std::string("NVRTC: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error") + "\n" + log;
    n.DestroyProgram(&prog);
    return false;
  }
  size_t sz = 0;
  rc = n.GetCUBINSize(prog, &sz);
  if (rc == 0 && sz > 0) {
    cubin->resize(sz);
    rc = n.GetCUBIN(prog, cubin->data());
  }
  n.DestroyProgram(&prog);
  if (rc != 0 || sz == 0) {
    *err = "NVRTC produced no cubin";
    return false;
  }
  return true;
}

}  // namespace sp

using namespace sp;

extern "C" {

int sp_jit_check(const char* source, int dtype, int64_t* cubin_bytes) {
  SP_CHECK_ARG(source != nullptr && (dtype == SP_F32 || dtype == SP_F64), "source / dtype");
  std::vector<char> cubin;
  std::string err;
  if (!jit_build(source, dtype, &cubin, &err)) {
    set_error("sp_jit_check: %s", err.c_str());
    return SP_ERR_ARG;
  }
  if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
  return SP_OK;
}

int sp_jit_compile(const char* source, int dtype, void** handle) {
  SP_CHECK_ARG(source != nullptr && handle != nullptr && (dtype == SP_F32 || dtype == SP_F64), "source / handle / dtype");
  JitObjective* j = new JitObjective();
  j->dtype = dtype;
  std::string err;
  if (!jit_build(source, dtype, &j->cubin, &err)) {
    set_error("sp_jit_compile: %s", err.c_str());
    delete j;
    return SP_ERR_ARG;
  }
  cudaError_t e = cudaLibraryLoadData(&j->lib, j->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&j->kern, j->lib, "sp_jit_eval_kernel");
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute((const void*)j->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    set_error("sp_jit_compile: loading the compiled objective failed: %s", cudaGetErrorString(e));
    if (j->lib) cudaLibraryUnload(j->lib);
    delete j;
    return SP_ERR_CUDA;
  }
  *handle = j;
  return SP_OK;
}

int sp_jit_eval(void* handle, int dtype, const void* X, int64_t P, int N, int64_t ld, const void* scale,
                const void* shift, void* f, void* stream) {
  SP_CHECK_ARG(handle != nullptr && X != nullptr && f != nullptr && P >= 1 && N >= 1 && ld >= N, "null pointer or bad shape");
  JitObjective* j = static_cast<JitObjective*>(handle);
  SP_CHECK_ARG(dtype == j->dtype, "dtype differs from the one the objective was compiled for");
  SP_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift come together");
  const size_t elem = dtype == SP_F32 ? 4 : 8;
  int stride = N | 1;  // odd row stride: 32 threads walking their rows hit 32 different banks
  const size_t budget = 200 * 1024;
  SP_CHECK_ARG((size_t)stride * elem <= budget, "ndim too large for the shared-memory row tile");
  int rows = (int)(budget / ((size_t)stride * elem));
  if (rows > 128) rows = 128;
  // several CTAs per SM when the tile is small: cap the tile so that >= 2 fit
  while (rows > 32 && (size_t)rows * stride * elem > 96 * 1024) rows >>= 1;
  const size_t smem = (size_t)rows * stride * elem;
  int64_t need = (P + rows - 1) / rows, cap = (int64_t)sm_count() * 4;
  const int grid = (int)(need < cap ? need : cap);
  long long P_ = P, ld_ = ld;
  void* args[] = {(void*)&X, &P_, &N, &ld_, (void*)&scale, (void*)&shift, &f, &rows, &stride};
  cudaError_t e = cudaLaunchKernel((const void*)j->kern, dim3(grid), dim3(128), args, smem, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("sp_jit_eval: %s", cudaGetErrorString(e));
    return SP_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return SP_OK;
}

int sp_jit_free(void* handle) {
  SP_CHECK_ARG(handle != nullptr, "null handle");
  JitObjective* j = static_cast<JitObjective*>(handle);
  if (j->lib) cudaLibraryUnload(j->lib);
  delete j;
  return SP_OK;
}

}  // extern "C"
