// Probe for the DRAM *reads* ncu attributes to the VD-CMA sampling kernel (which algorithmically reads nothing):
// write a P x 1024 fp32 matrix (P = 16384: 64 MiB) with the access patterns the kernels use and let ncu count
// dram__bytes_read / dram__bytes_write / L2 sector traffic per variant.
//   0 plain 16-byte stores, a warp per row (8 x 512 B per row)      1 the same with st.global.cs (evict-first)
//   2 variant 0 + two scattered 4-byte stores per row (yvn, arfit)  3 row staged in shared memory, one 4 KiB
//   cp.async.bulk (TMA) store per row                                4 variant 0 with a 256 MiB matrix (> L2)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/write_probe profiles/write_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int N = 1024;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* __restrict__ y, float* __restrict__ s1, float* __restrict__ s2, int P) {
  __shared__ __align__(128) float stage[8][N];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warp = blockIdx.x * 8 + w, nwarps = gridDim.x * 8;
  for (int row = warp; row < P; row += nwarps) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j0 = (c * 32 + lane) * 4;
      float4 v = make_float4(row + j0, row - j0, (float)c, (float)lane);
      acc += v.x;
      float4* dst = reinterpret_cast<float4*>(y + (size_t)row * N + j0);
      if (MODE == 1) __stcs(dst, v);
      else if (MODE == 3) *reinterpret_cast<float4*>(&stage[w][j0]) = v;
      else *dst = v;
    }
    if (MODE == 3) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y + (size_t)row * N),
                     "r"((uint32_t)__cvta_generic_to_shared(&stage[w][0])), "r"(N * 4)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    }
    if (MODE == 2 && lane == 0) {
      s1[row] = acc;
      s2[row] = -acc;
    }
  }
}

int main() {
  const int P = 16384, Pbig = 65536;
  float *y, *s1, *s2;
  cudaMalloc(&y, (size_t)Pbig * N * 4);
  cudaMalloc(&s1, Pbig * 4);
  cudaMalloc(&s2, Pbig * 4);
  const int grid = 148 * 3;
  for (int rep = 0; rep < 2; ++rep) {
    probe<0><<<grid, 256>>>(y, s1, s2, P);
    probe<1><<<grid, 256>>>(y, s1, s2, P);
    probe<2><<<grid, 256>>>(y, s1, s2, P);
    probe<3><<<grid, 256>>>(y, s1, s2, P);
    probe<0><<<grid, 256>>>(y, s1, s2, Pbig);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("write_probe: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
