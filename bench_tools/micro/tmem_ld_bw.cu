// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM for the epilogue's access pattern.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

// mode 0: ld x16 + wait each; mode 1: two ld x16 in flight (double buffer); mode 2: ld x32 + wait; mode 3: 4 x (ld x16) then one wait
__global__ void __launch_bounds__(256, 1) k(int nwarps, int mode, int iters, float* out, long long* cyc) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        for (int c = 0; c < 256; c += 16) { uint32_t v[16]; ld16(tb + c, v); tmem_ld_wait(); acc += __uint_as_float(v[0]) + __uint_as_float(v[15]); }
      } else if (mode == 1) {
        uint32_t v0[16], v1[16];
        ld16(tb, v0);
        for (int c = 0; c < 256; c += 32) {
          tmem_ld_wait(); ld16(tb + c + 16, v1); acc += __uint_as_float(v0[0]) + __uint_as_float(v0[15]);
          tmem_ld_wait(); if (c + 32 < 256) ld16(tb + c + 32, v0); acc += __uint_as_float(v1[0]) + __uint_as_float(v1[15]);
        }
      } else if (mode == 2) {
        for (int c = 0; c < 256; c += 32) { uint32_t v[32]; tmem_ld_32x32(tb + c, v); tmem_ld_wait(); acc += __uint_as_float(v[0]) + __uint_as_float(v[31]); }
      } else {
        for (int c = 0; c < 256; c += 64) {
          uint32_t a[16], b[16], d[16], e[16];
          ld16(tb + c, a); ld16(tb + c + 16, b); ld16(tb + c + 32, d); ld16(tb + c + 48, e);
          tmem_ld_wait();
          acc += __uint_as_float(a[0]) + __uint_as_float(b[0]) + __uint_as_float(d[0]) + __uint_as_float(e[15]);
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1024); cudaMallocManaged(&cyc, 8);
  const int iters = 200;
  for (int mode = 0; mode < 4; ++mode)
    for (int nw : {1, 4, 8}) {
      k<<<1, 256>>>(nw, mode, iters, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      const double bytes = (double)nw * iters * 256 * 32 * 4;
      printf("mode %d warps %d: %lld cycles, %.1f B/clk per SM, %.1f cycles per 128x256 fp32 tile-half-slot (128 KB)\n", mode, nw, *cyc,
             bytes / *cyc, 131072.0 / (bytes / *cyc));
    }
  return 0;
}
