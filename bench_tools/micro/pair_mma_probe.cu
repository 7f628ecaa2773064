// cta_group::2 MMA issue/throughput model (profiles/r02_micro.md, M6): cycles per tcgen05.mma as a function of N, of the
// accumulator pattern and of the commit frequency.  One cluster of two CTAs; operands are zeros in shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;

// n_acc: MMAs rotate over this many accumulators (columns acc * N); commit_every: one multicast commit per this many MMAs
// (to a scratch barrier nobody waits on), plus the final one the leader waits on.
template <int N, int n_acc, int commit_every, int pairmode>
__global__ void __launch_bounds__(128, 1) k(int n_mma, long long* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, scratch[8];
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 131072 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&scratch[i], 1); fence_mbar_init(); }
  if (warp == 1) {
    if (pairmode) { tmem_alloc_pair(&tbase, 512); tmem_relinquish_pair(); } else { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tb = tbase;
  if ((rank == 0 || !pairmode) && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(pairmode ? 256 : 128, (uint32_t)N);
    const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(smem)), bd = umma_desc_kmajor_sw128(smem_u32(smem + 65536));
    const long long t0 = clock64();
    for (int i0 = 0; i0 < n_mma; i0 += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t d = tb + (uint32_t)((j % n_acc) * N);
        // K-chunk j >> 2 of the activation tile (16 KB apart), stage j >> 2 of the weight ring (brows * 128 B apart)
        constexpr uint32_t brows = pairmode ? N / 2 : N;
        const uint64_t aj = ad + (uint64_t)(((j >> 2) * 16384) >> 4) + 2 * (j & 3);
        const uint64_t bj = bd + (uint64_t)((((j >> 2) + 4 * ((i0 >> 4) & 1)) * brows * 128 % 65536) >> 4) + 2 * (j & 3);
        if (pairmode) umma_bf16_ss_pair(d, aj, bj, idesc, 1u);
        else umma_bf16_ss(d, aj, bj, idesc, 1u);
        if (commit_every > 0 && (j % commit_every) == commit_every - 1) {
          if (pairmode) umma_commit_pair(&scratch[(j / commit_every) & 7], (uint16_t)0x3); else umma_commit(&scratch[(j / commit_every) & 7]);
        }
      }
    }
    const long long t1 = clock64();
    if (pairmode) umma_commit_pair(&bar, (uint16_t)0x3); else umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    res[0] = t1 - t0; res[1] = t2 - t0;
  } else if (pairmode && rank == 1 && threadIdx.x == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 1) { tc_fence_after(); if (pairmode) tmem_dealloc_pair(tb, 512); else tmem_dealloc(tb, 512); }
}


template <int N, int n_acc, int ce, int pairmode>
static void run(long long* res) {
  long long h[2];
  const int n_mma = 2048;
  auto kern = k<N, n_acc, ce, pairmode>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 131072 + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, n_mma, res);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost);
  printf("M6 %s M=%d N=%3d accumulators %d commit every %2d: %.1f cycles per MMA (issue loop %.1f)\n", pairmode ? "cta_group::2" : "cta_group::1",
         pairmode ? 256 : 128, N, n_acc, ce, (double)h[1] / n_mma, (double)h[0] / n_mma);
}
template <int N, int pairmode>
static void run_n(long long* res) {
  run<N, 1, 0, pairmode>(res); run<N, 1, 16, pairmode>(res); run<N, 1, 4, pairmode>(res); run<N, 1, 2, pairmode>(res); run<N, 1, 1, pairmode>(res);
  run<N, 2, 0, pairmode>(res); run<N, 2, 4, pairmode>(res);
}
int main() {
  long long* res;
  cudaMalloc(&res, 16);
  run_n<256, 1>(res); run_n<128, 1>(res); run_n<64, 1>(res);
  run_n<256, 0>(res); run_n<128, 0>(res); run_n<64, 0>(res);
  return 0;
}
