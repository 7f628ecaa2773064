// what does the producer thread pay per bulk copy?  (profiles/r02_micro.md, M5)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, int bytes, int mode, long long* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[16];
  __shared__ long long tl[64];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int q = 0;
    if (mode == 0) {            // 8 x (arrive_expect_tx + copy) on 8 barriers, then 8 waits
#pragma unroll
      for (int i = 0; i < 8; ++i) { mbar_arrive_expect_tx(&full[i], bytes); bulk_g2s(smem + i * bytes, src + i * bytes, bytes, &full[i]); tl[q++] = clock64() - t0; }
#pragma unroll
      for (int i = 0; i < 8; ++i) { mbar_wait(&full[i], 0); tl[q++] = clock64() - t0; }
    } else if (mode == 1) {     // one arrive_expect_tx, 8 copies on one barrier, one wait
      mbar_arrive_expect_tx(&full[0], 8 * bytes); tl[q++] = clock64() - t0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { bulk_g2s(smem + i * bytes, src + i * bytes, bytes, &full[0]); tl[q++] = clock64() - t0; }
      mbar_wait(&full[0], 0); tl[q++] = clock64() - t0;
    } else if (mode == 2) {     // plain barrier: arrive + wait, 8 times
#pragma unroll
      for (int i = 0; i < 8; ++i) { mbar_arrive(&full[1]); tl[q++] = clock64() - t0; mbar_wait(&full[1], i & 1); tl[q++] = clock64() - t0; }
    } else if (mode == 3) {     // copies first, arrive_expect_tx afterwards (tx-count may go negative transiently: allowed)
#pragma unroll
      for (int i = 0; i < 8; ++i) { bulk_g2s(smem + i * bytes, src + i * bytes, bytes, &full[i]); tl[q++] = clock64() - t0; }
#pragma unroll
      for (int i = 0; i < 8; ++i) { mbar_arrive_expect_tx(&full[i], bytes); tl[q++] = clock64() - t0; }
#pragma unroll
      for (int i = 0; i < 8; ++i) { mbar_wait(&full[i], 0); tl[q++] = clock64() - t0; }
    } else if (mode == 4) {     // try_wait on an already completed phase, 8 times
      mbar_arrive(&full[2]); 
#pragma unroll
      for (int i = 0; i < 8; ++i) { mbar_wait(&full[2], 0); tl[q++] = clock64() - t0; }
    } else if (mode == 5) {     // clock64 + shared store only
#pragma unroll
      for (int i = 0; i < 8; ++i) { tl[q++] = clock64() - t0; }
    }
    for (int i = 0; i < 64; ++i) res[i] = i < q ? tl[i] : -1;
  }
}
int main() {
  uint8_t* src; long long* res; long long h[64];
  cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  cudaMalloc(&res, 64 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 1024);
  const char* names[] = {"8 x (expect_tx + copy), 8 waits", "1 expect_tx, 8 copies, 1 wait", "8 x (arrive, wait) plain", "8 copies, 8 expect_tx, 8 waits", "8 waits on a completed phase", "clock + store"};
  for (int rep = 0; rep < 2; ++rep)
    for (int bytes : {4096, 16384})
      for (int mode = 0; mode < 6; ++mode) {
        k<<<1, 128, 196608 + 1024>>>(src, bytes, mode, res);
        cudaDeviceSynchronize(); cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
        printf("bytes %5d %-34s:", bytes, names[mode]);
        for (int i = 0; i < 64 && h[i] >= 0; ++i) printf(" %lld", h[i]);
        printf("\n");
      }
  return 0;
}
