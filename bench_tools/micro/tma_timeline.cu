// timeline of bulk copies issued by one thread: clock at every issue and at every observed completion
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, int bytes, int depth, int n_copies, int poll, long long* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[16];
  __shared__ long long tl[128];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int n = 0; n < n_copies + depth; ++n) {
      const int st = n % depth;
      if (n >= depth) {
        const uint32_t par = ((n / depth) - 1) & 1u;
        if (poll) { while (!mbar_try_wait(&full[st], par)) { } } else mbar_wait(&full[st], par);
        tl[2 * (n - depth) + 1] = clock64() - t0;
      }
      if (n < n_copies) {
        mbar_arrive_expect_tx(&full[st], bytes);
        bulk_g2s(smem + st * bytes, src + ((size_t)n * bytes) % (1 << 20), bytes, &full[st]);
        tl[2 * n] = clock64() - t0;
      }
    }
    for (int i = 0; i < 2 * n_copies; ++i) res[i] = tl[i];
  }
}
int main() {
  uint8_t* src; long long* res;
  cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  cudaMalloc(&res, 64 * 2 * 8); long long h[128];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 1024);
  for (int poll = 0; poll < 2; ++poll)
  for (int bytes : {4096, 16384})
    for (int depth : {1, 2, 4}) {
      k<<<1, 128, 196608 + 1024>>>(src, bytes, depth, 12, poll, res);
      cudaDeviceSynchronize(); cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
      printf("bytes %d depth %d poll %d\n  issue:", bytes, depth, poll);
      for (int i = 0; i < 12; ++i) printf(" %lld", h[2 * i]);
      printf("\n  done :");
      for (int i = 0; i < 12; ++i) printf(" %lld", h[2 * i + 1]);
      printf("\n");
    }
  return 0;
}
