// L2 -> shared memory bulk-copy service model per SM (profiles/r02_micro.md, M4): copies in flight x copy size x issuing
// threads, every SM streaming the same 1 MB region.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu && ./tma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;

// `nthr` issuing threads (one per warp), each with its own ring of `depth` buffers of `bytes`; each thread keeps `depth`
// copies in flight; `split` > 1 issues every buffer as `split` separate bulk copies on the same barrier.
__global__ void __launch_bounds__(128, 1) bulk_probe(const uint8_t* src, int bytes, int depth, int nthr, int split, int n_copies, long long* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[4][16];
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(&full[0][0] + i, 1); fence_mbar_init(); }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nthr) {
    uint8_t* base = smem + (size_t)w * depth * bytes;
    const int piece = bytes / split;
    const long long t0 = clock64();
    for (int n = 0; n < n_copies + depth; ++n) {
      const int st = n % depth;
      if (n >= depth) mbar_wait(&full[w][st], ((n / depth) - 1) & 1u);
      if (n < n_copies) {
        mbar_arrive_expect_tx(&full[w][st], bytes);
        const uint8_t* s = src + ((size_t)(n * nthr + w) * bytes) % (1 << 20);
        for (int p = 0; p < split; ++p) bulk_g2s(base + st * bytes + p * piece, s + p * piece, piece, &full[w][st]);
      }
    }
    res[blockIdx.x * 4 + w] = clock64() - t0;
  }
}

int main() {
  uint8_t* src; long long* res;
  cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  cudaMallocManaged(&res, 148 * 4 * 8);
  cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 1024);
  const int n = 1000;
  for (int grid : {1, 148})
    for (int nthr : {1, 2, 4})
      for (int bytes : {4096, 8192, 16384, 32768, 65536})
        for (int depth : {1, 2, 4, 8})
          for (int split : {1, 4}) {
            if ((size_t)bytes * depth * nthr > 196608) continue;
            if (grid == 148 && (depth == 1 || depth == 8)) continue;
            if (split == 4 && (nthr != 1 || depth != 2)) continue;
            bulk_probe<<<grid, 128, 196608 + 1024>>>(src, bytes, depth, nthr, split, n, res);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            double cyc = 0;
            for (int b = 0; b < grid; ++b) for (int w = 0; w < nthr; ++w) cyc = cyc > (double)res[b * 4 + w] ? cyc : (double)res[b * 4 + w];
            printf("M4 grid %3d threads %d copy %5d B x%d in flight %d/thread: %.0f cycles per copy-slot, %.1f B/clk/SM (%.0f B/clk chip)\n", grid, nthr,
                   bytes, split, depth, cyc / n, (double)n * bytes * nthr / cyc, (double)n * bytes * nthr / cyc * grid);
          }
  return 0;
}
