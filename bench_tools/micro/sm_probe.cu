// Per-SM micro-benchmarks behind the round-2 forward-kernel design (profiles/r02_micro.md):
//   M1  MUFU.SIN epilogue arithmetic: cycles per warp-element as a function of warps per sub-partition and of the fraction
//       of sines moved to the FMA pipe
//   M2  tcgen05.ld throughput while tcgen05.mma runs (accumulator read-out of one slot under the other slot's MMAs), as a
//       function of reader warps, load shape and loads in flight; with and without the sine work / shared-memory stores
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sm_probe sm_probe.cu && ./sm_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;

// ---------------------------------------------------------------------------------------------- M1
__device__ __forceinline__ float sin_poly(float u) {
  // u in half-revolutions (angle = pi * u): k = rint(u), f = u - k in [-0.5, 0.5], sin(pi u) = (-1)^k sin(pi f)
  const float v = u + 12582912.0f;
  const float kf = v - 12582912.0f;
  const float f = u - kf;
  const uint32_t sgn = __float_as_uint(v) << 31;
  const float f2 = f * f;
  float p = fmaf(-0.5958483f, f2, 2.5500992f);     // odd degree-7 fit of sin(pi f) on [-0.5, 0.5] (coefficients illustrative)
  p = fmaf(p, f2, -5.1677127f);
  p = fmaf(p, f2, 3.1415927f);
  return __uint_as_float(__float_as_uint(p * f) ^ sgn);
}
template <int kPoly /* elements of every 8 on the FMA pipe */, int kMode>
__global__ void __launch_bounds__(1024, 1) mufu_probe(int iters, float s, float t, float* out, long long* cyc) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (float)(threadIdx.x + i);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (kMode == 0) x[i] = __sinf(x[i]);                              // FMUL + MUFU
      else if ((i & 7) < kPoly) x[i] = sin_poly(fmaf(x[i], s, t));      // FMA pipe
      else x[i] = __sinf(fmaf(x[i], s, t));                             // FFMA + FMUL + MUFU
    }
    if (kMode == 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack_bf16x2(x[i], x[i + 1]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  float sum = __uint_as_float(acc & 0x3f800000u);
#pragma unroll
  for (int i = 0; i < 16; ++i) sum += x[i];
  out[threadIdx.x] = sum;
}

// ---------------------------------------------------------------------------------------------- M2
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st_v4(uint32_t smem_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct M2Out { long long mma_cycles; long long n_mma; long long rd_cycles[32]; long long rd_bytes[32]; };

// work: 0 = touch only, 1 = FMUL + MUFU.SIN + pack per element, 2 = the same + 16-byte stores into a swizzled row
// depth: x16 loads in flight per warp before the wait (1, 2, 4); shape32: use x32 loads (depth counts x32 loads)
template <int kDepth, bool kX32, int kWork, bool kPipe>
__global__ void __launch_bounds__(640, 1) ldtm_probe(int n_readers, int mma_on, int mma_n /*128 or 256*/, int n_batches, M2Out* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // operands: A = [128][64] sw128 chunk (16 KB), B = [256][64] sw128 chunk (32 KB); epilogue rows at 64 KB.. (64 KB)
  for (int i = threadIdx.x; i < 131072 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); stop = 0; fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (warp == 0) {
    if (elect_one()) {
      const long long t0 = clock64();
      if (mma_on) {
        const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)mma_n);
        const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(smem)), bd = umma_desc_kmajor_sw128(smem_u32(smem + 16384));
        for (int b = 0; b < n_batches; ++b) {
          if (b >= 2) mbar_wait(&bars[b & 1], ((b - 2) >> 1) & 1u);      // at most two batches in flight
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 16; ++k) umma_bf16_ss(tb + (uint32_t)(b & 1) * 0u, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, k != 0);
          umma_commit(&bars[b & 1]);
        }
        for (int b = max(n_batches - 2, 0); b < n_batches; ++b) mbar_wait(&bars[b & 1], (b >> 1) & 1u);
        out->n_mma = (long long)n_batches * 16;
      } else {
        while (clock64() - t0 < (long long)n_batches * 2048) { }
        out->n_mma = 0;
      }
      out->mma_cycles = clock64() - t0;
      stop = 1;
    }
  } else if (warp >= 4 && warp < 4 + n_readers) {
    const int r = warp - 4;
    const int quad = warp & 3;
    const int grp = r >> 2, ngrp = (n_readers + 3) >> 2;               // readers of one lane quarter split the 256 columns
    const int cols = 256 / ngrp;
    const uint32_t taddr = tb + 256u + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * cols);
    const uint32_t row = smem_u32(smem + 65536) + (uint32_t)((quad * 32 + lane) * 128);
    const int r7 = lane & 7;
    long long bytes = 0;
    float sink = 0.f;
    const long long t0 = clock64();
    constexpr int NE = kX32 ? 32 : 16;
    constexpr int STEP = NE * kDepth;
    auto issue = [&](uint32_t (&v)[kDepth][NE], int c) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        if constexpr (kX32) tmem_ld_32x32(taddr + c + d * 32, v[d]); else ld16(taddr + c + d * 16, v[d]);
      }
    };
    auto work = [&](uint32_t (&v)[kDepth][NE], int c) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        if (kWork == 0) {
          sink += __uint_as_float(v[d][0]) + __uint_as_float(v[d][NE - 1]);
        } else {
#pragma unroll
          for (int g = 0; g < NE / 8; ++g) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __sinf(__uint_as_float(v[d][g * 8 + i]));
            const uint32_t p0 = pack_bf16x2(o[0], o[1]), p1 = pack_bf16x2(o[2], o[3]), p2 = pack_bf16x2(o[4], o[5]), p3 = pack_bf16x2(o[6], o[7]);
            if (kWork == 2) st_v4(row + (uint32_t)((((c >> 3) + d * (NE / 8) + g) & 7) ^ r7) * 16u + (uint32_t)(((c >> 6) & 3) * 16384), p0, p1, p2, p3);
            else sink += __uint_as_float((p0 ^ p1 ^ p2 ^ p3) & 0x3f800000u);
          }
        }
      }
    };
    while (!stop) {
      if constexpr (kPipe) {
        uint32_t va[kDepth][NE], vb[kDepth][NE];
        issue(va, 0);
        for (int c = 0; c < cols; c += 2 * STEP) {
          tmem_ld_wait();
          issue(vb, (c + STEP) % cols);
          work(va, c);
          tmem_ld_wait();
          issue(va, (c + 2 * STEP) % cols);
          work(vb, c + STEP);
        }
        tmem_ld_wait();
      } else {
        for (int c = 0; c < cols; c += STEP) {
          uint32_t v[kDepth][NE];
          issue(v, c);
          tmem_ld_wait();
          work(v, c);
        }
      }
      bytes += (long long)cols * 32 * 4;
    }
    const long long t1 = clock64();
    if (lane == 0) { out->rd_cycles[r] = t1 - t0; out->rd_bytes[r] = bytes; }
    if (sink == 123.456f) out->n_mma = -1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int kDepth, bool kX32, int kWork, bool kPipe>
static void run_m2(const char* name, M2Out* o) {
  auto kern = ldtm_probe<kDepth, kX32, kWork, kPipe>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  for (int mma_on = 0; mma_on < 2; ++mma_on)
    for (int nr : {4, 8, 16}) {
      for (int mma_n : {128, 256}) {
        if (mma_n == 128 && (kWork != 0 || !mma_on)) continue;
        memset(o, 0, sizeof(M2Out));
        kern<<<1, 128 + 32 * nr, 131072 + 1024>>>(nr, mma_on, mma_n, 200, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        double b = 0, c = 0;
        for (int r = 0; r < nr; ++r) { b += (double)o->rd_bytes[r]; c = c > (double)o->rd_cycles[r] ? c : (double)o->rd_cycles[r]; }
        printf("M2 %-22s readers %2d mma %d (N=%d): ldtm %.1f B/clk/SM -> %.0f cycles per 128 KB accumulator;  mma %.1f cycles each\n", name, nr,
               mma_on, mma_n, b / c, 131072.0 / (b / c), o->n_mma > 0 ? (double)o->mma_cycles / (double)o->n_mma : 0.0);
      }
    }
}


// ---------------------------------------------------------------------------------------------- M4
// L2 -> shared memory bulk-copy latency and sustained rate per SM, every SM streaming the same 1 MB region (the weights).
__global__ void __launch_bounds__(128, 1) bulk_probe(const uint8_t* src, int stage_bytes, int nstage, int n_copies, long long* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    // latency of one isolated copy (after a warm-up copy)
    long long lat = 0;
    for (int i = 0; i < 5; ++i) {
      const long long t0 = clock64();
      mbar_arrive_expect_tx(&full[0], stage_bytes);
      bulk_g2s(smem, src + ((size_t)(i * 37 + blockIdx.x) * stage_bytes) % (1 << 20), stage_bytes, &full[0]);
      mbar_wait(&full[0], i & 1u);
      if (i) lat += clock64() - t0;
    }
    uint32_t ph[16];
    for (int i = 0; i < 16; ++i) ph[i] = (i == 0) ? 1u : 0u;
    const long long t0 = clock64();
    for (int n = 0; n < n_copies + nstage; ++n) {
      const int st = n % nstage;
      if (n >= nstage) { mbar_wait(&full[st], ph[st]); ph[st] ^= 1u; }
      if (n < n_copies) {
        mbar_arrive_expect_tx(&full[st], stage_bytes);
        bulk_g2s(smem + st * stage_bytes, src + ((size_t)n * stage_bytes) % (1 << 20), stage_bytes, &full[st]);
      }
    }
    const long long t1 = clock64();
    res[2 * blockIdx.x] = lat / 4;
    res[2 * blockIdx.x + 1] = t1 - t0;
  }
}
static void run_m4() {
  uint8_t* src; long long* res;
  cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  cudaMallocManaged(&res, 148 * 2 * 8);
  cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  for (int grid : {1, 148})
    for (int sb : {8192, 16384, 32768})
      for (int ring : {65536, 131072}) {
        const int nstage = ring / sb;
        if (nstage > 16) continue;
        const int n = 2000;
        bulk_probe<<<grid, 128, 131072 + 1024>>>(src, sb, nstage, n, res);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        double lat = 0, cyc = 0;
        for (int b = 0; b < grid; ++b) { lat += (double)res[2 * b]; cyc = cyc > (double)res[2 * b + 1] ? cyc : (double)res[2 * b + 1]; }
        printf("M4 grid %3d stage %5d B ring %6d B: isolated copy latency %.0f cycles; sustained %.1f B/clk/SM (%.1f B/clk chip)\n", grid, sb, ring,
               lat / grid, (double)n * sb / cyc, (double)n * sb / cyc * grid);
      }
}

template <int kPoly, int kMode>
static void run_m1(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  for (int nw : {4, 8, 16, 32}) {
    mufu_probe<kPoly, kMode><<<1, nw * 32>>>(iters, 1.0001f, 0.1f, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    const double wel = (double)iters * 16 * (nw / 4.0);      // warp-elements per sub-partition
    printf("M1 %-28s warps/SMSP %d: %.2f cycles per warp-element per SMSP  (tile-layer of 128x256: %.0f cycles)\n", name, nw / 4,
           (double)*cyc / wel, (double)*cyc / wel * 256.0);
  }
}

int main() {
  float* out; long long* cyc; M2Out* o;
  cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8); cudaMallocManaged(&o, sizeof(M2Out));
  run_m1<0, 0>("FMUL+MUFU", out, cyc);
  run_m1<0, 1>("FFMA+FMUL+MUFU", out, cyc);
  run_m1<0, 2>("FFMA+FMUL+MUFU+pack", out, cyc);
  run_m1<1, 2>("same, 1/8 poly", out, cyc);
  run_m1<2, 2>("same, 2/8 poly", out, cyc);
  run_m1<3, 2>("same, 3/8 poly", out, cyc);
  run_m1<8, 2>("all poly", out, cyc);
  run_m2<1, false, 0, false>("x16 d1 touch", o);
  run_m2<1, false, 0, true>("x16 d1 pipe touch", o);
  run_m2<2, false, 0, true>("x16 d2 pipe touch", o);
  run_m2<4, false, 0, false>("x16 d4 touch", o);
  run_m2<1, true, 0, true>("x32 d1 pipe touch", o);
  run_m2<1, false, 1, true>("x16 d1 pipe sin", o);
  run_m2<2, false, 1, true>("x16 d2 pipe sin", o);
  run_m2<1, true, 1, true>("x32 d1 pipe sin", o);
  run_m2<1, false, 2, true>("x16 d1 pipe sin+st", o);
  run_m2<2, false, 2, true>("x16 d2 pipe sin+st", o);
  run_m2<1, true, 2, true>("x32 d1 pipe sin+st", o);
  run_m4();
  return 0;
}
