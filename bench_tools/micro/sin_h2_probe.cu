// M9: sines of a 16-element accumulator chunk with kPoly of its 8 pairs evaluated on the FMA pipe in HALF precision
// (range reduction in fp32: 3 operations per element; the polynomial and the result as f16x2: 5 HFMA2 / HMUL2 per PAIR; the sign
// of odd half-turns by three integer instructions per pair) and the rest by FMUL + MUFU.SIN; every pair leaves as f16x2.
// Cycles per warp-element per sub-partition at 1, 2, 4, 8 warps per sub-partition, and the accuracy of both paths after
// rounding to fp16 (the activation format this variant needs) against bf16-rounded MUFU sines (the shipped format).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sin_h2_probe sin_h2_probe.cu && ./sin_h2_probe
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;

__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) {
  uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
// sin of two arguments (radians) -> f16x2.  n = rint(x / pi), f = x / pi - n in [-0.5, 0.5], sin(x) = (-1)^n sin(pi f),
// sin(pi f) = f (c1 + s (c3 + s (c5 + s c7))), s = f^2, evaluated in f16x2.
template <bool kPacked32>
__device__ __forceinline__ uint32_t sin2_h2(float x0, float x1) {
  constexpr float INV_PI = 0.318309886f, MAGIC = 12582912.0f;
  float t0, t1, f0, f1;
  if (kPacked32) {
    const float2 t = fma2(make_float2(x0, x1), make_float2(INV_PI, INV_PI), make_float2(MAGIC, MAGIC));
    const float2 n = add2(t, make_float2(-MAGIC, -MAGIC));
    const float2 f = fma2(make_float2(x0, x1), make_float2(INV_PI, INV_PI), make_float2(-n.x, -n.y));
    t0 = t.x; t1 = t.y; f0 = f.x; f1 = f.y;
  } else {
    t0 = fmaf(x0, INV_PI, MAGIC); t1 = fmaf(x1, INV_PI, MAGIC);
    const float n0 = t0 - MAGIC, n1 = t1 - MAGIC;
    f0 = fmaf(x0, INV_PI, -n0); f1 = fmaf(x1, INV_PI, -n1);
  }
  const uint32_t h = pack_f16x2(f0, f1);
  const uint32_t s = h2_mul(h, h);
  // c7 = -0.5993 (0xB8CB), c5 = 2.5502 (0x411A), c3 = -5.1677 (0xC52B), c1 = 3.1416 (0x4248)
  uint32_t q = h2_fma(s, 0xB8CBB8CBu, 0x411A411Au);
  q = h2_fma(q, s, 0xC52BC52Bu);
  q = h2_fma(q, s, 0x42484248u);
  uint32_t r = h2_mul(q, h);
  uint32_t v;
  asm("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(v) : "r"(__float_as_uint(t0)), "r"(__float_as_uint(t1)));   // byte0(t0) | byte0(t1) << 16
  v <<= 15;
  asm("lop3.b32 %0, %1, %2, 0x80008000, 0x78;" : "=r"(r) : "r"(r), "r"(v));                              // r ^ (v & signs)
  return r;
}

template <int kPoly, bool kPacked32>
__global__ void __launch_bounds__(1024, 1) probe(int iters, float s, float* out, long long* cyc) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.37f * (float)(threadIdx.x + i);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool poly = ((i + 1) * kPoly) / 8 != (i * kPoly) / 8;
      uint32_t y;
      if (poly) y = sin2_h2<kPacked32>(x[2 * i], x[2 * i + 1]);
      else y = pack_f16x2(__sinf(x[2 * i]), __sinf(x[2 * i + 1]));
      acc ^= y;
      const float2 yf = unpack_f16x2(y);
      x[2 * i] = fmaf(yf.x, s, x[2 * i]); x[2 * i + 1] = fmaf(yf.y, s, x[2 * i + 1]);   // next "accumulator" value (stands for the tcgen05.ld)
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

// errors after rounding to the activation format: [0] fp16 polynomial, [1] MUFU -> fp16, [2] MUFU -> bf16 (shipped); max and sum of squares
__global__ void accuracy(float lo, float hi, int n, float* maxerr, double* sq) {
  float w0 = 0.f, w1 = 0.f, w2 = 0.f; double q0 = 0, q1 = 0, q2 = 0;
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = lo + (hi - lo) * (float)i / (float)n;
    const double ref = sin((double)x);
    const float2 y = unpack_f16x2(sin2_h2<false>(x, -x));
    const float e0 = fmaxf(fabsf((float)(y.x - ref)), fabsf((float)(y.y + ref)));
    const float e1 = fabsf((float)(__half2float(__float2half_rn(__sinf(x))) - ref));
    const float e2 = fabsf((float)(__bfloat162float(__float2bfloat16_rn(__sinf(x))) - ref));
    w0 = fmaxf(w0, e0); w1 = fmaxf(w1, e1); w2 = fmaxf(w2, e2);
    q0 += (double)e0 * e0; q1 += (double)e1 * e1; q2 += (double)e2 * e2;
  }
  atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(w0));
  atomicMax(reinterpret_cast<int*>(maxerr + 1), __float_as_int(w1));
  atomicMax(reinterpret_cast<int*>(maxerr + 2), __float_as_int(w2));
  atomicAdd(sq, q0); atomicAdd(sq + 1, q1); atomicAdd(sq + 2, q2);
}

template <int kPoly, bool kPacked32>
static void run(float* out, long long* cyc) {
  const int iters = 2000;
  for (int nw : {4, 8, 16, 32}) {
    probe<kPoly, kPacked32><<<1, nw * 32>>>(iters, 1e-3f, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    const double wel = (double)iters * 16 * (nw / 4.0);
    printf("M9 %d of 8 pairs fp16-poly (%s reduction), warps/SMSP %d: %.2f cycles per warp-element per SMSP (tile-layer of 128x256: %.0f cycles)\n",
           kPoly, kPacked32 ? "f32x2" : "scalar", nw / 4, (double)*cyc / wel, (double)*cyc / wel * 256.0);
  }
}

int main() {
  float* out; long long* cyc; float* err; double* sq;
  cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8); cudaMallocManaged(&err, 16); cudaMallocManaged(&sq, 32);
  for (float range : {8.f, 64.f, 512.f}) {
    const int n = 1 << 22;
    for (int i = 0; i < 3; ++i) { err[i] = 0.f; sq[i] = 0.0; }
    accuracy<<<64, 256>>>(-range, range, n, err, sq);
    cudaDeviceSynchronize();
    printf("M9 accuracy on [-%g, %g] (max abs / rms): fp16 polynomial %.3e / %.3e, sin.approx -> fp16 %.3e / %.3e, sin.approx -> bf16 %.3e / %.3e\n",
           range, range, err[0], sqrt(sq[0] / n), err[1], sqrt(sq[1] / n), err[2], sqrt(sq[2] / n));
  }
  run<0, false>(out, cyc); run<2, false>(out, cyc); run<3, false>(out, cyc); run<4, false>(out, cyc); run<5, false>(out, cyc); run<8, false>(out, cyc);
  run<3, true>(out, cyc); run<4, true>(out, cyc); run<5, true>(out, cyc);
  return 0;
}
