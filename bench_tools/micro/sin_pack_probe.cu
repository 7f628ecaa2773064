// M8: sines of a 16-element accumulator chunk with kPoly of its 8 pairs evaluated by a PACKED polynomial (FFMA2 / FMUL2 /
// FADD2: two sines per issue slot on the FMA pipe) and the rest by FMUL + MUFU.SIN; every pair packed to bf16x2.
// Cycles per warp-element per sub-partition at 1, 2, 4, 8 warps per sub-partition, and the polynomial's accuracy.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sin_pack_probe sin_pack_probe.cu && ./sin_pack_probe
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "../../cips-3dplusplus_b200/csrc/sm100_ptx.cuh"
using namespace c3d::ptx;

// sin of two arguments (radians, |x| < 2^22 * 2 pi) with 10 packed FMA-pipe instructions: r = x / 2 pi, k = rint(r) by the
// magic-number trick, f = r - k in [-0.5, 0.5], sin(2 pi f) = f (c1 + s (c3 + s (c5 + s (c7 + s c9)))), s = f^2; |err| < 7e-6
__device__ __forceinline__ float2 sin2_poly(float2 x) {
  const float2 r = mul2(x, make_float2(0.15915494309189535f, 0.15915494309189535f));
  const float2 t = add2(r, make_float2(12582912.0f, 12582912.0f));
  const float2 k = add2(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = fma2(k, make_float2(-1.0f, -1.0f), r);
  const float2 s = mul2(f, f);
  float2 p = fma2(s, make_float2(32.78138732910156f, 32.78138732910156f), make_float2(-74.47799682617188f, -74.47799682617188f));
  p = fma2(p, s, make_float2(81.36681365966797f, 81.36681365966797f));
  p = fma2(p, s, make_float2(-41.331214904785156f, -41.331214904785156f));
  p = fma2(p, s, make_float2(6.283055782318115f, 6.283055782318115f));
  return mul2(p, f);
}

template <int kPoly>
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
      float2 y;
      // spread the polynomial pairs evenly over the chunk
      const bool poly = ((i + 1) * kPoly) / 8 != (i * kPoly) / 8;
      if (poly) y = sin2_poly(make_float2(x[2 * i], x[2 * i + 1]));
      else y = make_float2(__sinf(x[2 * i]), __sinf(x[2 * i + 1]));
      acc ^= pack_bf16x2(y.x, y.y);
      x[2 * i] = fmaf(y.x, s, x[2 * i]); x[2 * i + 1] = fmaf(y.y, s, x[2 * i + 1]);   // next "accumulator" value (stands for the tcgen05.ld)
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

__global__ void accuracy(float lo, float hi, int n, float* maxerr) {
  float worst = 0.f;
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = lo + (hi - lo) * (float)i / (float)n;
    const float2 y = sin2_poly(make_float2(x, -x));
    const double ref = sin((double)x);
    worst = fmaxf(worst, fmaxf(fabsf((float)(y.x - ref)), fabsf((float)(y.y + ref))));
  }
  atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(worst));
}
__global__ void accuracy_mufu(float lo, float hi, int n, float* maxerr) {
  float worst = 0.f;
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = lo + (hi - lo) * (float)i / (float)n;
    worst = fmaxf(worst, fabsf((float)(__sinf(x) - sin((double)x))));
  }
  atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(worst));
}

template <int kPoly>
static void run(float* out, long long* cyc) {
  const int iters = 2000;
  for (int nw : {4, 8, 16, 32}) {
    probe<kPoly><<<1, nw * 32>>>(iters, 1e-3f, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    const double wel = (double)iters * 16 * (nw / 4.0);
    printf("M8 %d of 8 pairs packed-poly, warps/SMSP %d: %.2f cycles per warp-element per SMSP (tile-layer of 128x256: %.0f cycles)\n",
           kPoly, nw / 4, (double)*cyc / wel, (double)*cyc / wel * 256.0);
  }
}

int main() {
  float* out; long long* cyc; float* err;
  cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8); cudaMallocManaged(&err, 8);
  for (float range : {8.f, 64.f, 512.f}) {
    err[0] = 0.f; err[1] = 0.f;
    accuracy<<<64, 256>>>(-range, range, 1 << 22, err);
    accuracy_mufu<<<64, 256>>>(-range, range, 1 << 22, err + 1);
    cudaDeviceSynchronize();
    printf("M8 accuracy on [-%g, %g]: packed polynomial max abs err %.3e, sin.approx %.3e\n", range, range, err[0], err[1]);
  }
  run<0>(out, cyc); run<1>(out, cyc); run<2>(out, cyc); run<3>(out, cyc); run<4>(out, cyc); run<5>(out, cyc); run<6>(out, cyc); run<8>(out, cyc);
  return 0;
}
