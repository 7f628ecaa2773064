// Shared definitions for libc3dpp: packed-weight blob layout, error plumbing, small device math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/c3d_abi.h"

#define C3D_MAGIC 0x43334450u  // "C3DP"

namespace c3d {

constexpr int W = C3D_W;            // hidden width
constexpr int KCHUNK = 64;          // bf16 elements per 128-byte swizzle row
constexpr int NCHUNK = W / KCHUNK;  // 4 K-chunks per layer

// ------------------------------------------------------------------------------------------
// Packed blob (device memory owned by the caller; produced by c3d_pack_weights).
//   header   : uint32 magic, int32 D
//   w0       : float4[256]      (W0[c][0..2], b0[c])                     volume_renderer.py:57,62
//   wvdir    : float4[256]      (Wview[c][256..258], 0)                  volume_renderer.py:111-113
//   bias     : float[(D+1)*256] bias of point layer l (l<D) / view layer (l==D)
//   wsig     : float[256]       sigma_linear.weight                      volume_renderer.py:115
//   wrgb     : float4[256]      (Wrgb[0][c], Wrgb[1][c], Wrgb[2][c], 0)  volume_renderer.py:114
//   scal     : float[8]         bsig, brgb[0..2], sigmoid_beta
//   film     : per layer l<=D:  GwT[256 k][256 c], BwT[256 k][256 c], gb[256], bb[256]
//   wT32     : per layer l in 1..D: WT[256 k][256 c] fp32 (view layer: columns 0..255 only)
//   w32      : per layer l in 1..D: W[256 c][256 k] fp32, reference orientation (backward GEMMs contract over c)
//   wf16h    : per layer l in 1..D: 4 K-chunks x [256 n][64 k] fp16(W), rows of 128 B, 16-byte units
//              XOR-swizzled with (n & 7)  == UMMA K-major SWIZZLE_128B image of the smem stage
//   wf16l    : the same layout for fp16(2^11 (W - fp16(W))): with wf16h the two-way fp16 split of the fp32 weights (22 bits
//              of mantissa; the scaled low part stays in the normal range), operands of the fp32-mode tensor-core MLP
//   heads16  : 4 K-chunks x [16 n][64 k] fp16, same swizzle: rows 0..2 = Wrgb, row 4 / 5 = hi / lo fp16 split
//              of sigma_linear.weight, other rows zero (one N=16 MMA serves the rgb and the sdf head)
//   w0img    : "wk16" [256 n][16 k] bf16, UMMA K-major no-swizzle (8x8 core matrices: 16 B per row, 128 B per
//              K-block, 256 B per 8 rows): the K=16 side operand of the layer-0 and view-layer MMAs.
//              k-slots 4j..4j+3 (j = x,y,z) = (hi, hi, lo, hi) of W0[n][j]  x  point tile (hi, mid, hi, lo) of the
//              normalised coordinate (split product exact to ~2^-18); slots 12..14 = bf16(Wview[n][256+j]) x
//              view-direction tile; slot 15 zero.  The point tile zeroes 12..15, the view tile zeroes 0..11.
//   wbf16T   : per layer l in 1..D: 4 chunks x [256 k][64 c] bf16, same swizzle: W_l transposed (rows = input channel k,
//              contraction over the output channel c) -- the A operand of the backward MMAs  g_h = W^T g_acc
//   bwd16    : three backward side images:
//              [0] K16 image [256 c][16]: slots 0..2 and 3..5 = bf16(Wrgb[j][c]) (pair with hi / lo of g_rgb), rest zero
//              [1] K16 image [256 k][16]: slots 6, 7 = bf16(sigma_linear.weight[k]) (pair with hi / lo of g_sdf), rest zero
//              [2] heads image 4 chunks x [16 n][64 c] (sw128): rows 0..2 = W0[c][j], rows 4..6 = Wview[c][256+j]
//
// 16-bit operand formats of the tensor-core ("bf16") mode.  Every product that reads the hidden-layer activation tile --
// the layer GEMMs, the heads, the compositing -- takes IEEE half operands (tcgen05 kind::f16 with a_format = b_format = F16:
// the same rate as bf16): the activations are sines in [-1, 1] and the weights are O(1e-2 .. 1), so fp16's range is ample and
// its 11-bit mantissa cuts the rounding of both operands by 8x (feature_map rel-L2 vs the reference 1.0e-2 -> 8e-4 at D = 8,
// style gradients 3.2e-2 -> 3e-3; tests/test_operand_format_cpu.py reproduces both figures on the CPU).  The K = 16 side
// products (coordinates, view directions, FiLM shift: split hi/lo values of wide range) and the whole backward (cotangents of
// unbounded range) stay bfloat16.
// ------------------------------------------------------------------------------------------
struct PackedLayout {
  size_t w0, wvdir, bias, wsig, wrgb, scal, film, wT32, w32, wf16h, rgb16, w0img, wbf16T, bwd16, wf16l, total;
  int D;
};
constexpr size_t FILM_LAYER_FLOATS = 2 * (size_t)W * W + 2 * W;
constexpr size_t WBF16_LAYER_BYTES = (size_t)W * W * 2;      // 131072
constexpr size_t WBF16_CHUNK_BYTES = (size_t)W * KCHUNK * 2; // 32768
constexpr size_t RGB16_BYTES = (size_t)16 * W * 2;           // 8192 (heads16)
constexpr size_t W0IMG_BYTES = (size_t)W * 16 * 2;           // 8192

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline PackedLayout packed_layout(int D) {
  PackedLayout L;
  L.D = D;
  size_t o = 256;  // header
  L.w0 = o;    o += sizeof(float4) * W;
  L.wvdir = o; o += sizeof(float4) * W;
  L.bias = o;  o += sizeof(float) * (size_t)(D + 1) * W;
  L.wsig = o;  o += sizeof(float) * W;
  L.wrgb = o;  o += sizeof(float4) * W;
  L.scal = o;  o += 256;
  L.film = o;  o += sizeof(float) * FILM_LAYER_FLOATS * (size_t)(D + 1);
  L.wT32 = o;  o += sizeof(float) * (size_t)W * W * (size_t)D;
  L.w32 = o;   o += sizeof(float) * (size_t)W * W * (size_t)D;
  o = align_up(o, 1024);
  L.wf16h = o; o += WBF16_LAYER_BYTES * (size_t)D;
  L.rgb16 = o; o += RGB16_BYTES;
  L.w0img = o; o += W0IMG_BYTES;
  L.wbf16T = o; o += WBF16_LAYER_BYTES * (size_t)D;
  L.bwd16 = o; o += 3 * W0IMG_BYTES;
  o = align_up(o, 1024);
  L.wf16l = o; o += WBF16_LAYER_BYTES * (size_t)D;
  L.total = align_up(o, 1024);
  return L;
}

// byte offset of element (n, k) inside one K-chunk image with `rows` rows (UMMA K-major SW128)
__host__ __device__ inline uint32_t sw128_offset(int n, int k /*0..63*/) {
  return (uint32_t)n * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7u)) << 4) + (((uint32_t)k & 7u) << 1);
}

// byte offset of element (r, k<16) of a K=16 operand in the UMMA K-major no-swizzle ("interleave") layout
__host__ __device__ inline uint32_t k16_offset(int r, int k) {
  return ((uint32_t)r >> 3) * 256u + ((uint32_t)k >> 3) * 128u + ((uint32_t)r & 7u) * 16u + ((uint32_t)k & 7u) * 2u;
}

// ------------------------------------------------------------------------------------------
// error plumbing (thread-local message; no exceptions across the ABI)
// ------------------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern thread_local int g_launches;
int fail(int code, const char* fmt, ...);
#define C3D_CHECK_ARG(cond, ...) do { if (!(cond)) return c3d::fail(C3D_ERR_ARG, __VA_ARGS__); } while (0)
#define C3D_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
  return c3d::fail(C3D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define C3D_LAUNCH_CHECK() do { c3d::g_launches++; C3D_CUDA(cudaGetLastError()); } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

// Ray set-up shared by every kernel that generates rays in-kernel (nerf_utils.py:39-63).
struct RayGeom {
  float ox, oy, oz;   // origin (c2w[:, 3])
  float dx, dy, dz;   // un-normalised world direction
  float vx, vy, vz;   // view direction (normalised d or d_cam)
  float dnorm;        // |d|
};
__device__ __forceinline__ RayGeom make_ray(const float* __restrict__ pose /*3x4*/, float focal, int img_size,
                                            int ray, bool static_viewdirs) {
  const int iy = ray / img_size, ix = ray - iy * img_size;
  const float half = 0.5f * (float)img_size;
  const float cx = ((float)ix + 0.5f - half) / focal;
  const float cy = -((float)iy + 0.5f - half) / focal;
  const float cz = -1.0f;
  RayGeom r;
  // products and sums rounded separately, as (d_cam * R).sum(-1) rounds them (nerf_utils.py:49), and -- being explicit --
  // identically in every kernel that inlines this function (no compiler-chosen FMA contraction)
  r.dx = __fadd_rn(__fadd_rn(__fmul_rn(cx, pose[0]), __fmul_rn(cy, pose[1])), __fmul_rn(cz, pose[2]));
  r.dy = __fadd_rn(__fadd_rn(__fmul_rn(cx, pose[4]), __fmul_rn(cy, pose[5])), __fmul_rn(cz, pose[6]));
  r.dz = __fadd_rn(__fadd_rn(__fmul_rn(cx, pose[8]), __fmul_rn(cy, pose[9])), __fmul_rn(cz, pose[10]));
  r.ox = pose[3]; r.oy = pose[7]; r.oz = pose[11];
  r.dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r.dx, r.dx), __fmul_rn(r.dy, r.dy)), __fmul_rn(r.dz, r.dz)));
  float sx = static_viewdirs ? cx : r.dx, sy = static_viewdirs ? cy : r.dy, sz = static_viewdirs ? cz : r.dz;
  float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz)));
  float inv = 1.0f / fmaxf(n, 1e-12f);
  r.vx = sx * inv; r.vy = sy * inv; r.vz = sz * inv;
  return r;
}
// sample depth k of N (nerf_utils.py:97-119, offset sampling): near*(1-t)+far*t (+ u*(far-near)/N)
__device__ __forceinline__ float sample_depth(float near, float far, int k, int N, float u) {
  // explicit, separately rounded operations: the same depth in every kernel, rounded as the reference's tensor ops round it
  const float step = (1.0f - 1.0f / (float)N) / (float)(N - 1 > 0 ? N - 1 : 1);
  const float t = __fmul_rn((float)k, step);
  const float z = __fadd_rn(__fmul_rn(near, 1.0f - t), __fmul_rn(far, t));
  if (u == 0.0f) return z;
  // perturb: lower + (upper-lower)*u with upper = z_{k+1} (far for the last sample)
  const float t1 = __fmul_rn((float)(k + 1), step);
  const float zu = (k + 1 < N) ? __fadd_rn(__fmul_rn(near, 1.0f - t1), __fmul_rn(far, t1)) : far;
  return __fadd_rn(z, __fmul_rn(zu - z, u));
}

}  // namespace c3d
