// fp32 parity mode of the FiLM-SIREN point MLP (volume_renderer.py:133-160) on the FP32 pipe.
// One CTA = 64 sample points through all layers; activations stay in shared memory between
// layers (transposed, [channel][point]); weights stream through a 32-row shared stage from the
// packed blob's transposed fp32 copy.  Per-point outputs (features, raw rgb, sdf) go to HBM and
// are composited by composite_fwd_kernel -- this is the unfused reference-precision path; the
// fused tensor-core path is fused_bf16_sm100.cuh.
#pragma once
#include "c3d_common.cuh"

namespace c3d {

constexpr int F32_TP = 64;        // points per CTA
constexpr int F32_LD = 68;        // padded row stride of actT (floats)
constexpr int F32_KC = 32;        // weight rows per stage
constexpr size_t F32_SMEM = sizeof(float) * ((size_t)W * F32_LD + (size_t)F32_KC * W + F32_TP * 8 + 4 * F32_TP * 4);

struct MlpF32Args {
  const uint8_t* blob; PackedLayout L;
  const float2* film; const float4* first; const float4* view;   // style_prep tables (image-indexed from img0)
  const float* pts;        // (imgs, pts_per_img, 3) world space
  const float* viewdirs;   // (imgs, n_rays, 3)
  const float* near; const float* far;  // (imgs)
  int n_samples; int pts_per_img; int tiles_per_img; int n_imgs;
  float* feat;             // (imgs*pts_per_img, 256)
  float* rgb;              // (imgs*pts_per_img, 3)
  float* sdf;              // (imgs*pts_per_img)
  float* save_acc;         // optional: (D, imgs*pts_per_img, 256) pre-FiLM accumulators of layers 1..D (backward)
  size_t save_stride;      // imgs*pts_per_img*256
};

__device__ __forceinline__ void f32_gemm_layer(const float* __restrict__ WT /*[256 k][256 c] global*/,
                                               const float* actT, float* wS, float (&acc)[8][8], int tx, int ty) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int kc = 0; kc < W; kc += F32_KC) {
    __syncthreads();   // previous stage fully consumed (also orders actT writes of the previous layer)
    const float4* src = reinterpret_cast<const float4*>(WT + (size_t)kc * W);
    float4* dst = reinterpret_cast<float4*>(wS);
#pragma unroll
    for (int i = 0; i < (F32_KC * W / 4) / 256; ++i) dst[threadIdx.x + i * 256] = __ldg(src + threadIdx.x + i * 256);
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < F32_KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(actT + (size_t)(kc + kk) * F32_LD + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(actT + (size_t)(kc + kk) * F32_LD + ty * 8 + 4);
      const float4 w0 = *reinterpret_cast<const float4*>(wS + kk * W + tx * 4);
      const float4 w1 = *reinterpret_cast<const float4*>(wS + kk * W + 128 + tx * 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }
  __syncthreads();     // all reads of actT for this layer are done -> safe to overwrite
}

__global__ void __launch_bounds__(256, 1) mlp_fp32_kernel(MlpF32Args a) {
  extern __shared__ __align__(16) float smem[];
  float* actT = smem;                           // [256][F32_LD]
  float* wS = actT + (size_t)W * F32_LD;        // [32][256]
  float* pS = wS + (size_t)F32_KC * W;          // [64][4]  normalised point
  float* vS = pS + F32_TP * 4;                  // [64][4]  view direction
  float* red = vS + F32_TP * 4;                 // [4][64][4]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int img = blockIdx.x / a.tiles_per_img;
  const int tile = blockIdx.x - img * a.tiles_per_img;
  const int p0 = tile * F32_TP;                 // first point of the tile within the image
  const int D = a.L.D;
  const size_t img_pt0 = (size_t)img * a.pts_per_img;
  const int n_rays = a.pts_per_img / a.n_samples;

  if (threadIdx.x < F32_TP) {
    const int p = min(p0 + (int)threadIdx.x, a.pts_per_img - 1);
    const float s = 2.0f / (a.far[img] - a.near[img]);         // normalize_points, nerf_utils.py:130
    const float* q = a.pts + (img_pt0 + p) * 3;
    pS[threadIdx.x * 4 + 0] = q[0] * s; pS[threadIdx.x * 4 + 1] = q[1] * s; pS[threadIdx.x * 4 + 2] = q[2] * s;
    const int ray = p / a.n_samples;
    const float* v = a.viewdirs + ((size_t)img * n_rays + ray) * 3;
    vS[threadIdx.x * 4 + 0] = v[0]; vS[threadIdx.x * 4 + 1] = v[1]; vS[threadIdx.x * 4 + 2] = v[2];
  }
  __syncthreads();

  int ch[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { ch[j] = tx * 4 + j; ch[4 + j] = 128 + tx * 4 + j; }

  // layer 0: sin(gamma0 * (W0 p + b0) + beta0), K = 3 on the FP32 pipe
  {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = a.first[(size_t)img * W + ch[j]];
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* q = pS + (ty * 8 + i) * 4;
        o[i] = sinf(fmaf(t.x, q[0], fmaf(t.y, q[1], fmaf(t.z, q[2], t.w))));
      }
      *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  float acc[8][8];
  const float* WT = reinterpret_cast<const float*>(a.blob + a.L.wT32);
  auto save = [&](int l) {
    if (!a.save_acc) return;
    float* base = a.save_acc + (size_t)(l - 1) * a.save_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = p0 + ty * 8 + i;
      if (p < a.pts_per_img) {
        float4* o = reinterpret_cast<float4*>(base + (img_pt0 + p) * W);
        o[tx] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        o[32 + tx] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
  };
  for (int l = 1; l < D; ++l) {
    f32_gemm_layer(WT + (size_t)(l - 1) * W * W, actT, wS, acc, tx, ty);
    save(l);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = a.film[((size_t)img * (D + 1) + l) * W + ch[j]];
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = sinf(fmaf(f.x, acc[i][j], f.y));
      *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  __syncthreads();
  // sdf head on h_{D-1}  (volume_renderer.py:148)
  {
    const int p = threadIdx.x & 63, q = threadIdx.x >> 6;
    const float* wsig = reinterpret_cast<const float*>(a.blob + a.L.wsig);
    float s = 0.f;
    for (int c = q * 64; c < q * 64 + 64; ++c) s = fmaf(wsig[c], actT[(size_t)c * F32_LD + p], s);
    red[(q * 64 + p) * 4] = s;
    __syncthreads();
    if (threadIdx.x < F32_TP && p0 + p < a.pts_per_img) {
      const float bsig = reinterpret_cast<const float*>(a.blob + a.L.scal)[0];
      a.sdf[img_pt0 + p0 + p] = red[p * 4] + red[(64 + p) * 4] + red[(128 + p) * 4] + red[(192 + p) * 4] + bsig;
    }
  }
  // view layer (volume_renderer.py:151-152): K = 256 through the stage + 3 view-direction columns
  f32_gemm_layer(WT + (size_t)(D - 1) * W * W, actT, wS, acc, tx, ty);
  save(D);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 f = a.film[((size_t)img * (D + 1) + D) * W + ch[j]];
    const float4 tv = a.view[(size_t)img * W + ch[j]];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* v = vS + (ty * 8 + i) * 4;
      const float vt = fmaf(tv.x, v[0], fmaf(tv.y, v[1], tv.z * v[2]));
      acc[i][j] = sinf(fmaf(f.x, acc[i][j], vt + f.y));
    }
    *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = p0 + ty * 8 + i;
    if (p < a.pts_per_img) {
      float4* o = reinterpret_cast<float4*>(a.feat + (img_pt0 + p) * W);
      o[tx] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      o[32 + tx] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
  __syncthreads();
  // rgb head (volume_renderer.py:154)
  {
    const int p = threadIdx.x & 63, q = threadIdx.x >> 6;
    const float4* wrgb = reinterpret_cast<const float4*>(a.blob + a.L.wrgb);
    float r = 0.f, g = 0.f, b = 0.f;
    for (int c = q * 64; c < q * 64 + 64; ++c) {
      const float4 w = wrgb[c];
      const float f = actT[(size_t)c * F32_LD + p];
      r = fmaf(w.x, f, r); g = fmaf(w.y, f, g); b = fmaf(w.z, f, b);
    }
    float* o = red + (q * 64 + p) * 4;
    o[0] = r; o[1] = g; o[2] = b;
    __syncthreads();
    if (threadIdx.x < F32_TP && p0 + p < a.pts_per_img) {
      const float* sc = reinterpret_cast<const float*>(a.blob + a.L.scal);
      float* out = a.rgb + (img_pt0 + p0 + p) * 3;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        out[j] = red[p * 4 + j] + red[(64 + p) * 4 + j] + red[(128 + p) * 4 + j] + red[(192 + p) * 4 + j] + sc[1 + j];
    }
  }
}

}  // namespace c3d
