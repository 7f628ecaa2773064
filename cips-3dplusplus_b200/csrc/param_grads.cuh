// Gradients w.r.t. the renderer's own parameters (training / stage-2 inversion; the reference gets them from autograd
// through volume_renderer.py:15-160).  FP32 pipe, fed by mlp_bwd_kernel<true>, which leaves per layer l = 0..D
//   H[l] = h_l          (layer outputs, (pts, 256))
//   G[l] = dL/d acc_l   (cotangent of the pre-FiLM product W_l h_{l-1}, (pts, 256))
// in the workspace.  Then
//   wgrad_gemm_kernel     dW_l[out][in]  = sum_p G[l][p][out] H[l-1][p][in]              l = 1..D   (split over points)
//   head_wgrad_kernel     skinny products: dW_0 = G[0]^T pts_n, dW_view[:,256:] = G[D]^T viewdirs,
//                         dW_rgb = g_rgb^T H[D], dw_sigma = g_sdf^T H[D-1], and the two head bias sums
//   film_param_bwd_kernel everything that follows from the per-image column sums g_film = (sum g_a acc, sum g_a):
//                         FiLM layer biases, gamma / beta LinearLayer weights and biases
// d sigmoid_beta comes out of composite_bwd_kernel.  All outputs are accumulated with atomics into zeroed buffers.
#pragma once
#include "c3d_common.cuh"

namespace c3d {

constexpr int WG_PC = 32;   // points per shared-memory stage

// grid (4 output tiles of 128x128, splits); block 256.  out[m * ldo + k] += sum_p A[p][m] B[p][k]
__global__ void __launch_bounds__(256) wgrad_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         long long n_pts, int pts_per_cta, float* __restrict__ out, int ldo) {
  __shared__ __align__(16) float As[WG_PC][128];
  __shared__ __align__(16) float Bs[WG_PC][128];
  const int m0 = (blockIdx.x >> 1) * 128, k0 = (blockIdx.x & 1) * 128;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long p_begin = (long long)blockIdx.y * pts_per_cta;
  const long long p_end = p_begin + pts_per_cta < n_pts ? p_begin + pts_per_cta : n_pts;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (long long p0 = p_begin; p0 < p_end; p0 += WG_PC) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = threadIdx.x + q * 256, r = idx >> 5, c4 = idx & 31;
      const bool ok = p0 + r < p_end;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(&As[r][0])[c4] = ok ? __ldcs(reinterpret_cast<const float4*>(A + (p0 + r) * W + m0) + c4) : z;
      reinterpret_cast<float4*>(&Bs[r][0])[c4] = ok ? __ldcs(reinterpret_cast<const float4*>(B + (p0 + r) * W + k0) + c4) : z;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < WG_PC; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[r][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[r][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]), b1 = *reinterpret_cast<const float4*>(&Bs[r][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      atomicAdd(out + (size_t)m * ldo + k, acc[i][j]);
    }
  }
}

struct HeadWgradArgs {
  const float* a;      // (imgs * pts_per_img / a_div, a_cols) small left factor; NULL: a single column of ones
  int a_cols, a_div;   // a_div = n_samples when `a` is per ray
  const float* H;      // (imgs * pts_per_img, 256)
  int pts_per_img, pts_per_cta;
  const float* near; const float* far;   // optional: scale the image's contribution by 2 / (far - near)
  float* out; int so_j, so_c;            // out[j * so_j + c * so_c] += sum_p a[p][j] H[p][c]
  float* out_sum;                        // optional (a_cols): += sum_p a[p][j]
};

// grid (splits, imgs); block 256 (one thread per channel c)
__global__ void __launch_bounds__(256) head_wgrad_kernel(HeadWgradArgs g) {
  const int img = blockIdx.y, c = threadIdx.x;
  const int p_begin = blockIdx.x * g.pts_per_cta;
  const int p_end = min(p_begin + g.pts_per_cta, g.pts_per_img);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, sum[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t base = (size_t)img * g.pts_per_img;
  for (int p = p_begin; p < p_end; ++p) {
    const float h = __ldcs(g.H + (base + p) * W + c);
    const float* ap = g.a ? g.a + ((base + p) / g.a_div) * g.a_cols : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < g.a_cols) { const float v = ap ? ap[j] : 1.0f; acc[j] = fmaf(v, h, acc[j]); sum[j] += v; }
  }
  const float scale = g.near ? 2.0f / (g.far[img] - g.near[img]) : 1.0f;
  if (p_begin < p_end) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < g.a_cols) {
        atomicAdd(g.out + (size_t)j * g.so_j + (size_t)c * g.so_c, acc[j] * scale);
        if (g.out_sum && c == 0) atomicAdd(g.out_sum + j, sum[j]);
      }
  }
}

// grid (D+1, 256 channels c); block 256 (style index k).  g_film (batch, D+1, 256, 2) = (G1, G2); film = (gamma, shift).
__global__ void __launch_bounds__(256) film_param_bwd_kernel(const uint8_t* __restrict__ blob, PackedLayout L, c3d_param_grads pg,
                                                             const float* __restrict__ g_film, const float2* __restrict__ film,
                                                             const float* __restrict__ styles, int batch) {
  const int l = blockIdx.x, c = blockIdx.y, k = threadIdx.x, D = L.D;
  const float bias = reinterpret_cast<const float*>(blob + L.bias)[l * W + c];
  float accg = 0.f, accb = 0.f, s_dg = 0.f, s_g2 = 0.f, s_b = 0.f;
  for (int b = 0; b < batch; ++b) {
    const size_t row = (size_t)b * (D + 1) + l;
    const float G1 = g_film[(row * W + c) * 2 + 0], G2 = g_film[(row * W + c) * 2 + 1];
    const float dg = fmaf(bias, G2, G1);                 // d gamma: a = gamma (acc + bias) + beta
    const float s = styles[row * W + k];
    accg = fmaf(dg, s, accg); accb = fmaf(G2, s, accb);
    s_dg += dg; s_g2 += G2; s_b = fmaf(film[row * W + c].x, G2, s_b);
  }
  float* gw = l < D ? pg.pts_gamma_weight[l] : pg.views_gamma_weight;
  float* bw = l < D ? pg.pts_beta_weight[l] : pg.views_beta_weight;
  gw[(size_t)c * W + k] = 15.0f * accg;                  // gamma = 15 (Gw s + gb) + 30, beta = 0.25 (Bw s + bb)
  bw[(size_t)c * W + k] = 0.25f * accb;
  if (k == 0) {
    (l < D ? pg.pts_gamma_bias[l] : pg.views_gamma_bias)[c] = 15.0f * s_dg;
    (l < D ? pg.pts_beta_bias[l] : pg.views_beta_bias)[c] = 0.25f * s_g2;
    (l < D ? pg.pts_bias[l] : pg.views_bias)[c] = s_b;
  }
}

}  // namespace c3d
