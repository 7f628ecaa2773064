// Second-order path of the eikonal regulariser (training): gradients of a loss on E = d sdf / d pts
// (exp/cips3d/nerf_utils.py:220-228, create_graph=True) w.r.t. the FiLM styles and the renderer parameters.
//
// With v = dL/dE fixed, dL/dtheta = d/dtheta sum_p v_p . E_p = d/dtheta sum_p D_v sdf_p, the directional derivative of sdf
// along v.  D_v sdf is one forward-mode (tangent) sweep through the point layers, and its theta-gradient is a reverse
// sweep over the primal and the tangent chain together ("reverse over forward"), all on the FP32 pipe:
//
//   primal   a_l = gamma_l (W_l h_{l-1}) + shift_l        h_l  = sin a_l            (mlp_fp32_kernel, saves u_l = W_l h_{l-1})
//   tangent  ad_l = gamma_l (W_l hd_{l-1})                hd_l = cos a_l * ad_l     (eik_tangent_kernel, saves ud_l = W_l hd_{l-1})
//            x = s p,  xd = s v,  S = w_sigma . hd_{D-1}
//   reverse  adj(ad_l) = adj(hd_l) cos a_l
//            adj(a_l)  = adj(h_l) cos a_l - adj(hd_l) sin a_l * ad_l
//            adj(gamma_l) += adj(a_l) u_l + adj(ad_l) ud_l ,  adj(shift_l) += adj(a_l)          -> g_film, as in backward.cuh
//            adj(h_{l-1})  = W_l^T (gamma_l adj(a_l)) ,  adj(hd_{l-1}) = W_l^T (gamma_l adj(ad_l))
//            adj(W_l)     += (gamma_l adj(a_l)) h_{l-1}^T + (gamma_l adj(ad_l)) hd_{l-1}^T          -> param_grads.cuh kernels
//   (eik_bwd_kernel; the four per-layer dumps H, Hd, G = gamma adj(a), Gd = gamma adj(ad) feed the weight-gradient GEMMs)
#pragma once
#include "c3d_common.cuh"
#include "mlp_fp32.cuh"

namespace c3d {

struct EikArgs {
  const uint8_t* blob; PackedLayout L;
  const float2* film; const float4* first;         // style_prep tables (image-indexed from img0)
  const float* pts; const float* v;                // (imgs, pts_per_img, 3): world-space points, cotangent of the eikonal term
  const float* near; const float* far;
  int pts_per_img, tiles_per_img;
  const float* save_acc;  size_t save_stride;      // (D, pts, 256): u_l of layers 1..D from mlp_fp32_kernel (index l-1)
  float* save_tacc;                                // (D-1, pts, 256): ud_l of layers 1..D-1 (index l-1), same stride
  // eik_bwd_kernel only
  float* g_film;                                   // (imgs, D+1, 256, 2) atomically accumulated (G1', G2')
  float* dump_h; float* dump_hd; float* dump_g; float* dump_gd;   // (D, pts, 256) each, layers 0..D-1, same stride
};

// smem: actT [256][F32_LD] | wS [32][256] | pS [64][4] | vS [64][4]
constexpr size_t EIKT_SMEM = sizeof(float) * ((size_t)W * F32_LD + (size_t)F32_KC * W + F32_TP * 8);

__global__ void __launch_bounds__(256, 1) eik_tangent_kernel(EikArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* actT = smem;
  float* wS = actT + (size_t)W * F32_LD;
  float* pS = wS + (size_t)F32_KC * W;
  float* vS = pS + F32_TP * 4;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int img = blockIdx.x / a.tiles_per_img;
  const int p0 = (blockIdx.x - img * a.tiles_per_img) * F32_TP;
  const int D = a.L.D;
  const size_t img_pt0 = (size_t)img * a.pts_per_img;
  if (threadIdx.x < F32_TP) {
    const int p = min(p0 + (int)threadIdx.x, a.pts_per_img - 1);
    const float s = 2.0f / (a.far[img] - a.near[img]);
    const float* q = a.pts + (img_pt0 + p) * 3;
    const float* u = a.v + (img_pt0 + p) * 3;
    pS[threadIdx.x * 4 + 0] = q[0] * s; pS[threadIdx.x * 4 + 1] = q[1] * s; pS[threadIdx.x * 4 + 2] = q[2] * s;
    vS[threadIdx.x * 4 + 0] = u[0] * s; vS[threadIdx.x * 4 + 1] = u[1] * s; vS[threadIdx.x * 4 + 2] = u[2] * s;
  }
  __syncthreads();
  int ch[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { ch[j] = tx * 4 + j; ch[4 + j] = 128 + tx * 4 + j; }
  // layer 0: hd_0 = cos(a_0) * (gamma W0 . xd)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = a.first[(size_t)img * W + ch[j]];
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* q = pS + (ty * 8 + i) * 4;
      const float* d = vS + (ty * 8 + i) * 4;
      o[i] = cosf(fmaf(t.x, q[0], fmaf(t.y, q[1], fmaf(t.z, q[2], t.w)))) * fmaf(t.x, d[0], fmaf(t.y, d[1], t.z * d[2]));
    }
    *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
  float acc[8][8];
  const float* WT = reinterpret_cast<const float*>(a.blob + a.L.wT32);
  for (int l = 1; l < D; ++l) {
    f32_gemm_layer(WT + (size_t)(l - 1) * W * W, actT, wS, acc, tx, ty);      // ud_l = W_l hd_{l-1}
    const float* up = a.save_acc + (size_t)(l - 1) * a.save_stride;
    float* tp = a.save_tacc + (size_t)(l - 1) * a.save_stride;
    float2 f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = a.film[((size_t)img * (D + 1) + l) * W + ch[j]];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = min(p0 + ty * 8 + i, a.pts_per_img - 1);
      const float4* ap = reinterpret_cast<const float4*>(up + (img_pt0 + p) * W);
      const float4 u0 = ap[tx], u1 = ap[32 + tx];
      const float uv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      if (p0 + ty * 8 + i < a.pts_per_img) {
        float4* o = reinterpret_cast<float4*>(tp + (img_pt0 + p) * W);
        o[tx] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        o[32 + tx] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = cosf(fmaf(f[j].x, uv[j], f[j].y)) * f[j].x * acc[i][j];   // hd_l
    }
    if (l < D - 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
        *reinterpret_cast<float4*>(actT + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
      }
    }
  }
}

// smem: gT [256][F32_LD] | gT2 [256][F32_LD] | wS [32][256] | pS, vS | red [8][256][2]
constexpr size_t EIKB_SMEM = sizeof(float) * (2 * (size_t)W * F32_LD + (size_t)F32_KC * W + F32_TP * 8 + 8 * W * 2);

__global__ void __launch_bounds__(256, 1) eik_bwd_kernel(EikArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* gT = smem;
  float* gT2 = gT + (size_t)W * F32_LD;
  float* wS = gT2 + (size_t)W * F32_LD;
  float* pS = wS + (size_t)F32_KC * W;
  float* vS = pS + F32_TP * 4;
  float* red = vS + F32_TP * 4;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int img = blockIdx.x / a.tiles_per_img;
  const int p0 = (blockIdx.x - img * a.tiles_per_img) * F32_TP;
  const int D = a.L.D;
  const size_t img_pt0 = (size_t)img * a.pts_per_img;
  if (threadIdx.x < F32_TP) {
    const int p = min(p0 + (int)threadIdx.x, a.pts_per_img - 1);
    const float s = 2.0f / (a.far[img] - a.near[img]);
    const float* q = a.pts + (img_pt0 + p) * 3;
    const float* u = a.v + (img_pt0 + p) * 3;
    pS[threadIdx.x * 4 + 0] = q[0] * s; pS[threadIdx.x * 4 + 1] = q[1] * s; pS[threadIdx.x * 4 + 2] = q[2] * s;
    vS[threadIdx.x * 4 + 0] = u[0] * s; vS[threadIdx.x * 4 + 1] = u[1] * s; vS[threadIdx.x * 4 + 2] = u[2] * s;
  }
  __syncthreads();
  int ch[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { ch[j] = tx * 4 + j; ch[4 + j] = 128 + tx * 4 + j; }
  int pidx[8]; bool pval[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { pval[i] = p0 + ty * 8 + i < a.pts_per_img; pidx[i] = min(p0 + ty * 8 + i, a.pts_per_img - 1); }
  auto store_rows = [&](float* base, const float (&x)[8]) {   // one point row of my 8 channels (two float4)
    float4* o = reinterpret_cast<float4*>(base);
    o[tx] = make_float4(x[0], x[1], x[2], x[3]);
    o[32 + tx] = make_float4(x[4], x[5], x[6], x[7]);
  };
  auto reduce_film = [&](int l, const float2 (&cs)[8]) {      // block column sums -> one atomic pair per channel
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[(ty * W + ch[j]) * 2 + 0] = cs[j].x; red[(ty * W + ch[j]) * 2 + 1] = cs[j].y; }
    __syncthreads();
    const int c = threadIdx.x;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) { s0 += red[(w8 * W + c) * 2 + 0]; s1 += red[(w8 * W + c) * 2 + 1]; }
    float* gf = a.g_film + (((size_t)img * (D + 1) + l) * W + c) * 2;
    atomicAdd(gf + 0, s0); atomicAdd(gf + 1, s1);
  };

  float gh[8][8], gt[8][8];                         // adj(h_l), adj(hd_l) of my (point, channel) block
  {
    const float* wsig = reinterpret_cast<const float*>(a.blob + a.L.wsig);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { gh[i][j] = 0.f; gt[i][j] = pval[i] ? wsig[ch[j]] : 0.f; }
  }
  const float* Wn = reinterpret_cast<const float*>(a.blob + a.L.w32);
  for (int l = D - 1; l >= 1; --l) {
    const float* up = a.save_acc + (size_t)(l - 1) * a.save_stride;
    const float* tp = a.save_tacc + (size_t)(l - 1) * a.save_stride;
    float2 f[8], cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { f[j] = a.film[((size_t)img * (D + 1) + l) * W + ch[j]]; cs[j] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const size_t row = (img_pt0 + pidx[i]) * W;
      const float4 u0 = reinterpret_cast<const float4*>(up + row)[tx], u1 = reinterpret_cast<const float4*>(up + row)[32 + tx];
      const float4 d0 = reinterpret_cast<const float4*>(tp + row)[tx], d1 = reinterpret_cast<const float4*>(tp + row)[32 + tx];
      const float uv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      float hv[8], hdv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float sn, cn;
        sincosf(fmaf(f[j].x, uv[j], f[j].y), &sn, &cn);
        const float ad = f[j].x * dv[j];
        const float adj_ad = pval[i] ? gt[i][j] * cn : 0.f;
        const float adj_a = pval[i] ? gh[i][j] * cn - gt[i][j] * sn * ad : 0.f;
        cs[j].x += adj_a * uv[j] + adj_ad * dv[j]; cs[j].y += adj_a;
        hv[j] = sn; hdv[j] = cn * ad;
        gh[i][j] = f[j].x * adj_a;                  // G[l]  = adj(u_l)
        gt[i][j] = f[j].x * adj_ad;                 // Gd[l] = adj(ud_l)
      }
      if (pval[i]) {
        store_rows(a.dump_h + (size_t)l * a.save_stride + row, hv);
        store_rows(a.dump_hd + (size_t)l * a.save_stride + row, hdv);
        float g0[8], g1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { g0[j] = gh[i][j]; g1[j] = gt[i][j]; }
        store_rows(a.dump_g + (size_t)l * a.save_stride + row, g0);
        store_rows(a.dump_gd + (size_t)l * a.save_stride + row, g1);
      }
    }
    // transposed cotangent tiles for the two GEMMs against the un-transposed W_l
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<float4*>(gT + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(gh[0][j], gh[1][j], gh[2][j], gh[3][j]);
      *reinterpret_cast<float4*>(gT + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(gh[4][j], gh[5][j], gh[6][j], gh[7][j]);
      *reinterpret_cast<float4*>(gT2 + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(gt[0][j], gt[1][j], gt[2][j], gt[3][j]);
      *reinterpret_cast<float4*>(gT2 + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(gt[4][j], gt[5][j], gt[6][j], gt[7][j]);
    }
    reduce_film(l, cs);
    f32_gemm_layer(Wn + (size_t)(l - 1) * W * W, gT, wS, gh, tx, ty);      // adj(h_{l-1})
    f32_gemm_layer(Wn + (size_t)(l - 1) * W * W, gT2, wS, gt, tx, ty);     // adj(hd_{l-1})
  }
  // ---- layer 0: u_0 = W0 x, ud_0 = W0 xd  (K = 3)
  {
    const float4* w0 = reinterpret_cast<const float4*>(a.blob + a.L.w0);
    float2 cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = w0[ch[j]];
      const float2 f = a.film[((size_t)img * (D + 1) + 0) * W + ch[j]];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* q = pS + (ty * 8 + i) * 4;
        const float* d = vS + (ty * 8 + i) * 4;
        const float u = fmaf(w.x, q[0], fmaf(w.y, q[1], w.z * q[2]));
        const float ud = fmaf(w.x, d[0], fmaf(w.y, d[1], w.z * d[2]));
        float sn, cn;
        sincosf(fmaf(f.x, u, f.y), &sn, &cn);
        const float ad = f.x * ud;
        const float adj_ad = pval[i] ? gt[i][j] * cn : 0.f;
        const float adj_a = pval[i] ? gh[i][j] * cn - gt[i][j] * sn * ad : 0.f;
        cs[j].x += adj_a * u + adj_ad * ud; cs[j].y += adj_a;
        if (pval[i]) {
          const size_t e = (img_pt0 + pidx[i]) * W + ch[j];
          a.dump_h[e] = sn; a.dump_hd[e] = cn * ad;
          a.dump_g[e] = f.x * adj_a; a.dump_gd[e] = f.x * adj_ad;
        }
      }
    }
    reduce_film(0, cs);
  }
}

}  // namespace c3d
