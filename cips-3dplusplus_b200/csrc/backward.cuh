// Backward of the NeRF branch (autograd of VolumeFeatureRenderer.forward, used by flip inversion:
// exp/cips3d/models/projector_v9.py:1143).  fp32 on the FP32 pipe; gradients w.r.t. styles (through FiLM
// gamma / beta), pts, rays_d (through |d| in the sample spacing), viewdirs.
//
//   mlp_fp32_kernel (save_acc)  forward recompute, keeps the pre-FiLM accumulators of layers 1..D in HBM
//   composite_bwd_kernel        d(volume_integration): per-point cotangents of rgb / sdf, compositing weights,
//                               d pts (through xyz / depth), d rays_d
//   mlp_bwd_kernel              64-point tiles back through the layers (GEMMs against the un-transposed weights),
//                               column sums for d gamma / d beta, d pts through layer 0, d viewdirs
//   film_bwd_kernel             (d gamma, d beta) -> d styles through the FiLM linears
//   raygen_bwd_kernel           POSES entry: (d pts, d rays_d, d viewdirs) -> d cam_poses, d focal
//
// Math: with w_k = alpha_k T_k, om_k = 1 - alpha_k + 1e-10,
//   dL/dalpha_k = gw_k T_k - (sum_{m>k} gw_m w_m) / om_k ,  gw_k = dL/dw_k
//   alpha = 1 - exp(-sigma delta) , sigma = s(-sdf/beta)/beta  =>  dalpha/dsdf = -delta (1-alpha) s (1-s) / beta^2
#pragma once
#include "c3d_common.cuh"
#include "kernels_aux.cuh"
#include "mlp_fp32.cuh"

namespace c3d {

struct CompositeBwdArgs {
  long long n_rays; int n_samples; int n_feat;
  const float* sigmoid_beta_ptr;
  const float* rgb; const float* sdf; const float* features; const float* z_vals; const float* rays_d; const float* pts;
  const float* g_rgb_map; const float* g_feature_map; const float* g_xyz; const float* g_mask; const float* g_sdf_in;
  const float* gdot;  // optional (n_rays, N): precomputed g_feature_map . features (tensor-core path); overrides `features`
  float* weights;     // (n_rays, N) out
  float* g_rgb;       // (n_rays, N, 3) out
  float* g_sdf;       // (n_rays, N) out
  float* g_features;  // (n_rays, N, n_feat) out or NULL (the MLP backward forms w * g_feature_map itself)
  float* g_pts;       // (n_rays, N, 3) out (overwritten)
  float* g_rays_d;    // (n_rays, 3) out (overwritten)
  float* g_beta;      // optional scalar: d sigmoid_beta, atomically accumulated
};

__global__ void __launch_bounds__(256) composite_bwd_kernel(CompositeBwdArgs p) {
  __shared__ float s_w[8][CMP_MAX_N], s_T[8][CMP_MAX_N], s_gw[8][CMP_MAX_N], s_a[8][CMP_MAX_N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * 8 + warp;
  if (ray >= p.n_rays) return;
  const int N = p.n_samples;
  const float beta = *p.sigmoid_beta_ptr, inv_beta = 1.0f / beta;
  const float* rd = p.rays_d + ray * 3;
  const float dnorm = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
  const float* z = p.z_vals + ray * N;
  const float* sdf = p.sdf + ray * N;
  float* W_ = s_w[warp]; float* T_ = s_T[warp]; float* GW = s_gw[warp]; float* AL = s_a[warp];
  // ---- forward recompute: alpha, T, w, xyz
  float carry = 1.0f, ax = 0.f, ay = 0.f, az = 0.f;
  for (int k0 = 0; k0 < N; k0 += 32) {
    const int k = k0 + lane;
    float one_minus = 1.0f, alpha = 0.f;
    if (k < N) {
      const float dist = (k + 1 < N ? z[k + 1] - z[k] : 1e10f) * dnorm;
      alpha = alpha_from_sdf<true>(sdf[k], inv_beta, dist);
      one_minus = 1.0f - alpha + 1e-10f;
    }
    float total;
    const float T = carry * warp_excl_prod(one_minus, lane, total);
    carry *= total;
    if (k < N) {
      const float w = alpha * T;
      W_[k] = w; T_[k] = T; AL[k] = alpha;
      const float* q = p.pts + (ray * N + k) * 3;
      ax = fmaf(w, q[0], ax); ay = fmaf(w, q[1], ay); az = fmaf(w, q[2], az);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
  }
  // cotangent of xyz including the depth path: depth = -|xyz|  (nerf_utils.py:335)
  float gx = p.g_xyz ? p.g_xyz[ray * 3 + 0] : 0.f, gy = p.g_xyz ? p.g_xyz[ray * 3 + 1] : 0.f, gz = p.g_xyz ? p.g_xyz[ray * 3 + 2] : 0.f;
  const float g_m0 = p.g_mask ? p.g_mask[ray * 2 + 0] : 0.f, g_depth = p.g_mask ? p.g_mask[ray * 2 + 1] : 0.f;
  const float nrm = sqrtf(ax * ax + ay * ay + az * az);
  if (nrm > 0.f) { const float c = -g_depth / nrm; gx = fmaf(c, ax, gx); gy = fmaf(c, ay, gy); gz = fmaf(c, az, gz); }
  const float gr0 = p.g_rgb_map ? p.g_rgb_map[ray * 3 + 0] : 0.f, gr1 = p.g_rgb_map ? p.g_rgb_map[ray * 3 + 1] : 0.f,
              gr2 = p.g_rgb_map ? p.g_rgb_map[ray * 3 + 2] : 0.f;
  __syncwarp();
  // ---- gw_k = dL/dw_k : feature dot products with lanes over channels
  const int C4 = p.n_feat >> 2;
  if (p.gdot || !(p.g_feature_map && p.features)) {          // precomputed (tensor-core path) or no cotangent on feature_map
    for (int k = lane; k < N; k += 32) GW[k] = p.gdot ? p.gdot[ray * N + k] : 0.f;
  } else {
    for (int k = 0; k < N; ++k) {
      float acc = 0.f;
      const float4* f = reinterpret_cast<const float4*>(p.features) + ((size_t)ray * N + k) * C4;
      const float4* g = reinterpret_cast<const float4*>(p.g_feature_map) + (size_t)ray * C4;
      for (int c = lane; c < C4; c += 32) {
        const float4 a = __ldcs(f + c), b = g[c];
        acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) GW[k] = acc;
    }
  }
  __syncwarp();
  float g_dn = 0.f, g_b = 0.f;
  for (int k0 = 0; k0 < N; k0 += 32) {
    const int k = k0 + lane;
    if (k < N) {
      const float* c = p.rgb + (ray * N + k) * 3;
      const float* q = p.pts + (ray * N + k) * 3;
      const float s0 = sigmoid_precise(c[0]), s1 = sigmoid_precise(c[1]), s2 = sigmoid_precise(c[2]);
      const float gw = GW[k] + 2.0f * (gr0 * s0 + gr1 * s1 + gr2 * s2) + gx * q[0] + gy * q[1] + gz * q[2] + (k == N - 1 ? g_m0 : 0.f);
      GW[k] = gw;
      const float w = W_[k];
      float* go = p.g_rgb + (ray * N + k) * 3;
      go[0] = 2.0f * gr0 * w * s0 * (1.0f - s0); go[1] = 2.0f * gr1 * w * s1 * (1.0f - s1); go[2] = 2.0f * gr2 * w * s2 * (1.0f - s2);
      float* gp = p.g_pts + (ray * N + k) * 3;
      gp[0] = w * gx; gp[1] = w * gy; gp[2] = w * gz;
      p.weights[ray * N + k] = w;
    }
  }
  __syncwarp();
  for (int k0 = 0; k0 < N; k0 += 32) {
    const int k = k0 + lane;
    if (k < N) {
      float S = 0.f;                                  // sum_{m>k} gw_m w_m
      for (int m = k + 1; m < N; ++m) S = fmaf(GW[m], W_[m], S);
      const float alpha = AL[k], om = 1.0f - alpha + 1e-10f;
      const float g_alpha = GW[k] * T_[k] - S / om;
      const float s = sigmoid_precise(-sdf[k] * inv_beta);
      const float sigma = s * inv_beta;
      const float dz = (k + 1 < N) ? z[k + 1] - z[k] : 1e10f;
      const float one_m_alpha = 1.0f - alpha;         // exactly 0 for the last sample (delta = 1e10)
      const float g_sigma = g_alpha * (dz * dnorm) * one_m_alpha;
      p.g_sdf[ray * N + k] = -g_sigma * s * (1.0f - s) * inv_beta * inv_beta + (p.g_sdf_in ? p.g_sdf_in[ray * N + k] : 0.f);
      g_dn = fmaf(g_alpha * sigma, one_m_alpha * dz, g_dn);
      // sigma = s(-sdf/beta)/beta  =>  d sigma / d beta = (s (1-s) sdf / beta - s) / beta^2
      g_b = fmaf(g_sigma, (s * (1.0f - s) * sdf[k] * inv_beta - s) * inv_beta * inv_beta, g_b);
      if (p.g_features) {
        // handled below (needs all lanes)
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { g_dn += __shfl_xor_sync(0xffffffffu, g_dn, o); g_b += __shfl_xor_sync(0xffffffffu, g_b, o); }
  if (lane == 0 && p.g_beta) atomicAdd(p.g_beta, g_b);
  if (lane == 0) {
    const float c = dnorm > 0.f ? g_dn / dnorm : 0.f;
    p.g_rays_d[ray * 3 + 0] = c * rd[0]; p.g_rays_d[ray * 3 + 1] = c * rd[1]; p.g_rays_d[ray * 3 + 2] = c * rd[2];
  }
  if (p.g_features && p.g_feature_map) {
    float4* go = reinterpret_cast<float4*>(p.g_features) + (size_t)ray * N * C4;
    const float4* g = reinterpret_cast<const float4*>(p.g_feature_map) + (size_t)ray * C4;
    for (int k = 0; k < N; ++k) {
      const float w = W_[k];
      for (int c = lane; c < C4; c += 32) { const float4 b = g[c]; go[(size_t)k * C4 + c] = make_float4(w * b.x, w * b.y, w * b.z, w * b.w); }
    }
  }
}

// ------------------------------------------------------------------------------------------
struct MlpBwdArgs {
  const uint8_t* blob; PackedLayout L;
  const float2* film; const float4* view;          // image-indexed from img0
  const float* pts; const float* viewdirs; const float* near; const float* far;
  int n_samples; int pts_per_img; int tiles_per_img;
  const float* save_acc; size_t save_stride;       // (D, pts, 256) accumulators of layers 1..D
  const float* weights;                            // (pts) compositing weights
  const float* g_feature_map;                      // (imgs, n_rays, 256) or NULL
  const float* g_rgb;                              // (pts, 3)
  const float* g_sdf;                              // (pts)
  float* g_film;                                   // (imgs, D+1, 256, 2): (sum g_a*acc, sum g_a), atomically accumulated
  float* g_pts;                                    // (pts, 3) += through layer 0
  float* g_viewdirs;                               // (imgs, n_rays, 3) atomically accumulated, or NULL
  float* dump_h; float* dump_g; size_t dump_stride; // kDump: (D+1, pts, 256) layer outputs / accumulator cotangents
};

// smem: gT [256][F32_LD] (cotangent of the pre-FiLM accumulators, transposed) | wS [32][256] | pS, vS | red [8][256][2]
constexpr size_t BWD_SMEM = sizeof(float) * ((size_t)W * F32_LD + (size_t)F32_KC * W + F32_TP * 8 + 8 * W * 2);

template <bool kDump>
__global__ void __launch_bounds__(256, 1) mlp_bwd_kernel(MlpBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* gT = smem;
  float* wS = gT + (size_t)W * F32_LD;
  float* pS = wS + (size_t)F32_KC * W;
  float* vS = pS + F32_TP * 4;
  float* red = vS + F32_TP * 4;                   // [8 warps][256 channels][2]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int img = blockIdx.x / a.tiles_per_img;
  const int tile = blockIdx.x - img * a.tiles_per_img;
  const int p0 = tile * F32_TP;
  const int D = a.L.D;
  const size_t img_pt0 = (size_t)img * a.pts_per_img;
  const int n_rays = a.pts_per_img / a.n_samples;
  const float nscale = 2.0f / (a.far[img] - a.near[img]);

  if (threadIdx.x < F32_TP) {
    const int p = min(p0 + (int)threadIdx.x, a.pts_per_img - 1);
    const float* q = a.pts + (img_pt0 + p) * 3;
    pS[threadIdx.x * 4 + 0] = q[0] * nscale; pS[threadIdx.x * 4 + 1] = q[1] * nscale; pS[threadIdx.x * 4 + 2] = q[2] * nscale;
    const float* v = a.viewdirs + ((size_t)img * n_rays + p / a.n_samples) * 3;
    vS[threadIdx.x * 4 + 0] = v[0]; vS[threadIdx.x * 4 + 1] = v[1]; vS[threadIdx.x * 4 + 2] = v[2];
  }
  __syncthreads();
  int ch[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { ch[j] = tx * 4 + j; ch[4 + j] = 128 + tx * 4 + j; }
  int pidx[8]; bool pval[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { pval[i] = p0 + ty * 8 + i < a.pts_per_img; pidx[i] = min(p0 + ty * 8 + i, a.pts_per_img - 1); }

  // kDump: row-major (point, 256) stores of my 8x8 block, two float4 per point
  auto dump_rows = [&](float* base, const float (&v)[8][8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (pval[i]) {
        float4* o = reinterpret_cast<float4*>(base + (img_pt0 + pidx[i]) * W);
        o[tx] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        o[32 + tx] = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
      }
  };
  float g[8][8];                                   // cotangent of the layer's output h_l for my (point, channel) block
  // ---- view layer output cotangent: g_f = w * g_feature_map[ray] + Wrgb^T g_rgb
  {
    const float4* wrgb = reinterpret_cast<const float4*>(a.blob + a.L.wrgb);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const size_t pt = img_pt0 + pidx[i];
      const float w = pval[i] ? a.weights[pt] : 0.f;
      const float gr = pval[i] ? a.g_rgb[pt * 3 + 0] : 0.f, gg = pval[i] ? a.g_rgb[pt * 3 + 1] : 0.f, gb = pval[i] ? a.g_rgb[pt * 3 + 2] : 0.f;
      const float* gF = a.g_feature_map ? a.g_feature_map + ((size_t)img * n_rays + pidx[i] / a.n_samples) * W : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 wr = wrgb[ch[j]];
        g[i][j] = (gF ? w * gF[ch[j]] : 0.f) + wr.x * gr + wr.y * gg + wr.z * gb;
      }
    }
  }
  const float* Wn = reinterpret_cast<const float*>(a.blob + a.L.w32);
  for (int l = D; l >= 1; --l) {
    // ---- through sin and FiLM of layer l:  a = scale*acc + shift (+ view term), g_a = g * cos(a)
    const float* accp = a.save_acc + (size_t)(l - 1) * a.save_stride;
    float2 cs[8];                                  // per-channel (sum g_a*acc, sum g_a) over my 8 points
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = make_float2(0.f, 0.f);
    float gv[8][3];
#pragma unroll
    for (int i = 0; i < 8; ++i) gv[i][0] = gv[i][1] = gv[i][2] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4* ap = reinterpret_cast<const float4*>(accp + (img_pt0 + pidx[i]) * W);
      const float4 a0 = ap[tx], a1 = ap[32 + tx];
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float* v = vS + (ty * 8 + i) * 4;
      float hv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = a.film[((size_t)img * (D + 1) + l) * W + ch[j]];
        float accv = av[j], arg;
        if (l == D) {
          const float4 wv = reinterpret_cast<const float4*>(a.blob + a.L.wvdir)[ch[j]];
          accv += fmaf(wv.x, v[0], fmaf(wv.y, v[1], wv.z * v[2]));            // acc + Wv . v ; d a / d gamma = this + b
          arg = fmaf(f.x, accv, f.y);
          float sn, cn;
          if (kDump) { sincosf(arg, &sn, &cn); hv[j] = sn; } else cn = cosf(arg);
          const float ga = pval[i] ? g[i][j] * cn : 0.f;
          const float gs = ga * f.x;
          gv[i][0] = fmaf(gs, wv.x, gv[i][0]); gv[i][1] = fmaf(gs, wv.y, gv[i][1]); gv[i][2] = fmaf(gs, wv.z, gv[i][2]);
          cs[j].x = fmaf(ga, accv, cs[j].x); cs[j].y += ga;
          g[i][j] = ga * f.x;
        } else {
          arg = fmaf(f.x, accv, f.y);
          float sn, cn;
          if (kDump) { sincosf(arg, &sn, &cn); hv[j] = sn; } else cn = cosf(arg);
          const float ga = pval[i] ? g[i][j] * cn : 0.f;
          cs[j].x = fmaf(ga, accv, cs[j].x); cs[j].y += ga;
          g[i][j] = ga * f.x;                       // cotangent of acc_l
        }
      }
      if (kDump && pval[i]) {                       // h_l row of this point
        float4* o = reinterpret_cast<float4*>(a.dump_h + (size_t)l * a.dump_stride + (img_pt0 + pidx[i]) * W);
        o[tx] = make_float4(hv[0], hv[1], hv[2], hv[3]);
        o[32 + tx] = make_float4(hv[4], hv[5], hv[6], hv[7]);
      }
    }
    if (kDump) dump_rows(a.dump_g + (size_t)l * a.dump_stride, g);
    if (l == D && a.g_viewdirs) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          float s = gv[i][q];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (tx == 0 && pval[i]) atomicAdd(a.g_viewdirs + ((size_t)img * n_rays + pidx[i] / a.n_samples) * 3 + q, s);
        }
      }
    }
    // column sums -> g_film (block reduction over the 8 warps, then one atomic per channel)
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[(ty * W + ch[j]) * 2 + 0] = cs[j].x; red[(ty * W + ch[j]) * 2 + 1] = cs[j].y; }
    // transposed cotangent tile for the GEMM
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<float4*>(gT + (size_t)ch[j] * F32_LD + ty * 8) = make_float4(g[0][j], g[1][j], g[2][j], g[3][j]);
      *reinterpret_cast<float4*>(gT + (size_t)ch[j] * F32_LD + ty * 8 + 4) = make_float4(g[4][j], g[5][j], g[6][j], g[7][j]);
    }
    __syncthreads();
    {
      const int c = threadIdx.x;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { s0 += red[(w8 * W + c) * 2 + 0]; s1 += red[(w8 * W + c) * 2 + 1]; }
      float* gf = a.g_film + (((size_t)img * (D + 1) + l) * W + c) * 2;
      atomicAdd(gf + 0, s0); atomicAdd(gf + 1, s1);
    }
    // ---- g_h_{l-1}[p][k] = sum_c g_acc[p][c] * W_l[c][k]   (+ sigma head on h_{D-1})
    f32_gemm_layer(Wn + (size_t)(l - 1) * W * W, gT, wS, g, tx, ty);
    if (l == D) {
      const float* wsig = reinterpret_cast<const float*>(a.blob + a.L.wsig);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gs = pval[i] ? a.g_sdf[img_pt0 + pidx[i]] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i][j] = fmaf(wsig[ch[j]], gs, g[i][j]);
      }
    }
  }
  // ---- layer 0: a_0 = gamma0 * (W0 p + b0) + beta0
  {
    const float4* w0 = reinterpret_cast<const float4*>(a.blob + a.L.w0);
    float2 cs[8];
    float gp[8][3];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) gp[i][0] = gp[i][1] = gp[i][2] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = w0[ch[j]];
      const float2 f = a.film[((size_t)img * (D + 1) + 0) * W + ch[j]];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* q = pS + (ty * 8 + i) * 4;
        const float accv = fmaf(w.x, q[0], fmaf(w.y, q[1], w.z * q[2]));
        float sn, cn;
        if (kDump) sincosf(fmaf(f.x, accv, f.y), &sn, &cn); else cn = cosf(fmaf(f.x, accv, f.y));
        const float ga = pval[i] ? g[i][j] * cn : 0.f;
        cs[j].x = fmaf(ga, accv, cs[j].x); cs[j].y += ga;
        const float gs = ga * f.x;
        if (kDump && pval[i]) {
          a.dump_h[(img_pt0 + pidx[i]) * W + ch[j]] = sn;
          a.dump_g[(img_pt0 + pidx[i]) * W + ch[j]] = gs;
        }
        gp[i][0] = fmaf(gs, w.x, gp[i][0]); gp[i][1] = fmaf(gs, w.y, gp[i][1]); gp[i][2] = fmaf(gs, w.z, gp[i][2]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[(ty * W + ch[j]) * 2 + 0] = cs[j].x; red[(ty * W + ch[j]) * 2 + 1] = cs[j].y; }
    __syncthreads();
    {
      const int c = threadIdx.x;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { s0 += red[(w8 * W + c) * 2 + 0]; s1 += red[(w8 * W + c) * 2 + 1]; }
      float* gf = a.g_film + (((size_t)img * (D + 1) + 0) * W + c) * 2;
      atomicAdd(gf + 0, s0); atomicAdd(gf + 1, s1);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float s = gp[i][q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tx == 0 && pval[i]) a.g_pts[(img_pt0 + pidx[i]) * 3 + q] += s * nscale;
      }
    }
  }
}

// g_styles[b][l][k] = 15 * sum_c g_gamma[c] Gw[c][k] + 0.25 * sum_c g_beta[c] Bw[c][k],
// g_gamma = G1 + bias * G2, g_beta = G2.   grid (D+1, batch), 256 threads (k).
__global__ void __launch_bounds__(256) film_bwd_kernel(const uint8_t* __restrict__ blob, PackedLayout L,
                                                        const float* __restrict__ g_film, float* __restrict__ g_styles) {
  __shared__ float gg[W], gb[W];
  const int l = blockIdx.x, b = blockIdx.y, D = L.D, k = threadIdx.x;
  const float* gf = g_film + (((size_t)b * (D + 1) + l) * W) * 2;
  const float bias = reinterpret_cast<const float*>(blob + L.bias)[l * W + k];
  gg[k] = 15.0f * (gf[k * 2 + 0] + bias * gf[k * 2 + 1]);
  gb[k] = 0.25f * gf[k * 2 + 1];
  __syncthreads();
  const float* f = reinterpret_cast<const float*>(blob + L.film) + (size_t)l * FILM_LAYER_FLOATS;
  const float4* GwT = reinterpret_cast<const float4*>(f + (size_t)k * W);              // row k: Gw[c][k] over c
  const float4* BwT = reinterpret_cast<const float4*>(f + (size_t)W * W + (size_t)k * W);
  float s = 0.f;
  for (int c4 = 0; c4 < W / 4; ++c4) {
    const float4 gw = GwT[c4], bw = BwT[c4];
    s = fmaf(gg[c4 * 4 + 0], gw.x, fmaf(gg[c4 * 4 + 1], gw.y, fmaf(gg[c4 * 4 + 2], gw.z, fmaf(gg[c4 * 4 + 3], gw.w, s))));
    s = fmaf(gb[c4 * 4 + 0], bw.x, fmaf(gb[c4 * 4 + 1], bw.y, fmaf(gb[c4 * 4 + 2], bw.z, fmaf(gb[c4 * 4 + 3], bw.w, s))));
  }
  g_styles[((size_t)b * (D + 1) + l) * W + k] = s;
}

// POSES entry: chain (d pts, d rays_d, d viewdirs) through Render.get_rays_in_world / get_points
// (nerf_utils.py:39-66,160-161).  One thread per ray; block reduction, then atomics into (b,3,4) and (b).
__global__ void __launch_bounds__(128) raygen_bwd_kernel(c3d_raygen_params p, const float* __restrict__ g_pts,
                                                          const float* __restrict__ g_rays_d, const float* __restrict__ g_viewdirs,
                                                          float* __restrict__ g_pose, float* __restrict__ g_focal) {
  __shared__ float red[4][13];
  const int hw = p.img_size * p.img_size;
  const int b = blockIdx.y;
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  float out[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) out[i] = 0.f;
  if (ray < hw) {
    const long long gid = (long long)b * hw + ray;
    const float* pose = p.cam_poses + (size_t)b * 12;
    const float f = p.focal[b];
    const int iy = ray / p.img_size, ix = ray - iy * p.img_size;
    const float half = 0.5f * (float)p.img_size;
    const float c[3] = {((float)ix + 0.5f - half) / f, -((float)iy + 0.5f - half) / f, -1.0f};
    const RayGeom r = make_ray(pose, f, p.img_size, ray, p.static_viewdirs != 0);
    const float u = p.ray_offset ? p.ray_offset[gid] : 0.f;
    float gd[3] = {g_rays_d[gid * 3 + 0], g_rays_d[gid * 3 + 1], g_rays_d[gid * 3 + 2]}, go[3] = {0.f, 0.f, 0.f};
    const int N = p.n_samples;
    for (int k = 0; k < N; ++k) {
      const float z = sample_depth(p.near[b], p.far[b], k, N, u);
      const float* gp = g_pts + (gid * N + k) * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) { go[i] += gp[i]; gd[i] = fmaf(gp[i], z, gd[i]); }
    }
    float gc[3] = {0.f, 0.f, 0.f};                  // cotangent of the camera-frame direction
    {
      // viewdirs = x / max(|x|, eps), x = d (or d_cam when static): J^T g = (g - v (v.g)) / |x|
      const float gvx = g_viewdirs[gid * 3 + 0], gvy = g_viewdirs[gid * 3 + 1], gvz = g_viewdirs[gid * 3 + 2];
      const float dot = r.vx * gvx + r.vy * gvy + r.vz * gvz;
      const float sx = p.static_viewdirs ? c[0] : r.dx, sy = p.static_viewdirs ? c[1] : r.dy, sz = p.static_viewdirs ? c[2] : r.dz;
      const float inv = 1.0f / fmaxf(sqrtf(sx * sx + sy * sy + sz * sz), 1e-12f);
      const float t3[3] = {(gvx - r.vx * dot) * inv, (gvy - r.vy * dot) * inv, (gvz - r.vz * dot) * inv};
#pragma unroll
      for (int i = 0; i < 3; ++i) { if (p.static_viewdirs) gc[i] += t3[i]; else gd[i] += t3[i]; }
    }
    // d = R c  ->  g_R[i][j] = gd[i] c[j],  g_c += R^T gd ;  o = pose[:,3]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { out[i * 4 + j] = gd[i] * c[j]; gc[j] = fmaf(pose[i * 4 + j], gd[i], gc[j]); }
      out[i * 4 + 3] = go[i];
    }
    out[12] = -(gc[0] * c[0] + gc[1] * c[1]) / f;   // c.x, c.y are proportional to 1/f
  }
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    float s = out[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 13) {
    const float s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    if (threadIdx.x < 12) { if (g_pose) atomicAdd(g_pose + (size_t)b * 12 + threadIdx.x, s); }
    else if (g_focal) atomicAdd(g_focal + b, s);
  }
}

}  // namespace c3d
