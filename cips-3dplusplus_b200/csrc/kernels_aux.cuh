// Auxiliary kernels: weight packing, FiLM style preparation, ray generation, standalone compositing.
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"

namespace c3d {

// ------------------------------------------------------------------------------------------
// pack_weights: reference state_dict -> packed blob (layout: c3d_common.cuh)
// ------------------------------------------------------------------------------------------
__global__ void pack_small_kernel(c3d_raw_params raw, uint8_t* __restrict__ blob, PackedLayout L) {
  const int c = threadIdx.x;  // 256 threads, 1 block
  const int D = raw.D;
  if (c == 0) {
    reinterpret_cast<uint32_t*>(blob)[0] = C3D_MAGIC;
    reinterpret_cast<int32_t*>(blob)[1] = D;
    float* s = reinterpret_cast<float*>(blob + L.scal);
    s[0] = raw.sigma_bias[0];
    s[1] = raw.rgb_bias[0]; s[2] = raw.rgb_bias[1]; s[3] = raw.rgb_bias[2];
    s[4] = raw.sigmoid_beta[0];
    s[5] = s[6] = s[7] = 0.f;
  }
  reinterpret_cast<float4*>(blob + L.w0)[c] =
      make_float4(raw.pts_weight[0][c * 3 + 0], raw.pts_weight[0][c * 3 + 1], raw.pts_weight[0][c * 3 + 2], raw.pts_bias[0][c]);
  reinterpret_cast<float4*>(blob + L.wvdir)[c] =
      make_float4(raw.views_weight[c * (W + 3) + W + 0], raw.views_weight[c * (W + 3) + W + 1],
                  raw.views_weight[c * (W + 3) + W + 2], 0.f);
  float* bias = reinterpret_cast<float*>(blob + L.bias);
  for (int l = 0; l < D; ++l) bias[l * W + c] = raw.pts_bias[l][c];
  bias[D * W + c] = raw.views_bias[c];
  reinterpret_cast<float*>(blob + L.wsig)[c] = raw.sigma_weight[c];
  reinterpret_cast<float4*>(blob + L.wrgb)[c] =
      make_float4(raw.rgb_weight[c], raw.rgb_weight[W + c], raw.rgb_weight[2 * W + c], 0.f);
  // FiLM biases
  for (int l = 0; l <= D; ++l) {
    float* f = reinterpret_cast<float*>(blob + L.film) + (size_t)l * FILM_LAYER_FLOATS;
    const float* gb = l < D ? raw.pts_gamma_bias[l] : raw.views_gamma_bias;
    const float* bb = l < D ? raw.pts_beta_bias[l] : raw.views_beta_bias;
    f[2 * W * W + c] = gb[c];
    f[2 * W * W + W + c] = bb[c];
  }
  // heads16 fp16 image: [16 n][64 k] x 4 chunks; rows 0..2 rgb head, rows 4/5 = hi/lo split of the sdf head
  uint8_t* r16 = blob + L.rgb16;
  const float ws = raw.sigma_weight[c];
  const float ws_hi = __half2float(__float2half_rn(ws));
  for (int n = 0; n < 16; ++n) {
    float v = n < 3 ? raw.rgb_weight[n * W + c] : (n == 4 ? ws_hi : (n == 5 ? ws - ws_hi : 0.f));
    const int ch = c >> 6, k = c & 63;
    *reinterpret_cast<__half*>(r16 + (size_t)ch * (16 * 128) + sw128_offset(n, k)) = __float2half_rn(v);
  }
  // wk16 image (see c3d_common.cuh): W0 split in slots 0..11, view-direction weights in 12..14
  uint8_t* w0i = blob + L.w0img;
  for (int j = 0; j < 3; ++j) {
    const float w = raw.pts_weight[0][c * 3 + j];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(w0i + k16_offset(c, 4 * j + 0)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(w0i + k16_offset(c, 4 * j + 1)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(w0i + k16_offset(c, 4 * j + 2)) = lo;
    *reinterpret_cast<__nv_bfloat16*>(w0i + k16_offset(c, 4 * j + 3)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(w0i + k16_offset(c, 12 + j)) = __float2bfloat16_rn(raw.views_weight[c * (W + 3) + W + j]);
  }
  *reinterpret_cast<__nv_bfloat16*>(w0i + k16_offset(c, 15)) = __float2bfloat16_rn(0.f);
  // backward side images (layout comment in c3d_common.cuh)
  uint8_t* b0 = blob + L.bwd16;
  uint8_t* b1 = b0 + W0IMG_BYTES;
  uint8_t* b2 = b1 + W0IMG_BYTES;
  for (int k = 0; k < 16; ++k) {
    const float v0 = k < 6 ? raw.rgb_weight[(k % 3) * W + c] : 0.f;
    const float v1 = (k == 6 || k == 7) ? raw.sigma_weight[c] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(b0 + k16_offset(c, k)) = __float2bfloat16_rn(v0);
    *reinterpret_cast<__nv_bfloat16*>(b1 + k16_offset(c, k)) = __float2bfloat16_rn(v1);
  }
  for (int n = 0; n < 16; ++n) {
    float v = 0.f;
    if (n < 3) v = raw.pts_weight[0][c * 3 + n];
    else if (n >= 4 && n < 7) v = raw.views_weight[c * (W + 3) + W + (n - 4)];
    *reinterpret_cast<__nv_bfloat16*>(b2 + (size_t)(c >> 6) * (16 * 128) + sw128_offset(n, c & 63)) = __float2bfloat16_rn(v);
  }
}

// grid (D+1 layers, 256/32 row tiles), block (32,8): transposes 32x32 tiles of the three 256x256
// matrices of layer l (FiLM gamma, FiLM beta, hidden weight) and writes the bf16 swizzled image.
__global__ void pack_matrix_kernel(c3d_raw_params raw, uint8_t* __restrict__ blob, PackedLayout L) {
  __shared__ float tile[32][33];
  const int l = blockIdx.x;        // 0..D
  const int D = raw.D;
  const int which = blockIdx.z;    // 0 gamma, 1 beta, 2 hidden weight
  const float* src; int ld = W;
  if (which == 0) src = l < D ? raw.pts_gamma_weight[l] : raw.views_gamma_weight;
  else if (which == 1) src = l < D ? raw.pts_beta_weight[l] : raw.views_beta_weight;
  else {
    if (l == 0) return;            // layer 0 is 256x3, lives in w0
    src = l < D ? raw.pts_weight[l] : raw.views_weight;
    ld = l < D ? W : W + 3;
  }
  const int r0 = blockIdx.y * 32;  // output-channel tile
  for (int c0 = 0; c0 < W; c0 += 32) {
    // load src[r0+ty*4+i][c0+tx]
    for (int i = threadIdx.y; i < 32; i += 8) tile[i][threadIdx.x] = src[(size_t)(r0 + i) * ld + c0 + threadIdx.x];
    __syncthreads();
    if (which < 2) {
      float* dst = reinterpret_cast<float*>(blob + L.film) + (size_t)l * FILM_LAYER_FLOATS + (size_t)which * W * W;
      for (int i = threadIdx.y; i < 32; i += 8) dst[(size_t)(c0 + i) * W + r0 + threadIdx.x] = tile[threadIdx.x][i];
    } else {
      float* dst = reinterpret_cast<float*>(blob + L.wT32) + (size_t)(l - 1) * W * W;
      for (int i = threadIdx.y; i < 32; i += 8) dst[(size_t)(c0 + i) * W + r0 + threadIdx.x] = tile[threadIdx.x][i];
      float* dsn = reinterpret_cast<float*>(blob + L.w32) + (size_t)(l - 1) * W * W;
      for (int i = threadIdx.y; i < 32; i += 8) dsn[(size_t)(r0 + i) * W + c0 + threadIdx.x] = tile[i][threadIdx.x];
      uint8_t* imgT = blob + L.wbf16T + (size_t)(l - 1) * WBF16_LAYER_BYTES;
      uint8_t* imgH = blob + L.wf16h + (size_t)(l - 1) * WBF16_LAYER_BYTES;
      uint8_t* imgL = blob + L.wf16l + (size_t)(l - 1) * WBF16_LAYER_BYTES;
      for (int i = threadIdx.y; i < 32; i += 8) {
        const int n = r0 + i, k = c0 + threadIdx.x;
        const size_t off = (size_t)(k >> 6) * WBF16_CHUNK_BYTES + sw128_offset(n, k & 63);
        const float wv = tile[i][threadIdx.x];
        const __half hi = __float2half_rn(wv);                  // fp16(W): forward operand; with the 2^11-scaled lo the fp32 mode's
        *reinterpret_cast<__half*>(imgH + off) = hi;
        *reinterpret_cast<__half*>(imgL + off) = __float2half_rn((wv - __half2float(hi)) * 2048.0f);
        // transposed image: row = input channel k, column = output channel n
        *reinterpret_cast<__nv_bfloat16*>(imgT + (size_t)(n >> 6) * WBF16_CHUNK_BYTES + sw128_offset(k, n & 63)) =
            __float2bfloat16_rn(tile[i][threadIdx.x]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// style_prep: styles (b, D+1, 256) -> FiLM tables   (volume_renderer.py:66-67, 77-83)
//   film  (b, D+1, 256) float2 : (gamma, gamma*bias + beta)   epilogue form sin(gamma*acc + shift)
//   first (b, 256) float4      : gamma0 * W0[c][0..2], shift0
//   view  (b, 256) float4      : gammaD * Wview[c][256..258], 0
// grid (D+1, ceil(b/8)), block 256 (thread = output channel), 8 images per block share the weights.
// ------------------------------------------------------------------------------------------
constexpr int SP_IMGS = 8;
constexpr int SP_KSPLIT = 2;                     // the 256-long dot products are split over 2 thread groups: the kernel is
constexpr int SP_THREADS = W * SP_KSPLIT;        // latency-bound (a 67 us fixed cost of every small-batch call otherwise)
__global__ void __launch_bounds__(SP_THREADS) style_prep_kernel(const uint8_t* __restrict__ blob, PackedLayout L,
                                                                 const float* __restrict__ styles, int batch,
                                                                 float2* __restrict__ film, float4* __restrict__ first,
                                                                 float4* __restrict__ view) {
  __shared__ float s[SP_IMGS][W];
  __shared__ float part[SP_KSPLIT - 1][SP_IMGS][2][W];
  const int l = blockIdx.x, D = L.D, c = threadIdx.x & (W - 1), ks = threadIdx.x / W;
  const int b0 = blockIdx.y * SP_IMGS;
  for (int i = ks; i < SP_IMGS; i += SP_KSPLIT) {
    const int b = b0 + i;
    s[i][c] = b < batch ? styles[((size_t)b * (D + 1) + l) * W + c] : 0.f;
  }
  __syncthreads();
  const float* f = reinterpret_cast<const float*>(blob + L.film) + (size_t)l * FILM_LAYER_FLOATS;
  const float* GwT = f;
  const float* BwT = f + (size_t)W * W;
  float g[SP_IMGS], be[SP_IMGS];
#pragma unroll
  for (int i = 0; i < SP_IMGS; ++i) g[i] = be[i] = 0.f;
  constexpr int KS = W / SP_KSPLIT;
#pragma unroll 16
  for (int kk = 0; kk < KS; ++kk) {
    const int k = ks * KS + kk;
    const float gw = GwT[(size_t)k * W + c], bw = BwT[(size_t)k * W + c];
#pragma unroll
    for (int i = 0; i < SP_IMGS; ++i) {
      g[i] = fmaf(s[i][k], gw, g[i]);
      be[i] = fmaf(s[i][k], bw, be[i]);
    }
  }
  if (ks > 0) {
#pragma unroll
    for (int i = 0; i < SP_IMGS; ++i) { part[ks - 1][i][0][c] = g[i]; part[ks - 1][i][1][c] = be[i]; }
  }
  __syncthreads();
  if (ks > 0) return;
#pragma unroll
  for (int q = 0; q < SP_KSPLIT - 1; ++q)
#pragma unroll
    for (int i = 0; i < SP_IMGS; ++i) { g[i] += part[q][i][0][c]; be[i] += part[q][i][1][c]; }
  const float gb = f[2 * W * W + c], bb = f[2 * W * W + W + c];
  const float bias = reinterpret_cast<const float*>(blob + L.bias)[l * W + c];
  const float4 w0 = reinterpret_cast<const float4*>(blob + L.w0)[c];
  const float4 wv = reinterpret_cast<const float4*>(blob + L.wvdir)[c];
#pragma unroll
  for (int i = 0; i < SP_IMGS; ++i) {
    const int b = b0 + i;
    if (b >= batch) break;
    const float gamma = 15.0f * (g[i] + gb) + 30.0f;   // LinearLayer(std_init=15, bias_init=30)
    const float beta = 0.25f * (be[i] + bb);           // LinearLayer(std_init=0.25)
    const float shift = fmaf(gamma, bias, beta);
    film[((size_t)b * (D + 1) + l) * W + c] = make_float2(gamma, shift);
    if (l == 0) first[(size_t)b * W + c] = make_float4(gamma * w0.x, gamma * w0.y, gamma * w0.z, shift);
    if (l == D) view[(size_t)b * W + c] = make_float4(gamma * wv.x, gamma * wv.y, gamma * wv.z, 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// Per-image weight images of the CTA-pair forward kernel (fused_pair_sm100.cuh): FiLM folded into the GEMM operands.
//   wimg[b][idx 0..D-1][kc 0..3][half 0..1] : 16 KB stage images [128 rows n][64 k] fp16, K-major SWIZZLE_128B,
//                                             value fp16(gamma_{b,idx+1}[n] * W_{idx+1}[n][k]),  n = 128 half + row
//   kimg[b][L 0..D][half 0..1]              : 4 KB K16 images [128 rows n][16 slots], UMMA K-major no-swizzle:
//        L = 0     slots 4j..4j+3 = (hi, hi, lo, hi) of gamma W0[n][j]   (x point tile (hi, mid, hi, lo))
//        L = D     slots j and 3+j = bf16(gamma Wview[n][256+j])         (x view tile (hi x3, lo x3))
//        all L     slots 12, 13 = hi / lo of the shift gamma b + beta    (x the two "ones" slots of either tile)
// ------------------------------------------------------------------------------------------
// grid (8 = kc*2 + half, D, batch), block 256: thread = (row, 32-element half of the 64-wide K-chunk)
__global__ void __launch_bounds__(256) film_weights_kernel(const uint8_t* __restrict__ blob, PackedLayout L,
                                                            const float2* __restrict__ film, uint8_t* __restrict__ wimg) {
  const int D = L.D, idx = blockIdx.y, b = blockIdx.z, kc = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int i = threadIdx.x >> 1, q = threadIdx.x & 1, n = half * 128 + i;
  const float gamma = film[((size_t)b * (D + 1) + idx + 1) * W + n].x;
  const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(blob + L.w32) + (size_t)idx * W * W +
                                                      (size_t)n * W + kc * 64 + q * 32);
  uint8_t* dst = wimg + (((size_t)b * D + idx) * 8 + blockIdx.x) * 16384 + (size_t)i * 128;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float4 x = src[2 * u], y = src[2 * u + 1];
    uint4 o;
    o.x = ptx::pack_f16x2(gamma * x.x, gamma * x.y); o.y = ptx::pack_f16x2(gamma * x.z, gamma * x.w);
    o.z = ptx::pack_f16x2(gamma * y.x, gamma * y.y); o.w = ptx::pack_f16x2(gamma * y.z, gamma * y.w);
    *reinterpret_cast<uint4*>(dst + (((q * 4 + u) ^ (i & 7)) << 4)) = o;
  }
}

// grid (D+1, batch), block 256 (thread = output channel n)
__global__ void __launch_bounds__(256) film_k16_kernel(const uint8_t* __restrict__ blob, PackedLayout L,
                                                        const float2* __restrict__ film, uint8_t* __restrict__ kimg) {
  const int D = L.D, l = blockIdx.x, b = blockIdx.y, n = threadIdx.x, half = n >> 7, i = n & 127;
  const float2 f = film[((size_t)b * (D + 1) + l) * W + n];
  float s[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) s[k] = 0.f;
  auto bf = [](float x) { return __bfloat162float(__float2bfloat16_rn(x)); };
  if (l == 0) {
    const float4 w0 = reinterpret_cast<const float4*>(blob + L.w0)[n];
    const float w[3] = {f.x * w0.x, f.x * w0.y, f.x * w0.z};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float hi = bf(w[j]), lo = bf(w[j] - hi);
      s[4 * j + 0] = hi; s[4 * j + 1] = hi; s[4 * j + 2] = lo; s[4 * j + 3] = hi;
    }
  }
  if (l == D) {
    const float4 wv = reinterpret_cast<const float4*>(blob + L.wvdir)[n];
    s[0] = s[3] = f.x * wv.x; s[1] = s[4] = f.x * wv.y; s[2] = s[5] = f.x * wv.z;
  }
  const float sh = bf(f.y);
  s[12] = sh; s[13] = f.y - sh;
  uint8_t* dst = kimg + (((size_t)b * (D + 1) + l) * 2 + half) * 4096 + k16_offset(i, 0);
  *reinterpret_cast<uint4*>(dst) = make_uint4(ptx::pack_bf16x2(s[0], s[1]), ptx::pack_bf16x2(s[2], s[3]),
                                              ptx::pack_bf16x2(s[4], s[5]), ptx::pack_bf16x2(s[6], s[7]));
  *reinterpret_cast<uint4*>(dst + 128) = make_uint4(ptx::pack_bf16x2(s[8], s[9]), ptx::pack_bf16x2(s[10], s[11]),
                                                    ptx::pack_bf16x2(s[12], s[13]), ptx::pack_bf16x2(s[14], s[15]));
}

// ------------------------------------------------------------------------------------------
// raygen: Render.prepare_nerf_inputs (nerf_utils.py:172-218); thread per sample point, so the (b,hw,N,3) / (b,hw,N)
// writes of a warp are contiguous (a thread per ray would write 12 N-byte runs 12 N bytes apart).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raygen_kernel(c3d_raygen_params p) {
  const int hw = p.img_size * p.img_size, N = p.n_samples;
  const long long total = (long long)p.batch * hw * N;
  const long long pid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pid >= total) return;
  const long long gid = pid / N;                       // ray index over the batch
  const int k = (int)(pid - gid * N);
  const int b = (int)(gid / hw), ray = (int)(gid - (long long)b * hw);
  const RayGeom r = make_ray(p.cam_poses + (size_t)b * 12, p.focal[b], p.img_size, ray, p.static_viewdirs != 0);
  const float u = p.ray_offset ? p.ray_offset[gid] : 0.f;
  if (k == 0) {
    if (p.rays_d) { float* o = p.rays_d + gid * 3; o[0] = r.dx; o[1] = r.dy; o[2] = r.dz; }
    if (p.viewdirs) { float* o = p.viewdirs + gid * 3; o[0] = r.vx; o[1] = r.vy; o[2] = r.vz; }
  }
  const float z = sample_depth(p.near[b], p.far[b], k, N, u);
  if (p.z_vals) p.z_vals[pid] = z;
  if (p.pts) {
    float* o = p.pts + pid * 3;
    o[0] = fmaf(r.dx, z, r.ox); o[1] = fmaf(r.dy, z, r.oy); o[2] = fmaf(r.dz, z, r.oz);
  }
}

// ------------------------------------------------------------------------------------------
// composite_forward: Render.volume_integration (nerf_utils.py:230-338), one warp per ray.
//   lanes = samples for the density -> alpha -> transmittance scan (shuffle prefix product),
//   lanes = channels (float4 each) for the weighted feature sum.  HBM-bound: every input byte
//   is read exactly once (1056*N + 1152 B/ray at 256 channels).
// ------------------------------------------------------------------------------------------
constexpr int CMP_MAX_N = 256;
__device__ __forceinline__ float warp_excl_prod(float v, int lane, float& total) {
  // inclusive product scan, then shift
  float x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x *= y;
  }
  total = __shfl_sync(0xffffffffu, x, 31);
  float e = __shfl_up_sync(0xffffffffu, x, 1);
  return lane == 0 ? 1.0f : e;
}

template <bool kPrecise>
__device__ __forceinline__ float alpha_from_sdf(float sdf, float inv_beta, float dist) {
  // sigma = sigmoid(-sdf/beta)/beta ; alpha = 1 - exp(-sigma*dist)      (nerf_utils.py:278,286)
  const float sg = kPrecise ? sigmoid_precise(-sdf * inv_beta) : sigmoidf_(-sdf * inv_beta);
  const float sigma = sg * inv_beta;
  return 1.0f - (kPrecise ? expf(-sigma * dist) : __expf(-sigma * dist));
}

__global__ void __launch_bounds__(256) composite_fwd_kernel(c3d_composite_params p) {
  __shared__ float s_w[8][CMP_MAX_N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * 8 + warp;
  if (ray >= p.n_rays) return;
  const int N = p.n_samples;
  const bool raw_density = (p.flags & C3D_COMPOSITE_RAW_DENSITY) != 0;     // with_sdf=False branch (nerf_utils.py:288-296)
  const float beta = p.sigmoid_beta_ptr ? *p.sigmoid_beta_ptr : p.sigmoid_beta;
  const float inv_beta = 1.0f / beta;
  const float* rd = p.rays_d + ray * 3;
  const float dnorm = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
  const float* z = p.z_vals + ray * N;
  const float* sdf = p.sdf + ray * N;
  float carry = 1.0f;
  // ---- pass 1: weights w_k = alpha_k T_k into shared memory
  float wsum_head = 0.f;                                                    // sum of w_0 .. w_{N-2} (force_background)
  for (int k0 = 0; k0 < N; k0 += 32) {
    const int k = k0 + lane;
    float one_minus = 1.0f, alpha = 0.f;
    if (k < N) {
      const float dist = (k + 1 < N ? z[k + 1] - z[k] : 1e10f) * dnorm;
      if (raw_density) {
        const float x = sdf[k];                                             // raw sigma (+ noise, added by the caller)
        const float sp = x > 20.0f ? x : log1pf(expf(x));                   // F.softplus (threshold 20)
        alpha = 1.0f - expf(-sp * dist);
      } else {
        alpha = alpha_from_sdf<true>(sdf[k], inv_beta, dist);
      }
      one_minus = 1.0f - alpha + 1e-10f;
    }
    float total;
    const float T = carry * warp_excl_prod(one_minus, lane, total);
    carry *= total;
    if (k < N) {
      const float w = alpha * T;
      s_w[warp][k] = w;
      if (k < N - 1) wsum_head += w;
    }
  }
  if (p.flags & C3D_COMPOSITE_FORCE_BACKGROUND) {                           // weights[..., -1] = 1 - sum(weights[..., :-1])
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum_head += __shfl_xor_sync(0xffffffffu, wsum_head, o);
    if (lane == 0) s_w[warp][N - 1] = 1.0f - wsum_head;
  }
  __syncwarp();
  // ---- pass 2: weighted sums of sigmoid(rgb) and of the points
  float a_rgb0 = 0.f, a_rgb1 = 0.f, a_rgb2 = 0.f, a_x = 0.f, a_y = 0.f, a_z = 0.f;
  for (int k = lane; k < N; k += 32) {
    const float w = s_w[warp][k];
    if (p.weights) p.weights[ray * N + k] = w;
    const float* c = p.rgb + (ray * N + k) * 3;
    a_rgb0 = fmaf(w, sigmoid_precise(c[0]), a_rgb0);
    a_rgb1 = fmaf(w, sigmoid_precise(c[1]), a_rgb1);
    a_rgb2 = fmaf(w, sigmoid_precise(c[2]), a_rgb2);
    const float* q = p.pts + (ray * N + k) * 3;
    a_x = fmaf(w, q[0], a_x); a_y = fmaf(w, q[1], a_y); a_z = fmaf(w, q[2], a_z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a_rgb0 += __shfl_xor_sync(0xffffffffu, a_rgb0, o);
    a_rgb1 += __shfl_xor_sync(0xffffffffu, a_rgb1, o);
    a_rgb2 += __shfl_xor_sync(0xffffffffu, a_rgb2, o);
    a_x += __shfl_xor_sync(0xffffffffu, a_x, o);
    a_y += __shfl_xor_sync(0xffffffffu, a_y, o);
    a_z += __shfl_xor_sync(0xffffffffu, a_z, o);
  }
  if (lane == 0) {
    float* o = p.rgb_map + ray * 3;
    o[0] = -1.0f + 2.0f * a_rgb0; o[1] = -1.0f + 2.0f * a_rgb1; o[2] = -1.0f + 2.0f * a_rgb2;
    float* x = p.xyz + ray * 3;
    x[0] = a_x; x[1] = a_y; x[2] = a_z;
    p.mask[ray * 2 + 0] = s_w[warp][N - 1];
    p.mask[ray * 2 + 1] = -sqrtf(a_x * a_x + a_y * a_y + a_z * a_z);
  }
  __syncwarp();
  if (p.features && p.n_feat > 0) {
    const int C4 = p.n_feat >> 2;                    // float4 per row
    const float4* f = reinterpret_cast<const float4*>(p.features) + (size_t)ray * N * C4;
    for (int c = lane; c < C4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = 0;
      for (; k + 4 <= N; k += 4) {                    // 4 independent 16-byte loads in flight per lane
        const float4 v0 = __ldcs(f + (size_t)(k + 0) * C4 + c);
        const float4 v1 = __ldcs(f + (size_t)(k + 1) * C4 + c);
        const float4 v2 = __ldcs(f + (size_t)(k + 2) * C4 + c);
        const float4 v3 = __ldcs(f + (size_t)(k + 3) * C4 + c);
        const float w0 = s_w[warp][k], w1 = s_w[warp][k + 1], w2 = s_w[warp][k + 2], w3 = s_w[warp][k + 3];
        acc.x = fmaf(w0, v0.x, acc.x); acc.y = fmaf(w0, v0.y, acc.y); acc.z = fmaf(w0, v0.z, acc.z); acc.w = fmaf(w0, v0.w, acc.w);
        acc.x = fmaf(w1, v1.x, acc.x); acc.y = fmaf(w1, v1.y, acc.y); acc.z = fmaf(w1, v1.z, acc.z); acc.w = fmaf(w1, v1.w, acc.w);
        acc.x = fmaf(w2, v2.x, acc.x); acc.y = fmaf(w2, v2.y, acc.y); acc.z = fmaf(w2, v2.z, acc.z); acc.w = fmaf(w2, v2.w, acc.w);
        acc.x = fmaf(w3, v3.x, acc.x); acc.y = fmaf(w3, v3.y, acc.y); acc.z = fmaf(w3, v3.z, acc.z); acc.w = fmaf(w3, v3.w, acc.w);
      }
      for (; k < N; ++k) {
        const float4 v = __ldcs(f + (size_t)k * C4 + c);
        const float w = s_w[warp][k];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      reinterpret_cast<float4*>(p.feature_map)[(size_t)ray * C4 + c] = acc;
    }
  }
}

}  // namespace c3d
