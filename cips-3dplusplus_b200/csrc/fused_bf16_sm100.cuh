// Fused NeRF-branch forward for sm_100a: rays -> samples -> FiLM-SIREN point MLP on tcgen05 tensor
// cores (bf16 operands, fp32 TMEM accumulators) -> SDF compositing -> (rgb, feature, sdf, mask, xyz)
// maps.  Per-point activations never leave the SM.
//
// Reference semantics: exp/cips3d/volume_renderer.py:133-160,192-283 and
// exp/cips3d/nerf_utils.py:17-218,230-338 (see DESIGN.md for the math and the data layout).
//
// One persistent CTA per SM, 12 warps:
//   warp 0  (1 lane)  weight producer : cp.async.bulk (TMA engine) of pre-swizzled 32 KB K-chunks of
//                                        the layer's bf16 weights into a 2-stage shared-memory ring
//                                        (multicast to the 2 CTAs of a cluster when kCluster == 2)
//   warp 1  (1 lane)  MMA issuer      : tcgen05.mma 128x256x16 (hidden layers), 2 x 128x128x16 with the
//                                        operand roles swapped (view layer -> channels on TMEM lanes),
//                                        128x16x16 (rgb head); tcgen05.commit -> mbarriers
//   warp 2            TMEM allocator (512 columns = two 128x256 fp32 accumulators)
//   warps 4-7 / 8-11  epilogue group 0 / 1: each owns one 128-point tile "slot" (activations 64 KB in
//                                        shared memory as the K-major SWIZZLE_128B A operand, one
//                                        accumulator).  While the tensor core runs slot X's layer, slot
//                                        Y's warps apply sin(gamma*acc+shift) and rewrite its A tile.
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"

namespace c3d { namespace fused {

using namespace c3d::ptx;

constexpr int NTHREADS = 384;
constexpr int TILE = 128;
constexpr int ACT_CHUNK = TILE * 128;          // 16384 B: [128 rows][64 bf16]
constexpr int ACT_BYTES = NCHUNK * ACT_CHUNK;  // 65536
constexpr int STAGE_BYTES = W * 128;           // 32768: [256 rows][64 bf16]
constexpr int NSTAGE = 2;
constexpr int RSLOTS = 32;                     // per-slot ray accumulators (rays per tile <= 128/N + 2)
constexpr int MIN_SAMPLES = 8;

// shared memory map (bytes from a 1024-aligned base)
constexpr int SM_ACT = 0;
constexpr int SM_STAGE = SM_ACT + 2 * ACT_BYTES;                 // 131072
constexpr int SM_RGB16 = SM_STAGE + NSTAGE * STAGE_BYTES;        // 196608
constexpr int SM_FILM = SM_RGB16 + (int)RGB16_BYTES;             // 204800  [slot][buf][256] float2
constexpr int SM_TAB0 = SM_FILM + 2 * 2 * W * 8;                 // 212992  [slot][256] float4
constexpr int SM_PT = SM_TAB0 + 2 * W * 16;                      // 221184  [slot][128] float4 (w, vx, vy, vz)
constexpr int SM_WSIG = SM_PT + 2 * TILE * 16;                   // 225280  [256] float
constexpr int SM_OM = SM_WSIG + W * 4;                           // 226304  [slot][128] float
constexpr int SM_FLAG = SM_OM + 2 * TILE * 4;                    // 227328  [slot][128] int
constexpr int SM_RAYACC = SM_FLAG + 2 * TILE * 4;                // 228352  [slot][RSLOTS][8] float
constexpr int SM_MISC = SM_RAYACC + 2 * RSLOTS * 8 * 4;          // 230400  barriers, tmem ptr, carries
constexpr int SM_TOTAL = SM_MISC + 256;                          // 230656
constexpr int SMEM_BYTES = SM_TOTAL + 1024;                      // + alignment slack

struct Args {
  const uint8_t* blob; PackedLayout L;
  const float2* film; const float4* first; const float4* view;   // style_prep tables, indexed by image
  int batch, n_rays, n_samples, D, img_size, static_viewdirs, input_kind;
  int unit_rays, units_per_img;
  const float* cam_poses; const float* focal; const float* near; const float* far; const float* ray_offset;
  const float* pts; const float* rays_d; const float* viewdirs; const float* z_vals;
  float* rgb_map; float* feature_map; float* sdf; float* mask; float* xyz; float* z_vals_out;
  int debug;   // bit 0: producer skips the weight copies (timing experiments only; results are garbage)
};

struct Misc {
  uint64_t full[NSTAGE], empty[NSTAGE], a_ready[2], acc_full[2];
  uint32_t tmem_base;
  float carry[2];
};

__device__ __forceinline__ int unit_tiles(const Args& a, int u) {
  const int r0 = (u % a.units_per_img) * a.unit_rays;
  const int nr = min(a.unit_rays, a.n_rays - r0);
  return (nr * a.n_samples + TILE - 1) / TILE;
}
__device__ __forceinline__ int slot_tiles(const Args& a, int slot, int nslots) {
  int t = 0;
  const int total = a.batch * a.units_per_img;
  for (int u = slot; u < total; u += nslots) t += unit_tiles(a, u);
  return t;
}

// 8 activations -> bf16 -> one 16-byte store into the K-major SWIZZLE_128B A tile (row = point)
__device__ __forceinline__ void store8(uint8_t* act_row /*act + row*128*/, int row7, int c8 /*0..31*/, const float (&o)[8]) {
  uint4 pk;
  pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
  pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
  *reinterpret_cast<uint4*>(act_row + (c8 >> 3) * ACT_CHUNK + (((c8 & 7) ^ row7) << 4)) = pk;
}

template <int kCluster>
__global__ void __launch_bounds__(NTHREADS, 1) fused_forward_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D, N = a.n_samples;
  const int JOBS = D + 1;                           // MMA jobs per tile: D-1 hidden, view, rgb
  const int nslots = 2 * gridDim.x;
  const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0u;

  // ---------------------------------------------------------------- one-time setup
  if (threadIdx.x == 32) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&misc->full[i], 1); mbar_init(&misc->empty[i], kCluster); }
    for (int i = 0; i < 2; ++i) { mbar_init(&misc->a_ready[i], TILE); mbar_init(&misc->acc_full[i], 1); }
    misc->carry[0] = misc->carry[1] = 1.0f;
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(&misc->tmem_base, 512); tmem_relinquish(); }
  {
    // resident tables: rgb-head bf16 image, sigma weights; zero the ray accumulators
    const uint4* src = reinterpret_cast<const uint4*>(a.blob + a.L.rgb16);
    uint4* dst = reinterpret_cast<uint4*>(smem + SM_RGB16);
    for (int i = threadIdx.x; i < (int)RGB16_BYTES / 16; i += NTHREADS) dst[i] = src[i];
    const float* ws = reinterpret_cast<const float*>(a.blob + a.L.wsig);
    float* wd = reinterpret_cast<float*>(smem + SM_WSIG);
    for (int i = threadIdx.x; i < W; i += NTHREADS) wd[i] = ws[i];
    float* ra = reinterpret_cast<float*>(smem + SM_RAYACC);
    for (int i = threadIdx.x; i < 2 * RSLOTS * 8; i += NTHREADS) ra[i] = 0.f;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = misc->tmem_base;

  // rounds of the merged job sequence; identical for every CTA of a cluster (weight ring is shared)
  int my_tiles[2];
  my_tiles[0] = slot_tiles(a, 2 * blockIdx.x + 0, nslots);
  my_tiles[1] = slot_tiles(a, 2 * blockIdx.x + 1, nslots);
  int max_tiles = max(my_tiles[0], my_tiles[1]);
  if (kCluster > 1) {
    const int peer = blockIdx.x ^ 1;
    max_tiles = max(max_tiles, max(slot_tiles(a, 2 * peer, nslots), slot_tiles(a, 2 * peer + 1, nslots)));
  }
  const int rounds = max_tiles * JOBS;

  if (warp == 0 && lane == 0) {
    // ============================================================ weight producer
    const uint8_t* wsrc = a.blob + a.L.wbf16;
    uint32_t n = 0;
    for (int g = 0; g < rounds; ++g) {
      const int j = g % JOBS;
      if (j >= D) continue;
      for (int s = 0; s < 2; ++s) {
        for (int c = 0; c < NCHUNK; ++c, ++n) {
          const uint32_t st = n & 1u, ph = (n >> 1) & 1u;
          mbar_wait(&misc->empty[st], ph ^ 1u);
          uint8_t* dst = smem + SM_STAGE + st * STAGE_BYTES;
          const uint8_t* src = wsrc + (size_t)j * WBF16_LAYER_BYTES + (size_t)c * WBF16_CHUNK_BYTES;
          mbar_arrive_expect_tx(&misc->full[st], STAGE_BYTES);
          if (kCluster == 1) {
            bulk_g2s(dst, src, STAGE_BYTES, &misc->full[st]);
          } else {
            const uint32_t half = STAGE_BYTES / 2;
            bulk_g2s_multicast(dst + cta_rank * half, src + cta_rank * half, half, &misc->full[st], (uint16_t)0x3);
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ============================================================ MMA issuer
    const uint32_t idesc_l = umma_idesc_bf16(128, 256), idesc_t = umma_idesc_bf16(128, 128), idesc_r = umma_idesc_bf16(128, 16);
    const uint32_t act_addr[2] = {smem_u32(smem + SM_ACT), smem_u32(smem + SM_ACT + ACT_BYTES)};
    const uint32_t stage_addr[2] = {smem_u32(smem + SM_STAGE), smem_u32(smem + SM_STAGE + STAGE_BYTES)};
    const uint32_t rgb_addr = smem_u32(smem + SM_RGB16);
    uint32_t n = 0, jobcnt[2] = {0u, 0u};
    for (int g = 0; g < rounds; ++g) {
      const int j = g % JOBS, tile_idx = g / JOBS;
      for (int s = 0; s < 2; ++s) {
        const bool real = tile_idx < my_tiles[s];
        const uint32_t tacc = tmem_base + (uint32_t)s * 256u;
        if (j < D) {
          if (real) { mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u); tc_fence_after(); }
          for (int c = 0; c < NCHUNK; ++c, ++n) {
            const uint32_t st = n & 1u, ph = (n >> 1) & 1u;
            mbar_wait(&misc->full[st], ph);
            tc_fence_after();
            if (real) {
              if (j < D - 1) {            // hidden layer: rows = points (A = activations), N = 256 channels
                const uint64_t ad = umma_desc_kmajor_sw128(act_addr[s] + c * ACT_CHUNK);
                const uint64_t bd = umma_desc_kmajor_sw128(stage_addr[st]);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tacc, ad + 2 * kk, bd + 2 * kk, idesc_l, (c | kk) != 0);
              } else {                    // view layer, operand roles swapped: rows = channels, N = 128 points
                const uint64_t bd = umma_desc_kmajor_sw128(act_addr[s] + c * ACT_CHUNK);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint64_t ad = umma_desc_kmajor_sw128(stage_addr[st] + h * (STAGE_BYTES / 2));
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tacc + (uint32_t)h * 128u, ad + 2 * kk, bd + 2 * kk, idesc_t, (c | kk) != 0);
                }
              }
            }
            if (kCluster == 1) umma_commit(&misc->empty[st]);
            else umma_commit_multicast(&misc->empty[st], (uint16_t)0x3);
          }
          if (real) { umma_commit(&misc->acc_full[s]); jobcnt[s]++; }
        } else if (real) {                // rgb head: rows = points, N = 16 (3 used)
          mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c) {
            const uint64_t ad = umma_desc_kmajor_sw128(act_addr[s] + c * ACT_CHUNK);
            const uint64_t bd = umma_desc_kmajor_sw128(rgb_addr + c * 2048);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tacc, ad + 2 * kk, bd + 2 * kk, idesc_r, (c | kk) != 0);
          }
          umma_commit(&misc->acc_full[s]);
          jobcnt[s]++;
        }
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue groups (thread = point row / channel)
    const int s = (warp - 4) >> 2;
    const int t = threadIdx.x - 128 - s * TILE;          // 0..127
    const int quad = warp & 3;
    const uint32_t bar_id = 1u + (uint32_t)s;
    const int slot = 2 * blockIdx.x + s;
    uint8_t* act = smem + SM_ACT + s * ACT_BYTES;
    uint8_t* act_row = act + t * 128;
    const int row7 = t & 7;
    float2* filmS = reinterpret_cast<float2*>(smem + SM_FILM) + s * 2 * W;
    float4* tab0S = reinterpret_cast<float4*>(smem + SM_TAB0) + s * W;
    float4* ptS = reinterpret_cast<float4*>(smem + SM_PT) + s * TILE;
    const float* wsigS = reinterpret_cast<const float*>(smem + SM_WSIG);
    float* omS = reinterpret_cast<float*>(smem + SM_OM) + s * TILE;
    int* flagS = reinterpret_cast<int*>(smem + SM_FLAG) + s * TILE;
    float* rayacc = reinterpret_cast<float*>(smem + SM_RAYACC) + s * RSLOTS * 8;
    const uint32_t tacc = tmem_base + (uint32_t)s * 256u + ((uint32_t)(quad * 32) << 16);
    const float* scal = reinterpret_cast<const float*>(a.blob + a.L.scal);
    const float bsig = scal[0], brgb0 = scal[1], brgb1 = scal[2], brgb2 = scal[3];
    const float inv_beta = 1.0f / scal[4];
    // view-layer state (thread = channel t and t+128)
    uint32_t xoroff[8];
#pragma unroll
    for (int jx = 0; jx < 8; ++jx) xoroff[jx] = (uint32_t)((((t & 63) >> 3) ^ jx) << 4) + (uint32_t)((t & 7) << 1);
    float cur[2] = {0.f, 0.f}, shift_ray[2] = {0.f, 0.f};
    uint32_t jobcnt = 0;
    int cur_img = -1;
    const int total_units = a.batch * a.units_per_img;

    for (int u = slot; u < total_units; u += nslots) {
      const int img = u / a.units_per_img;
      const int r0 = (u - img * a.units_per_img) * a.unit_rays;
      const int nr = min(a.unit_rays, a.n_rays - r0);
      const int npts = nr * N;
      const int ntiles = (npts + TILE - 1) / TILE;
      const float near = a.near[img], far = a.far[img];
      const float nscale = 2.0f / (far - near);
      if (img != cur_img) {                        // layer-0 table of this image
        named_bar_sync(bar_id, TILE);              // all readers of the previous table are done
        tab0S[t] = a.first[(size_t)img * W + t];
        tab0S[t + TILE] = a.first[(size_t)img * W + t + TILE];
        cur_img = img;
        named_bar_sync(bar_id, TILE);
      }
      const float2* film_img = a.film + (size_t)img * (D + 1) * W;
      cur[0] = cur[1] = 0.f;

      for (int tile = 0; tile < ntiles; ++tile) {
        // ------------------------------------------------ geometry of my point (nerf_utils.py:17-170)
        const int q = tile * TILE + t;
        const bool valid = q < npts;
        const int qc = valid ? q : npts - 1;
        const int rl = qc / N, k = qc - rl * N;
        const size_t gray = (size_t)img * a.n_rays + r0 + rl;
        float px, py, pz, vx, vy, vz, dist, zk;
        if (a.input_kind == C3D_INPUT_POSES) {
          const RayGeom rg = make_ray(a.cam_poses + (size_t)img * 12, a.focal[img], a.img_size, r0 + rl, a.static_viewdirs != 0);
          const float uo = a.ray_offset ? a.ray_offset[gray] : 0.f;
          zk = sample_depth(near, far, k, N, uo);
          const float z1 = (k + 1 < N) ? sample_depth(near, far, k + 1, N, uo) : 0.f;
          px = fmaf(rg.dx, zk, rg.ox); py = fmaf(rg.dy, zk, rg.oy); pz = fmaf(rg.dz, zk, rg.oz);
          vx = rg.vx; vy = rg.vy; vz = rg.vz;
          dist = ((k + 1 < N) ? (z1 - zk) : 1e10f) * rg.dnorm;
        } else {
          const float* pp = a.pts + (gray * N + k) * 3;
          px = pp[0]; py = pp[1]; pz = pp[2];
          const float* vv = a.viewdirs + gray * 3;
          vx = vv[0]; vy = vv[1]; vz = vv[2];
          const float* rd = a.rays_d + gray * 3;
          const float dn = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
          zk = a.z_vals[gray * N + k];
          dist = ((k + 1 < N) ? (a.z_vals[gray * N + k + 1] - zk) : 1e10f) * dn;
        }
        if (a.z_vals_out && valid) a.z_vals_out[gray * N + k] = zk;
        const float nx = px * nscale, ny = py * nscale, nz = pz * nscale;
        float sdf = 0.f;

        // ------------------------------------------------ layer 0 on the FP32 pipe (K = 3)
#pragma unroll 4
        for (int c8 = 0; c8 < 32; ++c8) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 tt = tab0S[c8 * 8 + i];
            o[i] = __sinf(fmaf(tt.x, nx, fmaf(tt.y, ny, fmaf(tt.z, nz, tt.w))));
            if (D == 1) sdf = fmaf(wsigS[c8 * 8 + i], o[i], sdf);
          }
          store8(act_row, row7, c8, o);
        }
        // film table of the first MMA layer (hidden l=1, or the view layer when D == 1) is fetched below
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&misc->a_ready[s]);

        // ------------------------------------------------ hidden layers 1..D-1 (thread = point)
        for (int l = 1; l < D; ++l) {
          float2* fS = filmS + (l & 1) * W;
          fS[t] = film_img[l * W + t];
          fS[t + TILE] = film_img[l * W + t + TILE];
          named_bar_sync(bar_id, TILE);
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          jobcnt++;
          tc_fence_after();
          const bool last = (l == D - 1);
          uint32_t v[2][32];
          tmem_ld_32x32(tacc, v[0]);
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            tmem_ld_wait();
            if (cc + 1 < 8) tmem_ld_32x32(tacc + (cc + 1) * 32, v[(cc + 1) & 1]);
            const uint32_t(&vv)[32] = v[cc & 1];
#pragma unroll
            for (int i8 = 0; i8 < 4; ++i8) {
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                const int c = cc * 32 + i8 * 8 + i;
                const float4 f = *reinterpret_cast<const float4*>(fS + c);   // (scale,shift) of c and c+1
                o[i] = __sinf(fmaf(__uint_as_float(vv[i8 * 8 + i]), f.x, f.y));
                o[i + 1] = __sinf(fmaf(__uint_as_float(vv[i8 * 8 + i + 1]), f.z, f.w));
                if (last) sdf = fmaf(wsigS[c], o[i], fmaf(wsigS[c + 1], o[i + 1], sdf));
              }
              store8(act_row, row7, cc * 4 + i8, o);
            }
          }
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(&misc->a_ready[s]);
        }

        // ------------------------------------------------ density -> alpha -> transmittance (nerf_utils.py:267-307)
        sdf += bsig;
        if (valid) a.sdf[gray * N + k] = sdf;
        const float sigma = sigmoid_precise(-sdf * inv_beta) * inv_beta;
        const float alpha = 1.0f - expf(-sigma * dist);
        const float om = 1.0f - alpha + 1e-10f;
        omS[t] = valid ? om : 1.0f;
        named_bar_sync(bar_id, TILE);
        const int first_row = t - k;               // row of sample 0 of my ray (negative: began in an earlier tile)
        float T = first_row < 0 ? misc->carry[s] : 1.0f;
        for (int m = max(first_row, 0); m < t; ++m) T *= omS[m];
        const float wgt = valid ? alpha * T : 0.f;
        ptS[t] = make_float4(wgt, vx, vy, vz);
        flagS[t] = (rl << 3) | (valid ? 4 : 0) | ((valid && k == N - 1) ? 2 : 0) | ((valid && k == 0) ? 1 : 0);
        named_bar_sync(bar_id, TILE);
        if (t == TILE - 1) misc->carry[s] = (k == N - 1) ? 1.0f : T * om;

        // ------------------------------------------------ view layer (thread = channel), features composited in registers
        {
          const float2 f0 = film_img[D * W + t], f1 = film_img[D * W + t + TILE];
          const float4 tv0 = a.view[(size_t)img * W + t], tv1 = a.view[(size_t)img * W + t + TILE];
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          jobcnt++;
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float scale = h ? f1.x : f0.x, shift0 = h ? f1.y : f0.y;
            const float4 tv = h ? tv1 : tv0;
            float cu = cur[h], sr = shift_ray[h];
            float* fout = a.feature_map + ((size_t)img * a.n_rays + r0) * W + t + h * TILE;
            uint8_t* abase = act + ((t + h * TILE) >> 6) * ACT_CHUNK;
            uint32_t v[2][32];
            tmem_ld_32x32(tacc + h * 128, v[0]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              tmem_ld_wait();
              if (cc + 1 < 4) tmem_ld_32x32(tacc + h * 128 + (cc + 1) * 32, v[(cc + 1) & 1]);
              const uint32_t(&vv)[32] = v[cc & 1];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int p = cc * 32 + i;
                const float4 pw = ptS[p];
                const int fl = flagS[p];
                if (fl & 1) sr = fmaf(tv.x, pw.y, fmaf(tv.y, pw.z, fmaf(tv.z, pw.w, shift0)));
                const float feat = __sinf(fmaf(__uint_as_float(vv[i]), scale, sr));
                cu = fmaf(pw.x, feat, cu);
                *reinterpret_cast<__nv_bfloat16*>(abase + p * 128 + xoroff[p & 7]) = __float2bfloat16_rn(feat);
                if (fl & 2) { fout[(size_t)(fl >> 3) * W] = cu; cu = 0.f; }
              }
            }
            cur[h] = cu; shift_ray[h] = sr;
          }
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(&misc->a_ready[s]);
        }

        // ------------------------------------------------ rgb head + per-ray sums (nerf_utils.py:315,329-336)
        {
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          jobcnt++;
          tc_fence_after();
          uint32_t v4[4];
          tmem_ld_32x4(tacc, v4);
          tmem_ld_wait();
          tc_fence_before();
          float vals[6];
          vals[0] = wgt * sigmoid_precise(__uint_as_float(v4[0]) + brgb0);
          vals[1] = wgt * sigmoid_precise(__uint_as_float(v4[1]) + brgb1);
          vals[2] = wgt * sigmoid_precise(__uint_as_float(v4[2]) + brgb2);
          vals[3] = wgt * px; vals[4] = wgt * py; vals[5] = wgt * pz;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int rid = __shfl_down_sync(0xffffffffu, rl, o);
            const bool same = (lane + o < 32) && (rid == rl);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const float y = __shfl_down_sync(0xffffffffu, vals[j], o);
              if (same) vals[j] += y;
            }
          }
          const int rprev = __shfl_up_sync(0xffffffffu, rl, 1);
          float* racc = rayacc + (rl & (RSLOTS - 1)) * 8;
          if (valid && (lane == 0 || rprev != rl)) {
#pragma unroll
            for (int j = 0; j < 6; ++j) atomicAdd(racc + j, vals[j]);
          }
          named_bar_sync(bar_id, TILE);
          if (valid && k == N - 1) {
            const float x = racc[3], y = racc[4], z = racc[5];
            float* o3 = a.rgb_map + gray * 3;
            o3[0] = -1.0f + 2.0f * racc[0]; o3[1] = -1.0f + 2.0f * racc[1]; o3[2] = -1.0f + 2.0f * racc[2];
            float* x3 = a.xyz + gray * 3;
            x3[0] = x; x3[1] = y; x3[2] = z;
            a.mask[gray * 2 + 0] = wgt;
            a.mask[gray * 2 + 1] = -sqrtf(x * x + y * y + z * z);
#pragma unroll
            for (int j = 0; j < 6; ++j) racc[j] = 0.f;
          }
        }
      }  // tiles
    }    // units
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------
// Self-test tile product through the same descriptors / layouts:  D[128][N] = A[128][K] * B[N][K]^T
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                                float* __restrict__ Dout, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // K/64 chunks x [128][64]
  uint8_t* sB = smem + ACT_BYTES;           // K/64 chunks x [N][64]
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const int nchunk = K / 64;
  for (int idx = threadIdx.x; idx < 128 * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    *reinterpret_cast<uint16_t*>(sA + (k >> 6) * ACT_CHUNK + sw128_offset(r, k & 63)) = A[idx];
  }
  for (int idx = threadIdx.x; idx < N * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    *reinterpret_cast<uint16_t*>(sB + (k >> 6) * (N * 128) + sw128_offset(r, k & 63)) = B[idx];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tbase, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)N);
    for (int c = 0; c < nchunk; ++c) {
      const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(sA + c * ACT_CHUNK));
      const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(sB + c * (N * 128)));
      for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tb, ad + 2 * kk, bd + 2 * kk, idesc, (c | kk) != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 4) {
    uint32_t v4[4];
    tmem_ld_32x4(taddr + c0, v4);
    tmem_ld_wait();
    for (int j = 0; j < 4; ++j) Dout[(size_t)threadIdx.x * N + c0 + j] = __uint_as_float(v4[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

}}  // namespace c3d::fused
