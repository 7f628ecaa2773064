// Fused NeRF-branch forward for sm_100a, version 2: every layer runs with the operand roles swapped
// (D^T[channel][point] = W * H^T) so TMEM lanes are channels and columns are points.
//
// Why: with rows = points (version 1) every epilogue thread needs the per-channel FiLM (scale, shift) of all 256
// columns; those warp-uniform shared-memory loads cost one wavefront per 4 bytes per warp and saturated the shared
// memory pipe at 24 % tensor utilisation (profiles/r01_fused_v1.md).  With lanes = channels the FiLM constants of a
// thread are two registers, the sdf / rgb heads and layer 0 become small MMAs, and the only shared-memory traffic
// of a hidden layer is the bf16 activation store.
//
// Reference semantics: exp/cips3d/volume_renderer.py:133-160,192-283, exp/cips3d/nerf_utils.py:17-218,230-338.
//
// Per 128-point tile the MMA issuer runs D+3 jobs; each is "wait a_ready[slot] -> MMAs -> commit acc_full[slot]"
// and the slot's epilogue group answers each with "wait acc_full -> epilogue -> arrive a_ready":
//   job 0        layer 0   : 2 x (128x128x16)  A = w0img (hi/lo split of W0), B = point tile (hi/mid/lo split)
//   job 1..D-1   hidden l  : 2 x 16 x (128x128x16), A = weight stage rows [128h,128h+128), B = activation tile
//   job D        heads/sdf : 16 x (128x16x16)  A = activation tile (rows = points), B = heads16 -> sdf per point
//   job D+1      view layer: like hidden; epilogue composites features in registers (fp32) per ray
//   job D+2      heads/rgb : as job D on the bf16 feature tile -> raw rgb per point
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"
#include "fused_bf16_sm100.cuh"   // Args, slot/unit helpers

namespace c3d { namespace fused2 {

using namespace c3d::ptx;
using fused::Args;
using fused::unit_tiles;
using fused::slot_tiles;

constexpr int NTHREADS = 384;
constexpr int TILE = 128;
constexpr int ACT_CHUNK = TILE * 128;          // 16384
constexpr int ACT_BYTES = NCHUNK * ACT_CHUNK;  // 65536
constexpr int STAGE_BYTES = W * 128;           // 32768
constexpr int NSTAGE = 2;
constexpr int RSLOTS = 32;
constexpr int P16_BYTES = TILE * 32;           // 4096: [128 points][16 k] bf16

constexpr int SM_ACT = 0;
constexpr int SM_STAGE = SM_ACT + 2 * ACT_BYTES;                 // 131072
constexpr int SM_HEADS = SM_STAGE + NSTAGE * STAGE_BYTES;        // 196608
constexpr int SM_W0 = SM_HEADS + (int)RGB16_BYTES;               // 204800
constexpr int SM_P16 = SM_W0 + (int)W0IMG_BYTES;                 // 212992  [slot][4096]
constexpr int SM_PT = SM_P16 + 2 * P16_BYTES;                    // 221184  [slot][128] float4 (w, vx, vy, vz)
constexpr int SM_OM = SM_PT + 2 * TILE * 16;                     // 225280  [slot][128] float
constexpr int SM_FLAG = SM_OM + 2 * TILE * 4;                    // 226304  [slot][128] int
constexpr int SM_RAYACC = SM_FLAG + 2 * TILE * 4;                // 227328  [slot][RSLOTS][8] float
constexpr int SM_MISC = SM_RAYACC + 2 * RSLOTS * 8 * 4;          // 229376
constexpr int SM_TOTAL = SM_MISC + 256;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;

struct Misc {
  uint64_t full[NSTAGE], empty[NSTAGE], a_ready[2], acc_full[2];
  uint32_t tmem_base;
  float carry[2];
};

// bf16(x) -> 2-byte shared-memory store (F2FP + STS.U16, no repacking)
__device__ __forceinline__ void st_bf16(uint32_t smem_addr, float x) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(x));
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(smem_addr), "h"((unsigned short)r) : "memory");
}

// packed weight-layer index streamed by job j of a tile (-1: the job uses resident operands only)
__device__ __forceinline__ int job_layer(int j, int D) {
  if (j >= 1 && j <= D - 1) return j - 1;   // hidden layer l = j  -> wbf16[l-1]
  if (j == D + 1) return D - 1;             // view layer          -> wbf16[D-1]
  return -1;
}

template <int kCluster>
__global__ void __launch_bounds__(NTHREADS, 1) fused_forward_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D, N = a.n_samples;
  const int JOBS = D + 3;
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
    const uint4* src = reinterpret_cast<const uint4*>(a.blob + a.L.rgb16);      // heads16 + w0img are adjacent
    uint4* dst = reinterpret_cast<uint4*>(smem + SM_HEADS);
    for (int i = threadIdx.x; i < (int)(RGB16_BYTES + W0IMG_BYTES) / 16; i += NTHREADS) dst[i] = src[i];
    float* ra = reinterpret_cast<float*>(smem + SM_RAYACC);
    for (int i = threadIdx.x; i < 2 * RSLOTS * 8; i += NTHREADS) ra[i] = 0.f;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = misc->tmem_base;

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
      const int layer = job_layer(g % JOBS, D);
      if (layer < 0) continue;
      for (int s = 0; s < 2; ++s) {
        for (int c = 0; c < NCHUNK; ++c, ++n) {
          const uint32_t st = n & 1u, ph = (n >> 1) & 1u;
          mbar_wait(&misc->empty[st], ph ^ 1u);
          uint8_t* dst = smem + SM_STAGE + st * STAGE_BYTES;
          const uint8_t* src = wsrc + (size_t)layer * WBF16_LAYER_BYTES + (size_t)c * WBF16_CHUNK_BYTES;
          if (a.debug & 1) { mbar_arrive(&misc->full[st]); continue; }     // timing experiment: no weight traffic
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
    const uint32_t idesc_t = umma_idesc_bf16(128, 128), idesc_h = umma_idesc_bf16(128, 16);
    const uint32_t act_addr[2] = {smem_u32(smem + SM_ACT), smem_u32(smem + SM_ACT + ACT_BYTES)};
    const uint32_t stage_addr[2] = {smem_u32(smem + SM_STAGE), smem_u32(smem + SM_STAGE + STAGE_BYTES)};
    const uint32_t p16_addr[2] = {smem_u32(smem + SM_P16), smem_u32(smem + SM_P16 + P16_BYTES)};
    const uint32_t heads_addr = smem_u32(smem + SM_HEADS), w0_addr = smem_u32(smem + SM_W0);
    uint32_t n = 0, jobcnt[2] = {0u, 0u};
    for (int g = 0; g < rounds; ++g) {
      const int j = g % JOBS, tile_idx = g / JOBS;
      const int layer = job_layer(j, D);
      for (int s = 0; s < 2; ++s) {
        const bool real = tile_idx < my_tiles[s];
        const uint32_t tacc = tmem_base + (uint32_t)s * 256u;
        if (layer >= 0) {
          if (real) { mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u); tc_fence_after(); }
          for (int c = 0; c < NCHUNK; ++c, ++n) {
            const uint32_t st = n & 1u, ph = (n >> 1) & 1u;
            mbar_wait(&misc->full[st], ph);
            tc_fence_after();
            if (real) {
              const uint64_t bd = umma_desc_kmajor_sw128(act_addr[s] + c * ACT_CHUNK);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t ad = umma_desc_kmajor_sw128(stage_addr[st] + h * (STAGE_BYTES / 2));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16_ss(tacc + (uint32_t)h * 128u, ad + 2 * kk, bd + 2 * kk, idesc_t, (c | kk) != 0);
              }
            }
            if (kCluster == 1) umma_commit(&misc->empty[st]);
            else umma_commit_multicast(&misc->empty[st], (uint16_t)0x3);
          }
          if (real) { umma_commit(&misc->acc_full[s]); jobcnt[s]++; }
        } else if (real) {
          mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u);
          tc_fence_after();
          if (j == 0) {                     // layer 0: K = 16 split product, channels on lanes
            const uint64_t bd = umma_desc_kmajor_k16(p16_addr[s]);
#pragma unroll
            for (int h = 0; h < 2; ++h)
              umma_bf16_ss(tacc + (uint32_t)h * 128u, umma_desc_kmajor_k16(w0_addr + h * 4096), bd, idesc_t, 0u);
          } else {                          // heads (sdf after the last hidden layer, rgb after the view layer)
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
              const uint64_t ad = umma_desc_kmajor_sw128(act_addr[s] + c * ACT_CHUNK);
              const uint64_t bd = umma_desc_kmajor_sw128(heads_addr + c * 2048);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tacc, ad + 2 * kk, bd + 2 * kk, idesc_h, (c | kk) != 0);
            }
          }
          umma_commit(&misc->acc_full[s]);
          jobcnt[s]++;
        }
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue groups
    const int s = (warp - 4) >> 2;
    const int t = threadIdx.x - 128 - s * TILE;          // 0..127: point row (point stages) / channel (layer stages)
    const int quad = warp & 3;
    const uint32_t bar_id = 1u + (uint32_t)s;
    const int slot = 2 * blockIdx.x + s;
    uint8_t* act = smem + SM_ACT + s * ACT_BYTES;
    uint8_t* p16 = smem + SM_P16 + s * P16_BYTES;
    float4* ptS = reinterpret_cast<float4*>(smem + SM_PT) + s * TILE;
    float* omS = reinterpret_cast<float*>(smem + SM_OM) + s * TILE;
    int* flagS = reinterpret_cast<int*>(smem + SM_FLAG) + s * TILE;
    float* rayacc = reinterpret_cast<float*>(smem + SM_RAYACC) + s * RSLOTS * 8;
    const uint32_t tacc = tmem_base + (uint32_t)s * 256u + ((uint32_t)(quad * 32) << 16);
    const float* scal = reinterpret_cast<const float*>(a.blob + a.L.scal);
    const float bsig = scal[0], brgb0 = scal[1], brgb1 = scal[2], brgb2 = scal[3];
    const float inv_beta = 1.0f / scal[4];
    // bf16 store of channel t (+128) of point p: act + chunk*16384 + p*128 + xoroff[p&7]
    uint32_t xoroff[8];
#pragma unroll
    for (int jx = 0; jx < 8; ++jx) xoroff[jx] = (uint32_t)((((t & 63) >> 3) ^ jx) << 4) + (uint32_t)((t & 7) << 1);
    const uint32_t abase_u32[2] = {smem_u32(act + (t >> 6) * ACT_CHUNK), smem_u32(act + ((t + TILE) >> 6) * ACT_CHUNK)};
    float cur[2] = {0.f, 0.f}, shift_ray[2] = {0.f, 0.f};
    uint32_t jobcnt = 0;
    const int total_units = a.batch * a.units_per_img;

    for (int u = slot; u < total_units; u += nslots) {
      const int img = u / a.units_per_img;
      const int r0 = (u - img * a.units_per_img) * a.unit_rays;
      const int nr = min(a.unit_rays, a.n_rays - r0);
      const int npts = nr * N;
      const int ntiles = (npts + TILE - 1) / TILE;
      const float near = a.near[img], far = a.far[img];
      const float nscale = 2.0f / (far - near);
      const float2* film_img = a.film + (size_t)img * (D + 1) * W;
      const float4 tv0 = a.view[(size_t)img * W + t], tv1 = a.view[(size_t)img * W + t + TILE];
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
        {
          // point tile for the layer-0 MMA: per coordinate (hi, mid, lo, hi, mid) in bf16 (24 bits of the fp32 value)
          const float pn[3] = {px * nscale, py * nscale, pz * nscale};
          float e[16];
#pragma unroll
          for (int jx = 0; jx < 3; ++jx) {
            const float hi = __bfloat162float(__float2bfloat16_rn(pn[jx]));
            const float r1 = pn[jx] - hi;
            const float mid = __bfloat162float(__float2bfloat16_rn(r1));
            const float lo = __bfloat162float(__float2bfloat16_rn(r1 - mid));
            e[4 * jx + 0] = hi; e[4 * jx + 1] = mid; e[4 * jx + 2] = hi; e[4 * jx + 3] = lo;
          }
          e[12] = e[13] = e[14] = e[15] = 0.f;
          uint4 lo8, hi8;                       // values are bf16-exact: the packing below does not round
          lo8.x = pack_bf16x2(e[0], e[1]); lo8.y = pack_bf16x2(e[2], e[3]);
          lo8.z = pack_bf16x2(e[4], e[5]); lo8.w = pack_bf16x2(e[6], e[7]);
          hi8.x = pack_bf16x2(e[8], e[9]); hi8.y = pack_bf16x2(e[10], e[11]);
          hi8.z = pack_bf16x2(e[12], e[13]); hi8.w = pack_bf16x2(e[14], e[15]);
          uint8_t* row = p16 + (t >> 3) * 256 + (t & 7) * 16;
          *reinterpret_cast<uint4*>(row) = lo8;
          *reinterpret_cast<uint4*>(row + 128) = hi8;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&misc->a_ready[s]);

        // ------------------------------------------------ layers 0..D-1: thread = channel t and t+128
        for (int l = 0; l < D; ++l) {
          const float2 f0 = film_img[l * W + t], f1 = film_img[l * W + t + TILE];
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          jobcnt++;
          tc_fence_after();
          {
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(tacc, v0);
#pragma unroll 1
            for (int hc = 0; hc < 4; ++hc) {               // 64 points of one channel half per iteration
              const int h = hc >> 1, p0 = (hc & 1) * 64;
              const float scale = h ? f1.x : f0.x, shift = h ? f1.y : f0.y;
              const uint32_t ab = abase_u32[h] + (uint32_t)p0 * 128u;
              tmem_ld_wait();
              tmem_ld_32x32(tacc + h * 128 + p0 + 32, v1);
#pragma unroll
              for (int i = 0; i < 32; ++i)
                st_bf16(ab + i * 128 + xoroff[i & 7], __sinf(fmaf(__uint_as_float(v0[i]), scale, shift)));
              tmem_ld_wait();
              if (hc < 3) tmem_ld_32x32(tacc + ((hc + 1) >> 1) * 128 + ((hc + 1) & 1) * 64, v0);
#pragma unroll
              for (int i = 0; i < 32; ++i)
                st_bf16(ab + (32 + i) * 128 + xoroff[i & 7], __sinf(fmaf(__uint_as_float(v1[i]), scale, shift)));
            }
          }
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(&misc->a_ready[s]);
        }

        // ------------------------------------------------ sdf head (thread = point) -> alpha -> transmittance
        mbar_wait(&misc->acc_full[s], jobcnt & 1u);
        jobcnt++;
        tc_fence_after();
        float sdf;
        {
          uint32_t v4[4];
          tmem_ld_32x4(tacc + 4, v4);              // heads16 rows 4, 5: hi / lo part of sigma_linear.weight
          tmem_ld_wait();
          tc_fence_before();
          sdf = __uint_as_float(v4[0]) + __uint_as_float(v4[1]) + bsig;
        }
        if (valid) a.sdf[gray * N + k] = sdf;
        const float sigma = sigmoid_precise(-sdf * inv_beta) * inv_beta;
        const float alpha = 1.0f - expf(-sigma * dist);
        const float om = 1.0f - alpha + 1e-10f;
        omS[t] = valid ? om : 1.0f;
        named_bar_sync(bar_id, TILE);
        const int first_row = t - k;
        float T = first_row < 0 ? misc->carry[s] : 1.0f;
        for (int m = max(first_row, 0); m < t; ++m) T *= omS[m];
        const float wgt = valid ? alpha * T : 0.f;
        ptS[t] = make_float4(wgt, vx, vy, vz);
        flagS[t] = (rl << 3) | (valid ? 4 : 0) | ((valid && k == N - 1) ? 2 : 0) | ((valid && k == 0) ? 1 : 0);
        named_bar_sync(bar_id, TILE);
        if (t == TILE - 1) misc->carry[s] = (k == N - 1) ? 1.0f : T * om;
        mbar_arrive(&misc->a_ready[s]);          // accumulator drained -> the view-layer MMAs may start

        // ------------------------------------------------ view layer (thread = channel), features composited in registers
        {
          const float2 f0 = film_img[D * W + t], f1 = film_img[D * W + t + TILE];
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          jobcnt++;
          tc_fence_after();
          {
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(tacc, v0);
            float* const fbase = a.feature_map + ((size_t)img * a.n_rays + r0) * W + t;
#pragma unroll 1
            for (int hc = 0; hc < 4; ++hc) {
              const int h = hc >> 1, p0 = (hc & 1) * 64;
              const float scale = h ? f1.x : f0.x, shift0 = h ? f1.y : f0.y;
              const float4 tv = h ? tv1 : tv0;
              float cu = cur[h], sr = shift_ray[h];
              float* fout = fbase + h * TILE;
              const uint32_t ab = abase_u32[h] + (uint32_t)p0 * 128u;
              tmem_ld_wait();
              tmem_ld_32x32(tacc + h * 128 + p0 + 32, v1);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float4 pw = ptS[p0 + i];
                const int fl = flagS[p0 + i];
                if (fl & 1) sr = fmaf(tv.x, pw.y, fmaf(tv.y, pw.z, fmaf(tv.z, pw.w, shift0)));
                const float feat = __sinf(fmaf(__uint_as_float(v0[i]), scale, sr));
                cu = fmaf(pw.x, feat, cu);
                st_bf16(ab + i * 128 + xoroff[i & 7], feat);
                if (fl & 2) { fout[(size_t)(fl >> 3) * W] = cu; cu = 0.f; }
              }
              tmem_ld_wait();
              if (hc < 3) tmem_ld_32x32(tacc + ((hc + 1) >> 1) * 128 + ((hc + 1) & 1) * 64, v0);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float4 pw = ptS[p0 + 32 + i];
                const int fl = flagS[p0 + 32 + i];
                if (fl & 1) sr = fmaf(tv.x, pw.y, fmaf(tv.y, pw.z, fmaf(tv.z, pw.w, shift0)));
                const float feat = __sinf(fmaf(__uint_as_float(v1[i]), scale, sr));
                cu = fmaf(pw.x, feat, cu);
                st_bf16(ab + (32 + i) * 128 + xoroff[i & 7], feat);
                if (fl & 2) { fout[(size_t)(fl >> 3) * W] = cu; cu = 0.f; }
              }
              cur[h] = cu; shift_ray[h] = sr;
            }
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
            for (int jx = 0; jx < 6; ++jx) {
              const float y = __shfl_down_sync(0xffffffffu, vals[jx], o);
              if (same) vals[jx] += y;
            }
          }
          const int rprev = __shfl_up_sync(0xffffffffu, rl, 1);
          float* racc = rayacc + (rl & (RSLOTS - 1)) * 8;
          if (valid && (lane == 0 || rprev != rl)) {
#pragma unroll
            for (int jx = 0; jx < 6; ++jx) atomicAdd(racc + jx, vals[jx]);
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
            for (int jx = 0; jx < 6; ++jx) racc[jx] = 0.f;
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

// Self-test of the K = 16 no-swizzle operand layout:  D[128][128] = A[128][16] * B[128][16]^T
// swap != 0 builds the descriptors with leading/stride byte offsets exchanged (diagnostic only).
__global__ void __launch_bounds__(128, 1) umma_k16_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                                    float* __restrict__ Dout, int swap) {
  __shared__ __align__(1024) uint8_t sA[4096];
  __shared__ __align__(1024) uint8_t sB[4096];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  for (int idx = threadIdx.x; idx < 128 * 16; idx += 128) {
    const int r = idx >> 4, k = idx & 15;
    *reinterpret_cast<uint16_t*>(sA + k16_offset(r, k)) = A[idx];
    *reinterpret_cast<uint16_t*>(sB + k16_offset(r, k)) = B[idx];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tbase, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    uint64_t ad = umma_desc_kmajor_k16(smem_u32(sA)), bd = umma_desc_kmajor_k16(smem_u32(sB));
    if (swap) {
      const uint64_t m = ((uint64_t)0x3FFF << 16) | ((uint64_t)0x3FFF << 32);
      const uint64_t sw = ((uint64_t)(256u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32);
      ad = (ad & ~m) | sw;
      bd = (bd & ~m) | sw;
    }
    umma_bf16_ss(tb, ad, bd, umma_idesc_bf16(128, 128), 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 128; c0 += 4) {
    uint32_t v4[4];
    tmem_ld_32x4(taddr + c0, v4);
    tmem_ld_wait();
    for (int jx = 0; jx < 4; ++jx) Dout[(size_t)threadIdx.x * 128 + c0 + jx] = __uint_as_float(v4[jx]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 128); }
}


// Self-test of MN-major SWIZZLE_128B operands:  D[128][N] = A[128][K] * B[N][K]^T with A and/or B stored [k][mn].
// variant bit 0: A MN-major, bit 1: B MN-major, bit 2: exchange the LBO / SBO fields (diagnostic).
__global__ void __launch_bounds__(128, 1) umma_mn_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                                   float* __restrict__ Dout, int N, int K, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 65536;
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const bool a_mn = variant & 1, b_mn = variant & 2;
  const uint32_t blk = (uint32_t)K * 128u;                   // bytes of one 64-wide mn block: K rows of 128 B
  for (int idx = threadIdx.x; idx < 128 * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    const uint32_t off = a_mn ? (uint32_t)(r >> 6) * blk + (uint32_t)k * 128u + (uint32_t)((((r & 63) >> 3) ^ (k & 7)) << 4) + (uint32_t)(r & 7) * 2u
                              : (uint32_t)(k >> 6) * ACT_CHUNK + sw128_offset(r, k & 63);
    *reinterpret_cast<uint16_t*>(sA + off) = A[idx];
  }
  for (int idx = threadIdx.x; idx < N * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    const uint32_t off = b_mn ? (uint32_t)(r >> 6) * blk + (uint32_t)k * 128u + (uint32_t)((((r & 63) >> 3) ^ (k & 7)) << 4) + (uint32_t)(r & 7) * 2u
                              : (uint32_t)(k >> 6) * (uint32_t)(N * 128) + sw128_offset(r, k & 63);
    *reinterpret_cast<uint16_t*>(sB + off) = B[idx];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tbase, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)N, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad, bd;
      if (a_mn) ad = umma_desc_mnmajor_sw128(smem_u32(sA) + ks * 2048, blk);
      else ad = umma_desc_kmajor_sw128(smem_u32(sA + (ks >> 2) * ACT_CHUNK)) + 2 * (ks & 3);
      if (b_mn) bd = umma_desc_mnmajor_sw128(smem_u32(sB) + ks * 2048, blk);
      else bd = umma_desc_kmajor_sw128(smem_u32(sB + (ks >> 2) * (N * 128))) + 2 * (ks & 3);
      if (variant & 4) {
        const uint64_t m = ((uint64_t)0x3FFF << 16) | ((uint64_t)0x3FFF << 32);
        const uint64_t sw = ((uint64_t)(1024u >> 4) << 16) | ((uint64_t)((blk >> 4) & 0x3FFF) << 32);
        if (a_mn) ad = (ad & ~m) | sw;
        if (b_mn) bd = (bd & ~m) | sw;
      }
      umma_bf16_ss(tb, ad, bd, idesc, ks != 0);
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
    for (int jx = 0; jx < 4; ++jx) Dout[(size_t)threadIdx.x * N + c0 + jx] = __uint_as_float(v4[jx]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

}}  // namespace c3d::fused2
