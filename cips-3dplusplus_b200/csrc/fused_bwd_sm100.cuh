// Fused backward of the point MLP on tcgen05 (bf16 operands, fp32 accumulation) -- the tensor-core counterpart of
// mlp_bwd_kernel (backward.cuh), same persistent two-slot structure as the forward kernel.
//
// Inputs per tile come from the forward kernel's save mode (fp16 acc tiles per layer; cos(arg) is recomputed here), from the
// compositing backward (per-point weights, d rgb, d sdf) and from the caller (d feature_map).  Orientation as in the
// forward: TMEM lanes = channels, columns = points, so for every layer the epilogue thread of channel c computes
//     g_a = g_h * cos(arg)          (cotangent of the SIREN argument)
//     G1[c] += g_a * acc, G2[c] += g_a       (column sums -> d gamma = G1 + b G2, d beta = G2; registers, no shuffles)
//     g_acc = gamma[c] * g_a  -> bf16 -> G^T tile (same [channel][point] swizzled layout as the forward's H^T tile)
// and one MMA job per layer propagates  g_h_{l-1}^T[k][p] = sum_c W_l[c][k] * g_acc[c][p]  with the transposed weight
// images streamed through the same 4 x 16 KB ring.
//
// Jobs per 128-point tile (D + 3):
//   0      view-layer cotangent  g_f^T = Wrgb^T g_rgb (K=16 side product) + gF^T Wgt (K=16: ray slots)
//   1      d viewdirs            D[p][j] = sum_c g_acc_D[c][p] Wview[c][256+j]   (heads-style, rows = points)
//   2..D+1 layers D..1           g_h^T = W^T g_acc  (+ sigma-head rank-1 term for the first of them)
//   D+2    d points              D[p][j] = sum_c g_acc_0[c][p] W0[c][j]
#pragma once
#include <cuda_fp16.h>
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"
#include "fused_common.cuh"
#include "fused_bf16_sm100.cuh"

namespace c3d { namespace fusedbwd {

using namespace c3d::ptx;
using fused::Args;
using fused::slot_tiles;
using fused::TILE;
using fused::ACT_BYTES;
using fused::ACT_PBLOCK;
using fused::tmem_ld_32x16;
using fused::st_v4;

constexpr int EGW = 8;                                            // epilogue warps per slot: a thread owns ONE channel
constexpr int EG_THREADS = EGW * 32;
constexpr int NTHREADS = 128 + 2 * EG_THREADS;                    // 4 role warps + two epilogue groups
constexpr int PF = 2;                    // iterations (of 16 points) whose saved-row loads are in flight per thread (1..4 measured)
constexpr int STAGE_BYTES = 16384;
constexpr int NSTAGE = 4;
constexpr int RAYS = 16;

constexpr int SM_ACT = 0;                                         // G^T tiles (the first 12 KB double as the job-0 operands)
constexpr int SM_STAGE = SM_ACT + 2 * ACT_BYTES;                  // 131072
constexpr int SM_IMG = SM_STAGE + NSTAGE * STAGE_BYTES;           // 196608  bwd16 images 0, 1, 2
constexpr int SM_AUX = SM_IMG + 3 * (int)W0IMG_BYTES;             // 221184  [slot][128 p][16] k16: d rgb / d sdf tile
constexpr int SM_MISC = SM_AUX + 2 * 4096;                        // 229376
constexpr int SM_TOTAL = SM_MISC + 256;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;

struct Misc {
  uint64_t full[NSTAGE], empty[NSTAGE], a_ready[2], acc_full[2];
  uint32_t tmem_base;
};

// ring job j of a tile streams the transposed weights of layer D - (j - 2)  (index layer-1 in wbf16T)
__device__ __forceinline__ int job_layer(int j, int D) { return (j >= 2 && j <= D + 1) ? (D - (j - 2)) - 1 : -1; }

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

template <int kCluster>
__global__ void __launch_bounds__(NTHREADS, 1) fused_backward_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int D = a.D, N = a.n_samples;
  const int JOBS = D + 3;
  const int nslots = 2 * gridDim.x;
  const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0u;

  if (threadIdx.x == 32) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&misc->full[i], 1); mbar_init(&misc->empty[i], kCluster); }
    for (int i = 0; i < 2; ++i) { mbar_init(&misc->a_ready[i], EG_THREADS); mbar_init(&misc->acc_full[i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(&misc->tmem_base, 512); tmem_relinquish(); }
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.blob + a.L.bwd16);
    uint4* dst = reinterpret_cast<uint4*>(smem + SM_IMG);
    for (int i = threadIdx.x; i < 3 * (int)W0IMG_BYTES / 16; i += NTHREADS) dst[i] = src[i];
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

  if (warp == 0) {
   if (elect_one()) {
    // ============================================================ weight producer (transposed images)
    const uint8_t* wsrc = a.blob + a.L.wbf16T;
    uint32_t n = 0;
    for (int g = 0; g < rounds; ++g) {
      const int layer = job_layer(g % JOBS, D);
      if (layer < 0) continue;
      for (int s = 0; s < 2; ++s) {
        for (int c = 0; c < 2 * NCHUNK; ++c, ++n) {
          const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
          mbar_wait(&misc->empty[st], ph ^ 1u);
          uint8_t* dst = smem + SM_STAGE + st * STAGE_BYTES;
          const uint8_t* src = wsrc + (size_t)layer * WBF16_LAYER_BYTES + (size_t)c * STAGE_BYTES;
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
   }
  } else if (warp == 1) {
   if (elect_one()) {
    // ============================================================ MMA issuer
    const uint32_t idesc_kk = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_l = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t idesc_h = umma_idesc_bf16(128, 16, 1, 0);
    const uint32_t act_addr[2] = {smem_u32(smem + SM_ACT), smem_u32(smem + SM_ACT + ACT_BYTES)};
    const uint32_t stage_base = smem_u32(smem + SM_STAGE);
    const uint32_t aux_addr[2] = {smem_u32(smem + SM_AUX), smem_u32(smem + SM_AUX + 4096)};
    const uint32_t img0 = smem_u32(smem + SM_IMG), img1 = img0 + (uint32_t)W0IMG_BYTES, img2 = img1 + (uint32_t)W0IMG_BYTES;
    uint32_t n = 0, jobcnt[2] = {0u, 0u};
    for (int g = 0; g < rounds; ++g) {
      const int j = g % JOBS, tile_idx = g / JOBS;
      const int layer = job_layer(j, D);
      for (int s = 0; s < 2; ++s) {
        const bool real = tile_idx < my_tiles[s];
        const uint32_t tacc = tmem_base + (uint32_t)s * 256u;
        if (layer >= 0) {
          if (real) {
            mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u);
            tc_fence_after();
            if (j == 2) {                 // first ring job: sigma head, g_h_{D-1} += w_sigma * g_sdf  (K = 16 side product)
              const uint64_t bd = umma_desc_kmajor_k16(aux_addr[s]);
#pragma unroll
              for (int h = 0; h < 2; ++h)
                umma_bf16_ss(tacc + (uint32_t)h * 128u, umma_desc_kmajor_k16(img1 + h * 4096), bd, idesc_kk, 0u);
            }
          }
          const uint32_t acc0 = (j == 2) ? 1u : 0u;
          for (int c = 0; c < 2 * NCHUNK; ++c, ++n) {
            const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
            mbar_wait(&misc->full[st], ph);
            tc_fence_after();
            if (real) {
              const int kc = c >> 1, h = c & 1;
              const uint64_t ad = umma_desc_kmajor_sw128(stage_base + st * STAGE_BYTES);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16_ss(tacc + (uint32_t)h * 128u, ad + 2 * kk,
                             umma_desc_mnmajor_sw128(act_addr[s] + kc * 8192 + kk * 2048, ACT_PBLOCK), idesc_l,
                             acc0 | (uint32_t)((kc | kk) != 0));
            }
            if (kCluster == 1) umma_commit(&misc->empty[st]);
            else umma_commit_multicast(&misc->empty[st], (uint16_t)0x3);
          }
          if (real) { umma_commit(&misc->acc_full[s]); jobcnt[s]++; }
        } else if (real) {
          mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u);
          tc_fence_after();
          if (j == 0) {
            // g_f^T = Wrgb^T g_rgb (image 0 x aux tile)  +  gF^T Wgt (both staged inside the still unused G^T tile)
            const uint64_t b1 = umma_desc_kmajor_k16(aux_addr[s]);
            const uint64_t b2 = umma_desc_kmajor_k16(act_addr[s] + 8192);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              umma_bf16_ss(tacc + (uint32_t)h * 128u, umma_desc_kmajor_k16(img0 + h * 4096), b1, idesc_kk, 0u);
              umma_bf16_ss(tacc + (uint32_t)h * 128u, umma_desc_kmajor_k16(act_addr[s] + h * 4096), b2, idesc_kk, 1u);
            }
          } else {                          // heads-style: rows = points; job 1 -> d viewdirs, job D+2 -> d points
#pragma unroll 2
            for (int ks = 0; ks < 16; ++ks)
              umma_bf16_ss(tacc, umma_desc_mnmajor_sw128(act_addr[s] + ks * 2048, ACT_PBLOCK),
                           umma_desc_kmajor_sw128(img2 + (ks >> 2) * 2048) + 2 * (ks & 3), idesc_h, ks != 0);
          }
          umma_commit(&misc->acc_full[s]);
          jobcnt[s]++;
        }
      }
    }
   }
  } else if (warp >= 4) {
    // ============================================================ epilogue groups
    const int s = (warp - 4) / EGW;
    const int te = threadIdx.x - 128 - s * EG_THREADS;      // 0 .. EG_THREADS-1
    const int t = te & (TILE - 1);                          // point of the tile (point role: te < TILE) / channel inside a half
    const int quad = warp & 3;
    const bool point_role = te < TILE;
    const int h = te >> 7, ch = te;                         // my channel half / channel
    const int slot = 2 * blockIdx.x + s;
    const uint32_t aux_u32 = smem_u32(smem + SM_AUX + s * 4096);
    const uint32_t act_u32 = smem_u32(smem + SM_ACT + s * ACT_BYTES);
    const uint32_t tacc = tmem_base + (uint32_t)s * 256u + ((uint32_t)(quad * 32) << 16);
    const uint32_t tcol = tacc + (uint32_t)h * 128u;
    const uint32_t row_u32 = act_u32 + (uint32_t)ch * 128u;
    const int c7 = t & 7;
    uint32_t jobcnt = 0;
    const int total_units = a.batch * a.units_per_img;
    // cycle counters of one epilogue thread (build with -DC3D_KERNEL_PROF, run with C3D_DEBUG=2), as in the forward kernel
#ifdef C3D_KERNEL_PROF
    const bool eprof = (a.debug & 2) != 0 && blockIdx.x == 0 && te == 0;
    long long e_setup = 0, e_wait = 0, e_epi = 0, e_other = 0, e_mark = clock64(), e_begin = e_mark;
    int e_tiles = 0;
#define C3D_BPROF(acc) do { if (eprof) { const long long now_ = clock64(); acc += now_ - e_mark; e_mark = now_; } } while (0)
#else
#define C3D_BPROF(acc) do { } while (0)
#endif

    for (int u = slot; u < total_units; u += nslots) {
      const int img = u / a.units_per_img;
      const int r0 = (u - img * a.units_per_img) * a.unit_rays;
      const int nr = min(a.unit_rays, a.n_rays - r0);
      const int npts = nr * N;
      const int ntiles = (npts + TILE - 1) / TILE;
      const float nscale = 2.0f / (a.far[img] - a.near[img]);
      const float2* film_img = a.film + (size_t)img * (D + 1) * W;
      float* gfilm_img = a.g_film + (size_t)img * (D + 1) * W * 2;

      for (int tile = 0; tile < ntiles; ++tile) {
        const long long tile_g = (long long)u * a.tiles_per_unit + tile;
        const int q = tile * TILE + t;
        const bool valid = q < npts;
        const int qc = valid ? q : npts - 1;
        const int rl = qc / N, k = qc - rl * N;
        const int rl0 = (tile * TILE) / N;
        const size_t gray = (size_t)img * a.n_rays + r0 + rl;
        const size_t gpt = gray * N + k;
        // ---- job-0 operands: (a) d rgb / d sdf tile of my point, (b) Wgt row of my point, (c) gF^T rows of my channels
        {
          const uint32_t row = (uint32_t)((t >> 3) * 256 + (t & 7) * 16);
          if (point_role) {
            const float wv = valid ? a.w_pt[gpt] : 0.f;
            float e[8];
#pragma unroll
            for (int jx = 0; jx < 3; ++jx) {
              const float gr = valid ? a.g_rgb_pt[gpt * 3 + jx] : 0.f;
              const float hi = __bfloat162float(__float2bfloat16_rn(gr));
              e[jx] = hi; e[3 + jx] = gr - hi;
            }
            const float gs = valid ? a.g_sdf_pt[gpt] : 0.f;
            e[6] = __bfloat162float(__float2bfloat16_rn(gs)); e[7] = gs - e[6];
            st_v4(aux_u32 + row, pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
            st_v4(aux_u32 + row + 128, 0u, 0u, 0u, 0u);
            // Wgt[point][ray slot] (K-major k16 tile at G^T + 8192)
            const int myslot = rl - rl0;
            uint32_t wr[8];
#pragma unroll
            for (int jx = 0; jx < 8; ++jx)
              wr[jx] = pack_bf16x2(2 * jx == myslot ? wv : 0.f, 2 * jx + 1 == myslot ? wv : 0.f);
            st_v4(act_u32 + 8192 + row, wr[0], wr[1], wr[2], wr[3]);
            st_v4(act_u32 + 8192 + row + 128, wr[4], wr[5], wr[6], wr[7]);
          }
          // gF^T[channel][ray slot] for my channel(s) (k16 image at G^T + 0 / + 4096)
          const int tile_end = min((tile + 1) * TILE, npts);
          {
            uint32_t gw[8];
#pragma unroll
            for (int jx = 0; jx < 8; ++jx) {
              float g0 = 0.f, g1 = 0.f;
              if (a.g_feature_map) {
                if ((rl0 + 2 * jx) * N < tile_end) g0 = a.g_feature_map[((size_t)img * a.n_rays + r0 + rl0 + 2 * jx) * W + t + TILE * h];
                if ((rl0 + 2 * jx + 1) * N < tile_end) g1 = a.g_feature_map[((size_t)img * a.n_rays + r0 + rl0 + 2 * jx + 1) * W + t + TILE * h];
              }
              gw[jx] = pack_bf16x2(g0, g1);
            }
            st_v4(act_u32 + h * 4096 + row, gw[0], gw[1], gw[2], gw[3]);
            st_v4(act_u32 + h * 4096 + row + 128, gw[4], gw[5], gw[6], gw[7]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&misc->a_ready[s]);
        C3D_BPROF(e_setup);

        // ---- layers D .. 0: cotangent through sin/FiLM, column sums, G^T tile for the next MMA job
        for (int l = D; l >= 0; --l) {
          C3D_BPROF(e_other);
          // my channel's saved fp16 accumulator rows of this layer: 16 point groups of 16 bytes.  They do not depend on the
          // MMAs, so the first PF iterations' loads are in flight while the accumulator is still being produced.
          const float scale = film_img[l * W + ch].x, shift = film_img[l * W + ch].y;
          const uint4* pacc = reinterpret_cast<const uint4*>(a.save_acc + (((size_t)l * a.n_tiles_g + tile_g) * 16 * W + ch) * 8);
          uint4 ac[PF + 1][2];
#pragma unroll
          for (int i = 0; i < PF; ++i) {
            ac[i][0] = __ldcs(pacc + (size_t)(2 * i) * W);
            ac[i][1] = __ldcs(pacc + (size_t)(2 * i + 1) * W);
          }
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          C3D_BPROF(e_wait);
          jobcnt++;
          tc_fence_after();
          {
            // two points per operation (FFMA2 / FMUL2 / FADD2): the pass is bound by issue slots, not by a pipe
            const float2 scale2 = make_float2(scale, scale), shift2 = make_float2(shift, shift);
            float2 G1v = make_float2(0.f, 0.f), G2v = G1v;
            uint32_t v[16];
#pragma unroll
            for (int cp = 0; cp < 8; ++cp) {               // 16 points per iteration = 2 point groups
              tmem_ld_32x16(tcol + cp * 16, v);
              if (cp + PF < 8) {
                ac[(cp + PF) % (PF + 1)][0] = __ldcs(pacc + (size_t)(2 * (cp + PF)) * W);
                ac[(cp + PF) % (PF + 1)][1] = __ldcs(pacc + (size_t)(2 * (cp + PF) + 1) * W);
              }
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const uint4 aq = ac[cp % (PF + 1)][g];
                const uint32_t aw[4] = {aq.x, aq.y, aq.z, aq.w};
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 acc2 = __half22float2(*reinterpret_cast<const __half2*>(&aw[i]));     // saved fp16 accumulators
                  const float2 arg = fma2(acc2, scale2, shift2);
                  const float2 gh = make_float2(__uint_as_float(v[g * 8 + 2 * i]), __uint_as_float(v[g * 8 + 2 * i + 1]));
                  const float2 ga = mul2(gh, make_float2(__cosf(arg.x), __cosf(arg.y)));
                  G2v = add2(G2v, ga);
                  G1v = fma2(ga, acc2, G1v);
                  const float2 og = mul2(ga, scale2);
                  o[i] = pack_bf16x2(og.x, og.y);
                }
                const int unit = (cp & 3) * 2 + g;            // 16-byte unit inside the 64-point block
                st_v4(row_u32 + (uint32_t)(cp >> 2) * ACT_PBLOCK + (uint32_t)((unit ^ c7) << 4), o[0], o[1], o[2], o[3]);
              }
            }
            const float G1 = G1v.x + G1v.y, G2 = G2v.x + G2v.y;
            atomicAdd(gfilm_img + ((size_t)l * W + ch) * 2 + 0, G1);
            atomicAdd(gfilm_img + ((size_t)l * W + ch) * 2 + 1, G2);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(&misc->a_ready[s]);
          C3D_BPROF(e_epi);
          if (l == D) {
            // ---- d viewdirs of my point (job 1: heads rows 4..6 = Wview[:, 256..258])
            mbar_wait(&misc->acc_full[s], jobcnt & 1u);
            jobcnt++;
            tc_fence_after();
            uint32_t v4[4];
            if (point_role) { tmem_ld_32x4(tacc + 4, v4); tmem_ld_wait(); }
            tc_fence_before();
            if (point_role && valid && a.g_viewdirs) {
#pragma unroll
              for (int jx = 0; jx < 3; ++jx) atomicAdd(a.g_viewdirs + gray * 3 + jx, __uint_as_float(v4[jx]));
            }
            mbar_arrive(&misc->a_ready[s]);
          }
        }
        // ---- d points (job D+2: heads rows 0..2 = W0)
        {
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          jobcnt++;
          tc_fence_after();
          uint32_t v4[4];
          if (point_role) { tmem_ld_32x4(tacc, v4); tmem_ld_wait(); }
          tc_fence_before();
          if (point_role && valid) {
            float* gp = a.g_pts + gpt * 3;
#pragma unroll
            for (int jx = 0; jx < 3; ++jx) gp[jx] += nscale * __uint_as_float(v4[jx]);
          }
        }
#ifdef C3D_KERNEL_PROF
        if (eprof) ++e_tiles;
#endif
        C3D_BPROF(e_other);
      }
    }
#ifdef C3D_KERNEL_PROF
    if (eprof)
      printf("c3d prof bwd eg(slot %d): total %lld tiles %d  setup %lld  wait_acc_full %lld  layer epilogues %lld  other(narrow waits, d pts) %lld\n",
             s, clock64() - e_begin, e_tiles, e_setup, e_wait, e_epi, e_other);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// gdot[point] = sum_c g_feature_map[ray][c] * feat[c][point]; feat = sin(scale * acc + shift) of the view layer, recomputed
// from its saved fp16 accumulator tiles; one warp per (tile, point group of 8), lanes over channels.
__global__ void __launch_bounds__(256) gdot_kernel(const Args a, float* __restrict__ gdot) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + warp;
  const long long total = a.n_tiles_g * 16;
  if (wid >= total) return;
  const long long tile_g = wid >> 4;
  const int pg = (int)(wid & 15);
  const int u = (int)(tile_g / a.tiles_per_unit), tile = (int)(tile_g - (long long)u * a.tiles_per_unit);
  const int N = a.n_samples;
  const int img = u / a.units_per_img;
  const int r0 = (u - img * a.units_per_img) * a.unit_rays;
  const int nr = min(a.unit_rays, a.n_rays - r0);
  const int npts = nr * N;
  const int q0 = tile * TILE + pg * 8;
  if (q0 >= npts) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const uint4* f = reinterpret_cast<const uint4*>(a.save_acc + (((size_t)a.D * a.n_tiles_g + tile_g) * 16 + pg) * W * 8);
  const float2* film = a.film + ((size_t)img * (a.D + 1) + a.D) * W;                 // view layer's (scale, shift)
  auto feat2 = [&](uint32_t w, float2 fs) {                                          // two saved accumulators -> outputs
    const float2 acc = __half22float2(*reinterpret_cast<const __half2*>(&w));
    return make_float2(__sinf(fmaf(acc.x, fs.x, fs.y)), __sinf(fmaf(acc.y, fs.x, fs.y)));
  };
  const int ray_first = q0 / N, ray_last = min(q0 + 7, npts - 1) / N;
  if (ray_first == ray_last) {
    // the group's 8 points lie on one ray (always when N % 8 == 0): its cotangent row is read once per channel
    const float* grow = a.g_feature_map + ((size_t)img * a.n_rays + r0 + ray_first) * W;
    uint4 fv[W / 32];
    float gv[W / 32];
#pragma unroll
    for (int j = 0; j < W / 32; ++j) { fv[j] = __ldcs(f + lane + 32 * j); gv[j] = grow[lane + 32 * j]; }
#pragma unroll
    for (int j = 0; j < W / 32; ++j) {
      const uint32_t fw[4] = {fv[j].x, fv[j].y, fv[j].z, fv[j].w};
      const float2 fs = film[lane + 32 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = feat2(fw[i], fs);
        acc[2 * i] = fmaf(gv[j], x.x, acc[2 * i]);
        acc[2 * i + 1] = fmaf(gv[j], x.y, acc[2 * i + 1]);
      }
    }
  } else {
    for (int c = lane; c < W; c += 32) {
      const uint4 fv = __ldcs(f + c);
      const uint32_t fw[4] = {fv.x, fv.y, fv.z, fv.w};
      const float2 fs = film[c];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int q = min(q0 + i, npts - 1);
        const float g = a.g_feature_map[((size_t)img * a.n_rays + r0 + q / N) * W + c];
        const float2 x2 = feat2(fw[i >> 1], fs);
        acc[i] = fmaf(g, (i & 1) ? x2.y : x2.x, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (lane < 8 && q0 + lane < npts) {
    const int q = q0 + lane;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v = (lane == i) ? acc[i] : v;
    gdot[((size_t)img * a.n_rays + r0 + q / N) * N + (q % N)] = v;
  }
}

}}  // namespace c3d::fusedbwd
