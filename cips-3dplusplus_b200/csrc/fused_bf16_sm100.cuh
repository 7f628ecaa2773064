// Fused NeRF-branch forward for sm_100a (third structure tried this round; profiles/r01_fused.md has the history).
//
// Reference semantics: exp/cips3d/volume_renderer.py:133-160,192-283, exp/cips3d/nerf_utils.py:17-218,230-338.
//
// Orientation: every layer computes D^T[channel][point] = W * H^T, so TMEM lanes are channels (the FiLM scale /
// shift of an epilogue thread are two registers) and TMEM columns are the 128 points of the tile.
//
// Activation tile (64 KB per slot): H^T stored [channel][point] in bf16, two 64-point blocks of [256 rows][128 B],
// 16-byte units XOR-swizzled with (channel & 7).  One buffer, three operand views:
//   * B operand, MN-major SWIZZLE_128B (N = points, K = channels)      -> hidden / view layer MMAs
//   * A operand, MN-major SWIZZLE_128B (M = points, K = channels)      -> sdf / rgb head MMAs (N = 16)
//   * A operand, K-major  SWIZZLE_128B (M = channels, K = points)      -> compositing MMA  F^T = feat^T * Wgt^T
// The epilogue therefore writes 8 consecutive points of its channel with one 16-byte store.
//
// Per 128-point tile the MMA issuer runs D+3 jobs ("wait a_ready -> MMAs -> commit acc_full"), the slot's
// epilogue group answers each ("wait acc_full -> epilogue -> arrive a_ready"):
//   job 0       layer 0    : K = 16 split product, A = wk16 (W0 hi/lo), B = point tile (hi/mid/lo)
//   job 1..D-1  hidden l   : 2 halves x 16 x (128x128x16), A = weight ring stage, B = H^T
//   job D       sdf head   : 16 x (128x16x16), A = H^T (rows = points), B = heads16
//   job D+1     view layer : hidden-style + one K = 16 MMA per half adding W_view[:,256:259] * viewdir
//   job D+2     post       : compositing MMAs (features summed per ray on the tensor core) + rgb head MMA
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"
#include "fused_common.cuh"

namespace c3d { namespace fused {


// epilogue warps per slot (template parameter kEgw): 4 -> each thread owns channels t and t+128;
// 8 -> warps 0-3 own channel half 0, warps 4-7 half 1 (and only warps 0-3 run the per-point stages)
__host__ __device__ constexpr int nthreads(int egw) { return 128 + 2 * egw * 32; }
constexpr int ACT_PBLOCK = 32768;
constexpr int STAGE_BYTES = 128 * 128;         // 16384: one channel half of a K-chunk, [128 rows][64 k]
constexpr int NSTAGE = 4;
constexpr int RAYS = 16;                       // rays touching one tile (n_samples >= 8)
constexpr int AUX_BYTES = 4096;                // per slot: point tile / view tile ([128][16] k16) or Wgt ([16][128] sw128)

constexpr int SM_ACT = 0;
constexpr int SM_STAGE = SM_ACT + 2 * ACT_BYTES;                 // 131072
constexpr int SM_HEADS = SM_STAGE + NSTAGE * STAGE_BYTES;        // 196608
constexpr int SM_WK16 = SM_HEADS + (int)RGB16_BYTES;             // 204800
constexpr int SM_AUX = SM_WK16 + (int)W0IMG_BYTES;               // 212992  [slot][4096]
constexpr int SM_OM = SM_AUX + 2 * AUX_BYTES;                    // 221184  [slot][128] float
constexpr int SM_RAYACC = SM_OM + 2 * TILE * 4;                  // 222208  [slot][RAYS*2][8] float
constexpr int SM_PART = SM_RAYACC + 2 * RAYS * 2 * 8 * 4;        // 224256  [slot][RAYS][4 warps][8] float: per-warp ray sums
constexpr int SM_MISC = SM_PART + 2 * RAYS * 4 * 8 * 4;          // 228352
constexpr int SM_TOTAL = SM_MISC + 256;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;

struct Misc {
  uint64_t full[NSTAGE], empty[NSTAGE], a_ready[2], acc_full[2];
  uint32_t tmem_base;
  float carry[2];
};

__device__ __forceinline__ int job_layer(int j, int D) {
  if (j >= 1 && j <= D - 1) return j - 1;
  if (j == D + 1) return D - 1;
  return -1;
}

__device__ __forceinline__ void st_f16(uint32_t smem_addr, float x) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(x));
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(smem_addr), "h"((unsigned short)r) : "memory");
}
__device__ __forceinline__ void st_v4(uint32_t smem_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// sin(acc * scale + shift) for 16 consecutive points of one channel -> fp16 -> two 16-byte stores into H^T.
__device__ __forceinline__ void epilogue16(const uint32_t (&v)[16], float scale, float shift, uint32_t row_addr,
                                           int u0, int c7) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __sinf(fmaf(__uint_as_float(v[g * 8 + i]), scale, shift));
    st_v4(row_addr + (uint32_t)(((u0 + g) ^ c7) << 4), pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]),
          pack_f16x2(o[4], o[5]), pack_f16x2(o[6], o[7]));
  }
}

// Same, additionally keeping what the backward needs: the pre-FiLM accumulator as fp16 in global memory, layout
// [point group][channel][8 points] so a warp writes 512 contiguous bytes per group.  The backward recomputes
// cos(scale * acc + shift) (and, for the view layer, the output sin(...) itself) from it: |acc| < 2, so fp16's 11 significant
// bits put the ~30 rad argument within 0.007 rad -- the accuracy a stored bf16 cosine has -- at a third of the bytes of storing
// accumulator, cosine and output.
__device__ __forceinline__ void epilogue16_save(const uint32_t (&v)[16], float scale, float shift, uint32_t row_addr,
                                                int u0, int c7, __nv_bfloat16* sacc) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __sinf(fmaf(__uint_as_float(v[g * 8 + i]), scale, shift));
    const uint4 po = make_uint4(pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]), pack_f16x2(o[4], o[5]), pack_f16x2(o[6], o[7]));
    st_v4(row_addr + (uint32_t)(((u0 + g) ^ c7) << 4), po.x, po.y, po.z, po.w);
    const size_t goff = (size_t)g * (W * 8);
    *reinterpret_cast<uint4*>(sacc + goff) =
        make_uint4(pack_f16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1])),
                   pack_f16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3])),
                   pack_f16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5])),
                   pack_f16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])));
  }
}

template <int kCluster, int kEgw, bool kSave = false>
__global__ void __launch_bounds__(nthreads(kEgw), 1) fused_forward_kernel(const Args a) {
  constexpr int EG_THREADS = kEgw * 32;
  constexpr int NTHREADS = nthreads(kEgw);
  constexpr int HPT = kEgw == 8 ? 1 : 2;               // channel halves per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int D = a.D, N = a.n_samples;
  const int JOBS = a.sdf_only ? D + 1 : D + 3;         // density-only pass: jobs 0..D (layer 0, hidden layers, sdf head)
  const int nslots = 2 * gridDim.x;
  const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0u;

  if (threadIdx.x == 32) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&misc->full[i], 1); mbar_init(&misc->empty[i], kCluster); }
    for (int i = 0; i < 2; ++i) { mbar_init(&misc->a_ready[i], EG_THREADS); mbar_init(&misc->acc_full[i], 1); }
    misc->carry[0] = misc->carry[1] = 1.0f;
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(&misc->tmem_base, 512); tmem_relinquish(); }
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.blob + a.L.rgb16);      // heads16 + wk16 are adjacent
    uint4* dst = reinterpret_cast<uint4*>(smem + SM_HEADS);
    for (int i = threadIdx.x; i < (int)(RGB16_BYTES + W0IMG_BYTES) / 16; i += NTHREADS) dst[i] = src[i];
    float* ra = reinterpret_cast<float*>(smem + SM_RAYACC);
    for (int i = threadIdx.x; i < 2 * RAYS * 2 * 8; i += NTHREADS) ra[i] = 0.f;
    float* rp = reinterpret_cast<float*>(smem + SM_PART);
    for (int i = threadIdx.x; i < 2 * RAYS * 4 * 8; i += NTHREADS) rp[i] = 0.f;
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
    // ============================================================ weight producer
    const uint8_t* wsrc = a.blob + a.L.wf16h;
    uint32_t n = 0;
    for (int g = 0; g < rounds; ++g) {
      const int layer = job_layer(g % JOBS, D);
      if (layer < 0) continue;
      for (int s = 0; s < 2; ++s) {
        for (int c = 0; c < 2 * NCHUNK; ++c, ++n) {          // stage = (K-chunk c>>1, channel half c&1): 16 KB contiguous
          const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
          mbar_wait(&misc->empty[st], ph ^ 1u);
          uint8_t* dst = smem + SM_STAGE + st * STAGE_BYTES;
          const uint8_t* src = wsrc + (size_t)layer * WBF16_LAYER_BYTES + (size_t)c * STAGE_BYTES;
#ifdef C3D_KERNEL_PROF
          if (a.debug & 1) { mbar_arrive(&misc->full[st]); continue; }     // timing experiment: skip the weight copies
#endif
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
    const uint32_t idesc_kk = umma_idesc_bf16(128, 128, 0, 0);     // both K-major (K = 16 side products)
    // every product that reads the activation tile takes IEEE half operands (c3d_common.cuh: "16-bit operand formats")
    const uint32_t idesc_l = umma_idesc_f16(128, 128, 0, 1);       // layers: A = weights K-major, B = H^T MN-major
    const uint32_t idesc_h = umma_idesc_f16(128, 16, 1, 0);        // heads : A = H^T MN-major (rows = points)
    const uint32_t idesc_c = umma_idesc_f16(128, 16, 0, 0);        // composite: A = feat^T K-major, B = Wgt K-major
    const uint32_t act_addr[2] = {smem_u32(smem + SM_ACT), smem_u32(smem + SM_ACT + ACT_BYTES)};
    const uint32_t stage_base = smem_u32(smem + SM_STAGE);
    const uint32_t aux_addr[2] = {smem_u32(smem + SM_AUX), smem_u32(smem + SM_AUX + AUX_BYTES)};
    const uint32_t heads_addr = smem_u32(smem + SM_HEADS), wk_addr = smem_u32(smem + SM_WK16);
    uint32_t n = 0, jobcnt[2] = {0u, 0u};
#ifdef C3D_KERNEL_PROF
    const bool prof = (a.debug & 2) != 0;
#else
    constexpr bool prof = false;
#endif
    long long t_ready = 0, t_full = 0, t_issue = 0, t_mark = clock64(), t_begin = t_mark;
#define C3D_PROF_ADD(acc) do { if (prof) { const long long now_ = clock64(); acc += now_ - t_mark; t_mark = now_; } } while (0)
    for (int g = 0; g < rounds; ++g) {
      const int j = g % JOBS, tile_idx = g / JOBS;
      const int layer = job_layer(j, D);
      for (int s = 0; s < 2; ++s) {
        const bool real = tile_idx < my_tiles[s];
        const uint32_t tacc = tmem_base + (uint32_t)s * 256u;
        if (layer >= 0) {
          if (real) {
            C3D_PROF_ADD(t_issue);
            mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u);
            C3D_PROF_ADD(t_ready);
            tc_fence_after();
            if (j == D + 1) {             // view layer: W_view[:,256:259] * viewdir first (K = 16), then accumulate
              const uint64_t bd = umma_desc_kmajor_k16(aux_addr[s]);
#pragma unroll
              for (int h = 0; h < 2; ++h)
                umma_bf16_ss(tacc + (uint32_t)h * 128u, umma_desc_kmajor_k16(wk_addr + h * 4096), bd, idesc_kk, 0u);
            }
          }
          const uint32_t acc0 = (j == D + 1) ? 1u : 0u;
          for (int c = 0; c < 2 * NCHUNK; ++c, ++n) {
            const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
            C3D_PROF_ADD(t_issue);
            mbar_wait(&misc->full[st], ph);
            C3D_PROF_ADD(t_full);
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
          C3D_PROF_ADD(t_issue);
          mbar_wait(&misc->a_ready[s], jobcnt[s] & 1u);
          C3D_PROF_ADD(t_ready);
          tc_fence_after();
          if (j == 0) {                     // layer 0
            const uint64_t bd = umma_desc_kmajor_k16(aux_addr[s]);
#pragma unroll
            for (int h = 0; h < 2; ++h)
              umma_bf16_ss(tacc + (uint32_t)h * 128u, umma_desc_kmajor_k16(wk_addr + h * 4096), bd, idesc_kk, 0u);
          } else {
            // These narrow MMAs are paced by the issuing thread and by the accumulate dependency, not by the tensor
            // pipe: descriptors are base + immediate, the loops are fully unrolled, and consecutive K-steps go to
            // different partial accumulators (2 per channel half for the compositing, 4 for the heads) that the
            // epilogue sums.
            if (j == D + 2) {               // compositing: F^T[c][ray] = sum_p feat^T[c][p] * Wgt[ray][p]
              const uint64_t fa = umma_desc_kmajor_sw128(act_addr[s]), wb = umma_desc_kmajor_sw128(aux_addr[s]);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  umma_bf16_ss(tacc + (uint32_t)(h * 2 + (ks & 1)) * 16u,
                               fa + (uint64_t)(((ks >> 2) * ACT_PBLOCK + h * 16384 + (ks & 3) * 32) >> 4),
                               wb + (uint64_t)(((ks >> 2) * 2048 + (ks & 3) * 32) >> 4), idesc_c, ks >= 2);
            }
            const uint32_t dcol = (j == D + 2) ? 64u : 0u;    // heads: sdf (job D) / rgb (job D+2)
            const uint64_t ha = umma_desc_mnmajor_sw128(act_addr[s], ACT_PBLOCK), hb = umma_desc_kmajor_sw128(heads_addr);
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
              umma_bf16_ss(tacc + dcol + (uint32_t)(ks & 3) * 16u, ha + (uint64_t)((ks * 2048) >> 4),
                           hb + (uint64_t)(((ks >> 2) * 2048 + (ks & 3) * 32) >> 4), idesc_h, ks >= 4);
          }
          umma_commit(&misc->acc_full[s]);
          jobcnt[s]++;
        }
      }
    }
    if (prof && (blockIdx.x % 21 == 0 || blockIdx.x == 1)) {
      C3D_PROF_ADD(t_issue);
      printf("c3d prof mma[blk %d]: total %lld  wait_a_ready %lld  wait_full %lld  issue %lld  rounds %d\n", (int)blockIdx.x, clock64() - t_begin,
             t_ready, t_full, t_issue, rounds);
    }
   }
  } else if (warp >= 4) {
    // ============================================================ epilogue groups
    const int s = (warp - 4) / kEgw;
    const int te = threadIdx.x - 128 - s * EG_THREADS;   // 0..EG_THREADS-1
    const int t = te & 127;                              // point row of the point stages; channel t (+128) of the layer stages
    const bool pt_role = te < TILE;                      // the per-point stages run on the first four warps of the slot
    const int quad = warp & 3;
    const uint32_t bar_id = 1u + (uint32_t)s;
    const int slot = 2 * blockIdx.x + s;
    uint8_t* aux = smem + SM_AUX + s * AUX_BYTES;
    const uint32_t aux_u32 = smem_u32(aux);
    float* omS = reinterpret_cast<float*>(smem + SM_OM) + s * TILE;
    float* rayacc = reinterpret_cast<float*>(smem + SM_RAYACC) + s * RAYS * 2 * 8;
    float* raypart = reinterpret_cast<float*>(smem + SM_PART) + s * RAYS * 4 * 8;
    const uint32_t tacc = tmem_base + (uint32_t)s * 256u + ((uint32_t)(quad * 32) << 16);
    const float* scal = reinterpret_cast<const float*>(a.blob + a.L.scal);
    const float bsig = scal[0], brgb0 = scal[1], brgb1 = scal[2], brgb2 = scal[3];
    const float inv_beta = 1.0f / scal[4];
    const uint32_t act_u32 = smem_u32(smem + SM_ACT + s * ACT_BYTES);
    const int c7 = t & 7;
    float carry_f[HPT];                                  // partial feature sums of a ray continuing into the next tile
#pragma unroll
    for (int hh = 0; hh < HPT; ++hh) carry_f[hh] = 0.f;
#ifdef C3D_KERNEL_PROF
    const bool eprof = (a.debug & 2) != 0 && blockIdx.x == 0 && te == 0;
#else
    constexpr bool eprof = false;
#endif
    long long e_wait = 0, e_epi = 0, e_pt = 0, e_mark = clock64(), e_begin = e_mark;
#define C3D_EPROF(acc) do { if (eprof) { const long long now_ = clock64(); acc += now_ - e_mark; e_mark = now_; } } while (0)
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
#pragma unroll
      for (int hh = 0; hh < HPT; ++hh) carry_f[hh] = 0.f;

      for (int tile = 0; tile < ntiles; ++tile) {
        // ------------------------------------------------ geometry of my point (nerf_utils.py:17-170)
        const long long tile_g = (long long)u * a.tiles_per_unit + tile;   // index of this tile in the saved tensors
        const int q = tile * TILE + t;
        const bool valid = q < npts;
        const int qc = valid ? q : npts - 1;
        const int rl = qc / N, k = qc - rl * N;
        const int rl0 = (tile * TILE) / N;                // first ray touching this tile
        const size_t gray = (size_t)img * a.n_rays + r0 + rl;
        float px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, dist = 0.f, zk = 0.f;
        if (!pt_role) {
        } else if (a.input_kind == C3D_INPUT_POSES) {
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
        if (pt_role && a.z_vals_out && valid) a.z_vals_out[gray * N + k] = zk;
        const uint32_t aux_row = aux_u32 + (uint32_t)((t >> 3) * 256 + (t & 7) * 16);
        if (pt_role) {
          // point tile of the layer-0 MMA: per coordinate (hi, mid, hi, lo); k-slots 12..15 zero
          const float pn[3] = {px * nscale, py * nscale, pz * nscale};
          float e[12];
#pragma unroll
          for (int jx = 0; jx < 3; ++jx) {
            const float hi = __bfloat162float(__float2bfloat16_rn(pn[jx]));
            const float r1 = pn[jx] - hi;
            const float mid = __bfloat162float(__float2bfloat16_rn(r1));
            const float lo = __bfloat162float(__float2bfloat16_rn(r1 - mid));
            e[4 * jx + 0] = hi; e[4 * jx + 1] = mid; e[4 * jx + 2] = hi; e[4 * jx + 3] = lo;
          }
          st_v4(aux_row, pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
          st_v4(aux_row + 128, pack_bf16x2(e[8], e[9]), pack_bf16x2(e[10], e[11]), 0u, 0u);
        }
        if (pt_role) fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&misc->a_ready[s]);

        float sdf = 0.f, wgt = 0.f;
        // ------------------------------------------------ layers 0..D (D = view layer): thread = channel t and t+128
        for (int l = 0; l <= D; ++l) {
          const float2 fa = film_img[l * W + t + (kEgw == 8 ? TILE * (te >> 7) : 0)];
          const float2 fb = kEgw == 8 ? fa : film_img[l * W + t + TILE];
          if (l == D) {
            // ---------------------------------------------- sdf head (thread = point) -> alpha -> transmittance
            mbar_wait(&misc->acc_full[s], jobcnt & 1u);
            jobcnt++;
            tc_fence_after();
            if (pt_role) {
              {
                uint32_t v4[4][4];                   // heads16 rows 4, 5 (hi / lo of sigma_linear.weight), 4 partial sums
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) tmem_ld_32x4(tacc + pp * 16 + 4, v4[pp]);
                tmem_ld_wait();
                tc_fence_before();
                sdf = bsig;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) sdf += __uint_as_float(v4[pp][0]) + __uint_as_float(v4[pp][1]);
              }
              if (valid) a.sdf[gray * N + k] = sdf;
              if (!a.sdf_only) {
              const float sigma = sigmoid_precise(-sdf * inv_beta) * inv_beta;
              const float alpha = 1.0f - expf(-sigma * dist);
              const float om = 1.0f - alpha + 1e-10f;
              omS[t] = valid ? om : 1.0f;
              // view-direction tile for the view-layer MMA (k-slots 12..14), reuses the point-tile buffer
              st_v4(aux_row, 0u, 0u, 0u, 0u);
              st_v4(aux_row + 128, 0u, 0u, pack_bf16x2(vx, vy), pack_bf16x2(vz, 0.f));
              named_bar_sync(bar_id, TILE);
              const int first_row = t - k;
              float T = first_row < 0 ? misc->carry[s] : 1.0f;
              for (int m = max(first_row, 0); m < t; ++m) T *= omS[m];
              wgt = valid ? alpha * T : 0.f;
              named_bar_sync(bar_id, TILE);
              if (t == TILE - 1) misc->carry[s] = (k == N - 1) ? 1.0f : T * om;
              fence_proxy_async_smem();
              }
            }
            if (a.sdf_only) break;                   // density-only pass: the tile ends with the sdf head
            mbar_arrive(&misc->a_ready[s]);
          }
          C3D_EPROF(e_pt);
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          C3D_EPROF(e_wait);
          jobcnt++;
          tc_fence_after();
          if (l == D && pt_role) {
            // the view-direction tile has been consumed: build Wgt[ray slot][point] (fp16, K-major SW128) in its place
            const int myslot = rl - rl0;
#pragma unroll
            for (int jx = 0; jx < RAYS; ++jx)
              st_f16(aux_u32 + (uint32_t)((t >> 6) * 2048) + sw128_offset(jx, t & 63), jx == myslot ? wgt : 0.f);
          }
#pragma unroll 1
          for (int hh = 0; hh < HPT; ++hh) {
            // my channel (lane `t` of accumulator half h), 128 points in 8 chunks of 16, TMEM loads double-buffered
            const int h = kEgw == 8 ? (te >> 7) : hh;
            const uint32_t tcol = tacc + (uint32_t)h * 128u;
            const uint32_t row_u32 = act_u32 + (uint32_t)(t + TILE * h) * 128u;
            uint32_t v0[16], v1[16];
            tmem_ld_32x16(tcol, v0);
#pragma unroll 2
            for (int cp = 0; cp < 4; ++cp) {               // 32 points per iteration
              const uint32_t row = row_u32 + (uint32_t)(cp >> 1) * ACT_PBLOCK;
              tmem_ld_wait();
              tmem_ld_32x16(tcol + cp * 32 + 16, v1);
              if (kSave) {
                // element offset of (layer l, this tile, point group 4*cp, my channel)
                const size_t so = ((((size_t)l * a.n_tiles_g + tile_g) * 16 + cp * 4) * W + (t + TILE * h)) * 8;
                epilogue16_save(v0, hh ? fb.x : fa.x, hh ? fb.y : fa.y, row, (cp & 1) * 4, c7, a.save_acc + so);
                tmem_ld_wait();
                if (cp < 3) tmem_ld_32x16(tcol + (cp + 1) * 32, v0);
                epilogue16_save(v1, hh ? fb.x : fa.x, hh ? fb.y : fa.y, row, (cp & 1) * 4 + 2, c7,
                                a.save_acc + so + 2 * W * 8);
              } else {
                epilogue16(v0, hh ? fb.x : fa.x, hh ? fb.y : fa.y, row, (cp & 1) * 4, c7);
                tmem_ld_wait();
                if (cp < 3) tmem_ld_32x16(tcol + (cp + 1) * 32, v0);
                epilogue16(v1, hh ? fb.x : fa.x, hh ? fb.y : fa.y, row, (cp & 1) * 4 + 2, c7);
              }
            }
          }
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(&misc->a_ready[s]);
          C3D_EPROF(e_epi);
        }

        // ------------------------------------------------ post job: composited features (thread = channel) + rgb (thread = point)
        if (!a.sdf_only) {
          mbar_wait(&misc->acc_full[s], jobcnt & 1u);
          jobcnt++;
          tc_fence_after();
          float rgbv[3] = {brgb0, brgb1, brgb2};         // raw rgb of my point (point-role threads): 4 partial sums
          {
            uint32_t v4[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) tmem_ld_32x4(tacc + 64 + pp * 16, v4[pp]);
            tmem_ld_wait();
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              rgbv[0] += __uint_as_float(v4[pp][0]); rgbv[1] += __uint_as_float(v4[pp][1]); rgbv[2] += __uint_as_float(v4[pp][2]);
            }
          }
          const int tile_end = min((tile + 1) * TILE, npts);      // first point index beyond this tile
#pragma unroll
          for (int hh = 0; hh < HPT; ++hh) {
            const int h = kEgw == 8 ? (te >> 7) : hh;
            uint32_t fv[16], fw[16];                     // composited features of my channel; column = ray slot; 2 partial sums
            tmem_ld_32x16(tacc + (uint32_t)(h * 2) * 16u, fv);
            tmem_ld_32x16(tacc + (uint32_t)(h * 2 + 1) * 16u, fw);
            tmem_ld_wait();
#pragma unroll
            for (int jx = 0; jx < 16; ++jx) fv[jx] = __float_as_uint(__uint_as_float(fv[jx]) + __uint_as_float(fw[jx]));
            // thread = channel t + 128 h, fv[jx] = ray slot jx.  (b, hw, 256): a warp writes 32 consecutive channels of a
            // ray; (b, 256, hw): a thread writes up to 16 consecutive rays of its channel
            const int ch = t + TILE * h;
            float* fbase = a.feat_nchw ? a.feature_map + ((size_t)img * W + ch) * a.n_rays + r0 + rl0
                                       : a.feature_map + ((size_t)img * a.n_rays + r0 + rl0) * W + ch;
            __nv_bfloat16* hbase = reinterpret_cast<__nv_bfloat16*>(a.feature_map) + ((size_t)img * W + ch) * a.n_rays + r0 + rl0;
            const size_t fstride = a.feat_nchw ? 1 : W;
            const bool fbf16 = a.feat_nchw == 2;          // (b, 256, hw) bf16: the decoder hand-off at half the bytes
            // ray slot jx covers the unit's points [(rl0 + jx) N, (rl0 + jx + 1) N): the first jdone slots end inside this tile
            // (complete rays: stored), slot jdone -- if it has points here -- continues in the next tile (carried).  A few
            // predicated instructions per slot (the first version spent ~460 instructions per warp and tile here).
            const int jdone = tile_end / N - rl0;
            const bool partial = (rl0 + jdone) * N < tile_end;
            fv[0] = __float_as_uint(__uint_as_float(fv[0]) + carry_f[hh]);
            if (fbf16) {
#pragma unroll
              for (int jx = 0; jx < RAYS; ++jx)
                if (jx < jdone) hbase[jx] = __float2bfloat16_rn(__uint_as_float(fv[jx]));
            } else {
#pragma unroll
              for (int jx = 0; jx < RAYS; ++jx)
                if (jx < jdone) fbase[(size_t)jx * fstride] = __uint_as_float(fv[jx]);
            }
            float cnext = 0.f;
#pragma unroll
            for (int jx = 0; jx < RAYS; ++jx)
              if (jx == jdone) cnext = __uint_as_float(fv[jx]);
            carry_f[hh] = partial ? cnext : 0.f;
          }
          tmem_ld_wait();
          tc_fence_before();
          if (pt_role) {
          // rgb / xyz / mask sums of my ray (nerf_utils.py:315,329-336)
          if (kSave && valid) {
            float* ro = a.rgb_pt + (gray * N + k) * 3;
            ro[0] = rgbv[0]; ro[1] = rgbv[1]; ro[2] = rgbv[2];
            a.w_pt[gray * N + k] = wgt;
          }
          float vals[6];
          vals[0] = wgt * sigmoid_precise(rgbv[0]);
          vals[1] = wgt * sigmoid_precise(rgbv[1]);
          vals[2] = wgt * sigmoid_precise(rgbv[2]);
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
          // The head lane of every (ray, warp) segment parks its partial sums in part[ray slot][warp]; after the barrier the
          // thread of the ray's last point in this tile adds the (at most four) partials in warp order to the ray's running
          // sums -- a fixed order, so the maps are bit-reproducible (shared-memory atomics were not for N > 32).
          const int rprev = __shfl_up_sync(0xffffffffu, rl, 1);
          float* racc = rayacc + (rl & (2 * RAYS - 1)) * 8;
          float* part = raypart + (rl - rl0) * 32;
          if (valid && (lane == 0 || rprev != rl)) {
#pragma unroll
            for (int jx = 0; jx < 6; ++jx) part[quad * 8 + jx] = vals[jx];
          }
          named_bar_sync(bar_id, TILE);
          if (valid && (k == N - 1 || q == tile_end - 1)) {
#pragma unroll
            for (int jx = 0; jx < 6; ++jx) {
              float acc = racc[jx];
#pragma unroll
              for (int w4 = 0; w4 < 4; ++w4) { acc += part[w4 * 8 + jx]; part[w4 * 8 + jx] = 0.f; }
              racc[jx] = acc;
            }
          }
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
        }
      }  // tiles
    }    // units
    if (eprof) {
      C3D_EPROF(e_pt);
      printf("c3d prof eg(slot %d): total %lld  wait_acc_full(layers) %lld  epilogue %lld  other(point stages, sdf/post waits) %lld\n",
             s, clock64() - e_begin, e_wait, e_epi, e_pt);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}}  // namespace c3d::fused
