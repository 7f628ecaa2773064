// Fused NeRF-branch forward for sm_100a, CTA-pair kernel (the default bf16 forward since round 2).
//
// Reference semantics: exp/cips3d/volume_renderer.py:133-160,192-283, exp/cips3d/nerf_utils.py:17-218,230-338.
//
// Structure (profiles/r02_fused.md and profiles/r02_micro.md have the measurements behind each choice)
//   * FiLM is folded into the GEMM.  film_weights_kernel (kernels_aux.cuh) writes, per image, fp16(gamma_c * W_l[c][k]) in
//     the stage layout plus a K = 16 side image holding the shift (gamma b + beta, hi/lo split, multiplied by two "ones"
//     slots of the point tile), the layer-0 weights and the view-direction columns.  The accumulator is the SIREN argument
//     itself, so the epilogue is sin -> fp16 -> store with no per-channel constants.
//   * Orientation D[point][channel] = H * W'^T: TMEM lanes are points.  A thread owns one point of its tile in every
//     stage (geometry, layer epilogues, sdf / transmittance, rgb) and writes its own rows of the K-major activation tile.
//   * Two CTAs of a cluster form a pair and issue tcgen05.mma.cta_group::2 (M = 256 points = 128 per CTA, N = 256): every
//     CTA stages only its 128 weight rows and never exchanges activations.  Per 128 points and layer an SM then moves
//     224 KB through shared memory (the single-CTA kernel: 448 KB -- 3500 cycles at 128 B/clk against 2048 tensor cycles,
//     which is what bounds it), and a 64 KB ring holds a WHOLE layer, so slot 1 re-reads the stages slot 0 consumed: one
//     stream of the layer per two tiles, the refill has a full job of slack.
//   * Per SM a 128-point tile-layer costs 2048 tensor cycles and 32768 MUFU.SIN = 2080 SFU cycles (sm_probe M1): both
//     pipes are needed all the time; with two accumulator slots (all 512 TMEM columns) the dependency chain
//     MMA -> sines -> MMA of a slot leaves each of them about half busy.  Two variants that shorten the chain were built
//     and measured slower (profiles/r02_fused.md): N = 128 half-layer jobs (the activation tile is then read twice: shared
//     memory binds) and a mid-epilogue hand-off with an event-driven issuer (mbarrier polling latency).
//   * The sdf head runs on the FP32 pipe inside the last hidden layer's epilogue (fp32 sines, 128 FFMA per thread), and
//     density -> alpha -> transmittance run under the view layer's MMAs: no separate job on the chain.
//
// Roles per CTA (640 threads): warp 0 weight producer (bulk copies of its halves), warp 1 MMA issuer (leader CTA; the
// whole warp runs the control flow, one elected lane executes the tcgen05 instructions so every operand stays in uniform
// registers -- with the loop inside `if (elect_one())` ptxas wrapped each UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
// waterfall) or barrier relay (peer CTA: forwards "my stage landed" to the leader), warp 2 TMEM allocator, warps 4-11 /
// 12-19 epilogue group of slot 0 / 1 (8 warps: two per sub-partition, what the SFU needs to stay saturated).  Barriers the
// leader waits on (full, kfull, a_ready) collect arrivals from both CTAs; barriers signalled by tcgen05.commit (empty,
// kempty, acc_full) are multicast to both.
//
// Per pair-tile (128 points per CTA) the issuer runs D+2 jobs:
//   job 0       layer 0    : K = 16 product of the point tile (hi, mid, hi, lo per coordinate, ones) with the K16 image
//   job 1..D-1  hidden l   : K16 product (shift) + 16 x (256x256x16), A = activation tile, B = weight ring
//   job D       view layer : K16 product (view direction, shift) + 16 x (256x256x16)
//   job D+1     post       : compositing MMAs (A = feat^T as MN-major view, B = Wgt; N = 32, each CTA reads its own 16
//                            columns) + rgb head
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"
#include "fused_common.cuh"

namespace c3d { namespace pairk {

using namespace c3d::ptx;
using fused::Args;
using fused::TILE;
using fused::ACT_BYTES;
using fused::ACT_CHUNK;

constexpr int EGW = 8;                         // epilogue warps per slot: warp (quad, grp) owns lanes 32 quad.. and columns 64 grp..
                                               // of each accumulator half; grp 0 threads also run the per-point stages
constexpr int NTHREADS = 128 + 2 * EGW * 32;   // 640
// 4 stages of 16 KB ([128 rows][64 k] of a K-chunk) = 64 KB = one layer of this CTA's weight rows (the producer and the relay pay
// ~200 cycles per stage: 8 KB stages starved the issuer).
constexpr int RING_BYTES = 65536;
constexpr int STAGE_BYTES = 16384;
constexpr int NSTAGE = 4;
constexpr int RAYS = 16;                       // rays touching one tile (n_samples >= 8)
constexpr int AUX_BYTES = 4096;                // per slot: point / view tile ([128][16] k16) or Wgt ([16][128] sw128)
constexpr int K16_BYTES = 4096;                // per slot: this CTA's part of the layer's K16 image ([half 2][64 rows][16])
constexpr int HEADS_BYTES = 4096;              // 4 K-chunks x [8 rows][64 k]: this CTA's half of heads16
constexpr size_t WIMG_LAYER_BYTES = (size_t)W * W * 2;   // per image and K=256 layer: [half 2][kc 4][rank 2][8 KB]
constexpr size_t KIMG_LAYER_BYTES = (size_t)W * 16 * 2;  // per image and layer 0..D: [rank 2][half 2][2 KB]

constexpr int SM_ACT = 0;
constexpr int SM_STAGE = SM_ACT + 2 * ACT_BYTES;                 // 131072
constexpr int SM_K16 = SM_STAGE + RING_BYTES;                    // 196608
constexpr int SM_HEADS = SM_K16 + 2 * K16_BYTES;                 // 204800
constexpr int SM_AUX = SM_HEADS + HEADS_BYTES;                   // 208896
constexpr int SM_OM = SM_AUX + 2 * AUX_BYTES;                    // 217088  [slot][128] float
constexpr int SM_RAYACC = SM_OM + 2 * TILE * 4;                  // 218112  [slot][RAYS*2][8] float
constexpr int SM_WSIG = SM_RAYACC + 2 * RAYS * 2 * 8 * 4;        // 220160  sigma_linear.weight, fp32 [256]
constexpr int SM_SDFP = SM_WSIG + W * 4;                         // 221184  [slot][128] float: sdf partial sums of column group 1
constexpr int SM_PART = SM_SDFP + 2 * TILE * 4;                  // 222208  [slot][RAYS][4 warps][8] float: per-warp ray sums
constexpr int SM_MISC = SM_PART + 2 * RAYS * 4 * 8 * 4;          // 226304
constexpr int SM_TOTAL = SM_MISC + 256;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;

struct Misc {
  uint64_t full[NSTAGE], empty[NSTAGE], kfull[2], kempty[2], a_ready[2];
  uint64_t acc_full[2];      // [slot]: the job's MMAs are complete
  uint32_t tmem_base;
  float carry[2];
};

// ---- static schedule: pair-slot ps handles pair-units ps, ps + n_pairslots, ...; a pair-unit is 2 * unit_rays rays of one
// image (the leader takes the first unit_rays); Args::units_per_img counts pair-units.
__device__ __forceinline__ int pu_tiles(const Args& a, int pu) {
  const int r0 = (pu % a.units_per_img) * 2 * a.unit_rays;
  const int nr = min(a.unit_rays, a.n_rays - r0);            // the leader's share is never smaller than the peer's
  return (nr * a.n_samples + TILE - 1) / TILE;
}
struct Cursor {
  int pu, tile, ntiles, total, step;
  __device__ __forceinline__ void init(const Args& a, int ps, int nps) {
    total = a.batch * a.units_per_img; step = nps; pu = ps; tile = 0;
#ifdef C3D_KERNEL_PROF
    if ((a.debug & 16) && (ps & 1)) pu = total;      // experiment: odd pair-slots idle (one slot per pair runs alone)
#endif
    ntiles = pu < total ? pu_tiles(a, pu) : 0;
  }
  __device__ __forceinline__ bool valid() const { return pu < total; }
  __device__ __forceinline__ void next_tile(const Args& a) {
    if (++tile >= ntiles) { pu += step; tile = 0; ntiles = pu < total ? pu_tiles(a, pu) : 0; }
  }
};
// Both slots of a pair run the same job index; when their pair-units belong to the same image they also use the same
// weights, so slot 1 re-reads the ring stages slot 0 consumed instead of streaming the layer a second time.
__device__ __forceinline__ bool same_weights(const Args& a, const Cursor& c0, const Cursor& c1) {
  return c0.valid() && c1.valid() && c0.pu / a.units_per_img == c1.pu / a.units_per_img;
}
// kind of job j: 0 = layer 0 (K16 only), 1 = K = 256 layer (hidden 1..D-1, view layer D), 3 = post
__device__ __forceinline__ int job_kind(int j, int D) { return j == 0 ? 0 : (j == D + 1 ? 3 : 1); }
__device__ __forceinline__ int job_film_layer(int j, int D) { return j; }        // jobs 0..D: index into kimg
__device__ __forceinline__ int job_w_layer(int j, int D) { return j - 1; }       // jobs 1..D: index into wimg

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
// descriptor of the same layout `bytes` further (start-address field, 16-byte units; no carry out of the field here)
__device__ __forceinline__ uint64_t desc_at(uint64_t base, uint32_t bytes) { return base + (uint64_t)(bytes >> 4); }

// sin of 16 consecutive channels of one point -> fp16 -> two 16-byte stores into the point's row (units u0, u0 + 1)
__device__ __forceinline__ void epilogue16(const uint32_t (&v)[16], uint32_t row_addr, int u0, int r7) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __sinf(__uint_as_float(v[g * 8 + i]));
    st_v4(row_addr + (uint32_t)(((u0 + g) ^ r7) << 4), pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]),
          pack_f16x2(o[4], o[5]), pack_f16x2(o[6], o[7]));
  }
}

// In-kernel cycle counters of the issuing thread and of one epilogue thread (development builds only: -DC3D_KERNEL_PROF).
#ifdef C3D_KERNEL_PROF
#define C3D_PROF_DECL(n) long long prof_t[n] = {}; long long prof_mark = clock64(); const long long prof_begin = prof_mark
#define C3D_PROF(i) do { const long long now_ = clock64(); prof_t[i] += now_ - prof_mark; prof_mark = now_; } while (0)
#else
#define C3D_PROF_DECL(n) do { } while (0)
#define C3D_PROF(i) do { } while (0)
#endif

// the same for the last hidden layer: also the sdf head  sum_c wsig[c] * sin(...)  over these 16 channels, from the fp32 sines
// (wsig: shared-memory address of the 16 weights; warp-uniform, so the loads are broadcasts)
__device__ __forceinline__ void epilogue16_sdf(const uint32_t (&v)[16], uint32_t row_addr, int u0, int r7, const float4* wsig,
                                               float (&acc)[4]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __sinf(__uint_as_float(v[g * 8 + i]));
    const float4 wa = wsig[2 * g], wb = wsig[2 * g + 1];
    acc[0] = fmaf(wa.x, o[0], acc[0]); acc[1] = fmaf(wa.y, o[1], acc[1]); acc[2] = fmaf(wa.z, o[2], acc[2]); acc[3] = fmaf(wa.w, o[3], acc[3]);
    acc[0] = fmaf(wb.x, o[4], acc[0]); acc[1] = fmaf(wb.y, o[5], acc[1]); acc[2] = fmaf(wb.z, o[6], acc[2]); acc[3] = fmaf(wb.w, o[7], acc[3]);
    st_v4(row_addr + (uint32_t)(((u0 + g) ^ r7) << 4), pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]),
          pack_f16x2(o[4], o[5]), pack_f16x2(o[6], o[7]));
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) fused_forward_pair_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int D = a.D, N = a.n_samples;
  const int JOBS = D + 2;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairslots = (int)gridDim.x;          // 2 per cluster
  const int ps0 = (int)(blockIdx.x >> 1) * 2;      // pair-slots of this cluster: ps0, ps0 + 1

  if (threadIdx.x == 32) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&misc->full[i], leader ? 2 : 1); mbar_init(&misc->empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&misc->kfull[i], leader ? 2 : 1);
      mbar_init(&misc->kempty[i], 1);
      mbar_init(&misc->a_ready[i], 2 * EGW);       // one arrival per epilogue warp of both CTAs (leader's copy is used)
      mbar_init(&misc->acc_full[i], 1);
    }
    misc->carry[0] = misc->carry[1] = 1.0f;
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc_pair(&misc->tmem_base, 512); tmem_relinquish_pair(); }
  {
    // this CTA's half of heads16: rows rank*8 .. rank*8+7 of every K-chunk (one 1 KB swizzle atom each)
    const uint8_t* src = a.blob + a.L.rgb16;
    uint4* dst = reinterpret_cast<uint4*>(smem + SM_HEADS);
    for (int i = threadIdx.x; i < HEADS_BYTES / 16; i += NTHREADS) {
      const int kc = i >> 6, u = i & 63;
      dst[i] = reinterpret_cast<const uint4*>(src + kc * 2048 + rank * 1024)[u];
    }
    float* ra = reinterpret_cast<float*>(smem + SM_RAYACC);
    for (int i = threadIdx.x; i < 2 * RAYS * 2 * 8; i += NTHREADS) ra[i] = 0.f;
    float* rp = reinterpret_cast<float*>(smem + SM_PART);
    for (int i = threadIdx.x; i < 2 * RAYS * 4 * 8; i += NTHREADS) rp[i] = 0.f;
    float* ws_ = reinterpret_cast<float*>(smem + SM_WSIG);
    for (int i = threadIdx.x; i < W; i += NTHREADS) ws_[i] = reinterpret_cast<const float*>(a.blob + a.L.wsig)[i];
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, misc->tmem_base, 0);     // provably warp-uniform

  if (warp == 0) {
   if (elect_one()) {
    // ============================================================ weight producer (this CTA's parts)
    Cursor c[2];
    c[0].init(a, ps0, n_pairslots); c[1].init(a, ps0 + 1, n_pairslots);
    int jb[2] = {0, 0};
    uint32_t n = 0, kcnt[2] = {0u, 0u};
    while (c[0].valid() || c[1].valid()) {
      const bool shared = same_weights(a, c[0], c[1]);
      for (int s = 0; s < 2; ++s) {
        if (!c[s].valid()) continue;
        const int j = jb[s], kind = job_kind(j, D);
        const size_t img = (size_t)(c[s].pu / a.units_per_img);
        if (kind <= 1) {
          mbar_wait(&misc->kempty[s], (kcnt[s] & 1u) ^ 1u);
          kcnt[s]++;
          mbar_arrive_expect_tx(&misc->kfull[s], K16_BYTES);
          bulk_g2s(smem + SM_K16 + s * K16_BYTES,
                   a.kimg + (img * (D + 1) + job_film_layer(j, D)) * KIMG_LAYER_BYTES + rank * K16_BYTES, K16_BYTES, &misc->kfull[s]);
          if (kind == 1 && !(shared && s == 1)) {
            const uint8_t* wl = a.wimg + (img * D + job_w_layer(j, D)) * WIMG_LAYER_BYTES + rank * STAGE_BYTES;
#pragma unroll 1
            for (int pc = 0; pc < NSTAGE; ++pc, ++n) {             // piece pc = half * 2 + K-chunk pair / K-chunk, in the order the issuer consumes
              const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
              mbar_wait(&misc->empty[st], ph ^ 1u);
              mbar_arrive_expect_tx(&misc->full[st], STAGE_BYTES);
              bulk_g2s(smem + SM_STAGE + st * STAGE_BYTES, wl + (size_t)pc * 2 * STAGE_BYTES, STAGE_BYTES, &misc->full[st]);
            }
          }
        }
        if (++jb[s] == JOBS) { jb[s] = 0; c[s].next_tile(a); }
      }
    }
   }
  } else if (warp == 1 && !leader) {
   if (elect_one()) {
    // ============================================================ relay: "my part landed" -> leader's full / kfull
    Cursor c[2];
    c[0].init(a, ps0, n_pairslots); c[1].init(a, ps0 + 1, n_pairslots);
    int jb[2] = {0, 0};
    uint32_t n = 0, kcnt[2] = {0u, 0u};
    const uint32_t r_full = mapa_u32(smem_u32(&misc->full[0]), 0u), r_kfull = mapa_u32(smem_u32(&misc->kfull[0]), 0u);
    while (c[0].valid() || c[1].valid()) {
      const bool shared = same_weights(a, c[0], c[1]);
      for (int s = 0; s < 2; ++s) {
        if (!c[s].valid()) continue;
        const int kind = job_kind(jb[s], D);
        if (kind <= 1) {
          mbar_wait(&misc->kfull[s], kcnt[s] & 1u);
          kcnt[s]++;
          mbar_arrive_remote(r_kfull + (uint32_t)s * 8u);
          if (kind == 1 && !(shared && s == 1)) {
#pragma unroll 1
            for (int pc = 0; pc < NSTAGE; ++pc, ++n) {
              const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
              mbar_wait(&misc->full[st], ph);
              mbar_arrive_remote(r_full + st * 8u);
            }
          }
        }
        if (++jb[s] == JOBS) { jb[s] = 0; c[s].next_tile(a); }
      }
    }
   }
  } else if (warp == 1) {
   {
    // ============================================================ MMA issuer (leader CTA, issues for the pair)
    // The whole warp runs the (warp-uniform) control flow and one elected lane executes the tcgen05 instructions: every
    // operand is then uniform by data flow and lives in uniform registers.  (With the loop inside `if (elect_one())`
    // ptxas wrapped each UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall -- ~170 cycles per MMA.)
    const bool issue = elect_one();
    // every product that reads the activation tile takes IEEE half operands (c3d_common.cuh: "16-bit operand formats");
    // the K16 side products (point / view tile x kimg) stay bf16 x bf16 and add into the same fp32 accumulator
    const uint32_t idesc_k = umma_idesc_bf16(256, 256, 0, 0);      // K16 side products: both K-major
    const uint32_t idesc_l = umma_idesc_f16(256, 256, 0, 0);       // layers: both K-major
    const uint32_t idesc_h = umma_idesc_f16(256, 16, 0, 0);        // heads: B = 8 rows of heads16 per CTA
    const uint32_t idesc_c = umma_idesc_f16(256, 32, 1, 0);        // compositing: A = feat^T (MN-major view), B = Wgt
    const uint32_t act_base = smem_u32(smem + SM_ACT);
    const uint32_t stage_base = smem_u32(smem + SM_STAGE);
    const uint32_t aux_base = smem_u32(smem + SM_AUX);
    const uint32_t k16_base = smem_u32(smem + SM_K16);
    const uint32_t heads_addr = smem_u32(smem + SM_HEADS);
    Cursor c[2];
    c[0].init(a, ps0, n_pairslots); c[1].init(a, ps0 + 1, n_pairslots);
    int jb[2] = {0, 0};
    uint32_t n = 0, kcnt[2] = {0u, 0u}, acnt[2] = {0u, 0u};
    uint32_t n_shared = 0;                 // first ring stage of slot 0's layer job when slot 1 re-reads the same weights
    C3D_PROF_DECL(8);                      // 0 issue, 1 wait a_ready (layer jobs), 2 wait a_ready (narrow jobs), 3 wait kfull, 4 wait full
    while (c[0].valid() || c[1].valid()) {
      const bool shared = same_weights(a, c[0], c[1]);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (!c[s].valid()) continue;
        const int j = jb[s], kind = job_kind(j, D);
        const uint32_t tacc = tmem_base + (uint32_t)s * 256u;
        const uint32_t act_addr = act_base + (uint32_t)s * ACT_BYTES, aux_addr = aux_base + (uint32_t)s * AUX_BYTES;
        const uint32_t k16_addr = k16_base + (uint32_t)s * K16_BYTES;
        C3D_PROF(0);
        mbar_wait_cluster(&misc->a_ready[s], acnt[s] & 1u);
        C3D_PROF(kind <= 1 ? 1 : 2);
        acnt[s]++;
        tc_fence_after();
        if (kind <= 1) {
          mbar_wait_cluster(&misc->kfull[s], kcnt[s] & 1u);
          C3D_PROF(3);
          kcnt[s]++;
          tc_fence_after();
          const bool reuse = shared && s == 1;       // the stages slot 0 just used hold my weights too: no wait, I release them
          if (kind == 1 && s == 0) n_shared = n;
          const uint32_t n0 = reuse ? n_shared : n;
          const uint64_t auxd = umma_desc_kmajor_k16(aux_addr);
          if (issue) {
            umma_bf16_ss_pair(tacc, auxd, umma_desc_kmajor_k16(k16_addr), idesc_k, 0u);
            umma_commit_pair(&misc->kempty[s], (uint16_t)0x3);
          }
          if (kind == 1) {
#pragma unroll
            for (int kc = 0; kc < NCHUNK; ++kc) {
              const uint32_t m = n0 + (uint32_t)kc;
              const uint32_t st = m % NSTAGE, ph = (m / NSTAGE) & 1u;
              if (!reuse) {
                C3D_PROF(0);
                mbar_wait_cluster(&misc->full[st], ph);
                C3D_PROF(4);
                tc_fence_after();
              }
              if (issue) {
                const uint64_t ad = umma_desc_kmajor_sw128(act_addr + kc * ACT_CHUNK);
                const uint64_t bd = umma_desc_kmajor_sw128(stage_base + st * STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16_ss_pair(tacc, ad + 2 * kk, bd + 2 * kk, idesc_l, 1u);
                if (!(shared && s == 0)) umma_commit_pair(&misc->empty[st], (uint16_t)0x3);
              }
            }
          }
          if (issue) umma_commit_pair(&misc->acc_full[s], (uint16_t)0x3);
          if (kind == 1 && !reuse) n += NSTAGE;
        } else if (issue) {
          // These narrow MMAs are latency-bound when chained on one accumulator, so consecutive K-steps go to different
          // partial accumulators (summed by the epilogue): 2 per channel half for the compositing, 4 for the heads.
          {                                 // compositing: F^T[c][ray] = sum_p feat[p][c] * Wgt[ray][p], per channel half
            const uint64_t fa = umma_desc_mnmajor_sw128(act_addr, ACT_CHUNK), wb = umma_desc_kmajor_sw128(aux_addr);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                umma_bf16_ss_pair(tacc + (uint32_t)(h * 2 + (ks & 1)) * 32u, desc_at(fa, 2 * h * ACT_CHUNK + ks * 2048),
                                  desc_at(wb, (ks >> 2) * 2048 + (ks & 3) * 32), idesc_c, ks >= 2);
          }
          const uint32_t dcol = 128u;                          // rgb head (rows 0..2 of heads16)
          const uint64_t ha = umma_desc_kmajor_sw128(act_addr), hb = umma_desc_kmajor_sw128(heads_addr);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            umma_bf16_ss_pair(tacc + dcol + (uint32_t)(ks & 3) * 16u, desc_at(ha, (ks >> 2) * ACT_CHUNK + (ks & 3) * 32),
                              desc_at(hb, (ks >> 2) * 1024 + (ks & 3) * 32), idesc_h, ks >= 4);
          umma_commit_pair(&misc->acc_full[s], (uint16_t)0x3);
        }
        __syncwarp();
        if (++jb[s] == JOBS) { jb[s] = 0; c[s].next_tile(a); }
      }
    }
#ifdef C3D_KERNEL_PROF
    C3D_PROF(0);
    if (blockIdx.x % 42 == 0 && issue)
      printf("c3d prof pair mma[blk %d]: total %lld  issue %lld  wait_a_ready layers %lld narrow %lld  wait_kfull %lld  wait_full %lld\n", (int)blockIdx.x,
             clock64() - prof_begin, prof_t[0], prof_t[1], prof_t[2], prof_t[3], prof_t[4]);
#endif
   }
  } else if (warp >= 4) {
    // ============================================================ epilogue groups (both CTAs)
    const int s = (warp - 4) / EGW;
    const int te = (int)threadIdx.x - 128 - s * EGW * 32;
    const int t = te & 127;                              // my point row of the tile = my TMEM lane
    const int grp = te >> 7;                             // my 64 columns of each accumulator half
    const bool ptg = grp == 0;                           // this thread also runs the per-point stages
    const int quad = warp & 3;
    const uint32_t bar_id = 1u + (uint32_t)s;
    uint8_t* aux = smem + SM_AUX + s * AUX_BYTES;
    const uint32_t aux_u32 = smem_u32(aux);
    float* omS = reinterpret_cast<float*>(smem + SM_OM) + s * TILE;
    float* sdfS = reinterpret_cast<float*>(smem + SM_SDFP) + s * TILE;
    float* rayacc = reinterpret_cast<float*>(smem + SM_RAYACC) + s * RAYS * 2 * 8;
    float* raypart = reinterpret_cast<float*>(smem + SM_PART) + s * RAYS * 4 * 8;
    float* scratch = a.feat_scratch + (size_t)(2 * blockIdx.x + s) * a.unit_rays * W;   // channel-major outputs: this slot's unit
    const uint32_t tacc = tmem_base + (uint32_t)s * 256u + ((uint32_t)(quad * 32) << 16);
    const float* scal = reinterpret_cast<const float*>(a.blob + a.L.scal);
    const float bsig = scal[0], brgb0 = scal[1], brgb1 = scal[2], brgb2 = scal[3];
    const float inv_beta = 1.0f / scal[4];
    const uint32_t act_u32 = smem_u32(smem + SM_ACT + s * ACT_BYTES);
    const uint32_t row_u32 = act_u32 + (uint32_t)t * 128u;
    const int r7 = t & 7;
    const uint32_t ready_remote = mapa_u32(smem_u32(&misc->a_ready[s]), 0u);
    // "my part of the tile is written": every thread fences its own stores, one arrival per warp on the leader's barrier
    auto arrive_ready = [&]() {
      fence_proxy_async_smem();          // my stores all went to my own CTA's shared memory
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(&misc->a_ready[s]); else mbar_arrive_remote(ready_remote); }
    };
    float carry_f = 0.f;                                 // partial feature sum (my channel) of a ray continuing into the next tile
    uint32_t cnt = 0;                                    // completed phases of acc_full[s]
    Cursor cur;
    cur.init(a, ps0 + s, n_pairslots);
    C3D_PROF_DECL(12);   // 8 post: TMEM read-out, 9 feature stores, 10 ray sums + outputs, 11 geometry compute; 0 post epilogue (+ Wgt build), 1 wait acc, 2 density stage, 4 sines + stores first K-chunk, 5 second, 6 arrive, 7 geometry

    for (; cur.valid(); ) {
      const int u = cur.pu;
      const int img = u / a.units_per_img;
      const int r0 = (u - img * a.units_per_img) * 2 * a.unit_rays + (int)rank * a.unit_rays;
      const int nr = max(0, min(a.unit_rays, a.n_rays - r0));   // 0: this CTA idles through the pair-unit
      const int npts = nr * N;
      const int ntiles = cur.ntiles;
      const float near = a.near[img], far = a.far[img];
      const float nscale = 2.0f / (far - near);
      carry_f = 0.f;

      for (int tile = 0; tile < ntiles; ++tile, cur.next_tile(a)) {
        // ------------------------------------------------ geometry of my point (nerf_utils.py:17-170)
        const int q = tile * TILE + t;
        const bool valid = q < npts;
        const int qc = valid ? q : max(npts - 1, 0);
        const int rl = qc / N, k = qc - rl * N;
        const int rl0 = (tile * TILE) / N;                // first ray touching this tile
        const size_t gray = (size_t)img * a.n_rays + min(r0 + rl, a.n_rays - 1);
        float px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, dist = 0.f, zk = 0.f;
        if (npts == 0 || !ptg) {
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
        if (a.z_vals_out && valid && ptg) a.z_vals_out[gray * N + k] = zk;
        const uint32_t aux_row = aux_u32 + (uint32_t)((t >> 3) * 256 + (t & 7) * 16);
        if (ptg) {
          // point tile of the K16 products of layers 0..D-1: per coordinate (hi, mid, hi, lo); slots 12, 13 = 1 (shift)
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
          st_v4(aux_row + 128, pack_bf16x2(e[8], e[9]), pack_bf16x2(e[10], e[11]), pack_bf16x2(1.0f, 1.0f), 0u);
        }
        C3D_PROF(11);
        arrive_ready();
        C3D_PROF(7);

        float sdf = 0.f, wgt = 0.f;
        float sdfa[4] = {0.f, 0.f, 0.f, 0.f};                // partial sums of the sdf head over my channels
        // ------------------------------------------------ layers 0..D (D = view layer)
        for (int l = 0; l <= D; ++l) {
          // my point: channels 128 grp .. + 127 (K-chunks 2 grp, 2 grp + 1 of the next layer's input); TMEM loads double-buffered
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            C3D_PROF(h == 0 ? 0 : 4);
            if (h == 0) {
              mbar_wait(&misc->acc_full[s], cnt & 1u);
              cnt++;
              tc_fence_after();
              if (l == D && ptg) {
                // all MMAs of the view layer are complete: the view tile has been consumed, build Wgt[ray slot][point]
                // (fp16, K-major SW128) in its place
                const int myslot = rl - rl0;
#pragma unroll
                for (int jx = 0; jx < RAYS; ++jx)
                  st_f16(aux_u32 + (uint32_t)((t >> 6) * 2048) + sw128_offset(jx, t & 63), jx == myslot ? wgt : 0.f);
              }
            }
            C3D_PROF(h == 0 ? 1 : 3);
            const int chunk = 2 * grp + h;
            const uint32_t tcol = tacc + (uint32_t)chunk * 64u;
            const uint32_t row = row_u32 + (uint32_t)chunk * ACT_CHUNK;
            uint32_t v0[16], v1[16];
            tmem_ld_32x16(tcol, v0);
            const float4* wsg = reinterpret_cast<const float4*>(smem + SM_WSIG) + chunk * 16;   // 64 weights of this K-chunk
            if (l == D - 1) {
#pragma unroll
              for (int cq = 0; cq < 2; ++cq) {             // 32 channels per iteration: units 4 cq .. 4 cq + 3 of my row
                tmem_ld_wait();
                tmem_ld_32x16(tcol + cq * 32 + 16, v1);
                epilogue16_sdf(v0, row, 4 * cq, r7, wsg + 8 * cq, sdfa);
                tmem_ld_wait();
                if (cq < 1) tmem_ld_32x16(tcol + (cq + 1) * 32, v0);
                epilogue16_sdf(v1, row, 4 * cq + 2, r7, wsg + 8 * cq + 4, sdfa);
              }
            } else {
#pragma unroll
              for (int cq = 0; cq < 2; ++cq) {
                tmem_ld_wait();
                tmem_ld_32x16(tcol + cq * 32 + 16, v1);
                epilogue16(v0, row, 4 * cq, r7);
                tmem_ld_wait();
                if (cq < 1) tmem_ld_32x16(tcol + (cq + 1) * 32, v0);
                epilogue16(v1, row, 4 * cq + 2, r7);
              }
            }
          }
          C3D_PROF(5);
          if (l == D - 1) {
            const float sdfp = (sdfa[0] + sdfa[1]) + (sdfa[2] + sdfa[3]);
            // the last hidden layer is stored: hand the tile to the view-layer job first (view tile in place of the point
            // tile: slots 0..2 / 3..5 = hi / lo of the view direction, 12, 13 = 1), then -- under its MMAs -- density,
            // alpha and transmittance of my point (thread = point; the column group 1 thread passes its half of the sdf sum)
            if (ptg) {
              const float h0 = __bfloat162float(__float2bfloat16_rn(vx)), h1 = __bfloat162float(__float2bfloat16_rn(vy)),
                          h2 = __bfloat162float(__float2bfloat16_rn(vz));
              st_v4(aux_row, pack_bf16x2(h0, h1), pack_bf16x2(h2, vx - h0), pack_bf16x2(vy - h1, vz - h2), 0u);
              st_v4(aux_row + 128, 0u, 0u, pack_bf16x2(1.0f, 1.0f), 0u);
            } else {
              sdfS[t] = sdfp;
            }
            arrive_ready();
            C3D_PROF(6);
            named_bar_sync(bar_id + 2u, 2 * TILE);         // all 8 warps of the slot: the partial sums are written
            if (ptg) {
              sdf = bsig + sdfp + sdfS[t];
              if (valid) a.sdf[gray * N + k] = sdf;
              const float sigma = sigmoid_precise(-sdf * inv_beta) * inv_beta;
              const float alpha = 1.0f - expf(-sigma * dist);
              const float om = 1.0f - alpha + 1e-10f;
              omS[t] = valid ? om : 1.0f;
              named_bar_sync(bar_id, TILE);
              const int first_row = t - k;
              float T = first_row < 0 ? misc->carry[s] : 1.0f;
              for (int m = max(first_row, 0); m < t; ++m) T *= omS[m];
              wgt = valid ? alpha * T : 0.f;
              named_bar_sync(bar_id, TILE);
              if (t == TILE - 1) misc->carry[s] = (k == N - 1) ? 1.0f : T * om;
            }
            sdfa[0] = sdfa[1] = sdfa[2] = sdfa[3] = 0.f;
            C3D_PROF(2);
          } else {
            arrive_ready();
            C3D_PROF(6);
          }
        }

        // ------------------------------------------------ post job: composited features (thread = channel) + rgb (thread = point)
        {
          C3D_PROF(0);
          mbar_wait(&misc->acc_full[s], cnt & 1u);
          C3D_PROF(1);
          cnt++;
          tc_fence_after();
          float rgbv[3] = {brgb0, brgb1, brgb2};         // raw rgb of my point: 4 partial sums
          if (ptg) {
            uint32_t v4[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) tmem_ld_32x4(tacc + 128 + pp * 16, v4[pp]);
            tmem_ld_wait();
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              rgbv[0] += __uint_as_float(v4[pp][0]); rgbv[1] += __uint_as_float(v4[pp][1]); rgbv[2] += __uint_as_float(v4[pp][2]);
            }
          }
          const int tile_end = min((tile + 1) * TILE, npts);      // first point index beyond this tile
          {
            const int hh = grp;                          // one channel half per thread: lane t <-> channel t + 128 hh
            uint32_t fv[16], fw[16];                     // my CTA's 16 ray-slot columns of channel t + 128 hh, 2 partial sums
            tmem_ld_32x16(tacc + (uint32_t)(hh * 2) * 32u + rank * 16u, fv);
            tmem_ld_32x16(tacc + (uint32_t)(hh * 2 + 1) * 32u + rank * 16u, fw);
            tmem_ld_wait();
#pragma unroll
            for (int jx = 0; jx < 16; ++jx) fv[jx] = __float_as_uint(__uint_as_float(fv[jx]) + __uint_as_float(fw[jx]));
            C3D_PROF(8);
            // thread = channel t + 128 hh, fv[jx] = ray slot jx.  (b, hw, 256): a warp writes 32 consecutive channels of a
            // ray (128 B); (b, 256, hw): a thread writes up to 16 consecutive rays of its channel (64 B)
            const int ray0 = min(r0 + rl0, a.n_rays - 1), ch = t + TILE * hh;
            // (b, hw, 256): straight to the destination(s) -- this rank's feature_map or, fused all-gather, every peer's gathered
            // tensor.  Channel-major layouts: into this slot's L2-resident scratch [unit ray][channel] first; the unit is
            // transposed with 16-byte stores when its last tile is done (below).  Either way a warp writes 32 consecutive
            // channels of a ray (128 B).
            const int ndst = (a.n_peers > 0 && !a.feat_nchw) ? a.n_peers : 1;
            const size_t gimg = (size_t)(a.n_peers > 0 ? a.gather_off : 0) + img;
            const size_t foff = a.feat_nchw ? (size_t)rl0 * W + ch : (gimg * a.n_rays + ray0) * W + ch;
            // ray slot jx covers the unit's points [(rl0 + jx) N, (rl0 + jx + 1) N): the first jdone slots end inside this tile
            // (complete rays: stored), slot jdone -- if it has points here -- continues in the next tile (carried).  Kept to a
            // few predicated instructions per slot: this block runs once per tile on the critical path of the slot's chain.
            const int jdone = tile_end / N - rl0;
            const bool partial = (rl0 + jdone) * N < tile_end;
            fv[0] = __float_as_uint(__uint_as_float(fv[0]) + carry_f);
            for (int pr = 0; pr < ndst; ++pr) {
              float* dst = (a.feat_nchw ? scratch : (a.n_peers > 0 ? reinterpret_cast<float*>(a.peer_feat[pr]) : a.feature_map)) + foff;
#pragma unroll
              for (int jx = 0; jx < RAYS; ++jx)
                if (jx < jdone) dst[jx * W] = __uint_as_float(fv[jx]);
            }
            float cnext = 0.f;
#pragma unroll
            for (int jx = 0; jx < RAYS; ++jx)
              if (jx == jdone) cnext = __uint_as_float(fv[jx]);
            carry_f = partial ? cnext : 0.f;
          }
          C3D_PROF(9);
          if (a.feat_nchw && tile == ntiles - 1) {
            // ---- the unit's rays are complete: scratch [ray][channel] -> (b, 256, hw), thread = channel, 8 rays per 16-byte
            // (bf16) / 2 x 16-byte (fp32) store; to every peer's gathered tensor when the all-gather is fused
            __threadfence_block();
            named_bar_sync(bar_id + 2u, 2 * TILE);
            const int ndst = a.n_peers > 0 ? a.n_peers : 1;
            const size_t gimg = (size_t)(a.n_peers > 0 ? a.gather_off : 0) + img;
            const size_t obase = (gimg * W + te) * a.n_rays + r0;
            const bool fbf16 = a.feat_nchw == 2;
            const bool vec = ((a.n_rays | r0) & 7) == 0;            // 16-byte aligned runs of 8 rays
            int ry = 0;
            for (; vec && ry + 8 <= nr; ry += 8) {
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = __ldcg(scratch + (size_t)(ry + i) * W + te);
              for (int pr = 0; pr < ndst; ++pr) {
                void* fb = a.n_peers > 0 ? a.peer_feat[pr] : (void*)a.feature_map;
                if (fbf16) {
                  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(fb) + obase + ry) =
                      make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                } else {
                  float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(fb) + obase + ry);
                  o4[0] = make_float4(x[0], x[1], x[2], x[3]);
                  o4[1] = make_float4(x[4], x[5], x[6], x[7]);
                }
              }
            }
            for (; ry < nr; ++ry) {                                  // ragged shapes: one ray at a time
              const float x = __ldcg(scratch + (size_t)ry * W + te);
              for (int pr = 0; pr < ndst; ++pr) {
                void* fb = a.n_peers > 0 ? a.peer_feat[pr] : (void*)a.feature_map;
                if (fbf16) reinterpret_cast<__nv_bfloat16*>(fb)[obase + ry] = __float2bfloat16_rn(x);
                else reinterpret_cast<float*>(fb)[obase + ry] = x;
              }
            }
            named_bar_sync(bar_id + 2u, 2 * TILE);                 // the scratch may be overwritten by the next unit
          }
          tmem_ld_wait();
          tc_fence_before();
          if (ptg) {
          // rgb / xyz / mask sums of my ray (nerf_utils.py:315,329-336)
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
            const int ndst = a.n_peers > 0 ? a.n_peers : 1;
            const size_t gr = gray + (a.n_peers > 0 ? (size_t)a.gather_off * a.n_rays : 0);
            for (int pr = 0; pr < ndst; ++pr) {
              float* o3 = (a.n_peers > 0 ? a.peer_rgb[pr] : a.rgb_map) + gr * 3;
              o3[0] = -1.0f + 2.0f * racc[0]; o3[1] = -1.0f + 2.0f * racc[1]; o3[2] = -1.0f + 2.0f * racc[2];
              float* x3 = (a.n_peers > 0 ? a.peer_xyz[pr] : a.xyz) + gr * 3;
              x3[0] = x; x3[1] = y; x3[2] = z;
              float* m2 = (a.n_peers > 0 ? a.peer_mask[pr] : a.mask) + gr * 2;
              m2[0] = wgt;
              m2[1] = -sqrtf(x * x + y * y + z * z);
            }
#pragma unroll
            for (int jx = 0; jx < 6; ++jx) racc[jx] = 0.f;
          }
          }
        }
        C3D_PROF(10);
      }  // tiles
    }    // pair-units
#ifdef C3D_KERNEL_PROF
    C3D_PROF(0);
    if (blockIdx.x < 2 && t == 0)
      printf("c3d prof pair eg[blk %d slot %d grp %d]: total %lld  post epilogue (+ Wgt build) %lld  wait acc %lld  density stage %lld  (%lld)  sines chunk 0 %lld chunk 1 %lld  arrive %lld  first arrive %lld | post: read-out %lld feature stores %lld ray sums %lld geometry %lld\n",
             (int)blockIdx.x, s, grp, clock64() - prof_begin, prof_t[0], prof_t[1], prof_t[2], prof_t[3], prof_t[4], prof_t[5], prof_t[6], prof_t[7], prof_t[8], prof_t[9], prof_t[10], prof_t[11]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

}}  // namespace c3d::pairk
