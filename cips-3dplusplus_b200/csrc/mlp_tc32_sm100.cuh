// fp32 parity mode of the FiLM-SIREN point MLP (volume_renderer.py:133-160) on the tensor cores of sm_100a.
//
// The fp32 products W_l h are formed from IEEE half-precision operands split two ways, x = hi + 2^-11 lo' with
// hi = fp16(x) and lo' = fp16(2^11 (x - hi)) (22 bits of mantissa, the scaled low part stays in fp16's normal range),
// three tcgen05 products per layer:
//     acc0 = Wh Ah          acc1 = Wh Al' + Wl' Ah          W h = acc0 + 2^-11 acc1
// The dropped term (Wl Al) is 2^-22 relative.  Measured on B200 (bench_tools/umma_accum_probe.py): a K = 256 chain in TMEM has
// an rms error of 1.2e-7 on |acc| ~ 0.6 -- the level of a CPU fp32 GEMM (1.7e-7); the two small products go to a second
// accumulator so that the 16-step main chain is not lengthened (TMEM truncates when it accumulates) and so that they can carry
// their own scale.  Round 2's first version split three ways in bf16 (six products per layer, 257 ms per configs[1] step);
// fp16's 11-bit mantissa needs half the tensor work for the same 2.4e-7.
// Sines run on the FP32 pipe: Cody-Waite reduction by pi (two constants, FMA) and an odd degree-9 polynomial, max error 1.1e-7
// on |x| <= 100 (sinf: 0.7e-7); FiLM is applied in the epilogue in fp32; layer 0 (K = 3), the view-direction columns and both
// heads run on the FP32 pipe as in mlp_fp32_kernel.  Per-point outputs (features, raw rgb, sdf) go to HBM for
// composite_fwd_kernel exactly as in the FP32-pipe kernel, whose argument struct this kernel shares.
//
// Structure: clusters of two CTAs, one pair-tile of 2 x 128 points at a time.  TMEM lanes are points (D[point][channel],
// tcgen05.mma.cta_group::2, M = 256, N = 256), so every CTA keeps the two split tiles of ITS 128 points in shared memory
// (2 x 64 KB, K-major SWIZZLE_128B, written in place by the epilogue) and stages only its 128 rows of each weight chunk
// (16 KB stages, a ring of five).  Warp 0 streams the stages with cp.async.bulk in the order the issuer consumes them -- per
// K-chunk: Wh (x Ah, Al'), Wl' (x Ah) -- 12 MMAs of 128 cycles per K-chunk against two 16 KB copies.  Warp 1 issues (leader
// CTA) or relays the peer's "stage landed" (follower), warp 2 owns TMEM, warps 4..19 are the epilogue: warp (quad, grp) handles
// TMEM lanes 32 quad.. (points) x columns 64 grp.. (channels).
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"
#include "mlp_fp32.cuh"
#include <type_traits>

namespace c3d { namespace tc32 {

using namespace c3d::ptx;

constexpr int TILE = 128;                       // points per CTA and tile
constexpr int EPW = 16;                         // epilogue warps
constexpr int NTHREADS = 128 + EPW * 32;        // 640
constexpr int SPLIT_BYTES = 65536;              // one split tile: [4 K-chunks][128 points][64 k] fp16
constexpr int CHUNK_BYTES = 16384;
constexpr int STAGE_BYTES = 16384;              // [128 weight rows of this CTA][64 k]
constexpr int NSTAGE = 5;
constexpr int NSPLIT = 2;
constexpr float LO_SCALE = 2048.0f, LO_UNSCALE = 1.0f / 2048.0f;
constexpr int SM_ACT = 0;                                   // hi, lo'
constexpr int SM_STAGE = NSPLIT * SPLIT_BYTES;              // 131072
constexpr int SM_RED = SM_STAGE + NSTAGE * STAGE_BYTES;     // 212992  [4 column groups][128] float: sdf partial sums
constexpr int SM_SCR = SM_RED + 4 * TILE * 4;               // 215040  [4 column groups][128] float4: rgb partial sums
constexpr int SM_MISC = SM_SCR + 4 * TILE * 16;             // 223232
constexpr int SMEM_BYTES = SM_MISC + 128;                   // 223360: no slack for re-alignment, see the trap below

struct Misc {
  uint64_t full[NSTAGE], empty[NSTAGE], a_ready, acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void st_v4(uint32_t smem_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// x[0..7] -> two 16-byte units (hi, scaled lo fp16 of 8 consecutive channels) at the same offset of the two split tiles
__device__ __forceinline__ void split_store8(const float* x, uint32_t addr_hi) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_f16x2(x[2 * i], x[2 * i + 1]);
    const float2 f = unpack_f16x2(h[i]);
    l[i] = pack_f16x2((x[2 * i] - f.x) * LO_SCALE, (x[2 * i + 1] - f.y) * LO_SCALE);
  }
  st_v4(addr_hi, h[0], h[1], h[2], h[3]);
  st_v4(addr_hi + SPLIT_BYTES, l[0], l[1], l[2], l[3]);
}

// sin(x) on the FP32 pipe: n = rint(x / pi) by the magic-number trick, r = x - n pi with pi = PI_HI + PI_LO (FMA: one rounding
// each), sign from the parity of n, odd degree-9 minimax polynomial on [-pi/2, pi/2].  Max error 1.1e-7 for |x| <= 100
// (bench_tools/sin_poly_fit.py).  Branch-free so that the 16 sines of a chunk interleave; beyond |x| = 8192 the reduction runs
// out of bits and the whole chunk is redone by libdevice (never on the reference's value ranges).
__device__ __forceinline__ float sin_poly(float x) {
  const float t = fmaf(x, 0.318309886f, 12582912.0f);
  const float n = t - 12582912.0f;
  float r = fmaf(n, -3.14159274f, x);
  r = fmaf(n, 8.742278e-8f, r);
  r = __uint_as_float(__float_as_uint(r) ^ (__float_as_uint(t) << 31));
  const float s = r * r;
  float q = fmaf(2.5943759e-06f, s, -1.9804016e-04f);
  q = fmaf(q, s, 8.3329808e-03f);
  q = fmaf(q, s, -1.6666655e-01f);
  return fmaf(r * s, q, r);
}
__device__ __noinline__ float sin_far(float x) { return sinf(x); }
template <int N>
__device__ __forceinline__ void sin_inplace(float (&x)[N]) {
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) m = fmaxf(m, fabsf(x[i]));
  if (__builtin_expect(m > 8192.0f, 0)) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = sin_far(x[i]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = sin_poly(x[i]);
  }
}

// in-kernel cycle counters of one epilogue thread and of the issuer (development builds only: -DC3D_KERNEL_PROF)
#ifdef C3D_KERNEL_PROF
#define TC32_PROF_DECL(n) long long pt_[n] = {}; long long pm_ = clock64(); const long long pb_ = pm_
#define TC32_PROF(i) do { const long long now_ = clock64(); pt_[i] += now_ - pm_; pm_ = now_; } while (0)
#else
#define TC32_PROF_DECL(n) do { } while (0)
#define TC32_PROF(i) do { } while (0)
#endif

__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc32_kernel(const MlpF32Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc* misc = reinterpret_cast<Misc*>(smem + SM_MISC);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int D = a.L.D;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int ncl = (int)(gridDim.x >> 1), cl = (int)(blockIdx.x >> 1);
  const int ptpi = (a.pts_per_img + 2 * TILE - 1) / (2 * TILE);     // pair-tiles per image
  const int total = a.n_imgs * ptpi;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();                     // the swizzled tiles need 1024-byte alignment

  if (threadIdx.x == 32) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&misc->full[i], leader ? 2 : 1); mbar_init(&misc->empty[i], 1); }
    mbar_init(&misc->a_ready, 2 * EPW);            // one arrival per epilogue warp of both CTAs (the leader's copy is used)
    mbar_init(&misc->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc_pair(&misc->tmem_base, 512); tmem_relinquish_pair(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, misc->tmem_base, 0);

  if (warp == 0) {
    if (elect_one()) {
      // ============================================================ weight producer: this CTA's 128 rows of every chunk
      const uint8_t* img[NSPLIT] = {a.blob + a.L.wf16h, a.blob + a.L.wf16l};
      uint32_t n = 0;
      for (int pt = cl; pt < total; pt += ncl)
        for (int l = 1; l <= D; ++l)
          for (int kc = 0; kc < NCHUNK; ++kc)
#pragma unroll
            for (int sp = 0; sp < NSPLIT; ++sp, ++n) {
              const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
              mbar_wait(&misc->empty[st], ph ^ 1u);
              mbar_arrive_expect_tx(&misc->full[st], STAGE_BYTES);
              bulk_g2s(smem + SM_STAGE + st * STAGE_BYTES,
                       img[sp] + (size_t)(l - 1) * WBF16_LAYER_BYTES + (size_t)kc * WBF16_CHUNK_BYTES + rank * STAGE_BYTES,
                       STAGE_BYTES, &misc->full[st]);
            }
    }
  } else if (warp == 1 && !leader) {
    if (elect_one()) {
      // ============================================================ relay: "my stage landed" -> the leader's full barrier
      const uint32_t r_full = mapa_u32(smem_u32(&misc->full[0]), 0u);
      uint32_t n = 0;
      for (int pt = cl; pt < total; pt += ncl)
        for (int i = 0; i < D * NCHUNK * NSPLIT; ++i, ++n) {
          const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
          mbar_wait(&misc->full[st], ph);
          mbar_arrive_remote(r_full + st * 8u);
        }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (leader; whole warp runs the control flow)
    const bool issue = elect_one();
    const uint32_t idesc = umma_idesc_f16(256, 256);
    const uint32_t act_base = smem_u32(smem + SM_ACT), stage_base = smem_u32(smem + SM_STAGE);
    const uint32_t acc0 = tmem_base, acc1 = tmem_base + 256u;
    uint32_t n = 0, acnt = 0;
    TC32_PROF_DECL(4);                                   // 0 issue, 1 wait a_ready, 2 wait full
    for (int pt = cl; pt < total; pt += ncl)
      for (int l = 1; l <= D; ++l) {
        TC32_PROF(0);
        mbar_wait_cluster(&misc->a_ready, acnt & 1u);
        TC32_PROF(1);
        acnt++;
        tc_fence_after();
#pragma unroll 1
        for (int kc = 0; kc < NCHUNK; ++kc)
#pragma unroll
          for (int sp = 0; sp < NSPLIT; ++sp, ++n) {
            const uint32_t st = n % NSTAGE, ph = (n / NSTAGE) & 1u;
            TC32_PROF(0);
            mbar_wait_cluster(&misc->full[st], ph);
            TC32_PROF(2);
            tc_fence_after();
            if (issue) {
              const uint64_t bd = umma_desc_kmajor_sw128(stage_base + st * STAGE_BYTES);
#pragma unroll
              for (int asp = 0; asp < NSPLIT - sp; ++asp) {       // Wh x (Ah, Al'), Wl' x Ah
                const uint64_t ad = umma_desc_kmajor_sw128(act_base + asp * SPLIT_BYTES + kc * CHUNK_BYTES);
                const bool main_acc = sp == 0 && asp == 0;
                const bool first = kc == 0 && sp == 0 && asp <= 1;      // first MMA into acc0 (asp 0) / acc1 (asp 1)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16_ss_pair(main_acc ? acc0 : acc1, ad + 2 * kk, bd + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
              }
              umma_commit_pair(&misc->empty[st], (uint16_t)0x3);
            }
          }
        if (issue) umma_commit_pair(&misc->acc_full, (uint16_t)0x3);
        __syncwarp();
      }
#ifdef C3D_KERNEL_PROF
    TC32_PROF(0);
    if (blockIdx.x == 0 && issue)
      printf("tc32 prof mma: total %lld  issue %lld  wait a_ready %lld  wait full %lld\n", clock64() - pb_, pt_[0], pt_[1], pt_[2]);
#endif
  } else if (warp >= 4) {
    // ============================================================ epilogue (both CTAs)
    const int e = warp - 4, quad = e & 3, grp = e >> 2;
    const int row = quad * 32 + lane;                          // my point of the tile = my TMEM lane
    const int c0 = grp * 64;                                   // my 64 channels = K-chunk grp of the next layer's operand
    const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, t1 = t0 + 256u;
    const uint32_t row_hi = smem_u32(smem + SM_ACT) + (uint32_t)grp * CHUNK_BYTES + (uint32_t)row * 128u;
    // feature staging of the view layer: my warp's own 2 x 4 KB of the split tiles (its 32 rows of K-chunk grp, hi and lo tile)
    // as [32 points][64 channels] fp32, rows 0..15 in the hi part, 16..31 in the lo part, 16-byte units XOR-swizzled with the row
    const uint32_t stage_warp = smem_u32(smem + SM_ACT) + (uint32_t)grp * CHUNK_BYTES + (uint32_t)(quad * 32) * 128u;
    const uint32_t stage_row = stage_warp + (uint32_t)((lane & 16) ? SPLIT_BYTES : 0) + (uint32_t)((lane & 15) * 256);
    const int r7 = row & 7;
    const uint32_t ready_remote = mapa_u32(smem_u32(&misc->a_ready), 0u);
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    const float* wsig = reinterpret_cast<const float*>(a.blob + a.L.wsig);
    const float4* wrgb = reinterpret_cast<const float4*>(a.blob + a.L.wrgb);
    const float* scal = reinterpret_cast<const float*>(a.blob + a.L.scal);
    const int n_rays = a.pts_per_img / a.n_samples;
    uint32_t cnt = 0;
    TC32_PROF_DECL(4);                                   // 0 layer 0 (+ tile set-up), 1 wait acc, 2 epilogues, 3 view layer + heads
    auto arrive_ready = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(&misc->a_ready); else mbar_arrive_remote(ready_remote); }
    };
    auto sdf_out = [&](float part, size_t gp, bool valid) {   // sum of the four column groups' partial dot products
      red[grp * TILE + row] = part;
      named_bar_sync(1, EPW * 32);
      if (grp == 0 && valid) a.sdf[gp] = ((red[row] + red[TILE + row]) + (red[2 * TILE + row] + red[3 * TILE + row])) + scal[0];
    };

    for (int pt = cl; pt < total; pt += ncl) {
      const int img = pt / ptpi;
      const int q = (pt - img * ptpi) * 2 * TILE + (int)rank * TILE + row;
      const bool valid = q < a.pts_per_img;
      const int qc = valid ? q : a.pts_per_img - 1;
      const size_t gp = (size_t)img * a.pts_per_img + qc;
      const float ns = 2.0f / (a.far[img] - a.near[img]);                   // normalize_points, nerf_utils.py:130
      const float px = a.pts[gp * 3 + 0] * ns, py = a.pts[gp * 3 + 1] * ns, pz = a.pts[gp * 3 + 2] * ns;
      const float* vd = a.viewdirs + ((size_t)img * n_rays + qc / a.n_samples) * 3;
      const float vx = vd[0], vy = vd[1], vz = vd[2];
      // ---- layer 0 on the FP32 pipe: sin(gamma0 (W0 p + b0) + beta0)
      {
        const float4* first = a.first + (size_t)img * W + c0;
        float sp = 0.f;
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(first + u * 8 + i);
            x[i] = fmaf(t.x, px, fmaf(t.y, py, fmaf(t.z, pz, t.w)));
          }
          sin_inplace(x);
          if (D == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sp = fmaf(__ldg(wsig + c0 + u * 8 + i), x[i], sp);
          }
          split_store8(x, row_hi + (uint32_t)((u ^ r7) << 4));
        }
        if (D == 1) sdf_out(sp, gp, valid);
      }
      arrive_ready();
      TC32_PROF(0);
      // ---- layers 1 .. D
      for (int l = 1; l <= D; ++l) {
        mbar_wait(&misc->acc_full, cnt & 1u);
        TC32_PROF(1);
        cnt++;
        tc_fence_after();
        const float2* film = a.film + ((size_t)img * (D + 1) + l) * W + c0;
        const float4* view = a.view + (size_t)img * W + c0;
        float* save = a.save_acc ? a.save_acc + (size_t)(l - 1) * a.save_stride + gp * W + c0 : nullptr;
        float sp = 0.f, rr = 0.f, rg = 0.f, rb = 0.f;
        // 16 channels per iteration; kind: 0 hidden layer, 1 last hidden layer (+ sdf head), 2 view layer (+ rgb head, features)
        auto chunk = [&](int j, auto kind_tag) {
          constexpr int kind = decltype(kind_tag)::value;
          uint32_t v0[16], v1[16];
          ld16(t0 + j * 16, v0);
          ld16(t1 + j * 16, v1);
          tmem_ld_wait();
          float x[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = fmaf(__uint_as_float(v1[i]), LO_UNSCALE, __uint_as_float(v0[i]));
          if (save && valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(save + j * 16)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(film + j * 16 + i));     // (gamma, shift) of two channels
            if (kind == 2) {                                  // view layer: + gammaD * Wview[:, 256..258] . viewdir
              const float4 tv = __ldg(view + j * 16 + i), tw = __ldg(view + j * 16 + i + 1);
              x[i] = fmaf(f.x, x[i], fmaf(tv.x, vx, fmaf(tv.y, vy, tv.z * vz)) + f.y);
              x[i + 1] = fmaf(f.z, x[i + 1], fmaf(tw.x, vx, fmaf(tw.y, vy, tw.z * vz)) + f.w);
            } else {
              x[i] = fmaf(f.x, x[i], f.y);
              x[i + 1] = fmaf(f.z, x[i + 1], f.w);
            }
          }
          sin_inplace(x);
          if (kind < 2) {
            if (kind == 1) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(wsig + c0 + j * 16 + i));
                sp = fmaf(w.x, x[i], sp); sp = fmaf(w.y, x[i + 1], sp); sp = fmaf(w.z, x[i + 2], sp); sp = fmaf(w.w, x[i + 3], sp);
              }
            }
            split_store8(x, row_hi + (uint32_t)(((2 * j) ^ r7) << 4));
            split_store8(x + 8, row_hi + (uint32_t)(((2 * j + 1) ^ r7) << 4));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 w = __ldg(wrgb + c0 + j * 16 + i);
              rr = fmaf(w.x, x[i], rr); rg = fmaf(w.y, x[i], rg); rb = fmaf(w.z, x[i], rb);
            }
            // features: staged in this warp's own (now dead) rows of the split tiles, written out row by row below -- a thread
            // storing its own 64 B of a 1 KB row made every STG.128 touch 32 different lines (8.2 k LSU wavefronts per tile-layer)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (uint32_t)(((4 * j + i) ^ (lane & 15)) << 4)),
                           "f"(x[4 * i]), "f"(x[4 * i + 1]), "f"(x[4 * i + 2]), "f"(x[4 * i + 3]) : "memory");
          }
        };
        if (l == D) {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) chunk(j, std::integral_constant<int, 2>{});
          __syncwarp();
          // my warp's [32 points][64 channels] block: 16 lanes per row (256 B contiguous), two rows per instruction
          const int qw = (pt - img * ptpi) * 2 * TILE + (int)rank * TILE + quad * 32;      // first point of my warp
#pragma unroll 4
          for (int kk = 0; kk < 16; ++kk) {
            const int rw = 2 * kk + (lane >> 4), u = lane & 15;
            const uint32_t src = stage_warp + (uint32_t)((rw & 16) ? SPLIT_BYTES : 0) + (uint32_t)((rw & 15) * 256) +
                                 (uint32_t)((u ^ (rw & 15)) << 4);
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src) : "memory");
            if (qw + rw < a.pts_per_img)
              *reinterpret_cast<float4*>(a.feat + ((size_t)img * a.pts_per_img + qw + rw) * W + c0 + u * 4) = v;
          }
          __syncwarp();
        } else if (l == D - 1) {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) chunk(j, std::integral_constant<int, 1>{});
        } else {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) chunk(j, std::integral_constant<int, 0>{});
        }
        if (l < D) {
          if (l == D - 1) sdf_out(sp, gp, valid);
          arrive_ready();
          TC32_PROF(2);
        } else {
          // rgb head: partial sums of the four column groups
          float4* scr = reinterpret_cast<float4*>(smem + SM_SCR);
          scr[grp * TILE + row] = make_float4(rr, rg, rb, 0.f);
          named_bar_sync(1, EPW * 32);
          if (grp == 0) {
            const float4 p0 = scr[row], p1 = scr[TILE + row], p2 = scr[2 * TILE + row], p3 = scr[3 * TILE + row];
            if (valid) {
              float* o = a.rgb + gp * 3;
              o[0] = ((p0.x + p1.x) + (p2.x + p3.x)) + scal[1];
              o[1] = ((p0.y + p1.y) + (p2.y + p3.y)) + scal[2];
              o[2] = ((p0.z + p1.z) + (p2.z + p3.z)) + scal[3];
            }
          }
          TC32_PROF(3);
        }
      }
    }
#ifdef C3D_KERNEL_PROF
    if (blockIdx.x == 0 && threadIdx.x == 128)
      printf("tc32 prof eg: total %lld  layer 0 %lld  wait acc %lld  epilogues %lld  view layer %lld\n", clock64() - pb_, pt_[0], pt_[1], pt_[2], pt_[3]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

}}  // namespace c3d::tc32
