// Shared pieces of the fused forward path: kernel arguments, the static tile schedule, and the tcgen05 layout
// self-test kernels (exercised by tests/test_gpu_parity.py through c3d_umma_selftest).
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"

namespace c3d { namespace fused {

using namespace c3d::ptx;

constexpr int TILE = 128;                      // points per tile
constexpr int ACT_CHUNK = TILE * 128;          // 16384 B: [128 rows][64 bf16], K-major SWIZZLE_128B
constexpr int ACT_BYTES = NCHUNK * ACT_CHUNK;  // 65536
constexpr int MIN_SAMPLES = 8;                 // <= 16 rays touch one 128-point tile

struct Args {
  const uint8_t* blob; PackedLayout L;
  const float2* film; const float4* first; const float4* view;   // style_prep tables, indexed by image
  int batch, n_rays, n_samples, D, img_size, static_viewdirs, input_kind;
  int unit_rays, units_per_img;
  const float* cam_poses; const float* focal; const float* near; const float* far; const float* ray_offset;
  const float* pts; const float* rays_d; const float* viewdirs; const float* z_vals;
  float* rgb_map; float* feature_map; float* sdf; float* mask; float* xyz; float* z_vals_out;
  int feat_nchw; // feature_map layout: 0 = (b, hw, 256) as the reference returns it, 1 = (b, 256, hw) as the decoder wants it
                 // (model_v3.py:1014); written directly from the compositing epilogue either way
  // fused all-gather (c3d_gather_out): the maps of image `img` go to image `gather_off + img` of every peer's gathered tensors
  // (peer memory mapped over NVLink); n_peers == 0: the four pointers above
  int n_peers, gather_off;
  float* feat_scratch;   // CTA-pair kernel, channel-major feature_map: [2 * grid slots][unit_rays][256] fp32, L2-resident
  void* peer_feat[C3D_MAX_PEERS]; float* peer_rgb[C3D_MAX_PEERS]; float* peer_mask[C3D_MAX_PEERS]; float* peer_xyz[C3D_MAX_PEERS];
  int sdf_only;  // density-only pass (coarse pass of the two-pass render): the tile ends after the sdf head -- no view layer,
                 // no rgb head, no compositing; only `sdf` (and `z_vals_out`) are written
  int debug;   // bit 0: producer skips the weight copies (timing experiments only; results are garbage)
  int stagger;                                // CTA-pair kernel: slot 1 starts this many cycles after slot 0
  const uint8_t* wimg; const uint8_t* kimg;   // CTA-pair kernel: per-image FiLM-folded weight images (film_weights_kernel)
  // ---- backward support (NULL in plain forward launches) ----
  // bf16 tiles [layer][tile_g][point group 16][channel 256][8 points]; tile_g = unit * tiles_per_unit + tile
  __nv_bfloat16* save_acc;    // pre-FiLM accumulators, layers 0..D, stored as fp16 bit patterns (the backward recomputes
                              // cos(scale * acc + shift) from them)
  float* rgb_pt;              // (b, n_rays, N, 3) raw rgb head output
  float* w_pt;                // (b, n_rays, N) compositing weights
  int tiles_per_unit; long long n_tiles_g;
  // backward kernel inputs / outputs
  const float* g_rgb_pt; const float* g_sdf_pt; const float* g_feature_map;
  float* g_film; float* g_pts; float* g_viewdirs;
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


// Self-test of the CTA-pair MMA (cta_group::2), launched as ONE cluster of two CTAs:
//   D[256][N] = A[256][K] * B[N][K]^T ; CTA r stages A rows r*128.. and B rows r*N/2.. ; the leader issues, both read back.
// variant bit 0: A stored MN-major ([k][m], the compositing operand view).
__global__ void __launch_bounds__(128, 1) umma_pair_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                                     float* __restrict__ Dout, int N, int K, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // K-major: K/64 chunks x [128][64] ; MN-major: 2 blocks of [K rows][64 m]
  uint8_t* sB = smem + 65536;               // K/64 chunks x [N/2][64]
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool a_mn = variant & 1;
  const int nh = N / 2;
  const uint32_t blk = (uint32_t)K * 128u;
  for (int idx = threadIdx.x; idx < 128 * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    const uint32_t off = a_mn ? (uint32_t)(r >> 6) * blk + (uint32_t)k * 128u + (uint32_t)((((r & 63) >> 3) ^ (k & 7)) << 4) + (uint32_t)(r & 7) * 2u
                              : (uint32_t)(k >> 6) * 16384u + sw128_offset(r, k & 63);
    *reinterpret_cast<uint16_t*>(sA + off) = A[(size_t)(rank * 128 + r) * K + k];
  }
  for (int idx = threadIdx.x; idx < nh * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    *reinterpret_cast<uint16_t*>(sB + (k >> 6) * (nh * 128) + sw128_offset(r, k & 63)) = B[(size_t)(rank * nh + r) * K + k];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc_pair(&tbase, 256); tmem_relinquish_pair(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = tbase;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, (uint32_t)N, a_mn ? 1u : 0u, 0u);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = a_mn ? umma_desc_mnmajor_sw128(smem_u32(sA) + ks * 2048, blk)
                               : umma_desc_kmajor_sw128(smem_u32(sA + (ks >> 2) * 16384)) + 2 * (ks & 3);
      const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(sB + (ks >> 2) * (nh * 128))) + 2 * (ks & 3);
      umma_bf16_ss_pair(tb, ad, bd, idesc, ks != 0);
    }
    umma_commit_pair(&bar, (uint16_t)0x3);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 4) {
    uint32_t v4[4];
    tmem_ld_32x4(taddr + c0, v4);
    tmem_ld_wait();
    for (int jx = 0; jx < 4; ++jx) Dout[(size_t)(rank * 128 + threadIdx.x) * N + c0 + jx] = __uint_as_float(v4[jx]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tb, 256); }
}

}}  // namespace c3d::fused
