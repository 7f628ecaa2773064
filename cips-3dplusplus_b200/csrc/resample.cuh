// Inverse-CDF importance resampling ("sample_pdf") and merged fine-pass depths -- EXTENSION, off by default.
//
// The reference renders in a single pass (nerf_utils.py:172-218) and has no sample_pdf; BASELINE.json's north star asks
// for one ("a warp-level CDF plus binary search").  This restates the canonical NeRF hierarchical sampling step
// (oracle/nerf_oracle.py::importance_depths) for the SDF compositing weights of Render.volume_integration
// (nerf_utils.py:267-307).
//
// Bound: HBM.  Per ray the kernels read z (4N) + weights or sdf (4N) [+ rays 24 B, + u 4K] and write z_fine (4K),
// z_merged 4(N+K) and the fine-pass points 12(N+K).  Two kernels behind c3d_sample_pdf:
//
//  * sample_pdf_lane_kernel (below, second half; the default whenever the rows of >= 64 rays fit shared memory):
//    lanes = rays.
//  * sample_pdf_kernel (first half; larger N / K, or C3D_RESAMPLE=warp): lanes = samples.  A block owns chunks of RB
//    consecutive rays; every input / output array of a chunk is one contiguous span, moved by the TMA engine as 1-D bulk
//    copies (cp.async.bulk global->shared with mbarrier complete_tx; shared->global bulk groups; the next chunk's loads
//    are issued while the stores drain).  Compute, one warp per ray --
//      lanes = samples: density -> alpha -> transmittance (shuffle prefix product), interior weights + 1e-5 -> sum
//      (butterfly) -> CDF (shuffle prefix sum with carry for N > 32);
//      lanes = new samples: per-lane binary search of u in the CDF (upper bound), linear interpolation in the bin;
//      merge: rank of every coarse / new depth in the union by binary search in the other (sorted) list.
#pragma once
#include "c3d_common.cuh"
#include "sm100_ptx.cuh"
#include "kernels_aux.cuh"

namespace c3d {
namespace resample {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int MAX_N = 256;
constexpr int MAX_K = 256;
constexpr size_t SMEM_BUDGET = 100 * 1024;   // two blocks per SM: one loads / stores while the other computes

// shared-memory chunk layout (all offsets in floats from the dynamic base, each section 16-byte aligned)
struct Layout {
  int RB;                               // rays per chunk (multiple of 4: every span is a multiple of 16 bytes)
  int z, win, u, ro, rd, cdf, fine, merged, pts, total;
};
__host__ __device__ inline Layout make_layout(int N, int K, bool has_u, bool has_o, bool has_d, bool want_fine,
                                              bool want_merged, bool want_pts, int rb_override = 0) {
  const int per_ray = N + N + (has_u ? K : 0) + (has_o ? 3 : 0) + (has_d ? 3 : 0) + N /*cdf*/ + K /*fine*/ +
                      (N + K) /*merged, also scratch*/ + (want_pts ? 3 * (N + K) : 0);
  int RB = (int)((SMEM_BUDGET - 64) / ((size_t)per_ray * 4));
  RB = RB > 64 ? 64 : RB;
  RB &= ~7;                             // one ray per warp and pass: whole passes only
  if (RB < 8) RB = 8;
  if (rb_override > 0) RB = rb_override;
  Layout L;
  L.RB = RB;
  int o = 0;
  L.z = o;      o += RB * N;
  L.win = o;    o += RB * N;
  L.u = o;      o += has_u ? RB * K : 0;
  L.ro = o;     o += has_o ? RB * 3 : 0;
  L.rd = o;     o += has_d ? RB * 3 : 0;
  L.cdf = o;    o += RB * N;
  L.fine = o;   o += RB * K;
  L.merged = o; o += RB * (N + K);
  L.pts = o;    o += want_pts ? RB * 3 * (N + K) : 0;
  L.total = o;
  (void)want_fine; (void)want_merged;
  return L;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float y = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += y;
  }
  return v;
}
// number of elements of the ascending array a[0..n) that are <= x (upper bound) / < x (lower bound)
__device__ __forceinline__ int count_le(const float* a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int count_lt(const float* a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo;
}

// One ray, one warp.  z, w: the ray's N coarse depths and (weights | sdf) in shared memory; w is overwritten by the
// compositing weights when they are derived from the sdf.
__device__ __forceinline__ void resample_ray(int lane, int N, int K, const float* __restrict__ z, float* __restrict__ w,
                                             bool from_sdf, float inv_beta, float dnorm, const float* __restrict__ u,
                                             float* __restrict__ cdf, float* __restrict__ fine,
                                             float* __restrict__ merged, float* __restrict__ pts, float ox, float oy,
                                             float oz, float dx, float dy, float dz) {
  // ---- compositing weights from the sdf (nerf_utils.py:267-307), the same arithmetic as composite_fwd_kernel
  if (from_sdf) {
    float carry = 1.0f;
    for (int k0 = 0; k0 < N; k0 += 32) {
      const int k = k0 + lane;
      float one_minus = 1.0f, alpha = 0.f;
      if (k < N) {
        const float dist = (k + 1 < N ? z[k + 1] - z[k] : 1e10f) * dnorm;
        alpha = alpha_from_sdf<true>(w[k], inv_beta, dist);
        one_minus = 1.0f - alpha + 1e-10f;
      }
      float total;
      const float T = carry * warp_excl_prod(one_minus, lane, total);
      carry *= total;
      if (k < N) w[k] = alpha * T;
    }
    __syncwarp();
  }
  // ---- PDF over the N-1 mid-point bins from the interior weights w[1..N-2] (+1e-5), CDF with a leading zero
  const int M = N - 1;                       // CDF entries == bin edges
  float part = 0.f;
  for (int j = lane; j < N - 2; j += 32) part += w[j + 1] + 1e-5f;
  const float wsum = warp_sum(part);
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int j0 = 0; j0 < N - 2; j0 += 32) {
    const int j = j0 + lane;
    const float v = j < N - 2 ? (w[j + 1] + 1e-5f) / wsum : 0.f;
    const float s = warp_incl_sum(v, lane);
    if (j < N - 2) cdf[j + 1] = carry + s;
    carry += __shfl_sync(0xffffffffu, s, 31);
  }
  __syncwarp();
  // ---- inverse CDF: per-lane binary search + interpolation between the bin's mid-point edges
  bool sorted = true;
  const float ustep = K > 1 ? 1.0f / (float)(K - 1) : 0.f;
  for (int j = lane; j < K; j += 32) {
    const float uj = u ? u[j] : (j == K - 1 && K > 1 ? 1.0f : (float)j * ustep);
    const int inds = count_le(cdf, M, uj);
    const int below = inds - 1 > 0 ? inds - 1 : 0;
    const int above = inds < M - 1 ? inds : M - 1;
    const float cb = cdf[below], ca = cdf[above];
    const float bb = 0.5f * (z[below + 1] + z[below]), ba = 0.5f * (z[above + 1] + z[above]);
    float denom = ca - cb;
    denom = denom < 1e-5f ? 1.0f : denom;
    const float t = (uj - cb) / denom;
    fine[j] = bb + t * (ba - bb);
  }
  __syncwarp();
  if (merged == nullptr && pts == nullptr) return;
  for (int j = lane; j < K; j += 32) sorted = sorted && (j == 0 || fine[j - 1] <= fine[j]);
  sorted = __all_sync(0xffffffffu, sorted);
  // ---- ascending union: rank of each depth = own index + number of depths of the other list before it
  //      (ties: coarse first).  Unsorted new depths (random u in arbitrary order) are ranked by counting.
  for (int i = lane; i < N; i += 32) {
    const float v = z[i];
    int r;
    if (sorted) r = i + count_lt(fine, K, v);
    else { r = i; for (int j = 0; j < K; ++j) r += fine[j] < v; }
    if (merged) merged[r] = v;
    if (pts) { pts[3 * r] = fmaf(dx, v, ox); pts[3 * r + 1] = fmaf(dy, v, oy); pts[3 * r + 2] = fmaf(dz, v, oz); }
  }
  for (int j = lane; j < K; j += 32) {
    const float v = fine[j];
    int r = count_le(z, N, v);
    if (sorted) r += j;
    else for (int i = 0; i < K; ++i) r += (fine[i] < v) || (fine[i] == v && i < j);
    if (merged) merged[r] = v;
    if (pts) { pts[3 * r] = fmaf(dx, v, ox); pts[3 * r + 1] = fmaf(dy, v, oy); pts[3 * r + 2] = fmaf(dz, v, oz); }
  }
}

// cooperative copies for chunks the bulk engine cannot take (ragged tail, unaligned spans)
__device__ __forceinline__ void coop_load(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += THREADS) dst[i] = src[i];
}
__device__ __forceinline__ void coop_store(float* __restrict__ dst, const float* src, int n) {
  for (int i = threadIdx.x; i < n; i += THREADS) dst[i] = src[i];
}

__global__ void __launch_bounds__(THREADS) sample_pdf_kernel(c3d_resample_params p, Layout L, int n_chunks) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t full_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.n_samples, K = p.n_importance, NK = N + K, RB = L.RB;
  const bool from_sdf = p.weights == nullptr;
  const float* win = from_sdf ? p.sdf : p.weights;
  const float beta = p.sigmoid_beta_ptr ? *p.sigmoid_beta_ptr : p.sigmoid_beta;
  const float inv_beta = 1.0f / beta;
  const bool want_out = p.z_merged != nullptr || p.pts_merged != nullptr;

  if (threadIdx.x == 0) { ptx::mbar_init(&full_bar, 1); ptx::fence_mbar_init(); }
  __syncthreads();

  auto chunk_rays = [&](int c) { const long long r0 = (long long)c * RB; const long long rem = p.n_rays - r0; return (int)(rem < RB ? rem : RB); };
  // bulk copies need 16-byte sizes: full chunks always qualify (RB % 4 == 0), a ragged tail only when its spans do
  auto bulk_ok = [&](int nr) { return (nr & 3) == 0; };
  auto issue_loads = [&](int c) {          // thread 0
    const long long r0 = (long long)c * RB;
    const int nr = chunk_rays(c);
    uint32_t bytes = 2u * nr * N * 4u + (p.u ? nr * K * 4u : 0u) + (p.rays_o ? nr * 12u : 0u) + (p.rays_d ? nr * 12u : 0u);
    ptx::mbar_arrive_expect_tx(&full_bar, bytes);
    ptx::bulk_g2s(smem + L.z, p.z_vals + r0 * N, nr * N * 4u, &full_bar);
    ptx::bulk_g2s(smem + L.win, win + r0 * N, nr * N * 4u, &full_bar);
    if (p.u) ptx::bulk_g2s(smem + L.u, p.u + r0 * K, nr * K * 4u, &full_bar);
    if (p.rays_o) ptx::bulk_g2s(smem + L.ro, p.rays_o + r0 * 3, nr * 12u, &full_bar);
    if (p.rays_d) ptx::bulk_g2s(smem + L.rd, p.rays_d + r0 * 3, nr * 12u, &full_bar);
  };

  uint32_t phase = 0;
  int c = blockIdx.x;
  if (c < n_chunks && bulk_ok(chunk_rays(c)) && threadIdx.x == 0) issue_loads(c);
  for (; c < n_chunks; c += gridDim.x) {
    const long long r0 = (long long)c * RB;
    const int nr = chunk_rays(c);
    const bool bulk = bulk_ok(nr);
    if (bulk) {
      ptx::mbar_wait(&full_bar, phase);
      phase ^= 1;
      if (threadIdx.x == 0) ptx::bulk_wait_group_read0();    // the previous chunk's stores have drained the out buffers
    } else {
      if (threadIdx.x == 0) ptx::bulk_wait_group_read0();
      __syncthreads();
      coop_load(smem + L.z, p.z_vals + r0 * N, nr * N);
      coop_load(smem + L.win, win + r0 * N, nr * N);
      if (p.u) coop_load(smem + L.u, p.u + r0 * K, nr * K);
      if (p.rays_o) coop_load(smem + L.ro, p.rays_o + r0 * 3, nr * 3);
      if (p.rays_d) coop_load(smem + L.rd, p.rays_d + r0 * 3, nr * 3);
    }
    __syncthreads();
    for (int i = warp; i < nr; i += WARPS) {
      float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, dnorm = 1.f;
      if (p.rays_o) { ox = smem[L.ro + 3 * i]; oy = smem[L.ro + 3 * i + 1]; oz = smem[L.ro + 3 * i + 2]; }
      if (p.rays_d) {
        dx = smem[L.rd + 3 * i]; dy = smem[L.rd + 3 * i + 1]; dz = smem[L.rd + 3 * i + 2];
        dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
      }
      resample_ray(lane, N, K, smem + L.z + i * N, smem + L.win + i * N, from_sdf, inv_beta, dnorm,
                   p.u ? smem + L.u + i * K : nullptr, smem + L.cdf + i * N, smem + L.fine + i * K,
                   want_out ? smem + L.merged + i * NK : nullptr, p.pts_merged ? smem + L.pts + i * 3 * NK : nullptr,
                   ox, oy, oz, dx, dy, dz);
    }
    if (bulk) ptx::fence_proxy_async_smem();                  // generic-proxy writes -> visible to the bulk stores
    __syncthreads();
    const int cn = c + gridDim.x;
    if (bulk) {
      if (threadIdx.x == 0) {
        if (p.z_fine) ptx::bulk_s2g(p.z_fine + r0 * K, smem + L.fine, nr * K * 4u);
        if (p.z_merged) ptx::bulk_s2g(p.z_merged + r0 * NK, smem + L.merged, nr * NK * 4u);
        if (p.pts_merged) ptx::bulk_s2g(p.pts_merged + r0 * NK * 3, smem + L.pts, nr * NK * 12u);
        ptx::bulk_commit_group();
      }
    } else {
      if (p.z_fine) coop_store(p.z_fine + r0 * K, smem + L.fine, nr * K);
      if (p.z_merged) coop_store(p.z_merged + r0 * NK, smem + L.merged, nr * NK);
      if (p.pts_merged) coop_store(p.pts_merged + r0 * NK * 3, smem + L.pts, nr * NK * 3);
    }
    // inputs of this chunk are dead (barrier above): prefetch the next chunk while the stores drain
    if (cn < n_chunks && bulk_ok(chunk_rays(cn)) && threadIdx.x == 0) issue_loads(cn);
  }
  if (threadIdx.x == 0) ptx::bulk_wait_group0();
}


// ------------------------------------------------------------------------------------------
// Lane-per-ray variant (default when a ray's working set is small, e.g. the N = K = 24 configuration): with 24 samples
// a warp-per-ray scan leaves lanes idle and pays the shuffle / control overhead once per ray; here a thread owns a ray
// and runs the prefix product, the CDF, a fixed-length binary search and an in-place backward merge serially in shared
// memory (~2.5x fewer issued instructions per ray).  Rows are staged in shared memory with an ODD stride, so the 32 rays of a warp hit 32
// different banks; global traffic stays fully coalesced: the block's input / output spans are contiguous and are moved
// by flat 16-byte / 4-byte accesses, with (row, column) recovered by a multiply-high division.
// Summation order here is the reference's own (sequential cumprod / cumsum).
// ------------------------------------------------------------------------------------------
struct LaneLayout {
  int TPB;                      // rays (= threads) per block
  int Sz, Sk, Sm;               // odd row strides (floats) of the weight / CDF, new-depth and depth (merged) rows
  int zm, w, fine, geom, total;
  uint32_t magN, magK, magNK;   // ceil(2^32 / d): e / d == __umulhi(e, mag) for the e < 2^17 used here
};
__host__ inline LaneLayout make_lane_layout(int N, int K, bool want_pts, int tpb) {
  LaneLayout L;
  L.TPB = tpb;
  L.Sz = N | 1; L.Sk = K | 1; L.Sm = (N + K) | 1;
  int o = 0;
  L.zm = o;     o += tpb * L.Sm;            // coarse depths in the first N slots; the union is merged in place from the back
  L.w = o;      o += tpb * L.Sz;
  L.fine = o;   o += tpb * L.Sk;
  L.geom = o;   o += want_pts ? tpb * 6 : 0;
  L.total = o;
  auto mag = [](int d) { return (uint32_t)((0x100000000ull + (uint64_t)d - 1) / (uint64_t)d); };
  L.magN = mag(N); L.magK = mag(K); L.magNK = mag(N + K);
  return L;
}

// branch-free upper bound with a trip count that depends on n only: number of a[0..n) <= x
__device__ __forceinline__ int count_le_fixed(const float* __restrict__ a, int n, float x) {
  int base = 0, len = n;
  while (len > 1) {
    const int half = len >> 1;
    base += a[base + half - 1] <= x ? half : 0;
    len -= half;
  }
  return base + (a[base] <= x ? 1 : 0);
}

template <int TPB>
__global__ void __launch_bounds__(TPB) sample_pdf_lane_kernel(c3d_resample_params p, LaneLayout L) {
  extern __shared__ __align__(128) float smem[];
  const int tid = threadIdx.x;
  const int N = p.n_samples, K = p.n_importance, NK = N + K, M = N - 1;
  const long long r0 = (long long)blockIdx.x * TPB;
  const int nr = (int)((p.n_rays - r0) < TPB ? (p.n_rays - r0) : TPB);
  const bool from_sdf = p.weights == nullptr;
  const bool want_union = p.z_merged != nullptr || p.pts_merged != nullptr;
  const float* gw = (from_sdf ? p.sdf : p.weights) + r0 * N;
  const float* gz = p.z_vals + r0 * N;

  // ---- stage the block's rows: flat coalesced reads, scattered into odd-stride rows
  {
    const int n_in = nr * N;
    const int n4 = n_in >> 2;                                  // the span starts 16-byte aligned (r0 * N * 4, TPB % 4 == 0)
    const float4* gz4 = reinterpret_cast<const float4*>(gz);
    const float4* gw4 = reinterpret_cast<const float4*>(gw);
    for (int i = tid; i < n4; i += TPB) {
      const float4 a = __ldcs(gz4 + i), b = __ldcs(gw4 + i);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
      int row = (int)__umulhi((uint32_t)(4 * i), L.magN), col = 4 * i - row * N;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        smem[L.zm + row * L.Sm + col] = av[c];
        smem[L.w + row * L.Sz + col] = bv[c];
        if (++col == N) { col = 0; ++row; }
      }
    }
    for (int e = 4 * n4 + tid; e < n_in; e += TPB) {
      const int row = (int)__umulhi((uint32_t)e, L.magN), col = e - row * N;
      smem[L.zm + row * L.Sm + col] = gz[e];
      smem[L.w + row * L.Sz + col] = gw[e];
    }
    if (p.u) {                                                   // draws land in the rows the samples will replace
      const float* gu = p.u + r0 * K;
      for (int e = tid; e < nr * K; e += TPB) {
        const int row = (int)__umulhi((uint32_t)e, L.magK), col = e - row * K;
        smem[L.fine + row * L.Sk + col] = __ldcs(gu + e);
      }
    }
    if (p.pts_merged)
      for (int e = tid; e < nr * 3; e += TPB) {
        const int row = e / 3, c = e - 3 * row;
        smem[L.geom + row * 6 + c] = p.rays_o[r0 * 3 + e];
        smem[L.geom + row * 6 + 3 + c] = p.rays_d[r0 * 3 + e];
      }
  }
  __syncthreads();

  float* __restrict__ zr = smem + L.zm + tid * L.Sm;           // z[0..N), later the merged union [0..N+K)
  float* __restrict__ wr = smem + L.w + tid * L.Sz;            // weights, then cdf[j] in wr[j]
  float* __restrict__ fr = smem + L.fine + tid * L.Sk;         // u_j (when given) until sample j replaces it
  bool sorted = true;
  if (tid < nr) {
    // ---- compositing weights from the sdf (nerf_utils.py:267-307): sequential cumprod, the reference's order
    if (from_sdf) {
      const float beta = p.sigmoid_beta_ptr ? *p.sigmoid_beta_ptr : p.sigmoid_beta;
      const float inv_beta = 1.0f / beta;
      const float* d = p.rays_d + (r0 + tid) * 3;
      const float dnorm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      float T = 1.0f, znext = zr[0];
#pragma unroll 4
      for (int k = 0; k < N; ++k) {
        const float zk = znext;
        znext = k + 1 < N ? zr[k + 1] : 0.f;
        const float dist = (k + 1 < N ? znext - zk : 1e10f) * dnorm;
        const float alpha = alpha_from_sdf<true>(wr[k], inv_beta, dist);
        wr[k] = alpha * T;
        T *= 1.0f - alpha + 1e-10f;
      }
    }
    // ---- PDF of the interior weights (+1e-5), CDF with a leading zero written over the weights
    float wsum = 0.f;
    for (int j = 1; j <= N - 2; ++j) wsum += wr[j] + 1e-5f;
    float c = 0.f;
    wr[0] = 0.f;
#pragma unroll 4
    for (int j = 1; j <= N - 2; ++j) { c += (wr[j] + 1e-5f) / wsum; wr[j] = c; }
    // ---- inverse CDF: searchsorted(cdf, u, right=True) with a fixed trip count, linear interpolation in the bin
    const bool has_u = p.u != nullptr;
    const float ustep = K > 1 ? 1.0f / (float)(K - 1) : 0.f;
    float prev = -1e30f;
#pragma unroll 2
    for (int j = 0; j < K; ++j) {
      const float uj = has_u ? fr[j] : (j == K - 1 && K > 1 ? 1.0f : (float)j * ustep);
      const int inds = count_le_fixed(wr, M, uj);
      const int below = inds - 1 > 0 ? inds - 1 : 0;
      const int above = inds < M - 1 ? inds : M - 1;
      const float cb = wr[below], ca = wr[above];
      const float bb = 0.5f * (zr[below + 1] + zr[below]), ba = 0.5f * (zr[above + 1] + zr[above]);
      float denom = ca - cb;
      denom = denom < 1e-5f ? 1.0f : denom;
      const float v = bb + ((uj - cb) / denom) * (ba - bb);
      fr[j] = v;
      sorted = sorted && (prev <= v);
      prev = v;
    }
  }
  // Draws in arbitrary order give new depths in arbitrary order: z_fine keeps that order (written out first), then the
  // ray's new depths are sorted in place for the merge.  Block-uniform decision, the common (ascending) case skips it.
  bool fine_stored = false;
  if (want_union && __syncthreads_or(tid < nr && !sorted)) {
    if (p.z_fine) {
      float* out = p.z_fine + r0 * K;
      for (int e = tid; e < nr * K; e += TPB) {
        const int row = (int)__umulhi((uint32_t)e, L.magK), col = e - row * K;
        __stcs(out + e, smem[L.fine + row * L.Sk + col]);
      }
      fine_stored = true;
      __syncthreads();
    }
    if (tid < nr && !sorted)
      for (int j = 1; j < K; ++j) {                              // insertion sort of one small row
        const float v = fr[j];
        int i = j - 1;
        while (i >= 0 && fr[i] > v) { fr[i + 1] = fr[i]; --i; }
        fr[i + 1] = v;
      }
  }
  // ---- ascending union, merged from the back into the depth row (ties: coarse first, i.e. new depth taken first here)
  if (want_union && tid < nr) {
    const float NEG = -3.0e38f;
    int i = N - 1, j = K - 1;
    float zi = zr[i], fj = fr[j];
    for (int o = NK - 1; o >= 0; --o) {
      const bool tf = fj >= zi;
      zr[o] = tf ? fj : zi;
      j -= tf ? 1 : 0;
      i -= tf ? 0 : 1;
      const float fn = fr[j > 0 ? j : 0], zn = zr[i > 0 ? i : 0];
      fj = tf ? (j >= 0 ? fn : NEG) : fj;
      zi = tf ? zi : (i >= 0 ? zn : NEG);
    }
  }
  __syncthreads();

  // ---- flat coalesced writes of the block's output spans
  if (p.z_fine && !fine_stored) {
    float* out = p.z_fine + r0 * K;
    for (int e = tid; e < nr * K; e += TPB) {
      const int row = (int)__umulhi((uint32_t)e, L.magK), col = e - row * K;
      __stcs(out + e, smem[L.fine + row * L.Sk + col]);
    }
  }
  if (p.z_merged) {
    float* out = p.z_merged + r0 * NK;
    for (int e = tid; e < nr * NK; e += TPB) {
      const int row = (int)__umulhi((uint32_t)e, L.magNK), col = e - row * NK;
      __stcs(out + e, smem[L.zm + row * L.Sm + col]);
    }
  }
  if (p.pts_merged) {
    float* out = p.pts_merged + r0 * NK * 3;
    for (int e = tid; e < nr * NK; e += TPB) {                 // one point per thread: 12 contiguous bytes
      const int row = (int)__umulhi((uint32_t)e, L.magNK), col = e - row * NK;
      const float zv = smem[L.zm + row * L.Sm + col];
      const float* g = smem + L.geom + row * 6;
      __stcs(out + 3 * e + 0, fmaf(g[3], zv, g[0]));
      __stcs(out + 3 * e + 1, fmaf(g[4], zv, g[1]));
      __stcs(out + 3 * e + 2, fmaf(g[5], zv, g[2]));
    }
  }
}

}  // namespace resample
}  // namespace c3d
