// Inverse-CDF importance resampling ("sample_pdf") and merged fine-pass depths -- EXTENSION, off by default.
//
// The reference renders in a single pass (nerf_utils.py:172-218) and has no sample_pdf; BASELINE.json's north star asks
// for one ("a warp-level CDF plus binary search").  This restates the canonical NeRF hierarchical sampling step
// (oracle/nerf_oracle.py::importance_depths) for the SDF compositing weights of Render.volume_integration
// (nerf_utils.py:267-307).
//
// Bound: HBM.  Per ray the kernel reads z (4N) + weights or sdf (4N) [+ rays 24 B, + u 4K] and writes z_fine (4K),
// z_merged 4(N+K) and the fine-pass points 12(N+K).  Data path: a block owns chunks of RB consecutive rays; every
// input/output array of a chunk is one contiguous span, moved by the TMA engine as 1-D bulk copies
// (cp.async.bulk global->shared with mbarrier complete_tx; shared->global bulk groups), so global traffic is issued as
// a few large transactions per chunk instead of 96-byte rows per warp.  Compute: one warp per ray --
//   lanes = samples: density -> alpha -> transmittance (shuffle prefix product), interior weights + 1e-5 -> sum
//   (butterfly) -> CDF (shuffle prefix sum with carry for N > 32);
//   lanes = new samples: per-lane binary search of u in the CDF (upper bound), linear interpolation in the bin;
//   merge: rank of every coarse / new depth in the union by binary search in the other (sorted) list.
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

}  // namespace resample
}  // namespace c3d
