// Small kernels that keep the flip-inversion step (projector_v9.py:998-1166) from being launch-bound: the step's big
// kernels take ~1 ms per 4 images, the ~130 PyTorch launches around them (camera set-up and its autograd, two gradient
// clippings, two Adam updates) as much again.
//
//   camera_kernel      Camera.generate_camera_params (`locations` mode, nerf_utils.py:369-378, 412-436): (azim, elev) ->
//                      camera-to-world 3x4, focal, near, far, and -- by forward-mode differentiation with two tangents --
//                      the Jacobian d pose / d (azim, elev), so the backward is one 12x2 contraction per camera.
//   adam_clip_kernel   torch.nn.utils.clip_grad_norm_ + torch.optim.Adam for a handful of small tensors in ONE launch
//                      (the latent group and the camera group of projector_v9.py:1016-1030, 1150-1153).
#pragma once
#include "c3d_common.cuh"

namespace c3d {
namespace invaux {

// value + two tangents (d/d azim, d/d elev)
struct D2 {
  float v, a, e;
};
__device__ __forceinline__ D2 mk(float v, float a = 0.f, float e = 0.f) { return D2{v, a, e}; }
__device__ __forceinline__ D2 operator+(D2 x, D2 y) { return mk(x.v + y.v, x.a + y.a, x.e + y.e); }
__device__ __forceinline__ D2 operator-(D2 x, D2 y) { return mk(x.v - y.v, x.a - y.a, x.e - y.e); }
__device__ __forceinline__ D2 operator*(D2 x, D2 y) { return mk(x.v * y.v, x.a * y.v + x.v * y.a, x.e * y.v + x.v * y.e); }
__device__ __forceinline__ D2 operator*(D2 x, float s) { return mk(x.v * s, x.a * s, x.e * s); }
struct V3 { D2 x, y, z; };
__device__ __forceinline__ V3 cross(V3 p, V3 q) {
  return V3{p.y * q.z - p.z * q.y, p.z * q.x - p.x * q.z, p.x * q.y - p.y * q.x};
}
// torch.nn.functional.normalize(v, dim=1, eps): v / max(|v|, eps); below eps the norm is a constant
__device__ __forceinline__ V3 normalize(V3 p, float eps) {
  const D2 n2 = p.x * p.x + p.y * p.y + p.z * p.z;
  const float n = sqrtf(n2.v);
  if (n < eps) { const float s = 1.0f / eps; return V3{p.x * s, p.y * s, p.z * s}; }
  const float inv = 1.0f / n;
  const D2 invn = mk(inv, -0.5f * n2.a * inv * inv * inv, -0.5f * n2.e * inv * inv * inv);   // d(1/sqrt(n2)) = -n2'/(2 n^3)
  return V3{p.x * invn, p.y * invn, p.z * invn};
}

// one thread per camera
__global__ void camera_kernel(const float* __restrict__ azim, const float* __restrict__ elev, int n, float img_size,
                              const float* __restrict__ fov_ang /* (n) degrees */, float fov_scalar, float dist_radius,
                              float* __restrict__ pose, float* __restrict__ focal, float* __restrict__ near,
                              float* __restrict__ far, float* __restrict__ jac /* (n,12,2) or NULL */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float az = azim[i], el = elev[i];
  const float sa = sinf(az), ca = cosf(az), se = sinf(el), ce = cosf(el);
  // camera on the unit sphere: dir = (cos e sin a, sin e, cos e cos a), loc = dist * dir, dist = 1
  const V3 dir{mk(ce * sa, ce * ca, -se * sa), mk(se, 0.f, ce), mk(ce * ca, -ce * sa, -se * ca)};
  const V3 up{mk(0.f), mk(1.f), mk(0.f)};
  const V3 zax = normalize(dir, 1e-5f);
  V3 xax = normalize(cross(up, zax), 1e-5f);
  const V3 yax = normalize(cross(zax, xax), 1e-5f);
  if (fabsf(xax.x.v) <= 5e-3f && fabsf(xax.y.v) <= 5e-3f && fabsf(xax.z.v) <= 5e-3f)       // degenerate-x fix (:428-431)
    xax = normalize(cross(yax, zax), 1e-5f);
  const D2 P[12] = {xax.x, yax.x, zax.x, dir.x, xax.y, yax.y, zax.y, dir.y, xax.z, yax.z, zax.z, dir.z};
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    pose[(size_t)i * 12 + k] = P[k].v;
    if (jac) { jac[((size_t)i * 12 + k) * 2 + 0] = P[k].a; jac[((size_t)i * 12 + k) * 2 + 1] = P[k].e; }
  }
  const float fov = (fov_ang ? fov_ang[i] : fov_scalar) * 3.14159265358979323846f / 180.0f;
  focal[i] = 0.5f * img_size / tanf(fov);
  near[i] = 1.0f - dist_radius;
  far[i] = 1.0f + dist_radius;
}

constexpr int ADAM_MAX_TENSORS = 8;
struct AdamArgs {
  int n_tensors;
  int group[ADAM_MAX_TENSORS];          // clipping / learning-rate group of each tensor (0 or 1)
  long long numel[ADAM_MAX_TENSORS];
  float* param[ADAM_MAX_TENSORS];
  const float* grad[ADAM_MAX_TENSORS];
  float* exp_avg[ADAM_MAX_TENSORS];
  float* exp_avg_sq[ADAM_MAX_TENSORS];
  const float* lr[2];                   // device scalars, one per group
  float* step;                          // device scalar: number of updates done so far (incremented here)
  float beta1, beta2, eps, max_norm;    // max_norm <= 0: no clipping
  float* grad_norm;                     // optional (2): the groups' gradient norms before clipping
};

// One block.  Pass 1: per-group L2 norm of the gradients (clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6))).
// Pass 2: Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) on the clipped gradients.
__global__ void __launch_bounds__(1024) adam_clip_kernel(AdamArgs a) {
  __shared__ float red[2][32];
  __shared__ float coef[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float s[2] = {0.f, 0.f};
  for (int t = 0; t < a.n_tensors; ++t) {
    const float* g = a.grad[t];
    float acc = 0.f;
    for (long long i = tid; i < a.numel[t]; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
    s[a.group[t]] += acc;
  }
#pragma unroll
  for (int gidx = 0; gidx < 2; ++gidx) {
    float v = s[gidx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[gidx][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int gidx = 0; gidx < 2; ++gidx) {
      float v = lane < (int)(blockDim.x >> 5) ? red[gidx][lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) {
        const float norm = sqrtf(v);
        coef[gidx] = a.max_norm > 0.f ? fminf(1.0f, a.max_norm / (norm + 1e-6f)) : 1.0f;
        if (a.grad_norm) a.grad_norm[gidx] = norm;
      }
    }
  }
  __syncthreads();
  const float step = *a.step + 1.0f;
  const float bc1 = 1.0f - powf(a.beta1, step), bc2 = 1.0f - powf(a.beta2, step);
  const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  for (int t = 0; t < a.n_tensors; ++t) {
    const float c = coef[a.group[t]];
    const float step_size = *a.lr[a.group[t]] / bc1;
    float* p = a.param[t]; float* m = a.exp_avg[t]; float* v = a.exp_avg_sq[t];
    const float* g = a.grad[t];
    for (long long i = tid; i < a.numel[t]; i += blockDim.x) {
      const float gi = g[i] * c;
      const float mi = a.beta1 * m[i] + (1.0f - a.beta1) * gi;
      const float vi = a.beta2 * v[i] + (1.0f - a.beta2) * gi * gi;
      m[i] = mi; v[i] = vi;
      p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + a.eps);
    }
  }
  __syncthreads();
  if (tid == 0) *a.step = step;
}

}  // namespace invaux
}  // namespace c3d
