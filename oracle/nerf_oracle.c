/* CPU restatement (plain C + OpenMP) of the CIPS-3D++ NeRF branch -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Same algorithm as oracle/nerf_oracle.py (which is pinned to vectors produced by the reference's own modules);
 * this file exists so that the CPU baseline timed next to the GPU path uses all host cores with a
 * cache-blocked loop nest instead of numpy temporaries.  tests/test_oracle_c.py checks it against the numpy
 * oracle and the golden vectors.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Reference lines restated (paths under exp/cips3d/):
 *   FiLM gamma/beta                 volume_renderer.py:66-67, 77-81
 *   FiLM-SIREN layers, heads        volume_renderer.py:70-85, 133-160
 *   normalize_points                nerf_utils.py:123-133
 *   volume_integration              nerf_utils.py:230-338
 *   ray / sample generation         nerf_utils.py:17-121, 135-170
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define W 256
#define PB 8 /* points per register block */

typedef struct {
  int D;
  /* per layer l in 0..D (l == D: view layer): transposed weight WT[k][c] (k < K_l), bias, FiLM linears (row-major) */
  const float* weight[17]; /* [0]: (256,3), hidden: (256,256), view: (256,259) -- reference layout (out,in) */
  const float* bias[17];
  const float* gamma_w[17];
  const float* gamma_b[17];
  const float* beta_w[17];
  const float* beta_b[17];
  const float* rgb_w; /* (3,256) */
  const float* rgb_b;
  const float* sigma_w; /* (1,256) */
  const float* sigma_b;
  const float* sigmoid_beta;
} oracle_params;

/* sin(x) for |x| < ~1e4: Cody-Waite reduction by pi, odd polynomial on [-pi/2, pi/2]; auto-vectorisable */
static inline float sin_poly(float x) {
  const float k = rintf(x * 0.31830988618379067f);
  float r = fmaf(k, -3.140625f, x);
  r = fmaf(k, -9.67502593994140625e-4f, r);
  r = fmaf(k, -1.509957990978376e-7f, r);
  const float r2 = r * r;
  float p = -2.5050759689e-8f;
  p = fmaf(p, r2, 2.7557314297e-6f);
  p = fmaf(p, r2, -1.9841270114e-4f);
  p = fmaf(p, r2, 8.3333337680e-3f);
  p = fmaf(p, r2, -1.6666667163e-1f);
  const float s = fmaf(r * r2, p, r);
  const int odd = ((int)k) & 1;
  return odd ? -s : s;
}
static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* out[p][c] = sum_k in[p][k] * WT[k][c], PB points at a time */
static void layer_gemm(const float* in, int ldin, const float* WT, int K, float* out /* [PB][W] */) {
  memset(out, 0, sizeof(float) * PB * W);
  for (int k = 0; k < K; ++k) {
    const float* w = WT + (size_t)k * W;
    float a[PB];
    for (int p = 0; p < PB; ++p) a[p] = in[(size_t)p * ldin + k];
    for (int p = 0; p < PB; ++p) {
      float* o = out + (size_t)p * W;
      const float ap = a[p];
#pragma omp simd
      for (int c = 0; c < W; ++c) o[c] = fmaf(ap, w[c], o[c]);
    }
  }
}

/* VolumeFeatureRenderer.forward (volume_renderer.py:192-283), reference tensor layouts */
int oracle_renderer_forward(const oracle_params* P, int b, int n_rays, int N, const float* pts, const float* rays_d,
                            const float* viewdirs, const float* z_vals, const float* near, const float* far,
                            const float* styles, float* rgb_map, float* feature_map, float* sdf_out, float* mask,
                            float* xyz, int nthreads) {
  const int D = P->D;
  if (D < 1 || D > 16 || N < 1 || N > 4096) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  /* transposed weights WT[l][k][c] */
  float* WT[17];
  int Kl[17];
  for (int l = 0; l <= D; ++l) {
    Kl[l] = l == 0 ? 3 : (l == D ? W + 3 : W);
    WT[l] = (float*)malloc(sizeof(float) * (size_t)Kl[l] * W);
    for (int c = 0; c < W; ++c)
      for (int k = 0; k < Kl[l]; ++k) WT[l][(size_t)k * W + c] = P->weight[l][(size_t)c * Kl[l] + k];
  }
  /* FiLM tables (b, D+1, 256): gamma = 15*(Gw s + gb) + 30, beta = 0.25*(Bw s + bb) */
  float* gam = (float*)malloc(sizeof(float) * (size_t)b * (D + 1) * W);
  float* bet = (float*)malloc(sizeof(float) * (size_t)b * (D + 1) * W);
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < b; ++i)
    for (int l = 0; l <= D; ++l) {
      const float* s = styles + ((size_t)i * (D + 1) + l) * W;
      for (int c = 0; c < W; ++c) {
        float g = 0.f, be = 0.f;
        const float* gw = P->gamma_w[l] + (size_t)c * W;
        const float* bw = P->beta_w[l] + (size_t)c * W;
#pragma omp simd reduction(+ : g, be)
        for (int k = 0; k < W; ++k) { g += s[k] * gw[k]; be += s[k] * bw[k]; }
        gam[((size_t)i * (D + 1) + l) * W + c] = 15.0f * (g + P->gamma_b[l][c]) + 30.0f;
        bet[((size_t)i * (D + 1) + l) * W + c] = 0.25f * (be + P->beta_b[l][c]);
      }
    }
  const float sbeta = P->sigmoid_beta[0];
  const long total_rays = (long)b * n_rays;
  const int NP = (N + PB - 1) / PB * PB;
#pragma omp parallel
  {
    float* h0 = (float*)aligned_alloc(64, sizeof(float) * (size_t)NP * (W + 8));
    float* h1 = (float*)aligned_alloc(64, sizeof(float) * (size_t)NP * (W + 8));
    float* acc = (float*)aligned_alloc(64, sizeof(float) * PB * W);
    float* sdf = (float*)malloc(sizeof(float) * NP);
    float* rgb = (float*)malloc(sizeof(float) * NP * 3);
    float* wts = (float*)malloc(sizeof(float) * NP);
    const int ld = W + 8;
#pragma omp for schedule(dynamic, 16)
    for (long r = 0; r < total_rays; ++r) {
      const int i = (int)(r / n_rays);
      const float nscale = 2.0f / (far[i] - near[i]);
      const float* g = gam + (size_t)i * (D + 1) * W;
      const float* be = bet + (size_t)i * (D + 1) * W;
      const float* P3 = pts + (size_t)r * N * 3;
      const float* vd = viewdirs + (size_t)r * 3;
      /* layer 0 input: normalised points, padded to NP rows */
      for (int p = 0; p < NP; ++p) {
        const int q = p < N ? p : N - 1;
        h0[(size_t)p * ld + 0] = P3[q * 3 + 0] * nscale;
        h0[(size_t)p * ld + 1] = P3[q * 3 + 1] * nscale;
        h0[(size_t)p * ld + 2] = P3[q * 3 + 2] * nscale;
      }
      float* in = h0;
      float* out = h1;
      for (int l = 0; l <= D; ++l) {
        if (l == D) { /* sdf head on h_{D-1}, then append the view direction (volume_renderer.py:148-152) */
          for (int p = 0; p < N; ++p) {
            float s = 0.f;
            const float* hp = in + (size_t)p * ld;
#pragma omp simd reduction(+ : s)
            for (int c = 0; c < W; ++c) s += hp[c] * P->sigma_w[c];
            sdf[p] = s + P->sigma_b[0];
          }
          for (int p = 0; p < NP; ++p) {
            in[(size_t)p * ld + W + 0] = vd[0]; in[(size_t)p * ld + W + 1] = vd[1]; in[(size_t)p * ld + W + 2] = vd[2];
          }
        }
        const float* gl = g + (size_t)l * W;
        const float* bl = be + (size_t)l * W;
        const float* bias = P->bias[l];
        for (int p0 = 0; p0 < NP; p0 += PB) {
          layer_gemm(in + (size_t)p0 * ld, ld, WT[l], Kl[l], acc);
          for (int p = 0; p < PB; ++p) {
            float* o = out + (size_t)(p0 + p) * ld;
            const float* a = acc + (size_t)p * W;
#pragma omp simd
            for (int c = 0; c < W; ++c) o[c] = sin_poly(fmaf(gl[c], a[c] + bias[c], bl[c]));
          }
        }
        float* t = in; in = out; out = t;
      }
      const float* feat = in; /* (N, 256) */
      for (int p = 0; p < N; ++p) {
        const float* f = feat + (size_t)p * ld;
        for (int j = 0; j < 3; ++j) {
          float s = 0.f;
          const float* w = P->rgb_w + (size_t)j * W;
#pragma omp simd reduction(+ : s)
          for (int c = 0; c < W; ++c) s += f[c] * w[c];
          rgb[p * 3 + j] = s + P->rgb_b[j];
        }
      }
      /* volume integration (nerf_utils.py:267-336) */
      const float* rd = rays_d + (size_t)r * 3;
      const float dn = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
      const float* z = z_vals + (size_t)r * N;
      float T = 1.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f, x = 0.f, y = 0.f, zz = 0.f;
      for (int p = 0; p < N; ++p) {
        const float dist = (p + 1 < N ? z[p + 1] - z[p] : 1e10f) * dn;
        const float sigma = sigmoidf_(-sdf[p] / sbeta) / sbeta;
        const float alpha = 1.0f - expf(-sigma * dist);
        const float w = alpha * T;
        T *= (1.0f - alpha + 1e-10f);
        wts[p] = w;
        c0 += w * sigmoidf_(rgb[p * 3 + 0]); c1 += w * sigmoidf_(rgb[p * 3 + 1]); c2 += w * sigmoidf_(rgb[p * 3 + 2]);
        x += w * P3[p * 3 + 0]; y += w * P3[p * 3 + 1]; zz += w * P3[p * 3 + 2];
        sdf_out[(size_t)r * N + p] = sdf[p];
      }
      rgb_map[r * 3 + 0] = -1.0f + 2.0f * c0; rgb_map[r * 3 + 1] = -1.0f + 2.0f * c1; rgb_map[r * 3 + 2] = -1.0f + 2.0f * c2;
      xyz[r * 3 + 0] = x; xyz[r * 3 + 1] = y; xyz[r * 3 + 2] = zz;
      mask[r * 2 + 0] = wts[N - 1];
      mask[r * 2 + 1] = -sqrtf(x * x + y * y + zz * zz);
      float* fm = feature_map + (size_t)r * W;
      memset(fm, 0, sizeof(float) * W);
      for (int p = 0; p < N; ++p) {
        const float w = wts[p];
        const float* f = feat + (size_t)p * ld;
#pragma omp simd
        for (int c = 0; c < W; ++c) fm[c] = fmaf(w, f[c], fm[c]);
      }
    }
    free(h0); free(h1); free(acc); free(sdf); free(rgb); free(wts);
  }
  for (int l = 0; l <= D; ++l) free(WT[l]);
  free(gam); free(bet);
  return 0;
}

/* Render.prepare_nerf_inputs (nerf_utils.py:172-218), unperturbed or with one offset per ray */
int oracle_prepare_inputs(int b, int img_size, int N, int static_viewdirs, const float* c2w, const float* focal,
                          const float* near, const float* far, const float* ray_offset, float* pts, float* rays_d,
                          float* viewdirs, float* z_vals) {
  const int hw = img_size * img_size;
#pragma omp parallel for schedule(static)
  for (long gid = 0; gid < (long)b * hw; ++gid) {
    const int i = (int)(gid / hw), ray = (int)(gid % hw);
    const int iy = ray / img_size, ix = ray % img_size;
    const float* M = c2w + (size_t)i * 12;
    const float f = focal[i], half = 0.5f * img_size;
    const float cx = ((float)ix + 0.5f - half) / f, cy = -((float)iy + 0.5f - half) / f, cz = -1.0f;
    const float dx = cx * M[0] + cy * M[1] + cz * M[2];
    const float dy = cx * M[4] + cy * M[5] + cz * M[6];
    const float dz = cx * M[8] + cy * M[9] + cz * M[10];
    rays_d[gid * 3 + 0] = dx; rays_d[gid * 3 + 1] = dy; rays_d[gid * 3 + 2] = dz;
    const float sx = static_viewdirs ? cx : dx, sy = static_viewdirs ? cy : dy, sz = static_viewdirs ? cz : dz;
    const float n = fmaxf(sqrtf(sx * sx + sy * sy + sz * sz), 1e-12f);
    viewdirs[gid * 3 + 0] = sx / n; viewdirs[gid * 3 + 1] = sy / n; viewdirs[gid * 3 + 2] = sz / n;
    const float u = ray_offset ? ray_offset[gid] : 0.f;
    const float step = (1.0f - 1.0f / N) / (N > 1 ? N - 1 : 1);
    for (int k = 0; k < N; ++k) {
      const float t = k * step, t1 = (k + 1) * step;
      float z = near[i] * (1.0f - t) + far[i] * t;
      if (ray_offset) {
        const float zu = k + 1 < N ? near[i] * (1.0f - t1) + far[i] * t1 : far[i];
        z = z + (zu - z) * u;
      }
      z_vals[gid * N + k] = z;
      pts[(gid * N + k) * 3 + 0] = M[3] + dx * z;
      pts[(gid * N + k) * 3 + 1] = M[7] + dy * z;
      pts[(gid * N + k) * 3 + 2] = M[11] + dz * z;
    }
  }
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
