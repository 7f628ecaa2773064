// libc3dpp.so -- C ABI entry points (see include/c3d_abi.h).  Validation + launches only; all math is in
// the kernel headers.  Nothing here allocates device memory or synchronises.
#include <stdlib.h>
#include <string.h>

#include "c3d_common.cuh"
#include "kernels_aux.cuh"
#include "mlp_fp32.cuh"
#include "fused_common.cuh"
#include "fused_bf16_sm100.cuh"
#include "fused_pair_sm100.cuh"
#include "mlp_tc32_sm100.cuh"
#include "backward.cuh"
#include "param_grads.cuh"
#include "eikonal.cuh"
#include "fused_bwd_sm100.cuh"
#include "resample.cuh"
#include "inversion_aux.cuh"

namespace c3d {
thread_local char g_err[512] = "";
thread_local int g_launches = 0;
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------------------------------
// Tuning options.  Read ONCE from the environment (first use), changed afterwards only through c3d_set_option -- no
// getenv on the launch path.
//   fwd       "pair" (default: CTA-pair kernel, fused_pair_sm100.cuh) | "v3" (single-CTA kernel, fused_bf16_sm100.cuh)
//   cluster   1 | 2 (default)   single-CTA kernels: weight multicast across a cluster of 2
//   grid      0 (default: one CTA per SM) | CTAs of the persistent kernels
//   egw       4 (default) | 8   single-CTA forward: epilogue warps per slot
//   bwd       "tc" (default) | "simt"   force the FP32-pipe backward
//   fp32      "tc" (default: fp32-mode MLP on the tensor cores, two-way fp16 split, mlp_tc32_sm100.cuh) | "simt" (FP32 pipe)
//   resample  "auto" (default) | "warp" | "lane";  resample_rb  rays per block of the warp kernel (0 = auto)
//   debug     bit mask for builds with -DC3D_KERNEL_PROF (ignored by release builds)
// ------------------------------------------------------------------------------------------
struct Options { int fwd_pair, cluster, grid, egw, bwd_simt, resample, resample_rb, debug, fp32_simt; };
static int parse_option(Options& o, const char* key, const char* val) {
  if (!key || !val) return -1;
  if (!strcmp(key, "fwd")) { if (!strcmp(val, "pair")) o.fwd_pair = 1; else if (!strcmp(val, "v3")) o.fwd_pair = 0; else return -1; }
  else if (!strcmp(key, "cluster")) { const int v = atoi(val); if (v != 1 && v != 2) return -1; o.cluster = v; }
  else if (!strcmp(key, "grid")) { const int v = atoi(val); if (v < 0) return -1; o.grid = v; }
  else if (!strcmp(key, "egw")) { const int v = atoi(val); if (v != 4 && v != 8) return -1; o.egw = v; }
  else if (!strcmp(key, "fp32")) { if (!strcmp(val, "simt")) o.fp32_simt = 1; else if (!strcmp(val, "tc")) o.fp32_simt = 0; else return -1; }
  else if (!strcmp(key, "bwd")) { if (!strcmp(val, "simt")) o.bwd_simt = 1; else if (!strcmp(val, "tc")) o.bwd_simt = 0; else return -1; }
  else if (!strcmp(key, "resample")) {
    if (!strcmp(val, "auto")) o.resample = 0; else if (!strcmp(val, "warp")) o.resample = 1; else if (!strcmp(val, "lane")) o.resample = 2; else return -1;
  }
  else if (!strcmp(key, "resample_rb")) { const int v = atoi(val); if (v < 0 || v % 4) return -1; o.resample_rb = v; }
  else if (!strcmp(key, "debug")) o.debug = atoi(val);
  else return -1;
  return 0;
}
static Options& options() {
  static Options o = [] {
    Options d{1, 2, 0, 4, 0, 0, 0, 0, 0};
    const char* keys[][2] = {{"C3D_FWD", "fwd"}, {"C3D_CLUSTER", "cluster"}, {"C3D_GRID", "grid"}, {"C3D_EGW", "egw"}, {"C3D_BWD", "bwd"},
                             {"C3D_FP32", "fp32"}, {"C3D_RESAMPLE", "resample"}, {"C3D_RESAMPLE_RB", "resample_rb"}, {"C3D_DEBUG", "debug"}};
    for (auto& k : keys) { const char* e = getenv(k[0]); if (e) parse_option(d, k[1], e); }
    return d;
  }();
  return o;
}

// per-device facts and one-time function attributes
constexpr int MAX_DEV = 64;
static int device_sms(int* dev_out = nullptr) {
  static int sms[MAX_DEV] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (dev_out) *dev_out = dev;
  int& n = sms[dev & (MAX_DEV - 1)];
  if (n == 0 && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize once per (call site, device): `slot` indexes the kernels sharing a call site
#define C3D_SMEM_ATTR(kern, bytes, slot, nslots)                                                              \
  do {                                                                                                        \
    static int done_[MAX_DEV][nslots] = {};                                                                   \
    int dev_ = 0;                                                                                             \
    device_sms(&dev_);                                                                                        \
    int& d_ = done_[dev_ & (MAX_DEV - 1)][slot];                                                              \
    if (d_ < (int)(bytes)) {                                                                                  \
      C3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));        \
      d_ = (int)(bytes);                                                                                      \
    }                                                                                                         \
  } while (0)
// density-only pass: none of the four map outputs is requested (only sdf / z_vals_out)
static bool fwd_sdf_only(const c3d_fwd_params* p) { return !p->gather && !p->rgb_map && !p->feature_map && !p->mask && !p->xyz; }

// the plain bf16 forward runs the CTA-pair kernel; the density-only pass (and the save-mode forward of the backward) run the
// single-CTA kernel, which has those modes
static bool fwd_uses_pair(const c3d_fwd_params* p) {
  if (fwd_sdf_only(p)) return false;
  return options().fwd_pair && p->mode == C3D_MODE_BF16 && p->n_samples >= fused::MIN_SAMPLES;
}

constexpr int PAIR_MAX_GRID = 160;          // CTAs of the persistent CTA-pair kernel (one per SM: 148 on B200)
constexpr int PAIR_MAX_UNIT_RAYS = 128;     // a unit holds at most 128 / gcd(N, 128) <= 128 rays once a batch shrinks it (forward_pair)
struct FwdWs {
  size_t film, first, view, wimg, kimg, scratch, chunk, total;
  int chunk_imgs;
  size_t c_feat, c_rgb, c_pts, c_rd, c_vd, c_z;   // offsets inside the fp32 chunk area
};
static FwdWs fwd_ws(const c3d_fwd_params* p) {
  FwdWs w;
  size_t o = 0;
  const size_t b = (size_t)p->batch;
  w.film = o;  o += align_up(b * (p->D + 1) * W * sizeof(float2), 256);
  w.first = o; o += align_up(b * W * sizeof(float4), 256);
  w.view = o;  o += align_up(b * W * sizeof(float4), 256);
  w.wimg = w.kimg = w.scratch = 0;
  if (fwd_uses_pair(p)) {
    w.wimg = o; o += align_up(b * p->D * pairk::WIMG_LAYER_BYTES, 1024);
    w.kimg = o; o += align_up(b * (p->D + 1) * pairk::KIMG_LAYER_BYTES, 1024);
    // channel-major feature_map: per slot one unit of rays in (ray, channel) order, transposed when the unit is complete
    w.scratch = o;
    if (p->feat_layout != C3D_FEAT_NHWC) o += align_up((size_t)2 * PAIR_MAX_GRID * PAIR_MAX_UNIT_RAYS * W * sizeof(float), 1024);
  }
  w.chunk = o;
  w.chunk_imgs = 0;
  w.c_feat = w.c_rgb = w.c_pts = w.c_rd = w.c_vd = w.c_z = 0;
  if (p->mode == C3D_MODE_FP32) {
    const size_t P = (size_t)p->n_rays * p->n_samples, R = (size_t)p->n_rays;
    const size_t per_img = align_up(P * W * 4, 256) + align_up(P * 3 * 4, 256) +
                           (p->input_kind == C3D_INPUT_POSES ? align_up(P * 12, 256) + 2 * align_up(R * 12, 256) + align_up(P * 4, 256) : 0);
    size_t ci = ((size_t)1 << 30) / per_img;
    if (ci < 1) ci = 1;
    if (ci > b) ci = b;
    w.chunk_imgs = (int)ci;
    size_t c = 0;
    w.c_feat = c; c += align_up(ci * P * W * 4, 256);
    w.c_rgb = c;  c += align_up(ci * P * 3 * 4, 256);
    if (p->input_kind == C3D_INPUT_POSES) {
      w.c_pts = c; c += align_up(ci * P * 12, 256);
      w.c_rd = c;  c += align_up(ci * R * 12, 256);
      w.c_vd = c;  c += align_up(ci * R * 12, 256);
      w.c_z = c;   c += align_up(ci * P * 4, 256);
    }
    o += c;
  }
  w.total = o;
  return w;
}

static int validate_fwd(const c3d_fwd_params* p) {
  C3D_CHECK_ARG(p != nullptr, "params is NULL");
  C3D_CHECK_ARG(p->abi_version == C3D_ABI_VERSION, "abi_version %d != library %d", p->abi_version, C3D_ABI_VERSION);
  C3D_CHECK_ARG(p->mode == C3D_MODE_FP32 || p->mode == C3D_MODE_BF16, "bad mode %d", p->mode);
  C3D_CHECK_ARG(p->input_kind == C3D_INPUT_POSES || p->input_kind == C3D_INPUT_POINTS, "bad input_kind %d", p->input_kind);
  C3D_CHECK_ARG(p->batch >= 1 && p->n_rays >= 1, "batch=%d n_rays=%d must be >= 1", p->batch, p->n_rays);
  C3D_CHECK_ARG(p->n_samples >= 2 && p->n_samples <= 256, "n_samples=%d outside [2,256]", p->n_samples);
  C3D_CHECK_ARG(p->D >= 1 && p->D <= C3D_MAX_LAYERS, "D=%d outside [1,%d]", p->D, C3D_MAX_LAYERS);
  C3D_CHECK_ARG((long long)p->batch * p->n_rays * p->n_samples < (1ll << 31), "batch*n_rays*n_samples overflows int32");
  C3D_CHECK_ARG(p->packed && p->styles && p->near && p->far, "packed/styles/near/far must be non-NULL");
  C3D_CHECK_ARG(p->sdf && ((p->rgb_map && p->feature_map && p->mask && p->xyz) || fwd_sdf_only(p) || p->gather),
                "outputs: sdf plus either all of rgb_map / feature_map / mask / xyz, or none of them (density-only pass)");
  if (p->gather) {
    const c3d_gather_out* g = p->gather;
    C3D_CHECK_ARG(fwd_uses_pair(p), "the fused all-gather is part of the bf16 CTA-pair forward kernel (MODE_BF16, n_samples >= %d, "
                                    "option fwd=pair)", fused::MIN_SAMPLES);
    C3D_CHECK_ARG(g->n_peers >= 1 && g->n_peers <= C3D_MAX_PEERS && g->image_offset >= 0, "gather: n_peers=%d image_offset=%d",
                  g->n_peers, g->image_offset);
    for (int i = 0; i < g->n_peers; ++i)
      C3D_CHECK_ARG(g->feature_map[i] && g->rgb_map[i] && g->mask[i] && g->xyz[i] && aligned16(g->feature_map[i]) &&
                    aligned16(g->rgb_map[i]) && aligned16(g->mask[i]) && aligned16(g->xyz[i]), "gather: destination %d is NULL or unaligned", i);
  }
  if (p->input_kind == C3D_INPUT_POSES) {
    C3D_CHECK_ARG(p->cam_poses && p->focal, "POSES input needs cam_poses and focal");
    C3D_CHECK_ARG(p->img_size >= 1 && p->n_rays == p->img_size * p->img_size, "n_rays=%d != img_size^2 (%d)", p->n_rays, p->img_size);
  } else {
    C3D_CHECK_ARG(p->pts && p->rays_d && p->viewdirs && p->z_vals, "POINTS input needs pts, rays_d, viewdirs, z_vals");
  }
  const void* ptrs[] = {p->packed, p->styles, p->cam_poses, p->pts, p->rays_d, p->viewdirs, p->z_vals, p->rgb_map,
                        p->feature_map, p->sdf, p->mask, p->xyz, p->workspace};
  for (const void* q : ptrs) C3D_CHECK_ARG(aligned16(q), "pointer %p is not 16-byte aligned", q);
  C3D_CHECK_ARG(p->feat_layout == C3D_FEAT_NHWC || p->feat_layout == C3D_FEAT_NCHW || p->feat_layout == C3D_FEAT_NCHW_BF16,
                "bad feat_layout %d", p->feat_layout);
  C3D_CHECK_ARG(p->feat_layout != C3D_FEAT_NCHW_BF16 || (p->mode == C3D_MODE_BF16 && p->n_samples >= fused::MIN_SAMPLES),
                "bf16 feature maps come from the tensor-core kernels only (MODE_BF16, n_samples >= %d)", fused::MIN_SAMPLES);
  const FwdWs w = fwd_ws(p);
  C3D_CHECK_ARG(p->workspace && p->workspace_bytes >= w.total, "workspace too small: %zu < %zu", p->workspace_bytes, w.total);
  return C3D_OK;
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int hw) {
  // in (b, hw, 256) -> out (b, 256, hw); 32x32 tiles through shared memory
  __shared__ float tile[32][33];
  const int b = blockIdx.z, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + i < hw) tile[i][threadIdx.x] = in[((size_t)b * hw + r0 + i) * W + c0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + threadIdx.x < hw) out[((size_t)b * W + c0 + i) * hw + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

static int launch_style_prep(const void* packed, int D, const float* styles, int batch, float* film, float* first,
                             float* view, cudaStream_t st) {
  const PackedLayout L = packed_layout(D);
  dim3 grid(D + 1, (batch + SP_IMGS - 1) / SP_IMGS);
  style_prep_kernel<<<grid, SP_THREADS, 0, st>>>(reinterpret_cast<const uint8_t*>(packed), L, styles, batch,
                                          reinterpret_cast<float2*>(film), reinterpret_cast<float4*>(first),
                                          reinterpret_cast<float4*>(view));
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

static int gcd_(int a, int b) { return b ? gcd_(b, a % b) : a; }

// Work unit = whole rays whose points fill whole 128-row tiles when possible (u0 rays); about 6 tiles per unit for
// large batches, fewer when that would leave slots of the persistent grid without work (small batches).
static int unit_rays_for(const c3d_fwd_params* p) {
  const int nsm = device_sms();
  const int u0 = 128 / gcd_(p->n_samples, 128);
  int ur = u0;
  while (ur * p->n_samples < 6 * 128) ur += u0;
  const long long want = 4ll * 2 * nsm;                       // >= 4 units per slot before units grow beyond u0
  while (ur > u0 && (long long)p->batch * ((p->n_rays + ur - 1) / ur) < want) ur -= u0;
  if (ur > p->n_rays) ur = p->n_rays;
  return ur;
}

static void fused_fill_args(fused::Args& a, const c3d_fwd_params* p, const float2* film, const float4* first,
                            const float4* view) {
  memset(&a, 0, sizeof(a));
  a.blob = reinterpret_cast<const uint8_t*>(p->packed);
  a.L = packed_layout(p->D);
  a.film = film; a.first = first; a.view = view;
  a.batch = p->batch; a.n_rays = p->n_rays; a.n_samples = p->n_samples; a.D = p->D;
  a.img_size = p->img_size; a.static_viewdirs = p->static_viewdirs; a.input_kind = p->input_kind;
  a.unit_rays = unit_rays_for(p);
  const int ur = a.unit_rays;
  a.units_per_img = (p->n_rays + ur - 1) / ur;
  a.tiles_per_unit = (ur * p->n_samples + fused::TILE - 1) / fused::TILE;
  a.n_tiles_g = (long long)a.batch * a.units_per_img * a.tiles_per_unit;
  a.cam_poses = p->cam_poses; a.focal = p->focal; a.near = p->near; a.far = p->far; a.ray_offset = p->ray_offset;
  a.pts = p->pts; a.rays_d = p->rays_d; a.viewdirs = p->viewdirs; a.z_vals = p->z_vals;
  a.rgb_map = p->rgb_map; a.feature_map = p->feature_map; a.sdf = p->sdf; a.mask = p->mask; a.xyz = p->xyz;
  a.z_vals_out = p->z_vals_out;
  a.sdf_only = fwd_sdf_only(p) ? 1 : 0;
  a.feat_nchw = p->feat_layout == C3D_FEAT_NCHW ? 1 : (p->feat_layout == C3D_FEAT_NCHW_BF16 ? 2 : 0);
  if (p->gather) {
    a.n_peers = p->gather->n_peers; a.gather_off = p->gather->image_offset;
    for (int i = 0; i < a.n_peers; ++i) {
      a.peer_feat[i] = p->gather->feature_map[i]; a.peer_rgb[i] = p->gather->rgb_map[i];
      a.peer_mask[i] = p->gather->mask[i]; a.peer_xyz[i] = p->gather->xyz[i];
    }
  }
  a.debug = options().debug;
}

// kind: 0 forward, 1 forward + save (backward support), 2 backward
static int fused_launch(const fused::Args& a, int kind, cudaStream_t st) {
  const int nsm = device_sms();
  const int cluster = options().cluster;
  const long long total_units = (long long)a.batch * a.units_per_img;
  int grid = nsm;
  if ((long long)grid * 2 > total_units) grid = (int)((total_units + 1) / 2);
  if (grid < 1) grid = 1;
  if (cluster == 2) grid = (grid + 1) & ~1;
  if (options().grid > 0) { grid = options().grid; if (cluster == 2) grid = (grid + 1) & ~1; }
  const int egw = (kind == 0 && options().egw == 8) ? 8 : 4;   // forward: epilogue warps per slot (tuning knob; 4 measured best)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kind == 2 ? fusedbwd::NTHREADS : fused::nthreads(egw));
  cfg.dynamicSmemBytes = kind == 2 ? fusedbwd::SMEM_BYTES : fused::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  void (*kern)(const fused::Args);
  int variant;                                       // index of `kern` among the eight kernels of this call site
  if (kind == 2) { kern = cluster == 2 ? fusedbwd::fused_backward_kernel<2> : fusedbwd::fused_backward_kernel<1>; variant = 0; }
  else if (kind == 1) { kern = cluster == 2 ? fused::fused_forward_kernel<2, 4, true> : fused::fused_forward_kernel<1, 4, true>; variant = 2; }
  else if (egw == 8) { kern = cluster == 2 ? fused::fused_forward_kernel<2, 8> : fused::fused_forward_kernel<1, 8>; variant = 4; }
  else { kern = cluster == 2 ? fused::fused_forward_kernel<2, 4> : fused::fused_forward_kernel<1, 4>; variant = 6; }
  variant += cluster == 2 ? 1 : 0;
  C3D_SMEM_ATTR(kern, cfg.dynamicSmemBytes, variant, 8);
  C3D_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

// CTA-pair forward: per-image weight images, then one persistent launch of 2-CTA clusters (one CTA per SM).
static int forward_pair(const c3d_fwd_params* p, const FwdWs& w, cudaStream_t st) {
  const int nsm = device_sms();
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  fused::Args a;
  fused_fill_args(a, p, reinterpret_cast<const float2*>(ws + w.film), reinterpret_cast<const float4*>(ws + w.first),
                  reinterpret_cast<const float4*>(ws + w.view));
  // pair-unit = 2 * unit_rays rays of one image; shrink units while the persistent grid would be short of work
  const int u0 = 128 / gcd_(p->n_samples, 128);
  int ur = u0;
  while (ur * p->n_samples < 6 * 128) ur += u0;
  while (ur > u0 && (long long)p->batch * ((p->n_rays + 2 * ur - 1) / (2 * ur)) < 4ll * nsm) ur -= u0;
  if (ur > p->n_rays) ur = p->n_rays;
  a.unit_rays = ur;
  a.units_per_img = (p->n_rays + 2 * ur - 1) / (2 * ur);
  a.wimg = ws + w.wimg; a.kimg = ws + w.kimg;
  a.feat_scratch = reinterpret_cast<float*>(ws + w.scratch);
  C3D_CHECK_ARG(ur <= PAIR_MAX_UNIT_RAYS, "unit of %d rays exceeds the scratch layout", ur);
  film_weights_kernel<<<dim3(8, p->D, p->batch), 256, 0, st>>>(a.blob, a.L, a.film, ws + w.wimg);
  C3D_LAUNCH_CHECK();
  film_k16_kernel<<<dim3(p->D + 1, p->batch), 256, 0, st>>>(a.blob, a.L, a.film, ws + w.kimg);
  C3D_LAUNCH_CHECK();
  const long long total_pu = (long long)p->batch * a.units_per_img;
  int grid = nsm & ~1;
  if ((long long)grid > ((total_pu + 1) & ~1ll)) grid = (int)((total_pu + 1) & ~1ll);
  if (options().grid > 0) grid = (options().grid + 1) & ~1;
  if (grid > PAIR_MAX_GRID) grid = PAIR_MAX_GRID;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(pairk::NTHREADS); cfg.dynamicSmemBytes = pairk::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  void (*kern)(const fused::Args) = pairk::fused_forward_pair_kernel;
  C3D_SMEM_ATTR(kern, pairk::SMEM_BYTES, 0, 1);
  C3D_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

static int forward_bf16(const c3d_fwd_params* p, const FwdWs& w, cudaStream_t st) {
  C3D_CHECK_ARG(p->n_samples >= fused::MIN_SAMPLES, "bf16 mode needs n_samples >= %d (got %d); use C3D_MODE_FP32",
                fused::MIN_SAMPLES, p->n_samples);
  if (fwd_uses_pair(p)) return forward_pair(p, w, st);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  fused::Args a;
  fused_fill_args(a, p, reinterpret_cast<const float2*>(ws + w.film), reinterpret_cast<const float4*>(ws + w.first),
                  reinterpret_cast<const float4*>(ws + w.view));
  return fused_launch(a, 0, st);
}

// the fp32-mode point MLP: tensor cores (two-way fp16 split, clusters of two CTAs) unless the option asks for the FP32 pipe
static int launch_mlp_fp32(MlpF32Args& m, int n_imgs, cudaStream_t st) {
  m.n_imgs = n_imgs;
  if (options().fp32_simt) {
    C3D_SMEM_ATTR(mlp_fp32_kernel, F32_SMEM, 0, 1);
    mlp_fp32_kernel<<<(unsigned)(n_imgs * m.tiles_per_img), 256, F32_SMEM, st>>>(m);
    C3D_LAUNCH_CHECK();
    return C3D_OK;
  }
  const long long pair_tiles = (long long)n_imgs * ((m.pts_per_img + 2 * tc32::TILE - 1) / (2 * tc32::TILE));
  int grid = device_sms() & ~1;
  if (pair_tiles * 2 < grid) grid = (int)(pair_tiles * 2);
  if (options().grid > 0) grid = (options().grid + 1) & ~1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc32::NTHREADS); cfg.dynamicSmemBytes = tc32::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  void (*kern)(const MlpF32Args) = tc32::mlp_tc32_kernel;
  C3D_SMEM_ATTR(kern, tc32::SMEM_BYTES, 0, 1);
  C3D_CUDA(cudaLaunchKernelEx(&cfg, kern, m));
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

static int forward_fp32(const c3d_fwd_params* p, const FwdWs& w, cudaStream_t st, float* feat_out) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  uint8_t* ck = ws + w.chunk;
  const size_t P = (size_t)p->n_rays * p->n_samples, R = (size_t)p->n_rays;
  for (int i0 = 0; i0 < p->batch; i0 += w.chunk_imgs) {
    const int ni = (p->batch - i0 < w.chunk_imgs) ? p->batch - i0 : w.chunk_imgs;
    const float *pts, *rays_d, *viewdirs, *z_vals;
    if (p->input_kind == C3D_INPUT_POSES) {
      c3d_raygen_params rg;
      memset(&rg, 0, sizeof(rg));
      rg.batch = ni; rg.img_size = p->img_size; rg.n_samples = p->n_samples; rg.static_viewdirs = p->static_viewdirs;
      rg.cam_poses = p->cam_poses + (size_t)i0 * 12; rg.focal = p->focal + i0; rg.near = p->near + i0; rg.far = p->far + i0;
      rg.ray_offset = p->ray_offset ? p->ray_offset + (size_t)i0 * R : nullptr;
      rg.pts = reinterpret_cast<float*>(ck + w.c_pts); rg.rays_d = reinterpret_cast<float*>(ck + w.c_rd);
      rg.viewdirs = reinterpret_cast<float*>(ck + w.c_vd); rg.z_vals = reinterpret_cast<float*>(ck + w.c_z);
      const long long n = (long long)ni * R * p->n_samples;
      raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rg);
      C3D_LAUNCH_CHECK();
      pts = rg.pts; rays_d = rg.rays_d; viewdirs = rg.viewdirs; z_vals = rg.z_vals;
      if (p->z_vals_out)
        C3D_CUDA(cudaMemcpyAsync(p->z_vals_out + (size_t)i0 * P, z_vals, (size_t)ni * P * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      pts = p->pts + (size_t)i0 * P * 3; rays_d = p->rays_d + (size_t)i0 * R * 3;
      viewdirs = p->viewdirs + (size_t)i0 * R * 3; z_vals = p->z_vals + (size_t)i0 * P;
    }
    MlpF32Args m;
    m.blob = reinterpret_cast<const uint8_t*>(p->packed);
    m.L = packed_layout(p->D);
    m.film = reinterpret_cast<const float2*>(ws + w.film) + (size_t)i0 * (p->D + 1) * W;
    m.first = reinterpret_cast<const float4*>(ws + w.first) + (size_t)i0 * W;
    m.view = reinterpret_cast<const float4*>(ws + w.view) + (size_t)i0 * W;
    m.pts = pts; m.viewdirs = viewdirs; m.near = p->near + i0; m.far = p->far + i0;
    m.n_samples = p->n_samples; m.pts_per_img = (int)P; m.tiles_per_img = (int)((P + F32_TP - 1) / F32_TP);
    m.feat = reinterpret_cast<float*>(ck + w.c_feat); m.rgb = reinterpret_cast<float*>(ck + w.c_rgb);
    m.sdf = p->sdf + (size_t)i0 * P;
    m.save_acc = nullptr; m.save_stride = 0;
    { const int rc = launch_mlp_fp32(m, ni, st); if (rc != C3D_OK) return rc; }
    if (fwd_sdf_only(p)) continue;                 // density-only pass: no compositing
    c3d_composite_params c;
    memset(&c, 0, sizeof(c));
    c.n_rays = (int64_t)ni * R; c.n_samples = p->n_samples; c.n_feat = W;
    c.sigmoid_beta_ptr = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p->packed) + m.L.scal) + 4;
    c.rgb = m.rgb; c.sdf = m.sdf; c.features = m.feat; c.z_vals = z_vals; c.rays_d = rays_d; c.pts = pts;
    c.rgb_map = p->rgb_map + (size_t)i0 * R * 3; c.feature_map = feat_out + (size_t)i0 * R * W;
    c.xyz = p->xyz + (size_t)i0 * R * 3; c.mask = p->mask + (size_t)i0 * R * 2;
    composite_fwd_kernel<<<(unsigned)((c.n_rays + 7) / 8), 256, 0, st>>>(c);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

}  // namespace c3d

using namespace c3d;

extern "C" {

int c3d_abi_version(void) { return C3D_ABI_VERSION; }

int c3d_set_option(const char* key, const char* value) {
  Options o = options();
  if (parse_option(o, key, value) != 0) return fail(C3D_ERR_ARG, "c3d_set_option: unknown option or value %s=%s", key ? key : "(null)", value ? value : "(null)");
  options() = o;
  return C3D_OK;
}
const char* c3d_last_error(void) { return g_err; }
int c3d_last_launch_count(void) { return g_launches; }

size_t c3d_packed_bytes(int32_t D) {
  if (D < 1 || D > C3D_MAX_LAYERS) return 0;
  return packed_layout(D).total;
}

int c3d_pack_weights(const c3d_raw_params* raw, void* packed, size_t packed_bytes, c3d_stream_t stream) {
  C3D_CHECK_ARG(raw && packed, "raw/packed is NULL");
  C3D_CHECK_ARG(raw->D >= 1 && raw->D <= C3D_MAX_LAYERS, "D=%d outside [1,%d]", raw->D, C3D_MAX_LAYERS);
  const PackedLayout L = packed_layout(raw->D);
  C3D_CHECK_ARG(packed_bytes >= L.total, "packed buffer too small: %zu < %zu", packed_bytes, L.total);
  C3D_CHECK_ARG(aligned16(packed), "packed buffer must be 16-byte aligned");
  for (int l = 0; l < raw->D; ++l)
    C3D_CHECK_ARG(raw->pts_weight[l] && raw->pts_bias[l] && raw->pts_gamma_weight[l] && raw->pts_gamma_bias[l] &&
                      raw->pts_beta_weight[l] && raw->pts_beta_bias[l], "pts layer %d has a NULL tensor", l);
  C3D_CHECK_ARG(raw->views_weight && raw->views_bias && raw->views_gamma_weight && raw->views_gamma_bias &&
                    raw->views_beta_weight && raw->views_beta_bias && raw->rgb_weight && raw->rgb_bias &&
                    raw->sigma_weight && raw->sigma_bias && raw->sigmoid_beta, "a head/view tensor is NULL");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  pack_small_kernel<<<1, 256, 0, st>>>(*raw, reinterpret_cast<uint8_t*>(packed), L);
  C3D_LAUNCH_CHECK();
  pack_matrix_kernel<<<dim3(raw->D + 1, W / 32, 3), dim3(32, 8), 0, st>>>(*raw, reinterpret_cast<uint8_t*>(packed), L);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

size_t c3d_workspace_bytes(const c3d_fwd_params* p) {
  if (!p || p->batch < 1 || p->n_rays < 1 || p->n_samples < 1 || p->D < 1) return 0;
  FwdWs w = fwd_ws(p);
  // (b, 256, hw) features: the tensor-core kernels write them directly; only the fp32 parity path stages + transposes
  size_t extra = (p->feat_layout == C3D_FEAT_NCHW && p->mode == C3D_MODE_FP32) ? align_up((size_t)p->batch * p->n_rays * W * 4, 256) : 0;
  return w.total + extra;
}

int c3d_style_prep(const void* packed, int32_t D, const float* styles, int32_t batch, float* film, float* first,
                     float* view, c3d_stream_t stream) {
  C3D_CHECK_ARG(packed && styles && film && first && view && batch >= 1, "NULL argument or batch < 1");
  C3D_CHECK_ARG(D >= 1 && D <= C3D_MAX_LAYERS, "D=%d outside [1,%d]", D, C3D_MAX_LAYERS);
  return launch_style_prep(packed, D, styles, batch, film, first, view, reinterpret_cast<cudaStream_t>(stream));
}

int c3d_nerf_forward(const c3d_fwd_params* p, c3d_stream_t stream) {
  g_launches = 0;
  int rc = validate_fwd(p);
  if (rc != C3D_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const FwdWs w = fwd_ws(p);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  rc = launch_style_prep(p->packed, p->D, p->styles, p->batch, reinterpret_cast<float*>(ws + w.film),
                         reinterpret_cast<float*>(ws + w.first), reinterpret_cast<float*>(ws + w.view), st);
  if (rc != C3D_OK) return rc;
  float* feat_out = p->feature_map;
  // channel-major features come straight from the compositing epilogue of the tensor-core kernels (no extra launch, no
  // extra bytes); the fp32 parity path composites one warp per ray and goes through a staging buffer + transposition
  const bool nchw = p->feat_layout == C3D_FEAT_NCHW && !fwd_sdf_only(p) && p->mode == C3D_MODE_FP32;
  if (nchw) {
    C3D_CHECK_ARG(p->workspace_bytes >= c3d_workspace_bytes(p), "workspace too small for NCHW staging");
    feat_out = reinterpret_cast<float*>(ws + w.total);
  }
  c3d_fwd_params q = *p;
  q.feature_map = feat_out;
  rc = (p->mode == C3D_MODE_BF16) ? forward_bf16(&q, w, st) : forward_fp32(&q, w, st, feat_out);
  if (rc != C3D_OK) return rc;
  if (nchw) {
    dim3 grid((p->n_rays + 31) / 32, W / 32, p->batch);
    nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, st>>>(feat_out, p->feature_map, p->n_rays);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

int c3d_raygen(const c3d_raygen_params* p, c3d_stream_t stream) {
  C3D_CHECK_ARG(p && p->batch >= 1 && p->img_size >= 1 && p->n_samples >= 2, "bad raygen sizes");
  C3D_CHECK_ARG(p->cam_poses && p->focal && p->near && p->far, "cam_poses/focal/near/far must be non-NULL");
  const long long n = (long long)p->batch * p->img_size * p->img_size * p->n_samples;
  C3D_CHECK_ARG(n < (1ll << 39), "too many sample points");
  raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

int c3d_composite_forward(const c3d_composite_params* p, c3d_stream_t stream) {
  C3D_CHECK_ARG(p && p->n_rays >= 1, "n_rays must be >= 1");
  C3D_CHECK_ARG(p->n_samples >= 1 && p->n_samples <= CMP_MAX_N, "n_samples=%d outside [1,%d]", p->n_samples, CMP_MAX_N);
  C3D_CHECK_ARG(p->n_feat >= 0 && p->n_feat % 4 == 0, "n_feat=%d must be a multiple of 4", p->n_feat);
  C3D_CHECK_ARG(p->rgb && p->sdf && p->z_vals && p->rays_d && p->pts, "rgb/sdf/z_vals/rays_d/pts must be non-NULL");
  C3D_CHECK_ARG(p->rgb_map && p->xyz && p->mask, "rgb_map/xyz/mask outputs must be non-NULL");
  C3D_CHECK_ARG(!p->features || p->feature_map, "feature_map output missing");
  C3D_CHECK_ARG(aligned16(p->features) && aligned16(p->feature_map), "features must be 16-byte aligned");
  C3D_CHECK_ARG((p->flags & ~3) == 0, "unknown composite flags 0x%x", p->flags);
  C3D_CHECK_ARG((p->flags & C3D_COMPOSITE_RAW_DENSITY) || p->sigmoid_beta_ptr || p->sigmoid_beta > 0.f, "sigmoid_beta must be > 0");
  composite_fwd_kernel<<<(unsigned)((p->n_rays + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

int c3d_sample_pdf(const c3d_resample_params* p, c3d_stream_t stream) {
  C3D_CHECK_ARG(p && p->n_rays >= 1, "n_rays must be >= 1");
  C3D_CHECK_ARG(p->n_samples >= 3 && p->n_samples <= resample::MAX_N, "n_samples=%d outside [3,%d]", p->n_samples,
                resample::MAX_N);
  C3D_CHECK_ARG(p->n_importance >= 1 && p->n_importance <= resample::MAX_K, "n_importance=%d outside [1,%d]",
                p->n_importance, resample::MAX_K);
  C3D_CHECK_ARG(p->z_vals, "z_vals must be non-NULL");
  C3D_CHECK_ARG(p->weights || (p->sdf && p->rays_d), "either weights, or sdf and rays_d, must be given");
  C3D_CHECK_ARG(p->weights || p->sigmoid_beta_ptr || p->sigmoid_beta > 0.f, "sigmoid_beta must be > 0");
  C3D_CHECK_ARG(p->z_fine || p->z_merged || p->pts_merged, "no output requested");
  C3D_CHECK_ARG(!p->pts_merged || (p->rays_o && p->rays_d), "pts_merged needs rays_o and rays_d");
  C3D_CHECK_ARG(aligned16(p->z_vals) && aligned16(p->weights) && aligned16(p->sdf) && aligned16(p->rays_d) &&
                aligned16(p->rays_o) && aligned16(p->u) && aligned16(p->z_fine) && aligned16(p->z_merged) &&
                aligned16(p->pts_merged), "all pointers must be 16-byte aligned");
  c3d_resample_params q = *p;
  if (q.weights) q.sdf = nullptr;
  if (!q.pts_merged) q.rays_o = nullptr;
  if (q.weights && !q.pts_merged) q.rays_d = nullptr;
  // lanes = rays when 128 (or 64) rays' working sets fit shared memory twice per SM, else lanes = samples
  // (C3D_RESAMPLE=warp|lane forces one of them for A/B runs)
  {
    const bool force_warp = options().resample == 1;
    for (int tpb = 128; tpb >= 64 && !force_warp; tpb >>= 1) {
      const resample::LaneLayout LL = resample::make_lane_layout(q.n_samples, q.n_importance, q.pts_merged != nullptr, tpb);
      const size_t smem = (size_t)LL.total * sizeof(float);
      if (smem > 110 * 1024) continue;
      const long long blocks = (q.n_rays + tpb - 1) / tpb;
      C3D_CHECK_ARG(blocks < (1ll << 31), "too many rays");
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
      if (tpb == 128) {
        C3D_SMEM_ATTR(resample::sample_pdf_lane_kernel<128>, smem, 0, 1);
        resample::sample_pdf_lane_kernel<128><<<(unsigned)blocks, 128, smem, st>>>(q, LL);
      } else {
        C3D_SMEM_ATTR(resample::sample_pdf_lane_kernel<64>, smem, 0, 1);
        resample::sample_pdf_lane_kernel<64><<<(unsigned)blocks, 64, smem, st>>>(q, LL);
      }
      C3D_LAUNCH_CHECK();
      return C3D_OK;
    }
    C3D_CHECK_ARG(options().resample != 2, "resample=lane: the working set does not fit shared memory");
  }
  const int rb_env = options().resample_rb;
  const resample::Layout L = resample::make_layout(q.n_samples, q.n_importance, q.u != nullptr, q.rays_o != nullptr,
                                                   q.rays_d != nullptr, q.z_fine != nullptr, q.z_merged != nullptr,
                                                   q.pts_merged != nullptr, rb_env);
  const size_t smem = (size_t)L.total * sizeof(float);
  C3D_CHECK_ARG(smem <= 200 * 1024, "resampling chunk does not fit shared memory (%zu bytes)", smem);
  C3D_SMEM_ATTR(resample::sample_pdf_kernel, smem, 0, 1);
  const long long n_chunks = (q.n_rays + L.RB - 1) / L.RB;
  C3D_CHECK_ARG(n_chunks < (1ll << 31), "too many rays");
  const int sms = device_sms();
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  const unsigned grid = (unsigned)(n_chunks < (long long)sms * per_sm ? n_chunks : (long long)sms * per_sm);
  resample::sample_pdf_kernel<<<grid, resample::THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(q, L, (int)n_chunks);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

int c3d_camera_params(const float* azim, const float* elev, int32_t n, int32_t img_size, const float* fov_ang, float fov_scalar,
                      float dist_radius, float* cam_poses, float* focal, float* near, float* far, float* jac,
                      c3d_stream_t stream) {
  C3D_CHECK_ARG(n >= 1 && img_size >= 1, "n and img_size must be >= 1");
  C3D_CHECK_ARG(azim && elev && cam_poses && focal && near && far, "azim/elev and the four outputs must be non-NULL");
  C3D_CHECK_ARG(fov_ang || fov_scalar > 0.f, "fov must be > 0");
  invaux::camera_kernel<<<(unsigned)((n + 63) / 64), 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      azim, elev, n, (float)img_size, fov_ang, fov_scalar, dist_radius, cam_poses, focal, near, far, jac);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

int c3d_adam_clip_step(const c3d_adam_params* p, c3d_stream_t stream) {
  C3D_CHECK_ARG(p && p->n_tensors >= 1 && p->n_tensors <= invaux::ADAM_MAX_TENSORS, "n_tensors outside [1,%d]",
                invaux::ADAM_MAX_TENSORS);
  C3D_CHECK_ARG(p->step && p->lr[0] && p->lr[1], "step and both learning-rate scalars must be non-NULL");
  C3D_CHECK_ARG(p->beta1 >= 0.f && p->beta1 < 1.f && p->beta2 >= 0.f && p->beta2 < 1.f && p->eps > 0.f, "bad Adam constants");
  invaux::AdamArgs a;
  memset(&a, 0, sizeof(a));
  a.n_tensors = p->n_tensors;
  for (int t = 0; t < p->n_tensors; ++t) {
    C3D_CHECK_ARG(p->group[t] == 0 || p->group[t] == 1, "tensor %d: group must be 0 or 1", t);
    C3D_CHECK_ARG(p->numel[t] >= 1 && p->param[t] && p->grad[t] && p->exp_avg[t] && p->exp_avg_sq[t], "tensor %d: NULL pointer or empty", t);
    a.group[t] = p->group[t]; a.numel[t] = p->numel[t]; a.param[t] = p->param[t]; a.grad[t] = p->grad[t];
    a.exp_avg[t] = p->exp_avg[t]; a.exp_avg_sq[t] = p->exp_avg_sq[t];
  }
  a.lr[0] = p->lr[0]; a.lr[1] = p->lr[1]; a.step = p->step;
  a.beta1 = p->beta1; a.beta2 = p->beta2; a.eps = p->eps; a.max_norm = p->max_norm; a.grad_norm = p->grad_norm;
  invaux::adam_clip_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

int c3d_umma_selftest(const uint16_t* a, const uint16_t* b, float* d, int32_t N, int32_t K, int32_t variant,
                      c3d_stream_t stream) {
  C3D_CHECK_ARG(a && b && d, "NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (K == 16) {       // layer-0 operand layout (K-major, no swizzle); variant 1 swaps LBO/SBO (diagnostic)
    C3D_CHECK_ARG(N == 128, "K=16 self-test needs N=128 (got %d)", N);
    fused::umma_k16_selftest_kernel<<<1, 128, 0, st>>>(a, b, d, variant);
    C3D_LAUNCH_CHECK();
    return C3D_OK;
  }
  C3D_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0, "N=%d must be a multiple of 16 in [16,256]", N);
  C3D_CHECK_ARG(K >= 64 && K <= 256 && K % 64 == 0, "K=%d must be 16 or a multiple of 64 in [64,256]", K);
  if (variant & 8) {   // CTA-pair MMA (cta_group::2): a is (256,K), d is (256,N); bit 0: A MN-major
    const int smem_p = 65536 + (N / 2) * 128 * (K / 64) + 1024;
    C3D_CUDA(cudaFuncSetAttribute(fused::umma_pair_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_p; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    C3D_CUDA(cudaLaunchKernelEx(&cfg, fused::umma_pair_selftest_kernel, a, b, d, (int)N, (int)K, (int)(variant & 1)));
    C3D_LAUNCH_CHECK();
    return C3D_OK;
  }
  if (variant != 0) {  // MN-major operand layouts (bit 0: A, bit 1: B, bit 2: diagnostic LBO/SBO exchange)
    C3D_CHECK_ARG(N % 64 == 0, "MN-major self-test needs N %% 64 == 0 (got %d)", N);
    const int smem_mn = 2 * 65536 + 1024;
    C3D_CUDA(cudaFuncSetAttribute(fused::umma_mn_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_mn));
    fused::umma_mn_selftest_kernel<<<1, 128, smem_mn, st>>>(a, b, d, N, K, variant);
    C3D_LAUNCH_CHECK();
    return C3D_OK;
  }
  const int smem = fused::ACT_BYTES + N * 128 * (K / 64) + 1024;
  C3D_CUDA(cudaFuncSetAttribute(fused::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  fused::umma_selftest_kernel<<<1, 128, smem, st>>>(a, b, d, N, K);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

}  // extern "C"

namespace c3d {

struct BwdWs {
  size_t film, first, view, g_film, chunk, total;
  int chunk_imgs;
  size_t c_acc, c_feat, c_rgb, c_sdf, c_w, c_grgb, c_gsdf, c_pts, c_rd, c_vd, c_z, c_gpts, c_grd, c_gvd, c_dh, c_dg;
};
static BwdWs bwd_ws(const c3d_bwd_params* bp) {
  const c3d_fwd_params* p = &bp->fwd;
  BwdWs w;
  memset(&w, 0, sizeof(w));
  size_t o = 0;
  const size_t b = (size_t)p->batch, P = (size_t)p->n_rays * p->n_samples, R = (size_t)p->n_rays, D = (size_t)p->D;
  w.film = o;   o += align_up(b * (D + 1) * W * sizeof(float2), 256);
  w.first = o;  o += align_up(b * W * sizeof(float4), 256);
  w.view = o;   o += align_up(b * W * sizeof(float4), 256);
  w.g_film = o; o += align_up(b * (D + 1) * W * sizeof(float2), 256);
  w.chunk = o;
  const bool poses = p->input_kind == C3D_INPUT_POSES;
  const bool pgrads = bp->g_params != nullptr;
  const size_t per_img = P * (D + 1) * W * 4 + P * (12 + 4 + 4 + 12 + 4) + P * 12 + 2 * R * 12 +
                         (poses ? P * 12 + 2 * R * 12 + P * 4 : 0) + (pgrads ? 2 * P * (D + 1) * W * 4 : 0) + 4096;
  size_t ci = ((size_t)2 << 30) / per_img;
  if (ci < 1) ci = 1;
  if (ci > b) ci = b;
  w.chunk_imgs = (int)ci;
  size_t c = 0;
  w.c_acc = c;  c += align_up(ci * P * D * W * 4, 256);
  w.c_feat = c; c += align_up(ci * P * W * 4, 256);
  w.c_rgb = c;  c += align_up(ci * P * 12, 256);
  w.c_sdf = c;  c += align_up(ci * P * 4, 256);
  w.c_w = c;    c += align_up(ci * P * 4, 256);
  w.c_grgb = c; c += align_up(ci * P * 12, 256);
  w.c_gsdf = c; c += align_up(ci * P * 4, 256);
  w.c_gpts = c; c += align_up(ci * P * 12, 256);      // scratch when the caller does not ask for g_pts / POSES mode
  w.c_grd = c;  c += align_up(ci * R * 12, 256);
  w.c_gvd = c;  c += align_up(ci * R * 12, 256);
  if (poses) {
    w.c_pts = c; c += align_up(ci * P * 12, 256);
    w.c_rd = c;  c += align_up(ci * R * 12, 256);
    w.c_vd = c;  c += align_up(ci * R * 12, 256);
    w.c_z = c;   c += align_up(ci * P * 4, 256);
  }
  if (pgrads) {
    w.c_dh = c; c += align_up(ci * P * (D + 1) * W * 4, 256);
    w.c_dg = c; c += align_up(ci * P * (D + 1) * W * 4, 256);
  }
  w.total = o + c;
  return w;
}

// bf16 mode differentiates through the tensor-core kernels (forward with save + fused backward); fp32 mode and
// n_samples < 8 use the FP32-pipe kernels.  C3D_BWD=simt forces the latter (A/B runs).
static bool bwd_uses_tensor_path(const c3d_bwd_params* bp) {
  if (options().bwd_simt) return false;
  if (bp->g_params) return false;                    // parameter gradients come from the FP32-pipe kernels
  return bp->fwd.mode == C3D_MODE_BF16 && bp->fwd.n_samples >= fused::MIN_SAMPLES;
}

struct BwdTcWs {
  size_t film, first, view, g_film, chunk, total;
  int chunk_imgs, unit_rays, units_per_img, tiles_per_unit;
  size_t c_acc, c_cos, c_feat, c_rgbpt, c_wpt, c_sdfpt, c_gdot, c_grgb, c_gsdf, c_wts, c_orgb, c_ofeat, c_omask, c_oxyz,
      c_gpts, c_grd, c_gvd, c_pts, c_rd, c_vd, c_z;
};
static BwdTcWs bwd_tc_ws(const c3d_bwd_params* bp) {
  const c3d_fwd_params* p = &bp->fwd;
  BwdTcWs w;
  memset(&w, 0, sizeof(w));
  const size_t b = (size_t)p->batch, P = (size_t)p->n_rays * p->n_samples, R = (size_t)p->n_rays, D = (size_t)p->D;
  const int ur = unit_rays_for(p);
  w.unit_rays = ur;
  w.units_per_img = (p->n_rays + ur - 1) / ur;
  w.tiles_per_unit = (ur * p->n_samples + fused::TILE - 1) / fused::TILE;
  const size_t tiles_img = (size_t)w.units_per_img * w.tiles_per_unit;
  size_t o = 0;
  w.film = o;   o += align_up(b * (D + 1) * W * sizeof(float2), 256);
  w.first = o;  o += align_up(b * W * sizeof(float4), 256);
  w.view = o;   o += align_up(b * W * sizeof(float4), 256);
  w.g_film = o; o += align_up(b * (D + 1) * W * sizeof(float2), 256);
  w.chunk = o;
  const bool poses = p->input_kind == C3D_INPUT_POSES;
  const size_t per_img = tiles_img * 65536 * (D + 1) + P * (12 + 4 + 4 + 4 + 12 + 4 + 4 + 12) + R * (12 + 1024 + 8 + 12 + 24) +
                         (poses ? P * 16 + R * 24 : 0) + 8192;
  size_t ci = ((size_t)4 << 30) / per_img;
  if (ci < 1) ci = 1;
  if (ci > b || bp->fwd_saved) ci = b;          // saved forward: the whole batch is one chunk that lives until the backward
  w.chunk_imgs = (int)ci;
  size_t c = 0;
  auto take = [&](size_t bytes) { const size_t at = c; c += align_up(bytes, 256); return at; };
  w.c_acc = take(ci * tiles_img * 65536 * (D + 1));
  w.c_cos = 0;                                    // cos(arg) is recomputed from the fp16 accumulators in the backward
  w.c_feat = 0;                                   // the view layer's output is recomputed from its accumulators (gdot_kernel)
  w.c_rgbpt = take(ci * P * 12); w.c_wpt = take(ci * P * 4); w.c_sdfpt = take(ci * P * 4); w.c_gdot = take(ci * P * 4);
  w.c_grgb = take(ci * P * 12); w.c_gsdf = take(ci * P * 4); w.c_wts = take(ci * P * 4);
  w.c_orgb = take(ci * R * 12); w.c_ofeat = take(ci * R * W * 4); w.c_omask = take(ci * R * 8); w.c_oxyz = take(ci * R * 12);
  w.c_gpts = take(ci * P * 12); w.c_grd = take(ci * R * 12); w.c_gvd = take(ci * R * 12);
  if (poses) { w.c_pts = take(ci * P * 12); w.c_rd = take(ci * R * 12); w.c_vd = take(ci * R * 12); w.c_z = take(ci * P * 4); }
  w.total = o + c;
  return w;
}

static int validate_bwd(const c3d_bwd_params* bp) {
  C3D_CHECK_ARG(bp != nullptr, "params is NULL");
  const c3d_fwd_params* p = &bp->fwd;
  C3D_CHECK_ARG(p->abi_version == C3D_ABI_VERSION, "abi_version %d != library %d", p->abi_version, C3D_ABI_VERSION);
  C3D_CHECK_ARG(p->input_kind == C3D_INPUT_POSES || p->input_kind == C3D_INPUT_POINTS, "bad input_kind %d", p->input_kind);
  C3D_CHECK_ARG(p->batch >= 1 && p->n_rays >= 1, "batch=%d n_rays=%d must be >= 1", p->batch, p->n_rays);
  C3D_CHECK_ARG(p->n_samples >= 2 && p->n_samples <= 256, "n_samples=%d outside [2,256]", p->n_samples);
  C3D_CHECK_ARG(p->D >= 1 && p->D <= C3D_MAX_LAYERS, "D=%d outside [1,%d]", p->D, C3D_MAX_LAYERS);
  C3D_CHECK_ARG((long long)p->batch * p->n_rays * p->n_samples < (1ll << 31), "batch*n_rays*n_samples overflows int32");
  C3D_CHECK_ARG(p->packed && p->styles && p->near && p->far, "packed/styles/near/far must be non-NULL");
  C3D_CHECK_ARG(p->feat_layout == C3D_FEAT_NHWC, "backward takes g_feature_map in (b,hw,256) layout");
  if (p->input_kind == C3D_INPUT_POSES) {
    C3D_CHECK_ARG(p->cam_poses && p->focal, "POSES input needs cam_poses and focal");
    C3D_CHECK_ARG(p->img_size >= 1 && p->n_rays == p->img_size * p->img_size, "n_rays=%d != img_size^2", p->n_rays);
  } else {
    C3D_CHECK_ARG(p->pts && p->rays_d && p->viewdirs && p->z_vals, "POINTS input needs pts, rays_d, viewdirs, z_vals");
  }
  if (bp->g_params) {
    const c3d_param_grads* g = bp->g_params;
    bool all = g->views_weight && g->views_bias && g->views_gamma_weight && g->views_gamma_bias && g->views_beta_weight &&
               g->views_beta_bias && g->rgb_weight && g->rgb_bias && g->sigma_weight && g->sigma_bias && g->sigmoid_beta;
    for (int l = 0; l < p->D; ++l)
      all = all && g->pts_weight[l] && g->pts_bias[l] && g->pts_gamma_weight[l] && g->pts_gamma_bias[l] &&
            g->pts_beta_weight[l] && g->pts_beta_bias[l];
    C3D_CHECK_ARG(all, "g_params must carry every parameter gradient pointer of layers 0..D-1, the view layer and the heads");
  }
  const size_t need = bwd_uses_tensor_path(bp) ? bwd_tc_ws(bp).total : bwd_ws(bp).total;
  C3D_CHECK_ARG(p->workspace && p->workspace_bytes >= need, "workspace too small: %zu < %zu", p->workspace_bytes, need);
  C3D_CHECK_ARG(aligned16(p->workspace) && aligned16(bp->g_feature_map), "workspace / g_feature_map must be 16-byte aligned");
  return C3D_OK;
}

}  // namespace c3d

extern "C" {

size_t c3d_backward_workspace_bytes(const c3d_bwd_params* p) {
  if (!p || p->fwd.batch < 1 || p->fwd.n_rays < 1 || p->fwd.n_samples < 1 || p->fwd.D < 1) return 0;
  if (p->fwd_saved && !bwd_uses_tensor_path(p)) return 0;       // saved forward: tensor-core path only
  return bwd_uses_tensor_path(p) ? bwd_tc_ws(p).total : bwd_ws(p).total;
}

static int backward_simt(const c3d_bwd_params* bp, cudaStream_t st);
static int backward_tc(const c3d_bwd_params* bp, cudaStream_t st, int phases);

int c3d_nerf_backward(const c3d_bwd_params* bp, c3d_stream_t stream) {
  g_launches = 0;
  int rc = validate_bwd(bp);
  if (rc != C3D_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bp->fwd_saved) {
    C3D_CHECK_ARG(bp->fwd_saved == 1 && bwd_uses_tensor_path(bp), "fwd_saved is only valid on the tensor-core backward path");
    return backward_tc(bp, st, 2);
  }
  return bwd_uses_tensor_path(bp) ? backward_tc(bp, st, 3) : backward_simt(bp, st);
}

static int backward_simt(const c3d_bwd_params* bp, cudaStream_t st) {
  int rc;
  const c3d_fwd_params* p = &bp->fwd;
  const BwdWs w = bwd_ws(bp);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  uint8_t* ck = ws + w.chunk;
  const size_t P = (size_t)p->n_rays * p->n_samples, R = (size_t)p->n_rays;
  const int D = p->D;
  const PackedLayout L = packed_layout(D);
  const bool poses = p->input_kind == C3D_INPUT_POSES;
  float2* film = reinterpret_cast<float2*>(ws + w.film);
  float4* view = reinterpret_cast<float4*>(ws + w.view);
  float* g_film = reinterpret_cast<float*>(ws + w.g_film);
  rc = launch_style_prep(p->packed, D, p->styles, p->batch, reinterpret_cast<float*>(film),
                         reinterpret_cast<float*>(ws + w.first), reinterpret_cast<float*>(view), st);
  if (rc != C3D_OK) return rc;
  C3D_CUDA(cudaMemsetAsync(g_film, 0, (size_t)p->batch * (D + 1) * W * sizeof(float2), st));
  C3D_SMEM_ATTR(mlp_fp32_kernel, F32_SMEM, 0, 1);
  C3D_SMEM_ATTR(mlp_bwd_kernel<false>, BWD_SMEM, 0, 1);
  C3D_SMEM_ATTR(mlp_bwd_kernel<true>, BWD_SMEM, 0, 1);
  const c3d_param_grads* pg = bp->g_params;
  if (pg) {                                          // everything below accumulates with atomics
    for (int l = 0; l < D; ++l) C3D_CUDA(cudaMemsetAsync(pg->pts_weight[l], 0, sizeof(float) * W * (l == 0 ? 3 : W), st));
    C3D_CUDA(cudaMemsetAsync(pg->views_weight, 0, sizeof(float) * W * (W + 3), st));
    C3D_CUDA(cudaMemsetAsync(pg->rgb_weight, 0, sizeof(float) * 3 * W, st));
    C3D_CUDA(cudaMemsetAsync(pg->rgb_bias, 0, sizeof(float) * 3, st));
    C3D_CUDA(cudaMemsetAsync(pg->sigma_weight, 0, sizeof(float) * W, st));
    C3D_CUDA(cudaMemsetAsync(pg->sigma_bias, 0, sizeof(float), st));
    C3D_CUDA(cudaMemsetAsync(pg->sigmoid_beta, 0, sizeof(float), st));
  }
  const float* beta_ptr = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p->packed) + L.scal) + 4;
  for (int i0 = 0; i0 < p->batch; i0 += w.chunk_imgs) {
    const int ni = (p->batch - i0 < w.chunk_imgs) ? p->batch - i0 : w.chunk_imgs;
    const float *pts, *rays_d, *viewdirs, *z_vals;
    c3d_raygen_params rg;
    memset(&rg, 0, sizeof(rg));
    if (poses) {
      rg.batch = ni; rg.img_size = p->img_size; rg.n_samples = p->n_samples; rg.static_viewdirs = p->static_viewdirs;
      rg.cam_poses = p->cam_poses + (size_t)i0 * 12; rg.focal = p->focal + i0; rg.near = p->near + i0; rg.far = p->far + i0;
      rg.ray_offset = p->ray_offset ? p->ray_offset + (size_t)i0 * R : nullptr;
      rg.pts = reinterpret_cast<float*>(ck + w.c_pts); rg.rays_d = reinterpret_cast<float*>(ck + w.c_rd);
      rg.viewdirs = reinterpret_cast<float*>(ck + w.c_vd); rg.z_vals = reinterpret_cast<float*>(ck + w.c_z);
      const long long n = (long long)ni * R * p->n_samples;
      raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rg);
      C3D_LAUNCH_CHECK();
      pts = rg.pts; rays_d = rg.rays_d; viewdirs = rg.viewdirs; z_vals = rg.z_vals;
    } else {
      pts = p->pts + (size_t)i0 * P * 3; rays_d = p->rays_d + (size_t)i0 * R * 3;
      viewdirs = p->viewdirs + (size_t)i0 * R * 3; z_vals = p->z_vals + (size_t)i0 * P;
    }
    // gradient destinations of this chunk (scratch when the caller passed NULL, always scratch in POSES mode)
    float* g_pts = (!poses && bp->g_pts) ? bp->g_pts + (size_t)i0 * P * 3 : reinterpret_cast<float*>(ck + w.c_gpts);
    float* g_rd = (!poses && bp->g_rays_d) ? bp->g_rays_d + (size_t)i0 * R * 3 : reinterpret_cast<float*>(ck + w.c_grd);
    float* g_vd = (!poses && bp->g_viewdirs) ? bp->g_viewdirs + (size_t)i0 * R * 3 : reinterpret_cast<float*>(ck + w.c_gvd);
    C3D_CUDA(cudaMemsetAsync(g_vd, 0, (size_t)ni * R * 12, st));
    // 1. forward recompute with saved accumulators
    MlpF32Args m;
    m.blob = reinterpret_cast<const uint8_t*>(p->packed); m.L = L;
    m.film = film + (size_t)i0 * (D + 1) * W;
    m.first = reinterpret_cast<const float4*>(ws + w.first) + (size_t)i0 * W;
    m.view = view + (size_t)i0 * W;
    m.pts = pts; m.viewdirs = viewdirs; m.near = p->near + i0; m.far = p->far + i0;
    m.n_samples = p->n_samples; m.pts_per_img = (int)P; m.tiles_per_img = (int)((P + F32_TP - 1) / F32_TP);
    m.feat = reinterpret_cast<float*>(ck + w.c_feat); m.rgb = reinterpret_cast<float*>(ck + w.c_rgb);
    m.sdf = reinterpret_cast<float*>(ck + w.c_sdf);
    m.save_acc = reinterpret_cast<float*>(ck + w.c_acc); m.save_stride = (size_t)ni * P * W; m.n_imgs = ni;
    { const int rc = launch_mlp_fp32(m, ni, st); if (rc != C3D_OK) return rc; }   // tensor cores unless fp32=simt
    // 2. compositing backward
    CompositeBwdArgs c;
    memset(&c, 0, sizeof(c));
    c.n_rays = (long long)ni * R; c.n_samples = p->n_samples; c.n_feat = W; c.sigmoid_beta_ptr = beta_ptr;
    c.rgb = m.rgb; c.sdf = m.sdf; c.features = m.feat; c.z_vals = z_vals; c.rays_d = rays_d; c.pts = pts;
    c.g_rgb_map = bp->g_rgb_map ? bp->g_rgb_map + (size_t)i0 * R * 3 : nullptr;
    c.g_feature_map = bp->g_feature_map ? bp->g_feature_map + (size_t)i0 * R * W : nullptr;
    c.g_xyz = bp->g_xyz ? bp->g_xyz + (size_t)i0 * R * 3 : nullptr;
    c.g_mask = bp->g_mask ? bp->g_mask + (size_t)i0 * R * 2 : nullptr;
    c.g_sdf_in = bp->g_sdf ? bp->g_sdf + (size_t)i0 * P : nullptr;
    c.weights = reinterpret_cast<float*>(ck + w.c_w); c.g_rgb = reinterpret_cast<float*>(ck + w.c_grgb);
    c.g_sdf = reinterpret_cast<float*>(ck + w.c_gsdf); c.g_features = nullptr;
    c.g_pts = g_pts; c.g_rays_d = g_rd; c.g_beta = pg ? pg->sigmoid_beta : nullptr;
    composite_bwd_kernel<<<(unsigned)((c.n_rays + 7) / 8), 256, 0, st>>>(c);
    C3D_LAUNCH_CHECK();
    // 3. MLP backward
    MlpBwdArgs mb;
    mb.blob = m.blob; mb.L = L; mb.film = m.film; mb.view = m.view;
    mb.pts = pts; mb.viewdirs = viewdirs; mb.near = m.near; mb.far = m.far;
    mb.n_samples = p->n_samples; mb.pts_per_img = (int)P; mb.tiles_per_img = m.tiles_per_img;
    mb.save_acc = m.save_acc; mb.save_stride = m.save_stride;
    mb.weights = c.weights; mb.g_feature_map = c.g_feature_map; mb.g_rgb = c.g_rgb; mb.g_sdf = c.g_sdf;
    mb.g_film = g_film + (size_t)i0 * (D + 1) * W * 2; mb.g_pts = g_pts; mb.g_viewdirs = g_vd;
    mb.dump_h = pg ? reinterpret_cast<float*>(ck + w.c_dh) : nullptr;
    mb.dump_g = pg ? reinterpret_cast<float*>(ck + w.c_dg) : nullptr;
    mb.dump_stride = (size_t)ni * P * W;
    if (pg) mlp_bwd_kernel<true><<<(unsigned)(ni * mb.tiles_per_img), 256, BWD_SMEM, st>>>(mb);
    else mlp_bwd_kernel<false><<<(unsigned)(ni * mb.tiles_per_img), 256, BWD_SMEM, st>>>(mb);
    C3D_LAUNCH_CHECK();
    // 3b. parameter gradients of this chunk from the dumped layer outputs H[l] / accumulator cotangents G[l]
    if (pg) {
      const long long np = (long long)ni * P;
      const int splits = (int)((np + 4095) / 4096 < 74 ? (np + 4095) / 4096 : 74);
      const int per_cta = (int)(((np + splits - 1) / splits + WG_PC - 1) / WG_PC * WG_PC);
      for (int l = 1; l <= D; ++l) {
        wgrad_gemm_kernel<<<dim3(4, splits), 256, 0, st>>>(mb.dump_g + (size_t)l * mb.dump_stride, mb.dump_h + (size_t)(l - 1) * mb.dump_stride,
                                                          np, per_cta, l < D ? pg->pts_weight[l] : pg->views_weight, l < D ? W : W + 3);
        C3D_LAUNCH_CHECK();
      }
      HeadWgradArgs h;
      memset(&h, 0, sizeof(h));
      h.pts_per_img = (int)P;
      const int hs = (int)((P + 2047) / 2048 < 64 ? (P + 2047) / 2048 : 64);
      h.pts_per_cta = (int)((P + hs - 1) / hs);
      // W_0 (256,3): G[0]^T (pts * 2/(far-near))
      h.a = pts; h.a_cols = 3; h.a_div = 1; h.H = mb.dump_g; h.near = m.near; h.far = m.far;
      h.out = pg->pts_weight[0]; h.so_j = 1; h.so_c = 3; h.out_sum = nullptr;
      head_wgrad_kernel<<<dim3(hs, ni), 256, 0, st>>>(h);
      C3D_LAUNCH_CHECK();
      // view-direction columns of W_view (256,259)
      h.a = viewdirs; h.a_div = p->n_samples; h.H = mb.dump_g + (size_t)D * mb.dump_stride; h.near = h.far = nullptr;
      h.out = pg->views_weight + W; h.so_j = 1; h.so_c = W + 3;
      head_wgrad_kernel<<<dim3(hs, ni), 256, 0, st>>>(h);
      C3D_LAUNCH_CHECK();
      // rgb head (3,256) on the view layer's output, sigma head (1,256) on h_{D-1}
      h.a = c.g_rgb; h.a_div = 1; h.H = mb.dump_h + (size_t)D * mb.dump_stride; h.out = pg->rgb_weight; h.so_j = W; h.so_c = 1;
      h.out_sum = pg->rgb_bias;
      head_wgrad_kernel<<<dim3(hs, ni), 256, 0, st>>>(h);
      C3D_LAUNCH_CHECK();
      h.a = c.g_sdf; h.a_cols = 1; h.H = mb.dump_h + (size_t)(D - 1) * mb.dump_stride; h.out = pg->sigma_weight; h.out_sum = pg->sigma_bias;
      head_wgrad_kernel<<<dim3(hs, ni), 256, 0, st>>>(h);
      C3D_LAUNCH_CHECK();
    }
    // 4. POSES entry: chain to the camera
    if (poses && (bp->g_cam_poses || bp->g_focal)) {
      dim3 grid((unsigned)((R + 127) / 128), ni);
      raygen_bwd_kernel<<<grid, 128, 0, st>>>(rg, g_pts, g_rd, g_vd, bp->g_cam_poses ? bp->g_cam_poses + (size_t)i0 * 12 : nullptr,
                                              bp->g_focal ? bp->g_focal + i0 : nullptr);
      C3D_LAUNCH_CHECK();
    }
  }
  if (bp->g_styles) {
    film_bwd_kernel<<<dim3(D + 1, p->batch), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(p->packed), L, g_film, bp->g_styles);
    C3D_LAUNCH_CHECK();
  }
  if (pg) {
    film_param_bwd_kernel<<<dim3(D + 1, W), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(p->packed), L, *pg, g_film, film,
                                                           p->styles, p->batch);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

// phases: 1 = forward half (style tables, rays, save-mode forward), 2 = backward half, 3 = both per chunk (recompute).
// With bp->fwd_saved the workspace holds the whole batch as one chunk: c3d_nerf_forward_save runs phase 1 writing the
// caller's output tensors, c3d_nerf_backward then runs phase 2 on the tiles that forward left in the workspace.
static int backward_tc(const c3d_bwd_params* bp, cudaStream_t st, int phases) {
  int rc;
  const c3d_fwd_params* p = &bp->fwd;
  const BwdTcWs w = bwd_tc_ws(bp);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  uint8_t* ck = ws + w.chunk;
  const size_t P = (size_t)p->n_rays * p->n_samples, R = (size_t)p->n_rays;
  const int D = p->D;
  const PackedLayout L = packed_layout(D);
  const bool poses = p->input_kind == C3D_INPUT_POSES;
  const bool user_out = bp->fwd_saved && (phases & 1);       // the forward half produces the caller's outputs
  float2* film = reinterpret_cast<float2*>(ws + w.film);
  float4* first = reinterpret_cast<float4*>(ws + w.first);
  float4* view = reinterpret_cast<float4*>(ws + w.view);
  float* g_film = reinterpret_cast<float*>(ws + w.g_film);
  if (phases & 1) {
    rc = launch_style_prep(p->packed, D, p->styles, p->batch, reinterpret_cast<float*>(film), reinterpret_cast<float*>(first),
                           reinterpret_cast<float*>(view), st);
    if (rc != C3D_OK) return rc;
  }
  if (phases & 2) C3D_CUDA(cudaMemsetAsync(g_film, 0, (size_t)p->batch * (D + 1) * W * sizeof(float2), st));
  const float* beta_ptr = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p->packed) + L.scal) + 4;
  for (int i0 = 0; i0 < p->batch; i0 += w.chunk_imgs) {
    const int ni = (p->batch - i0 < w.chunk_imgs) ? p->batch - i0 : w.chunk_imgs;
    c3d_fwd_params q = *p;                         // this chunk as a POINTS-entry launch
    q.batch = ni; q.input_kind = C3D_INPUT_POINTS;
    q.styles = p->styles + (size_t)i0 * (D + 1) * W; q.near = p->near + i0; q.far = p->far + i0;
    c3d_raygen_params rg;
    memset(&rg, 0, sizeof(rg));
    if (poses) {
      rg.batch = ni; rg.img_size = p->img_size; rg.n_samples = p->n_samples; rg.static_viewdirs = p->static_viewdirs;
      rg.cam_poses = p->cam_poses + (size_t)i0 * 12; rg.focal = p->focal + i0; rg.near = q.near; rg.far = q.far;
      rg.ray_offset = p->ray_offset ? p->ray_offset + (size_t)i0 * R : nullptr;
      rg.pts = reinterpret_cast<float*>(ck + w.c_pts); rg.rays_d = reinterpret_cast<float*>(ck + w.c_rd);
      rg.viewdirs = reinterpret_cast<float*>(ck + w.c_vd); rg.z_vals = reinterpret_cast<float*>(ck + w.c_z);
      if (phases & 1) {
        const long long n = (long long)ni * R * p->n_samples;
        raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rg);
        C3D_LAUNCH_CHECK();
      }
      q.pts = rg.pts; q.rays_d = rg.rays_d; q.viewdirs = rg.viewdirs; q.z_vals = rg.z_vals;
    } else {
      q.pts = p->pts + (size_t)i0 * P * 3; q.rays_d = p->rays_d + (size_t)i0 * R * 3;
      q.viewdirs = p->viewdirs + (size_t)i0 * R * 3; q.z_vals = p->z_vals + (size_t)i0 * P;
    }
    // the caller's feature_map (either layout) is written by the save-mode forward itself; the recompute pass writes a
    // scratch copy in (b, hw, 256)
    if (!user_out) q.feat_layout = C3D_FEAT_NHWC;
    q.rgb_map = user_out ? p->rgb_map : reinterpret_cast<float*>(ck + w.c_orgb);
    q.feature_map = user_out ? p->feature_map : reinterpret_cast<float*>(ck + w.c_ofeat);
    q.mask = user_out ? p->mask : reinterpret_cast<float*>(ck + w.c_omask);
    q.xyz = user_out ? p->xyz : reinterpret_cast<float*>(ck + w.c_oxyz);
    q.sdf = reinterpret_cast<float*>(ck + w.c_sdfpt); q.z_vals_out = nullptr;
    float* g_pts = (!poses && bp->g_pts) ? bp->g_pts + (size_t)i0 * P * 3 : reinterpret_cast<float*>(ck + w.c_gpts);
    float* g_rd = (!poses && bp->g_rays_d) ? bp->g_rays_d + (size_t)i0 * R * 3 : reinterpret_cast<float*>(ck + w.c_grd);
    float* g_vd = (!poses && bp->g_viewdirs) ? bp->g_viewdirs + (size_t)i0 * R * 3 : reinterpret_cast<float*>(ck + w.c_gvd);
    const float* gF = bp->g_feature_map ? bp->g_feature_map + (size_t)i0 * R * W : nullptr;
    fused::Args a;
    fused_fill_args(a, &q, film + (size_t)i0 * (D + 1) * W, first + (size_t)i0 * W, view + (size_t)i0 * W);
    a.save_acc = reinterpret_cast<__nv_bfloat16*>(ck + w.c_acc);
    a.rgb_pt = reinterpret_cast<float*>(ck + w.c_rgbpt); a.w_pt = reinterpret_cast<float*>(ck + w.c_wpt);
    a.g_feature_map = gF;
    if (phases & 1) {
      // 1. forward on the tensor cores, keeping bf16 acc / cos tiles per layer and per-point rgb / weights
      rc = fused_launch(a, 1, st);
      if (rc != C3D_OK) return rc;
      if (user_out) {                               // single chunk (i0 == 0): hand the remaining outputs to the caller
        C3D_CUDA(cudaMemcpyAsync(p->sdf, q.sdf, (size_t)ni * P * 4, cudaMemcpyDeviceToDevice, st));
        if (poses && p->z_vals_out)
          C3D_CUDA(cudaMemcpyAsync(p->z_vals_out, q.z_vals, (size_t)ni * P * 4, cudaMemcpyDeviceToDevice, st));
      }
    }
    if (!(phases & 2)) continue;
    C3D_CUDA(cudaMemsetAsync(g_vd, 0, (size_t)ni * R * 12, st));
    // 2. d(volume_integration): needs g_feature_map . feat per point
    float* gdot = reinterpret_cast<float*>(ck + w.c_gdot);
    if (gF) {
      const long long warps = a.n_tiles_g * 16;
      fusedbwd::gdot_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(a, gdot);
      C3D_LAUNCH_CHECK();
    }
    CompositeBwdArgs c;
    memset(&c, 0, sizeof(c));
    c.n_rays = (long long)ni * R; c.n_samples = p->n_samples; c.n_feat = W; c.sigmoid_beta_ptr = beta_ptr;
    c.rgb = a.rgb_pt; c.sdf = q.sdf; c.features = nullptr; c.gdot = gF ? gdot : nullptr;
    c.z_vals = q.z_vals; c.rays_d = q.rays_d; c.pts = q.pts;
    c.g_rgb_map = bp->g_rgb_map ? bp->g_rgb_map + (size_t)i0 * R * 3 : nullptr;
    c.g_feature_map = nullptr;
    c.g_xyz = bp->g_xyz ? bp->g_xyz + (size_t)i0 * R * 3 : nullptr;
    c.g_mask = bp->g_mask ? bp->g_mask + (size_t)i0 * R * 2 : nullptr;
    c.g_sdf_in = bp->g_sdf ? bp->g_sdf + (size_t)i0 * P : nullptr;
    c.weights = reinterpret_cast<float*>(ck + w.c_wts); c.g_rgb = reinterpret_cast<float*>(ck + w.c_grgb);
    c.g_sdf = reinterpret_cast<float*>(ck + w.c_gsdf); c.g_features = nullptr; c.g_pts = g_pts; c.g_rays_d = g_rd;
    composite_bwd_kernel<<<(unsigned)((c.n_rays + 7) / 8), 256, 0, st>>>(c);
    C3D_LAUNCH_CHECK();
    // 3. fused tensor-core backward of the MLP
    a.g_rgb_pt = c.g_rgb; a.g_sdf_pt = c.g_sdf; a.g_film = g_film + (size_t)i0 * (D + 1) * W * 2;
    a.g_pts = g_pts; a.g_viewdirs = g_vd;
    rc = fused_launch(a, 2, st);
    if (rc != C3D_OK) return rc;
    if (poses && (bp->g_cam_poses || bp->g_focal)) {
      dim3 grid((unsigned)((R + 127) / 128), ni);
      raygen_bwd_kernel<<<grid, 128, 0, st>>>(rg, g_pts, g_rd, g_vd, bp->g_cam_poses ? bp->g_cam_poses + (size_t)i0 * 12 : nullptr,
                                              bp->g_focal ? bp->g_focal + i0 : nullptr);
      C3D_LAUNCH_CHECK();
    }
  }
  if ((phases & 2) && bp->g_styles) {
    film_bwd_kernel<<<dim3(D + 1, p->batch), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(p->packed), L, g_film, bp->g_styles);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

int c3d_nerf_forward_save(const c3d_bwd_params* bp, c3d_stream_t stream) {
  g_launches = 0;
  C3D_CHECK_ARG(bp && bp->fwd_saved == 1, "c3d_nerf_forward_save needs fwd_saved == 1");
  C3D_CHECK_ARG(bwd_uses_tensor_path(bp), "the saved-forward path exists for the tensor-core backward only "
                                          "(bf16 mode, n_samples >= 8, no parameter gradients)");
  const c3d_fwd_params* p = &bp->fwd;
  C3D_CHECK_ARG(p->rgb_map && p->feature_map && p->sdf && p->mask && p->xyz, "all five output pointers are required");
  C3D_CHECK_ARG(aligned16(p->rgb_map) && aligned16(p->feature_map) && aligned16(p->sdf) && aligned16(p->mask) &&
                aligned16(p->xyz) && aligned16(p->z_vals_out), "output pointers must be 16-byte aligned");
  const int layout = p->feat_layout;
  C3D_CHECK_ARG(layout == C3D_FEAT_NHWC || layout == C3D_FEAT_NCHW, "bad feat_layout %d", layout);
  c3d_bwd_params b = *bp;                      // validate_bwd checks the backward's own (NHWC cotangent) convention
  b.fwd.feat_layout = C3D_FEAT_NHWC;
  int rc = validate_bwd(&b);
  if (rc != C3D_OK) return rc;
  return backward_tc(bp, reinterpret_cast<cudaStream_t>(stream), 1);
}

}  // extern "C" (helpers above are static)

namespace c3d {

struct EikWs {
  size_t film, first, view, g_film, chunk, total;
  int chunk_imgs;
  size_t c_acc, c_tacc, c_feat, c_rgb, c_sdf, c_h, c_hd, c_g, c_gd;
};
static EikWs eik_ws(const c3d_bwd_params* bp) {
  const c3d_fwd_params* p = &bp->fwd;
  EikWs w;
  memset(&w, 0, sizeof(w));
  size_t o = 0;
  const size_t b = (size_t)p->batch, P = (size_t)p->n_rays * p->n_samples, D = (size_t)p->D;
  w.film = o;   o += align_up(b * (D + 1) * W * sizeof(float2), 256);
  w.first = o;  o += align_up(b * W * sizeof(float4), 256);
  w.view = o;   o += align_up(b * W * sizeof(float4), 256);
  w.g_film = o; o += align_up(b * (D + 1) * W * sizeof(float2), 256);
  w.chunk = o;
  const size_t per_img = P * W * 4 * (6 * D + 1) + P * 16 + 8192;
  size_t ci = ((size_t)4 << 30) / per_img;
  if (ci < 1) ci = 1;
  if (ci > b) ci = b;
  w.chunk_imgs = (int)ci;
  size_t c = 0;
  auto take = [&](size_t bytes) { const size_t at = c; c += align_up(bytes, 256); return at; };
  w.c_acc = take(ci * P * D * W * 4); w.c_tacc = take(ci * P * D * W * 4);
  w.c_feat = take(ci * P * W * 4); w.c_rgb = take(ci * P * 12); w.c_sdf = take(ci * P * 4);
  w.c_h = take(ci * P * D * W * 4); w.c_hd = take(ci * P * D * W * 4);
  w.c_g = take(ci * P * D * W * 4); w.c_gd = take(ci * P * D * W * 4);
  w.total = o + c;
  return w;
}

}  // namespace c3d

extern "C" {

size_t c3d_eikonal_workspace_bytes(const c3d_bwd_params* p) {
  if (!p || p->fwd.batch < 1 || p->fwd.n_rays < 1 || p->fwd.n_samples < 1 || p->fwd.D < 1) return 0;
  return eik_ws(p).total;
}

int c3d_eikonal_backward(const c3d_bwd_params* bp, const float* g_eik, c3d_stream_t stream) {
  g_launches = 0;
  C3D_CHECK_ARG(bp != nullptr && g_eik != nullptr, "params / g_eik is NULL");
  const c3d_fwd_params* p = &bp->fwd;
  C3D_CHECK_ARG(p->abi_version == C3D_ABI_VERSION, "abi_version %d != library %d", p->abi_version, C3D_ABI_VERSION);
  C3D_CHECK_ARG(p->input_kind == C3D_INPUT_POINTS, "the eikonal path takes C3D_INPUT_POINTS inputs");
  C3D_CHECK_ARG(p->batch >= 1 && p->n_rays >= 1 && p->n_samples >= 1, "batch / n_rays / n_samples must be >= 1");
  C3D_CHECK_ARG(p->D >= 1 && p->D <= C3D_MAX_LAYERS, "D=%d outside [1,%d]", p->D, C3D_MAX_LAYERS);
  C3D_CHECK_ARG((long long)p->batch * p->n_rays * p->n_samples < (1ll << 31), "batch*n_rays*n_samples overflows int32");
  C3D_CHECK_ARG(p->packed && p->styles && p->near && p->far && p->pts && p->viewdirs, "packed/styles/near/far/pts/viewdirs must be non-NULL");
  C3D_CHECK_ARG(bp->g_styles || bp->g_params, "nothing to compute: g_styles and g_params are both NULL");
  const EikWs w = eik_ws(bp);
  C3D_CHECK_ARG(p->workspace && p->workspace_bytes >= w.total, "workspace too small: %zu < %zu", p->workspace_bytes, w.total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  uint8_t* ck = ws + w.chunk;
  const size_t P = (size_t)p->n_rays * p->n_samples, R = (size_t)p->n_rays;
  const int D = p->D;
  const PackedLayout L = packed_layout(D);
  float2* film = reinterpret_cast<float2*>(ws + w.film);
  float* g_film = reinterpret_cast<float*>(ws + w.g_film);
  int rc = launch_style_prep(p->packed, D, p->styles, p->batch, reinterpret_cast<float*>(film),
                             reinterpret_cast<float*>(ws + w.first), reinterpret_cast<float*>(ws + w.view), st);
  if (rc != C3D_OK) return rc;
  C3D_CUDA(cudaMemsetAsync(g_film, 0, (size_t)p->batch * (D + 1) * W * sizeof(float2), st));
  const c3d_param_grads* pg = bp->g_params;
  if (pg) {
    for (int l = 0; l < D; ++l) C3D_CUDA(cudaMemsetAsync(pg->pts_weight[l], 0, sizeof(float) * W * (l == 0 ? 3 : W), st));
    C3D_CUDA(cudaMemsetAsync(pg->views_weight, 0, sizeof(float) * W * (W + 3), st));
    C3D_CUDA(cudaMemsetAsync(pg->rgb_weight, 0, sizeof(float) * 3 * W, st));
    C3D_CUDA(cudaMemsetAsync(pg->rgb_bias, 0, sizeof(float) * 3, st));
    C3D_CUDA(cudaMemsetAsync(pg->sigma_weight, 0, sizeof(float) * W, st));
    C3D_CUDA(cudaMemsetAsync(pg->sigma_bias, 0, sizeof(float), st));
    C3D_CUDA(cudaMemsetAsync(pg->sigmoid_beta, 0, sizeof(float), st));
  }
  C3D_SMEM_ATTR(mlp_fp32_kernel, F32_SMEM, 0, 1);
  C3D_SMEM_ATTR(eik_tangent_kernel, EIKT_SMEM, 0, 1);
  C3D_SMEM_ATTR(eik_bwd_kernel, EIKB_SMEM, 0, 1);
  for (int i0 = 0; i0 < p->batch; i0 += w.chunk_imgs) {
    const int ni = (p->batch - i0 < w.chunk_imgs) ? p->batch - i0 : w.chunk_imgs;
    const float* pts = p->pts + (size_t)i0 * P * 3;
    // 1. primal forward with saved accumulators
    MlpF32Args m;
    m.blob = reinterpret_cast<const uint8_t*>(p->packed); m.L = L;
    m.film = film + (size_t)i0 * (D + 1) * W;
    m.first = reinterpret_cast<const float4*>(ws + w.first) + (size_t)i0 * W;
    m.view = reinterpret_cast<const float4*>(ws + w.view) + (size_t)i0 * W;
    m.pts = pts; m.viewdirs = p->viewdirs + (size_t)i0 * R * 3; m.near = p->near + i0; m.far = p->far + i0;
    m.n_samples = p->n_samples; m.pts_per_img = (int)P; m.tiles_per_img = (int)((P + F32_TP - 1) / F32_TP);
    m.feat = reinterpret_cast<float*>(ck + w.c_feat); m.rgb = reinterpret_cast<float*>(ck + w.c_rgb);
    m.sdf = reinterpret_cast<float*>(ck + w.c_sdf);
    m.save_acc = reinterpret_cast<float*>(ck + w.c_acc); m.save_stride = (size_t)ni * P * W; m.n_imgs = ni;
    { const int rc = launch_mlp_fp32(m, ni, st); if (rc != C3D_OK) return rc; }   // tensor cores unless fp32=simt
    // 2. tangent sweep along v, 3. reverse sweep over both chains
    EikArgs e;
    memset(&e, 0, sizeof(e));
    e.blob = m.blob; e.L = L; e.film = m.film; e.first = m.first;
    e.pts = pts; e.v = g_eik + (size_t)i0 * P * 3; e.near = m.near; e.far = m.far;
    e.pts_per_img = (int)P; e.tiles_per_img = m.tiles_per_img;
    e.save_acc = m.save_acc; e.save_stride = m.save_stride; e.save_tacc = reinterpret_cast<float*>(ck + w.c_tacc);
    e.g_film = g_film + (size_t)i0 * (D + 1) * W * 2;
    e.dump_h = reinterpret_cast<float*>(ck + w.c_h); e.dump_hd = reinterpret_cast<float*>(ck + w.c_hd);
    e.dump_g = reinterpret_cast<float*>(ck + w.c_g); e.dump_gd = reinterpret_cast<float*>(ck + w.c_gd);
    eik_tangent_kernel<<<(unsigned)(ni * e.tiles_per_img), 256, EIKT_SMEM, st>>>(e);
    C3D_LAUNCH_CHECK();
    eik_bwd_kernel<<<(unsigned)(ni * e.tiles_per_img), 256, EIKB_SMEM, st>>>(e);
    C3D_LAUNCH_CHECK();
    // 4. weight gradients from the dumps
    if (pg) {
      const long long np = (long long)ni * P;
      const int splits = (int)((np + 4095) / 4096 < 74 ? (np + 4095) / 4096 : 74);
      const int per_cta = (int)(((np + splits - 1) / splits + WG_PC - 1) / WG_PC * WG_PC);
      for (int l = 1; l < D; ++l) {
        wgrad_gemm_kernel<<<dim3(4, splits), 256, 0, st>>>(e.dump_g + (size_t)l * e.save_stride, e.dump_h + (size_t)(l - 1) * e.save_stride,
                                                          np, per_cta, pg->pts_weight[l], W);
        C3D_LAUNCH_CHECK();
        wgrad_gemm_kernel<<<dim3(4, splits), 256, 0, st>>>(e.dump_gd + (size_t)l * e.save_stride, e.dump_hd + (size_t)(l - 1) * e.save_stride,
                                                          np, per_cta, pg->pts_weight[l], W);
        C3D_LAUNCH_CHECK();
      }
      HeadWgradArgs h;
      memset(&h, 0, sizeof(h));
      h.pts_per_img = (int)P;
      const int hs = (int)((P + 2047) / 2048 < 64 ? (P + 2047) / 2048 : 64);
      h.pts_per_cta = (int)((P + hs - 1) / hs);
      h.a_cols = 3; h.a_div = 1; h.near = m.near; h.far = m.far; h.out = pg->pts_weight[0]; h.so_j = 1; h.so_c = 3;
      h.a = pts; h.H = e.dump_g;                       // adj(u_0)^T x
      head_wgrad_kernel<<<dim3(hs, ni), 256, 0, st>>>(h);
      C3D_LAUNCH_CHECK();
      h.a = e.v; h.H = e.dump_gd;                      // adj(ud_0)^T xd
      head_wgrad_kernel<<<dim3(hs, ni), 256, 0, st>>>(h);
      C3D_LAUNCH_CHECK();
      h.a = nullptr; h.a_cols = 1; h.near = h.far = nullptr; h.H = e.dump_hd + (size_t)(D - 1) * e.save_stride;
      h.out = pg->sigma_weight; h.so_j = W; h.so_c = 1;   // S = w_sigma . hd_{D-1}
      head_wgrad_kernel<<<dim3(hs, ni), 256, 0, st>>>(h);
      C3D_LAUNCH_CHECK();
    }
  }
  if (bp->g_styles) {
    film_bwd_kernel<<<dim3(D + 1, p->batch), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(p->packed), L, g_film, bp->g_styles);
    C3D_LAUNCH_CHECK();
  }
  if (pg) {
    film_param_bwd_kernel<<<dim3(D + 1, W), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(p->packed), L, *pg, g_film, film,
                                                           p->styles, p->batch);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

int c3d_composite_backward(const c3d_composite_params* p, c3d_stream_t stream) {
  C3D_CHECK_ARG(p && p->n_rays >= 1, "n_rays must be >= 1");
  C3D_CHECK_ARG(p->n_samples >= 1 && p->n_samples <= CMP_MAX_N, "n_samples=%d outside [1,%d]", p->n_samples, CMP_MAX_N);
  C3D_CHECK_ARG(p->n_feat >= 0 && p->n_feat % 4 == 0, "n_feat=%d must be a multiple of 4", p->n_feat);
  C3D_CHECK_ARG(p->rgb && p->sdf && p->z_vals && p->rays_d && p->pts, "rgb/sdf/z_vals/rays_d/pts must be non-NULL");
  C3D_CHECK_ARG(p->flags == 0, "composite backward covers the with_sdf=True branch only (flags must be 0)");
  C3D_CHECK_ARG(p->sigmoid_beta_ptr, "composite backward needs sigmoid_beta_ptr (device scalar)");
  C3D_CHECK_ARG(p->weights && p->g_rgb && p->g_sdf && p->g_pts && p->g_rays_d, "weights/g_rgb/g_sdf/g_pts/g_rays_d outputs are required");
  C3D_CHECK_ARG(!p->features || !p->g_feature_map || p->g_features, "g_features output missing");
  CompositeBwdArgs c;
  memset(&c, 0, sizeof(c));
  c.n_rays = p->n_rays; c.n_samples = p->n_samples; c.n_feat = p->n_feat; c.sigmoid_beta_ptr = p->sigmoid_beta_ptr;
  c.rgb = p->rgb; c.sdf = p->sdf; c.features = p->features; c.z_vals = p->z_vals; c.rays_d = p->rays_d; c.pts = p->pts;
  c.g_rgb_map = p->g_rgb_map; c.g_feature_map = p->g_feature_map; c.g_xyz = p->g_xyz; c.g_mask = p->g_mask;
  c.weights = p->weights; c.g_rgb = p->g_rgb; c.g_sdf = p->g_sdf; c.g_features = p->g_features; c.g_pts = p->g_pts;
  c.g_rays_d = p->g_rays_d; c.g_beta = p->g_sigmoid_beta;   // accumulated (+=): zero it before the call
  composite_bwd_kernel<<<(unsigned)((p->n_rays + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(c);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

}  // extern "C"
