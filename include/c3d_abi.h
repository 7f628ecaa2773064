/* c3d_abi.h -- C ABI of libc3dpp.so: the B200-native NeRF branch of CIPS-3D++.
 *
 * The reference has NO native code / FFI on this path (it is ~60 ATen calls issued from
 * Python); the entry points below replace these reference interfaces:
 *
 *   c3d_nerf_forward      VolumeFeatureRenderer.forward        exp/cips3d/volume_renderer.py:192-283
 *                         (+ Render.prepare_nerf_inputs        exp/cips3d/nerf_utils.py:172-218 when
 *                            input_kind == C3D_INPUT_POSES)
 *   c3d_nerf_backward     autograd of the same forward (flip inversion, projector_v9.py:1143)
 *   c3d_raygen            Render.prepare_nerf_inputs           exp/cips3d/nerf_utils.py:172-218
 *   c3d_style_prep        FiLMSiren.gamma / .beta LinearLayers exp/cips3d/volume_renderer.py:66-67,77-81
 *   c3d_composite_forward Render.volume_integration            exp/cips3d/nerf_utils.py:230-338
 *   c3d_composite_backward autograd of volume_integration
 *   c3d_sample_pdf        (extension, no reference counterpart: the reference renders in one pass; canonical NeRF
 *                         hierarchical sampling for the weights of Render.volume_integration nerf_utils.py:267-307)
 *   c3d_pack_weights      (new) re-lays the reference state_dict (volume_renderer.py:107-115,183)
 *                         into the kernel's packed blob; re-derivable from the fp32 state dict
 *
 * Conventions
 *   - plain C, no C++ / torch types; every pointer is a DEVICE pointer to fp32 data unless noted
 *   - ownership: all buffers (inputs, outputs, packed weights, workspace) are allocated and
 *     owned by the caller; the library never allocates, frees or retains a pointer
 *   - errors: 0 = success, negative = failure; text via c3d_last_error() (thread-local)
 *   - streams: work is enqueued on the given CUDA stream; the library never synchronises
 *   - alignment: all pointers 16-byte aligned; W (hidden width) is fixed at 256
 */
#ifndef C3D_ABI_H_
#define C3D_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C3D_ABI_VERSION 11 /* 11: packed-blob layout 2 (fp16 weight images; a blob packed by an older library must be re-packed) */
#define C3D_MAX_LAYERS 16
#define C3D_W 256

typedef struct CUstream_st* c3d_stream_t; /* == cudaStream_t */

enum { C3D_OK = 0, C3D_ERR_ARG = -1, C3D_ERR_CUDA = -2, C3D_ERR_UNSUPPORTED = -3 };
/* Arithmetic of the point MLP.  C3D_MODE_FP32: fp32 parity mode (products formed on the tensor cores from two-way fp16 split
 * operands, fp32-accurate; option fp32=simt selects the FP32-pipe kernel).  C3D_MODE_BF16: the 16-bit tensor-core mode (the
 * reference's autocast counterpart; since ABI 11 its hidden-layer operands are IEEE half, not bfloat16: same rate, 8x finer). */
enum { C3D_MODE_FP32 = 0, C3D_MODE_BF16 = 1 };
enum { C3D_INPUT_POSES = 0, C3D_INPUT_POINTS = 1 };
/* feature_map layout: (b,hw,256) fp32 as the reference's renderer returns it; (b,256,hw) fp32 as the decoder consumes it
 * (model_v3.py:1014); (b,256,hw) bf16 -- the same hand-off at half the bytes (inference only: MODE_BF16, no backward).  The
 * tensor-core kernels write any of them straight from the compositing epilogue. */
enum { C3D_FEAT_NHWC = 0, C3D_FEAT_NCHW = 1, C3D_FEAT_NCHW_BF16 = 2 };

/* Reference state_dict tensors (names: SURVEY.md 3.4), fp32, contiguous, row-major (out,in). */
typedef struct c3d_raw_params {
  int32_t D;                                        /* N_layers_renderer, 1..C3D_MAX_LAYERS */
  int32_t _pad;
  const float* pts_weight[C3D_MAX_LAYERS];          /* [0]:(256,3)  [i>0]:(256,256) */
  const float* pts_bias[C3D_MAX_LAYERS];            /* (256) */
  const float* pts_gamma_weight[C3D_MAX_LAYERS];    /* (256,256) */
  const float* pts_gamma_bias[C3D_MAX_LAYERS];      /* (256) */
  const float* pts_beta_weight[C3D_MAX_LAYERS];     /* (256,256) */
  const float* pts_beta_bias[C3D_MAX_LAYERS];       /* (256) */
  const float* views_weight;                        /* (256,259) */
  const float* views_bias;
  const float* views_gamma_weight;
  const float* views_gamma_bias;
  const float* views_beta_weight;
  const float* views_beta_bias;
  const float* rgb_weight;                          /* (3,256) */
  const float* rgb_bias;                            /* (3) */
  const float* sigma_weight;                        /* (1,256) */
  const float* sigma_bias;                          /* (1) */
  const float* sigmoid_beta;                        /* (1) */
} c3d_raw_params;

/* Gradients w.r.t. the same tensors (same shapes), written by c3d_nerf_backward when requested.  Either the whole
 * set is given (every pointer of layers 0..D-1, the view layer, both heads and sigmoid_beta non-NULL) or none. */
typedef struct c3d_param_grads {
  float* pts_weight[C3D_MAX_LAYERS];
  float* pts_bias[C3D_MAX_LAYERS];
  float* pts_gamma_weight[C3D_MAX_LAYERS];
  float* pts_gamma_bias[C3D_MAX_LAYERS];
  float* pts_beta_weight[C3D_MAX_LAYERS];
  float* pts_beta_bias[C3D_MAX_LAYERS];
  float* views_weight;
  float* views_bias;
  float* views_gamma_weight;
  float* views_gamma_bias;
  float* views_beta_weight;
  float* views_beta_bias;
  float* rgb_weight;
  float* rgb_bias;
  float* sigma_weight;
  float* sigma_bias;
  float* sigmoid_beta;
} c3d_param_grads;

/* Fused all-gather of the rendered maps (multi-GPU serving: every rank ends up with every rank's maps; what the reference
 * does with torch.distributed around its generator, scripts/gen_images.py:57-91, train_v10.py evaluation).  Instead of a
 * collective after the kernel, the compositing epilogue of the bf16 forward kernel stores every finished ray straight into
 * the GATHERED tensors of all peers -- pointers into peer memory mapped over NVLink / NVSwitch (CUDA IPC, symmetric
 * memory; this rank included) -- so the transfer rides under the render and needs no SMs of its own.  The caller owns
 * the buffers, places them identically on every rank, and synchronises the ranks (a barrier after the launch has
 * completed) before any rank reads.  Layouts / dtypes as c3d_fwd_params (feature_map per feat_layout), leading dimension
 * = all images of all ranks; this call's first image lands at index image_offset. */
#define C3D_MAX_PEERS 16
typedef struct c3d_gather_out {
  int32_t n_peers;                       /* destinations, 1..C3D_MAX_PEERS (this rank is one of them) */
  int32_t image_offset;                  /* index of this call's image 0 in the gathered tensors */
  void* feature_map[C3D_MAX_PEERS];      /* per destination: base of the gathered feature tensor */
  float* rgb_map[C3D_MAX_PEERS];
  float* mask[C3D_MAX_PEERS];
  float* xyz[C3D_MAX_PEERS];
} c3d_gather_out;

typedef struct c3d_fwd_params {
  int32_t abi_version;       /* C3D_ABI_VERSION */
  int32_t mode;              /* C3D_MODE_* */
  int32_t input_kind;        /* C3D_INPUT_* */
  int32_t feat_layout;       /* C3D_FEAT_* */
  int32_t batch;             /* images (latent, pose) pairs */
  int32_t n_rays;            /* rays per image (hw; any count >= 1 for C3D_INPUT_POINTS) */
  int32_t n_samples;         /* N, 2..256 */
  int32_t D;                 /* must equal the packed blob's D */
  int32_t img_size;          /* POSES: n_rays == img_size^2 */
  int32_t static_viewdirs;   /* POSES: nerf_utils.py:58-61 */
  const void* packed;        /* c3d_pack_weights output */
  const float* styles;       /* (batch, D+1, 256) */
  /* C3D_INPUT_POSES */
  const float* cam_poses;    /* (batch,3,4) camera-to-world */
  const float* focal;        /* (batch) */
  const float* near;         /* (batch)  also used by C3D_INPUT_POINTS for normalisation */
  const float* far;          /* (batch) */
  const float* ray_offset;   /* NULL (eval) or (batch,n_rays) U[0,1) draws: perturb (nerf_utils.py:105-119) */
  /* C3D_INPUT_POINTS (reference layouts, contiguous) */
  const float* pts;          /* (batch,n_rays,N,3) world space */
  const float* rays_d;       /* (batch,n_rays,3) */
  const float* viewdirs;     /* (batch,n_rays,3) */
  const float* z_vals;       /* (batch,n_rays,N) */
  /* outputs.  sdf is always written; rgb_map / feature_map / mask / xyz are either all given or all NULL -- the latter is
   * a density-only pass (the coarse pass of the optional two-pass render): the kernel stops after the sdf head. */
  float* rgb_map;            /* (batch,n_rays,3) */
  float* feature_map;        /* (batch,n_rays,256) or (batch,256,n_rays) */
  float* sdf;                /* (batch,n_rays,N)  [reference shape (b,hw,N,1)] */
  float* mask;               /* (batch,n_rays,2): background weight, depth = -|xyz| */
  float* xyz;                /* (batch,n_rays,3) */
  float* z_vals_out;         /* optional (POSES): (batch,n_rays,N) sample depths, or NULL */
  void* workspace;           /* >= c3d_workspace_bytes() */
  size_t workspace_bytes;
  const c3d_gather_out* gather;  /* NULL, or (MODE_BF16, n_samples >= 8, no density-only pass): rgb_map / feature_map / mask / xyz
                                  * go to the gathered tensors of every peer INSTEAD of the four pointers above (which may
                                  * then be NULL; sdf and z_vals_out stay local) */
} c3d_fwd_params;

/* Cotangents in, gradients out. Any gradient pointer may be NULL (not computed). */
typedef struct c3d_bwd_params {
  c3d_fwd_params fwd;        /* same inputs as the forward call (outputs unused; recomputed) */
  const float* g_rgb_map;    /* (batch,n_rays,3) or NULL (= zero) */
  const float* g_feature_map;/* layout as fwd.feat_layout, or NULL */
  const float* g_mask;       /* (batch,n_rays,2) or NULL */
  const float* g_xyz;        /* (batch,n_rays,3) or NULL */
  const float* g_sdf;        /* (batch,n_rays,N) or NULL */
  float* g_styles;           /* (batch,D+1,256) */
  float* g_pts;              /* POINTS: (batch,n_rays,N,3) */
  float* g_rays_d;           /* POINTS: (batch,n_rays,3) */
  float* g_viewdirs;         /* POINTS: (batch,n_rays,3) */
  float* g_cam_poses;        /* POSES: (batch,3,4) */
  float* g_focal;            /* POSES: (batch) */
  const c3d_param_grads* g_params; /* HOST pointer or NULL: also return the gradients of the renderer's parameters
                                      (training; runs the FP32-pipe backward in either mode) */
  int32_t fwd_saved;         /* 0: the backward recomputes the forward in chunks of <= 4 GiB of workspace (default).
                                1: the workspace holds the whole batch; c3d_nerf_forward_save filled it (and produced the
                                   forward outputs), c3d_nerf_backward on the SAME workspace then skips the recomputation.
                                   Tensor-core path only (bf16 mode, n_samples >= 8, g_params == NULL). */
  int32_t _pad;
} c3d_bwd_params;

typedef struct c3d_raygen_params {
  int32_t batch, img_size, n_samples, static_viewdirs;
  const float* cam_poses;    /* (batch,3,4) */
  const float* focal;        /* (batch) */
  const float* near;         /* (batch) */
  const float* far;          /* (batch) */
  const float* ray_offset;   /* NULL or (batch,img_size^2) */
  float* pts;                /* (batch,hw,N,3) */
  float* rays_d;             /* (batch,hw,3) */
  float* viewdirs;           /* (batch,hw,3) */
  float* z_vals;             /* (batch,hw,N) */
} c3d_raygen_params;

/* the two branches of Render.volume_integration that the v10 configs leave unused (nerf_utils.py:288-296, 309-310) */
enum { C3D_COMPOSITE_RAW_DENSITY = 1,      /* with_sdf=False: alpha = 1 - exp(-softplus(sigma) * dist) */
       C3D_COMPOSITE_FORCE_BACKGROUND = 2  /* weights[..., -1] = 1 - sum(weights[..., :-1]) */ };
typedef struct c3d_composite_params {
  int64_t n_rays;            /* total rays (batch*hw) */
  int32_t n_samples;
  int32_t n_feat;            /* feature channels, multiple of 4, <= 256 (0: no features) */
  float sigmoid_beta;        /* used when sigmoid_beta_ptr == NULL */
  int32_t flags;             /* C3D_COMPOSITE_* (forward only); 0 = the with_sdf=True branch the v10 configs use */
  const float* sigmoid_beta_ptr; /* device scalar or NULL */
  const float* rgb;          /* (n_rays,N,3) raw rgb head output */
  const float* sdf;          /* (n_rays,N): sdf, or the raw density (+ noise) with C3D_COMPOSITE_RAW_DENSITY */
  const float* features;     /* (n_rays,N,n_feat) or NULL */
  const float* z_vals;       /* (n_rays,N) */
  const float* rays_d;       /* (n_rays,3) */
  const float* pts;          /* (n_rays,N,3) */
  float* rgb_map;            /* (n_rays,3) */
  float* feature_map;        /* (n_rays,n_feat) */
  float* xyz;                /* (n_rays,3) */
  float* mask;               /* (n_rays,2) */
  float* weights;            /* optional (n_rays,N) or NULL */
  /* backward only */
  const float* g_rgb_map; const float* g_feature_map; const float* g_xyz; const float* g_mask;
  float* g_rgb; float* g_sdf; float* g_features; float* g_pts; float* g_rays_d; float* g_sigmoid_beta;
} c3d_composite_params;

/* EXTENSION (off by default; the reference has no importance resampling -- parity unpinned, oracle:
 * oracle/nerf_oracle.py::importance_depths).  Per ray: PDF over the mid-points of the N coarse depths from the interior
 * compositing weights (+1e-5), K new depths by inverse-CDF sampling, the ascending union of both and its points. */
typedef struct c3d_resample_params {
  int64_t n_rays;            /* total rays */
  int32_t n_samples;         /* N coarse samples per ray, 3..256 */
  int32_t n_importance;      /* K new samples per ray, 1..256 */
  float sigmoid_beta;        /* used when weights == NULL and sigmoid_beta_ptr == NULL */
  int32_t _pad;
  const float* sigmoid_beta_ptr; /* device scalar or NULL */
  const float* z_vals;       /* (n_rays,N) ascending coarse depths */
  const float* weights;      /* (n_rays,N) compositing weights, or NULL: derived from sdf as volume_integration does */
  const float* sdf;          /* (n_rays,N), used when weights == NULL */
  const float* rays_d;       /* (n_rays,3): required when weights == NULL (|d| scales the spacing) or pts_merged != NULL */
  const float* rays_o;       /* (n_rays,3): required when pts_merged != NULL */
  const float* u;            /* (n_rays,K) draws in [0,1], or NULL: K evenly spaced values (eval / unperturbed) */
  float* z_fine;             /* (n_rays,K) new depths in the order of u, or NULL */
  float* z_merged;           /* (n_rays,N+K) ascending union, or NULL */
  float* pts_merged;         /* (n_rays,N+K,3) rays_o + rays_d * z_merged, or NULL */
} c3d_resample_params;

int c3d_abi_version(void);
const char* c3d_last_error(void);

/* Tuning options (kernel variants for A/B runs and tests; the defaults are what ships).  The library reads the matching
 * C3D_* environment variables once, at first use; afterwards only this call changes an option -- nothing on the launch path
 * touches the environment.  Keys / values: "fwd" = "pair" | "v3"; "cluster" = "1" | "2"; "grid" = CTAs (0: one per SM);
 * "egw" = "4" | "8"; "bwd" = "tc" | "simt"; "resample" = "auto" | "warp" | "lane"; "resample_rb" = rays per block;
 * "debug" = bit mask (development builds).  Process-wide, not thread-safe against concurrent launches.
 * Replaces: nothing in the reference (its path is Python -> ATen). */
int c3d_set_option(const char* key, const char* value);

size_t c3d_packed_bytes(int32_t D);
int c3d_pack_weights(const c3d_raw_params* raw, void* packed, size_t packed_bytes, c3d_stream_t stream);

size_t c3d_workspace_bytes(const c3d_fwd_params* p);
int c3d_nerf_forward(const c3d_fwd_params* p, c3d_stream_t stream);

size_t c3d_backward_workspace_bytes(const c3d_bwd_params* p);
int c3d_nerf_backward(const c3d_bwd_params* p, c3d_stream_t stream);
/* Forward of a step that will be differentiated (flip inversion, projector_v9.py:1100-1143): same outputs as
 * c3d_nerf_forward, computed once by the save-mode kernel into a workspace of c3d_backward_workspace_bytes(p) bytes
 * (p->fwd_saved == 1) that the caller keeps until it calls c3d_nerf_backward with the same p (+ cotangents). */
int c3d_nerf_forward_save(const c3d_bwd_params* p, c3d_stream_t stream);

/* Second-order path of the eikonal regulariser (training; nerf_utils.py:220-228 with create_graph=True): given
 * g_eik = dL/dE (batch,n_rays,N,3) for E = d sdf / d pts, returns dL/d styles (p->g_styles) and dL/d parameters
 * (p->g_params); C3D_INPUT_POINTS only, the other gradient pointers and the cotangents in p are ignored. */
size_t c3d_eikonal_workspace_bytes(const c3d_bwd_params* p);
int c3d_eikonal_backward(const c3d_bwd_params* p, const float* g_eik, c3d_stream_t stream);

int c3d_raygen(const c3d_raygen_params* p, c3d_stream_t stream);
/* film: (batch, D+1, 256, 2) = (gamma, gamma*bias+beta); first: (batch,256,4); view: (batch,256,4) */
int c3d_style_prep(const void* packed, int32_t D, const float* styles, int32_t batch, float* film, float* first,
                   float* view, c3d_stream_t stream);
int c3d_composite_forward(const c3d_composite_params* p, c3d_stream_t stream);
int c3d_composite_backward(const c3d_composite_params* p, c3d_stream_t stream);

/* Inverse-CDF importance resampling; at least one output must be non-NULL.  One kernel launch. */
int c3d_sample_pdf(const c3d_resample_params* p, c3d_stream_t stream);

/* ---- flip-inversion driver helpers (projector_v9.py:998-1166): keep the optimisation step from being launch-bound ---- */

/* Camera.generate_camera_params, `locations` mode (nerf_utils.py:369-378, 412-436): azim / elev (n) -> cam_poses (n,3,4),
 * focal / near / far (n); fov_ang: device (n) degrees, or NULL to use fov_scalar.  jac (n,12,2) or NULL receives
 * d cam_poses / d (azim, elev) (forward-mode), so the backward is a 12x2 contraction per camera.  One launch. */
int c3d_camera_params(const float* azim, const float* elev, int32_t n, int32_t img_size, const float* fov_ang, float fov_scalar,
                      float dist_radius, float* cam_poses, float* focal, float* near, float* far, float* jac,
                      c3d_stream_t stream);

/* torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (defaults: no weight decay, no amsgrad) for up to 8 small fp32 tensors in
 * two groups (group 0: latents, group 1: cameras; each with its own gradient-norm clipping and its own learning rate), one
 * launch.  lr: two DEVICE scalars; step: DEVICE scalar counting the updates done so far (incremented by the call). */
typedef struct c3d_adam_params {
  int32_t n_tensors;                     /* 1..8 */
  int32_t group[8];                      /* 0 or 1 */
  int64_t numel[8];
  float* param[8];
  const float* grad[8];
  float* exp_avg[8];
  float* exp_avg_sq[8];
  const float* lr[2];
  float* step;
  float beta1, beta2, eps, max_norm;     /* max_norm <= 0: no clipping */
  float* grad_norm;                      /* optional (2): gradient norms of the groups before clipping, or NULL */
} c3d_adam_params;
int c3d_adam_clip_step(const c3d_adam_params* p, c3d_stream_t stream);

/* Launch statistics of the most recent c3d_nerf_forward on this thread (kernel launches it made). */
int c3d_last_launch_count(void);

/* Self-test hook: one 128x{N}x{K} tcgen05 tile product through the library's own shared-memory
 * layouts; a: (128,K) bf16 bits row-major, b: (N,K) bf16 bits row-major, d: (128,N) fp32.
 * variant selects the operand roles exercised by the fused kernel (0: layer tile, 1: transposed
 * view-layer tile, 2: rgb-head tile). */
int c3d_umma_selftest(const uint16_t* a, const uint16_t* b, float* d, int32_t N, int32_t K,
                      int32_t variant, c3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* C3D_ABI_H_ */
