"""CPU oracle for the CIPS-3D++ NeRF branch -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy (float32) restatement of the reference's single-pass FiLM-SIREN / SDF
volume renderer.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
path (``cips3dpp_b200``) never does and fails loudly when the CUDA library is absent.

Parity pin: the reference ships no golden vectors or asserting tests for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference's own
modules imported live in the build container (``tests/golden/make_golden.py`` writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them).

Every function cites the reference lines it follows (paths relative to the
reference checkout, ``exp/cips3d/...``).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _f32(x):
    return np.asarray(x, dtype=F32)


def sigmoid(x):
    x = _f32(x)
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def l2_normalize(x, eps):
    """torch.nn.functional.normalize(p=2, dim=-1): x / max(||x||, eps)."""
    x = _f32(x)
    n = np.sqrt(np.sum(x * x, axis=-1, keepdims=True, dtype=F32), dtype=F32)
    return (x / np.maximum(n, F32(eps))).astype(F32)


# --------------------------------------------------------------------------------------
# Camera  (nerf_utils.py:343-436, `locations` and `sweep` modes)
# --------------------------------------------------------------------------------------
def generate_camera_params(locations, img_size, fov_ang=6.0, dist_radius=0.12):
    """nerf_utils.py:369-378 (intrinsics from `locations`) and :412-436 (extrinsics).

    locations: (b, 2) [azim, elev];  fov_ang: scalar or (b, 1).
    returns extrinsics (b,3,4), focal (b,1,1), near (b,1,1), far (b,1,1), viewpoint (b,2)
    """
    loc = _f32(locations)
    azim = loc[:, 0:1]
    elev = loc[:, 1:2]
    b = azim.shape[0]
    dist = np.ones((b, 1), F32)
    near = (dist - F32(dist_radius))[..., None]
    far = (dist + F32(dist_radius))[..., None]
    fov_angle = (_f32(fov_ang) * np.ones((b, 1), F32)).reshape(-1, 1) * F32(np.pi) / F32(180)
    focal = (F32(0.5) * F32(img_size) / np.tan(fov_angle, dtype=F32))[..., None]
    viewpoint = np.concatenate([azim, elev], 1)

    x = np.cos(elev) * np.sin(azim)
    y = np.sin(elev)
    z = np.cos(elev) * np.cos(azim)
    camera_dir = np.stack([x, y, z], axis=1).reshape(-1, 3).astype(F32)
    camera_loc = dist * camera_dir
    up = np.array([[0, 1, 0]], F32) * np.ones_like(dist)
    z_axis = l2_normalize(camera_dir, 1e-5)
    x_axis = l2_normalize(np.cross(up, z_axis), 1e-5)
    y_axis = l2_normalize(np.cross(z_axis, x_axis), 1e-5)
    # degenerate-x fix, nerf_utils.py:428-431
    is_close = np.all(np.abs(x_axis) <= F32(5e-3), axis=1, keepdims=True)
    if is_close.any():
        repl = l2_normalize(np.cross(y_axis, z_axis), 1e-5)
        x_axis = np.where(is_close, repl, x_axis)
    R = np.stack([x_axis, y_axis, z_axis], axis=1)          # rows are the axes
    T = camera_loc[:, :, None]
    extrinsics = np.concatenate([R.transpose(0, 2, 1), T], -1).astype(F32)
    return extrinsics, focal.astype(F32), near.astype(F32), far.astype(F32), viewpoint


def sweep_locations(batch, azim_range=0.3, elev_range=0.15, elev_u=None):
    """nerf_utils.py:379-386 -- 8 yaw steps per latent, one elevation draw per latent.

    elev_u: (batch,) uniform(0,1) draws (the reference uses torch.rand); returns (batch*8, 2).
    """
    k = np.arange(8, dtype=F32)
    azim = np.tile((-F32(azim_range) + (F32(2 * azim_range) / F32(7)) * k).reshape(-1, 1), (batch, 1))
    u = np.zeros(batch, F32) + F32(0.5) if elev_u is None else _f32(elev_u)
    elev = (-F32(elev_range) + F32(2 * elev_range) * np.repeat(u.reshape(batch, 1), 8, axis=1)).reshape(-1, 1)
    return np.concatenate([azim, elev], 1).astype(F32)


# --------------------------------------------------------------------------------------
# Rays / samples / points  (nerf_utils.py:17-218)
# --------------------------------------------------------------------------------------
def get_rays_in_world(focal, img_size, c2w, static_viewdirs=False):
    """nerf_utils.py:17-66. focal (b,1,1), c2w (b,3,4) -> rays_o, rays_d, viewdirs (b,h,w,3)."""
    focal = _f32(focal)
    c2w = _f32(c2w)
    S = int(img_size)
    lin = np.linspace(0.5, S - 0.5, S, dtype=F32)
    y, x = np.meshgrid(lin, lin, indexing="ij")
    x = x[None]
    y = y[None]
    b = focal.shape[0]
    d_cam = np.stack(
        [(x - F32(S * 0.5)) / focal, -(y - F32(S * 0.5)) / focal, -np.ones((b, S, S), F32)], axis=-1
    ).astype(F32)
    rays_d = np.sum(d_cam[..., None, :] * c2w[:, None, None, :3, :3], axis=-1, dtype=F32)
    rays_o = np.broadcast_to(c2w[:, None, None, :3, -1], rays_d.shape).astype(F32)
    viewdirs = l2_normalize(d_cam if static_viewdirs else rays_d, 1e-12)
    return rays_o, rays_d.astype(F32), viewdirs


def get_z_vals(near, far, b, h, w, N_samples, t_rand=None):
    """nerf_utils.py:68-121 with offset_sampling=True.

    t_rand: None (perturb=False) or (b,h,w,1) uniform draws (perturb=True, :105-110,119).
    """
    near = _f32(near).reshape(b, 1, 1, 1) * np.ones((b, h, w, 1), F32)
    far = _f32(far).reshape(b, 1, 1, 1) * np.ones((b, h, w, 1), F32)
    t_vals = np.linspace(0.0, 1.0 - 1.0 / N_samples, N_samples, dtype=F32).reshape(1, 1, 1, -1)
    z_vals = near * (F32(1.0) - t_vals) + far * t_vals
    if t_rand is not None:
        upper = np.concatenate([z_vals[..., 1:], far], -1)
        lower = z_vals
        z_vals = lower + (upper - lower) * _f32(t_rand)
    return z_vals.astype(F32)


def get_points(rays_o, rays_d, z_vals):
    """nerf_utils.py:135-170."""
    return (_f32(rays_o)[..., None, :] + _f32(rays_d)[..., None, :] * _f32(z_vals)[..., None]).astype(F32)


def normalize_points(pts, near, far):
    """nerf_utils.py:123-133."""
    pts = _f32(pts)
    shape = [-1] + [1] * (pts.ndim - 1)
    return (pts * F32(2) / (_f32(far) - _f32(near)).reshape(shape)).astype(F32)


def prepare_nerf_inputs(focal, img_size, cam_poses, near, far, N_samples, t_rand=None, static_viewdirs=False):
    """nerf_utils.py:172-218 -> pts (b,h,w,N,3), rays_d, viewdirs (b,h,w,3), z_vals (b,h,w,N)."""
    rays_o, rays_d, viewdirs = get_rays_in_world(focal, img_size, cam_poses, static_viewdirs)
    b, h, w, _ = rays_d.shape
    z_vals = get_z_vals(near, far, b, h, w, N_samples, t_rand)
    pts = get_points(rays_o, rays_d, z_vals)
    return pts, rays_d, viewdirs, z_vals


# --------------------------------------------------------------------------------------
# FiLM-SIREN point MLP  (volume_renderer.py:15-160)
# --------------------------------------------------------------------------------------
def linear_layer(x, weight, bias, std_init=1.0, bias_init=0.0):
    """volume_renderer.py:32-35:  std_init * (x W^T + b) + bias_init."""
    return (F32(std_init) * (_f32(x) @ _f32(weight).T + _f32(bias)) + F32(bias_init)).astype(F32)


def film_params(style, p):
    """volume_renderer.py:66-67: gamma = 15*Linear(w)+30, beta = 0.25*Linear(w).  style (b,256)."""
    gamma = linear_layer(style, p["gamma.weight"], p["gamma.bias"], 15.0, 30.0)
    beta = linear_layer(style, p["beta.weight"], p["beta.bias"], 0.25, 0.0)
    return gamma, beta


def film_siren(x, style, p):
    """volume_renderer.py:70-85: sin(gamma * (x W^T + b) + beta); x is (b, ..., Cin)."""
    out = _f32(x) @ _f32(p["weight"]).T + _f32(p["bias"])
    gamma, beta = film_params(style, p)
    shape = [gamma.shape[0]] + [1] * (out.ndim - 2) + [-1]
    return np.sin(gamma.reshape(shape) * out + beta.reshape(shape), dtype=F32)


def _sub(params, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


def points_forward(params, pts_n, viewdirs_pt, styles):
    """volume_renderer.py:133-160.  pts_n (b,...,3) normalised, viewdirs_pt (b,...,3), styles (b,D+1,256).

    returns rgb (b,...,3), sdf (b,...,1), features (b,...,256)
    """
    D = sum(1 for k in params if k.startswith("network.pts_linears.") and k.endswith(".weight")
            and k.count(".") == 3)
    h = _f32(pts_n)
    for i in range(D):
        h = film_siren(h, styles[:, i], _sub(params, f"network.pts_linears.{i}."))
    sdf = linear_layer(h, params["network.sigma_linear.weight"], params["network.sigma_linear.bias"])
    hv = np.concatenate([h, _f32(viewdirs_pt)], -1)
    feat = film_siren(hv, styles[:, -1], _sub(params, "network.views_linears."))
    rgb = linear_layer(feat, params["network.rgb_linear.weight"], params["network.rgb_linear.bias"])
    return rgb, sdf, feat


# --------------------------------------------------------------------------------------
# Volume integration  (nerf_utils.py:230-338, with_sdf=True branch)
# --------------------------------------------------------------------------------------
def volume_integration(rgb, sdf, features, z_vals, rays_d, pts, sigmoid_beta, with_sdf=True, force_background=False):
    """nerf_utils.py:230-338. Shapes (..., n, c) / (..., n) / (..., 3).  `with_sdf=False`: `sdf` holds the raw density
    (softplus branch, :288-296, noise-free); `force_background`: :309-310."""
    rgb, sdf, z_vals, rays_d, pts = map(_f32, (rgb, sdf, z_vals, rays_d, pts))
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    d_norm = np.sqrt(np.sum(rays_d * rays_d, axis=-1, keepdims=True, dtype=F32), dtype=F32)
    dists = np.concatenate([dists, np.broadcast_to(F32(1e10), d_norm.shape)], -1) * d_norm
    if with_sdf:
        beta = F32(np.asarray(sigmoid_beta, F32).reshape(-1)[0])
        sigma = sigmoid(-sdf / beta) / beta
    else:
        sigma = np.where(sdf > F32(20), sdf, np.log1p(np.exp(np.minimum(sdf, F32(20)), dtype=F32), dtype=F32)).astype(F32)
    alpha = (F32(1) - np.exp(-sigma * dists[..., None], dtype=F32)).astype(F32)
    ones = np.ones_like(alpha[..., :1, :])
    vis = np.cumprod(np.concatenate([ones, F32(1) - alpha + F32(1e-10)], axis=-2), axis=-2, dtype=F32)[..., :-1, :]
    weights = (alpha * vis).astype(F32)
    if force_background:
        weights = weights.copy()
        weights[..., -1, :] = F32(1) - np.sum(weights[..., :-1, :], axis=-2, dtype=F32)
    rgb_map = (F32(-1) + F32(2) * np.sum(weights * sigmoid(rgb), axis=-2, dtype=F32)).astype(F32)
    feature_map = None if features is None else np.sum(weights * _f32(features), axis=-2, dtype=F32)
    xyz = np.sum(weights * pts, axis=-2, dtype=F32)
    mask = weights[..., -1, :]
    depth = -np.sqrt(np.sum(xyz * xyz, axis=-1, keepdims=True, dtype=F32), dtype=F32)
    mask = np.concatenate([mask, depth], -1).astype(F32)
    return rgb_map, feature_map, xyz, mask, weights


def renderer_forward(params, pts, rays_d, viewdirs, z_vals, near, far, styles):
    """VolumeFeatureRenderer.forward, volume_renderer.py:192-283 (return_eikonal=False).

    pts (b,hw,N,3) world-space, rays_d/viewdirs (b,hw,3), z_vals (b,hw,N), near/far (b,1,1),
    styles (b,D+1,256).  returns rgb_map (b,hw,3), feature_map (b,hw,256), sdf (b,hw,N,1),
    mask (b,hw,2), xyz (b,hw,3).
    """
    pts = _f32(pts)
    pts_n = normalize_points(pts, near, far)
    vd = np.broadcast_to(_f32(viewdirs)[..., None, :], pts_n.shape)      # run_network, :298-299
    rgb, sdf, feat = points_forward(params, pts_n, vd, _f32(styles))
    rgb_map, feature_map, xyz, mask, _ = volume_integration(
        rgb, sdf, feat, z_vals, rays_d, pts, params["sigmoid_beta"])
    return rgb_map, feature_map, sdf, mask, xyz


def render(params, cam_poses, focal, near, far, styles, img_size=64, N_samples=24,
           static_viewdirs=False, t_rand=None, ray_idx=None):
    """prepare_nerf_inputs + flatten (model_v3.py:930-951) + renderer forward.

    ray_idx: optional 1-D index array selecting a subset of the hw rays (rays are independent).
    """
    pts, rays_d, viewdirs, z_vals = prepare_nerf_inputs(
        focal, img_size, cam_poses, near, far, N_samples, t_rand, static_viewdirs)
    b = pts.shape[0]
    pts = pts.reshape(b, -1, N_samples, 3)
    rays_d = rays_d.reshape(b, -1, 3)
    viewdirs = viewdirs.reshape(b, -1, 3)
    z_vals = z_vals.reshape(b, -1, N_samples)
    if ray_idx is not None:
        pts, rays_d, viewdirs, z_vals = pts[:, ray_idx], rays_d[:, ray_idx], viewdirs[:, ray_idx], z_vals[:, ray_idx]
    out = renderer_forward(params, pts, rays_d, viewdirs, z_vals, near, far, styles)
    return out + (z_vals,)


# --------------------------------------------------------------------------------------
# Parameter construction with the reference's init distributions (volume_renderer.py:16-30, 56-67)
# --------------------------------------------------------------------------------------
def init_params(D=8, W=256, style_dim=256, seed=0):
    """Random-init parameters with the reference's distributions (NOT its RNG stream)."""
    rng = np.random.default_rng(seed)

    def uni(shape, a):
        return rng.uniform(-a, a, size=shape).astype(F32)

    def lin(prefix, out_dim, in_dim, p):
        std = np.sqrt(2.0 / (1 + 0.2 ** 2)) / np.sqrt(in_dim)          # kaiming_normal_(a=0.2), :24-25
        p[prefix + "weight"] = (0.25 * rng.normal(0, std, size=(out_dim, in_dim))).astype(F32)
        p[prefix + "bias"] = uni((out_dim,), np.sqrt(1 / in_dim))

    def film(prefix, cin, p, first=False):
        p[prefix + "weight"] = uni((W, cin), 1 / 3 if first else np.sqrt(6 / cin) / 25)
        p[prefix + "bias"] = uni((W,), np.sqrt(1 / cin))
        lin(prefix + "gamma.", W, style_dim, p)
        lin(prefix + "beta.", W, style_dim, p)

    p = {"sigmoid_beta": np.full((1,), 0.1, F32)}
    for i in range(D):
        film(f"network.pts_linears.{i}.", 3 if i == 0 else W, p, first=(i == 0))
    film("network.views_linears.", W + 3, p)
    p["network.rgb_linear.weight"] = uni((3, W), np.sqrt(6 / W) / 25)
    p["network.rgb_linear.bias"] = uni((3,), np.sqrt(1 / W))
    p["network.sigma_linear.weight"] = uni((1, W), np.sqrt(6 / W) / 25)
    p["network.sigma_linear.bias"] = uni((1,), np.sqrt(1 / W))
    return p


def flops_per_point(D, W=256):
    """algorithmic FLOPs per sample point."""
    return 2 * (3 * W + (D - 1) * W * W + W * 1 + (W + 3) * W + W * 3)


# --------------------------------------------------------------------------------------
# EXTENSION -- inverse-CDF importance resampling + fine pass.  **PARITY UNPINNED**
#
# BASELINE.json's north star names a `sample_pdf` / fine pass, but the reference has none
# (`grep -rE "sample_pdf|searchsorted|importance"` over /root/reference is empty;
# CIPS-3D++ renders in one pass, nerf_utils.py:172-218).  There is therefore no reference
# output to pin these functions to.  They restate the canonical published algorithm
# (Mildenhall et al., "NeRF", ECCV 2020, section 5.2 "hierarchical volume sampling", as used by
# pi-GAN, the code base CIPS-3D descends from) and are cross-checked in tests/test_resample_cpu.py
# against an independent torch.searchsorted restatement.  The extension is off by default and
# never changes the single-pass outputs.
# --------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_importance, u=None):
    """Inverse-transform sampling of a piecewise-constant PDF.

    bins (..., M) ascending bin edges, weights (..., M-1) non-negative bin weights, u (..., K) draws in [0,1]
    (None: K evenly spaced values 0..1 = the unperturbed / eval mode).  Returns samples (..., K).
    Steps: weights + 1e-5 -> pdf -> cdf with a leading 0 -> searchsorted(cdf, u, right=True) ->
    linear interpolation inside the bin, a CDF step below 1e-5 is treated as 1.
    """
    bins, weights = _f32(bins), _f32(weights)
    K = int(n_importance)
    w = weights + F32(1e-5)
    pdf = (w / np.sum(w, axis=-1, keepdims=True, dtype=F32)).astype(F32)
    cdf = np.cumsum(pdf, axis=-1, dtype=F32)
    cdf = np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)              # (..., M)
    M = cdf.shape[-1]
    if u is None:
        u = np.broadcast_to(np.linspace(0.0, 1.0, K, dtype=F32), cdf.shape[:-1] + (K,))
    u = _f32(u)
    inds = np.sum(cdf[..., None, :] <= u[..., :, None], axis=-1)              # searchsorted(right=True)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, M - 1)
    cdf_b, cdf_a = np.take_along_axis(cdf, below, -1), np.take_along_axis(cdf, above, -1)
    bin_b, bin_a = np.take_along_axis(bins, below, -1), np.take_along_axis(bins, above, -1)
    denom = cdf_a - cdf_b
    denom = np.where(denom < F32(1e-5), F32(1), denom).astype(F32)
    t = ((u - cdf_b) / denom).astype(F32)
    return (bin_b + t * (bin_a - bin_b)).astype(F32)


def compositing_weights(sdf, z_vals, rays_d, sigmoid_beta):
    """w_k = alpha_k T_k exactly as volume_integration computes them (nerf_utils.py:267-307): the PDF the fine
    pass samples from when the density comes from an SDF.  sdf, z_vals (..., N); rays_d (..., 3)."""
    sdf, z_vals, rays_d = _f32(sdf), _f32(z_vals), _f32(rays_d)
    beta = F32(np.asarray(sigmoid_beta, F32).reshape(-1)[0])
    d_norm = np.sqrt(np.sum(rays_d * rays_d, axis=-1, keepdims=True, dtype=F32), dtype=F32)
    dists = np.concatenate([z_vals[..., 1:] - z_vals[..., :-1], np.broadcast_to(F32(1e10), d_norm.shape)], -1) * d_norm
    sigma = sigmoid(-sdf / beta) / beta
    alpha = (F32(1) - np.exp(-sigma * dists, dtype=F32)).astype(F32)
    vis = np.cumprod(np.concatenate([np.ones_like(alpha[..., :1]), F32(1) - alpha + F32(1e-10)], -1), -1,
                     dtype=F32)[..., :-1]
    return (alpha * vis).astype(F32)


def importance_depths(z_vals, weights, n_importance, u=None):
    """Fine-pass depths of one ray set: PDF over the mid-points of the coarse depths from the interior weights
    (first and last dropped), K new depths, and the ascending union of coarse and new depths.
    z_vals, weights (..., N) -> z_fine (..., K), z_merged (..., N+K)."""
    z_vals, weights = _f32(z_vals), _f32(weights)
    mids = (F32(0.5) * (z_vals[..., 1:] + z_vals[..., :-1])).astype(F32)
    z_fine = sample_pdf(mids, weights[..., 1:-1], n_importance, u)
    z_merged = np.sort(np.concatenate([z_vals, z_fine], -1), axis=-1)
    return z_fine, z_merged


def render_hierarchical(params, cam_poses, focal, near, far, styles, img_size=64, N_samples=24, N_importance=24,
                        static_viewdirs=False, u=None, ray_idx=None):
    """Coarse pass -> importance_depths -> second pass over the merged depths (same network: pi-GAN style).
    Returns the fine pass's (rgb_map, feature_map, sdf, mask, xyz, z_merged) and the coarse pass's tuple."""
    coarse = render(params, cam_poses, focal, near, far, styles, img_size, N_samples, static_viewdirs, None, ray_idx)
    z_vals = coarse[5]
    rays_o, rays_d, viewdirs = get_rays_in_world(focal, img_size, cam_poses, static_viewdirs)
    b = z_vals.shape[0]
    rays_o, rays_d, viewdirs = (t.reshape(b, -1, 3) for t in (rays_o, rays_d, viewdirs))
    if ray_idx is not None:
        rays_o, rays_d, viewdirs = rays_o[:, ray_idx], rays_d[:, ray_idx], viewdirs[:, ray_idx]
    w = compositing_weights(coarse[2][..., 0], z_vals, rays_d, params["sigmoid_beta"])
    _, z_m = importance_depths(z_vals, w, N_importance, u)
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * z_m[..., None]).astype(F32)
    fine = renderer_forward(params, pts, rays_d, viewdirs, z_m, near, far, styles)
    return fine + (z_m,), coarse
