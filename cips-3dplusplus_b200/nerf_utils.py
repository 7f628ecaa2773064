"""Host-side mirrors of the reference's static helpers `nerf_utils.Render` / `nerf_utils.Camera`
(exp/cips3d/nerf_utils.py) so callers that use them directly keep working on the new path.

* `Render.prepare_nerf_inputs`      -> c3d_raygen (CUDA)
* `Render.volume_integration`       -> c3d_composite_forward (CUDA)
* `Camera.generate_camera_params` stays PyTorch glue on purpose: it is ~30 scalar ops per image and must
  remain differentiable w.r.t. azimuth/elevation for flip inversion (projector_v9.py:209).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import _abi


def _c(t, *shape):
    return t.to(torch.float32).reshape(*shape).contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Render:
    # The three helpers below are plain PyTorch (a handful of elementwise ops on (b, h, w, .) tensors): the fused kernel
    # generates rays and samples itself, these exist for callers of the reference's static API (mlp_init_pass, apps).
    @staticmethod
    def get_rays_in_world(focal, img_size, c2w, static_viewdirs=False):
        """nerf_utils.py:17-66 -> rays_o, rays_d, viewdirs, each (b, h, w, 3); rays_d is not normalised."""
        b, S, dev = c2w.shape[0], int(img_size), c2w.device
        lin = torch.linspace(0.5, S - 0.5, S, device=dev)
        yy, xx = torch.meshgrid(lin, lin, indexing="ij")
        f = focal.reshape(b, 1, 1).to(torch.float32)
        d_cam = torch.stack([(xx[None] - S / 2) / f, -(yy[None] - S / 2) / f, -torch.ones(b, S, S, device=dev)], -1)
        rays_d = (d_cam[..., None, :] * c2w[:, None, None, :3, :3]).sum(-1)
        rays_o = c2w[:, None, None, :3, 3].expand(rays_d.shape)
        viewdirs = F.normalize(d_cam if static_viewdirs else rays_d, dim=-1)
        return rays_o, rays_d, viewdirs

    @staticmethod
    def get_z_vals(near, far, rays_d, N_samples, perturb=True, offset_sampling=True):
        """nerf_utils.py:68-121 -> z_vals (b, h, w, N): offset sampling (one draw per ray) or stratified sampling."""
        b, h, w, _ = rays_d.shape
        dev = rays_d.device
        near = near.reshape(b, 1, 1, 1).to(torch.float32).expand(b, h, w, 1)
        far = far.reshape(b, 1, 1, 1).to(torch.float32).expand(b, h, w, 1)
        hi = 1.0 - 1.0 / N_samples if offset_sampling else 1.0
        t = torch.linspace(0.0, hi, N_samples, device=dev).view(1, 1, 1, -1)
        z = near * (1.0 - t) + far * t
        if perturb:
            if offset_sampling:
                upper, lower = torch.cat([z[..., 1:], far], -1), z.detach()
                t_rand = torch.rand(b, h, w, 1, device=dev)
            else:
                mids = 0.5 * (z[..., 1:] + z[..., :-1])
                upper, lower = torch.cat([mids, z[..., -1:]], -1), torch.cat([z[..., :1], mids], -1)
                t_rand = torch.rand(z.shape, device=dev)
            z = lower + (upper - lower) * t_rand
        return z

    @staticmethod
    def get_points(rays_o, rays_d, z_vals):
        """nerf_utils.py:135-170 -> pts (b, h, w, N, 3)."""
        return rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)

    @staticmethod
    def prepare_nerf_inputs(focal, img_size, cam_poses, near, far, N_samples, perturb, static_viewdirs=False,
                            **kwargs):
        """nerf_utils.py:172-218 -> pts (b,h,w,N,3), rays_d (b,h,w,3), viewdirs (b,h,w,3), z_vals (b,h,w,N)."""
        if cam_poses.device.type != "cuda":
            raise RuntimeError("Render.prepare_nerf_inputs needs CUDA tensors (no CPU fallback)")
        lib = _abi.load()
        b, S, N, dev = cam_poses.shape[0], int(img_size), int(N_samples), cam_poses.device
        f = dict(dtype=torch.float32, device=dev)
        pts = torch.empty(b, S, S, N, 3, **f)
        rays_d = torch.empty(b, S, S, 3, **f)
        viewdirs = torch.empty(b, S, S, 3, **f)
        z_vals = torch.empty(b, S, S, N, **f)
        off = _c(torch.rand(b, S, S, 1, device=dev), b, S * S) if perturb else None     # nerf_utils.py:110
        keep = [_c(cam_poses, b, 3, 4), _c(focal, b), _c(near, b), _c(far, b)]
        P = _abi.RaygenParams()
        P.batch, P.img_size, P.n_samples, P.static_viewdirs = b, S, N, int(bool(static_viewdirs))
        P.cam_poses, P.focal, P.near, P.far = (t.data_ptr() for t in keep)
        P.ray_offset = None if off is None else off.data_ptr()
        P.pts, P.rays_d, P.viewdirs, P.z_vals = pts.data_ptr(), rays_d.data_ptr(), viewdirs.data_ptr(), z_vals.data_ptr()
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_raygen(P, _stream()), "c3d_raygen")
        return pts, rays_d, viewdirs, z_vals

    @staticmethod
    def normalize_points(pts, near, far):
        """nerf_utils.py:123-133 (one multiply; kept as a torch op because callers pass autograd tensors)."""
        shape = [-1] + [1] * (pts.dim() - 1)
        return pts * 2 / (far - near).view(*shape)

    @staticmethod
    def volume_integration(rgb, sdf, features, z_vals, rays_d, pts, with_sdf=True, sigmoid_beta=None,
                           return_eikonal=False, raw_noise_std=0.0, force_background=False):
        """nerf_utils.py:230-338 -> rgb_map, feature_map, xyz, mask, eikonal_term(None).  Forward only (no autograd).
        `with_sdf=False` takes `sdf` as the raw density (softplus, plus N(0, raw_noise_std) noise when asked);
        `force_background` puts the remaining weight on the last sample.  For the eikonal term use
        `NerfBranch.forward(return_eikonal=True)` (or `Render.get_eikonal_term`)."""
        if return_eikonal:
            raise NotImplementedError("use NerfBranch.forward(return_eikonal=True) or Render.get_eikonal_term")
        if rgb.device.type != "cuda":
            raise RuntimeError("Render.volume_integration needs CUDA tensors (no CPU fallback)")
        lib = _abi.load()
        lead, N, dev = rgb.shape[:-2], rgb.shape[-2], rgb.device
        R = int(math.prod(lead))
        f = dict(dtype=torch.float32, device=dev)
        C = 0 if features is None else features.shape[-1]
        keep = [_c(rgb, R, N, 3), _c(sdf, R, N), _c(z_vals, R, N), _c(rays_d, R, 3), _c(pts, R, N, 3)]
        feats = None if features is None else _c(features, R, N, C)
        rgb_map, xyz, mask = torch.empty(R, 3, **f), torch.empty(R, 3, **f), torch.empty(R, 2, **f)
        feature_map = None if features is None else torch.empty(R, C, **f)
        P = _abi.CompositeParams()
        P.n_rays, P.n_samples, P.n_feat = R, N, C
        P.flags = (0 if with_sdf else _abi.COMPOSITE_RAW_DENSITY) | (_abi.COMPOSITE_FORCE_BACKGROUND if force_background else 0)
        if not with_sdf:
            if raw_noise_std > 0:
                keep[1] = keep[1] + torch.randn_like(keep[1]) * raw_noise_std        # nerf_utils.py:292-294
            sigmoid_beta = 1.0 if sigmoid_beta is None else sigmoid_beta            # unused by this branch
        if torch.is_tensor(sigmoid_beta):
            sb = _c(sigmoid_beta, 1)
            keep.append(sb)
            P.sigmoid_beta_ptr = sb.data_ptr()
        else:
            P.sigmoid_beta = float(sigmoid_beta)
        P.rgb, P.sdf, P.z_vals, P.rays_d, P.pts = (t.data_ptr() for t in keep[:5])
        P.features = None if feats is None else feats.data_ptr()
        P.rgb_map, P.xyz, P.mask = rgb_map.data_ptr(), xyz.data_ptr(), mask.data_ptr()
        P.feature_map = None if feature_map is None else feature_map.data_ptr()
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_composite_forward(P, _stream()), "c3d_composite_forward")
        rs = lambda t: None if t is None else t.reshape(*lead, t.shape[-1])
        return rs(rgb_map), rs(feature_map), rs(xyz), rs(mask), None


    @staticmethod
    def get_eikonal_term(pts, sdf):
        """nerf_utils.py:220-228: d sdf / d pts by autograd, for `sdf` produced from `pts` (requires_grad) by
        `NerfBranch.forward`.  First order only -- the custom backward is not differentiable a second time; the training
        regulariser (a loss on this term) goes through `NerfBranch.forward(return_eikonal=True)` instead, whose double
        backward is a dedicated kernel (c3d_eikonal_backward)."""
        return torch.autograd.grad(outputs=sdf, inputs=pts, grad_outputs=torch.ones_like(sdf), retain_graph=True,
                                   only_inputs=True)[0]

    @staticmethod
    def importance_depths(z_vals, N_importance, weights=None, sdf=None, rays_d=None, sigmoid_beta=None, rays_o=None,
                          u=None, perturb=False, return_pts=False):
        """EXTENSION (no reference counterpart: the reference renders in one pass) -> c3d_sample_pdf (CUDA).

        Canonical NeRF hierarchical sampling: PDF over the mid-points of the coarse depths from the interior
        compositing weights, `N_importance` new depths by inverse-CDF sampling (`u`: draws in [0,1] of shape
        (..., N_importance); None -> evenly spaced, or uniform random when `perturb`), and the ascending union.
        The weights are either given (`weights`, (..., N)) or derived in-kernel from `sdf` (..., N[,1]), `rays_d` and
        `sigmoid_beta` exactly as `volume_integration` derives them.  Nothing here is differentiable (the new depths
        are constants of the fine pass, as in NeRF / pi-GAN).  Returns dict(z_fine, z_merged[, pts])."""
        if z_vals.device.type != "cuda":
            raise RuntimeError("Render.importance_depths needs CUDA tensors (no CPU fallback)")
        lib = _abi.load()
        lead, N, K, dev = z_vals.shape[:-1], z_vals.shape[-1], int(N_importance), z_vals.device
        R = int(math.prod(lead))
        f = dict(dtype=torch.float32, device=dev)
        keep = {"z": _c(z_vals.detach(), R, N)}
        P = _abi.ResampleParams()
        P.n_rays, P.n_samples, P.n_importance = R, N, K
        if weights is not None:
            keep["w"] = _c(weights.detach(), R, N)
            P.weights = keep["w"].data_ptr()
        else:
            if sdf is None or rays_d is None or sigmoid_beta is None:
                raise ValueError("give either `weights`, or `sdf`, `rays_d` and `sigmoid_beta`")
            keep["sdf"] = _c(sdf.detach(), R, N)
            P.sdf = keep["sdf"].data_ptr()
            if torch.is_tensor(sigmoid_beta):
                keep["sb"] = _c(sigmoid_beta.detach(), 1)
                P.sigmoid_beta_ptr = keep["sb"].data_ptr()
            else:
                P.sigmoid_beta = float(sigmoid_beta)
        if rays_d is not None:
            keep["d"] = _c(rays_d.detach(), R, 3)
            P.rays_d = keep["d"].data_ptr()
        if return_pts:
            if rays_o is None or rays_d is None:
                raise ValueError("return_pts needs rays_o and rays_d")
            keep["o"] = _c(rays_o.detach().expand(*lead, 3), R, 3)
            P.rays_o = keep["o"].data_ptr()
        if u is None and perturb:
            u = torch.rand(R, K, **f)
        if u is not None:
            keep["u"] = _c(u.detach(), R, K)
            P.u = keep["u"].data_ptr()
        P.z_vals = keep["z"].data_ptr()
        z_fine, z_merged = torch.empty(R, K, **f), torch.empty(R, N + K, **f)
        pts = torch.empty(R, N + K, 3, **f) if return_pts else None
        P.z_fine, P.z_merged = z_fine.data_ptr(), z_merged.data_ptr()
        P.pts_merged = None if pts is None else pts.data_ptr()
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_sample_pdf(P, _stream()), "c3d_sample_pdf")
        out = dict(z_fine=z_fine.reshape(*lead, K), z_merged=z_merged.reshape(*lead, N + K))
        if pts is not None:
            out["pts"] = pts.reshape(*lead, N + K, 3)
        return out


class _CameraFn(torch.autograd.Function):
    """(azim, elev) -> camera-to-world poses in one launch (c3d_camera_params); the kernel also returns the Jacobian
    d pose / d (azim, elev) by forward-mode differentiation, so the backward is a 12x2 contraction per camera."""

    @staticmethod
    def forward(ctx, azim, elev, img_size, fov, dist_radius):
        lib = _abi.load()
        n, dev = azim.numel(), azim.device
        f = dict(dtype=torch.float32, device=dev)
        az, el = _c(azim, n), _c(elev, n)
        pose, jac = torch.empty(n, 3, 4, **f), torch.empty(n, 12, 2, **f)
        focal, near, far = torch.empty(n, 1, 1, **f), torch.empty(n, 1, 1, **f), torch.empty(n, 1, 1, **f)
        fov_t = _c(fov, n) if torch.is_tensor(fov) else None
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_camera_params(az.data_ptr(), el.data_ptr(), n, int(img_size),
                                             None if fov_t is None else fov_t.data_ptr(),
                                             0.0 if fov_t is not None else float(fov), float(dist_radius), pose.data_ptr(),
                                             focal.data_ptr(), near.data_ptr(), far.data_ptr(), jac.data_ptr(), _stream()),
                       "c3d_camera_params")
        ctx.save_for_backward(jac)
        ctx.shapes = (azim.shape, elev.shape)
        ctx.mark_non_differentiable(focal, near, far)
        return pose, focal, near, far

    @staticmethod
    def backward(ctx, g_pose, g_focal, g_near, g_far):
        (jac,) = ctx.saved_tensors
        g = torch.einsum("nk,nkj->nj", g_pose.reshape(-1, 12).to(torch.float32), jac)
        return g[:, 0].reshape(ctx.shapes[0]), g[:, 1].reshape(ctx.shapes[1]), None, None, None


class Camera:
    @staticmethod
    def generate_camera_params(img_size, device, batch=1, locations=None, sweep=False, uniform=False,
                               azim_range=0.3, elev_range=0.15, fov_ang=6, dist_radius=0.12, up=None):
        """nerf_utils.py:343-436: camera on the unit sphere looking at the origin.

        Returns extrinsics (b,3,4), focal (b,1,1), near (b,1,1), far (b,1,1), viewpoint (b,2).
        `up` (b,3) or (1,3): the up vector of `generate_camera_params_v1` (nerf_utils.py:466-560); None = (0,1,0).
        """
        def rng(r):
            return (r[0], r[1]) if isinstance(r, (list, tuple)) else (-r, r)

        if locations is not None:
            azim, elev = locations[:, 0:1], locations[:, 1:2]
            n = azim.shape[0]
            if locations.device.type == "cuda" and up is None and not (torch.is_tensor(fov_ang) and fov_ang.requires_grad):
                # one launch instead of ~35 (and ~60 more in its autograd): what keeps small inversion batches from being
                # launch-bound; same values as the PyTorch glue below, which stays the path for CPU tensors / custom `up`
                fov = fov_ang.to(locations.device).reshape(-1).expand(n) if torch.is_tensor(fov_ang) else fov_ang
                pose, focal, near, far = _CameraFn.apply(azim, elev, img_size, fov, dist_radius)
                return pose, focal, near, far, torch.cat([azim, elev], 1)
        elif sweep:
            (a0, a1), (e0, e1) = rng(azim_range), rng(elev_range)
            azim = (a0 + (a1 - a0) / 7 * torch.arange(8, device=device)).view(-1, 1).repeat(batch, 1)
            elev = e0 + (e1 - e0) * torch.rand(batch, 1, device=device).repeat(1, 8).view(-1, 1)
            n = batch * 8
        else:
            if uniform:
                (a0, a1), (e0, e1) = rng(azim_range), rng(elev_range)
                azim = a0 + (a1 - a0) * torch.rand(batch, 1, device=device)
                elev = e0 + (e1 - e0) * torch.rand(batch, 1, device=device)
            else:
                azim = azim_range * torch.randn(batch, 1, device=device)
                elev = elev_range * torch.randn(batch, 1, device=device)
            n = batch
        dist = torch.ones(n, 1, device=device)
        near, far = (dist - dist_radius).unsqueeze(-1), (dist + dist_radius).unsqueeze(-1)
        if torch.is_tensor(fov_ang):
            fov = fov_ang.to(device).reshape(-1, 1) * torch.ones(n, 1, device=device)
        else:
            fov = fov_ang * torch.ones(n, 1, device=device)
        focal = 0.5 * img_size / torch.tan(fov * math.pi / 180).unsqueeze(-1)
        viewpoint = torch.cat([azim, elev], 1)

        cam_dir = torch.stack([torch.cos(elev) * torch.sin(azim), torch.sin(elev), torch.cos(elev) * torch.cos(azim)],
                              dim=1).view(-1, 3)
        cam_loc = dist * cam_dir
        if up is None:
            up = torch.zeros(n, 3, device=device)                # (0, 1, 0), built without a host copy (graph-capturable)
            up[:, 1] = 1.0
        else:
            up = up.to(device=device, dtype=torch.float32).reshape(-1, 3).expand(n, 3)
        z_ax = F.normalize(cam_dir, eps=1e-5)
        x_ax = F.normalize(torch.cross(up, z_ax, dim=1), eps=1e-5)
        y_ax = F.normalize(torch.cross(z_ax, x_ax, dim=1), eps=1e-5)
        # degenerate-x fix (:428-431) applied as an unconditional select: no host sync, same values
        degenerate = (x_ax.abs() <= 5e-3).all(dim=1, keepdim=True)
        x_ax = torch.where(degenerate, F.normalize(torch.cross(y_ax, z_ax, dim=1), eps=1e-5), x_ax)
        rot = torch.stack([x_ax, y_ax, z_ax], dim=2)              # columns are the camera axes
        extrinsics = torch.cat([rot, cam_loc[:, :, None]], dim=-1)
        return extrinsics, focal, near, far, viewpoint

    @staticmethod
    def generate_camera_params_v1(img_size, device, batch=1, locations=None, sweep=False, uniform=False, azim_range=0.3,
                                  elev_range=0.15, fov_ang=6, dist_radius=0.12, up=None):
        """nerf_utils.py:466-560 (used by render_video_web_v10.py:1636): `generate_camera_params` with a caller-given up
        vector."""
        return Camera.generate_camera_params(img_size, device, batch=batch, locations=locations, sweep=sweep, uniform=uniform,
                                             azim_range=azim_range, elev_range=elev_range, fov_ang=fov_ang,
                                             dist_radius=dist_radius, up=up)

    @staticmethod
    def get_camera2world(cam2world, trans, homo=False):
        """nerf_utils.py:439-463: axis-angle rotation (..., 3) + translation (..., 3) -> extrinsics (..., 3, 4) or, with
        `homo`, (..., 4, 4).  The reference calls pytorch3d's `axis_angle_to_matrix`; the same map (Rodrigues' formula,
        series expansion near zero) is written out here so the helper has no third-party dependency."""
        assert cam2world.shape[:-1] == trans.shape[:-1]
        prefix = cam2world.shape[:-1]
        theta = cam2world.norm(dim=-1, keepdim=True)
        small = theta < 1e-4
        t2 = theta * theta
        a = torch.where(small, 1.0 - t2 / 6.0, torch.sin(theta) / theta.clamp_min(1e-30))               # sin(t)/t
        b = torch.where(small, 0.5 - t2 / 24.0, (1.0 - torch.cos(theta)) / t2.clamp_min(1e-30))        # (1-cos t)/t^2
        x, y, z = cam2world.unbind(-1)
        zero = torch.zeros_like(x)
        K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], -1).reshape(*prefix, 3, 3)
        eye = torch.eye(3, dtype=cam2world.dtype, device=cam2world.device).expand(*prefix, 3, 3)
        rot = eye + a[..., None] * K + b[..., None] * (K @ K)
        ext = torch.cat([rot, trans.reshape(*prefix, 3, 1)], dim=-1)
        if homo:
            last = torch.zeros(*prefix, 1, 4, dtype=ext.dtype, device=ext.device)
            last[..., 0, 3] = 1.0
            ext = torch.cat([ext, last], dim=-2)
        return ext

