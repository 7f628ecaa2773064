"""TEST INFRASTRUCTURE: differentiable fp32 torch restatement of the reference NeRF branch
(exp/cips3d/nerf_utils.py:17-218,230-338; exp/cips3d/volume_renderer.py:15-160,192-283), used only to drive the same
optimisation loop through autograd as a reference for loss-curve parity.  Checked against the golden vectors in
tests/test_gpu_backward.py::test_torch_ref_matches_golden."""
import torch
import torch.nn.functional as F


def render_thumb(params, pose, focal, near, far, styles, S, N, static_viewdirs):
    dev = pose.device
    b = pose.shape[0]
    D = styles.shape[1] - 1
    lin = torch.linspace(0.5, S - 0.5, S, device=dev)
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    f = focal.reshape(b, 1, 1)
    d_cam = torch.stack([(xx[None] - S / 2) / f, -(yy[None] - S / 2) / f, -torch.ones(b, S, S, device=dev)], -1)
    rays_d = (d_cam[..., None, :] * pose[:, None, None, :3, :3]).sum(-1).reshape(b, S * S, 3)
    viewdirs = F.normalize(d_cam.reshape(b, S * S, 3) if static_viewdirs else rays_d, dim=-1)
    t = torch.linspace(0.0, 1.0 - 1.0 / N, N, device=dev).view(1, 1, N)
    z = near.reshape(b, 1, 1) * (1 - t) + far.reshape(b, 1, 1) * t
    z = z.expand(b, S * S, N)
    pts = pose[:, None, None, :3, 3] + rays_d[:, :, None, :] * z[..., None]
    return forward(params, pts, rays_d, viewdirs, z, near, far, styles)


def forward(params, pts, rays_d, viewdirs, z, near, far, styles):
    b = pts.shape[0]
    D = styles.shape[1] - 1
    h = pts * 2 / (far - near).reshape(b, 1, 1, 1)

    def film(x, pre, s):
        out = F.linear(x, params[pre + "weight"], params[pre + "bias"])
        gamma = 15 * F.linear(s, params[pre + "gamma.weight"], params[pre + "gamma.bias"]) + 30
        beta = 0.25 * F.linear(s, params[pre + "beta.weight"], params[pre + "beta.bias"])
        return torch.sin(gamma.view(b, 1, 1, -1) * out + beta.view(b, 1, 1, -1))

    for i in range(D):
        h = film(h, f"network.pts_linears.{i}.", styles[:, i])
    sdf = F.linear(h, params["network.sigma_linear.weight"], params["network.sigma_linear.bias"])
    feat = film(torch.cat([h, viewdirs[:, :, None, :].expand(-1, -1, h.shape[2], -1)], -1), "network.views_linears.",
                styles[:, -1])
    rgb = F.linear(feat, params["network.rgb_linear.weight"], params["network.rgb_linear.bias"])
    beta_s = params["sigmoid_beta"]
    dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], -1) * rays_d.norm(dim=-1, keepdim=True)
    sigma = torch.sigmoid(-sdf / beta_s) / beta_s
    alpha = 1 - torch.exp(-sigma * dists[..., None])
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1, :]), 1 - alpha + 1e-10], -2), -2)[..., :-1, :]
    w = alpha * T
    rgb_map = -1 + 2 * (w * torch.sigmoid(rgb)).sum(-2)
    fmap = (w * feat).sum(-2)
    xyz = (w * pts).sum(-2)
    mask = torch.cat([w[..., -1, :], -xyz.norm(dim=-1, keepdim=True)], -1)
    return rgb_map, fmap, sdf, mask, xyz
