"""Why the tensor-core mode feeds IEEE half (not bfloat16) operands to every product that reads the hidden-layer activations
(csrc/c3d_common.cuh, "16-bit operand formats"): a CPU emulation of the kernels' operand rounding -- activations and hidden-layer
weights rounded to the 16-bit format with a straight-through derivative, fp32 accumulation, FiLM / sine / compositing in fp32
(tests/torch_ref.py otherwise) -- against the reference's own outputs and autograd gradients (tests/golden, D = 8).
bfloat16 operands reproduce what round 1 measured on the GPU (feature_map 1.06e-2, style gradients 3.3e-2: beyond SURVEY 8(d)'s
2e-2); fp16 operands are an order of magnitude inside it.  Test infrastructure only: nothing here is on the product path."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_case, load_weights, rel_l2


def _ste(x, dt):
    return x + (x.to(dt).float() - x).detach()


def _forward(params, pts, rays_d, viewdirs, z, near, far, styles, dt):
    b, D = pts.shape[0], styles.shape[1] - 1
    h = pts * 2 / (far - near).reshape(b, 1, 1, 1)

    def film(x, pre, s, first=False):
        w = params[pre + "weight"]
        if first:
            out = F.linear(x, w, params[pre + "bias"])
        elif x.shape[-1] > 256:
            out = F.linear(_ste(x[..., :256], dt), _ste(w[:, :256], dt)) + F.linear(x[..., 256:], w[:, 256:]) + params[pre + "bias"]
        else:
            out = F.linear(_ste(x, dt), _ste(w, dt), params[pre + "bias"])
        gamma = 15 * F.linear(s, params[pre + "gamma.weight"], params[pre + "gamma.bias"]) + 30
        beta = 0.25 * F.linear(s, params[pre + "beta.weight"], params[pre + "beta.bias"])
        return torch.sin(gamma.view(b, 1, 1, -1) * out + beta.view(b, 1, 1, -1))

    for i in range(D):
        h = film(h, f"network.pts_linears.{i}.", styles[:, i], first=(i == 0))
    sdf = F.linear(h, params["network.sigma_linear.weight"], params["network.sigma_linear.bias"])
    feat = film(torch.cat([h, viewdirs[:, :, None, :].expand(-1, -1, h.shape[2], -1)], -1), "network.views_linears.", styles[:, -1])
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
    return rgb_map, fmap, mask, xyz


@pytest.mark.parametrize("fmt", ["bfloat16", "float16"])
def test_operand_rounding_of_the_tensor_core_mode(fmt):
    torch.set_num_threads(min(8, torch.get_num_threads()))
    c = load_case("ffhq_d8_n24_grads_static")
    params = {k: torch.from_numpy(v) for k, v in load_weights(int(c["D"])).items()}
    t = lambda k: torch.from_numpy(np.ascontiguousarray(c[k]))
    styles, pts, rays_d, viewdirs = (t(k).clone().requires_grad_(True) for k in ("styles", "pts", "rays_d", "viewdirs"))
    rgb_map, fmap, mask, xyz = _forward(params, pts, rays_d, viewdirs, t("z_vals"), t("near"), t("far"), styles, getattr(torch, fmt))
    loss = (rgb_map * t("cot_rgb_map")).sum() + 0.05 * (fmap * t("cot_feature_map")).sum() + (mask * t("cot_mask")).sum() \
        + (xyz * t("cot_xyz")).sum()
    loss.backward()
    e_feat = rel_l2(fmap.detach().numpy(), c["feature_map"])
    e_styles, e_pts = rel_l2(styles.grad.numpy(), c["g_styles"]), rel_l2(pts.grad.numpy(), c["g_pts"])
    print(fmt, dict(feature_map=e_feat, g_styles=e_styles, g_pts=e_pts))
    if fmt == "bfloat16":     # what the round-1 kernels measured on B200: 1.06e-2 / 3.26e-2 / 3.01e-2
        assert 5e-3 < e_feat < 2e-2 and 2e-2 < e_styles < 5e-2 and 2e-2 < e_pts < 5e-2
    else:                     # the shipped operand format
        assert e_feat < 2e-3 and e_styles < 6e-3 and e_pts < 6e-3
