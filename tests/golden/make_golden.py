"""Generate golden vectors from the UNMODIFIED reference NeRF branch (run in the build container).

    python tests/golden/make_golden.py            # needs /root/reference (or $C3D_REFERENCE)

The reference's `exp/cips3d/nerf_utils.py` and `exp/cips3d/volume_renderer.py` are imported
as they are; two import-only third-party modules they pull in (`tl2.tl2_utils`, used for a
`__repr__`; `pytorch3d.transforms`, used by an unrelated camera helper) are replaced by empty
in-memory stubs.  Nothing from the reference is written into this repository except the
numeric inputs/outputs below.

Outputs (committed):
  weights_seed0.npz   D=8 renderer state dict (torch.manual_seed(0) init, values rounded to
                      fp16-representable numbers so they store exactly in 2 bytes; D=2 / D=6
                      cases reuse layers 0..D-1 + views/rgb/sigma heads of the same dict)
  case_*.npz          inputs (256-ray subsets of a 64x64 image) and reference outputs
  camera.npz          Camera.generate_camera_params + prepare_nerf_inputs checks
  camera_v1.npz       `--camera-v1`: Camera.generate_camera_params_v1 with caller-given up vectors
  volint_modes.npz    `--volint`: Render.volume_integration with with_sdf=False / force_background
  pgrads_*.npz        `--param-grads`: gradients w.r.t. every renderer parameter for the two *_grads cases, computed
                      by the reference's autograd on the committed case inputs / cotangents (fp32; at D=8 the
                      256x256 matrices are kept for layers 1 and 7 only, to bound the fixture size)
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("C3D_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    tl2 = types.ModuleType("tl2")
    tl2_utils = types.ModuleType("tl2.tl2_utils")
    tl2_utils.get_class_repr = lambda self, prefix="": f"{prefix}.{type(self).__name__}"
    tl2.tl2_utils = tl2_utils
    p3d = types.ModuleType("pytorch3d")
    p3d_tr = types.ModuleType("pytorch3d.transforms")
    p3d.transforms = p3d_tr
    sys.modules.update({"tl2": tl2, "tl2.tl2_utils": tl2_utils, "pytorch3d": p3d, "pytorch3d.transforms": p3d_tr})
    sys.path.insert(0, REF)
    import exp.cips3d.nerf_utils as nu
    import exp.cips3d.volume_renderer as vr
    return nu, vr


FFHQ = dict(fov_ang=6, dist_radius=0.12)
CARS = dict(fov_ang=15, dist_radius=0.3)


def sub_state(sd8, D):
    """D-layer state dict built from the first D point layers of the D=8 one."""
    out = {}
    for k, v in sd8.items():
        if k.startswith("network.pts_linears."):
            if int(k.split(".")[2]) < D:
                out[k] = v
        else:
            out[k] = v
    return out


def main():
    nu, vr = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    R8 = vr.VolumeFeatureRenderer(N_layers_renderer=8, input_dim=3, hidden_dim=256, style_dim=256,
                                  view_dim=3, with_sdf=True, output_features=True)
    with torch.no_grad():
        for p in R8.parameters():
            p.copy_(p.half().float())
    sd8 = {k: v.detach().clone() for k, v in R8.state_dict().items()}
    np.savez(os.path.join(HERE, "weights_seed0.npz"),
             **{k: v.numpy().astype(np.float16) for k, v in sd8.items()})

    def renderer(D):
        R = vr.VolumeFeatureRenderer(N_layers_renderer=D, input_dim=3, hidden_dim=256, style_dim=256,
                                     view_dim=3, with_sdf=True, output_features=True)
        R.load_state_dict(sub_state(sd8, D), strict=True)
        return R.eval()

    ray_idx = torch.arange(0, 4096, 16) + (torch.arange(256) // 4) % 16      # 256 rays spread over the image

    def case(name, D, N, cam, locs, static_viewdirs=False, perturb=False, wplus=False, seed=1,
             sigmoid_beta=None, grads=False):
        g = torch.Generator().manual_seed(seed)
        R = renderer(D)
        if sigmoid_beta is not None:
            with torch.no_grad():
                R.sigmoid_beta.fill_(sigmoid_beta)
        locs_t = torch.tensor(locs, dtype=torch.float32)
        b = locs_t.shape[0]
        c2w, focal, near, far, _ = nu.Camera.generate_camera_params(
            img_size=64, device="cpu", locations=locs_t, **cam)
        torch.manual_seed(seed + 100)         # perturb draw (torch.rand inside get_z_vals)
        pts, rays_d, viewdirs, z_vals = nu.Render.prepare_nerf_inputs(
            focal=focal, img_size=64, cam_poses=c2w, near=near, far=far, N_samples=N,
            perturb=perturb, static_viewdirs=static_viewdirs)
        pts = pts.reshape(b, 4096, N, 3)[:, ray_idx].contiguous()
        rays_d = rays_d.reshape(b, 4096, 3)[:, ray_idx].contiguous()
        viewdirs = viewdirs.reshape(b, 4096, 3)[:, ray_idx].contiguous()
        z_vals = z_vals.reshape(b, 4096, N)[:, ray_idx].contiguous()
        if wplus:
            styles = 0.6 * torch.randn(b, D + 1, 256, generator=g)
        else:
            styles = (0.6 * torch.randn(b, 1, 256, generator=g)).repeat(1, D + 1, 1)
        out = {}
        if grads:
            styles.requires_grad_(True)
            pts.requires_grad_(True)
            rays_d.requires_grad_(True)
            viewdirs.requires_grad_(True)
            rgb_map, feat, sdf, mask, xyz, _ = R(pts=pts, rays_d=rays_d, viewdirs=viewdirs, z_vals=z_vals,
                                                 near=near, far=far, styles=styles)
            cot = {k: torch.randn(v.shape, generator=g) for k, v in
                   dict(rgb_map=rgb_map, feature_map=feat, mask=mask, xyz=xyz).items()}
            loss = (rgb_map * cot["rgb_map"]).sum() + (feat * cot["feature_map"]).sum() * 0.05 \
                + (mask * cot["mask"]).sum() + (xyz * cot["xyz"]).sum()
            gs = torch.autograd.grad(loss, [styles, pts, rays_d, viewdirs])
            out.update({"cot_" + k: v.numpy() for k, v in cot.items()})
            out.update(g_styles=gs[0].numpy(), g_pts=gs[1].numpy(), g_rays_d=gs[2].numpy(),
                       g_viewdirs=gs[3].numpy(), loss=np.float32(loss.item()))
        else:
            with torch.no_grad():
                rgb_map, feat, sdf, mask, xyz, _ = R(pts=pts, rays_d=rays_d, viewdirs=viewdirs, z_vals=z_vals,
                                                     near=near, far=far, styles=styles)
        out.update(
            D=np.int32(D), N=np.int32(N), static_viewdirs=np.int32(static_viewdirs),
            sigmoid_beta=R.sigmoid_beta.detach().numpy(), ray_idx=ray_idx.numpy().astype(np.int32),
            locations=locs_t.numpy(), c2w=c2w.numpy(), focal=focal.numpy(), near=near.numpy(), far=far.numpy(),
            pts=pts.detach().numpy(), rays_d=rays_d.detach().numpy(), viewdirs=viewdirs.detach().numpy(),
            z_vals=z_vals.numpy(), styles=styles.detach().numpy(),
            rgb_map=rgb_map.detach().numpy(), feature_map=feat.detach().numpy(), sdf=sdf.detach().numpy(),
            mask=mask.detach().numpy(), xyz=xyz.detach().numpy())
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
        print(name, {k: float(np.abs(out[k]).mean()) for k in ("rgb_map", "feature_map", "sdf", "mask", "xyz")})

    case("ffhq_d8_n24", 8, 24, FFHQ, [[0.25, -0.1]])
    case("ffhq_d2_n24", 2, 24, FFHQ, [[0.25, -0.1]])
    case("ffhq_d2_n128_static", 2, 128, FFHQ, [[0.25, -0.1]], static_viewdirs=True)
    case("cars_d6_n24", 6, 24, CARS, [[0.25, -0.1]])
    case("ffhq_d8_n24_b2_wplus_perturb", 8, 24, FFHQ, [[-0.3, 0.12], [0.0, 0.0]], perturb=True, wplus=True)
    case("cars_d6_n36_b2_beta", 6, 36, CARS, [[3.0, 0.16], [-1.5, 0.0]], wplus=True, sigmoid_beta=0.03)
    case("ffhq_d2_n24_grads", 2, 24, FFHQ, [[0.2, 0.05], [-0.2, -0.05]], wplus=True, grads=True)
    case("ffhq_d8_n24_grads_static", 8, 24, FFHQ, [[0.1, 0.1]], static_viewdirs=True, grads=True)

    # survey anchors: un-rounded seed-0 weights, full image, styles = randn after camera
    torch.manual_seed(0)
    Rr = vr.VolumeFeatureRenderer(N_layers_renderer=8, input_dim=3, hidden_dim=256, style_dim=256,
                                  view_dim=3, with_sdf=True, output_features=True)
    c2w, focal, near, far, _ = nu.Camera.generate_camera_params(
        img_size=64, device="cpu", locations=torch.zeros(1, 2), **FFHQ)
    pts, rays_d, viewdirs, z_vals = nu.Render.prepare_nerf_inputs(
        focal=focal, img_size=64, cam_poses=c2w, near=near, far=far, N_samples=24, perturb=False,
        static_viewdirs=False)
    styles = torch.randn(1, 9, 256)
    with torch.no_grad():
        o = Rr(pts=pts.reshape(1, 4096, 24, 3), rays_d=rays_d.reshape(1, 4096, 3),
               viewdirs=viewdirs.reshape(1, 4096, 3), z_vals=z_vals.reshape(1, 4096, 24),
               near=near, far=far, styles=styles)
    print("survey anchor mean-abs:", [float(t.abs().mean()) for t in o[:5]])

    # camera + ray generation goldens
    cam = {}
    for tag, cfg, locs in (("ffhq", FFHQ, [[0.0, 0.0], [0.25, -0.1], [-0.3, 0.15]]),
                           ("cars", CARS, [[3.14, 0.0], [-1.2, 0.1674], [0.0, 1.5707963]])):
        locs_t = torch.tensor(locs, dtype=torch.float32)
        c2w, focal, near, far, vp = nu.Camera.generate_camera_params(
            img_size=64, device="cpu", locations=locs_t, **cfg)
        for sv in (False, True):
            pts, rays_d, viewdirs, z_vals = nu.Render.prepare_nerf_inputs(
                focal=focal, img_size=64, cam_poses=c2w, near=near, far=far, N_samples=24, perturb=False,
                static_viewdirs=sv)
            cam.update({f"{tag}_sv{int(sv)}_pts": pts[:, ::8, ::8].numpy(),
                        f"{tag}_sv{int(sv)}_rays_d": rays_d[:, ::4, ::4].numpy(),
                        f"{tag}_sv{int(sv)}_viewdirs": viewdirs[:, ::4, ::4].numpy(),
                        f"{tag}_sv{int(sv)}_z_vals": z_vals[:, ::8, ::8].numpy()})
        cam.update({f"{tag}_locations": locs_t.numpy(), f"{tag}_c2w": c2w.numpy(), f"{tag}_focal": focal.numpy(),
                    f"{tag}_near": near.numpy(), f"{tag}_far": far.numpy()})
    np.savez_compressed(os.path.join(HERE, "camera.npz"), **cam)
    camera_v1(nu)


def volint_modes(nu=None):
    """Render.volume_integration in the branches the v10 configs leave unused (with_sdf=False, force_background) ->
    volint_modes.npz (random inputs, noise-free)."""
    if nu is None:
        nu, _ = import_reference()
    g = torch.Generator().manual_seed(11)
    R, N, C = 40, 24, 16
    rgb = torch.randn(R, N, 3, generator=g)
    raw = torch.randn(R, N, 1, generator=g) * 3.0
    feat = torch.randn(R, N, C, generator=g)
    z = torch.sort(0.7 + 0.6 * torch.rand(R, N, generator=g), dim=-1).values
    rd = torch.randn(R, 3, generator=g)
    pts = torch.randn(R, N, 3, generator=g)
    out = {"rgb": rgb, "raw": raw, "features": feat, "z_vals": z, "rays_d": rd, "pts": pts}
    for tag, kw in (("raw", dict(with_sdf=False)), ("raw_bg", dict(with_sdf=False, force_background=True)),
                    ("sdf_bg", dict(with_sdf=True, sigmoid_beta=torch.tensor([0.1]), force_background=True))):
        sdf_in = raw if not kw["with_sdf"] else raw * 0.05
        r = nu.Render.volume_integration(rgb=rgb, sdf=sdf_in, features=feat, z_vals=z, rays_d=rd, pts=pts, **kw)
        out.update({f"{tag}_rgb_map": r[0], f"{tag}_feature_map": r[1], f"{tag}_xyz": r[2], f"{tag}_mask": r[3]})
    np.savez_compressed(os.path.join(HERE, "volint_modes.npz"), **{k: v.numpy() for k, v in out.items()})


def camera_v1(nu=None):
    """Camera.generate_camera_params_v1 (caller-given up vector, nerf_utils.py:466-560) -> camera_v1.npz."""
    if nu is None:
        nu, _ = import_reference()
    locs = torch.tensor([[0.0, 0.0], [0.4, -0.2], [-2.5, 0.3], [3.0, 0.05]], dtype=torch.float32)
    ups = torch.tensor([[0.0, 1.0, 0.0], [0.1, 0.9, -0.2], [0.0, 0.0, 1.0], [-0.3, 1.0, 0.3]], dtype=torch.float32)
    c2w, focal, near, far, vp = nu.Camera.generate_camera_params_v1(img_size=64, device="cpu", locations=locs, up=ups, **CARS)
    np.savez_compressed(os.path.join(HERE, "camera_v1.npz"), locations=locs.numpy(), up=ups.numpy(), c2w=c2w.numpy(),
                        focal=focal.numpy(), near=near.numpy(), far=far.numpy(), viewpoint=vp.numpy())


def param_grads():
    """Reference autograd w.r.t. the renderer parameters on the committed *_grads cases (same loss as `case`)."""
    nu, vr = import_reference()
    torch.set_num_threads(os.cpu_count())
    w = np.load(os.path.join(HERE, "weights_seed0.npz"))
    sd8 = {k: torch.from_numpy(w[k].astype(np.float32)) for k in w.files}
    for name in ("ffhq_d2_n24_grads", "ffhq_d8_n24_grads_static"):
        c = np.load(os.path.join(HERE, f"case_{name}.npz"))
        D = int(c["D"])
        R = vr.VolumeFeatureRenderer(N_layers_renderer=D, input_dim=3, hidden_dim=256, style_dim=256,
                                     view_dim=3, with_sdf=True, output_features=True)
        R.load_state_dict(sub_state(sd8, D), strict=True)
        t = lambda k: torch.from_numpy(c[k])
        rgb_map, feat, sdf, mask, xyz, _ = R(pts=t("pts"), rays_d=t("rays_d"), viewdirs=t("viewdirs"),
                                             z_vals=t("z_vals"), near=t("near"), far=t("far"), styles=t("styles"))
        loss = (rgb_map * t("cot_rgb_map")).sum() + (feat * t("cot_feature_map")).sum() * 0.05 \
            + (mask * t("cot_mask")).sum() + (xyz * t("cot_xyz")).sum()
        assert abs(loss.item() - float(c["loss"])) <= 1e-4 * abs(float(c["loss"])), (loss.item(), float(c["loss"]))
        names = [k for k, _ in R.named_parameters()]
        gs = torch.autograd.grad(loss, [p for _, p in R.named_parameters()])
        keep = lambda k, g: g.numel() < 65536 or D <= 2 or "views" in k or k.split(".")[2] in ("1", str(D - 1))   # size
        # the eikonal term of the same inputs (nerf_utils.py:220-228), first order
        pts_e = t("pts").clone()
        eik = R(pts=pts_e, rays_d=t("rays_d"), viewdirs=t("viewdirs"), z_vals=t("z_vals"), near=t("near"), far=t("far"),
                styles=t("styles"), return_eikonal=True)[5]
        # ... and the second-order gradients of an eikonal loss (train_v10-style regulariser) w.r.t. styles and parameters
        st_e = t("styles").clone().requires_grad_(True)
        eik2 = R(pts=t("pts").clone(), rays_d=t("rays_d"), viewdirs=t("viewdirs"), z_vals=t("z_vals"), near=t("near"),
                 far=t("far"), styles=st_e, return_eikonal=True)[5]
        loss_e = ((eik2.norm(dim=-1) - 1) ** 2).mean()
        ps = [p for _, p in R.named_parameters()]
        ge = torch.autograd.grad(loss_e, [st_e] + ps, allow_unused=True)
        eik_grads = {"eik_g_styles": ge[0].numpy(), "eik_loss": np.float32(loss_e.item())}
        for k, g_ in zip(names, ge[1:]):
            if g_ is not None and keep(k, g_):
                eik_grads["eik_g_" + k] = g_.numpy()
        np.savez_compressed(os.path.join(HERE, f"pgrads_{name}.npz"), eikonal_term=eik.detach().numpy(), **eik_grads,
                            **{k: g.numpy() for k, g in zip(names, gs) if keep(k, g)})
        print(name, {k: float(g.abs().mean()) for k, g in zip(names, gs) if "weight" not in k or "pts_linears.1." in k})


def mlp_init_case():
    """VolumeFeatureRenderer.mlp_init_pass (volume_renderer.py:569-634) on a small image with the stratified-sampling
    draw pinned (torch.rand is patched to return the stored tensor)."""
    nu, vr = import_reference()
    w = np.load(os.path.join(HERE, "weights_seed0.npz"))
    sd8 = {k: torch.from_numpy(w[k].astype(np.float32)) for k in w.files}
    D, S, N, b = 2, 8, 12, 2
    R = vr.VolumeFeatureRenderer(N_layers_renderer=D, input_dim=3, hidden_dim=256, style_dim=256, view_dim=3,
                                 with_sdf=True, output_features=True)
    R.load_state_dict(sub_state(sd8, D), strict=True)
    g = torch.Generator().manual_seed(21)
    locs = torch.tensor([[0.2, -0.1], [-0.25, 0.1]])
    c2w, focal, near, far, _ = nu.Camera.generate_camera_params(img_size=S, device="cpu", locations=locs, **FFHQ)
    styles = 0.6 * torch.randn(b, D + 1, 256, generator=g)
    t_rand = torch.rand(b, S, S, N, generator=g)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: t_rand.clone()
    try:
        sdf, target = R.mlp_init_pass(cam_poses=c2w, focals=focal, img_size=S, near=near, far=far, styles=styles,
                                      nerf_cfg=dict(N_samples=N))
    finally:
        torch.rand = real_rand
    loss = ((sdf - target) ** 2).mean()
    gs = torch.autograd.grad(loss, [R.network.pts_linears[1].weight, R.network.sigma_linear.weight, R.network.pts_linears[0].gamma.bias])
    np.savez_compressed(os.path.join(HERE, "mlp_init_pass.npz"), c2w=c2w.numpy(), focal=focal.numpy(), near=near.numpy(),
                        far=far.numpy(), styles=styles.numpy(), t_rand=t_rand.numpy(), sdf=sdf.detach().numpy(),
                        target=target.numpy(), loss=np.float32(loss.item()), g_w1=gs[0].numpy(), g_wsigma=gs[1].numpy(),
                        g_gamma0_bias=gs[2].numpy(), D=np.int32(D), S=np.int32(S), N=np.int32(N))
    print("mlp_init_pass", sdf.shape, float(loss))


if __name__ == "__main__":
    if "--param-grads" in sys.argv:
        param_grads()
        mlp_init_case()
    elif "--camera-v1" in sys.argv:
        camera_v1()
    elif "--volint" in sys.argv:
        volint_modes()
    else:
        main()
