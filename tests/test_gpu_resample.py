"""GPU parity tests of the importance-resampling EXTENSION (c3d_sample_pdf through the C ABI, NerfBranch.render_hierarchical)
against the CPU oracle (oracle/nerf_oracle.py::sample_pdf / importance_depths / render_hierarchical).

The reference has no sample_pdf or fine pass: PARITY UNPINNED -- the oracle restates the published NeRF
algorithm and is cross-checked on CPU in tests/test_resample_cpu.py.  Tolerances (resample_checks.py): new depths within 1e-4
absolute wherever the PDF carries mass, CDF-space 2e-5 elsewhere; the merged depths are a bit-exact sort of the union."""
import numpy as np
import pytest
import torch

from conftest import load_case, load_weights, rel_l2
from oracle import nerf_oracle as O
from resample_checks import check_samples, det_u, make_rays

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(_dev())


def _module(D, precision, sigmoid_beta=None):
    import cips3dpp_b200 as c3d
    m = c3d.NerfBranch(D, precision=precision)
    sd = {k: torch.from_numpy(v) for k, v in load_weights(D).items()}
    if sigmoid_beta is not None:
        sd["sigmoid_beta"] = torch.from_numpy(np.asarray(sigmoid_beta, np.float32).reshape(1))
    m.load_state_dict(sd, strict=True)
    return m.to(_dev()).eval().requires_grad_(False)


# (rays, N, K): ragged ray counts (tail chunk, non-multiple-of-4 tail), N and K above one warp, the limits 3 / 1 / 256
SHAPES = [(1, 24, 24), (7, 24, 24), (64, 24, 24), (1000, 24, 24), (4099, 24, 24), (130, 3, 1), (257, 33, 40),
          (200, 128, 64), (90, 256, 256), (77, 24, 128), (300, 100, 7)]


@pytest.mark.parametrize("R,N,K", SHAPES)
@pytest.mark.parametrize("u_mode", ["det", "sorted", "random"])
@pytest.mark.parametrize("variant", ["auto", "warp"])
def test_sample_pdf_given_weights(R, N, K, u_mode, variant, monkeypatch):
    """Both kernels: lanes = rays (the default whenever the rows of 64+ rays fit shared memory) and lanes = samples."""
    import cips3dpp_b200 as c3d
    if variant != "auto":
        c3d._abi.set_options(resample=variant)
    z, w = make_rays(R, N, seed=R + N + K, peaked=(R % 2 == 0))
    rng = np.random.default_rng(K)
    u = None
    if u_mode != "det":
        u = rng.random((R, K)).astype(np.float32)
        if u_mode == "sorted":
            u = np.sort(u, -1)
    ro = rng.normal(0, 1, (R, 3)).astype(np.float32)
    rd = rng.normal(0, 1, (R, 3)).astype(np.float32)
    out = c3d.Render.importance_depths(_t(z), K, weights=_t(w), rays_d=_t(rd), rays_o=_t(ro),
                                       u=None if u is None else _t(u), return_pts=True)
    torch.cuda.synchronize()
    z_fine, z_merged, pts = (out[k].cpu().numpy() for k in ("z_fine", "z_merged", "pts"))
    o_fine, _ = O.importance_depths(z, w, K, u)
    check_samples(z, w, z_fine, det_u(R, K) if u is None else u, o_fine)
    # the merge only moves values: bit-exact ascending union of the coarse depths and the kernel's own new depths
    np.testing.assert_array_equal(z_merged, np.sort(np.concatenate([z, z_fine], -1), -1))
    np.testing.assert_allclose(pts, ro[:, None] + rd[:, None] * z_merged[..., None], atol=2e-6, rtol=0)


@pytest.mark.parametrize("R,N,K", [(500, 24, 24), (101, 128, 64), (64, 40, 216)])
@pytest.mark.parametrize("variant", ["auto", "warp"])
def test_sample_pdf_weights_from_sdf(R, N, K, variant, monkeypatch):
    """weights == NULL: the kernel derives w = alpha * T from the sdf exactly as volume_integration does."""
    import cips3dpp_b200 as c3d
    if variant != "auto":
        c3d._abi.set_options(resample=variant)
    rng = np.random.default_rng(R)
    z, _ = make_rays(R, N, seed=R)
    # a surface crossing somewhere along the ray (sdf changes sign), or none
    cross = rng.uniform(0.8, 1.2, (R, 1)).astype(np.float32)
    sdf = ((cross - z) * rng.uniform(0.5, 3.0, (R, 1))).astype(np.float32)
    rd = rng.normal(0, 1, (R, 3)).astype(np.float32)
    beta = np.float32(0.05)
    out = c3d.Render.importance_depths(_t(z), K, sdf=_t(sdf)[..., None], rays_d=_t(rd),
                                       sigmoid_beta=torch.tensor([beta], device=_dev()))
    torch.cuda.synchronize()
    w = O.compositing_weights(sdf, z, rd, beta)
    o_fine, _ = O.importance_depths(z, w, K, None)
    z_fine, z_merged = out["z_fine"].cpu().numpy(), out["z_merged"].cpu().numpy()
    check_samples(z, w, z_fine, det_u(R, K), o_fine)
    np.testing.assert_array_equal(z_merged, np.sort(np.concatenate([z, z_fine], -1), -1))
    assert "pts" not in out


def test_sample_pdf_full_size_properties():
    """BASELINE configs[1] size (256 images x 4096 rays, N = K = 24): size-independent properties on the device."""
    import cips3dpp_b200 as c3d
    R, N, K = 256 * 4096, 24, 24
    g = torch.Generator(device=_dev()).manual_seed(0)
    z = 0.88 + 0.24 * (torch.arange(N, device=_dev())[None] + torch.rand(R, 1, device=_dev(), generator=g)) / N
    w = torch.rand(R, N, device=_dev(), generator=g) ** 8
    u = torch.rand(R, K, device=_dev(), generator=g)
    out = c3d.Render.importance_depths(z, K, weights=w, u=u)
    zf, zm = out["z_fine"], out["z_merged"]
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    assert torch.isfinite(zf).all()
    assert (zf >= mids[:, :1] - 1e-6).all() and (zf <= mids[:, -1:] + 1e-6).all()
    assert torch.equal(zm, torch.sort(torch.cat([z, zf], -1), -1).values)
    # monotone in u: sorting u sorts the samples
    us, order = torch.sort(u, -1)
    assert (torch.gather(zf, -1, order).diff(dim=-1) >= -1e-6).all()
    # idempotence / determinism
    again = c3d.Render.importance_depths(z, K, weights=w, u=u)
    assert torch.equal(again["z_merged"], zm) and torch.equal(again["z_fine"], zf)


@pytest.mark.parametrize("case,N,K", [("ffhq_d2_n24", 24, 24), ("cars_d6_n24", 24, 40)])
def test_render_hierarchical_fp32_matches_oracle(case, N, K):
    """Coarse pass -> importance depths -> fine pass on the merged depths, full 64x64 image on the GPU, compared with the
    oracle's two-pass render at the golden ray subset."""
    c = load_case(case)
    D = int(c["D"])
    m = _module(D, "fp32", c["sigmoid_beta"])
    with torch.no_grad():
        out = m.render_hierarchical(_t(c["c2w"]), _t(c["focal"]), _t(c["near"]), _t(c["far"]), _t(c["styles"]),
                                    img_size=64, N_samples=N, N_importance=K,
                                    static_viewdirs=bool(c["static_viewdirs"]), coarse_maps=True)
    params = dict(load_weights(D), sigmoid_beta=np.asarray(c["sigmoid_beta"], np.float32).reshape(1))
    idx = c["ray_idx"].astype(np.int64)
    fine, coarse = O.render_hierarchical(params, c["c2w"], c["focal"], c["near"], c["far"], c["styles"], 64, N, K,
                                         bool(c["static_viewdirs"]), None, idx)
    tidx = torch.from_numpy(idx).to(_dev())
    g = lambda k: out[k][:, tidx].cpu().numpy()
    assert out["z_vals"].shape == (c["c2w"].shape[0], 4096, N + K) and out["sdf"].shape[-2:] == (N + K, 1)
    dz = np.abs(g("z_vals") - fine[5])
    # flat-PDF bins (resample_checks.py) and the u = 1 end point (cdf[-1] rounds to either side of 1) may move by up to a
    # bin; everything else is within the 1e-4 depth tolerance
    width = float((c["far"] - c["near"]).max()) / N
    assert (dz <= 1e-4).mean() > 0.97 and dz.max() <= width + 1e-4
    assert rel_l2(g("feature_map"), fine[1]) < 1e-3
    assert rel_l2(g("rgb_map"), fine[0]) < 1e-3
    assert rel_l2(g("xyz"), fine[4]) < 1e-3
    assert np.abs(g("mask")[..., 1] - fine[3][..., 1]).max() < 1e-4
    assert rel_l2(out["coarse"]["feature_map"][:, tidx].cpu().numpy(), coarse[1]) < 1e-3


def test_render_hierarchical_bf16_close_to_fp32():
    """bf16 mode: the coarse sdf is bf16-perturbed, so the new depths move; the rendered maps stay within the bf16 bound."""
    c = load_case("ffhq_d2_n24")
    args = (_t(c["c2w"]), _t(c["focal"]), _t(c["near"]), _t(c["far"]), _t(c["styles"]))
    with torch.no_grad():
        a = _module(2, "fp32").render_hierarchical(*args, img_size=64, N_samples=24, N_importance=24)
        b = _module(2, "bf16").render_hierarchical(*args, img_size=64, N_samples=24, N_importance=24)
    assert rel_l2(b["feature_map"].cpu().numpy(), a["feature_map"].cpu().numpy()) < 2e-2
    assert rel_l2(b["rgb_map"].cpu().numpy(), a["rgb_map"].cpu().numpy()) < 2e-2


def test_render_hierarchical_gradients_flow_through_fine_pass():
    """The new depths are constants; styles and camera gradients of the fine pass match torch autograd of the reference
    formulation evaluated at the same merged depths."""
    import cips3dpp_b200 as c3d
    import torch_ref
    c = load_case("ffhq_d2_n24")
    m = _module(2, "fp32")
    S, N, K = 16, 24, 24
    pose = _t(c["c2w"]).clone().requires_grad_(True)
    styles = _t(c["styles"]).clone().requires_grad_(True)
    focal, near, far = _t(c["focal"]) * S / 64, _t(c["near"]), _t(c["far"])
    out = m.render_hierarchical(pose, focal, near, far, styles, img_size=S, N_samples=N, N_importance=K)
    g = torch.Generator(device=_dev()).manual_seed(3)
    cot = torch.randn(out["feature_map"].shape, device=_dev(), generator=g)
    (out["feature_map"] * cot).sum().backward()
    z = out["z_vals"].detach()
    assert not z.requires_grad
    params = {k: _t(v) for k, v in load_weights(2).items()}
    pose2, styles2 = pose.detach().clone().requires_grad_(True), styles.detach().clone().requires_grad_(True)
    rays_o, rays_d, viewdirs = c3d.Render.get_rays_in_world(focal, S, pose2)
    b = pose.shape[0]
    rays_o, rays_d, viewdirs = (t.reshape(b, S * S, 3) for t in (rays_o, rays_d, viewdirs))
    pts = rays_o[:, :, None] + rays_d[:, :, None] * z[..., None]
    ref = torch_ref.forward(params, pts, rays_d, viewdirs, z, near, far, styles2)
    (ref[1] * cot).sum().backward()
    assert rel_l2(out["feature_map"].detach().cpu().numpy(), ref[1].detach().cpu().numpy()) < 1e-3
    assert rel_l2(styles.grad.cpu().numpy(), styles2.grad.cpu().numpy()) < 1e-3
    assert rel_l2(pose.grad.cpu().numpy(), pose2.grad.cpu().numpy()) < 1e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", ["ffhq_d2_n24", "ffhq_d8_n24", "cars_d6_n24"])
def test_density_only_pass_matches_full_render(case, precision):
    """The coarse pass of the two-pass render stops after the sdf head: same sdf and depths as the full render, no maps;
    the two-pass result does not depend on which coarse pass produced the densities."""
    import cips3dpp_b200 as c3d
    c = load_case(case)
    m = _module(int(c["D"]), precision, c["sigmoid_beta"])
    args = (_t(c["c2w"]), _t(c["focal"]), _t(c["near"]), _t(c["far"]), _t(c["styles"]))
    kw = dict(img_size=64, N_samples=int(c["N"]), static_viewdirs=bool(c["static_viewdirs"]))
    with torch.no_grad():
        pair = m.render(*args, **kw)                              # default bf16 forward: the CTA-pair kernel
        # the density-only pass is a mode of the single-CTA kernel: bit-exact against that kernel's full render
        c3d._abi.set_options(fwd="v3")
        full = m.render(*args, **kw)
        dens = m.render(*args, density_only=True, **kw)
        a = m.render_hierarchical(*args, N_importance=24, **kw)
        b = m.render_hierarchical(*args, N_importance=24, coarse_maps=True, **kw)
    assert set(dens) == {"sdf", "z_vals"} and set(a["coarse"]) == {"sdf", "z_vals"}
    assert torch.equal(dens["z_vals"], full["z_vals"])
    assert torch.equal(dens["sdf"], full["sdf"])               # same kernel, same arithmetic up to the sdf head
    assert rel_l2(dens["sdf"].cpu().numpy(), pair["sdf"].cpu().numpy()) < (1e-5 if precision == "fp32" else 2e-2)
    assert torch.equal(a["z_vals"], b["z_vals"]) and torch.equal(a["feature_map"], b["feature_map"])
    assert "feature_map" in b["coarse"]


def test_density_only_needs_consistent_outputs():
    import cips3dpp_b200 as c3d
    lib = c3d._abi.load()
    P = c3d._abi.FwdParams()
    P.abi_version, P.mode, P.input_kind, P.batch, P.n_rays, P.n_samples, P.D, P.img_size = c3d._abi.ABI_VERSION, 1, 0, 1, 16, 24, 2, 4
    t = torch.zeros(64, device=_dev())
    P.packed = P.styles = P.near = P.far = P.cam_poses = P.focal = t.data_ptr()
    P.sdf, P.rgb_map = t.data_ptr(), t.data_ptr()               # one map without the others
    assert lib.c3d_nerf_forward(P, None) == -1 and b"density-only" in lib.c3d_last_error()


def test_render_hierarchical_perturbed_and_wide():
    """Training-style sampling (random per-ray offset in the coarse pass, random draws for the new depths) and a wide
    fine pass (24 + 104 = 128 merged samples): finite maps, ascending merged depths inside [near, far], reproducible
    under the same seed."""
    c = load_case("ffhq_d2_n24")
    m = _module(2, "bf16")
    args = (_t(c["c2w"]), _t(c["focal"]), _t(c["near"]), _t(c["far"]), _t(c["styles"]))
    outs = []
    for _ in range(2):
        torch.manual_seed(21)
        with torch.no_grad():
            outs.append(m.render_hierarchical(*args, img_size=32, N_samples=24, N_importance=104, perturb=True))
    a, b = outs
    z = a["z_vals"]
    assert z.shape == (1, 1024, 128) and torch.isfinite(a["feature_map"]).all() and torch.isfinite(a["rgb_map"]).all()
    assert (z.diff(dim=-1) >= 0).all()
    assert z.min() >= float(c["near"].min()) - 1e-6 and z.max() <= float(c["far"].max()) + 1e-6
    assert torch.equal(a["z_vals"], b["z_vals"]) and torch.equal(a["feature_map"], b["feature_map"])
    with torch.no_grad():
        torch.manual_seed(22)
        other = m.render_hierarchical(*args, img_size=32, N_samples=24, N_importance=104, perturb=True)
    assert not torch.equal(other["z_vals"], a["z_vals"])
