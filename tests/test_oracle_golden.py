"""Pin the numpy oracle against vectors produced by the reference's own modules
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import CASES, GRAD_CASES, load_case, load_weights, rel_l2
from oracle import nerf_oracle as O


@pytest.mark.parametrize("name", CASES + GRAD_CASES)
def test_renderer_forward_matches_reference(name):
    c = load_case(name)
    params = load_weights(int(c["D"]))
    params["sigmoid_beta"] = c["sigmoid_beta"].astype(np.float32)
    rgb_map, feat, sdf, mask, xyz = O.renderer_forward(
        params, c["pts"], c["rays_d"], c["viewdirs"], c["z_vals"], c["near"], c["far"], c["styles"])
    # fp32 restatement: op-order noise only (~1e-7)
    assert rel_l2(feat, c["feature_map"]) < 2e-5
    assert rel_l2(rgb_map, c["rgb_map"]) < 2e-5
    assert rel_l2(sdf, c["sdf"]) < 2e-5
    assert rel_l2(xyz, c["xyz"]) < 2e-5
    assert np.abs(mask - c["mask"]).max() < 1e-5          # [:,0] background prob, [:,1] depth


@pytest.mark.parametrize("name", ["ffhq_d8_n24", "cars_d6_n24", "ffhq_d2_n128_static"])
def test_render_from_pose_matches_reference(name):
    """camera -> rays -> points -> renderer, all restated, from (azim, elev) only."""
    c = load_case(name)
    params = load_weights(int(c["D"]))
    cam = dict(fov_ang=15, dist_radius=0.3) if name.startswith("cars") else dict(fov_ang=6, dist_radius=0.12)
    c2w, focal, near, far, _ = O.generate_camera_params(c["locations"], 64, **cam)
    np.testing.assert_allclose(c2w, c["c2w"], atol=1e-6)
    np.testing.assert_allclose(focal, c["focal"], rtol=1e-6)
    rgb_map, feat, sdf, mask, xyz, z_vals = O.render(
        params, c2w, focal, near, far, c["styles"], 64, int(c["N"]),
        static_viewdirs=bool(c["static_viewdirs"]), ray_idx=c["ray_idx"])
    assert np.abs(z_vals - c["z_vals"]).max() < 1e-6
    assert rel_l2(feat, c["feature_map"]) < 5e-5
    assert rel_l2(rgb_map, c["rgb_map"]) < 5e-5
    assert np.abs(mask[..., 1] - c["mask"][..., 1]).max() < 1e-5


def test_camera_and_rays_match_reference(golden_dir):
    z = np.load(f"{golden_dir}/camera.npz")
    for tag, cam in (("ffhq", dict(fov_ang=6, dist_radius=0.12)), ("cars", dict(fov_ang=15, dist_radius=0.3))):
        c2w, focal, near, far, _ = O.generate_camera_params(z[f"{tag}_locations"], 64, **cam)
        np.testing.assert_allclose(c2w, z[f"{tag}_c2w"], atol=2e-6)
        np.testing.assert_allclose(focal, z[f"{tag}_focal"], rtol=1e-6)
        np.testing.assert_allclose(near, z[f"{tag}_near"], rtol=0, atol=0)
        np.testing.assert_allclose(far, z[f"{tag}_far"], rtol=0, atol=0)
        for sv in (0, 1):
            pts, rays_d, viewdirs, z_vals = O.prepare_nerf_inputs(
                z[f"{tag}_focal"], 64, z[f"{tag}_c2w"], near, far, 24, None, bool(sv))
            np.testing.assert_allclose(pts[:, ::8, ::8], z[f"{tag}_sv{sv}_pts"], atol=1e-6)
            np.testing.assert_allclose(rays_d[:, ::4, ::4], z[f"{tag}_sv{sv}_rays_d"], atol=1e-6)
            np.testing.assert_allclose(viewdirs[:, ::4, ::4], z[f"{tag}_sv{sv}_viewdirs"], atol=1e-6)
            np.testing.assert_allclose(z_vals[:, ::8, ::8], z[f"{tag}_sv{sv}_z_vals"], atol=1e-6)


def test_survey_known_answers():
    """SURVEY.md 8(c) sanity anchors that do not depend on weights."""
    c2w, focal, near, far, _ = O.generate_camera_params(np.zeros((1, 2), np.float32), 64, 6, 0.12)
    assert abs(float(focal[0, 0, 0]) - 304.46) < 0.01
    assert abs(float(near[0, 0, 0]) - 0.88) < 1e-6 and abs(float(far[0, 0, 0]) - 1.12) < 1e-6
    pts, rays_d, viewdirs, z_vals = O.prepare_nerf_inputs(focal, 64, c2w, near, far, 24)
    np.testing.assert_allclose(z_vals[0, 0, 0, :4], [0.88, 0.89, 0.90, 0.91], atol=1e-6)
    np.testing.assert_allclose(pts[0, 0, 0, 0], [-0.0910, 0.0910, 0.1200], atol=1e-4)
    _, focal_c, _, _, _ = O.generate_camera_params(np.array([[0.25, -0.1]], np.float32), 64, 15, 0.3)
    assert abs(float(focal_c[0, 0, 0]) - 119.4256) < 1e-3
    assert O.flops_per_point(8) == 1053696 and O.flops_per_point(6) == 791552 and O.flops_per_point(2) == 267264


def test_perturbed_z_vals():
    c = load_case("ffhq_d8_n24_b2_wplus_perturb")
    near, far = c["near"], c["far"]
    base = O.get_z_vals(near, far, 2, 1, 256, 24)
    step = (far - near).reshape(2, 1, 1, 1) / 24
    t = (c["z_vals"].reshape(2, 1, 256, 24) - base) / step
    assert t.min() >= -1e-4 and t.max() <= 1 + 1e-4
    assert np.abs(t - t[..., :1]).max() < 1e-3          # one shared offset per ray (nerf_utils.py:110)
    z2 = O.get_z_vals(near, far, 2, 1, 256, 24, t_rand=t[..., :1])
    assert np.abs(z2 - c["z_vals"].reshape(2, 1, 256, 24)).max() < 1e-6


def test_volume_integration_unused_branches_match_reference():
    """with_sdf=False (softplus density) and force_background (nerf_utils.py:288-296, 309-310) vs the reference's own output
    (tests/golden/volint_modes.npz, written by make_golden.py --volint)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "volint_modes.npz"))
    for tag, with_sdf, bg, scale in (("raw", False, False, 1.0), ("raw_bg", False, True, 1.0), ("sdf_bg", True, True, 0.05)):
        out = O.volume_integration(g["rgb"], g["raw"] * np.float32(scale), g["features"], g["z_vals"], g["rays_d"], g["pts"],
                                   np.float32(0.1), with_sdf=with_sdf, force_background=bg)
        for a, k in zip(out[:4], ("rgb_map", "feature_map", "xyz", "mask")):
            np.testing.assert_allclose(a, g[f"{tag}_{k}"], atol=3e-6, rtol=2e-5)
