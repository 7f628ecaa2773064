"""CPU-side checks of the boundary: the library builds for sm_100a, loads, and exports every symbol that
include/c3d_abi.h declares; argument validation that needs no GPU; host-side mirrors fail loudly on CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import cips3dpp_b200 as c3d
    c3d._abi.build_library()
    return c3d._abi.load()


def test_header_symbols_are_exported(lib):
    import cips3dpp_b200 as c3d
    hdr = open(os.path.join(ROOT, "include", "c3d_abi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(c3d_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 12
    raw = ctypes.CDLL(c3d._abi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in c3d_abi.h but not exported"
    assert declared == set(c3d._abi.EXPORTS), declared ^ set(c3d._abi.EXPORTS)


def test_abi_version_and_sizes(lib):
    import cips3dpp_b200 as c3d
    assert lib.c3d_abi_version() == c3d._abi.ABI_VERSION
    assert lib.c3d_packed_bytes(0) == 0 and lib.c3d_packed_bytes(17) == 0
    assert lib.c3d_packed_bytes(8) > 8 * 256 * 256 * 2
    assert lib.c3d_packed_bytes(8) % 1024 == 0


def test_struct_sizes_match_header(lib):
    """ctypes mirrors must have the C layout (all members are int32 / pointers / size_t, natural alignment)."""
    import cips3dpp_b200 as c3d
    a = c3d._abi
    assert ctypes.sizeof(a.RawParams) == 8 + 6 * 16 * 8 + 11 * 8
    assert ctypes.sizeof(a.FwdParams) == 10 * 4 + 18 * 8 + 8 + 8
    assert ctypes.sizeof(a.GatherOut) == 8 + 4 * 16 * 8
    assert ctypes.sizeof(a.BwdParams) == ctypes.sizeof(a.FwdParams) + 12 * 8 + 8
    assert ctypes.sizeof(a.ParamGrads) == ctypes.sizeof(a.RawParams) - 8
    assert ctypes.sizeof(a.RaygenParams) == 4 * 4 + 9 * 8
    assert ctypes.sizeof(a.CompositeParams) == 8 + 4 + 4 + 4 + 4 + 22 * 8
    assert ctypes.sizeof(a.AdamParams) == 4 + 8 * 4 + 4 + 8 * 8 * 5 + 2 * 8 + 8 + 4 * 4 + 8   # int32[9] padded to 40, then 8-byte members


def test_argument_validation_without_gpu(lib):
    import cips3dpp_b200 as c3d
    P = c3d._abi.FwdParams()
    assert lib.c3d_nerf_forward(P, None) == -1
    assert b"abi_version" in lib.c3d_last_error()
    P.abi_version = c3d._abi.ABI_VERSION
    P.mode = 7
    assert lib.c3d_nerf_forward(P, None) == -1 and b"mode" in lib.c3d_last_error()
    P.mode, P.batch, P.n_rays, P.n_samples, P.D = 1, 1, 16, 300, 2
    assert lib.c3d_nerf_forward(P, None) == -1 and b"n_samples" in lib.c3d_last_error()
    C = c3d._abi.CompositeParams()
    C.n_rays, C.n_samples = 4, 1000
    assert lib.c3d_composite_forward(C, None) == -1 and b"n_samples" in lib.c3d_last_error()
    assert lib.c3d_umma_selftest(None, None, None, 256, 256, 0, None) == -1


def test_module_is_dropin_shaped_and_has_no_cpu_path():
    import copy
    import cips3dpp_b200 as c3d
    m = c3d.NerfBranch(N_layers_renderer=2, input_dim=3, hidden_dim=256, style_dim=256, view_dim=3,
                       with_sdf=True, output_features=True)
    keys = set(m.state_dict())
    want = {"sigmoid_beta", "network.rgb_linear.weight", "network.rgb_linear.bias", "network.sigma_linear.weight",
            "network.sigma_linear.bias"}
    for pre in ("network.pts_linears.0.", "network.pts_linears.1.", "network.views_linears."):
        want |= {pre + s for s in ("weight", "bias", "gamma.weight", "gamma.bias", "beta.weight", "beta.bias")}
    assert keys == want
    assert m.network.pts_linears[0].weight.shape == (256, 3)
    assert m.network.views_linears.weight.shape == (256, 259)
    assert m.N_layers_renderer == 2 and m.sigmoid_beta.shape == (1,)
    m2 = copy.deepcopy(m)
    m2.load_state_dict(m.state_dict(), strict=True)
    x = torch.zeros(1, 4, 24, 3)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(pts=x, rays_d=x[:, :, 0], viewdirs=x[:, :, 0], z_vals=x[..., 0], near=torch.zeros(1, 1, 1),
          far=torch.ones(1, 1, 1), styles=torch.zeros(1, 3, 256))
    with pytest.raises(RuntimeError, match="no CPU"):
        c3d.Render.prepare_nerf_inputs(torch.ones(1, 1, 1), 8, torch.zeros(1, 3, 4), torch.zeros(1, 1, 1),
                                       torch.ones(1, 1, 1), 24, False)


def test_init_distributions_follow_reference():
    """volume_renderer.py:19-27,56-63: value ranges of the random init."""
    import cips3dpp_b200 as c3d
    torch.manual_seed(0)
    m = c3d.NerfBranch(3)
    net = m.network
    assert net.pts_linears[0].weight.abs().max() <= 1 / 3 and net.pts_linears[0].weight.abs().max() > 0.3
    lim = float(np.sqrt(6 / 256) / 25)
    for w in (net.pts_linears[1].weight, net.rgb_linear.weight, net.sigma_linear.weight):
        assert w.abs().max() <= lim * (1 + 1e-6) and w.abs().max() > 0.9 * lim
    assert net.views_linears.weight.abs().max() <= float(np.sqrt(6 / 259) / 25) * (1 + 1e-6)
    g = net.pts_linears[1].gamma.weight
    assert abs(float(g.std()) - 0.25 * np.sqrt(2 / (1 + 0.04)) / 16) < 2e-3
    assert float(m.sigmoid_beta) == pytest.approx(0.1)


def test_camera_mirror_matches_oracle_on_cpu():
    import cips3dpp_b200 as c3d
    from oracle import nerf_oracle as O
    locs = np.array([[0.0, 0.0], [0.25, -0.1], [-3.0, 0.16], [0.0, 1.5707963]], np.float32)
    out = c3d.Camera.generate_camera_params(64, "cpu", locations=torch.from_numpy(locs), fov_ang=6, dist_radius=0.12)
    ref = O.generate_camera_params(locs, 64, 6, 0.12)
    for a, b in zip(out, ref):
        np.testing.assert_allclose(a.numpy(), b, atol=2e-6, rtol=1e-6)
    ext, focal, near, far, vp = c3d.Camera.generate_camera_params(64, "cpu", batch=3, sweep=True)
    assert ext.shape == (24, 3, 4) and focal.shape == (24, 1, 1) and vp.shape == (24, 2)
    np.testing.assert_allclose(vp[:8, 0].numpy(), np.linspace(-0.3, 0.3, 8), atol=1e-6)
    loc = torch.tensor([[0.1, 0.05]], requires_grad=True)
    ext = c3d.Camera.generate_camera_params(64, "cpu", locations=loc)[0]
    ext.sum().backward()
    assert loc.grad is not None and loc.grad.abs().sum() > 0


def test_camera_v1_matches_reference_golden():
    """Camera.generate_camera_params_v1 (caller-given up vector, nerf_utils.py:466-560) vs the reference's own output
    (tests/golden/camera_v1.npz, written by make_golden.py --camera-v1)."""
    import cips3dpp_b200 as c3d
    g = np.load(os.path.join(ROOT, "tests", "golden", "camera_v1.npz"))
    out = c3d.Camera.generate_camera_params_v1(64, "cpu", locations=torch.from_numpy(g["locations"]),
                                               up=torch.from_numpy(g["up"]), fov_ang=15, dist_radius=0.3)
    for a, k in zip(out, ("c2w", "focal", "near", "far", "viewpoint")):
        np.testing.assert_allclose(a.numpy(), g[k], atol=2e-6, rtol=1e-6)
    base = c3d.Camera.generate_camera_params(64, "cpu", locations=torch.from_numpy(g["locations"]), fov_ang=15, dist_radius=0.3)
    v1 = c3d.Camera.generate_camera_params_v1(64, "cpu", locations=torch.from_numpy(g["locations"]), fov_ang=15, dist_radius=0.3)
    assert all(torch.equal(a, b) for a, b in zip(base, v1))


def test_get_camera2world_matches_rotation_vectors():
    """Camera.get_camera2world (nerf_utils.py:439-463; the reference delegates to pytorch3d's axis_angle_to_matrix, absent
    here): Rodrigues' formula against scipy's rotation-vector conversion, including tiny and near-pi angles."""
    from scipy.spatial.transform import Rotation
    import cips3dpp_b200 as c3d
    rng = np.random.default_rng(0)
    rv = rng.normal(0, 1.5, (64, 3))
    rv[0] = 0.0
    rv[1] = [1e-7, -2e-7, 5e-8]
    rv[2] = np.array([0.6, 0.0, 0.8]) * (np.pi - 1e-3)
    tr = rng.normal(0, 1, (64, 3))
    ext = c3d.Camera.get_camera2world(torch.from_numpy(rv), torch.from_numpy(tr), homo=True).numpy()
    assert ext.shape == (64, 4, 4)
    np.testing.assert_allclose(ext[:, :3, :3], Rotation.from_rotvec(rv).as_matrix(), atol=1e-12)
    np.testing.assert_allclose(ext[:, :3, 3], tr)
    np.testing.assert_array_equal(ext[:, 3], np.broadcast_to([0.0, 0.0, 0.0, 1.0], (64, 4)))
    e32 = c3d.Camera.get_camera2world(torch.from_numpy(rv).float().reshape(8, 8, 3), torch.from_numpy(tr).float().reshape(8, 8, 3))
    assert e32.shape == (8, 8, 3, 4)
    np.testing.assert_allclose(e32.reshape(64, 3, 4)[:, :, :3].numpy(), Rotation.from_rotvec(rv).as_matrix(), atol=2e-6)
