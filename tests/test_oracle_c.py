"""The C restatement (oracle/nerf_oracle.c, the timed CPU baseline) against the reference's golden vectors and
the numpy oracle.  CPU only."""
import numpy as np
import pytest

from conftest import CASES, load_case, load_weights, rel_l2
from oracle import c_oracle, nerf_oracle as O


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference_golden(name):
    c = load_case(name)
    params = load_weights(int(c["D"]))
    params["sigmoid_beta"] = c["sigmoid_beta"].astype(np.float32)
    rgb_map, feat, sdf, mask, xyz = c_oracle.renderer_forward(
        params, c["pts"], c["rays_d"], c["viewdirs"], c["z_vals"], c["near"], c["far"], c["styles"])
    assert rel_l2(feat, c["feature_map"]) < 5e-5
    assert rel_l2(rgb_map, c["rgb_map"]) < 5e-5
    assert rel_l2(sdf, c["sdf"]) < 5e-5
    assert rel_l2(xyz, c["xyz"]) < 5e-5
    assert np.abs(mask - c["mask"]).max() < 1e-5


def test_c_prepare_inputs_matches_numpy_oracle():
    locs = np.array([[0.25, -0.1], [-1.3, 0.1]], np.float32)
    c2w, focal, near, far, _ = O.generate_camera_params(locs, 16, 15, 0.3)
    u = np.random.default_rng(0).uniform(size=(2, 16, 16, 1)).astype(np.float32)
    for sv, t in ((False, None), (True, u)):
        a = c_oracle.prepare_inputs(c2w, focal, near, far, 16, 24, sv, t)
        b = O.prepare_nerf_inputs(focal, 16, c2w, near, far, 24, t, sv)
        np.testing.assert_allclose(a[0].reshape(b[0].shape), b[0], atol=2e-6)
        np.testing.assert_allclose(a[1].reshape(b[1].shape), b[1], atol=1e-6)
        np.testing.assert_allclose(a[2].reshape(b[2].shape), b[2], atol=1e-6)
        np.testing.assert_allclose(a[3].reshape(b[3].shape), b[3], atol=1e-6)
