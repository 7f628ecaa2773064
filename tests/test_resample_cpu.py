"""CPU checks of the importance-resampling EXTENSION's oracle (oracle/nerf_oracle.py::sample_pdf, importance_depths).

The reference has no sample_pdf, so there are no reference vectors: PARITY UNPINNED.  The numpy oracle
is cross-checked here against an independent torch restatement of the published algorithm (cumsum + torch.searchsorted
+ gather, the formulation NeRF and pi-GAN use) and against properties of inverse-transform sampling."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from resample_checks import check_samples, det_u, make_rays


def torch_sample_pdf(bins, weights, K, u=None):
    """Independent restatement (NeRF, Mildenhall et al. 2020, sec. 5.2) with torch.searchsorted."""
    bins, weights = torch.as_tensor(bins), torch.as_tensor(weights)
    weights = weights + 1e-5
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    if u is None:
        u = torch.linspace(0.0, 1.0, K).expand(*cdf.shape[:-1], K)
    u = torch.as_tensor(u).contiguous()
    inds = torch.searchsorted(cdf.contiguous(), u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_g = torch.stack([torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)], -1)
    bins_g = torch.stack([torch.gather(bins, -1, below), torch.gather(bins, -1, above)], -1)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return (bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])).numpy()


@pytest.mark.parametrize("N,K", [(3, 1), (24, 24), (33, 40), (128, 64)])
@pytest.mark.parametrize("det", [True, False])
def test_oracle_matches_torch_searchsorted(N, K, det):
    z, w = make_rays(200, N, seed=N * 10 + K)
    u = None if det else np.random.default_rng(5).random((200, K)).astype(np.float32)
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    ours = O.sample_pdf(mids, w[:, 1:-1], K, u)
    ref = torch_sample_pdf(mids, w[:, 1:-1], K, u)
    covered = check_samples(z, w, ours, det_u(200, K) if det else u, ref)     # tolerances: resample_checks.py
    assert covered > 0.5
    assert np.median(np.abs(ours - ref)) <= 1e-7


def test_sampling_properties():
    N, K, R = 24, 64, 300
    z, w = make_rays(R, N, seed=1)
    u = np.sort(np.random.default_rng(2).random((R, K)).astype(np.float32), -1)
    z_fine, z_merged = O.importance_depths(z, w, K, u)
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    assert (z_fine >= mids[:, :1] - 1e-6).all() and (z_fine <= mids[:, -1:] + 1e-6).all()
    assert (np.diff(z_fine, axis=-1) >= -1e-6).all()                      # monotone in u
    assert (np.diff(z_merged, axis=-1) >= 0).all()
    np.testing.assert_array_equal(z_merged, np.sort(np.concatenate([z, z_fine], -1), -1))
    # uniform weights -> samples uniformly spread over [mids[0], mids[-1]]
    zu, _ = O.importance_depths(z, np.ones_like(w), K, None)
    expect = mids[:, :1] + (mids[:, -1:] - mids[:, :1]) * np.linspace(0, 1, K, dtype=np.float32)[None]
    assert np.abs(zu - expect).max() < 1e-5
    # all the weight on sample k -> (nearly) every new depth falls into the bin around z_k
    k = 10
    wk = np.zeros_like(w); wk[:, k] = 1.0
    zk, _ = O.importance_depths(z, wk, K, u)
    inside = (zk >= mids[:, k - 1:k] - 1e-6) & (zk <= mids[:, k:k + 1] + 1e-6)
    assert inside.mean() > 0.99


def test_compositing_weights_match_volume_integration():
    rng = np.random.default_rng(0)
    R, N = 50, 24
    z, _ = make_rays(R, N, seed=3)
    sdf = rng.normal(0, 0.1, (R, N)).astype(np.float32)
    rays_d = rng.normal(0, 1, (R, 3)).astype(np.float32)
    rgb = rng.normal(0, 1, (R, N, 3)).astype(np.float32)
    pts = rng.normal(0, 1, (R, N, 3)).astype(np.float32)
    w_ref = O.volume_integration(rgb, sdf[..., None], None, z, rays_d, pts, 0.1)[4][..., 0]
    np.testing.assert_allclose(O.compositing_weights(sdf, z, rays_d, 0.1), w_ref, rtol=0, atol=1e-7)


def test_sample_pdf_argument_validation_without_gpu():
    import cips3dpp_b200 as c3d
    c3d._abi.build_library()
    lib = c3d._abi.load()
    P = c3d._abi.ResampleParams()
    assert ctypes.sizeof(P) == 8 + 4 + 4 + 4 + 4 + 10 * 8
    assert lib.c3d_sample_pdf(P, None) == -1 and b"n_rays" in lib.c3d_last_error()
    P.n_rays, P.n_samples, P.n_importance = 10, 2, 8
    assert lib.c3d_sample_pdf(P, None) == -1 and b"n_samples" in lib.c3d_last_error()
    P.n_samples, P.n_importance = 24, 0
    assert lib.c3d_sample_pdf(P, None) == -1 and b"n_importance" in lib.c3d_last_error()
    P.n_importance = 24
    assert lib.c3d_sample_pdf(P, None) == -1 and b"z_vals" in lib.c3d_last_error()
    with pytest.raises(RuntimeError, match="CUDA"):
        c3d.Render.importance_depths(torch.zeros(2, 24), 8, weights=torch.zeros(2, 24))
