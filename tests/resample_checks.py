"""Shared checks for the importance-resampling extension (used by the CPU oracle tests and the GPU parity tests).

Inverse-CDF sampling is ill-conditioned where the PDF is (nearly) flat: a depth moves by  d(cdf) / density, so the 1e-7
rounding differences between two fp32 prefix sums become ~1e-4 of depth inside a bin that carries only the +1e-5 floor.
(The opposite holds in a heavy bin: there the CDF is steep and one fp32 ulp of depth is already ~5e-5 of CDF.)
The checks therefore are
  (a) position space, the north star's tolerance: |z - z_ref| <= 1e-4 for every sample whose bin holds >= 1e-3 of the
      PDF mass (the bins importance sampling exists for), and
  (b) every remaining sample must satisfy (a) or, in CDF space, give back its u within 2e-5 when pushed through the
      float64 piecewise-linear CDF (1e-5 of that is the algorithm's own slack: a CDF step below 1e-5 is treated as 1,
      which pins the sample to the bin edge).
"""
import numpy as np

POS_TOL, CDF_TOL, MASS_MIN = 1e-4, 2e-5, 1e-3


def make_rays(R, N, seed, peaked=True):
    """Ascending depths (offset sampling, nerf_utils.py:97-119) and surface-like or diffuse weights."""
    rng = np.random.default_rng(seed)
    near, far = 0.88, 1.12
    z = (near + (far - near) * np.arange(N) / N).astype(np.float32)[None].repeat(R, 0)
    z = z + (rng.random((R, 1)) * (far - near) / N).astype(np.float32)
    if peaked:                                    # most of the weight in 1-3 neighbouring samples
        c = rng.integers(0, N, size=(R, 1))
        w = np.exp(-0.5 * ((np.arange(N)[None] - c) / rng.uniform(0.4, 2.0, (R, 1))) ** 2)
        w = w / w.sum(-1, keepdims=True) * rng.uniform(0.2, 1.0, (R, 1))
    else:
        w = rng.random((R, N)) / N
    return z, w.astype(np.float32)


def det_u(R, K):
    return np.broadcast_to(np.linspace(0.0, 1.0, K, dtype=np.float32), (R, K))


def check_samples(z, w, z_fine, u, z_ref):
    """z, w (R,N) coarse depths / weights; z_fine (R,K) samples under test; u (R,K); z_ref (R,K) oracle samples.
    Returns the fraction of samples covered by the position-space check."""
    z64, w64 = z.astype(np.float64), w.astype(np.float64)
    mids = 0.5 * (z64[:, 1:] + z64[:, :-1])
    ww = w64[:, 1:-1] + 1e-5
    pdf = ww / ww.sum(-1, keepdims=True)
    cdf = np.concatenate([np.zeros_like(pdf[:, :1]), np.cumsum(pdf, -1)], -1)
    assert np.isfinite(z_fine).all()
    assert (z_fine >= mids[:, :1] - 1e-6).all() and (z_fine <= mids[:, -1:] + 1e-6).all(), "sample outside the bins"
    F = np.stack([np.interp(z_fine[r].astype(np.float64), mids[r], cdf[r]) for r in range(z.shape[0])])
    err_cdf = np.abs(F - u.astype(np.float64))
    idx = np.clip((mids[:, None, :] <= z_ref[:, :, None].astype(np.float64)).sum(-1) - 1, 0, pdf.shape[-1] - 1)
    mass = np.take_along_axis(pdf, idx, -1)
    ok = mass >= MASS_MIN
    err = np.abs(z_fine.astype(np.float64) - z_ref.astype(np.float64))
    assert err[ok].max(initial=0.0) <= POS_TOL, f"depth error {err[ok].max():.3e} in a weighted bin"
    bad = (err > POS_TOL) & (err_cdf > CDF_TOL)
    assert not bad.any(), f"{bad.sum()} samples off in depth ({err[bad].max():.3e}) and in CDF space ({err_cdf[bad].max():.3e})"
    return float(ok.mean())
