"""BASELINE configs[2] for real (run on the B200 box): the UNMODIFIED reference `model_v3.Generator` -- mapping networks,
ray chunking, re-layout glue and the modulated-conv decoder up to 1024 x 1024 -- once with its own `VolumeFeatureRenderer`
and once with `use_b200_nerf_branch(G)`, on the same latents, cameras and decoder noise.

The reference modules come from oracle/_ref (placed there by oracle/make_ref.sh in the build container; git-ignored, travels
with the snapshot) behind the import stubs of tests/ref_stubs.py; the decoder's two `op` CUDA extensions are replaced by
pure-torch stand-ins in BOTH arms (decoder side, out of scope).  Skipped where the reference modules are absent.

Tolerances (north star): fp32 mode 1e-3 rel-L2 on thumb_rgb / rgb, 1e-4 abs on depth; bf16 mode 2e-2 rel-L2."""
import copy

import numpy as np
import pytest
import torch

import ref_stubs
from conftest import rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_stubs.available(), reason="reference modules not present (oracle/make_ref.sh)")]


@pytest.fixture(scope="module")
def ref():
    model_v3, nerf_utils = ref_stubs.import_model_v3()
    return model_v3, nerf_utils


def _inputs(nerf_utils, G, S, b, dev, seed):
    g = torch.Generator().manual_seed(seed)
    zs = [torch.randn(b, 256, generator=g).to(dev), torch.randn(b, 256, generator=g).to(dev)]
    loc = torch.stack([0.6 * torch.rand(b, generator=g) - 0.3, 0.3 * torch.rand(b, generator=g) - 0.15], 1).to(dev)
    pose, focal, near, far, _ = nerf_utils.Camera.generate_camera_params(img_size=S, device=dev, locations=loc, fov_ang=6,
                                                                         dist_radius=0.12)
    torch.manual_seed(seed + 1)
    noise_bufs = G.create_noise_bufs(start_size=S, device=dev)
    return dict(zs=zs, cam_poses=pose, focals=focal, img_size=S, near=near, far=far, truncation=1, return_sdf=True,
                return_xyz=True, noise_bufs=noise_bufs, nerf_cfg=dict(N_samples=24, perturb=False, static_viewdirs=False))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("D,size_end,ups,b,chunk", [
    (2, 1024, (128, 256, 512, 1024), 2, None),     # the shipped r1024 generator: 64^2 NeRF branch -> 1024^2 image
    (8, 64, (), 3, None),                          # the D = 8 renderer of the r64 stage
    (2, 256, (128, 256), 1, 1001),                 # batch 1 with an odd ray chunk: unaligned views of rays_d (ADVICE)
])
def test_reference_generator_with_and_without_the_b200_branch(ref, precision, D, size_end, ups, b, chunk):
    import cips3dpp_b200 as c3d
    model_v3, nerf_utils = ref
    dev = torch.device("cuda:0")
    torch.manual_seed(D)
    G = ref_stubs.build_generator(model_v3, D=D, size_end=size_end, upsample_list=ups).to(dev).eval().requires_grad_(False)
    G2 = c3d.use_b200_nerf_branch(copy.deepcopy(G), precision=precision)
    assert isinstance(G2.renderer, c3d.NerfBranch) and not any(p.requires_grad for p in G2.renderer.parameters())
    kw = _inputs(nerf_utils, G, 64, b, dev, seed=3)
    with torch.no_grad():
        want = G(N_rays_forward=chunk, **kw)
        got = G2(N_rays_forward=chunk, **kw)
    torch.cuda.synchronize()
    assert set(want.keys()) == set(got.keys())
    assert got["rgb"].shape == (b, 3, size_end, size_end)
    tol = 1e-3 if precision == "fp32" else 2e-2
    errs = {}
    for k in ("thumb_rgb", "rgb", "xyz", "sdf", "mask", "depth"):
        a, r = got[k].float().cpu().numpy(), want[k].float().cpu().numpy()
        assert a.shape == r.shape, k
        errs[k] = rel_l2(a, r)
    print("dropin", precision, D, size_end, b, chunk, {k: f"{v:.2e}" for k, v in errs.items()})
    for k in ("thumb_rgb", "rgb", "xyz"):
        assert errs[k] < tol, (k, errs)
    assert errs["sdf"] < (1e-3 if precision == "fp32" else 4.5e-2), errs
    d = np.abs(got["depth"].cpu().numpy() - want["depth"].cpu().numpy()).max()
    assert d < (1e-4 if precision == "fp32" else 2e-3), d
