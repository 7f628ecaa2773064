"""Full-size pins (run on the B200 box): the workloads `bench.py` times -- BASELINE configs[1] (FFHQ, D=8, N=24, yaw sweep) and
configs[3] (CompCars cameras, D=6) -- rendered on the GPU and compared, image by image and over ALL 4096 rays, with the C
restatement of the reference (oracle/nerf_oracle.c, itself pinned to the reference's vectors by tests/test_oracle_c.py).

Tolerances (BASELINE.json north_star / SURVEY.md 8d): fp32 mode depths <= 1e-4 abs, maps <= 1e-3 rel-L2; bf16 mode <= 2e-2 rel-L2.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import c_oracle, nerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definitions only; nothing is timed here)

pytestmark = pytest.mark.gpu

# eight images of the step: different latents and different poses of the sweep (first / last / middle)
PICK = {"c2": [0, 3, 7, 60, 129, 190, 250, 255], "c4": [0, 1, 5, 11, 16, 23, 30, 31]}


@pytest.mark.parametrize("fwd", ["pair", "v3"])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("cfg_name", ["c2", "c4"])
def test_baseline_workload_images_match_c_oracle(cfg_name, precision, fwd):
    import cips3dpp_b200 as c3d
    if precision == "fp32" and fwd == "v3":
        pytest.skip("the forward-kernel option only affects bf16 mode")
    c3d._abi.set_options(fwd=fwd)
    cfg = bench.CONFIGS[cfg_name]
    D, N, S = cfg["D"], cfg["N"], bench.IMG
    c2w, focal, near, far, styles = (x[PICK[cfg_name]] for x in bench.workload(cfg))
    params = O.init_params(D, seed=0)
    pts, rd, vd, z = c_oracle.prepare_inputs(c2w, focal, near, far, S, N)
    ref = dict(zip(("rgb_map", "feature_map", "sdf", "mask", "xyz"),
                   c_oracle.renderer_forward(params, pts, rd, vd, z, near, far, styles)))
    dev = torch.device("cuda:0")
    m = c3d.NerfBranch(D, precision=precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    m = m.to(dev).eval().requires_grad_(False)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    with torch.no_grad():
        out = m.render(t(c2w), t(focal), t(near), t(far), t(styles), img_size=S, N_samples=N)
    torch.cuda.synchronize()
    got = {k: out[k].cpu().numpy() for k in ("rgb_map", "feature_map", "sdf", "mask", "xyz", "z_vals")}
    assert np.abs(got["z_vals"] - z).max() < 1e-4                      # sample depths: same closed form, either precision
    tol = 1e-3 if precision == "fp32" else 2e-2                         # the north star's bounds
    # what the 16-bit mode measures on B200 with IEEE half operands, both kernels, worst image (+ 50 %): c2 feature_map 9.5e-4,
    # rgb_map 6.9e-4, xyz 5.9e-5, sdf 1.3e-3, depth map 8.3e-6; c4 6.4e-4, 4.2e-4, 6.3e-5, 1.3e-3, 4.7e-5
    measured = dict(feature_map=1.5e-3, rgb_map=1.1e-3, xyz=1.0e-4, sdf=2.0e-3, depth=7.5e-5)
    worst = {}
    for i in range(len(PICK[cfg_name])):                               # per image: no averaging over the batch
        for k in ("feature_map", "rgb_map", "xyz", "sdf"):
            e = rel_l2(got[k][i], ref[k][i])
            worst[k] = max(worst.get(k, 0.0), e)
            assert e < tol, (cfg_name, precision, "image", PICK[cfg_name][i], k, e)
        d = float(np.abs(got["mask"][i, :, 1] - ref["mask"][i, :, 1]).max())   # depth map
        worst["depth"] = max(worst.get("depth", 0.0), d)
        assert d < (1e-4 if precision == "fp32" else 2e-3), (cfg_name, precision, "image", PICK[cfg_name][i], "depth", d)
    print(cfg_name, precision, fwd, {k: f"{v:.2e}" for k, v in worst.items()})
    if precision == "bf16":
        for k, v in worst.items():
            assert v < measured[k], (cfg_name, fwd, k, v)


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4] at full size: 16 synthetic targets + flips (32 images of 64 x 64 rays, N = 24, D = 2), 200 steps.
def _inversion_setup(D=2, n=16):
    import cips3dpp_b200 as c3d
    dev = torch.device("cuda:0")
    params_np = O.init_params(D, seed=0)
    params = {k: torch.from_numpy(v).to(dev) for k, v in params_np.items()}
    g = torch.Generator().manual_seed(7)
    w_true = (0.6 * torch.randn(n, 1, 256, generator=g)).repeat(1, D + 1, 1).to(dev)
    az = (0.3 * (torch.rand(n, 1, 1, generator=g) - 0.5)).to(dev) * torch.tensor([[[1.0], [-1.0]]], device=dev)
    el = (0.1 * (torch.rand(n, 1, 1, generator=g) - 0.5)).to(dev).expand(n, 2, 1).contiguous()

    def module(prec):
        m = c3d.NerfBranch(D, precision=prec)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in params_np.items()})
        return m.to(dev).eval().requires_grad_(False)
    return dev, params, w_true, az, el, module


def test_inversion_loss_curves_at_full_size_within_one_percent():
    """North star: inversion loss curves within 1 %.  `FlipInversion` through libc3dpp (bf16 eager, bf16 CUDA graph, fp32)
    against the SAME loop driven through torch autograd of the reference restatement (tests/torch_ref.py, fp32)."""
    import cips3dpp_b200 as c3d
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch_ref
    D, S, N, steps, n = 2, 64, 24, 200, 16
    dev, params, w_true, az, el, module = _inversion_setup(D, n)
    w0 = torch.zeros(1, D + 1, 256, device=dev)
    with torch.no_grad():
        targets = c3d.FlipInversion(module("fp32"), img_size=S, N_samples=N).render_thumbs(w_true, az, el)[0::2].contiguous()

    class RefRenderer:                                                # same .render API, torch autograd inside
        def render(self, pose, focal, near, far, styles, img_size, N_samples, static_viewdirs):
            outs = [torch_ref.render_thumb(params, pose[i:i + 8], focal[i:i + 8], near[i:i + 8], far[i:i + 8], styles[i:i + 8],
                                           img_size, N_samples, static_viewdirs)[0] for i in range(0, pose.shape[0], 8)]
            return dict(rgb_map=torch.cat(outs, 0))
    ref = c3d.FlipInversion(RefRenderer(), img_size=S, N_samples=N, num_steps=steps).run(targets, w0)["losses"].cpu().numpy()
    assert ref[-1] < 0.5 * ref[0]                                      # the loop optimises
    for prec, graph, bound in (("bf16", False, 1e-2), ("bf16", True, 1e-2), ("fp32", False, 1e-4)):
        ours = c3d.FlipInversion(module(prec), img_size=S, N_samples=N, num_steps=steps).run(targets, w0, cuda_graph=graph)
        rel = np.abs(ours["losses"].cpu().numpy() - ref) / np.abs(ref)
        print("inversion", prec, "graph" if graph else "eager", f"max rel {rel.max():.2e} mean rel {rel.mean():.2e}")
        assert rel.max() < bound, (prec, graph, rel.max())


def test_inversion_with_reference_shaped_loss_within_one_percent():
    """The reference's inversion loss is perceptual and reaches the renderer through BOTH maps (projector_v9.py:230-246,
    1106-1137: `renderer_detach=False` -> features -> decoder -> VGG, plus 50x the thumb term).  Same loop, seeded random-init
    VGG16 conv stack + a stand-in decoder entry (tests/inversion_loss.py), so the kernel's backward receives a dense
    `g_feature_map` on every step: 16 targets + flips, 200 steps, against torch autograd of the reference restatement.
    Measured on B200 (TF32 off and deterministic cuDNN in the loss networks): at step 0 fp32 mode agrees to 6e-7 and the 16-bit
    mode to 4e-5, but this loop is chaotic and the torch arm is not reproducible: over five runs of this very test its own
    final loss was 4.2558, 4.2596, 4.2650, 4.2692, 4.2964 (1 % spread; autograd's atomics).  Against whichever trajectory the
    torch arm produced, fp32 mode measured a maximum over the 200 steps of 0.59 - 1.2 % (mean 0.1 - 0.5 %) with either fp32
    kernel (FP32 pipe: 0.72 %, tensor cores: 0.59 - 1.2 %), the 16-bit mode 0.66 - 1.2 % (mean 0.3 - 0.5 %; with bfloat16 forward
    operands in round 1: 2.3 - 2.5 %, mean 0.9 %) -- both modes sit at the noise floor of the comparison.  The north star's 1 % is
    asserted where the comparison is reproducible (the thumb-MSE loop above: 4e-6 / 5e-4); here the bound is the floor plus
    margin: every step < 2.5 %, mean < 1 %, either mode."""
    import cips3dpp_b200 as c3d
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch_ref
    from inversion_loss import RefShapedLoss
    D, S, N, steps, n = 2, 64, 24, 200, 16
    dev, params, w_true, az, el, module = _inversion_setup(D, n)
    torch.backends.cudnn.allow_tf32 = False            # the loss networks (cuDNN convs) in full fp32 in both arms: TF32 alone
    torch.backends.cuda.matmul.allow_tf32 = False      # moves this loss by 1e-4 per evaluation and 0.7 % along the trajectory
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False   # the loss networks' backward, both arms
    loss_fn = RefShapedLoss(seed=3).to(dev)
    w0 = torch.zeros(1, D + 1, 256, device=dev)
    with torch.no_grad():
        targets = c3d.FlipInversion(module("fp32"), img_size=S, N_samples=N).render_thumbs(w_true, az, el)[0::2].contiguous()

    class RefRenderer:                                                # same .render API, torch autograd inside
        def render(self, pose, focal, near, far, styles, img_size, N_samples, static_viewdirs, features_nchw=False):
            rgb, feat = [], []
            for i in range(0, pose.shape[0], 8):                      # chunks of 8 images bound the autograd graph's memory
                o = torch_ref.render_thumb(params, pose[i:i + 8], focal[i:i + 8], near[i:i + 8], far[i:i + 8], styles[i:i + 8],
                                           img_size, N_samples, static_viewdirs)
                rgb.append(o[0]); feat.append(o[1])
            f = torch.cat(feat, 0)
            return dict(rgb_map=torch.cat(rgb, 0), feature_map=f.transpose(1, 2) if features_nchw else f)
    kw = dict(img_size=S, N_samples=N, num_steps=steps, loss_fn=loss_fn, loss_on_features=True)
    ref = c3d.FlipInversion(RefRenderer(), **kw).run(targets, w0)["losses"].cpu().numpy()
    assert ref[-1] < 0.7 * ref[0]                                      # the loop optimises
    for prec, graph, bound in (("bf16", True, 2.5e-2), ("fp32", False, 2.5e-2)):
        ours = c3d.FlipInversion(module(prec), **kw).run(targets, w0, cuda_graph=graph)
        rel = np.abs(ours["losses"].cpu().numpy() - ref) / np.abs(ref)
        msg = (f"inversion (reference-shaped loss) {prec} {'graph' if graph else 'eager'}: first {ref[0]:.4f} last {ref[-1]:.4f} "
               f"max rel {rel.max():.2e} at step {int(rel.argmax())} mean rel {rel.mean():.2e} rel@[0,50,100,150,199] "
               f"{[float(f'{rel[i]:.2e}') for i in (0, 50, 100, 150, 199)]}")
        print(msg, flush=True)
        assert rel.max() < bound, msg
        assert rel.mean() < 1e-2, msg
