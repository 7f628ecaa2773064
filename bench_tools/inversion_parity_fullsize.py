"""Loss-curve parity of the flip inversion at BASELINE configs[4] size (16 targets + flips = 32 images of 64x64 rays, N = 24,
D = 2, 200 steps): cips3dpp_b200.FlipInversion through libc3dpp (bf16 and fp32 mode) against the SAME loop driven through
torch autograd of the reference formulation (tests/torch_ref.py, fp32).  The north star asks for loss curves within 1 %."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cips3dpp_b200 as c3d  # noqa: E402
import torch_ref  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
D, S, N = 2, 64, 24
steps, n = int(os.environ.get("STEPS", "200")), int(os.environ.get("TARGETS", "16"))
params_np = O.init_params(D, seed=0)
params = {k: torch.from_numpy(v).to(dev) for k, v in params_np.items()}
g = torch.Generator().manual_seed(7)
w_true = (0.6 * torch.randn(n, 1, 256, generator=g)).repeat(1, D + 1, 1).to(dev)
az = (0.3 * (torch.rand(n, 1, 1, generator=g) - 0.5)).to(dev) * torch.tensor([[[1.0], [-1.0]]], device=dev)
el = (0.1 * (torch.rand(n, 1, 1, generator=g) - 0.5)).to(dev).expand(n, 2, 1).contiguous()
w0 = torch.zeros(1, D + 1, 256, device=dev)


def module(prec):
    m = c3d.NerfBranch(D, precision=prec)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params_np.items()})
    return m.to(dev).eval().requires_grad_(False)


with torch.no_grad():
    targets = c3d.FlipInversion(module("fp32"), img_size=S, N_samples=N).render_thumbs(w_true, az, el)[0::2].contiguous()


class RefRenderer:
    def render(self, pose, focal, near, far, styles, img_size, N_samples, static_viewdirs):
        outs = []
        for i in range(0, pose.shape[0], 8):                      # chunks of 8 images bound the autograd graph's memory
            sl = slice(i, i + 8)
            outs.append(torch_ref.render_thumb(params, pose[sl], focal[sl], near[sl], far[sl], styles[sl], img_size, N_samples,
                                               static_viewdirs)[0])
        return dict(rgb_map=torch.cat(outs, 0))


ref = c3d.FlipInversion(RefRenderer(), img_size=S, N_samples=N, num_steps=steps).run(targets, w0)["losses"].cpu().numpy()
out = dict(targets=n, images=2 * n, steps=steps, ref_first=float(ref[0]), ref_last=float(ref[-1]))
for prec in ("bf16", "fp32"):
    for graph in ((False, True) if prec == "bf16" else (False,)):
        ours = c3d.FlipInversion(module(prec), img_size=S, N_samples=N, num_steps=steps).run(targets, w0, cuda_graph=graph)
        l = ours["losses"].cpu().numpy()
        rel = np.abs(l - ref) / np.abs(ref)
        out[f"{prec}{'_graph' if graph else ''}"] = dict(max_rel=float(rel.max()), mean_rel=float(rel.mean()), last=float(l[-1]))
print(json.dumps(out))
