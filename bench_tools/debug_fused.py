"""Dev helper (GPU box): run one golden case through one kernel variant in its own process and print errors.
usage: python bench_tools/debug_fused.py CASE PRECISION [points|poses]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case, load_weights, rel_l2  # noqa: E402
import cips3dpp_b200 as c3d  # noqa: E402

case, prec = sys.argv[1], sys.argv[2]
kind = sys.argv[3] if len(sys.argv) > 3 else "points"
c = load_case(case)
D = int(c["D"])
dev = torch.device("cuda:0")
m = c3d.NerfBranch(D, precision=prec)
sd = {k: torch.from_numpy(v) for k, v in load_weights(D).items()}
sd["sigmoid_beta"] = torch.from_numpy(c["sigmoid_beta"].astype(np.float32).reshape(1))
m.load_state_dict(sd)
m = m.to(dev).eval().requires_grad_(False)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
with torch.no_grad():
    if kind == "points":
        o = m(pts=t(c["pts"]), rays_d=t(c["rays_d"]), viewdirs=t(c["viewdirs"]), z_vals=t(c["z_vals"]),
              near=t(c["near"]), far=t(c["far"]), styles=t(c["styles"]))
        out = dict(rgb_map=o[0], feature_map=o[1], sdf=o[2], mask=o[3], xyz=o[4])
    else:
        out = m.render(t(c["c2w"]), t(c["focal"]), t(c["near"]), t(c["far"]), t(c["styles"]), img_size=64,
                       N_samples=int(c["N"]), static_viewdirs=bool(c["static_viewdirs"]))
        idx = torch.from_numpy(c["ray_idx"].astype(np.int64)).to(dev)
        out = {k: v[:, idx] for k, v in out.items() if v is not None}
torch.cuda.synchronize()
res = {}
for k in ("rgb_map", "feature_map", "sdf", "mask", "xyz"):
    a = out[k].cpu().numpy()
    res[k] = (rel_l2(a, c[k]), float(np.abs(a - c[k]).max()), bool(np.isfinite(a).all()))
print(case, prec, kind, "cluster=" + os.environ.get("C3D_CLUSTER", "default"),
      {k: f"rel {v[0]:.3e} max {v[1]:.3e} finite {v[2]}" for k, v in res.items()})
if res["feature_map"][0] > 0.05:
    a = out["feature_map"].cpu().numpy()
    print(" feat[0,0,:8] got", a[0, 0, :8], "want", c["feature_map"][0, 0, :8])
    print(" feat[0,5,:8] got", a[0, 5, :8], "want", c["feature_map"][0, 5, :8])
    print(" sdf[0,0,:6] got", out["sdf"].cpu().numpy()[0, 0, :6, 0], "want", c["sdf"][0, 0, :6, 0])
    per_ray = np.linalg.norm(a - c["feature_map"], axis=-1) / np.linalg.norm(c["feature_map"], axis=-1)
    print(" per-ray rel err (first 40):", np.round(per_ray[0, :40], 3))
