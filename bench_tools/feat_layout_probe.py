"""Render time of BASELINE configs[1] per feature-map layout (CUDA events, L2 flushed): (b,hw,256) fp32 / (b,256,hw) fp32 / (b,256,hw) bf16."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, cips3dpp_b200 as c3d
from oracle import nerf_oracle as O
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda:0")
m = c3d.NerfBranch(cfg["D"], precision="bf16")
m.load_state_dict({k: torch.from_numpy(v) for k, v in O.init_params(cfg["D"], seed=0).items()})
m = m.to(dev).eval().requires_grad_(False); m.cache_packed = True
args = [torch.from_numpy(x).to(dev) for x in bench.workload(cfg)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for name, kw in (("nhwc_fp32", {}), ("nchw_fp32", dict(features_nchw=True)), ("nchw_bf16", dict(features_nchw="bf16"))):
    ms = []
    for i in range(13):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.render(*args, img_size=64, N_samples=cfg["N"], **kw); e1.record()
        torch.cuda.synchronize()
        if i >= 3: ms.append(e0.elapsed_time(e1))
    res[name] = round(float(np.mean(ms)), 3)
print(json.dumps(res))
