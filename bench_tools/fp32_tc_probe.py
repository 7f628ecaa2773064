"""fp32 mode: tensor-core MLP (option fp32=tc, default) against the FP32-pipe kernel (fp32=simt) and the C oracle on full
64x64 images of BASELINE configs[1] / [3], plus the time of a 256-image step."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cips3dpp_b200 as c3d
from oracle import nerf_oracle as O
from oracle import c_oracle
dev = torch.device("cuda:0")
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
for D, N, nimg in ((8, 24, 2), (2, 24, 2), (6, 36, 1)):
    S = 64
    params = O.init_params(D=D, seed=3)
    rng = np.random.default_rng(D)
    locs = np.stack([rng.uniform(-0.3, 0.3, nimg), rng.uniform(-0.15, 0.15, nimg)], 1).astype(np.float32)
    c2w, focal, near, far, _ = O.generate_camera_params(locs, S, 6, 0.12)
    styles = (0.6 * rng.standard_normal((nimg, D + 1, 256))).astype(np.float32)
    pts, rd, vd, z = c_oracle.prepare_inputs(c2w, focal, near, far, S, N)
    ref = dict(zip(("rgb_map", "feature_map", "sdf", "mask", "xyz"), c_oracle.renderer_forward(params, pts, rd, vd, z, near, far, styles)))
    m = c3d.NerfBranch(D, precision="fp32")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    m = m.to(dev).eval().requires_grad_(False)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    outs = {}
    for mode in ("tc", "simt"):
        c3d._abi.set_options(fp32=mode)
        with torch.no_grad():
            out = m.render(t(c2w), t(focal), t(near), t(far), t(styles), img_size=S, N_samples=N)
        torch.cuda.synchronize()
        outs[mode] = {k: out[k].cpu().numpy() for k in ("rgb_map", "feature_map", "sdf", "mask", "xyz")}
        e = {k: rel(outs[mode][k], ref[k]) for k in ("feature_map", "rgb_map", "xyz", "sdf")}
        e["depth_max_abs"] = float(np.abs(outs[mode]["mask"][..., 1] - ref["mask"][..., 1]).max())
        print(f"D={D} N={N} fp32={mode} vs C oracle:", json.dumps({k: float(f"{v:.3e}") for k, v in e.items()}))
    print(f"D={D} tc vs simt: feature_map {rel(outs['tc']['feature_map'], outs['simt']['feature_map'].astype(np.float64)):.3e}")
# time of a configs[1] step
D, N, S, B = 8, 24, 64, 256
params = O.init_params(D=D, seed=0)
m = c3d.NerfBranch(D, precision="fp32")
m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
m = m.to(dev).eval().requires_grad_(False); m.cache_packed = True
rng = np.random.default_rng(0)
locs = np.stack([rng.uniform(-0.3, 0.3, B), rng.uniform(-0.15, 0.15, B)], 1).astype(np.float32)
c2w, focal, near, far, _ = O.generate_camera_params(locs, S, 6, 0.12)
styles = torch.from_numpy((0.6 * rng.standard_normal((B, D + 1, 256))).astype(np.float32)).to(dev)
args = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (c2w, focal, near, far)]
for mode in ("tc", "simt"):
    c3d._abi.set_options(fp32=mode)
    with torch.no_grad():
        m.render(*args, styles, img_size=S, N_samples=N)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 3 if mode == "tc" else 1
        e0.record()
        for _ in range(K): m.render(*args, styles, img_size=S, N_samples=N)
        e1.record(); torch.cuda.synchronize()
    print(f"fp32={mode}: {e0.elapsed_time(e1) / K:.1f} ms per 256-image step (D=8, N=24)")
