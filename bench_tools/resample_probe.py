"""Device timing of c3d_sample_pdf (importance resampling) against the HBM roofline, plus the two-pass render.
Run on the GPU box:  python bench_tools/resample_probe.py [--rays 1048576]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cips3dpp_b200 as c3d  # noqa: E402


def time_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2], t[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=256 * 4096)
    ap.add_argument("--render", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    R = a.rays
    for N, K in ((24, 24), (128, 128)):
        Rn = R if N == 24 else R // 8
        z = 0.88 + 0.24 * (torch.arange(N, device=dev)[None] + torch.rand(Rn, 1, device=dev)) / N
        sdf = (1.0 + 0.1 * torch.rand(Rn, 1, device=dev) - z) * 2.0
        w = torch.rand(Rn, N, device=dev) ** 8
        rd = torch.randn(Rn, 3, device=dev)
        ro = torch.randn(Rn, 3, device=dev)
        u = torch.rand(Rn, K, device=dev)
        sb = torch.tensor([0.1], device=dev)
        cases = {
            "weights->z": (lambda: c3d.Render.importance_depths(z, K, weights=w), 4 * N * 2 + 4 * K + 4 * (N + K)),
            "sdf->z+pts": (lambda: c3d.Render.importance_depths(z, K, sdf=sdf, rays_d=rd, rays_o=ro, sigmoid_beta=sb,
                                                                return_pts=True),
                           4 * N * 2 + 24 + 4 * K + 16 * (N + K)),
            "sdf,u->z+pts": (lambda: c3d.Render.importance_depths(z, K, sdf=sdf, rays_d=rd, rays_o=ro, sigmoid_beta=sb,
                                                                  u=u, return_pts=True),
                             4 * N * 2 + 24 + 4 * K + 4 * K + 16 * (N + K)),
        }
        for name, (fn, bpr) in cases.items():
            med, mn = time_ms(fn)
            print(json.dumps({"kernel": "sample_pdf_kernel", "case": name, "rays": Rn, "N": N, "K": K, "ms": round(med, 4),
                              "ms_min": round(mn, 4), "bytes_per_ray": bpr, "GBps": round(bpr * Rn / med / 1e6, 1),
                              "frac_hbm": round(bpr * Rn / med / 1e6 / peak, 3), "rays_per_s": round(Rn / med * 1e3),
                              "variant": os.environ.get("C3D_RESAMPLE", "auto")}), flush=True)
    if a.render:
        for D, b in ((8, 256), (2, 256)):
            m = c3d.NerfBranch(D, precision="bf16").to(dev).eval().requires_grad_(False)
            pose, focal, near, far, _ = c3d.Camera.generate_camera_params(64, dev, batch=b // 8, sweep=True)
            styles = 0.6 * torch.randn(b, D + 1, 256, device=dev)
            with torch.no_grad():
                one, _ = time_ms(lambda: m.render(pose, focal, near, far, styles, img_size=64, N_samples=24), reps=5, warm=2)
                two, _ = time_ms(lambda: m.render_hierarchical(pose, focal, near, far, styles, img_size=64, N_samples=24,
                                                               N_importance=24), reps=5, warm=2)
            print(json.dumps({"render": f"D={D} b={b}", "single_pass_ms": round(one, 3), "hierarchical_ms": round(two, 3),
                              "images_per_s_hier": round(b / two * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
