"""POINTS entry (the reference Generator's call: pts / rays_d / viewdirs / z_vals tensors) vs POSES entry (rays in-kernel)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cips3dpp_b200 as c3d
dev = torch.device("cuda:0")


def time_ms(fn, reps=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for D, b, prec in ((8, 256, "bf16"), (2, 256, "bf16"), (8, 16, "fp32"), (2, 16, "fp32")):
    m = c3d.NerfBranch(D, precision=prec).to(dev).eval().requires_grad_(False)
    pose, focal, near, far, _ = c3d.Camera.generate_camera_params(64, dev, batch=b // 8, sweep=True)
    styles = 0.6 * torch.randn(b, D + 1, 256, device=dev)
    with torch.no_grad():
        pts, rays_d, viewdirs, z = c3d.Render.prepare_nerf_inputs(focal=focal, img_size=64, cam_poses=pose, near=near, far=far,
                                                                  N_samples=24, perturb=False)
        pts, rays_d, viewdirs, z = pts.reshape(b, 4096, 24, 3), rays_d.reshape(b, 4096, 3), viewdirs.reshape(b, 4096, 3), z.reshape(b, 4096, 24)
        t_pose = time_ms(lambda: m.render(pose, focal, near, far, styles, img_size=64, N_samples=24))
        t_pts = time_ms(lambda: m(pts=pts, rays_d=rays_d, viewdirs=viewdirs, z_vals=z, near=near, far=far, styles=styles))
        t_prep = time_ms(lambda: c3d.Render.prepare_nerf_inputs(focal=focal, img_size=64, cam_poses=pose, near=near, far=far,
                                                                N_samples=24, perturb=False))
    print(json.dumps(dict(D=D, images=b, precision=prec, poses_entry_ms=round(t_pose, 3), points_entry_ms=round(t_pts, 3),
                          prepare_nerf_inputs_ms=round(t_prep, 3), images_per_s_points=round(b / t_pts * 1e3))), flush=True)
