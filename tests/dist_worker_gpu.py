"""Worker of tests/test_gpu_dist.py (one process per GPU under torchrun, NCCL over NVLink): the two collectives of the path.
 1. sharded render + dist.gather_maps (all_gather_into_tensor of rendered maps) == the same images rendered on one rank;
 2. flip inversion with one latent shared by all targets of all ranks (dist.allreduce_grads on d loss / d w) follows the
    loss curve of a single process fitting all targets.
Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cips3dpp_b200 as c3d  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402  (weights with the reference's init distributions; test infrastructure)

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
D = 2
m = c3d.NerfBranch(D, precision="bf16")
m.load_state_dict({k: torch.from_numpy(v) for k, v in O.init_params(D).items()})
m = m.to(dev).eval().requires_grad_(False)

# ---- 1. sharded render + all-gather of the maps (ragged: 7 images over the ranks)
n_img = 7
g = torch.Generator().manual_seed(1)
styles = (0.6 * torch.randn(n_img, D + 1, 256, generator=g)).to(dev)
loc = torch.stack([torch.linspace(-0.3, 0.3, n_img), torch.linspace(-0.1, 0.1, n_img)], 1).to(dev)
pose, focal, near, far, _ = c3d.Camera.generate_camera_params(64, dev, locations=loc, fov_ang=6, dist_radius=0.12)
s, e = c3d.dist.shard_range(n_img, rank, world)
with torch.no_grad():
    mine = m.render(pose[s:e], focal[s:e], near[s:e], far[s:e], styles[s:e], img_size=64, N_samples=24, features_nchw=True)
    feat = c3d.dist.gather_maps(mine["feature_map"], n_img)
    rgb = c3d.dist.gather_maps(mine["rgb_map"], n_img)
    full = m.render(pose, focal, near, far, styles, img_size=64, N_samples=24, features_nchw=True)
err_feat = float((feat - full["feature_map"]).norm() / full["feature_map"].norm())
err_rgb = float((rgb - full["rgb_map"]).norm() / full["rgb_map"].norm())

# ---- 1b. FUSED all-gather: the kernel's epilogue stores this rank's maps into the gathered tensors of every rank (peer memory)
fused = {}
try:
    Bf = 3
    nf = world * Bf
    gf = torch.Generator().manual_seed(2)
    styles_f = (0.6 * torch.randn(nf, D + 1, 256, generator=gf)).to(dev)
    loc_f = torch.stack([torch.linspace(-0.3, 0.3, nf), torch.linspace(-0.1, 0.1, nf)], 1).to(dev)
    pose_f, focal_f, near_f, far_f, _ = c3d.Camera.generate_camera_params(64, dev, locations=loc_f, fov_ang=6, dist_radius=0.12)
    gm = c3d.dist.GatheredMaps(Bf, 64 * 64, features="bf16")
    sl = slice(rank * Bf, (rank + 1) * Bf)
    with torch.no_grad():
        r = m.render(pose_f[sl], focal_f[sl], near_f[sl], far_f[sl], styles_f[sl], img_size=64, N_samples=24, gather=gm)
        gm.barrier()
        torch.cuda.synchronize()
        full_f = m.render(pose_f, focal_f, near_f, far_f, styles_f, img_size=64, N_samples=24, features_nchw="bf16")
    fused = dict(fused_feat_equal=bool(torch.equal(gm.feature_map, full_f["feature_map"])),
                 fused_rgb_equal=bool(torch.equal(gm.rgb_map, full_f["rgb_map"])),
                 fused_mask_equal=bool(torch.equal(gm.mask, full_f["mask"])), fused_xyz_equal=bool(torch.equal(gm.xyz, full_f["xyz"])),
                 fused_launches=int(m.last_launch_count))
    # the same tensors filled by peer-to-peer copies of the rank's shard (DMA engines)
    gm2 = c3d.dist.GatheredMaps(Bf, 64 * 64, features="bf16")
    with torch.no_grad():
        mine_f = m.render(pose_f[sl], focal_f[sl], near_f[sl], far_f[sl], styles_f[sl], img_size=64, N_samples=24, features_nchw="bf16")
    gm2.push(mine_f)
    gm2.barrier()
    torch.cuda.synchronize()
    fused["p2p_equal"] = bool(torch.equal(gm2.feature_map, full_f["feature_map"]) and torch.equal(gm2.rgb_map, full_f["rgb_map"])
                              and torch.equal(gm2.mask, full_f["mask"]) and torch.equal(gm2.xyz, full_f["xyz"]))
    ok = torch.tensor([int(all(v for k, v in fused.items() if k.endswith("equal")))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)                  # every rank must have received every rank's maps
    fused["fused_all_ranks_ok"] = bool(ok.item())
    del gm, gm2
except Exception as ex:  # noqa: BLE001  (symmetric memory unavailable on this box: reported, the NCCL path still covers the gather)
    fused = dict(fused_error=f"{type(ex).__name__}: {ex}"[:300])

# timing of the gather at the BASELINE configs[1] size per rank (256 images in total: 1 GiB of feature maps gathered on every rank)
big = torch.empty(256 // world, 256, 4096, device=dev).normal_()
for _ in range(2):
    c3d.dist.gather_maps(big, big.shape[0] * world)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = c3d.dist.gather_maps(big, big.shape[0] * world)
e1.record(); torch.cuda.synchronize()
gather_ms = e0.elapsed_time(e1) / 5
gather_gb = out.numel() * 4 / 1e9
del big, out

# ---- 2. shared-latent inversion: gradient all-reduce across ranks vs one process with all targets
n_t, steps = 4, 12
gt = torch.Generator().manual_seed(7)
targets = (torch.rand(n_t, 3, 32, 32, generator=gt) * 2 - 1).to(dev)
w0 = torch.zeros(1, D + 1, 256, device=dev)
ts, te = c3d.dist.shard_range(n_t, rank, world)
inv = c3d.FlipInversion(m, img_size=32, N_samples=24, num_steps=steps, shared_latent=True)
r_dist = inv.run(targets[ts:te], w0)
loss_dist = r_dist["losses"].clone()
dist.all_reduce(loss_dist)                                  # sum of the per-rank partial losses
w_dist = r_dist["w"]

dist.barrier()
dist.destroy_process_group()                                # single-process reference: no group -> no all-reduce
if rank == 0:
    r_one = c3d.FlipInversion(m, img_size=32, N_samples=24, num_steps=steps, shared_latent=True).run(targets, w0)
    rel = ((loss_dist - r_one["losses"]).abs() / r_one["losses"].abs()).max().item()
    w_err = float((w_dist - r_one["w"]).norm() / r_one["w"].norm())
    print(json.dumps(dict(world=world, **fused, gather_feat_rel=err_feat, gather_rgb_rel=err_rgb, gather_ms=gather_ms,
                          gather_GB=gather_gb, gather_GBps=gather_gb / (gather_ms * 1e-3),
                          inv_loss_rel=rel, inv_w_rel=w_err, first_loss=float(r_one["losses"][0]),
                          last_loss=float(r_one["losses"][-1]))))
