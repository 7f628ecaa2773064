"""BASELINE configs[4] shape: flip inversion of 16 targets (+ flips = 32 images per step), sharded over the ranks of a
torchrun launch (targets are independent: no data-path collective).  Reports the max-over-ranks time per optimisation
step (CUDA-graph replay of camera glue + forward + loss + backward + clipping + Adam; CUDA events around the loop)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cips3dpp_b200 as c3d
from oracle import nerf_oracle as O

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
D, targets, steps = int(os.environ.get("D", "2")), int(os.environ.get("TARGETS", "16")), int(os.environ.get("STEPS", "200"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lo, hi = c3d.dist.shard_range(targets, rank, world)
n = hi - lo
m = c3d.NerfBranch(D, precision="bf16")
m.load_state_dict({k: torch.from_numpy(v) for k, v in O.init_params(D).items()})
m = m.to(dev).eval().requires_grad_(False)
g = torch.Generator().manual_seed(5)
tgt_all = torch.rand(targets, 3, 64, 64, generator=g) * 2 - 1
tgt = tgt_all[lo:hi].to(dev)
w0 = torch.zeros(1, D + 1, 256, device=dev)
inv = c3d.FlipInversion(m, img_size=64, N_samples=24, num_steps=10)
inv.run(tgt, w0, cuda_graph=True)                                   # warm-up: allocator, capture path
inv.num_steps = steps
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
out = inv.run(tgt, w0, cuda_graph=True)
torch.cuda.synchronize()
ms = out["events"][0].elapsed_time(out["events"][1]) / steps         # device time of the replay loop (CUDA events)
x = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(x, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps(dict(what="flip inversion step (CUDA graph)", D=D, targets=targets, images_per_step=2 * targets, n_gpus=world,
                          ms_per_step=float(x.item()), images_per_s=2 * targets / (float(x.item()) * 1e-3),
                          final_loss_rank0=float(out["losses"][-1]))))
if world > 1:
    dist.destroy_process_group()
