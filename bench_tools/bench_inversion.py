"""Dev benchmark: one flip-inversion step (forward + backward through the NeRF branch) at BASELINE config 5 shape
per GPU: 2 targets + flips = 4 images of 64x64 rays, N=24 (16 targets / 8 GPUs)."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cips3dpp_b200 as c3d
from oracle import nerf_oracle as O
D = int(os.environ.get("D", "2")); n = int(os.environ.get("TARGETS", "2")); prec = os.environ.get("PREC", "bf16")
dev = torch.device("cuda:0")
m = c3d.NerfBranch(D, precision=prec)
m.load_state_dict({k: torch.from_numpy(v) for k, v in O.init_params(D).items()})
m = m.to(dev).eval().requires_grad_(False)
inv = c3d.FlipInversion(m, img_size=64, N_samples=24, num_steps=1)
w = torch.zeros(n, D + 1, 256, device=dev, requires_grad=True)
az = torch.zeros(n, 2, 1, device=dev, requires_grad=True); el = torch.zeros(n, 2, 1, device=dev, requires_grad=True)
tgt = torch.rand(2 * n, 3, 64, 64, device=dev) * 2 - 1
def step():
    th = inv.render_thumbs(w, az, el)
    loss = ((th - tgt) ** 2).mean()
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
with torch.no_grad():
    e0.record()
    for _ in range(K): inv.render_thumbs(w, az, el)
    e1.record(); torch.cuda.synchronize()
print(json.dumps(dict(D=D, targets=n, images=2 * n, precision=prec, ms_fwd_bwd=ms, ms_fwd=e0.elapsed_time(e1) / K,
                      images_per_s=2 * n / (ms * 1e-3))))
# the whole optimisation step (camera glue, forward, loss, backward, clipping, Adam) eagerly and as a CUDA graph,
# 200 steps (BASELINE configs[4]), device time from the CUDA events FlipInversion.run records around its loop
tg = tgt[0::2].contiguous()
w0 = torch.zeros(1, D + 1, 256, device=dev)
inv.num_steps = 3
inv.run(tg, w0)                                                      # warm: optimiser / allocator first-use costs
inv.num_steps = int(os.environ.get("STEPS", "200"))
for graph in (False, True):
    r = inv.run(tg, w0, cuda_graph=graph)
    torch.cuda.synchronize()
    ms = r["events"][0].elapsed_time(r["events"][1]) / inv.num_steps
    print(json.dumps(dict(full_step_cuda_graph=graph, steps=inv.num_steps, ms_per_step=ms, images_per_s=2 * n / (ms * 1e-3),
                          save_fwd_gb=os.environ.get("C3D_SAVE_FWD_GB", "48"), final_loss=float(r["losses"][-1]))))
