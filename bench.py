#!/usr/bin/env python
"""Benchmark of the NeRF-branch hot path (BASELINE.json metric: rays/s & images/s at 64x64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c2d2|c3|c4|c1|c5]

One "step" = one pass of the hot path over one batch: BASELINE.json configs[1] -- FFHQ v10 NeRF branch
(D=8, N=24 samples, 64x64 rays), 32 latents x 8-pose yaw sweep = 256 images, bf16, random-init weights,
synthetic latents.  N>1: launched by torchrun, every rank renders its own 256 images (weak scaling, no
data-path collective), time = max over ranks.

Prints ONE JSON line (see the keys at the bottom).  `--impl reference` times the CPU restatement of the
reference path (oracle/nerf_oracle.c, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (D, N, latents, poses/latent, cam cfg, description)
    "c2": dict(D=8, N=24, latents=32, sweep=8, fov=6.0, radius=0.12, azim=0.3, elev=0.15,
               desc="FFHQ v10 NeRF branch 64x64, D=8, N=24, 32 latents x 8-pose yaw sweep (BASELINE configs[1])"),
    "c2d2": dict(D=2, N=24, latents=32, sweep=8, fov=6.0, radius=0.12, azim=0.3, elev=0.15,
                 desc="FFHQ v10 NeRF branch 64x64 at the shipped r1024 depth D=2, N=24, 32 latents x 8-pose yaw sweep"),
    "c3": dict(D=2, N=128, latents=8, sweep=1, fov=6.0, radius=0.12, azim=0.3, elev=0.15,
               desc="FFHQ v10 multi-view render NeRF branch 64x64, D=2, N=128, 8 images (BASELINE configs[2], NeRF part)"),
    # BASELINE configs[2] for real: the unmodified reference Generator (mapping networks, ray chunking, re-layout, modulated-
    # conv decoder to 1024 x 1024; modules from oracle/_ref, see oracle/make_ref.sh) around the B200 NeRF branch
    "c3full": dict(D=2, N=128, latents=8, sweep=1, fov=6.0, radius=0.12, azim=0.3, elev=0.15,
                   desc="FFHQ v10 full multi-view render 1024x1024: B200 NeRF branch at 64x64 (D=2, N=128) + the reference's "
                        "modulated-conv decoder, 8 images per GPU, sharded by latent (BASELINE configs[2])"),
    "c4": dict(D=6, N=24, latents=32, sweep=1, fov=15.0, radius=0.3, azim=3.14, elev=0.0837,
               desc="CompCars v10 NeRF branch 64x64, D=6, N=24, batch 32 (BASELINE configs[3])"),
    "c1": dict(D=8, N=24, latents=1, sweep=1, fov=6.0, radius=0.12, azim=0.0, elev=0.0,
               desc="FFHQ v10 NeRF branch 64x64, D=8, N=24, batch 1 (BASELINE configs[0])"),
    # flip inversion: a step = one optimisation step (forward + backward through the NeRF branch + Adam) over 16 targets
    # and their flips (32 images of 64x64 rays) per GPU; see run_inversion
    "c5": dict(D=2, N=24, latents=16, sweep=2, fov=6.0, radius=0.12, azim=0.3, elev=0.15,
               desc="flip inversion step, D=2, N=24, 16 synthetic targets + flips = 32 images per GPU (BASELINE configs[4] shape)"),
}
IMG = 64



OPERANDS = ("tcgen05 kind::f16, fp32 accumulate: IEEE half operands wherever the hidden-layer activations are read (layers, heads, "
            "compositing; same dense peak as bf16, 8x finer mantissa), bfloat16 K=16 side operands and backward")


def _dtype(precision):
    """Arithmetic type of the path: the 16-bit tensor-core mode (API name "bf16", BASELINE's bf16 mode) multiplies fp16 operands."""
    return "f16" if precision == "bf16" else "f32"


def workload(cfg, seed_latent=1, seed_pose=2):
    """Synthetic inputs of the configured shape (numpy, host): poses, focal, near, far, styles."""
    from oracle import nerf_oracle as O
    rng_l, rng_p = np.random.default_rng(seed_latent), np.random.default_rng(seed_pose)
    L, S = cfg["latents"], cfg["sweep"]
    if S == 8:
        locs = O.sweep_locations(L, cfg["azim"], cfg["elev"], rng_p.uniform(size=L))
    else:
        locs = np.stack([rng_p.uniform(-cfg["azim"], cfg["azim"], L * S),
                         rng_p.uniform(-cfg["elev"], cfg["elev"], L * S)], 1).astype(np.float32)
    c2w, focal, near, far, _ = O.generate_camera_params(locs, IMG, cfg["fov"], cfg["radius"])
    w = (0.6 * rng_l.standard_normal((L, 1, 256))).astype(np.float32)           # one latent per identity
    styles = np.repeat(np.repeat(w, cfg["D"] + 1, axis=1), S, axis=0)            # (L*S, D+1, 256)
    return c2w, focal.reshape(-1), near.reshape(-1), far.reshape(-1), np.ascontiguousarray(styles)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:                                    # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm overrides it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(cfg, n_images, reps=1, seeds=(1, 2), keep=None):
    """oracle/nerf_oracle.c (kind 'port': the reference path is Python, nothing to compile) on host cores.
    `keep` (a dict) receives the oracle's outputs for these images: the caller checks the GPU maps against them."""
    from oracle import c_oracle, nerf_oracle as O
    c2w, focal, near, far, styles = workload(cfg, seed_latent=seeds[0], seed_pose=seeds[1])
    params = O.init_params(cfg["D"], seed=0)
    packed = c_oracle.pack_params(params)
    sl = slice(0, n_images)
    times = []
    nthr = host_threads()
    for _ in range(reps + 1):                                      # first pass = warm-up
        t0 = time.perf_counter()
        pts, rd, vd, z = c_oracle.prepare_inputs(c2w[sl], focal[sl], near[sl], far[sl], IMG, cfg["N"])
        outs = c_oracle.renderer_forward(params, pts, rd, vd, z, near[sl], far[sl], styles[sl], nthreads=nthr, packed=packed)
        times.append(time.perf_counter() - t0)
    if keep is not None:
        keep.update(rgb_map=outs[0], feature_map=outs[1], sdf=outs[2], mask=outs[3], xyz=outs[4], z_vals=z)
    t = min(times[1:])
    out = dict(value=n_images * IMG * IMG / t, unit="rays/s", cores=nthr, kind="port",
               sample=f"{n_images} of {c2w.shape[0]} images of the step ({t:.2f} s, best of {reps})",
               images_per_s=n_images / t, host_cpus=os.cpu_count())
    live = live_reference_cpu(cfg, params, c2w[:2], focal[:2], near[:2], far[:2], styles[:2], nthr)
    if live is not None:
        out["reference_torch"] = live
    return out, t


def live_reference_cpu(cfg, params, c2w, focal, near, far, styles, nthr):
    """The LIVE reference beside the port: the unmodified `VolumeFeatureRenderer` + `Render.prepare_nerf_inputs` (PyTorch, CPU,
    all host threads) on two images of the step.  Available where oracle/make_ref.sh has placed the reference modules."""
    try:
        import torch
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ref_stubs
        if not ref_stubs.available():
            return None
        model_v3, nerf_utils = ref_stubs.import_model_v3()
        from exp.cips3d.volume_renderer import VolumeFeatureRenderer
        torch.set_num_threads(nthr)
        r = VolumeFeatureRenderer(N_layers_renderer=cfg["D"], input_dim=3, hidden_dim=256, style_dim=256, view_dim=3,
                                  with_sdf=True, output_features=True).eval()
        r.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
        t_ = lambda x: torch.from_numpy(np.ascontiguousarray(x))
        times = []
        with torch.no_grad():
            for _ in range(2):
                t0 = time.perf_counter()
                pts, rays_d, viewdirs, z_vals = nerf_utils.Render.prepare_nerf_inputs(
                    focal=t_(focal).view(-1, 1, 1), img_size=IMG, cam_poses=t_(c2w), near=t_(near).view(-1, 1, 1),
                    far=t_(far).view(-1, 1, 1), N_samples=cfg["N"], perturb=False)
                r(pts=pts, rays_d=rays_d, viewdirs=viewdirs, z_vals=z_vals, near=t_(near).view(-1, 1, 1), far=t_(far).view(-1, 1, 1),
                  styles=t_(styles))
                times.append(time.perf_counter() - t0)
        t = min(times)
        return dict(value=2 * IMG * IMG / t, unit="rays/s", images_per_s=2 / t, cores=nthr, kind="reference",
                    sample=f"2 images through the unmodified PyTorch VolumeFeatureRenderer on the host cores ({t:.2f} s)")
    except Exception as ex:  # noqa: BLE001  (the port's number stands on its own)
        return dict(error=f"{type(ex).__name__}: {ex}"[:200])


def run_reference(args, cfg, rank, world):
    """The reference arm: the reference's own implementation of the path on the box's host cores, all threads, on a bounded
    sample of the workload per step.  Where oracle/make_ref.sh has placed the reference modules (oracle/_ref) this is the
    LIVE reference -- the unmodified PyTorch `VolumeFeatureRenderer` + `Render.prepare_nerf_inputs` (kind "reference");
    otherwise the C / OpenMP port of the same algorithm, oracle/nerf_oracle.c (kind "port", roughly 2x faster than PyTorch)."""
    if rank != 0:
        return
    from oracle import c_oracle, nerf_oracle as O
    c2w, focal, near, far, styles = workload(cfg)
    params = O.init_params(cfg["D"], seed=0)
    nthr = host_threads()
    n_img = max(1, min(cfg["latents"] * cfg["sweep"], 4))
    sl = slice(0, n_img)
    live = None
    try:
        import torch
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ref_stubs
        if ref_stubs.available():
            model_v3, nerf_utils = ref_stubs.import_model_v3()
            from exp.cips3d.volume_renderer import VolumeFeatureRenderer
            torch.set_num_threads(nthr)
            live = VolumeFeatureRenderer(N_layers_renderer=cfg["D"], input_dim=3, hidden_dim=256, style_dim=256, view_dim=3,
                                         with_sdf=True, output_features=True).eval()
            live.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
            t_ = lambda x: torch.from_numpy(np.ascontiguousarray(x))
            tin = dict(focal=t_(focal[sl]).view(-1, 1, 1), cam_poses=t_(c2w[sl]), near=t_(near[sl]).view(-1, 1, 1),
                       far=t_(far[sl]).view(-1, 1, 1), styles=t_(styles[sl]))
    except Exception:  # noqa: BLE001
        live = None
    packed = c_oracle.pack_params(params)

    def step():
        if live is not None:
            with torch.no_grad():
                pts, rays_d, viewdirs, z_vals = nerf_utils.Render.prepare_nerf_inputs(
                    focal=tin["focal"], img_size=IMG, cam_poses=tin["cam_poses"], near=tin["near"], far=tin["far"],
                    N_samples=cfg["N"], perturb=False)
                live(pts=pts, rays_d=rays_d, viewdirs=viewdirs, z_vals=z_vals, near=tin["near"], far=tin["far"], styles=tin["styles"])
        else:
            pts, rd, vd, z = c_oracle.prepare_inputs(c2w[sl], focal[sl], near[sl], far[sl], IMG, cfg["N"])
            c_oracle.renderer_forward(params, pts, rd, vd, z, near[sl], far[sl], styles[sl], nthreads=nthr, packed=packed)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        step()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    v = n_img * IMG * IMG / t
    kind = "reference" if live is not None else "port"
    what = ("the unmodified PyTorch VolumeFeatureRenderer + Render.prepare_nerf_inputs (oracle/_ref)" if live is not None
            else "oracle/nerf_oracle.c (C / OpenMP port of the reference algorithm)")
    cb = dict(value=v, unit="rays/s", cores=nthr, kind=kind, images_per_s=n_img / t, host_cpus=os.cpu_count(),
              sample=f"each step = {n_img} of {c2w.shape[0]} images of the workload through {what}")
    print(json.dumps({
        "impl": "reference", "metric": "nerf_branch_rays_per_s", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "images_per_s": n_img / t,
        "config": {"workload": cfg["desc"], "sample_images_per_step": n_img, "img_size": IMG},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_inversion(args, cfg, rank, world, local_rank):
    """--config c5: one JSON line for the flip-inversion step (cips3dpp_b200.FlipInversion, stage 1 of projector_v9.py:
    862-1166).  value = rays/s through forward + backward + optimiser with targets resident (whole step replayed as a CUDA
    graph, CUDA events around the replay loop); e2e = the same loop with the targets copied pinned host -> device before and
    the loss copied back to pinned host memory after every step; roofline = algorithmic 2 F FLOPs (forward + input-gradient GEMMs) of the step / its time."""
    import torch
    import torch.distributed as dist
    import cips3dpp_b200 as c3d
    from oracle import nerf_oracle as O
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, N, n_t = cfg["D"], cfg["N"], cfg["latents"]
    # --shared-latent (BASELINE configs[4] as written): the 16 targets are SHARDED over the ranks (strong scaling), one latent
    # is shared by all of them, and every step all-reduces its gradient over NCCL / NVLink (projector_v9.py:1050)
    shared = bool(args.shared_latent) and world > 1
    if shared:
        n_t = max(1, n_t // world)
    m = c3d.NerfBranch(D, precision=args.precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in O.init_params(D, seed=0).items()}, strict=True)
    m = m.to(dev).eval().requires_grad_(False)
    m.cache_packed = True                        # frozen weights: pack them once (the default re-packs on every call)
    g = torch.Generator().manual_seed(5 + rank)
    host_t = (torch.rand(n_t, 3, IMG, IMG, generator=g) * 2 - 1).pin_memory()
    tgt = host_t.to(dev)
    w0 = torch.zeros(1, D + 1, 256, device=dev)
    inv = c3d.FlipInversion(m, img_size=IMG, N_samples=N, num_steps=max(args.warmup, 3), shared_latent=shared)
    inv.run(tgt, w0)                                                  # warm-up (eager), then the graph path once
    inv.run(tgt, w0, cuda_graph=True)
    wq = torch.zeros(n_t, D + 1, 256, device=dev, requires_grad=True)       # count this library's launches of one step
    aq = torch.zeros(n_t, 2, 1, device=dev, requires_grad=True)
    th = inv.render_thumbs(wq, aq, aq.detach().clone().requires_grad_(True))
    launches = m.last_launch_count
    th.sum().backward()
    launches += m.last_launch_count + 2                                    # + c3d_camera_params, c3d_adam_clip_step
    inv.num_steps = args.steps

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync()
    r = inv.run(tgt, w0, cuda_graph=True)
    sync()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_step = r["events"][0].elapsed_time(r["events"][1]) / args.steps
    r2 = inv.run(tgt, w0, cuda_graph=True, host_targets=host_t)
    sync()
    ms_e2e = r2["events"][0].elapsed_time(r2["events"][1]) / args.steps
    if world > 1:
        t = torch.tensor([ms_step, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        imgs = 2 * n_t
        rays = imgs * IMG * IMG
        flops = 2 * O.flops_per_point(D) * rays * N
        print(json.dumps({
            "metric": "nerf_branch_rays_per_s", "value": world * rays / (ms_step * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if shared else "weak", "vs_baseline": None, "dtype": _dtype(args.precision), "data": "synthetic",
            "images_per_s": world * imgs / (ms_step * 1e-3),
            "config": {"workload": cfg["desc"], "images_per_gpu": imgs,
                       "collective": "all-reduce (NCCL, captured in the step's CUDA graph) of the shared latent's gradient every step"
                                     if shared else "none (each rank fits its own targets)", "rays_per_image": IMG * IMG, "samples_per_ray": N,
                       "layers": D, "step": "forward + backward (styles, cameras) + clipping + Adam, CUDA-graph replay",
                       "l2": "not flushed: the step's working set (saved tiles, 0.3 GB per image) is far larger than L2",
                       "final_loss": float(r["losses"][-1])},
            "e2e": {"value": world * rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(host_t.numel() * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches * args.steps),
            "roofline": {"bound": "tensor", "achieved": flops / (ms_step * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": flops / (ms_step * 1e-3) / 1e12 / peak_tf, "traffic": None,
                         "kernel": "whole step (fused_forward_kernel<save> + fused_backward_kernel dominate)",
                         "flops_per_launch": flops,
                         "peak_source": "measured sustained (MEASURED_PEAKS.json)" if peaks else "fallback"},
            "clocks": sampler.summary(),
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_full_render(args, cfg, rank, world, local_rank):
    """BASELINE configs[2]: the reference `model_v3.Generator.forward` (exp/cips3d/models/model_v3.py:875-1042; called by the
    apps at render_video_web_v10.py:1806-1824) with `use_b200_nerf_branch(G)`: 8 latents per GPU, 64 x 64 rays, N = 128
    samples, the reference decoder up to 1024 x 1024.  A step = one `G.forward` over the rank's 8 images.  value = rays/s of
    the whole pipeline with inputs resident; e2e = latents / cameras from pinned host memory in, the 1024 x 1024 images back to
    pinned host memory.  The same generator with the reference's own renderer is timed beside it on the same GPU
    (`reference_on_gpu`), and its CPU run is the `cpu_baseline` (kind "reference": the live PyTorch reference).
    The decoder's two `op` CUDA extensions are pure-torch stand-ins in every arm (tests/ref_stubs.py)."""
    import copy
    import torch
    import torch.distributed as dist
    import cips3dpp_b200 as c3d
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_stubs
    if not ref_stubs.available():
        if rank == 0:
            print(json.dumps({"metric": "nerf_branch_rays_per_s", "unavailable": "reference modules not found (run oracle/make_ref.sh "
                              "in the build container: oracle/_ref travels with the snapshot)", "config": {"workload": cfg["desc"]}}))
        return
    model_v3, nerf_utils = ref_stubs.import_model_v3()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, N, B = cfg["D"], cfg["N"], cfg["latents"]
    torch.manual_seed(0)
    G_ref = ref_stubs.build_generator(model_v3, D=D, size_end=1024, upsample_list=(128, 256, 512, 1024)).to(dev).eval().requires_grad_(False)
    G = c3d.use_b200_nerf_branch(copy.deepcopy(G_ref), precision=args.precision)
    G.renderer.cache_packed = True
    g = torch.Generator().manual_seed(11 + rank)                        # each rank renders its own latents (sharded by latent)
    host = dict(z0=torch.randn(B, 256, generator=g).pin_memory(), z1=torch.randn(B, 256, generator=g).pin_memory(),
                loc=torch.stack([cfg["azim"] * (2 * torch.rand(B, generator=g) - 1),
                                 cfg["elev"] * (2 * torch.rand(B, generator=g) - 1)], 1).pin_memory())
    torch.manual_seed(1)
    noise = G.create_noise_bufs(start_size=IMG, device=dev)
    nerf_cfg = dict(N_samples=N, perturb=False, static_viewdirs=False)
    out_host = torch.empty(B, 3, 1024, 1024).pin_memory()

    def forward(gen, z0, z1, loc):
        pose, focal, near, far, _ = nerf_utils.Camera.generate_camera_params(img_size=IMG, device=dev, locations=loc,
                                                                             fov_ang=cfg["fov"], dist_radius=cfg["radius"])
        return gen(zs=[z0, z1], cam_poses=pose, focals=focal, img_size=IMG, near=near, far=far, truncation=1, return_xyz=True,
                   noise_bufs=noise, nerf_cfg=nerf_cfg)
    res = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            sync()
            evs = []
            for _ in range(steps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(); e1.record(stream)
                evs.append((e0, e1))
            sync()
        return float(np.mean([a.elapsed_time(b) for a, b in evs]))

    def step():
        return forward(G, res["z0"], res["z1"], res["loc"])

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out_host.copy_(forward(G, d["z0"], d["z1"], d["loc"])["rgb"], non_blocking=True)

    def step_branch():                                                   # the NeRF branch of the same step, alone
        pose, focal, near, far, _ = nerf_utils.Camera.generate_camera_params(img_size=IMG, device=dev, locations=res["loc"],
                                                                             fov_ang=cfg["fov"], dist_radius=cfg["radius"])
        return G.renderer.render(pose, focal, near, far, styles_b, img_size=IMG, N_samples=N)
    styles_b = torch.zeros(B, D + 1, 256, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step = timed(step, args.steps, max(args.warmup, 3))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = G.renderer.last_launch_count
    ms_e2e = timed(step_e2e, args.steps, 2)
    ms_branch = timed(step_branch, args.steps, 2)
    ms_ref_gpu = timed(lambda: forward(G_ref, res["z0"], res["z1"], res["loc"]), max(2, args.steps // 2), 2)
    with torch.no_grad():
        a_, r_ = step(), forward(G_ref, res["z0"], res["z1"], res["loc"])
    rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm())
    parity = {"against": "the same reference Generator with its own VolumeFeatureRenderer, same GPU, same inputs",
              "rgb_1024_rel_l2": rel(a_["rgb"], r_["rgb"]), "thumb_rgb_rel_l2": rel(a_["thumb_rgb"], r_["thumb_rgb"]),
              "xyz_rel_l2": rel(a_["xyz"], r_["xyz"]), "depth_max_abs": float((a_["depth"] - r_["depth"]).abs().max())}
    if world > 1:
        t = torch.tensor([ms_step, ms_e2e, ms_branch], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e, ms_branch = (float(x) for x in t)
    if rank == 0:
        rays = B * IMG * IMG
        line = {
            "metric": "nerf_branch_rays_per_s", "value": world * rays / (ms_step * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": _dtype(args.precision), "data": "synthetic", "images_per_s": world * B / (ms_step * 1e-3),
            "config": {"workload": cfg["desc"], "images_per_gpu": B, "rays_per_image": IMG * IMG, "samples_per_ray": N, "layers": D,
                       "image_size": 1024, "weights": "random-init (reference constructors)",
                       "decoder": "reference modules, `op` CUDA extensions replaced by pure-torch stand-ins in every arm",
                       "l2": "flushed (256 MiB write) between timed steps, outside the event pair"},
            "e2e": {"value": world * rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(sum(v.numel() * 4 for v in host.values())), "d2h_bytes_per_step": int(out_host.numel() * 4)},
            "gpu_launches": int(launches * args.steps),
            "breakdown_ms": {"nerf_branch": ms_branch, "mapping_decoder_and_glue": ms_step - ms_branch},
            "reference_on_gpu": {"ms_per_step": ms_ref_gpu, "images_per_s": B / (ms_ref_gpu * 1e-3),
                                 "what": "the same Generator with the reference's PyTorch VolumeFeatureRenderer (fp32, cuBLAS + ATen)"},
            "parity": parity,
            "roofline": {"bound": "tensor", "kernel": "fused_forward_pair_kernel (NeRF-branch part of the step)", "unit": "TFLOP/s",
                         "achieved": None, "peak": None, "frac": None, "traffic": None},
            "clocks": sampler.summary(),
        }
        from oracle import nerf_oracle as O
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = O.flops_per_point(D) * rays * N / (ms_branch * 1e-3) / 1e12
        line["roofline"].update(achieved=ach, peak=peak_tf, frac=ach / peak_tf,
                                peak_source="measured sustained (MEASURED_PEAKS.json)" if peaks else "fallback")
        if world == 1 and not args.no_cpu_baseline:
            G_cpu = copy.deepcopy(G_ref).cpu()
            noise_cpu = [n.cpu() for n in noise]
            torch.set_num_threads(host_threads())
            t0 = time.perf_counter()
            with torch.no_grad():
                pose, focal, near, far, _ = nerf_utils.Camera.generate_camera_params(img_size=IMG, device="cpu", locations=host["loc"][:1],
                                                                                     fov_ang=cfg["fov"], dist_radius=cfg["radius"])
                G_cpu(zs=[host["z0"][:1], host["z1"][:1]], cam_poses=pose, focals=focal, img_size=IMG, near=near, far=far,
                      truncation=1, return_xyz=True, noise_bufs=noise_cpu, nerf_cfg=nerf_cfg)
            tc = time.perf_counter() - t0
            line["cpu_baseline"] = dict(value=IMG * IMG / tc, unit="rays/s", cores=host_threads(), kind="reference",
                                        sample=f"1 of the step's {B} images through the live PyTorch reference Generator on the host "
                                               f"cores ({tc:.1f} s)", images_per_s=1 / tc)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def side_measurements(m, params, devt, D, N, dev, timed):
    """Two side lines SURVEY.md section 8(d) asks for, at N=1 only, outside the headline timed region:
    (1) the standalone compositing kernel against the HBM roofline on reference-layout inputs
    ((1056 N + 1152) algorithmic bytes per ray); (2) the reference's own formulation of the path as plain PyTorch
    (cuBLAS + ATen, fp32 as the reference runs it, and under bf16 autocast) on the same GPU, on 8 images."""
    import torch
    import cips3dpp_b200 as c3d
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch_ref
    out = {}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    R = 16 * IMG * IMG
    g = torch.Generator(device=dev).manual_seed(3)
    rgb = torch.randn(R, N, 3, device=dev, generator=g)
    sdf = 0.1 * torch.randn(R, N, 1, device=dev, generator=g)
    feat = torch.randn(R, N, 256, device=dev, generator=g)
    z = torch.sort(0.88 + 0.24 * torch.rand(R, N, device=dev, generator=g), dim=-1).values
    rd = torch.randn(R, 3, device=dev, generator=g)
    pts = torch.randn(R, N, 3, device=dev, generator=g)
    beta = torch.tensor([0.1], device=dev)
    ms, ms_min, _ = timed(lambda: c3d.Render.volume_integration(rgb, sdf, feat, z, rd, pts, sigmoid_beta=beta), 10, 3)
    nbytes = (1056 * N + 1152) * R
    hbm = peaks.get("hbm_gbs", 6500.0)
    out["composite_kernel"] = {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "frac": nbytes / (ms * 1e-3) / 1e9 / hbm, "rays": R, "ms": ms,
                               "bytes_per_ray": 1056 * N + 1152, "rays_per_s": R / (ms * 1e-3)}
    del rgb, sdf, feat, z, rd, pts
    # (1b) importance resampling (EXTENSION, off in the headline path): c3d_sample_pdf on one step's rays, weights derived
    # from the sdf, outputs z_fine + merged depths + fine-pass points; algorithmic bytes 8N + 24 + 4K + 16(N+K) per ray
    Rr, K = 256 * IMG * IMG, N
    zc = 0.88 + 0.24 * (torch.arange(N, device=dev)[None] + torch.rand(Rr, 1, device=dev, generator=g)) / N
    sdf_c = (1.0 + 0.1 * torch.rand(Rr, 1, device=dev, generator=g) - zc) * 2.0
    rd, ro = torch.randn(Rr, 3, device=dev, generator=g), torch.randn(Rr, 3, device=dev, generator=g)
    ms, _, _ = timed(lambda: c3d.Render.importance_depths(zc, K, sdf=sdf_c, rays_d=rd, rays_o=ro, sigmoid_beta=beta,
                                                          return_pts=True), 10, 3)
    bpr = 8 * N + 24 + 4 * K + 16 * (N + K)
    out["resample_kernel"] = {"bound": "hbm", "achieved": bpr * Rr / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                              "frac": bpr * Rr / (ms * 1e-3) / 1e9 / hbm, "rays": Rr, "ms": ms, "bytes_per_ray": bpr,
                              "n_samples": N, "n_importance": K, "rays_per_s": Rr / (ms * 1e-3),
                              "kernel": "sample_pdf_lane_kernel"}
    del zc, sdf_c, rd, ro
    nb = min(8, devt[0].shape[0])
    tp = {k: torch.from_numpy(v).to(dev) for k, v in params.items()}
    a = [t[:nb] for t in devt]

    def torch_path():
        with torch.no_grad():
            return torch_ref.render_thumb(tp, a[0], a[1], a[2], a[3], a[4], IMG, N, False)

    def torch_path_bf16():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return torch_path()
    res = {}
    for name, fn in (("fp32", torch_path), ("bf16_autocast", torch_path_bf16)):
        ms, _, _ = timed(fn, 3, 2)
        res[name] = {"images_per_s": nb / (ms * 1e-3), "rays_per_s": nb * IMG * IMG / (ms * 1e-3), "ms": ms}
    out["torch_gpu_baseline"] = {"what": "reference formulation as plain PyTorch (cuBLAS + ATen) on this GPU",
                                 "images": nb, **res}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the compositing-kernel and torch-on-GPU side lines")
    ap.add_argument("--shared-latent", action="store_true", help="c5, N > 1: shard the 16 targets over the ranks (strong scaling), "
                    "one latent shared by all, gradient all-reduce every step")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: leave the all-gather of the rendered maps out of the timed region")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if args.config == "c5":
        run_inversion(args, cfg, rank, world, local_rank)
        return
    if args.config == "c3full":
        run_full_render(args, cfg, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    import cips3dpp_b200 as c3d
    from oracle import nerf_oracle as O

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c3d._abi.load()

    D, N = cfg["D"], cfg["N"]
    params = O.init_params(D, seed=0)
    m = c3d.NerfBranch(D, precision=args.precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    m = m.to(dev).eval().requires_grad_(False)
    m.cache_packed = True                        # serving: frozen weights are packed once (the default re-packs on every call)
    c2w, focal, near, far, styles = workload(cfg, seed_latent=1 + rank, seed_pose=2 + rank)
    B = c2w.shape[0]
    host = [torch.from_numpy(x).pin_memory() for x in (c2w, focal, near, far, styles)]
    devt = [h.to(dev) for h in host]
    m.packed_weights()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_resident():
        return m.render(devt[0], devt[1], devt[2], devt[3], devt[4], img_size=IMG, N_samples=N)

    out_host = {k: torch.empty(s, dtype=torch.float32).pin_memory()
                for k, s in (("rgb_map", (B, IMG * IMG, 3)), ("mask", (B, IMG * IMG, 2)), ("xyz", (B, IMG * IMG, 3)))}

    def step_e2e():
        d = [h.to(dev, non_blocking=True) for h in host]
        out = m.render(d[0], d[1], d[2], d[3], d[4], img_size=IMG, N_samples=N)
        for k, hbuf in out_host.items():
            hbuf.copy_(out[k], non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        evs = []
        barrier()
        t_wall = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)                                            # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        wall = time.perf_counter() - t_wall
        ms = [a.elapsed_time(b) for a, b in evs]
        return float(np.mean(ms)), float(np.min(ms)), wall

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step, ms_min, wall = timed(step_resident, args.steps, max(args.warmup, 3))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches_per_step = m.last_launch_count
    ms_e2e_serial, _, _ = timed(step_e2e, args.steps, 2)

    # End-to-end throughput as a serving loop runs it: the step's results go device -> host on a copy stream while the
    # next step (its own host -> device upload included) already runs.  One event pair around all K steps; every copy of
    # every step is inside it (the region ends only when the last download has finished).
    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_loop(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            flush.fill_(1)
            d = [h.to(dev, non_blocking=True) for h in host]
            out = m.render(d[0], d[1], d[2], d[3], d[4], img_size=IMG, N_samples=N)
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                for k, hbuf in out_host.items():
                    out[k].record_stream(copy_stream)
                    hbuf.copy_(out[k], non_blocking=True)
        stream.wait_stream(copy_stream)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) / steps
    e2e_loop(2)
    ms_e2e = e2e_loop(args.steps)

    # the dominant kernel alone: step minus the (tiny) FiLM style_prep launch, measured live
    film = torch.empty(B, D + 1, 256, 2, device=dev)
    first = torch.empty(B, 256, 4, device=dev)
    view = torch.empty(B, 256, 4, device=dev)
    lib = c3d._abi.load()

    def sp():
        c3d._abi.check(lib.c3d_style_prep(m.packed_weights().data_ptr(), D, devt[4].data_ptr(), B, film.data_ptr(),
                                          first.data_ptr(), view.data_ptr(), stream.cuda_stream), "c3d_style_prep")
    ms_sp, _, _ = timed(sp, args.steps, 2)

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # N > 1: the collective the north star names for this path -- every rank ends up with every rank's rendered maps
    # (all_gather_into_tensor over NVLink) -- inside the timed region: bf16 channel-major feature maps (written in that
    # form by the kernel) + fp32 rgb / mask / xyz, on a communication stream under the NEXT step's render.
    comm = None
    if world > 1 and not args.no_gather:
        comm_stream = torch.cuda.Stream(device=dev)
        full_feat = [torch.empty(world * B, 256, IMG * IMG, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        full_small = [torch.empty(world * B, IMG * IMG, 8, device=dev) for _ in range(2)]

        def render_and_gather(i, overlap=True):
            out = m.render(devt[0], devt[1], devt[2], devt[3], devt[4], img_size=IMG, N_samples=N, features_nchw="bf16")
            small = torch.cat([out["rgb_map"], out["mask"], out["xyz"]], -1)
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(done)
                out["feature_map"].record_stream(comm_stream)
                small.record_stream(comm_stream)
                dist.all_gather_into_tensor(full_feat[i & 1], out["feature_map"])
                dist.all_gather_into_tensor(full_small[i & 1], small)
            if not overlap:
                stream.wait_stream(comm_stream)

        def comm_loop(steps, overlap=True):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(stream)
            for i in range(steps):
                flush.fill_(1)
                render_and_gather(i, overlap)
            stream.wait_stream(comm_stream)
            e1.record(stream)
            barrier()
            return e0.elapsed_time(e1) / steps
        comm_loop(3)
        ms_comm = maxr(comm_loop(args.steps))
        ms_serial = maxr(comm_loop(args.steps, overlap=False))
        del full_feat, full_small
        # Without a collective kernel: dist.GatheredMaps (symmetric memory: every rank's gathered tensors mapped into every
        # process over NVLink / NVSwitch).  (a) "p2p": after the render, peer-to-peer copies of the rank's shard into every
        # rank's tensors on a communication stream -- DMA engines, no SMs, so they run under the next step's persistent
        # kernel; one cross-rank barrier per step.  (b) "fused": the kernel's compositing epilogue stores the maps into every
        # rank's tensors itself.  Two buffer sets alternate (a consumer may still read the previous step's maps).
        ms_fused = ms_p2p = None
        sym_err = None
        try:
            gms = [c3d.dist.GatheredMaps(B, IMG * IMG, features="bf16") for _ in range(2)]

            def sym_loop(steps, fused):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e0.record(stream)
                for i in range(steps):
                    flush.fill_(1)
                    gm = gms[i & 1]
                    if fused:
                        m.render(devt[0], devt[1], devt[2], devt[3], devt[4], img_size=IMG, N_samples=N, gather=gm)
                        gm.barrier()
                    else:
                        out = m.render(devt[0], devt[1], devt[2], devt[3], devt[4], img_size=IMG, N_samples=N, features_nchw="bf16")
                        done = torch.cuda.Event()
                        done.record(stream)
                        comm_stream.wait_event(done)
                        for k in ("feature_map", "rgb_map", "mask", "xyz"):
                            out[k].record_stream(comm_stream)
                        gm.push(out, comm_stream)
                        with torch.cuda.stream(comm_stream):
                            gm.barrier()
                stream.wait_stream(comm_stream)
                e1.record(stream)
                barrier()
                return e0.elapsed_time(e1) / steps
            sym_loop(3, False)
            ms_p2p = maxr(sym_loop(args.steps, False))
            sym_loop(3, True)
            ms_fused = maxr(sym_loop(args.steps, True))
        except Exception as ex:  # noqa: BLE001  (no symmetric memory on this box: the NCCL path is the fallback measurement)
            sym_err = f"{type(ex).__name__}: {ex}"[:200]
        gbytes = (world * B) * (256 * IMG * IMG * 2 + IMG * IMG * 8 * 4)
        base = maxr(ms_step)
        comm = {"what": "every rank ends up with every rank's maps: bf16 (b,256,hw) feature maps + fp32 rgb/mask/xyz, every step",
                "gathered_bytes_per_rank_per_step": int(gbytes), "ms_per_step_no_comm": base,
                "p2p": {"how": "peer-to-peer copies (DMA engines) of the rank's shard into every rank's symmetric-memory tensors on a "
                               "communication stream under the next step's render + one cross-rank barrier per step",
                        "ms_per_step": ms_p2p, "exposed_ms": None if ms_p2p is None else ms_p2p - base, "error": sym_err},
                "fused": {"how": "the kernel's epilogue stores into every rank's tensors (peer memory) + one cross-rank barrier per step",
                          "ms_per_step": ms_fused, "exposed_ms": None if ms_fused is None else ms_fused - base},
                "nccl": {"how": "all_gather_into_tensor on a communication stream under the next step's render",
                         "ms_per_step": ms_comm, "exposed_ms": ms_comm - base, "ms_per_step_not_overlapped": ms_serial,
                         "gather_ms_alone": ms_serial - base, "algbw_GBps_alone": gbytes / max(ms_serial - base, 1e-6) / 1e6},
                "headline_uses": min((v, k) for k, v in (("p2p", ms_p2p), ("fused", ms_fused), ("nccl", ms_comm)) if v is not None)[1]}

    ms_step, ms_e2e, ms_sp_m, ms_e2e_serial = maxr(ms_step), maxr(ms_e2e), maxr(ms_sp), maxr(ms_e2e_serial)
    ms_kernel_step = ms_step
    if comm is not None:                                      # the headline at N > 1 includes the gather of the maps
        ms_step = comm[comm["headline_uses"]]["ms_per_step"]
    extras = {}
    if world == 1 and not args.no_extras:
        extras = side_measurements(m, params, devt, D, N, dev, timed)
    rays_step = B * IMG * IMG
    value = world * rays_step / (ms_step * 1e-3)
    e2e_value = world * rays_step / (ms_e2e * 1e-3)

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured sustained (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md sustained)"
    flops_launch = O.flops_per_point(D) * rays_step * N
    ms_kernel = max(ms_kernel_step - ms_sp_m, 1e-6)
    achieved = flops_launch / (ms_kernel * 1e-3) / 1e12
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf)).get(f"{args.config}_{args.precision}")

    if rank == 0:
        line = {
            "metric": "nerf_branch_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": _dtype(args.precision), "data": "synthetic",
            "images_per_s": world * B / (ms_step * 1e-3), "ms_per_step_min": ms_min,
            "config": {"workload": cfg["desc"], "images_per_gpu": B, "rays_per_image": IMG * IMG, "samples_per_ray": N,
                       "layers": D, "weights": "random-init (reference distributions)", "sampling": "eval (unperturbed)",
                       "l2": "flushed (256 MiB write) between timed steps, outside the event pair",
                       "d2h": "rgb_map+mask+xyz to pinned host (feature_map stays on device for the decoder)",
                       "arithmetic": OPERANDS if args.precision == "bf16" else
                       "fp32 mode: two-way fp16 split operands, three tcgen05 products per layer, FiLM / sines in fp32"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": ms_e2e, "ms_per_step_unpipelined": ms_e2e_serial,
                    "how": "pinned host -> device upload, render, device -> host download of rgb_map+mask+xyz every step; "
                           "downloads overlap the next step on a copy stream; one event pair around all steps "
                           "(L2 flush fill included)",
                    "h2d_bytes_per_step": int(sum(h.numel() * 4 for h in host)),
                    "d2h_bytes_per_step": int(sum(h.numel() * 4 for h in out_host.values()))},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": traffic,
                         "kernel": "fused_forward_pair_kernel + its two per-image weight-image kernels (step minus style_prep)"
                         if args.precision == "bf16" else
                         "raygen_kernel + mlp_tc32_kernel + composite_fwd_kernel per image chunk (algorithmic FLOPs; the tensor "
                         "cores execute three fp16 products per hidden layer, i.e. ~3x these)",
                         "peak_source": peak_src, "kernel_ms": ms_kernel, "flops_per_launch": flops_launch,
                         "frac_of_nominal_2250": achieved / 2250.0,
                         "frac_of_burst": achieved / peaks.get("bf16_tflops", 1590.0)},
            "clocks": sampler.summary(),
        }
        line.update(extras)
        if comm is not None:
            line["comm"] = comm
            line["config"]["multi_gpu"] = "images sharded by (latent, pose); value includes the per-step all-gather of the maps (see comm)"
        if world == 1 and not args.no_cpu_baseline:
            # the CPU arm renders the first images of this very step: its outputs pin the GPU maps of the timed workload
            ref = {}
            n_chk = min(B, 8)
            cb, _ = cpu_baseline(cfg, n_chk, reps=2, seeds=(1 + rank, 2 + rank), keep=ref)
            line["cpu_baseline"] = cb
            with torch.no_grad():
                got = step_resident()
            torch.cuda.synchronize()
            rel = lambda a_, b_: float(np.linalg.norm(a_.astype(np.float64) - b_) / max(np.linalg.norm(b_.astype(np.float64)), 1e-30))
            g = {k: got[k][:n_chk].cpu().numpy() for k in ("rgb_map", "feature_map", "sdf", "mask", "xyz", "z_vals")}
            line["parity"] = {
                "against": f"oracle/nerf_oracle.c on the first {n_chk} images of the timed step (full 64x64, all rays)",
                "feature_map_rel_l2": rel(g["feature_map"], ref["feature_map"]), "rgb_map_rel_l2": rel(g["rgb_map"], ref["rgb_map"]),
                "xyz_rel_l2": rel(g["xyz"], ref["xyz"]), "sdf_rel_l2": rel(g["sdf"], ref["sdf"]),
                "sample_depth_max_abs": float(np.abs(g["z_vals"] - ref["z_vals"]).max()),
                "depth_map_max_abs": float(np.abs(g["mask"][..., 1] - ref["mask"][..., 1]).max()),
                "tolerance": {"bf16": "2e-2 rel-L2", "fp32": "1e-3 rel-L2, 1e-4 depth"}[args.precision],
            }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
