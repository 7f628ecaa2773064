"""Batched flip GAN inversion through the NeRF branch -- the caller of the forward+backward hot path
(reference: StyleGAN2Projector_Flip.project_wplus stage 1, exp/cips3d/models/projector_v9.py:862-1166, where only
the cameras and w_render move; `_G_forward` :191-257; lr ramp `_get_cur_lr` :154-166).

What is mirrored: per target one W+ latent `w_render (D+1, 256)` shared by the image and its horizontal flip
(`w.repeat(2,1,1)`, :1050), one (azim, elev) per view, Adam(betas=(0.9, 0.999)), cosine ramp-down / linear ramp-up
learning rate, gradient-norm clipping at 10, loss on the 64x64 thumbnail.  What is not: the 2-D decoder and the
VGG perceptual network stay the reference's (out of scope) -- pass `loss_fn` to plug them in.
Targets are independent, so ranks take disjoint targets and no gradient all-reduce is needed unless
`shared_latent=True` (one latent fitted to all targets of all ranks), which uses `dist.allreduce_grads`.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _abi
from . import dist as c3d_dist
from .nerf_utils import Camera


class _FusedClipAdam:
    """Gradient-norm clipping + Adam for the step's few small tensors in ONE launch (c3d_adam_clip_step): group 0 = the
    latents, group 1 = the cameras, each clipped on its own norm and stepped with its own learning rate (projector_v9.py:
    1016-1030, 1150-1153 uses two torch optimisers and two clip_grad_norm_ calls: ~40 launches per step).  Learning rates
    and the step counter live on the device, so the update is CUDA-graph capturable."""

    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-8, max_norm=10.0):
        self.tensors = [(t, g) for g, ts in enumerate(groups) for t in ts]
        dev = self.tensors[0][0].device
        self.m = [torch.zeros_like(t) for t, _ in self.tensors]
        self.v = [torch.zeros_like(t) for t, _ in self.tensors]
        self.lr = torch.zeros(2, device=dev)
        self.step_count = torch.zeros((), device=dev)
        self.betas, self.eps, self.max_norm = betas, eps, max_norm

    def step(self):
        lib = _abi.load()
        P = _abi.AdamParams()
        P.n_tensors = len(self.tensors)
        keep = []
        for i, (t, g) in enumerate(self.tensors):
            if t.grad is None or not t.is_contiguous():
                raise RuntimeError("fused update needs contiguous parameters with gradients")
            grad = t.grad if t.grad.is_contiguous() else t.grad.contiguous()
            keep.append(grad)
            P.group[i], P.numel[i] = g, t.numel()
            P.param[i], P.grad[i], P.exp_avg[i], P.exp_avg_sq[i] = t.data_ptr(), grad.data_ptr(), self.m[i].data_ptr(), \
                self.v[i].data_ptr()
        P.lr[0], P.lr[1] = self.lr.data_ptr(), self.lr.data_ptr() + 4
        P.step = self.step_count.data_ptr()
        P.beta1, P.beta2, P.eps, P.max_norm = self.betas[0], self.betas[1], self.eps, self.max_norm
        dev = self.lr.device
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_adam_clip_step(P, torch.cuda.current_stream().cuda_stream), "c3d_adam_clip_step")

    def reset(self):
        for t in self.m + self.v:
            t.zero_()
        self.step_count.zero_()


def lr_ramp(step, num_steps, lr0, rampdown=0.25, rampup=0.05):
    """projector_v9.py:154-166."""
    t = step / num_steps
    r = min(1.0, (1.0 - t) / rampdown)
    r = 0.5 - 0.5 * math.cos(r * math.pi)
    return lr0 * r * min(1.0, t / rampup)


def thumb_mse(thumb, target):
    return ((thumb - target) ** 2).mean(dim=(1, 2, 3)).sum()


class FlipInversion:
    def __init__(self, renderer, img_size=64, N_samples=24, cam_cfg=None, lr_latent=0.02, lr_cam=0.01, num_steps=200,
                 loss_fn=thumb_mse, clip=10.0, shared_latent=False, static_viewdirs=True, fused_update=True,
                 loss_on_features=False):
        self.renderer, self.img_size, self.N = renderer, img_size, N_samples
        self.cam_cfg = dict(fov_ang=6, dist_radius=0.12) if cam_cfg is None else dict(cam_cfg)
        self.lr_latent, self.lr_cam, self.num_steps = lr_latent, lr_cam, num_steps
        self.loss_fn, self.clip, self.shared_latent, self.static_viewdirs = loss_fn, clip, shared_latent, static_viewdirs
        # the reference's loss sees the thumb AND the decoder's image, i.e. the 256-channel feature map (projector_v9.py:230-246,
        # renderer_detach=False): with loss_on_features the loss is called as loss_fn(thumbs, targets, features) with features
        # (n*2, 256, S, S) -- the decoder / VGG stay the caller's -- and its cotangent flows back through feature_map
        self.loss_on_features = loss_on_features
        # clipping + both Adam updates as one kernel (CUDA tensors); False (or C3D_INV_FUSED=0, for A/B runs): torch.optim
        self.fused_update = fused_update and os.environ.get("C3D_INV_FUSED", "1") != "0"
        # stage 1 optimises latents and cameras only: with a frozen renderer the packed weights are built once
        if hasattr(renderer, "cache_packed") and not any(p.requires_grad for p in renderer.parameters()):
            renderer.cache_packed = True

    def render_maps(self, w, azim, elev, features=False):
        """w (n, D+1, 256); azim, elev (n, 2, 1) -> thumbs (n*2, 3, S, S) [and features (n*2, 256, S, S)], differentiable."""
        n, S = w.shape[0], self.img_size
        loc = torch.cat([azim.reshape(-1, 1), elev.reshape(-1, 1)], 1)
        pose, focal, near, far, _ = Camera.generate_camera_params(S, w.device, locations=loc, **self.cam_cfg)
        styles = w.repeat_interleave(2, dim=0)                        # image and flip share the latent
        kw = dict(features_nchw=True) if features else {}
        out = self.renderer.render(pose, focal, near, far, styles, img_size=S, N_samples=self.N,
                                   static_viewdirs=self.static_viewdirs, **kw)
        thumbs = out["rgb_map"].reshape(n * 2, S, S, 3).permute(0, 3, 1, 2)
        return (thumbs, out["feature_map"].reshape(n * 2, -1, S, S)) if features else thumbs

    def render_thumbs(self, w, azim, elev):
        return self.render_maps(w, azim, elev)

    def run(self, targets, w_init, azim_init=None, elev_init=None, callback=None, cuda_graph=False, host_targets=None):
        """targets (n, 3, S, S) in [-1, 1]; w_init (1 or n, D+1, 256).  Returns dict(w, azim, elev, losses, events); `events` is a
        pair of CUDA events around the optimisation loop (after a synchronize, `events[0].elapsed_time(events[1]) /
        num_steps` is the device time per step).

        `cuda_graph=True` captures one whole optimisation step (camera glue, forward, loss, backward, clipping,
        both Adam updates) into a CUDA graph and replays it `num_steps` times; only the two learning rates are
        written from the host between replays.  It needs frozen renderer weights and a capturable `loss_fn`; with a latent
        shared across ranks the NCCL all-reduce of its gradient is captured in the graph too.

        `host_targets` (pinned host tensor shaped like `targets`) makes every step end to end: the step's targets are
        copied host -> device before it and its loss is copied back to (pinned) host memory after it -- asynchronously on
        the same stream, so the loop still never waits for the host (what bench.py's `e2e` times)."""
        dev, n = targets.device, targets.shape[0]
        tgt = torch.stack([targets, targets.flip(-1)], 1).reshape(n * 2, *targets.shape[1:])
        nw = 1 if self.shared_latent else n
        w = w_init.detach().expand(nw, -1, -1).clone().requires_grad_(True)
        azim = (torch.zeros(n, 2, 1, device=dev) if azim_init is None else azim_init.detach().clone()).requires_grad_(True)
        elev = (torch.zeros(n, 2, 1, device=dev) if elev_init is None else elev_init.detach().clone()).requires_grad_(True)
        multi_rank = self.shared_latent and torch.distributed.is_available() and torch.distributed.is_initialized()
        # with a latent shared across ranks the gradient all-reduce (NCCL) is part of the step; it is captured into the CUDA
        # graph like every other launch of the step (every rank captures and replays the same sequence)
        fused = self.fused_update and dev.type == "cuda"
        device_lr = fused or cuda_graph                              # learning rates live on the device
        if fused:
            upd = _FusedClipAdam([[w], [azim, elev]], betas=(0.9, 0.999), max_norm=self.clip)
            lr_dev = upd.lr
        elif cuda_graph:
            lr_dev = torch.zeros(2, device=dev)
            opt_w = torch.optim.Adam([w], betas=(0.9, 0.999), lr=lr_dev[0], capturable=True)
            opt_c = torch.optim.Adam([azim, elev], betas=(0.9, 0.999), lr=lr_dev[1], capturable=True)
        else:
            opt_w = torch.optim.Adam([w], betas=(0.9, 0.999), lr=self.lr_latent)
            opt_c = torch.optim.Adam([azim, elev], betas=(0.9, 0.999), lr=self.lr_cam)
        if device_lr:
            lrs = torch.tensor([[lr_ramp(s, self.num_steps, l0) for l0 in (self.lr_latent, self.lr_cam)]
                                for s in range(self.num_steps)], dtype=torch.float32).to(dev)

        def one_step():
            wn = w.expand(n, -1, -1) if self.shared_latent else w
            if self.loss_on_features:
                thumbs, feats = self.render_maps(wn, azim, elev, features=True)
                loss = self.loss_fn(thumbs, tgt, feats)
            else:
                loss = self.loss_fn(self.render_maps(wn, azim, elev), tgt)
            w.grad = azim.grad = elev.grad = None
            loss.backward()
            if multi_rank:
                c3d_dist.allreduce_grads([w.grad])
            if fused:
                upd.step()                                           # both clippings + both Adam updates: one launch
            else:
                torch.nn.utils.clip_grad_norm_([w], self.clip)
                torch.nn.utils.clip_grad_norm_([azim, elev], self.clip)
                opt_w.step()
                opt_c.step()
            return loss.detach()

        losses = []
        if cuda_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                            # warm-up at lr = 0: parameters do not move
                for _ in range(2):
                    one_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            if fused:                                                # forget the warm-up in the Adam moments
                upd.reset()
            else:
                for opt in (opt_w, opt_c):
                    for st in opt.state.values():
                        for v in st.values():
                            if torch.is_tensor(v):
                                v.zero_()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss_static = one_step()
        if host_targets is not None:
            losses_host = torch.empty(self.num_steps, dtype=torch.float32).pin_memory()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        for step in range(self.num_steps):
            if device_lr:
                lr_dev.copy_(lrs[step])
            else:
                for opt, lr0 in ((opt_w, self.lr_latent), (opt_c, self.lr_cam)):
                    for g in opt.param_groups:
                        g["lr"] = lr_ramp(step, self.num_steps, lr0)
            if host_targets is not None:
                t = host_targets.to(dev, non_blocking=True)
                tgt.copy_(torch.stack([t, t.flip(-1)], 1).reshape(tgt.shape))
            if cuda_graph:
                graph.replay()
                losses.append(loss_static.clone())
            else:
                losses.append(one_step())
            if host_targets is not None:                             # device -> host read of the step's result
                losses_host[step:step + 1].copy_(losses[-1].reshape(1), non_blocking=True)
            if callback is not None:
                callback(step, losses[-1])
        ev[1].record()
        return dict(w=w.detach(), azim=azim.detach(), elev=elev.detach(), losses=torch.stack(losses), events=ev)
