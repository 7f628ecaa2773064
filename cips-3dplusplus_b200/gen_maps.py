"""Multi-GPU generation of NeRF-branch maps for FID-scale runs -- the NeRF half of the reference's `gen_images`
(exp/cips3d/scripts/gen_images.py:33-91): one process per GPU, every rank draws its own latents and cameras, renders
`batch_gpu` images per step and writes them under the reference's interleaved numbering
`idx_b * batch_gpu * world + idx_i * world + rank` (:83).  The 2-D decoder stays the reference's: pass it as `decoder`
(a callable `(feature_map_nchw, maps) -> images`) to get final images, otherwise the 64x64 maps are saved.

    torchrun --nproc-per-node 8 -m cips3dpp_b200.gen_maps --weights renderer.pt --num-imgs 50000 --out fake/

No data-path collective: ranks only meet at the final barrier.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from .nerf_utils import Camera


def image_index(idx_b, idx_i, batch_gpu, world_size, rank):
    """gen_images.py:83."""
    return idx_b * batch_gpu * world_size + idx_i * world_size + rank


def gen_maps(renderer, style_fn, cam_cfg, num_imgs, batch_gpu, out_dir=None, decoder=None, rank=0, world_size=1,
             img_size=64, N_samples=24, static_viewdirs=False, seed=0, device=None, keep=False, N_importance=0):
    """Render `num_imgs` images in total over `world_size` ranks.

    style_fn(batch, generator) -> styles (batch, D+1, 256) for the NeRF branch (the reference maps z through
    `Generator.style`, model_v3.py:1420-1433; any callable works).  cam_cfg: keyword arguments of
    `Camera.generate_camera_params` (fov_ang, dist_radius, azim_range, elev_range, uniform, ...).
    `N_importance > 0` renders with the optional coarse + fine pass (`NerfBranch.render_hierarchical`, an extension: the
    reference renders in one pass).  Returns the list of (index, path or arrays) this rank produced."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + rank)
    if out_dir is not None and rank == 0:
        os.makedirs(out_dir, exist_ok=True)
    if world_size > 1 and torch.distributed.is_initialized():
        torch.distributed.barrier()
    batch_size = batch_gpu * world_size
    produced = []
    with torch.no_grad():
        for idx_b in range((num_imgs + batch_size - 1) // batch_size):
            styles = style_fn(batch_gpu, g)
            torch.manual_seed(seed * 7919 + idx_b * world_size + rank)          # camera draws (torch.randn / rand inside)
            pose, focal, near, far, _ = Camera.generate_camera_params(img_size, device, batch=batch_gpu, **cam_cfg)
            kw = dict(img_size=img_size, N_samples=N_samples, static_viewdirs=static_viewdirs, features_nchw=decoder is not None)
            if N_importance > 0:
                maps = renderer.render_hierarchical(pose, focal, near, far, styles, N_importance=N_importance, **kw)
            else:
                maps = renderer.render(pose, focal, near, far, styles, **kw)
            images = None
            if decoder is not None:
                images = decoder(maps["feature_map"].view(batch_gpu, -1, img_size, img_size), maps)
            for idx_i in range(batch_gpu):
                idx = image_index(idx_b, idx_i, batch_gpu, world_size, rank)
                if idx >= num_imgs:
                    continue
                if images is not None:
                    rec = dict(image=images[idx_i].float().cpu().numpy())
                else:
                    thumb = maps["rgb_map"][idx_i].reshape(img_size, img_size, 3)
                    rec = dict(thumb=((thumb.clamp(-1, 1) + 1) * 127.5).round().to(torch.uint8).cpu().numpy(),
                               depth=maps["mask"][idx_i, :, 1].reshape(img_size, img_size).cpu().numpy())
                if out_dir is not None:
                    path = os.path.join(out_dir, f"{idx:0>5}.npz")
                    np.savez_compressed(path, **rec)
                    produced.append((idx, path))
                else:
                    produced.append((idx, rec if keep else None))
    if world_size > 1 and torch.distributed.is_initialized():
        torch.distributed.barrier()
    return produced


def gaussian_styles(D, mean=-0.02, std=0.62):
    """W+ latents with the statistics of the shipped `datasets/cars/style_render.pkl`: one latent per
    image, repeated over the D+1 FiLM layers (model_v3.py:1416)."""
    def fn(batch, generator):
        w = mean + std * torch.randn(batch, 1, 256, device=generator.device, generator=generator)
        return w.repeat(1, D + 1, 1)
    return fn


def main():
    from . import NerfBranch
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--weights", required=True, help="torch-saved state dict of the renderer (reference names)")
    ap.add_argument("--layers", type=int, required=True, help="N_layers_renderer of the checkpoint")
    ap.add_argument("--num-imgs", type=int, default=50000)
    ap.add_argument("--batch-gpu", type=int, default=64)
    ap.add_argument("--out", required=True)
    ap.add_argument("--n-samples", type=int, default=24)
    ap.add_argument("--n-importance", type=int, default=0, help="> 0: coarse + fine render (extension, off by default)")
    ap.add_argument("--fov", type=float, default=6.0)
    ap.add_argument("--dist-radius", type=float, default=0.12)
    ap.add_argument("--azim-range", type=float, default=0.3)
    ap.add_argument("--elev-range", type=float, default=0.15)
    ap.add_argument("--uniform", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    m = NerfBranch(a.layers)
    m.load_state_dict(torch.load(a.weights, map_location="cpu"), strict=True)
    m = m.cuda().eval().requires_grad_(False)
    m.cache_packed = True                            # frozen checkpoint: pack the weights once
    cam = dict(fov_ang=a.fov, dist_radius=a.dist_radius, azim_range=a.azim_range, elev_range=a.elev_range, uniform=a.uniform)
    out = gen_maps(m, gaussian_styles(a.layers), cam, a.num_imgs, a.batch_gpu, out_dir=a.out, rank=rank, world_size=world,
                   N_samples=a.n_samples, seed=a.seed, N_importance=a.n_importance)
    if rank == 0:
        print(f"rank 0 wrote {len(out)} of {a.num_imgs} images to {a.out}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
