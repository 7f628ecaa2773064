"""Multi-GPU plumbing: the path shards by (latent, pose) image with no data-path collective
(gen_images.py:57-91 does the same in the reference).  One process per GPU; NCCL is used only to
all-gather rendered maps and to all-reduce latent/camera gradients in batched inversion."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous split of n_items across ranks; earlier ranks take the remainder."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_maps(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather per-rank image shards (dim 0) back into the full (n_items, ...) tensor."""
    ws = dist.get_world_size(group)
    if ws == 1:
        return local
    counts = [shard_range(n_items, r, ws) for r in range(ws)]
    mx = max(e - s for s, e in counts)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = local.new_empty((ws * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + (e - s)] for r, (s, e) in enumerate(counts)], 0)


def allreduce_grads(tensors, group=None):
    """One flattened SUM all-reduce for a list of (small) gradient tensors, in place."""
    if dist.get_world_size(group) == 1 or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view_as(t))
        o += n
    return tensors


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class GatheredMaps:
    """Peer-mapped (symmetric-memory) destination of the all-gather of rendered maps, filled over NVLink / NVSwitch without a
    collective kernel.  Two ways to fill it:
      * `push(maps, stream)`: peer-to-peer copies of this rank's shard into every rank's tensors (DMA engines; the way bench.py
        gathers -- it overlaps with the next step's render, which leaves no SM for an NCCL kernel);
      * `NerfBranch.render(..., gather=gm)`: FUSED -- the kernel's compositing epilogue stores every finished ray straight
        into the tensors of every rank.  Bit-exact and launch-free, but the channel-major layout makes those stores 2-byte
        scattered writes, which NVLink serves badly (measured 1.5x the render time at 2 GPUs): kept for the (b, hw, 256)
        layout and small batches, not the default.

    One process per GPU, equal shards: rank r's images land at [r * batch_per_rank, (r + 1) * batch_per_rank).  After the
    launch, `barrier()` (cross-rank, on the current stream) makes every rank's stores visible before any rank reads; use two
    instances alternately when the next step should render while the previous result is still being consumed.
    Tensors (the full batch of all ranks): feature_map (B, 256, hw) bfloat16 [features='bf16'], float32 [features='nchw'] or
    (B, hw, 256) float32 [features='nhwc']; rgb_map (B, hw, 3); mask (B, hw, 2); xyz (B, hw, 3)."""

    def __init__(self, batch_per_rank: int, n_rays: int, features: str = "bf16", group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _abi
        group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _abi.MAX_PEERS:
            raise ValueError(f"at most {_abi.MAX_PEERS} peers")
        self.batch_per_rank, self.n_rays, self.features = batch_per_rank, n_rays, features
        B = self.world * batch_per_rank
        fdt = torch.bfloat16 if features == "bf16" else torch.float32
        fshape = (B, n_rays, 256) if features == "nhwc" else (B, 256, n_rays)
        spec = [("feature_map", fshape, fdt), ("rgb_map", (B, n_rays, 3), torch.float32), ("mask", (B, n_rays, 2), torch.float32),
                ("xyz", (B, n_rays, 3), torch.float32)]
        offs, total = {}, 0
        for name, shape, dt in spec:
            offs[name] = total
            n = 1
            for d in shape:
                n *= d
            total += (n * torch.empty((), dtype=dt).element_size() + 255) // 256 * 256
        dev = torch.device("cuda", torch.cuda.current_device())
        self._arena = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        self._hdl = symm_mem.rendezvous(self._arena, group)
        base = [int(p) for p in self._hdl.buffer_ptrs]
        self._ptrs = {name: [b + offs[name] for b in base] for name, _, _ in spec}
        self._elem_off = {}
        for name, shape, dt in spec:
            n = 1
            for d in shape:
                n *= d
            esz = torch.empty((), dtype=dt).element_size()
            self._elem_off[name] = offs[name] // esz                 # storage offset of the tensor in elements of its dtype
            setattr(self, name, self._arena[offs[name]: offs[name] + n * esz].view(dt).view(shape))
        self.layout = {"nhwc": _abi.FEAT_NHWC, "nchw": _abi.FEAT_NCHW, "bf16": _abi.FEAT_NCHW_BF16}[features]

    def push(self, maps, stream=None):
        """Copy-engine all-gather: this rank's freshly rendered maps (dict with feature_map / rgb_map / mask / xyz, this rank's
        images only) are copied into this rank's slot of EVERY rank's gathered tensors with peer-to-peer copies over NVLink
        (cudaMemcpyAsync on mapped peer memory: DMA engines, no SMs -- so it runs under the next step's persistent render
        kernel, which occupies every SM).  Enqueued on `stream` (default: current), which must already wait for the render."""
        if not hasattr(self, "_peer_views"):
            self._peer_views = []
            for r in range(self.world):
                views = {}
                for name in ("feature_map", "rgb_map", "mask", "xyz"):
                    t = getattr(self, name)
                    views[name] = self._hdl.get_buffer(r, t.shape, t.dtype, self._elem_off[name])
                self._peer_views.append(views)
        s0 = self.rank * self.batch_per_rank
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        with ctx:
            for k in range(self.world):
                r = (self.rank + k) % self.world                      # start with the own copy, then ring order: spreads the links
                for name in ("feature_map", "rgb_map", "mask", "xyz"):
                    self._peer_views[r][name][s0: s0 + self.batch_per_rank].copy_(maps[name], non_blocking=True)

    def struct(self):
        from . import _abi
        g = _abi.GatherOut()
        g.n_peers, g.image_offset = self.world, self.rank * self.batch_per_rank
        for name in ("feature_map", "rgb_map", "mask", "xyz"):
            arr = getattr(g, name)
            for i, p in enumerate(self._ptrs[name]):
                arr[i] = p
        return g

    def barrier(self):
        """Cross-rank barrier on the current CUDA stream: all ranks' stores into this rank's tensors are complete after it."""
        self._hdl.barrier()
