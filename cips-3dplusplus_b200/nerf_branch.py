"""NerfBranch: drop-in replacement of the reference `VolumeFeatureRenderer`
(exp/cips3d/volume_renderer.py:163-303) whose forward runs in libc3dpp.so on a B200.

Same constructor arguments, same parameter names (`sigmoid_beta`, `network.pts_linears.{i}.*`,
`network.views_linears.*`, `network.rgb_linear.*`, `network.sigma_linear.*` -- reference checkpoints load
with strict=True), same forward signature and return tuple.  `Generator.renderer = NerfBranch.from_reference(
G.renderer)` is the whole integration.

There is no PyTorch fallback: CPU tensors or a missing library raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _abi

W = 256


# ---------------------------------------------------------------------------------------------
# parameter containers with the reference names and init distributions
# ---------------------------------------------------------------------------------------------
class _Linear(nn.Module):
    """LinearLayer parameters (volume_renderer.py:15-30). `kind`: 'style' (0.25*kaiming, a=0.2) or 'freq'."""

    def __init__(self, in_dim, out_dim, kind):
        super().__init__()
        if kind == "freq":
            lim = math.sqrt(6.0 / in_dim) / 25.0
            w = torch.empty(out_dim, in_dim).uniform_(-lim, lim)
        else:
            w = 0.25 * nn.init.kaiming_normal_(torch.empty(out_dim, in_dim), a=0.2, mode="fan_in",
                                               nonlinearity="leaky_relu")
        self.weight = nn.Parameter(w)
        lim = math.sqrt(1.0 / in_dim)
        self.bias = nn.Parameter(torch.empty(out_dim).uniform_(-lim, lim))


class _FiLM(nn.Module):
    """FiLMSiren parameters (volume_renderer.py:39-67)."""

    def __init__(self, in_ch, out_ch, style_dim, is_first=False):
        super().__init__()
        lim = 1.0 / 3.0 if is_first else math.sqrt(6.0 / in_ch) / 25.0
        self.weight = nn.Parameter(torch.empty(out_ch, in_ch).uniform_(-lim, lim))
        lim = math.sqrt(1.0 / in_ch)
        self.bias = nn.Parameter(torch.empty(out_ch).uniform_(-lim, lim))
        self.gamma = _Linear(style_dim, out_ch, "style")
        self.beta = _Linear(style_dim, out_ch, "style")


class _SirenParams(nn.Module):
    """SirenGenerator parameters (volume_renderer.py:89-116)."""

    def __init__(self, D, hidden, style_dim, input_ch, view_ch):
        super().__init__()
        self.pts_linears = nn.ModuleList(
            [_FiLM(input_ch, hidden, style_dim, is_first=True)] +
            [_FiLM(hidden, hidden, style_dim) for _ in range(D - 1)])
        self.views_linears = _FiLM(view_ch + hidden, hidden, style_dim)
        self.rgb_linear = _Linear(hidden, 3, "freq")
        self.sigma_linear = _Linear(hidden, 1, "freq")


def _f32c(t):
    return _aligned(t.detach().to(torch.float32).contiguous())


def _aligned(t):
    """The C ABI wants 16-byte aligned pointers; a contiguous view with a storage offset (e.g. the reference's ray
    chunks `rays_d[:, i:i+n]` at batch 1, model_v3.py:1233-1249) may not be -- copy those."""
    return t.clone() if t.data_ptr() % 16 else t


# ---------------------------------------------------------------------------------------------
# autograd bridge
# ---------------------------------------------------------------------------------------------
class _NerfFn(torch.autograd.Function):
    """forward = c3d_nerf_forward, backward = c3d_nerf_backward (recomputes activations in-kernel)."""

    @staticmethod
    def forward(ctx, module, kind, meta, styles, a0, a1, a2, a3, near, far, *params):
        # kind POINTS: a0..a3 = pts, rays_d, viewdirs, z_vals ; kind POSES: a0, a1, a2 = cam_poses, focal, ray_offset
        # A step that will be differentiated runs the save-mode forward once and keeps its workspace for the backward
        # (no recomputation) when that fits the memory budget; otherwise plain forward + chunked recomputation.
        saved = None if any(p.requires_grad for p in params) else \
            module._launch_forward_save(kind, meta, styles, a0, a1, a2, a3, near, far)
        if saved is not None:
            outs, ctx.saved_ws = saved
        else:
            outs, ctx.saved_ws = module._launch_forward(kind, meta, styles, a0, a1, a2, a3, near, far), None
        ctx.module, ctx.kind, ctx.meta = module, kind, meta
        ctx.set_materialize_grads(False)             # an output the loss never touched has no cotangent (not a tensor of zeros)
        ctx.save_for_backward(styles, a0, a1, a2 if a2 is not None else styles.new_empty(0),
                              a3 if a3 is not None else styles.new_empty(0), near, far)
        ctx.has = (a2 is not None, a3 is not None)
        ctx.mark_non_differentiable(outs[5]) if outs[5] is not None else None
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, g_rgb, g_feat, g_sdf, g_mask, g_xyz, g_z):
        styles, a0, a1, a2, a3, near, far = ctx.saved_tensors
        a2 = a2 if ctx.has[0] else None
        a3 = a3 if ctx.has[1] else None
        grads = ctx.module._launch_backward(ctx.kind, ctx.meta, styles, a0, a1, a2, a3, near, far,
                                            g_rgb, g_feat, g_sdf, g_mask, g_xyz, ctx.needs_input_grad,
                                            saved_ws=ctx.saved_ws)
        return (None, None, None) + grads


class _EikonalFn(torch.autograd.Function):
    """Eikonal term E = d sdf / d pts (nerf_utils.py:220-228).  forward: one c3d_nerf_backward launch with a cotangent of
    ones on sdf (points are independent).  backward: c3d_eikonal_backward, the reverse-over-forward sweep that returns
    dL/d styles and dL/d parameters for a loss on E (the training-time double backward).  No gradient is returned for
    `pts` (a Hessian-vector product the reference's training never uses: cameras are sampled, not optimised)."""

    @staticmethod
    def forward(ctx, module, meta, styles, pts, rays_d, viewdirs, z_vals, near, far, *params):
        ones = torch.ones(meta[0], meta[1], meta[2], dtype=torch.float32, device=pts.device)
        needs = (False, False, False, False, True) + (False,) * 5 + (False,) * len(params)
        grads = module._launch_backward(_abi.INPUT_POINTS, meta, styles, pts, rays_d, viewdirs, z_vals, near, far,
                                        None, None, ones, None, None, needs)
        ctx.module, ctx.meta = module, meta
        ctx.save_for_backward(styles, pts, rays_d, viewdirs, z_vals, near, far)
        return grads[1]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_eik):
        styles, pts, rays_d, viewdirs, z_vals, near, far = ctx.saved_tensors
        needs = ctx.needs_input_grad
        g_styles, g_params = ctx.module._launch_eikonal_backward(ctx.meta, styles, pts, rays_d, viewdirs, z_vals, near, far,
                                                                 g_eik, needs[2], needs[9:])
        return (None, None, g_styles, None, None, None, None, None, None) + tuple(g_params)


class NerfBranch(nn.Module):
    def __init__(self, N_layers_renderer, input_dim=3, hidden_dim=256, style_dim=256, view_dim=3, with_sdf=True,
                 output_features=True, precision="bf16", **kwargs):
        super().__init__()
        if (input_dim, hidden_dim, style_dim, view_dim) != (3, W, W, 3) or not with_sdf or not output_features:
            raise ValueError("NerfBranch supports input_dim=3, hidden_dim=256, style_dim=256, view_dim=3, "
                             "with_sdf=True, output_features=True (the v10 configs)")
        if not 1 <= N_layers_renderer <= _abi.MAX_LAYERS:
            raise ValueError(f"N_layers_renderer must be in [1, {_abi.MAX_LAYERS}]")
        self.N_layers_renderer = N_layers_renderer
        self.input_dim, self.hidden_dim, self.style_dim, self.view_dim = input_dim, hidden_dim, style_dim, view_dim
        self.with_sdf, self.output_features = with_sdf, output_features
        self.precision = precision
        self.sigmoid_beta = nn.Parameter(0.1 * torch.ones(1))
        self.network = _SirenParams(N_layers_renderer, hidden_dim, style_dim, input_dim, view_dim)
        self._cache = None          # (key, packed blob); never pickled / deep-copied
        # The packed blob is re-derived from the parameters on EVERY call unless `cache_packed` is set.  Tensor version
        # counters cannot be trusted as a cache key: writes through `.data` (the reference's EMA update,
        # exp/cips3d/utils.py:79, every training iteration) do not bump them.  Serving code that knows its weights are
        # frozen sets `cache_packed = True` (then the key is (data_ptr, _version) per parameter) or calls
        # `invalidate_cache()` after weight surgery.
        self.cache_packed = False

    # -- construction helpers ------------------------------------------------------------------
    @classmethod
    def from_reference(cls, renderer, precision="bf16"):
        """Build from a reference VolumeFeatureRenderer (or anything with the same state_dict)."""
        m = cls(renderer.N_layers_renderer, precision=precision)
        m.load_state_dict(renderer.state_dict(), strict=True)
        p = next(renderer.parameters())
        m = m.to(p.device)
        # a frozen / eval generator (projector_v9.py:68 `deepcopy(G).eval().requires_grad_(False)`) stays frozen / eval:
        # otherwise every call would take the parameter-gradient path
        src = dict(renderer.named_parameters())
        for name, q in m.named_parameters():
            q.requires_grad_(src[name].requires_grad)
        m.train(renderer.training)
        return m

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_cache"] = None
        return d

    def extra_repr(self):
        return f"N_layers_renderer={self.N_layers_renderer}, precision={self.precision}"

    # -- packed weights -------------------------------------------------------------------------
    def _ordered_params(self):
        net = self.network
        ps = []
        for layer in list(net.pts_linears) + [net.views_linears]:
            ps += [layer.weight, layer.bias, layer.gamma.weight, layer.gamma.bias, layer.beta.weight, layer.beta.bias]
        ps += [net.rgb_linear.weight, net.rgb_linear.bias, net.sigma_linear.weight, net.sigma_linear.bias,
               self.sigmoid_beta]
        return ps

    def _fill_param_struct(self, raw, tensors):
        """Device pointers of `tensors` (in `_ordered_params` order) into a RawParams / ParamGrads struct."""
        D = self.N_layers_renderer
        it = iter(tensors)
        for l in range(D + 1):
            w, b, gw, gb, bw, bb = (next(it) for _ in range(6))
            if l < D:
                raw.pts_weight[l], raw.pts_bias[l] = w.data_ptr(), b.data_ptr()
                raw.pts_gamma_weight[l], raw.pts_gamma_bias[l] = gw.data_ptr(), gb.data_ptr()
                raw.pts_beta_weight[l], raw.pts_beta_bias[l] = bw.data_ptr(), bb.data_ptr()
            else:
                raw.views_weight, raw.views_bias = w.data_ptr(), b.data_ptr()
                raw.views_gamma_weight, raw.views_gamma_bias = gw.data_ptr(), gb.data_ptr()
                raw.views_beta_weight, raw.views_beta_bias = bw.data_ptr(), bb.data_ptr()
        rw, rb, sw, sb, sbeta = (next(it) for _ in range(5))
        raw.rgb_weight, raw.rgb_bias = rw.data_ptr(), rb.data_ptr()
        raw.sigma_weight, raw.sigma_bias, raw.sigmoid_beta = sw.data_ptr(), sb.data_ptr(), sbeta.data_ptr()

    def invalidate_cache(self):
        """Drop the packed blob (only matters with `cache_packed = True`)."""
        self._cache = None

    def packed_weights(self):
        """Kernel-layout weight blob.  Re-packed (two small kernels, in stream order) on every call; with
        `cache_packed = True` only when a parameter's (data_ptr, _version) changed."""
        ps = self._ordered_params()
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("NerfBranch parameters must live on a CUDA device (no CPU path)")
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self.cache_packed and self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        lib = _abi.load()
        D = self.N_layers_renderer
        keep = [_f32c(p) for p in ps]
        raw = _abi.RawParams()
        raw.D = D
        self._fill_param_struct(raw, keep)
        nbytes = lib.c3d_packed_bytes(D)
        old = self._cache[1] if self._cache is not None else None
        # the blob is rewritten in place in stream order (earlier launches on this stream have consumed it)
        blob = old if (old is not None and old.device == dev and old.numel() == nbytes
                       and not torch.cuda.is_current_stream_capturing()) else \
            torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_pack_weights(raw, _abi.ptr(blob), nbytes, torch.cuda.current_stream().cuda_stream),
                       "c3d_pack_weights")
        self._cache = (key, blob, keep)
        return blob

    # -- launches ---------------------------------------------------------------------------------
    def _mode(self, n_samples=None):
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if n_samples is not None and n_samples < _abi.MIN_SAMPLES_BF16:
            return _abi.MODE_FP32          # the tensor-core tiling needs >= 8 samples per ray; fewer run the fp32 kernels
        return _abi.MODE_BF16 if self.precision == "bf16" else _abi.MODE_FP32

    def _fill_common(self, P, kind, meta, styles, a0, a1, a2, a3, near, far):
        b, n_rays, N, img_size, static_viewdirs, nchw = meta
        P.abi_version = _abi.ABI_VERSION
        P.mode = self._mode(N)
        P.input_kind = kind
        P.feat_layout = int(nchw)                                  # _abi.FEAT_NHWC / FEAT_NCHW / FEAT_NCHW_BF16
        P.batch, P.n_rays, P.n_samples, P.D = b, n_rays, N, self.N_layers_renderer
        P.img_size, P.static_viewdirs = img_size, int(static_viewdirs)
        P.packed = self.packed_weights().data_ptr()
        P.styles, P.near, P.far = styles.data_ptr(), near.data_ptr(), far.data_ptr()
        if kind == _abi.INPUT_POINTS:
            P.pts, P.rays_d, P.viewdirs, P.z_vals = a0.data_ptr(), a1.data_ptr(), a2.data_ptr(), a3.data_ptr()
        else:
            P.cam_poses, P.focal = a0.data_ptr(), a1.data_ptr()
            P.ray_offset = a2.data_ptr() if a2 is not None else None

    def _launch_forward(self, kind, meta, styles, a0, a1, a2, a3, near, far, density_only=False, gather=None):
        lib = _abi.load()
        b, n_rays, N, img_size, static_viewdirs, nchw = meta
        dev = styles.device
        f = dict(dtype=torch.float32, device=dev)
        sdf = torch.empty(b, n_rays, N, 1, **f)
        z_out = torch.empty(b, n_rays, N, **f) if kind == _abi.INPUT_POSES else None
        P = _abi.FwdParams()
        self._fill_common(P, kind, meta, styles, a0, a1, a2, a3, near, far)
        if density_only:                                            # map pointers stay NULL: the kernel stops after the sdf head
            rgb_map = feat = mask = xyz = None
            P.sdf = sdf.data_ptr()
        elif gather is not None:                                    # fused all-gather: the maps go to every peer's tensors
            rgb_map = feat = mask = xyz = None
            P.sdf = sdf.data_ptr()
            gstruct = gather.struct()
            P.gather = C.pointer(gstruct)
        else:
            rgb_map = torch.empty(b, n_rays, 3, **f)
            feat = torch.empty((b, W, n_rays) if nchw else (b, n_rays, W), device=dev,
                               dtype=torch.bfloat16 if nchw == _abi.FEAT_NCHW_BF16 else torch.float32)
            mask = torch.empty(b, n_rays, 2, **f)
            xyz = torch.empty(b, n_rays, 3, **f)
            P.rgb_map, P.feature_map, P.sdf, P.mask, P.xyz = (t.data_ptr() for t in (rgb_map, feat, sdf, mask, xyz))
        P.z_vals_out = z_out.data_ptr() if z_out is not None else None
        nws = lib.c3d_workspace_bytes(P)
        ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
        P.workspace, P.workspace_bytes = ws.data_ptr(), nws
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_nerf_forward(P, torch.cuda.current_stream().cuda_stream), "c3d_nerf_forward")
        self.last_launch_count = lib.c3d_last_launch_count()
        return rgb_map, feat, sdf, mask, xyz, z_out

    def _launch_forward_save(self, kind, meta, styles, a0, a1, a2, a3, near, far):
        """c3d_nerf_forward_save: the forward of a differentiated step, run once in save mode.  Returns (outputs, workspace)
        or None when the path does not apply (fp32 mode, n_samples < 8, C3D_BWD=simt) or the whole-batch workspace exceeds
        the budget `C3D_SAVE_FWD_GB` (default 48 GiB of the 180 GB; 0 disables)."""
        if self._mode(meta[2]) != _abi.MODE_BF16:
            return None
        budget = float(os.environ.get("C3D_SAVE_FWD_GB", "48")) * (1 << 30)
        if budget <= 0:
            return None
        lib = _abi.load()
        b, n_rays, N, img_size, static_viewdirs, nchw = meta
        dev = styles.device
        B = _abi.BwdParams()
        self._fill_common(B.fwd, kind, meta, styles, a0, a1, a2, a3, near, far)
        B.fwd_saved = 1
        nws = lib.c3d_backward_workspace_bytes(B)
        if nws == 0 or nws > budget:
            return None
        try:                                                         # memory is tight: plain forward + chunked recomputation
            ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        except torch.cuda.OutOfMemoryError:
            return None
        f = dict(dtype=torch.float32, device=dev)
        rgb_map = torch.empty(b, n_rays, 3, **f)
        feat = torch.empty((b, W, n_rays) if nchw else (b, n_rays, W), **f)
        sdf = torch.empty(b, n_rays, N, 1, **f)
        mask = torch.empty(b, n_rays, 2, **f)
        xyz = torch.empty(b, n_rays, 3, **f)
        z_out = torch.empty(b, n_rays, N, **f) if kind == _abi.INPUT_POSES else None
        P = B.fwd
        P.rgb_map, P.feature_map, P.sdf, P.mask, P.xyz = (t.data_ptr() for t in (rgb_map, feat, sdf, mask, xyz))
        P.z_vals_out = z_out.data_ptr() if z_out is not None else None
        P.workspace, P.workspace_bytes = ws.data_ptr(), nws
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_nerf_forward_save(B, torch.cuda.current_stream().cuda_stream), "c3d_nerf_forward_save")
        self.last_launch_count = lib.c3d_last_launch_count()
        return (rgb_map, feat, sdf, mask, xyz, z_out), ws

    def _launch_backward(self, kind, meta, styles, a0, a1, a2, a3, near, far, g_rgb, g_feat, g_sdf, g_mask, g_xyz,
                         needs, saved_ws=None):
        """c3d_nerf_backward: recomputes the forward in fp32 and returns gradients for styles and the geometric
        inputs (POINTS: pts, rays_d, viewdirs; POSES: cam_poses, focal)."""
        lib = _abi.load()
        dev = styles.device
        b, n_rays, N, img_size, static_viewdirs, nchw = meta
        want_params = any(needs[10:])
        if nchw and g_feat is not None:
            g_feat = g_feat.transpose(1, 2)
        B = _abi.BwdParams()
        self._fill_common(B.fwd, kind, (b, n_rays, N, img_size, static_viewdirs, False), styles, a0, a1, a2, a3, near, far)
        cot = [None if g is None else g.to(torch.float32).contiguous() for g in (g_rgb, g_feat, g_mask, g_xyz, g_sdf)]
        B.g_rgb_map, B.g_feature_map, B.g_mask, B.g_xyz, B.g_sdf = (None if g is None else g.data_ptr() for g in cot)
        # needs_input_grad indices: module, kind, meta, styles, a0, a1, a2, a3, near, far, *params
        g_styles = torch.empty_like(styles) if needs[3] else None
        B.g_styles = None if g_styles is None else g_styles.data_ptr()
        g_a0 = g_a1 = g_a2 = None
        if kind == _abi.INPUT_POINTS:
            g_a0 = torch.empty_like(a0) if needs[4] else None
            g_a1 = torch.empty_like(a1) if needs[5] else None
            g_a2 = torch.zeros_like(a2) if needs[6] else None
            B.g_pts = None if g_a0 is None else g_a0.data_ptr()
            B.g_rays_d = None if g_a1 is None else g_a1.data_ptr()
            B.g_viewdirs = None if g_a2 is None else g_a2.data_ptr()
        else:
            g_a0 = torch.zeros_like(a0) if needs[4] else None
            g_a1 = torch.zeros_like(a1) if needs[5] else None
            B.g_cam_poses = None if g_a0 is None else g_a0.data_ptr()
            B.g_focal = None if g_a1 is None else g_a1.data_ptr()
        g_params, pg = [], None
        if want_params:                                              # training: the FP32-pipe backward also fills these
            g_params = [torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                        for p in self._ordered_params()]
            pg = _abi.ParamGrads()
            self._fill_param_struct(pg, g_params)
            B.g_params = C.cast(C.pointer(pg), C.c_void_p)
        if saved_ws is not None:                                     # tiles left by c3d_nerf_forward_save: no recomputation
            B.fwd_saved = 1
            nws, ws = saved_ws.numel(), saved_ws
        else:
            nws = lib.c3d_backward_workspace_bytes(B)
            ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
        B.fwd.workspace, B.fwd.workspace_bytes = ws.data_ptr(), nws
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_nerf_backward(B, torch.cuda.current_stream().cuda_stream), "c3d_nerf_backward")
        self.last_launch_count = lib.c3d_last_launch_count()
        g_params = [g.to(p.dtype) if need else None
                    for g, p, need in zip(g_params, self._ordered_params(), needs[10:])] if want_params \
            else [None] * (len(needs) - 10)
        return (g_styles, g_a0, g_a1, g_a2, None, None, None) + tuple(g_params)

    def _launch_eikonal_backward(self, meta, styles, pts, rays_d, viewdirs, z_vals, near, far, g_eik, need_styles,
                                 need_params):
        """c3d_eikonal_backward: dL/d styles and dL/d parameters for a loss on the eikonal term (FP32 pipe)."""
        lib = _abi.load()
        dev = styles.device
        b, n_rays, N = meta[0], meta[1], meta[2]
        B = _abi.BwdParams()
        self._fill_common(B.fwd, _abi.INPUT_POINTS, (b, n_rays, N, 0, False, False), styles, pts, rays_d, viewdirs, z_vals,
                          near, far)
        g_eik = g_eik.to(torch.float32).reshape(b, n_rays, N, 3).contiguous()
        g_styles = torch.empty_like(styles) if need_styles else None
        B.g_styles = None if g_styles is None else g_styles.data_ptr()
        want_params = any(need_params)
        g_params, pg = [None] * len(need_params), None
        if want_params:
            g_all = [torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                     for p in self._ordered_params()]
            pg = _abi.ParamGrads()
            self._fill_param_struct(pg, g_all)
            B.g_params = C.cast(C.pointer(pg), C.c_void_p)
        if g_styles is None and not want_params:
            return None, g_params
        nws = lib.c3d_eikonal_workspace_bytes(B)
        ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
        B.fwd.workspace, B.fwd.workspace_bytes = ws.data_ptr(), nws
        with torch.cuda.device(dev):
            _abi.check(lib.c3d_eikonal_backward(B, g_eik.data_ptr(), torch.cuda.current_stream().cuda_stream),
                       "c3d_eikonal_backward")
        self.last_launch_count = lib.c3d_last_launch_count()
        if want_params:                                              # after the launch: `.to` copies for non-fp32 parameters
            g_params = [g.to(p.dtype) if need else None for g, p, need in zip(g_all, self._ordered_params(), need_params)]
        return g_styles, g_params

    def _run(self, kind, meta, styles, a0, a1, a2, a3, near, far):
        tensors = [styles, a0, a1, near, far] + [t for t in (a2, a3) if t is not None]
        for t in tensors:
            if t.device.type != "cuda":
                raise RuntimeError("NerfBranch needs CUDA tensors: there is no CPU fallback")
        need_grad = torch.is_grad_enabled() and (
            any(t.requires_grad for t in tensors) or any(p.requires_grad for p in self.parameters()))
        if not need_grad:
            return self._launch_forward(kind, meta, styles, a0, a1, a2, a3, near, far)
        if meta[5] == _abi.FEAT_NCHW_BF16:
            raise RuntimeError("features_nchw='bf16' is an inference-only hand-off (no backward): render under torch.no_grad()")
        params = [p for p in self._ordered_params()]
        return _NerfFn.apply(self, kind, meta, styles, a0, a1, a2, a3, near, far, *params)

    # -- public API -----------------------------------------------------------------------------
    def forward(self, pts, rays_d, viewdirs, z_vals, near, far, styles=None, return_eikonal=False,
                N_samples_forward=None):
        """VolumeFeatureRenderer.forward (volume_renderer.py:192-283).

        pts (b,hw,N,3) or (b,h,w,N,3); rays_d, viewdirs (b,hw,3); z_vals (b,hw,N); near, far (b,1,1);
        styles (b,D+1,256).  Returns (rgb_map, feature_map, sdf, mask, xyz, eikonal_term=None).
        `N_samples_forward` (ray chunking to bound the unfused path's memory) is accepted and ignored:
        the fused kernel keeps per-point activations on chip.
        """
        if styles is None:
            raise ValueError("styles is required")
        lead = pts.shape[:-2]
        b, N = pts.shape[0], pts.shape[-2]
        n_rays = int(math.prod(lead[1:]))
        c = lambda t, *s: _aligned(t.to(torch.float32).reshape(*s).contiguous())
        meta = (b, n_rays, N, 0, False, False)
        rgb_map, feat, sdf, mask, xyz, _ = self._run(
            _abi.INPUT_POINTS, meta, c(styles, b, self.N_layers_renderer + 1, W), c(pts, b, n_rays, N, 3),
            c(rays_d, b, n_rays, 3), c(viewdirs, b, n_rays, 3), c(z_vals, b, n_rays, N), c(near, b), c(far, b))
        eik = None
        if return_eikonal:
            # d sdf / d pts (nerf_utils.py:220-228), differentiable w.r.t. styles and parameters (see _EikonalFn)
            args = (c(styles, b, self.N_layers_renderer + 1, W), c(pts, b, n_rays, N, 3), c(rays_d, b, n_rays, 3),
                    c(viewdirs, b, n_rays, 3), c(z_vals, b, n_rays, N), c(near, b), c(far, b))
            eik = _EikonalFn.apply(self, meta, *args, *self._ordered_params()).reshape(*lead, N, 3)
        if len(lead) > 2:
            rgb_map, feat, mask, xyz = (t.reshape(*lead, t.shape[-1]) for t in (rgb_map, feat, mask, xyz))
            sdf = sdf.reshape(*lead, N, 1)
        return rgb_map, feat, sdf, mask, xyz, eik

    def render(self, cam_poses, focal, near, far, styles, img_size=64, N_samples=24, static_viewdirs=False,
               perturb=False, ray_offset=None, features_nchw=False, density_only=False, gather=None):
        """Fused fast path = Render.prepare_nerf_inputs (nerf_utils.py:172-218) + forward, rays generated
        in-kernel.  cam_poses (b,3,4), focal/near/far (b,1,1) or (b,).  Returns a dict of maps; `z_vals`
        are the sample depths the kernel used.  `features_nchw`: False -> feature_map (b, hw, 256) as the reference's
        renderer returns it; True -> (b, 256, hw), the layout the decoder consumes (model_v3.py:1014); "bf16" -> the same in
        bfloat16 (inference only) -- each written directly by the kernel's compositing epilogue.  `density_only=True` (no autograd) stops after the sdf head and returns
        only `sdf` and `z_vals` -- the coarse pass of `render_hierarchical`."""
        b = cam_poses.shape[0]
        n_rays = img_size * img_size
        c = lambda t, *s: _aligned(t.to(torch.float32).reshape(*s).contiguous())
        if perturb and ray_offset is None:
            ray_offset = torch.rand(b, img_size, img_size, 1, device=cam_poses.device)   # nerf_utils.py:110
        ro = None if ray_offset is None else c(ray_offset, b, n_rays)
        if features_nchw == "bf16" and self._mode(N_samples) != _abi.MODE_BF16:
            raise ValueError("features_nchw='bf16' needs precision='bf16' and N_samples >= 8 (it comes from the tensor-core kernels)")
        layout = _abi.FEAT_NCHW_BF16 if features_nchw == "bf16" else (_abi.FEAT_NCHW if features_nchw else _abi.FEAT_NHWC)
        meta = (b, n_rays, N_samples, img_size, bool(static_viewdirs), layout)
        args = (_abi.INPUT_POSES, meta, c(styles, b, self.N_layers_renderer + 1, W), c(cam_poses, b, 3, 4), c(focal, b),
                ro, None, c(near, b), c(far, b))
        if density_only:
            if any(t.device.type != "cuda" for t in (styles, cam_poses)):
                raise RuntimeError("NerfBranch needs CUDA tensors: there is no CPU fallback")
            with torch.no_grad():
                out = self._launch_forward(*args, density_only=True)
            return dict(sdf=out[2], z_vals=out[5])
        if gather is not None:
            # multi-GPU serving: `gather` is a dist.GatheredMaps -- the kernel writes rgb_map / feature_map / mask / xyz of this
            # rank's images into the gathered tensors of EVERY rank (peer memory); inference only, bf16 mode.  The caller runs
            # gather.barrier() before reading gather.feature_map etc.
            if any(t.device.type != "cuda" for t in (styles, cam_poses)):
                raise RuntimeError("NerfBranch needs CUDA tensors: there is no CPU fallback")
            if b != gather.batch_per_rank or n_rays != gather.n_rays:
                raise ValueError("gather buffers were sized for another batch / image size")
            meta = meta[:5] + (gather.layout,)
            with torch.no_grad():
                out = self._launch_forward(args[0], meta, *args[2:], gather=gather)
            return dict(sdf=out[2], z_vals=out[5], gathered=gather)
        rgb_map, feat, sdf, mask, xyz, z = self._run(*args)
        return dict(rgb_map=rgb_map, feature_map=feat, sdf=sdf, mask=mask, xyz=xyz, z_vals=z)

    def render_hierarchical(self, cam_poses, focal, near, far, styles, img_size=64, N_samples=24, N_importance=24,
                            static_viewdirs=False, perturb=False, ray_offset=None, u=None, features_nchw=False,
                            coarse_maps=False):
        """EXTENSION, off by default (the reference renders in one pass; BASELINE.json's north star asks for a
        `sample_pdf` + fine pass): coarse pass (`render`, no grad) -> `Render.importance_depths` (c3d_sample_pdf: PDF from
        the coarse compositing weights, N_importance new depths, ascending union) -> second pass over the N_samples +
        N_importance merged depths with the same network (pi-GAN style; the point MLP is pointwise, so evaluating the
        union equals merging the coarse and fine outputs).  The new depths are constants of the fine pass; gradients flow
        through the fine pass to styles, parameters and -- through the rays -- to cam_poses / focal.
        The coarse pass only has to produce densities, so by default it is a density-only launch (the kernel stops after the
        sdf head: no view layer, rgb head or compositing); `coarse_maps=True` renders its maps too.
        Returns the fine pass's dict of maps (`z_vals` = merged depths) with the coarse pass's dict under "coarse"."""
        from .nerf_utils import Render
        if N_samples + N_importance > 256:
            raise ValueError("N_samples + N_importance must be <= 256")
        b, n_rays, N2 = cam_poses.shape[0], img_size * img_size, N_samples + N_importance
        c = lambda t, *s: _aligned(t.to(torch.float32).reshape(*s).contiguous())
        with torch.no_grad():
            coarse = self.render(cam_poses, focal, near, far, styles, img_size=img_size, N_samples=N_samples,
                                 static_viewdirs=static_viewdirs, perturb=perturb, ray_offset=ray_offset,
                                 density_only=not coarse_maps)
        rays_o, rays_d, viewdirs = Render.get_rays_in_world(focal=focal, img_size=img_size, c2w=cam_poses,
                                                            static_viewdirs=static_viewdirs)
        rays_o, rays_d, viewdirs = (t.reshape(b, n_rays, 3) for t in (rays_o, rays_d, viewdirs))
        geom_grad = torch.is_grad_enabled() and (cam_poses.requires_grad or focal.requires_grad)
        imp = Render.importance_depths(coarse["z_vals"], N_importance, sdf=coarse["sdf"], rays_d=rays_d,
                                       sigmoid_beta=self.sigmoid_beta, rays_o=rays_o, u=u, perturb=perturb,
                                       return_pts=not geom_grad)
        z = imp["z_merged"]
        pts = Render.get_points(rays_o, rays_d, z) if geom_grad else imp["pts"]
        meta = (b, n_rays, N2, 0, False, bool(features_nchw))
        rgb_map, feat, sdf, mask, xyz, _ = self._run(
            _abi.INPUT_POINTS, meta, c(styles, b, self.N_layers_renderer + 1, W), c(pts, b, n_rays, N2, 3),
            c(rays_d, b, n_rays, 3), c(viewdirs, b, n_rays, 3), z, c(near, b), c(far, b))
        return dict(rgb_map=rgb_map, feature_map=feat, sdf=sdf, mask=mask, xyz=xyz, z_vals=z, z_fine=imp["z_fine"],
                    coarse=coarse)

    def mlp_init_pass(self, cam_poses, focals, img_size, near, far, styles, nerf_cfg):
        """Sphere-initialisation pass (volume_renderer.py:569-634): sdf of stratified samples (`offset_sampling=False`)
        and its target `|pts| - (far - near) / 4`; differentiable w.r.t. the renderer parameters (FP32-pipe backward)."""
        from .nerf_utils import Render
        rays_o, rays_d, viewdirs = Render.get_rays_in_world(focal=focals, img_size=img_size, c2w=cam_poses)
        z_vals = Render.get_z_vals(near=near, far=far, rays_d=rays_d, N_samples=nerf_cfg["N_samples"],
                                   offset_sampling=False)
        pts = Render.get_points(rays_o, rays_d, z_vals)
        b = pts.shape[0]
        sdf = self.forward(pts=pts, rays_d=rays_d.reshape(b, -1, 3), viewdirs=viewdirs.reshape(b, -1, 3),
                           z_vals=z_vals.reshape(b, -1, z_vals.shape[-1]), near=near, far=far, styles=styles)[2]
        sdf = sdf.squeeze(-1)
        target_values = pts.detach().norm(dim=-1) - (far - near).view(-1, 1, 1, 1) / 4
        return sdf, target_values
