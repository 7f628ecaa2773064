"""TEST INFRASTRUCTURE: import the UNMODIFIED reference generator (exp/cips3d/models/model_v3.py) in the build
container behind import-only stubs for third-party packages that are not installed (tl2, pytorch3d, trimesh,
skimage) and a pure-torch stand-in for the reference's own `op` CUDA extension (decoder side, out of scope).
Nothing here is used by the product."""
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the reference checkout (build container), or the three modules oracle/make_ref.sh placed under oracle/_ref (GPU box)
REF = os.environ.get("C3D_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/exp/cips3d")
                                          else os.path.join(_ROOT, "oracle", "_ref"))


def available():
    return os.path.isdir(os.path.join(REF, "exp", "cips3d"))


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, n):
        return _Any()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope, self.scale = negative_slope, scale

    def forward(self, x):
        return fused_leaky_relu(x, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(x, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    if bias is not None:
        x = x + bias.view(1, -1, *([1] * (x.ndim - 2)))
    return F.leaky_relu(x, negative_slope) * scale


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """Pure-torch stand-in for the reference's `op.upfirdn2d` CUDA extension (decoder side, out of scope): zero-insertion
    upsampling, padding / cropping, FIR filtering with the flipped kernel (a true convolution), decimation."""
    b, c, h, w = input.shape
    x = input.reshape(b * c, 1, h, w)
    if up > 1:
        y = x.new_zeros(b * c, 1, h * up, w * up)
        y[:, :, ::up, ::up] = x
        x = y
    p0, p1 = pad
    x = F.pad(x, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    x = x[:, :, max(-p0, 0):x.shape[2] - max(-p1, 0), max(-p0, 0):x.shape[3] - max(-p1, 0)]
    k = torch.flip(kernel, [0, 1])[None, None].to(x.dtype)
    x = F.conv2d(x, k)[:, :, ::down, ::down]
    return x.reshape(b, c, x.shape[2], x.shape[3])


def import_model_v3():
    class _Reg:
        def register(self, *a, **k):
            return lambda cls: cls

    _stub("trimesh")
    _stub("tl2")
    _stub("tl2.tl2_utils", get_class_repr=lambda self, prefix="": type(self).__name__,
          dict2string=lambda dict_obj, **k: str(dict_obj), print_repr=lambda *a, **k: None)
    _stub("tl2.proj")
    _stub("tl2.proj.fvcore", MODEL_REGISTRY=_Reg())
    _stub("tl2.proj.pytorch")
    _stub("tl2.proj.pytorch.torch_utils")
    _stub("pytorch3d")
    _stub("pytorch3d.renderer", TexturesUV=_Any, look_at_view_transform=_Any(), FoVPerspectiveCameras=_Any)
    _stub("pytorch3d.structures", Meshes=_Any)
    _stub("pytorch3d.transforms", matrix_to_euler_angles=_Any(), axis_angle_to_matrix=_Any())
    _stub("op", FusedLeakyReLU=FusedLeakyReLU, fused_leaky_relu=fused_leaky_relu, upfirdn2d=upfirdn2d)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import exp  # noqa: F401
    import exp.stylesdf  # noqa: F401
    _stub("exp.stylesdf.utils", create_cameras=_Any(), create_mesh_renderer=_Any(), add_textures=_Any(),
          create_depth_mesh_renderer=_Any())
    from exp.cips3d.models import model_v3
    from exp.cips3d import nerf_utils
    return model_v3, nerf_utils


def build_generator(model_v3, D=2, size_end=64, upsample_list=()):
    """FFHQ v10 generator (configs/train_cips3d_ffhq_v10.yaml: G_cfg, train_r1024_r64_ks1), random-init."""
    return model_v3.Generator(
        enable_decoder=True, freeze_renderer=False, renderer_detach=True, predict_rgb_residual=False, scale_factor=1,
        renderer_cfg=dict(N_layers_renderer=D, input_dim=3, hidden_dim=256, view_dim=3, with_sdf=True, output_features=True),
        mapping_renderer_cfg=dict(z_dim=256, style_dim=256, N_layers=3),
        decoder_cfg=dict(size_start=4, size_end=size_end, in_channel=256, channel_multiplier=2, project_noise=False,
                         upsample_list=list(upsample_list), kernel_size=1),
        mapping_decoder_cfg=dict(style_dim=512, lr_mul_mapping=0.01, N_layers=5))
