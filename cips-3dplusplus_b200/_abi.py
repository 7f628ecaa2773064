"""ctypes binding of libc3dpp.so (include/c3d_abi.h).  No CPU fallback: if the library is missing or a
call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("C3D_LIB") or os.path.join(HERE, "libc3dpp.so")   # C3D_LIB: A/B builds (bench_tools)
SRC = os.path.join(HERE, "csrc", "c3d_abi.cu")
ABI_VERSION = 11
MAX_LAYERS = 16
MAX_PEERS = 16
MIN_SAMPLES_BF16 = 8     # fused::MIN_SAMPLES (csrc/fused_common.cuh)
MODE_FP32, MODE_BF16 = 0, 1
INPUT_POSES, INPUT_POINTS = 0, 1
FEAT_NHWC, FEAT_NCHW, FEAT_NCHW_BF16 = 0, 1, 2

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]

_fp = C.c_void_p


class RawParams(C.Structure):
    _fields_ = [("D", C.c_int32), ("_pad", C.c_int32)] + \
        [(n, _fp * MAX_LAYERS) for n in ("pts_weight", "pts_bias", "pts_gamma_weight", "pts_gamma_bias",
                                         "pts_beta_weight", "pts_beta_bias")] + \
        [(n, _fp) for n in ("views_weight", "views_bias", "views_gamma_weight", "views_gamma_bias",
                            "views_beta_weight", "views_beta_bias", "rgb_weight", "rgb_bias",
                            "sigma_weight", "sigma_bias", "sigmoid_beta")]


class ParamGrads(C.Structure):
    _fields_ = [(n, _fp * MAX_LAYERS) for n in ("pts_weight", "pts_bias", "pts_gamma_weight", "pts_gamma_bias",
                                                "pts_beta_weight", "pts_beta_bias")] + \
        [(n, _fp) for n in ("views_weight", "views_bias", "views_gamma_weight", "views_gamma_bias",
                            "views_beta_weight", "views_beta_bias", "rgb_weight", "rgb_bias",
                            "sigma_weight", "sigma_bias", "sigmoid_beta")]


class GatherOut(C.Structure):
    _fields_ = [("n_peers", C.c_int32), ("image_offset", C.c_int32)] + \
        [(n, _fp * MAX_PEERS) for n in ("feature_map", "rgb_map", "mask", "xyz")]


class FwdParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("abi_version", "mode", "input_kind", "feat_layout", "batch", "n_rays",
                                         "n_samples", "D", "img_size", "static_viewdirs")] + \
        [(n, _fp) for n in ("packed", "styles", "cam_poses", "focal", "near", "far", "ray_offset",
                            "pts", "rays_d", "viewdirs", "z_vals",
                            "rgb_map", "feature_map", "sdf", "mask", "xyz", "z_vals_out", "workspace")] + \
        [("workspace_bytes", C.c_size_t), ("gather", C.POINTER(GatherOut))]


class BwdParams(C.Structure):
    _fields_ = [("fwd", FwdParams)] + \
        [(n, _fp) for n in ("g_rgb_map", "g_feature_map", "g_mask", "g_xyz", "g_sdf",
                            "g_styles", "g_pts", "g_rays_d", "g_viewdirs", "g_cam_poses", "g_focal", "g_params")] + \
        [("fwd_saved", C.c_int32), ("_pad", C.c_int32)]


class RaygenParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "img_size", "n_samples", "static_viewdirs")] + \
        [(n, _fp) for n in ("cam_poses", "focal", "near", "far", "ray_offset", "pts", "rays_d", "viewdirs", "z_vals")]


COMPOSITE_RAW_DENSITY, COMPOSITE_FORCE_BACKGROUND = 1, 2


class CompositeParams(C.Structure):
    _fields_ = [("n_rays", C.c_int64), ("n_samples", C.c_int32), ("n_feat", C.c_int32),
                ("sigmoid_beta", C.c_float), ("flags", C.c_int32)] + \
        [(n, _fp) for n in ("sigmoid_beta_ptr", "rgb", "sdf", "features", "z_vals", "rays_d", "pts",
                            "rgb_map", "feature_map", "xyz", "mask", "weights",
                            "g_rgb_map", "g_feature_map", "g_xyz", "g_mask",
                            "g_rgb", "g_sdf", "g_features", "g_pts", "g_rays_d", "g_sigmoid_beta")]


class ResampleParams(C.Structure):
    _fields_ = [("n_rays", C.c_int64), ("n_samples", C.c_int32), ("n_importance", C.c_int32),
                ("sigmoid_beta", C.c_float), ("_pad", C.c_int32)] + \
        [(n, _fp) for n in ("sigmoid_beta_ptr", "z_vals", "weights", "sdf", "rays_d", "rays_o", "u",
                            "z_fine", "z_merged", "pts_merged")]


class AdamParams(C.Structure):
    _fields_ = [("n_tensors", C.c_int32), ("group", C.c_int32 * 8), ("numel", C.c_int64 * 8), ("param", _fp * 8),
                ("grad", _fp * 8), ("exp_avg", _fp * 8), ("exp_avg_sq", _fp * 8), ("lr", _fp * 2), ("step", _fp),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("max_norm", C.c_float), ("grad_norm", _fp)]


EXPORTS = {
    "c3d_abi_version": (C.c_int, []),
    "c3d_last_error": (C.c_char_p, []),
    "c3d_last_launch_count": (C.c_int, []),
    "c3d_set_option": (C.c_int, [C.c_char_p, C.c_char_p]),
    "c3d_packed_bytes": (C.c_size_t, [C.c_int32]),
    "c3d_pack_weights": (C.c_int, [C.POINTER(RawParams), _fp, C.c_size_t, _fp]),
    "c3d_workspace_bytes": (C.c_size_t, [C.POINTER(FwdParams)]),
    "c3d_nerf_forward": (C.c_int, [C.POINTER(FwdParams), _fp]),
    "c3d_backward_workspace_bytes": (C.c_size_t, [C.POINTER(BwdParams)]),
    "c3d_nerf_backward": (C.c_int, [C.POINTER(BwdParams), _fp]),
    "c3d_nerf_forward_save": (C.c_int, [C.POINTER(BwdParams), _fp]),
    "c3d_eikonal_workspace_bytes": (C.c_size_t, [C.POINTER(BwdParams)]),
    "c3d_eikonal_backward": (C.c_int, [C.POINTER(BwdParams), _fp, _fp]),
    "c3d_raygen": (C.c_int, [C.POINTER(RaygenParams), _fp]),
    "c3d_style_prep": (C.c_int, [_fp, C.c_int32, _fp, C.c_int32, _fp, _fp, _fp, _fp]),
    "c3d_composite_forward": (C.c_int, [C.POINTER(CompositeParams), _fp]),
    "c3d_composite_backward": (C.c_int, [C.POINTER(CompositeParams), _fp]),
    "c3d_camera_params": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, _fp, C.c_float, C.c_float, _fp, _fp, _fp, _fp, _fp, _fp]),
    "c3d_adam_clip_step": (C.c_int, [C.POINTER(AdamParams), _fp]),
    "c3d_sample_pdf": (C.c_int, [C.POINTER(ResampleParams), _fp]),
    "c3d_umma_selftest": (C.c_int, [_fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, _fp]),
}

_lib = None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into libc3dpp.so (in-tree).  nvcc cross-compiles without a GPU."""
    srcs = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "c3d_abi.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


def load():
    """Load libc3dpp.so; raises if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(this package has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    v = lib.c3d_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"libc3dpp.so ABI version {v} != binding {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def set_options(**kw):
    """Kernel-variant options for A/B runs and tests (c3d_set_option): fwd='pair'|'v3', cluster=1|2, grid=N, egw=4|8,
    bwd='tc'|'simt', resample='auto'|'warp'|'lane', resample_rb=N, debug=mask.  The library reads the C3D_* environment
    variables once at first use; after that only this call changes an option."""
    lib = load()
    for k, v in kw.items():
        check(lib.c3d_set_option(k.encode(), str(v).encode()), "c3d_set_option")


class C3DError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().c3d_last_error().decode("utf-8", "replace")
        raise C3DError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a contiguous float32/uint8 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
