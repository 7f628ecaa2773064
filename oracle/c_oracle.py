"""ctypes wrapper of oracle/nerf_oracle.c -- TEST INFRASTRUCTURE (see the header of that file)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libnerf_oracle.so")
_fp = C.POINTER(C.c_float)


class OracleParams(C.Structure):
    _fields_ = [("D", C.c_int)] + [(n, _fp * 17) for n in ("weight", "bias", "gamma_w", "gamma_b", "beta_w", "beta_b")] + \
        [(n, _fp) for n in ("rgb_w", "rgb_b", "sigma_w", "sigma_b", "sigmoid_beta")]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "nerf_oracle.c")):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(LIB)
        _lib.oracle_renderer_forward.restype = C.c_int
        _lib.oracle_prepare_inputs.restype = C.c_int
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_fp)


def pack_params(params):
    """params: reference-named dict of float32 arrays -> (OracleParams, keepalive list)."""
    D = sum(1 for k in params if k.startswith("network.pts_linears.") and k.endswith(".gamma.weight"))
    P = OracleParams()
    P.D = D
    keep = []

    def f(k):
        a = np.ascontiguousarray(params[k], np.float32)
        keep.append(a)
        return _p(a)

    for l in range(D + 1):
        pre = f"network.pts_linears.{l}." if l < D else "network.views_linears."
        P.weight[l], P.bias[l] = f(pre + "weight"), f(pre + "bias")
        P.gamma_w[l], P.gamma_b[l] = f(pre + "gamma.weight"), f(pre + "gamma.bias")
        P.beta_w[l], P.beta_b[l] = f(pre + "beta.weight"), f(pre + "beta.bias")
    P.rgb_w, P.rgb_b = f("network.rgb_linear.weight"), f("network.rgb_linear.bias")
    P.sigma_w, P.sigma_b = f("network.sigma_linear.weight"), f("network.sigma_linear.bias")
    P.sigmoid_beta = f("sigmoid_beta")
    return P, keep


def renderer_forward(params, pts, rays_d, viewdirs, z_vals, near, far, styles, nthreads=0, packed=None):
    lib = load()
    P, keep = packed if packed is not None else pack_params(params)
    c = lambda a: np.ascontiguousarray(a, np.float32)
    pts, rays_d, viewdirs, z_vals, styles = map(c, (pts, rays_d, viewdirs, z_vals, styles))
    near, far = c(near).reshape(-1), c(far).reshape(-1)
    b, n_rays, N = pts.shape[0], pts.shape[1], pts.shape[2]
    rgb_map = np.empty((b, n_rays, 3), np.float32)
    feat = np.empty((b, n_rays, 256), np.float32)
    sdf = np.empty((b, n_rays, N, 1), np.float32)
    mask = np.empty((b, n_rays, 2), np.float32)
    xyz = np.empty((b, n_rays, 3), np.float32)
    rc = lib.oracle_renderer_forward(C.byref(P), b, n_rays, N, _p(pts), _p(rays_d), _p(viewdirs), _p(z_vals), _p(near),
                                     _p(far), _p(styles), _p(rgb_map), _p(feat), _p(sdf), _p(mask), _p(xyz), int(nthreads))
    assert rc == 0
    return rgb_map, feat, sdf, mask, xyz


def prepare_inputs(c2w, focal, near, far, img_size, N, static_viewdirs=False, ray_offset=None):
    lib = load()
    c = lambda a: np.ascontiguousarray(a, np.float32)
    c2w, focal, near, far = c(c2w), c(focal).reshape(-1), c(near).reshape(-1), c(far).reshape(-1)
    b, hw = c2w.shape[0], img_size * img_size
    pts = np.empty((b, hw, N, 3), np.float32)
    rays_d = np.empty((b, hw, 3), np.float32)
    viewdirs = np.empty((b, hw, 3), np.float32)
    z_vals = np.empty((b, hw, N), np.float32)
    ro = None if ray_offset is None else c(ray_offset).reshape(-1)
    lib.oracle_prepare_inputs(b, img_size, N, int(static_viewdirs), _p(c2w), _p(focal), _p(near), _p(far),
                              None if ro is None else _p(ro), _p(pts), _p(rays_d), _p(viewdirs), _p(z_vals))
    return pts, rays_d, viewdirs, z_vals


def num_threads():
    return load().oracle_num_threads()
