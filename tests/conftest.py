import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_weights(D):
    """Golden D=8 state dict (fp16-exact values) cut down to D point layers, as float32 numpy."""
    z = np.load(os.path.join(GOLDEN, "weights_seed0.npz"))
    out = {}
    for k in z.files:
        if k.startswith("network.pts_linears.") and int(k.split(".")[2]) >= D:
            continue
        out[k] = z[k].astype(np.float32)
    return out


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    return {k: z[k] for k in z.files}


CASES = ["ffhq_d8_n24", "ffhq_d2_n24", "ffhq_d2_n128_static", "cars_d6_n24",
         "ffhq_d8_n24_b2_wplus_perturb", "cars_d6_n36_b2_beta"]
GRAD_CASES = ["ffhq_d2_n24_grads", "ffhq_d8_n24_grads_static"]


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _default_kernel_options():
    """Kernel-variant options set by a test (c3d_set_option is process-wide) do not leak into the next one."""
    yield
    import cips3dpp_b200 as c3d
    if c3d._abi._lib is not None:
        c3d._abi.set_options(fwd="pair", cluster=2, grid=0, egw=4, bwd="tc", fp32="tc", resample="auto", resample_rb=0, debug=0)
