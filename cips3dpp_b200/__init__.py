"""Import name of the `cips-3dplusplus_b200/` package directory (a hyphen is not importable): this package's search path IS
that directory, so `cips3dpp_b200.nerf_branch` etc. are the modules there; the names below mirror its `__init__.py`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cips-3dplusplus_b200")]

from . import _abi                                  # noqa: E402
from .nerf_branch import NerfBranch                 # noqa: E402
from .nerf_utils import Render, Camera              # noqa: E402
from .patch import use_b200_nerf_branch             # noqa: E402
from . import dist                                  # noqa: E402
from .inversion import FlipInversion                # noqa: E402
from .gen_maps import gen_maps                      # noqa: E402

__all__ = ["NerfBranch", "Render", "Camera", "use_b200_nerf_branch", "dist", "FlipInversion", "gen_maps", "_abi"]
