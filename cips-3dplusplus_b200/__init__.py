"""cips-3dplusplus_b200: the NeRF branch of CIPS-3D++ on B200 (sm_100a) behind the reference's module API.

Import as `cips3dpp_b200` (the directory name has a hyphen; `cips3dpp_b200/` is the import alias).
"""
from . import _abi
from .nerf_branch import NerfBranch
from .nerf_utils import Render, Camera
from .patch import use_b200_nerf_branch
from . import dist
from .inversion import FlipInversion
from .gen_maps import gen_maps

__all__ = ["NerfBranch", "Render", "Camera", "use_b200_nerf_branch", "dist", "FlipInversion", "gen_maps", "_abi"]
