"""Swap a reference generator's NeRF branch for the B200 one (the whole integration is one attribute)."""
from __future__ import annotations

from .nerf_branch import NerfBranch


def use_b200_nerf_branch(generator, precision="bf16"):
    """`generator`: an exp/cips3d/models/model_v3.Generator (or model/_v1/_v2/_v3_finetune) instance.

    Replaces `generator.renderer` (model_v3.py:848-851) by a NerfBranch holding the same weights; parameter
    names under `renderer.` are unchanged, so `load_state_dict(strict=True)`, checkpoints and optimiser
    grouping (train_v10.py:1110) keep working.  Returns the generator.
    """
    old = generator.renderer
    if isinstance(old, NerfBranch):
        old.precision = precision
        return generator
    generator.renderer = NerfBranch.from_reference(old, precision=precision)
    return generator
