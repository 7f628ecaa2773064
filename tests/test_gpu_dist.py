"""The path's two collectives on real GPUs (NCCL): all-gather of rendered maps and the gradient all-reduce of batched
inversion with a shared latent.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_gather_maps_and_grad_allreduce_nccl_world2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "dist_worker_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    print(line)
    assert line["world"] == 2
    # sharding does not change a ray's arithmetic beyond the tile partition (bf16 compositing per tile)
    assert line["gather_feat_rel"] < 1e-3 and line["gather_rgb_rel"] < 1e-3
    # summed partial losses and the shared latent follow the single-process run (north star: loss curves within 1 %)
    assert line["inv_loss_rel"] < 1e-2 and line["inv_w_rel"] < 2e-2
    assert line["last_loss"] < line["first_loss"]
    # fused all-gather: the kernel stored every rank's maps into every rank's gathered tensors (peer memory over NVLink) --
    # bit-identical to rendering all images on one GPU, with no collective launched
    assert "fused_error" not in line, line.get("fused_error")
    assert line["fused_all_ranks_ok"] and line["fused_feat_equal"] and line["fused_rgb_equal"] and line["fused_xyz_equal"]
    assert line["p2p_equal"]                        # the copy-engine gather (GatheredMaps.push) fills the same tensors
