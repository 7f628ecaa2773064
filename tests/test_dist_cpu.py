"""Multi-process plumbing on CPU (gloo, world_size 2): sharding, map gathering, gradient all-reduce, and the
bench.py reference arm under torchrun semantics (rank 0 prints, other ranks exit silently)."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from cips3dpp_b200.dist import shard_range
    for n in (0, 1, 7, 256, 257):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cips3dpp_b200.dist import allreduce_grads, gather_maps, shard_range
    n = 5                                              # ragged: rank 0 gets 3 images, rank 1 gets 2
    s, e = shard_range(n, rank, world)
    full = torch.arange(n * 4 * 3, dtype=torch.float32).reshape(n, 4, 3)
    got = gather_maps(full[s:e].clone(), n)
    ok_gather = torch.equal(got, full)
    g1, g2 = torch.full((2, 3), float(rank + 1)), torch.full((4,), 10.0 * (rank + 1))
    allreduce_grads([g1, g2])
    ok_reduce = torch.equal(g1, torch.full((2, 3), 3.0)) and torch.equal(g2, torch.full((4,), 30.0))
    q.put((rank, ok_gather, ok_reduce))
    dist.destroy_process_group()


def test_gather_and_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(g and r for _, g, r in res), res


def test_reference_arm_under_torchrun_env():
    """`bench.py --impl reference --gpus 2`: rank 0 prints the JSON line, rank 1 prints nothing and exits 0."""
    env0 = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    env1 = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--config", "c1"]
    r1 = subprocess.run(cmd, env=env1, capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run(cmd, env=env0, capture_output=True, text=True, timeout=600)
    assert r0.returncode == 0, r0.stderr
    line = json.loads([l for l in r0.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "nerf_branch_rays_per_s" and line["value"] > 0
    # "reference" where the live reference modules are available (build container, oracle/_ref on the GPU box), else the port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["n_gpus"] == 2


class _FakeRenderer:
    """Stands in for NerfBranch.render on CPU: the host logic of gen_maps (numbering, seeding, file output) is under test."""
    def __init__(self):
        self.calls = 0

    def render(self, pose, focal, near, far, styles, img_size, N_samples, static_viewdirs=False, features_nchw=False):
        self.calls += 1
        b = pose.shape[0]
        hw = img_size * img_size
        assert styles.shape == (b, 3, 256) and focal.shape[0] == b
        rgb = pose[:, :, 3].reshape(b, 1, 3).expand(b, hw, 3).clone()          # camera origin as a recognisable colour
        return dict(rgb_map=rgb, mask=torch.zeros(b, hw, 2), feature_map=torch.zeros(b, 256, hw))


def _gen_worker(rank, world, port, out_dir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cips3dpp_b200.gen_maps import gaussian_styles, gen_maps
    r = _FakeRenderer()
    got = gen_maps(r, gaussian_styles(2), dict(fov_ang=6, dist_radius=0.12), num_imgs=11, batch_gpu=3, out_dir=out_dir,
                   rank=rank, world_size=world, img_size=4, N_samples=8, device="cpu")
    q.put((rank, [i for i, _ in got], r.calls))
    dist.destroy_process_group()


def test_gen_maps_numbering_covers_all_images_gloo_world2(tmp_path):
    """gen_images.py:57-91 semantics: interleaved numbering, ceil(num/batch) steps on every rank, extras dropped."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gen_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = sorted(i for _, ids, _ in res for i in ids)
    assert idx == list(range(11))
    assert all(calls == 2 for _, _, calls in res)                  # ceil(11 / (3 * 2)) steps on both ranks
    files = sorted(os.listdir(tmp_path))
    assert files == [f"{i:0>5}.npz" for i in range(11)]
    import numpy as np
    z = np.load(os.path.join(tmp_path, files[0]))
    assert z["thumb"].shape == (4, 4, 3) and z["thumb"].dtype == np.uint8
