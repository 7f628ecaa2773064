"""Dev helper: MN-major operand self-test over variants."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cips3dpp_b200 as c3d
lib = c3d._abi.load()
dev = torch.device("cuda:0")
for (N, K) in ((128, 256), (64, 128), (128, 64)):
    rng = np.random.default_rng(N + K)
    a = rng.integers(-4, 5, size=(128, K)).astype(np.float32)
    b = rng.integers(-4, 5, size=(N, K)).astype(np.float32)
    ta = torch.from_numpy(a).to(dev).to(torch.bfloat16).view(torch.int16)
    tb = torch.from_numpy(b).to(dev).to(torch.bfloat16).view(torch.int16)
    for variant in (1, 2, 3, 5, 6, 7):
        d = torch.full((128, N), float("nan"), device=dev)
        c3d._abi.check(lib.c3d_umma_selftest(ta.data_ptr(), tb.data_ptr(), d.data_ptr(), N, K, variant,
                                             torch.cuda.current_stream().cuda_stream), "selftest")
        torch.cuda.synchronize()
        err = np.abs(d.cpu().numpy() - a @ b.T).max()
        print(f"N={N} K={K} variant={variant} (A_mn={variant&1} B_mn={(variant>>1)&1} swap={(variant>>2)&1}) max err {err}")
