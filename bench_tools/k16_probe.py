"""Dev helper: K=16 operand-layout self-test, both LBO/SBO assignments."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cips3dpp_b200 as c3d
lib = c3d._abi.load()
rng = np.random.default_rng(16)
a = rng.integers(-4, 5, size=(128, 16)).astype(np.float32)
b = rng.integers(-4, 5, size=(128, 16)).astype(np.float32)
dev = torch.device("cuda:0")
ta = torch.from_numpy(a).to(dev).to(torch.bfloat16).view(torch.int16)
tb = torch.from_numpy(b).to(dev).to(torch.bfloat16).view(torch.int16)
for variant in (0, 1):
    d = torch.full((128, 128), float("nan"), device=dev)
    c3d._abi.check(lib.c3d_umma_selftest(ta.data_ptr(), tb.data_ptr(), d.data_ptr(), 128, 16, variant,
                                         torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    err = np.abs(d.cpu().numpy() - a @ b.T).max()
    print("k16 variant", variant, "max err", err)
