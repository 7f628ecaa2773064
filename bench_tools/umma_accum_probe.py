"""How exact is a K = 256 bf16 x bf16 -> fp32 tcgen05 accumulation?  (Feasibility check for an fp32-accurate MLP on the tensor
cores by 3-way bf16 splitting: the operands are exact, what matters is how TMEM accumulates.)  Uses c3d_umma_selftest."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cips3dpp_b200 as c3d
lib = c3d._abi.load()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
K = N = 256
a = rng.uniform(-1, 1, size=(128, K)).astype(np.float32)            # activations (sines)
b = (rng.standard_normal((N, K)) * (1.0 / 16)).astype(np.float32)    # weights ~ U(-sqrt(6/256), ..)/...: |W h| ~ O(1)

def split3(x):
    t = torch.from_numpy(x)
    hi = t.to(torch.bfloat16); r1 = t - hi.float()
    mid = r1.to(torch.bfloat16); r2 = r1 - mid.float()
    lo = r2.to(torch.bfloat16)
    return hi, mid, lo

def mma(x, y):
    d = torch.full((128, N), float("nan"), device=dev)
    tx = x.contiguous().to(dev).view(torch.int16); ty = y.contiguous().to(dev).view(torch.int16)
    c3d._abi.check(lib.c3d_umma_selftest(tx.data_ptr(), ty.data_ptr(), d.data_ptr(), N, K, 0, torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    return d.cpu().numpy().astype(np.float64)

A = split3(a); B = split3(b)
exact = a.astype(np.float64) @ b.astype(np.float64).T
ref32 = (torch.from_numpy(a) @ torch.from_numpy(b).T).numpy().astype(np.float64)      # CPU fp32 GEMM
# one product, chained over K = 256 in TMEM, against the exact product of the same bf16 values
hh = mma(A[0], B[0]); hh_exact = A[0].double().numpy() @ B[0].double().numpy().T
e = hh - hh_exact
print(f"hi*hi, K=256 chained in TMEM: |acc| rms {np.sqrt((hh_exact**2).mean()):.3f}; error mean {e.mean():+.3e} rms {np.sqrt((e**2).mean()):.3e} max {np.abs(e).max():.3e}"
      f"; sign-correlated bias (mean of e*sign(acc)) {np.mean(e*np.sign(hh_exact)):+.3e};  fp32 eps*|acc| ~ {6e-8*np.sqrt((hh_exact**2).mean()):.1e}")
pairs = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]
tot6 = sum(mma(A[i], B[j]) for i, j in pairs)
tot9 = tot6 + sum(mma(A[i], B[j]) for i, j in [(1, 2), (2, 1), (2, 2)])
tot3 = sum(mma(A[i], B[j]) for i, j in pairs[:3])
for name, t in (("3 products (hi*hi, hi*mid, mid*hi)", tot3), ("6 products", tot6), ("9 products", tot9)):
    e = t - exact
    print(f"{name}, each K=256 chain in its own accumulator, summed on the host: error rms {np.sqrt((e**2).mean()):.3e} max {np.abs(e).max():.3e}")
e = ref32 - exact
print(f"CPU fp32 GEMM: error rms {np.sqrt((e**2).mean()):.3e} max {np.abs(e).max():.3e}")
