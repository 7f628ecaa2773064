"""Coefficients of the odd degree-9 sine polynomial of csrc/mlp_tc32_sm100.cuh (sin_poly): Lawson-iterated weighted least squares
on [0, 1.6], evaluated with the kernel's fp32 operation order against float64 sin on [-100, 100]."""
import numpy as np
from numpy.polynomial import chebyshev as C
# fit g(s) = (sin(r)/r - 1)/s,  s = r^2, r in [0, 1.6]: sin r = r + r*s*g(s), g degree 3 (c3..c9)
R = 1.6
k = np.arange(4000); r = R*np.cos(np.pi*(k+0.5)/4000); r = np.abs(r)+1e-9
s = r*r
g = (np.sin(r)/r - 1)/s
# weighted LS for abs error of sin: err = r*s*(g_fit - g) -> weight r*s ; iterate (Remez-like via IRLS not needed)
best=None
for deg in (3,4):
    A = np.vander(s, deg+1, increasing=True)
    w = r*s
    c,*_ = np.linalg.lstsq(A*w[:,None], g*w, rcond=None)
    # Lawson iterations toward minimax
    lw = np.ones_like(w)
    for it in range(200):
        c,*_ = np.linalg.lstsq(A*(w*np.sqrt(lw))[:,None], g*w*np.sqrt(lw), rcond=None)
        e = np.abs((A@c-g)*w); lw = lw*e/ e.mean(); lw/=lw.sum()/len(lw)
    c32 = c.astype(np.float32)
    # evaluate in fp32
    x = np.linspace(-100,100,2000001).astype(np.float32)
    def fsin(x, c32):
        f=np.float32
        t = (x*f(0.318309886)).astype(np.float32)  # not fma-exact, fine
        n = np.rint(t).astype(np.float32)
        rr = (x.astype(np.float64) - n.astype(np.float64)*np.float64(f(3.14159274))).astype(np.float32)
        rr = (rr.astype(np.float64) - n.astype(np.float64)*np.float64(f(-8.742278e-8))).astype(np.float32)
        sgn = np.where(n.astype(np.int64)&1, f(-1), f(1))
        rr = rr*sgn
        ss = rr*rr
        q = np.full_like(ss, c32[-1])
        for cc in c32[-2::-1]: q = (q*ss+cc).astype(np.float32)
        rs = rr*ss
        return (rs.astype(np.float64)*q + rr).astype(np.float32)
    y = fsin(x, c32); ref = np.sin(x.astype(np.float64))
    print(deg, [float(v) for v in c32], "max abs err", np.abs(y-ref).max(), "rms", np.sqrt(((y-ref)**2).mean()))
    import math
ysin = np.sin(x.astype(np.float32)); print("np.sin fp32 err", np.abs(ysin-ref).max())
