"""ncu target: the GEMM variants added in the third session of round 2, two launches each (the second is the captured
one): 320-wide pair tiles (conv 1280->1280 + temb + residual at 8x12; FF out-projection 4608x5120->1280), the
activation-stationary QKV (73728x320->960) and GEGLU (73728x320->2560) GEMMs, the one-slab six-stage short-K pair kernel
(18432x640->1920)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
h16 = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()  # noqa: E731

n, h, w, c = 48, 8, 12, 1280
x, res, temb = h16(n * h * w, c), h16(n * h * w, c), h16(3, c)
wt, b = ops.pack_conv3x3(h16(c, c, 3, 3, sc=0.01)), h16(c)
for _ in range(2):
    ops.conv3x3(x, wt, n, h, w, bias=b, rowbias=temb, rowbias_group=16 * h * w, residual=res)
for rows, k, nn, r in [(4608, 5120, 1280, True), (73728, 320, 960, False), (18432, 640, 1920, False)]:
    xs, wl = h16(rows, k), ops.pack_linear(h16(nn, k, sc=k ** -0.5))
    rr = h16(rows, nn) if r else None
    for _ in range(2):
        ops.linear(xs, wl, bias=h16(nn) if r else None, residual=rr)
xs = h16(73728, 320)
wp, bp = ops.pack_geglu(h16(2560, 320, sc=320 ** -0.5), h16(2560, sc=0.3))
for _ in range(2):
    ops.linear(xs, wp, bias=bp, geglu=True)
torch.cuda.synchronize()
print("ok")
