"""One conv3x3 320->320 [48,32,48] GEMM, a few launches (ncu target for tile/cluster experiments)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import ops  # noqa: E402

dev = torch.device("cuda")
n, h, w, c = 48, 32, 48, 320
x = torch.randn(n * h * w, c, device=dev).half()
wt = ops.pack_conv3x3(torch.randn(c, c, 3, 3, device=dev) * 0.02)
b = torch.zeros(c, device=dev).half()
for _ in range(4):
    y = ops.conv3x3(x, wt, n, h, w, bias=b)
torch.cuda.synchronize()
print("ok")
