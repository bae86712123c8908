"""Linear 73728x320 -> 320 (+bias), a few launches (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import ops  # noqa: E402

dev = torch.device("cuda")
rows, k, n = 73728, 320, int(os.environ.get("N", "320"))
x = torch.randn(rows, k, device=dev).half()
w = ops.pack_linear(torch.randn(n, k, device=dev) * 0.05)
b = torch.zeros(n, device=dev).half()
for _ in range(4):
    y = ops.linear(x, w, bias=b)
torch.cuda.synchronize()
print("ok")
