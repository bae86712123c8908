"""Spatial self-attention S=1536 d=40 (48 frames x 8 heads), a few launches (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import ops  # noqa: E402

dev = torch.device("cuda")
n, s, c = 48, 1536, 320
qkv = torch.randn(n * s, 3 * c, device=dev).half()
for _ in range(3):
    a = ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], n_batch=n, s_q=s, s_kv=s, heads=8, d=40, q_ld=3 * c,
                      kv_ld=3 * c)
torch.cuda.synchronize()
print("ok")
