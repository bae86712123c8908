"""The hot kernels at their config-2 shapes, a few launches each — the command captured by `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import ops  # noqa: E402

dev = torch.device("cuda")
n, h, w, c = 48, 32, 48, 320
rows = n * h * w
x = torch.randn(rows, c, device=dev).half()
res = torch.randn(rows, c, device=dev).half()
wt = ops.pack_conv3x3(torch.randn(c, c, 3, 3, device=dev) * 0.02)
wl = ops.pack_linear(torch.randn(c, c, device=dev) * 0.05)
wqkv = ops.pack_linear(torch.randn(3 * c, c, device=dev) * 0.05)
wg, bg = ops.pack_geglu(torch.randn(8 * c, c, device=dev) * 0.05, torch.zeros(8 * c, device=dev))
b = torch.zeros(c, device=dev).half()
g1 = torch.ones(c, device=dev).half()
for _ in range(3):
    y = ops.conv3x3(x, wt, n, h, w, bias=b, residual=res)          # gemm_tc_persistent<160,...> conv
    z = ops.linear(x, wl, bias=b, residual=res)                    # small-K linear + residual
    qkv = ops.linear(x, wqkv)                                      # qkv projection
    gg = ops.linear(x, wg, bias=bg, geglu=True)                    # GEGLU
    a = ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], n_batch=n, s_q=h * w, s_kv=h * w, heads=8, d=40,
                      q_ld=3 * c, kv_ld=3 * c)                     # spatial self-attention S=1536 d=40
    t = ops.temporal_attention(qkv, 3, 16, h * w, c, 8)            # temporal attention
    gn = ops.groupnorm(x, g1, b, n, h * w, 32, 16, 1e-5, True)     # 5-D GroupNorm + SiLU
    ln = ops.layernorm(x, g1, b)                                   # LayerNorm
torch.cuda.synchronize()
print("ok")
