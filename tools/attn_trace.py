"""clock64 trace of the first CTA of the persistent attention kernel: where does a 128-key block's time go?
Slots: 0 softmax sees S(g) | 1 softmax hands S back | 9 softmax starts waiting for PV(g-1) | 2 sees PV(g-1) done |
10 P stored | 3 fence done, P signalled | 4 MMA thread sees S free | 5 MMA thread has K(g) (QK issue) | 6 MMA thread
sees P(g) | 7 PV(g) issued | 8 producer got the ring slot for block g."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import lib, ops  # noqa: E402

dev = torch.device("cuda")
n, s, c = 48, 1536, 320
qkv = torch.randn(n * s, 3 * c, device=dev).half()
L = lib.load()
L.ivv_debug_attn_trace.argtypes = [ctypes.c_void_p]
L.ivv_debug_attn_trace.restype = None
buf = torch.zeros(64, 16, dtype=torch.int64, device=dev)
args = dict(n_batch=n, s_q=s, s_kv=s, heads=8, d=40, q_ld=3 * c, kv_ld=3 * c)
for _ in range(3):
    ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], **args)
torch.cuda.synchronize()
L.ivv_debug_attn_trace(buf.data_ptr())
ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], **args)
torch.cuda.synchronize()
L.ivv_debug_attn_trace(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
names = {0: "S seen", 1: "S freed", 9: "wait PV", 2: "PV seen", 10: "P stored", 3: "P signal", 4: "mma:Sfree", 5: "mma:QK",
         6: "mma:P seen", 7: "mma:PV iss", 8: "tma:slot", 11: "out:start", 12: "out:PV ok", 13: "out:done"}
order = [8, 5, 0, 1, 4, 9, 2, 10, 3, 6, 7, 11, 12, 13]
print("block " + " ".join(f"{names[k]:>10s}" for k in order))
for g in range(10, 27):
    print(f"{g:5d} " + " ".join(f"{(int(t[g, k]) - t0) if t[g, k] > 0 else -1:10d}" for k in order))
