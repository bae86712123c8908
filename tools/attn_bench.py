"""ivv_attention timing for the UNet's attention shapes under every kernel variant (env switches), CUDA-event timed.
Each variant runs in its own subprocess (the switches are read per call, but a hung variant must not take the rest)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# (n_batch, s_q, s_kv, heads, d, kv_div)
SHAPES = [(48, 1536, 1536, 8, 40, 1), (48, 1536, 77, 8, 40, 16), (48, 384, 384, 8, 80, 1), (48, 96, 96, 8, 160, 1),
          (48, 384, 77, 8, 80, 16), (48, 96, 77, 8, 160, 16), (48, 24, 24, 8, 160, 1)]
VARIANTS = [
    {"IVV_ATTN_PAIR": "0"},
    {"IVV_ATTN_PAIR": "1", "IVV_ATTN_MODE": "3"},
    {"IVV_ATTN_PAIR": "1", "IVV_ATTN_MODE": "3", "IVV_ATTN_PAIR_SHORT": "1"},
    {"IVV_ATTN_PAIR": "1", "IVV_ATTN_MODE": "3", "IVV_ATTN_POLY": "1"},
]
if os.environ.get("ATTN_BENCH_ONE"):
    SHAPES = SHAPES[:2]
if os.environ.get("ATTN_BENCH_SMALL"):  # the d = 80 / 160 shapes only, default vs the persistent one-tile kernel
    SHAPES = SHAPES[2:]
    VARIANTS = [{}, {"IVV_ATTN_PERSIST1": "1"}]


def child():
    from insv2v_b200 import ops
    dev = torch.device("cuda")
    tag = " ".join(f"{k[9:]}={v}" for k, v in os.environ.items() if k.startswith("IVV_ATTN_"))
    for n, sq, skv, heads, d, kv_div in SHAPES:
        c = heads * d
        q = torch.randn(n * sq, c, device=dev).half()
        kv = torch.randn(n // kv_div * skv, 2 * c, device=dev).half()
        args = dict(n_batch=n, s_q=sq, s_kv=skv, heads=heads, d=d, q_ld=c, kv_ld=2 * c, kv_div=kv_div)
        for _ in range(5):
            out = ops.attention(q, kv[:, :c], kv[:, c:], **args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            ops.attention(q, kv[:, :c], kv[:, c:], out=out, **args)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        fl = 4.0 * n * heads * sq * skv * d
        # accuracy against fp32 SDPA on a slice of the batch
        nn = min(n, kv_div * 2)
        qf = q[:nn * sq].float().reshape(nn, sq, heads, d).transpose(1, 2)
        kf, vf = (t.float().reshape(-1, skv, heads, d).transpose(1, 2).repeat_interleave(kv_div, dim=0)[:nn]
                  for t in (kv[:, :c], kv[:, c:]))
        ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1) @ vf).transpose(1, 2).reshape(nn * sq, c)
        err = (out[:nn * sq].float() - ref).abs().max().item()
        print(f"[{tag}] n={n} sq={sq} skv={skv} d={d}: {us:8.1f} us  {fl / us * 1e-6:7.1f} TFLOP/s  max|err|={err:.2e}",
              flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for v in VARIANTS:
            try:
                subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **v), timeout=180)
            except subprocess.TimeoutExpired:
                print("TIMEOUT", v, flush=True)
