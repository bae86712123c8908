"""A/B timing of the transformer-block linears of one UNet forward (QKV, out-projections with residual, feed-forward
GEGLU / out) under tuning switches, one subprocess per setting, operands rotated over more than L2 can hold.
Usage: python tools/linear_ab.py [KEY=VALUE[,KEY=VALUE...]] ...     e.g.  python tools/linear_ab.py "" IVV_CL4=1
       (IVV_LIB_PATH=<other .so> as a setting times another build of the library)"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, rows, k, n, bias, residual, geglu)
SHAPES = [
    ("qkv0", 73728, 320, 960, False, False, False), ("qkv1", 18432, 640, 1920, False, False, False),
    ("qkv2", 4608, 1280, 3840, False, False, False), ("q0", 73728, 320, 320, False, False, False),
    ("res0", 73728, 320, 320, True, True, False), ("res1", 18432, 640, 640, True, True, False),
    ("res2", 4608, 1280, 1280, True, True, False), ("res3", 1152, 1280, 1280, True, True, False),
    ("ffout0", 73728, 1280, 320, True, True, False), ("ffout1", 18432, 2560, 640, True, True, False),
    ("ffout2", 4608, 5120, 1280, True, True, False),
    ("geglu0", 73728, 320, 2560, True, False, True), ("geglu1", 18432, 640, 5120, True, False, True),
    ("geglu2", 4608, 1280, 10240, True, False, True),
]


def child():
    from insv2v_b200 import ops
    dev = torch.device("cuda")
    tag = " ".join(f"{k[4:]}={os.path.basename(v)}" for k, v in sorted(os.environ.items()) if k.startswith("IVV_"))
    total = 0.0
    for name, rows, k, n, bias, res, geglu in SHAPES:
        n_out = n // 2 if geglu else n
        nbuf = max(2, min(8, int(3e8 // (rows * (k + 2 * n_out) * 2)) + 1))
        xs = [torch.randn(rows, k, device=dev).half() for _ in range(nbuf)]
        rs = [torch.randn(rows, n_out, device=dev).half() for _ in range(nbuf)] if res else [None] * nbuf
        outs = [torch.empty(rows, n_out, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        w = torch.randn(n, k, device=dev) * k ** -0.5
        b = torch.randn(n, device=dev).half()
        if geglu:
            wt, bp = ops.pack_geglu(w, b)
        else:
            wt, bp = ops.pack_linear(w), (b if bias else None)

        def call(i):
            ops.linear(xs[i], wt, bias=bp, residual=rs[i], geglu=geglu, out=outs[i])

        for i in range(nbuf):
            call(i)
        torch.cuda.synchronize()
        ref = (xs[0].float() @ w.half().float().t())
        if geglu:
            ref = ref + b.float()
            ref = ref[:, :n_out] * torch.nn.functional.gelu(ref[:, n_out:])
        else:
            if bias:
                ref = ref + b.float()
            if res:
                ref = ref + rs[0].float()
        err = float((outs[0].float() - ref).norm() / ref.norm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for i in range(reps):
            call(i % nbuf)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        total += us
        tf = 2.0 * rows * k * n / us / 1e6
        print(f"[{tag:28s}] {name:7s} {rows:6d}x{k:5d}->{n:5d}: {us:8.1f} us {tf:7.1f} TFLOP/s  rel-L2 {err:.2e}", flush=True)
    print(f"[{tag:28s}] sum {total:8.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        settings = [dict(kv.split("=") for kv in a.split(",") if kv) for a in (sys.argv[1:] or [""])]
        for st in settings:
            env = {k: v for k, v in os.environ.items() if k not in st}
            try:
                subprocess.run([sys.executable, __file__, "child"], env=dict(env, **st), timeout=240)
            except subprocess.TimeoutExpired:
                print(f"[{st}] TIMEOUT", flush=True)
