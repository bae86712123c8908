"""A/B timing of GroupNorm(+SiLU) on the UNet's shapes (rows, channels, frames per group), operands rotated over more
than L2 holds; one subprocess per setting.  Usage: python tools/gn_ab.py "" IVV_GN_FUSED=1"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (n_img, hw, c, frames_per_group)
SHAPES = [(48, 1536, 320, 16), (48, 384, 640, 1), (48, 384, 640, 16), (48, 96, 1280, 1), (48, 96, 1280, 16),
          (48, 24, 1280, 1), (48, 24, 1280, 16), (48, 96, 2560, 16), (48, 24, 2560, 16)]


def child():
    from insv2v_b200 import ops
    dev = torch.device("cuda")
    tag = " ".join(f"{k[4:]}={os.path.basename(v)}" for k, v in sorted(os.environ.items()) if k.startswith("IVV_"))
    tot = 0.0
    for n, hw, c, fpg in SHAPES:
        rows = n * hw
        nbuf = max(2, min(16, int(3e8 // (rows * c * 4)) + 1))
        xs = [torch.randn(rows, c, device=dev).half() for _ in range(nbuf)]
        outs = [torch.empty_like(xs[0]) for _ in range(nbuf)]
        g, b = torch.randn(c, device=dev).half(), torch.randn(c, device=dev).half()
        for i in range(nbuf):
            ops.groupnorm(xs[i], g, b, n, hw, 32, fpg, 1e-5, True, out=outs[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 60
        e0.record()
        for i in range(reps):
            ops.groupnorm(xs[i % nbuf], g, b, n, hw, 32, fpg, 1e-5, True, out=outs[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        tot += us
        print(f"[{tag:16s}] groupnorm rows={rows:6d} c={c:5d} fpg={fpg:2d}: {us:7.1f} us {4.0 * rows * c / us / 1e3:7.0f} GB/s",
              flush=True)
    print(f"[{tag:16s}] sum {tot:7.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for a in (sys.argv[1:] or [""]):
            st = dict(kv.split("=") for kv in a.split(",") if kv)
            env = {k: v for k, v in os.environ.items() if k not in st}
            try:
                subprocess.run([sys.executable, __file__, "child"], env=dict(env, **st), timeout=200)
            except subprocess.TimeoutExpired:
                print(f"[{st}] TIMEOUT", flush=True)
