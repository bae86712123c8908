"""Time ivv_gemm for a few hot shapes under every tile width (IVV_FORCE_BN) — tile-selection tuning aid."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [(73728, 320, 320, 1), (73728, 320, 960, 1), (73728, 1280, 320, 1), (18432, 640, 640, 1), (18432, 640, 1920, 1),
          (4608, 1280, 1280, 1), (4608, 1280, 3840, 1), (1152, 1280, 1280, 9), (4608, 1280, 1280, 9),
          (73728, 320, 320, 9), (18432, 640, 640, 9), (16384, 512, 512, 9)]


def child():
    from insv2v_b200 import ops
    dev = torch.device("cuda")
    for rows, k, n, taps in SHAPES:
        nbuf = max(2, min(8, int(3e8 // (rows * (k + n) * 2)) + 1))
        xs = [torch.randn(rows, k, device=dev).half() for _ in range(nbuf)]
        outs = [torch.empty(rows, n, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        if taps == 1:
            w = ops.pack_linear(torch.randn(n, k, device=dev) * 0.03)
            geo = dict(n_img=1, h=1, w=rows)
        else:
            w = ops.pack_conv3x3(torch.randn(n, k, 3, 3, device=dev) * 0.01)
            hw = {73728: (48, 32, 48), 18432: (48, 16, 24), 4608: (48, 8, 12), 1152: (48, 4, 6), 16384: (16, 32, 32)}[rows]
            geo = dict(n_img=hw[0], h=hw[1], w=hw[2])
        for i in range(nbuf):
            ops.gemm(xs[i], w, c=k, taps=taps, out=outs[i], **geo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for i in range(reps):
            ops.gemm(xs[i % nbuf], w, c=k, taps=taps, out=outs[i % nbuf], **geo)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        print(f"BN={os.environ.get('IVV_FORCE_BN', 'auto'):>4s} rows={rows:6d} k={k * taps:6d} n={n:5d}: {us:8.1f} us "
              f"{2.0 * rows * k * taps * n / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for pair in ["1", "0"]:
            for bn in ["auto", "128", "160", "256"]:
                env = dict(os.environ, IVV_PAIR=pair)
                if bn != "auto":
                    env["IVV_FORCE_BN"] = bn
                print(f"--- IVV_PAIR={pair}", flush=True)
                subprocess.run([sys.executable, __file__, "child"], env=env)
