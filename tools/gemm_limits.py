"""Where does the GEMM time go? Same launches with (a) everything, (b) no MMA issue (TMA + epilogue only),
(c) no TMA loads (MMA + epilogue only). Tuning experiment; outputs of (b)/(c) are garbage by construction."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(73728, 320, 320, 9), (73728, 320, 320, 1), (73728, 320, 960, 1), (18432, 640, 640, 1), (4608, 1280, 1280, 1),
          (4608, 1280, 3840, 1)]


def child():
    from insv2v_b200 import ops
    dev = torch.device("cuda")
    for rows, k, n, taps in SHAPES:
        x = torch.randn(rows, k, device=dev).half()
        out = torch.empty(rows, n, device=dev, dtype=torch.float16)
        if taps == 1:
            w = ops.pack_linear(torch.randn(n, k, device=dev) * 0.03)
            geo = dict(n_img=1, h=1, w=rows)
        else:
            w = ops.pack_conv3x3(torch.randn(n, k, 3, 3, device=dev) * 0.01)
            hw = {73728: (48, 32, 48), 18432: (48, 16, 24), 4608: (48, 8, 12), 16384: (16, 32, 32)}[rows]
            geo = dict(n_img=hw[0], h=hw[1], w=hw[2])
        for _ in range(3):
            ops.gemm(x, w, c=k, taps=taps, out=out, **geo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            ops.gemm(x, w, c=k, taps=taps, out=out, **geo)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 30
        print(f"skip={os.environ.get('IVV_DEBUG_SKIP', '0')} pair={os.environ.get('IVV_PAIR', '1')} rows={rows:6d} "
              f"k={k * taps:6d} n={n:5d}: {us:8.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for pair in ["1"]:
            for skip in ["0", "1", "2", "3"]:
                subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, IVV_PAIR=pair, IVV_DEBUG_SKIP=skip))
