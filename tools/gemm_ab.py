"""A/B of the GEMM tuning switches on the hot shapes of the UNet/VAE, one subprocess per setting, rotating buffers so
that operands are not artificially L2-resident. Setting: IVV_HALO (one activation box per filter column for 3x3
convolutions); extra settings can be passed as KEY=VALUE arguments after "ab".
Usage: python tools/gemm_ab.py            (prints one line per shape and setting)"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (frames, h, w, c_in, n_out, taps, residual)
SHAPES = [
    (48, 32, 48, 320, 320, 9, False), (48, 32, 48, 320, 320, 9, True), (48, 32, 48, 640, 320, 9, False),
    (48, 32, 48, 960, 320, 9, False), (48, 32, 48, 640, 640, 9, False),
    (48, 16, 24, 640, 640, 9, False), (48, 16, 24, 1280, 640, 9, True), (48, 16, 24, 1280, 1280, 9, False),
    (16, 32, 48, 512, 512, 9, False), (4, 128, 192, 256, 256, 9, False), (1, 256, 384, 128, 128, 9, False),
    (1, 1, 73728, 320, 320, 1, True), (1, 1, 18432, 640, 640, 1, True), (1, 1, 4608, 1280, 1280, 1, True),
    (1, 1, 73728, 1280, 320, 1, True), (1, 1, 18432, 2560, 640, 1, True), (1, 1, 1152, 1280, 1280, 1, True),
]
# GEMM_AB_ONLY=wide: the shapes the 320-wide pair tiles are meant for (and two they must not slow down)
WIDE_SHAPES = [
    (48, 8, 12, 1280, 1280, 9, True), (48, 8, 12, 2560, 1280, 9, False), (48, 8, 12, 1920, 1280, 9, False),
    (48, 8, 12, 640, 1280, 9, False), (1, 1, 4608, 5120, 1280, 1, True), (1, 1, 18432, 2560, 640, 1, True),
    (1, 1, 73728, 2880, 320, 1, False),
]
SETTINGS = [dict(IVV_HALO="0"), dict(IVV_HALO="1")]
# temporal attention (clips, frames, pixels, channels): the four UNet levels
TATTN = [(3, 16, 1536, 320), (3, 16, 384, 640), (3, 16, 96, 1280), (3, 16, 24, 1280)]


def child():
    from insv2v_b200 import ops
    dev = torch.device("cuda")
    tag = " ".join(f"{k[4:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("IVV_"))
    only = os.environ.get("GEMM_AB_ONLY", "")  # "linear": the 1x1 shapes only, "conv": the 3x3 ones, "none": neither
    for n, h, w, c, n_out, taps, res in (WIDE_SHAPES if only == "wide" else SHAPES):
        if (only == "linear" and taps != 1) or (only == "conv" and taps == 1) or only == "none":
            continue
        rows = n * h * w
        nbuf = max(2, min(8, int(3e8 // (rows * (c + 2 * n_out) * 2)) + 1))
        xs = [torch.randn(rows, c, device=dev).half() for _ in range(nbuf)]
        rs = [torch.randn(rows, n_out, device=dev).half() for _ in range(nbuf)] if res else [None] * nbuf
        outs = [torch.empty(rows, n_out, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        if taps == 1:
            wt = ops.pack_linear(torch.randn(n_out, c, device=dev) * 0.03)
        else:
            wt = ops.pack_conv3x3(torch.randn(n_out, c, 3, 3, device=dev) * 0.01)
        bias = torch.randn(n_out, device=dev).half()

        def call(i):
            ops.gemm(xs[i], wt, n_img=n, h=h, w=w, c=c, taps=taps, bias=bias, residual=rs[i], out=outs[i])

        for i in range(nbuf):
            call(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for i in range(reps):
            call(i % nbuf)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        tf = 2.0 * rows * c * taps * n_out / us / 1e6
        print(f"[{tag:22s}] {n:3d}x{h:3d}x{w:5d} c={c:5d} n={n_out:5d} taps={taps} res={int(res)}: {us:8.1f} us "
              f"{tf:7.1f} TFLOP/s", flush=True)


    for clips, frames, hw, c in ([] if only else TATTN):
        rows = clips * frames * hw
        nbuf = max(2, min(8, int(3e8 // (rows * 4 * c * 2)) + 1))
        qs = [torch.randn(rows, 3 * c, device=dev).half() for _ in range(nbuf)]
        os_ = [torch.empty(rows, c, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        for i in range(nbuf):
            ops.temporal_attention(qs[i], clips, frames, hw, c, 8, out=os_[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            ops.temporal_attention(qs[i % nbuf], clips, frames, hw, c, 8, out=os_[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 40
        print(f"[{tag:22s}] temporal attention rows={rows:6d} c={c:5d}: {us:8.1f} us {8.0 * rows * c / us / 1e3:8.0f} GB/s",
              flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        # extra settings: each argument is a comma-separated KEY=VALUE list, e.g.  IVV_DS=0  IVV_DS=1
        settings = SETTINGS if len(sys.argv) == 1 else [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]]
        for st in settings:
            env = {k: v for k, v in os.environ.items() if k not in st}
            subprocess.run([sys.executable, __file__, "child"], env=dict(env, **st))
