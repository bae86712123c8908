"""clock64 trace of CTA 0 of the persistent GEMM kernel (ivv_debug_gemm_trace): per tile, when did the producer issue its
first / last load, when did the MMA thread get the accumulator and commit the tile, and where did epilogue group 0 spend
its time (top of tile, previous store drained, accumulator seen, residual seen, last chunk done, group barrier passed,
store issued). Usage: python tools/gemm_trace.py [rows k n res]   (default 73728 320 320 1)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import lib, ops  # noqa: E402

dev = torch.device("cuda")
rows, k, n, res = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (73728, 320, 320, 1)))
nbuf = 6
xs = [torch.randn(rows, k, device=dev).half() for _ in range(nbuf)]
rs = [torch.randn(rows, n, device=dev).half() if res else None for _ in range(nbuf)]
outs = [torch.empty(rows, n, device=dev, dtype=torch.float16) for _ in range(nbuf)]
w = ops.pack_linear(torch.randn(n, k, device=dev) * 0.03)
bias = torch.randn(n, device=dev).half()
L = lib.load()
L.ivv_debug_gemm_trace.argtypes = [ctypes.c_void_p]
L.ivv_debug_gemm_trace.restype = None
buf = torch.zeros(32, 16, dtype=torch.int64, device=dev)


LN = os.environ.get("LN", "0") == "1"  # consumer of a folded LayerNorm (+ per-frame rowbias with PE=1)
extra = {}
if LN:
    stats = ops.row_stats(rows, k, dev)
    ops.linear(torch.randn(rows, k, device=dev).half(), ops.pack_linear(torch.randn(k, k, device=dev) * 0.03),
               row_stats_out=stats)
    extra = dict(ln=(stats, torch.randn(n, device=dev).half(), 1e-5))
    if os.environ.get("PE", "0") == "1":
        extra.update(rowbias=torch.randn(16, n, device=dev).half(), rowbias_group=1536, rowbias_mod=16)


def call(i):
    ops.gemm(xs[i], w, n_img=1, h=1, w=rows, c=k, bias=bias, residual=rs[i], out=outs[i], **extra)


for i in range(nbuf):
    call(i)
torch.cuda.synchronize()
L.ivv_debug_gemm_trace(buf.data_ptr())
call(0)
torch.cuda.synchronize()
L.ivv_debug_gemm_trace(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
names = ["tma:first", "tma:last", "mma:acc ok", "mma:commit", "epi:top", "epi:drained", "epi:acc", "epi:res", "epi:chunks",
         "epi:bar", "epi:stored", "c0:bias req", "c0:acc regs", "c0:bias add", "c0:written"]
if os.environ.get("IVV_EPI2", "1") != "0" and n % 160 == 0 and k <= 1280:  # v3 pair kernel (gemm_tc_pair160_kernel)
    names = ["tma:first", "tma:last", "mma:acc ok", "mma:commit", "epi:top", "epi:acc", "epi:regs", "epi:slab ok",
             "epi:written", "st:full", "st:issued", "st:drained", "st:res req", "mma:stage0", "mma:stageN"]
print(f"rows={rows} k={k} n={n} res={res} ln={int(LN)}  (SM clocks since the first stamp of CTA 0)")
print("tile " + " ".join(f"{s:>11s}" for s in names))
for g in range(32):
    if not (t[g] > 0).any():
        break
    print(f"{g:4d} " + " ".join(f"{(int(t[g, j]) - t0) if t[g, j] > 0 else -1:11d}" for j in range(len(names))))
