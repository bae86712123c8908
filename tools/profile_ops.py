"""Per-op device timing of one eager UNet3D forward (and optionally the VAE decode) at the config-2 shape:
CUDA events around every C-ABI call, aggregated by op and shape. Warm L2, back-to-back launches — the closest view of
what the captured CUDA graph executes. Usage: python tools/profile_ops.py [unet|vae]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from insv2v_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "unet"
dev = torch.device("cuda")
unet, vae = bench.build_models(dev)
unet.use_cuda_graph = False
x = torch.randn(3, 8, bench.FRAMES, bench.LAT_H, bench.LAT_W, device=dev)
ctx = torch.randn(3, 77, 768, device=dev)
t = torch.full((3,), 981.0, device=dev)
z = torch.randn(bench.FRAMES, 4, bench.LAT_H, bench.LAT_W, device=dev)


def run():
    if which == "unet":
        unet(x, t, encoder_hidden_states=ctx)
    else:
        vae.decode(z)


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run()
e1.record()
torch.cuda.synchronize()
print(f"eager {which} pass: {e0.elapsed_time(e1):.2f} ms (includes Python launch overhead)")
ops.Prof.enabled = True
run()
agg = ops.Prof.report()
ops.Prof.enabled = False
tot = sum(v[1] for v in agg.values())
print(f"sum of per-op device times: {tot:.2f} ms over {sum(v[0] for v in agg.values())} calls")
print(f"{'op / shape':64s} {'n':>4s} {'ms':>8s} {'avg us':>8s} {'TFLOP/s':>8s} {'GB/s':>8s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    n, ms, fl, nb = v
    print(f"{str(k)[:64]:64s} {n:4d} {ms:8.3f} {1e3 * ms / n:8.1f} {fl / ms / 1e9:8.1f} {nb / ms / 1e6:8.0f} "
          f"{100 * ms / tot:5.1f}%")
