"""One eager (non-graph) UNet3D forward at the config-2 shape + one 16-frame VAE decode: the command profiled by ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda")
unet, vae = bench.build_models(dev)
unet.use_cuda_graph = False
F_, H_, W_ = [int(v) for v in os.environ.get("SHAPE", f"{bench.FRAMES},{bench.LAT_H},{bench.LAT_W}").split(",")]
x = torch.randn(3, 8, F_, H_, W_, device=dev)
ctx = torch.randn(3, 77, 768, device=dev)
t = torch.full((3,), 981.0, device=dev)
reps = int(os.environ.get("REPS", "1"))
for _ in range(reps):
    y = unet(x, t, encoder_hidden_states=ctx).sample
if os.environ.get("DECODE", "1") == "1":
    img = vae.decode(torch.randn(F_, 4, H_, W_, device=dev))
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
