"""One eager UNet3D forward at a given latent shape under the default kernel selection and with the third-session GEMM
variants switched off (IVV_WIDE=0 IVV_AS=0 IVV_SLAB1=0 IVV_KB2=0), one subprocess each, same seeded weights and inputs: the outputs
must agree to fp16 rounding (same K order, different tiles). Covers shapes that have no CPU golden (48x72 latents).
Usage: python tools/variant_consistency.py [F,H,W ...]      (default 16,48,72 and 16,32,48)"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(shape, path):
    import bench
    dev = torch.device("cuda")
    unet, _ = bench.build_models(dev)
    unet.use_cuda_graph = False
    f, h, w = shape
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 8, f, h, w, generator=g).to(dev)
    ctx = torch.randn(3, 77, 768, generator=g).to(dev)
    t = torch.full((3,), 481.0, device=dev)
    y = unet(x, t, encoder_hidden_states=ctx).sample
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    torch.save(y.float().cpu(), path)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(tuple(int(v) for v in sys.argv[2].split(",")), sys.argv[3])
        sys.exit(0)
    shapes = sys.argv[1:] or ["16,48,72", "16,32,48"]
    off = dict(IVV_WIDE="0", IVV_AS="0", IVV_SLAB1="0", IVV_KB2="0")
    worst = 0.0
    for sh in shapes:
        outs = []
        for tag, env in (("default", {}), ("off", off)):
            path = f"/tmp/variant_{tag}.pt"
            subprocess.run([sys.executable, __file__, "child", sh, path], env=dict(os.environ, **env), check=True)
            outs.append(torch.load(path))
        a, b = outs
        rel = float((a - b).norm() / b.norm())
        worst = max(worst, rel)
        print(f"shape {sh}: rel-L2(default vs variants off) = {rel:.3e}, max abs diff {float((a - b).abs().max()):.3e}, "
              f"|out| mean {float(b.abs().mean()):.3e}")
    assert worst < 1e-3, worst
    print("variant consistency ok")
