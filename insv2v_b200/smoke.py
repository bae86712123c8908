"""__graft_entry__.smoke(): one small pass of the hot path on cuda:0 (UNet3D forward through the C-ABI kernels, one
DDIM step, VAE decode of the result, one flow-warp) checked against the CPU oracle."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import insv2v_oracle as O  # checker only
    from insv2v_b200 import lib
    from insv2v_b200.flow_utils import warp_image
    from insv2v_b200.unet import UNet3DConditionModel
    from insv2v_b200.vae import AutoencoderKL

    lib.load()
    torch.cuda.set_device(0)
    gold = os.path.join(ROOT, "tests", "golden")
    cfg = O.UNET_CONFIG_MICRO
    schema = {k: tuple(v) for k, v in json.load(open(os.path.join(gold, "schema_unet_micro.json"))).items()}
    sd = O.seeded_state_dict(schema, seed=100)
    unet = UNet3DConditionModel(**cfg)
    unet.load_state_dict(sd, strict=True)
    unet = unet.cuda().eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 8, 4, 16, 16, generator=g)
    ctx = torch.randn(3, 77, cfg["cross_attention_dim"], generator=g)
    t = torch.tensor([981, 981, 981])
    y = unet(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda()).sample.cpu()
    with torch.no_grad():
        ref = O.unet3d_forward(sd, cfg, x, t, ctx)
    rel = ((y - ref).norm() / ref.norm()).item()
    print(f"[smoke] unet micro: rel_l2 vs oracle = {rel:.3e}")
    assert rel < 1e-2, rel

    vschema = {k: tuple(v) for k, v in json.load(open(os.path.join(gold, "schema_vae_tiny.json"))).items()}
    vsd = O.seeded_state_dict(vschema, seed=200)
    vae = AutoencoderKL(**O.VAE_CONFIG_TINY, lossconfig=None)
    vae.load_state_dict(vsd, strict=False)
    vae = vae.cuda().eval()
    z = y[0].permute(1, 0, 2, 3).contiguous()  # 4 frames [4, 4, 16, 16]
    img = vae.decode(z.cuda()).cpu()
    with torch.no_grad():
        iref = O.vae_decode(vsd, O.VAE_CONFIG_TINY, z)
    rel = ((img - iref).norm() / iref.norm()).item()
    print(f"[smoke] vae decode tiny: rel_l2 vs oracle = {rel:.3e}")
    assert rel < 1e-2, rel

    flow = torch.randn(4, 2, 16, 16, generator=g) * 3
    wv = warp_image(z.cuda(), flow.cuda()).cpu()
    err = (wv - O.warp_image(z, flow)).abs().max().item()
    print(f"[smoke] warp_image: max abs err vs oracle = {err:.3e}")
    assert err < 1e-4, err
    print(f"[smoke] ok — {lib.LAUNCH_COUNT} C-ABI kernel launches from {lib.LIB_PATH}")
