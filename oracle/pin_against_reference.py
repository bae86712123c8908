"""TEST INFRASTRUCTURE — pins oracle/insv2v_oracle.py against the reference's own code and mints tests/golden/.

Runs ONLY in the build container (needs /root/reference, which does not travel to the GPU box):
    python oracle/pin_against_reference.py
It imports the reference's modules UNCHANGED from /root/reference (through the diffusers / pytorch_lightning stand-ins in
oracle/shim), loads the seeded weights of oracle.seeded_state_dict into them (strict=True: this is also the proof that
the state-dict schema in tests/golden/schema_*.json is the reference's), runs reference and oracle on the same seeded
inputs, asserts agreement, and writes the reference's outputs as golden fixtures.
"""
import json
import os
import sys
from functools import partial

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("IVV_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from oracle import insv2v_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5


def seeded(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def close(name, a, b, tol=TOL):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    print(f"  {name}: max|ref-oracle| = {err:.3e} (max|ref| = {ref:.3e})")
    assert err <= tol * max(1.0, ref), f"{name}: oracle disagrees with the reference ({err:.3e})"


def build_ref_unet(cfg):
    from modules.video_unet_temporal.unet import UNet3DConditionModel
    kw = {k: v for k, v in cfg.items()}
    kw["motion_module_kwargs"] = dict(cfg["motion_module_kwargs"])
    kw["motion_module_kwargs"]["attention_block_types"] = list(cfg["motion_module_kwargs"]["attention_block_types"])
    return UNet3DConditionModel(**kw).eval()


def build_ref_decoder(cfg):
    import contextlib
    import io
    from modules.vqvae.model import Decoder
    with contextlib.redirect_stdout(io.StringIO()):
        dec = Decoder(**cfg["ddconfig"])
    ed, zc = cfg["embed_dim"], cfg["ddconfig"]["z_channels"]

    class RefVAE(torch.nn.Module):  # AutoencoderKL.decode without Lightning (autoencoder.py:97-100)
        def __init__(self):
            super().__init__()
            self.decoder = dec
            self.post_quant_conv = torch.nn.Conv2d(ed, zc, 1)

        def decode(self, z):
            return self.decoder(self.post_quant_conv(z))
    return RefVAE().eval()


def schema_of(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    meta = {}

    # ---------------- UNet schemas + goldens ----------------
    for tag, cfg in (("micro", O.UNET_CONFIG_MICRO), ("tiny", O.UNET_CONFIG_TINY), ("full", O.UNET_CONFIG_FULL)):
        ref = build_ref_unet(cfg)
        schema = schema_of(ref)
        with open(os.path.join(GOLD, f"schema_unet_{tag}.json"), "w") as f:
            json.dump(schema, f, indent=0, sort_keys=True)
        print(f"[unet {tag}] {len(schema)} tensors, {sum(v.numel() for v in ref.state_dict().values()) / 1e6:.2f} M")
        if tag == "full":
            del ref
            continue
        sd = O.seeded_state_dict(schema, seed=100)
        ref.load_state_dict(sd, strict=True)
        cd = cfg["cross_attention_dim"]
        cases = {"a": dict(shape=(3, 8, 4, 16, 16), t=[981, 981, 981], vsi=0),
                 "b": dict(shape=(1, 8, 2, 12, 20), t=[21], vsi=3)}  # odd size -> forward_upsample_size path
        if tag == "tiny":
            cases = {"a": dict(shape=(3, 8, 4, 16, 24), t=[501, 501, 501], vsi=0)}
        for cname, c in cases.items():
            x = seeded(c["shape"], 1)
            ctx = seeded((c["shape"][0], 77, cd), 2)
            t = torch.tensor(c["t"], dtype=torch.long)
            y_ref = ref(x, t, encoder_hidden_states=ctx, video_start_index=c["vsi"]).sample
            y_or = O.unet3d_forward(sd, cfg, x, t, ctx, video_start_index=c["vsi"])
            close(f"unet {tag}/{cname}", y_or, y_ref)
            torch.save({"out": y_ref.clone(), "shape": c["shape"], "t": c["t"], "vsi": c["vsi"], "weight_seed": 100,
                        "x_seed": 1, "ctx_seed": 2}, os.path.join(GOLD, f"unet_{tag}_{cname}.pt"))
        # pe length guard (motion_module.py:237-240)
        try:
            ref(seeded((1, 8, 4, 16, 16), 1), torch.tensor([1]), encoder_hidden_states=seeded((1, 77, cd), 2),
                video_start_index=30)
            raised = False
        except ValueError:
            raised = True
        print("  reference raises ValueError for video_start_index=30, f=4:", raised)
        meta[f"unet_{tag}_vsi30_raises"] = raised

        if tag == "micro":
            # ---------------- samplers (pl_trainer/inference/inference.py) ----------------
            from pl_trainer.inference.inference import InferenceIP2PVideo, InferenceIP2PVideoOpticalFlow
            steps = 3
            lat = seeded((1, 6, 4, 16, 16), 11)
            cond = seeded((1, 6, 4, 16, 16), 12)
            tc, tu = seeded((1, 77, cd), 13), seeded((1, 77, cd), 14)
            unet_fn = lambda x, t, c: O.unet3d_forward(sd, cfg, x, t, c)  # noqa: E731
            pipe = InferenceIP2PVideo(ref, scheduler="ddim", num_ddim_steps=steps)
            meta["ddim_timesteps_3"] = [int(t) for t in pipe.scheduler.timesteps]
            meta["ddim_timesteps_50"] = [int(t) for t in
                                         InferenceIP2PVideo(ref, scheduler="ddim", num_ddim_steps=50).scheduler.timesteps]
            assert meta["ddim_timesteps_3"] == O.ddim_timesteps(3) and meta["ddim_timesteps_50"] == O.ddim_timesteps(50)
            r1 = pipe(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, text_cfg=7.5, img_cfg=1.5)["latent"]
            o1 = O.sample_ip2p_video(unet_fn, lat, tc, tu, cond, 7.5, 1.5, steps)
            close("sampler first clip", o1, r1, 1e-4)
            lref = seeded((1, 2, 4, 16, 16), 15)
            r2 = pipe.second_clip_forward(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, latent_ref=lref,
                                          noise_correct_step=0.7, text_cfg=7.5, img_cfg=1.5)["latent"]
            o2 = O.sample_ip2p_video(unet_fn, lat, tc, tu, cond, 7.5, 1.5, steps, latent_ref=lref,
                                     noise_correct_step=0.7)
            close("sampler second clip (mean)", o2, r2, 1e-4)
            flows = [seeded((2, 2, 128, 128), 20 + q, 6.0) for q in range(4)]
            pf = InferenceIP2PVideoOpticalFlow.__new__(InferenceIP2PVideoOpticalFlow)
            InferenceIP2PVideo.__init__(pf, ref, scheduler="ddim", num_ddim_steps=steps)  # skip RAFTFlow().cuda()
            pf.obtain_flow_batched = lambda ref_images, query_images: [partial(pf.obtain_delta_noise, flow=fl)
                                                                       for fl in flows]
            r3 = pf.second_clip_forward(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, latent_ref=lref,
                                        ref_images=torch.zeros(1, 2, 3, 8, 8), query_images=torch.zeros(1, 4, 3, 8, 8),
                                        noise_correct_step=0.7, text_cfg=7.5, img_cfg=1.5)["latent"]
            o3 = O.sample_ip2p_video(unet_fn, lat, tc, tu, cond, 7.5, 1.5, steps, latent_ref=lref,
                                     noise_correct_step=0.7, flows=flows)
            close("sampler second clip (flow)", o3, r3, 1e-4)
            torch.save({"first": r1.clone(), "second_mean": r2.clone(), "second_flow": r3.clone(), "steps": steps,
                        "seeds": dict(lat=11, cond=12, tc=13, tu=14, lref=15, flow0=20), "weight_seed": 100,
                        "text_cfg": 7.5, "img_cfg": 1.5, "noise_correct_step": 0.7},
                       os.path.join(GOLD, "sampler_micro.pt"))
        del ref

    # ---------------- VAE decode ----------------
    for tag, cfg in (("tiny", O.VAE_CONFIG_TINY), ("full", O.VAE_CONFIG_FULL)):
        ref = build_ref_decoder(cfg)
        schema = schema_of(ref)
        with open(os.path.join(GOLD, f"schema_vae_{tag}.json"), "w") as f:
            json.dump(schema, f, indent=0, sort_keys=True)
        print(f"[vae {tag}] {len(schema)} tensors, {sum(v.numel() for v in ref.state_dict().values()) / 1e6:.2f} M")
        if tag == "full":
            continue
        sd = O.seeded_state_dict(schema, seed=200)
        ref.load_state_dict(sd, strict=True)
        z = seeded((2, 4, 8, 12), 3)
        y_ref = ref.decode(z)
        close("vae decode tiny", O.vae_decode(sd, cfg, z), y_ref)
        torch.save({"out": y_ref.clone(), "z_shape": (2, 4, 8, 12), "z_seed": 3, "weight_seed": 200},
                   os.path.join(GOLD, "vae_tiny.pt"))

    # ---------------- flow utils ----------------
    from misc_utils.flow_utils import resize_flow, warp_image
    img = seeded((4, 4, 32, 48), 5)
    flow = seeded((4, 2, 32, 48), 6, 6.0)
    big = seeded((4, 2, 256, 384), 7, 5.0)
    w_ref, r_ref = warp_image(img, flow), resize_flow(big, (32, 48))
    r2_ref = resize_flow(big[:, :, :100, :90].contiguous(), (37, 53))
    close("warp_image", O.warp_image(img, flow), w_ref)
    close("resize_flow", O.resize_flow(big, (32, 48)), r_ref)
    w3 = warp_image(img[0], flow[0])  # 3-D inputs are promoted (flow_utils.py:34-37)
    assert w3.shape == (1, 4, 32, 48)
    torch.save({"warp": w_ref.clone(), "resize": r_ref.clone(), "resize_general": r2_ref.clone(),
                "seeds": dict(img=5, flow=6, big=7)}, os.path.join(GOLD, "flow.pt"))

    with open(os.path.join(GOLD, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
