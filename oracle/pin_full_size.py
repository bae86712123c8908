"""TEST INFRASTRUCTURE — goldens at the BENCHMARKED shapes, minted from the reference's own classes.

Runs ONLY in the build container (needs /root/reference; takes ~40 min of CPU on 8 cores):
    python oracle/pin_full_size.py [unet] [vae] [encoder] [flow] [c1] [ddpm]
Every section builds the reference's module unchanged (through oracle/shim), loads oracle.seeded_state_dict weights
(strict=True), runs it in fp32 on seeded inputs, asserts that the oracle restatement agrees, and writes the REFERENCE's
output to tests/golden/*_full*.pt:

  unet    UNet3DConditionModel.forward at BASELINE configs[1] shape [3,8,16,32,48] (unet.py:296-434)
  vae     AutoencoderKL.decode, full ddconfig (ch 128, mult 1,2,4,4), z [2,4,32,48] -> [2,3,256,384]
  encoder AutoencoderKL.encode moments: Encoder.forward + quant_conv (vqvae/model.py:211-302, autoencoder.py:89-95),
          tiny config [2,3,64,96] and full config [1,3,128,192]
  flow    InferenceIP2PVideoOpticalFlow.second_clip_forward, 3 DDIM steps at [1,16,4,32,48], R=4, 12 synthetic flows
          [4,2,256,384] (configs[2]; inference.py:313-398)
  c1      configs[0]: 8-frame 256x256 clip, DDIM-20, InferenceIP2PVideo.__call__ + per-frame decode
          (instruct_p2p_video.py:66-79)
  ddpm    scheduler='ddpm' (the entry script's choice, insv2v_run_loveu_tgve.py:64-74), guidance_rescale, start_time
          and the all_latent / all_pred lists on the micro UNet
"""
import os
import sys
import time
from functools import partial

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import pin_against_reference as P  # noqa: E402  (sets sys.path for the shim + reference)

O = P.O
GOLD = P.GOLD
seeded = P.seeded


def _schema(name):
    import json
    return {k: tuple(v) for k, v in json.load(open(os.path.join(GOLD, f"schema_{name}.json"))).items()}


def build_ref_encoder(cfg):
    import contextlib
    import io
    from modules.vqvae.model import Encoder
    with contextlib.redirect_stdout(io.StringIO()):
        enc = Encoder(**cfg["ddconfig"])
    ed, zc = cfg["embed_dim"], cfg["ddconfig"]["z_channels"]

    class RefEnc(torch.nn.Module):  # AutoencoderKL.encode up to the moments (autoencoder.py:89-92)
        def __init__(self):
            super().__init__()
            self.encoder = enc
            self.quant_conv = torch.nn.Conv2d(2 * zc, 2 * ed, 1)

        def moments(self, x):
            return self.quant_conv(self.encoder(x))
    return RefEnc().eval()


def pin_unet():
    cfg = O.UNET_CONFIG_FULL
    ref = P.build_ref_unet(cfg)
    sd = O.seeded_state_dict(_schema("unet_full"), seed=7)
    ref.load_state_dict(sd, strict=True)
    shape = (3, 8, 16, 32, 48)
    x, ctx = seeded(shape, 31), seeded((3, 77, 768), 32)
    t = torch.tensor([981, 981, 981], dtype=torch.long)
    t0 = time.time()
    y_ref = ref(x, t, encoder_hidden_states=ctx).sample
    print(f"  reference UNet3D forward {shape}: {time.time() - t0:.1f}s")
    y_or = O.unet3d_forward(sd, cfg, x, t, ctx)
    P.close("unet full c2", y_or, y_ref)
    torch.save({"out": y_ref.clone(), "shape": shape, "t": [981] * 3, "vsi": 0, "weight_seed": 7, "x_seed": 31,
                "ctx_seed": 32}, os.path.join(GOLD, "unet_full_c2.pt"))


def pin_vae():
    cfg = O.VAE_CONFIG_FULL
    ref = P.build_ref_decoder(cfg)
    sd_all = O.seeded_state_dict(_schema("vae_full"), seed=201)
    ref.load_state_dict(sd_all, strict=True)
    z = seeded((2, 4, 32, 48), 33)
    y_ref = ref.decode(z)
    P.close("vae decode full", O.vae_decode(sd_all, cfg, z), y_ref)
    torch.save({"out": y_ref.clone(), "z_shape": (2, 4, 32, 48), "z_seed": 33, "weight_seed": 201},
               os.path.join(GOLD, "vae_full_decode.pt"))


def pin_encoder():
    import json
    for tag, cfg, xshape in (("tiny", O.VAE_CONFIG_TINY, (2, 3, 64, 96)), ("full", O.VAE_CONFIG_FULL, (1, 3, 128, 192))):
        ref = build_ref_encoder(cfg)
        schema = P.schema_of(ref)
        with open(os.path.join(GOLD, f"schema_vae_encoder_{tag}.json"), "w") as f:
            json.dump(schema, f, indent=0, sort_keys=True)
        sd = O.seeded_state_dict(schema, seed=300)
        ref.load_state_dict(sd, strict=True)
        x = seeded(xshape, 34)
        m_ref = ref.moments(x)
        P.close(f"vae encode moments {tag}", O.vae_encode_moments(sd, cfg, x), m_ref)
        torch.save({"moments": m_ref.clone(), "x_shape": xshape, "x_seed": 34, "weight_seed": 300},
                   os.path.join(GOLD, f"vae_encoder_{tag}.pt"))


def _flow_pipe(ref, steps, flows, scheduler="ddim"):
    from pl_trainer.inference.inference import InferenceIP2PVideo, InferenceIP2PVideoOpticalFlow
    pf = InferenceIP2PVideoOpticalFlow.__new__(InferenceIP2PVideoOpticalFlow)
    InferenceIP2PVideo.__init__(pf, ref, scheduler=scheduler, num_ddim_steps=steps)  # skip RAFTFlow().cuda()
    pf.obtain_flow_batched = lambda ref_images, query_images: [partial(pf.obtain_delta_noise, flow=fl) for fl in flows]
    return pf


def pin_flow():
    cfg = O.UNET_CONFIG_FULL
    ref = P.build_ref_unet(cfg)
    sd = O.seeded_state_dict(_schema("unet_full"), seed=7)
    ref.load_state_dict(sd, strict=True)
    steps, ncs = 3, 0.5
    lat, cond = seeded((1, 16, 4, 32, 48), 41), seeded((1, 16, 4, 32, 48), 42)
    tc, tu = seeded((1, 77, 768), 43), seeded((1, 77, 768), 44)
    lref = seeded((1, 4, 4, 32, 48), 45)
    flows = [seeded((4, 2, 256, 384), 50 + q, 5.0) for q in range(12)]
    pf = _flow_pipe(ref, steps, flows)
    t0 = time.time()
    out = pf.second_clip_forward(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, latent_ref=lref,
                                 ref_images=torch.zeros(1, 4, 3, 8, 8), query_images=torch.zeros(1, 12, 3, 8, 8),
                                 noise_correct_step=ncs, text_cfg=7.5, img_cfg=1.5)
    print(f"  reference flow sampler, {steps} steps at full size: {time.time() - t0:.1f}s")
    torch.save({"latent": out["latent"].clone(), "all_latent": [a.clone() for a in out["all_latent"]],
                "steps": steps, "noise_correct_step": ncs, "text_cfg": 7.5, "img_cfg": 1.5, "weight_seed": 7,
                "seeds": dict(lat=41, cond=42, tc=43, tu=44, lref=45, flow0=50)},
               os.path.join(GOLD, "sampler_full_flow.pt"))


def pin_c1():
    from pl_trainer.inference.inference import InferenceIP2PVideo
    cfg = O.UNET_CONFIG_FULL
    ref = P.build_ref_unet(cfg)
    sd = O.seeded_state_dict(_schema("unet_full"), seed=7)
    ref.load_state_dict(sd, strict=True)
    vcfg = O.VAE_CONFIG_FULL
    vae = P.build_ref_decoder(vcfg)
    vsd = O.seeded_state_dict(_schema("vae_full"), seed=201)
    vae.load_state_dict(vsd, strict=True)
    steps = 20
    lat, cond = seeded((1, 8, 4, 32, 32), 61), seeded((1, 8, 4, 32, 32), 62)
    tc, tu = seeded((1, 77, 768), 63), seeded((1, 77, 768), 64)
    pipe = InferenceIP2PVideo(ref, scheduler="ddim", num_ddim_steps=steps)
    t0 = time.time()
    out = pipe(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, text_cfg=7.5, img_cfg=1.5)
    t_s = time.time() - t0
    # decode_latent_to_image (instruct_p2p_video.py:66-79): per frame, latent / 0.18215 first (diffusion.py:247-249)
    t0 = time.time()
    frames = torch.stack([vae.decode(out["latent"][:, i] / 0.18215) for i in range(lat.shape[1])], dim=1)
    t_d = time.time() - t0
    print(f"  reference configs[0]: DDIM-20 {t_s:.1f}s + 8 frame decodes {t_d:.1f}s on {torch.get_num_threads()} threads")
    torch.save({"latent": out["latent"].clone(), "all_latent": [a.clone() for a in out["all_latent"]],
                "frames": frames.to(torch.float16), "steps": steps, "text_cfg": 7.5, "img_cfg": 1.5,
                "unet_seed": 7, "vae_seed": 201, "seeds": dict(lat=61, cond=62, tc=63, tu=64),
                "cpu_seconds": dict(sampler=t_s, decode=t_d, threads=torch.get_num_threads())},
               os.path.join(GOLD, "c1_e2e.pt"))


def pin_ddpm():
    from pl_trainer.inference.inference import InferenceIP2PVideo
    cfg = O.UNET_CONFIG_MICRO
    ref = P.build_ref_unet(cfg)
    sd = O.seeded_state_dict(_schema("unet_micro"), seed=100)
    ref.load_state_dict(sd, strict=True)
    cd = cfg["cross_attention_dim"]
    steps = 4
    lat, cond = seeded((1, 6, 4, 16, 16), 11), seeded((1, 6, 4, 16, 16), 12)
    tc, tu = seeded((1, 77, cd), 13), seeded((1, 77, cd), 14)
    lref = seeded((1, 2, 4, 16, 16), 15)
    flows = [seeded((2, 2, 128, 128), 20 + q, 6.0) for q in range(4)]
    unet_fn = lambda x, t, c: O.unet3d_forward(sd, cfg, x, t, c)  # noqa: E731
    gold = {"steps": steps, "seeds": dict(lat=11, cond=12, tc=13, tu=14, lref=15, flow0=20), "weight_seed": 100,
            "text_cfg": 7.5, "img_cfg": 1.5, "noise_seed": 77}
    pipe = InferenceIP2PVideo(ref, scheduler="ddpm", num_ddim_steps=steps)
    gold["ddpm_timesteps_4"] = [int(t) for t in pipe.scheduler.timesteps]
    gold["ddpm_timesteps_20"] = [int(t) for t in
                                 InferenceIP2PVideo(ref, scheduler="ddpm", num_ddim_steps=20).scheduler.timesteps]
    assert gold["ddpm_timesteps_20"] == O.ddpm_timesteps(20)

    def run(fn, **kw):
        torch.manual_seed(77)  # DDPMScheduler.step draws its variance noise from the global CPU generator
        out = fn(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, text_cfg=7.5, img_cfg=1.5, **kw)
        return {"latent": out["latent"].clone(), "all_latent": [a.clone() for a in out["all_latent"]],
                "all_pred": [a.clone() for a in out["all_pred"]]}

    def orc(**kw):
        torch.manual_seed(77)
        return O.sample_ip2p_video(unet_fn, lat, tc, tu, cond, 7.5, 1.5, steps, return_all=True, **kw)

    gold["ddpm_first"] = run(pipe)
    o = orc(scheduler="ddpm")
    P.close("ddpm first clip", o["latent"], gold["ddpm_first"]["latent"], 1e-4)
    P.close("ddpm first clip all_pred[1]", o["all_pred"][1], gold["ddpm_first"]["all_pred"][1], 1e-4)
    gold["ddpm_rescale_start1"] = run(pipe, guidance_rescale=0.7, start_time=1)
    o = orc(scheduler="ddpm", guidance_rescale=0.7, start_time=1)
    P.close("ddpm guidance_rescale start_time=1", o["latent"], gold["ddpm_rescale_start1"]["latent"], 1e-4)
    gold["ddpm_second_mean"] = run(pipe.second_clip_forward, latent_ref=lref, noise_correct_step=0.5)
    o = orc(scheduler="ddpm", latent_ref=lref, noise_correct_step=0.5)
    P.close("ddpm second clip (mean)", o["latent"], gold["ddpm_second_mean"]["latent"], 1e-4)
    pf = _flow_pipe(ref, steps, flows, scheduler="ddpm")
    gold["ddpm_second_flow"] = run(pf.second_clip_forward, latent_ref=lref, noise_correct_step=0.5,
                                   ref_images=torch.zeros(1, 2, 3, 8, 8), query_images=torch.zeros(1, 4, 3, 8, 8),
                                   guidance_rescale=0.3)
    o = orc(scheduler="ddpm", latent_ref=lref, noise_correct_step=0.5, flows=flows, guidance_rescale=0.3)
    P.close("ddpm second clip (flow, rescale)", o["latent"], gold["ddpm_second_flow"]["latent"], 1e-4)
    pd = InferenceIP2PVideo(ref, scheduler="ddim", num_ddim_steps=steps)
    gold["ddim_rescale_start2"] = run(pd, guidance_rescale=0.5, start_time=2)
    o = orc(scheduler="ddim", guidance_rescale=0.5, start_time=2)
    P.close("ddim guidance_rescale start_time=2", o["latent"], gold["ddim_rescale_start2"]["latent"], 1e-4)
    torch.save(gold, os.path.join(GOLD, "sampler_micro_ddpm.pt"))


def main():
    torch.set_grad_enabled(False)
    what = sys.argv[1:] or ["ddpm", "encoder", "vae", "unet", "flow", "c1"]
    for w in what:
        print(f"[{w}]")
        t0 = time.time()
        globals()[f"pin_{w}"]()
        print(f"[{w}] done in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
