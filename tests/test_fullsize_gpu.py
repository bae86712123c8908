"""Parity AT THE BENCHMARKED SHAPES and tolerance calibration.

Golden outputs come from the REFERENCE's own classes run in fp32 on the CPU at full size (oracle/pin_full_size.py):
  * UNet3DConditionModel.forward at BASELINE configs[1] shape [3,8,16,32,48]
  * AutoencoderKL.decode with the real ddconfig, 32x48 latents -> 256x384 frames
  * Encoder + quant_conv moments (tiny and real ddconfig)
  * InferenceIP2PVideoOpticalFlow.second_clip_forward, 3 DDIM steps at [1,16,4,32,48] with 12x4 synthetic flows
    (configs[2])
  * configs[0]: 8-frame 256x256 DDIM-20 end to end (sampler + per-frame decode)

Calibration: the product computes in fp16 storage / fp32 accumulation, the golden is fp32. What that precision regime
costs is measured, not assumed: the oracle (= the reference's arithmetic, pinned) is run on the SAME B200 under
`torch.autocast(float16)` with PyTorch SDPA — the reference's own fp16 path (gradio_demo.py:98,
pl_trainer/instruct_p2p_video.py:31-66) — against the same fp32 golden, and the product must stay within 1.5x of that
error (and under an absolute gate). Measured values are written to gpurun_out/calibration.json.
"""
import json
import os
import time

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import ROOT, err_stats, golden, schema, seeded

pytestmark = pytest.mark.gpu

UNET_REL_L2 = 4e-3     # absolute gate, whole UNet forward (round 1 asserted 1e-2; measured 2.2-2.7e-3)
VAE_REL_L2 = 3e-3
RATIO = 1.5            # product error <= RATIO x error of the reference's own fp16-autocast path
_CAL = {}


def _record(name, **kw):
    _CAL[name] = kw
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "calibration.json"), "w") as f:
            json.dump(_CAL, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _O():
    from oracle import insv2v_oracle as O
    return O


def _sdpa_core(q, k, v, heads):
    """The reference's fp16 attention core: diffusers AttnProcessor2_0 = F.scaled_dot_product_attention."""
    b, sq, c = q.shape
    d = c // heads
    qh, kh, vh = (t.reshape(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    return F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, sq, c)


class _RefFp16:
    """The oracle's functions on the GPU under fp16 autocast + SDPA (the reference's own mixed-precision path)."""

    def __init__(self):
        self.O = _O()

    def __enter__(self):
        self.saved = self.O.attention_core
        self.O.attention_core = _sdpa_core
        self.ac = torch.autocast("cuda", dtype=torch.float16)
        self.ac.__enter__()
        return self.O

    def __exit__(self, *a):
        self.ac.__exit__(*a)
        self.O.attention_core = self.saved


def _gate(name, got, ref_fp16, gold, abs_gate):
    e, r = err_stats(got, gold), err_stats(ref_fp16, gold)
    print(f"[{name}] product rel_l2={e['rel_l2']:.3e} max_abs={e['max_abs']:.3e} | reference fp16-autocast "
          f"rel_l2={r['rel_l2']:.3e} max_abs={r['max_abs']:.3e} | max|gold|={e['max_ref']:.3e}")
    _record(name, product_rel_l2=e["rel_l2"], product_max_abs=e["max_abs"], ref_fp16_rel_l2=r["rel_l2"],
            ref_fp16_max_abs=r["max_abs"], max_ref=e["max_ref"], abs_gate=abs_gate, ratio_gate=RATIO)
    assert torch.isfinite(got).all()
    assert e["rel_l2"] <= abs_gate, f"{name}: rel-L2 {e['rel_l2']:.3e} > {abs_gate}"
    assert e["rel_l2"] <= RATIO * r["rel_l2"] + 2e-4, \
        f"{name}: product error {e['rel_l2']:.3e} > {RATIO} x reference-fp16 error {r['rel_l2']:.3e}"
    assert e["max_abs"] <= RATIO * r["max_abs"] + 2e-3 * e["max_ref"], \
        f"{name}: product max error {e['max_abs']:.3e} vs reference-fp16 {r['max_abs']:.3e}"


@pytest.fixture(scope="module")
def full_unet():
    from insv2v_b200.unet import UNet3DConditionModel
    O = _O()
    sd = O.seeded_state_dict(schema("unet_full"), seed=7)
    m = UNet3DConditionModel(**O.UNET_CONFIG_FULL)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    return m, sd_gpu


@pytest.fixture(scope="module")
def full_vae():
    from insv2v_b200.vae import AutoencoderKL
    O = _O()
    sd = O.seeded_state_dict(schema("vae_full"), seed=201)
    vae = AutoencoderKL(**O.VAE_CONFIG_FULL, lossconfig=None)
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "quant_conv.")) for k in missing)
    return vae.cuda().eval(), {k: v.cuda() for k, v in sd.items()}


def test_unet_config2_shape_vs_reference_golden(full_unet):
    """UNet3DConditionModel.forward at [3,8,16,32,48] (the shape bench.py times) vs the reference's own class."""
    m, sd = full_unet
    O = _O()
    g = golden("unet_full_c2.pt")
    x, ctx = seeded(g["shape"], g["x_seed"]).cuda(), seeded((3, 77, 768), g["ctx_seed"]).cuda()
    t = torch.tensor(g["t"], device="cuda")
    y = m(x, t, encoder_hidden_states=ctx).sample
    assert y.shape == g["out"].shape and y.dtype == torch.float32
    with torch.no_grad(), _RefFp16() as Oh:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y16 = Oh.unet3d_forward(sd, O.UNET_CONFIG_FULL, x, t, ctx).float()
        torch.cuda.synchronize()
        t_ref = time.perf_counter() - t0
        t0 = time.perf_counter()
        Oh.unet3d_forward(sd, O.UNET_CONFIG_FULL, x, t, ctx)
        torch.cuda.synchronize()
        t_ref = min(t_ref, time.perf_counter() - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m(x, t, encoder_hidden_states=ctx)
    e1.record()
    torch.cuda.synchronize()
    _gate("unet [3,8,16,32,48]", y, y16, g["out"], UNET_REL_L2)
    _CAL["unet [3,8,16,32,48]"].update(pytorch_gpu_fp16_autocast_ms=1e3 * t_ref, product_ms=e0.elapsed_time(e1) / 3)
    _record("unet [3,8,16,32,48]", **_CAL["unet [3,8,16,32,48]"])
    print(f"  PyTorch fp16 autocast + SDPA on this GPU: {1e3 * t_ref:.1f} ms / forward; product "
          f"{e0.elapsed_time(e1) / 3:.1f} ms")


def test_vae_full_config_decode_vs_reference_golden(full_vae):
    """AutoencoderKL.decode, real ddconfig (ch 128, mult 1,2,4,4; d=512 S=1536 attention), 32x48 -> 256x384."""
    vae, sd = full_vae
    O = _O()
    g = golden("vae_full_decode.pt")
    z = seeded(g["z_shape"], g["z_seed"]).cuda()
    y = vae.decode(z)
    assert y.shape == g["out"].shape
    with torch.no_grad(), _RefFp16() as Oh:
        y16 = Oh.vae_decode(sd, O.VAE_CONFIG_FULL, z).float()
    _gate("vae decode 256x384", y, y16, g["out"], VAE_REL_L2)
    # the bench decodes 16 frames in one batch: per-frame GroupNorm / attention make batching exact per frame
    z16 = torch.cat([z, seeded((14, 4, 32, 48), 99).cuda()], dim=0)
    y_b = vae.decode(z16)
    e = err_stats(y_b[:2], g["out"])
    print(f"[vae decode 16-frame batch, frames 0-1] rel_l2={e['rel_l2']:.3e}")
    assert e["rel_l2"] <= VAE_REL_L2 and torch.isfinite(y_b).all()


@pytest.mark.parametrize("tag", ["tiny", "full"])
def test_vae_encode_moments_vs_reference_golden(tag):
    """Encoder.forward + quant_conv (vqvae/model.py:275-302, autoencoder.py:89-91) vs the reference's Encoder."""
    from insv2v_b200.vae import AutoencoderKL
    O = _O()
    cfg = {"tiny": O.VAE_CONFIG_TINY, "full": O.VAE_CONFIG_FULL}[tag]
    g = golden(f"vae_encoder_{tag}.pt")
    sd = O.seeded_state_dict(schema(f"vae_encoder_{tag}"), seed=g["weight_seed"])
    vae = AutoencoderKL(**cfg, lossconfig=None)
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("decoder.", "post_quant_conv.")) for k in missing)
    vae = vae.cuda().eval()
    x = seeded(g["x_shape"], g["x_seed"]).cuda()
    got = vae.encode_moments(x)
    sdg = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad(), _RefFp16() as Oh:
        m16 = Oh.vae_encode_moments(sdg, cfg, x).float()
    _gate(f"vae encode moments {tag}", got, m16, g["moments"], VAE_REL_L2)
    # encode() = mean + std * randn drawn on the CPU generator and moved to the device (autoencoder.py:22)
    torch.manual_seed(5)
    zs = vae.encode(x)
    torch.manual_seed(5)
    mean, logvar = torch.chunk(got, 2, dim=1)
    want = mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * torch.randn(mean.shape).cuda()
    assert torch.allclose(zs, want, rtol=0, atol=1e-6)


def _flow_inputs(g):
    s = g["seeds"]
    lat, cond = seeded((1, 16, 4, 32, 48), s["lat"]).cuda(), seeded((1, 16, 4, 32, 48), s["cond"]).cuda()
    tc, tu = seeded((1, 77, 768), s["tc"]).cuda(), seeded((1, 77, 768), s["tu"]).cuda()
    lref = seeded((1, 4, 4, 32, 48), s["lref"]).cuda()
    flows = [seeded((4, 2, 256, 384), s["flow0"] + q, 5.0).cuda() for q in range(12)]
    return lat, cond, tc, tu, lref, flows


def test_flow_sampler_full_size_vs_reference_golden(full_unet):
    """configs[2]: 3 DDIM steps of InferenceIP2PVideoOpticalFlow.second_clip_forward at 16f 256x384 with synthetic
    [12][4,2,256,384] flows; every step's latent vs the reference (all_latent)."""
    from insv2v_b200.inference import InferenceIP2PVideoOpticalFlow
    m, sd = full_unet
    O = _O()
    g = golden("sampler_full_flow.pt")
    lat, cond, tc, tu, lref, flows = _flow_inputs(g)
    pipe = InferenceIP2PVideoOpticalFlow(m, scheduler="ddim", num_ddim_steps=g["steps"], flow_estimator=object())
    pipe.obtain_flow_batched = lambda ref_images, query_images: flows  # synthetic flows stand in for RAFT
    out = pipe.second_clip_forward(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, latent_ref=lref,
                                   ref_images=torch.zeros(1, 4, 3, 8, 8), query_images=torch.zeros(1, 12, 3, 8, 8),
                                   noise_correct_step=g["noise_correct_step"], text_cfg=g["text_cfg"],
                                   img_cfg=g["img_cfg"])
    assert len(out["all_latent"]) == g["steps"] and len(out["all_pred"]) == g["steps"]

    def unet_fn(x, t, c):
        return O.unet3d_forward(sd, O.UNET_CONFIG_FULL, x, t, c).float()
    with torch.no_grad(), _RefFp16() as Oh:
        ref16 = Oh.sample_ip2p_video(unet_fn, lat, tc, tu, cond, g["text_cfg"], g["img_cfg"], g["steps"],
                                     latent_ref=lref, noise_correct_step=g["noise_correct_step"], flows=flows,
                                     return_all=True)
    for i in range(g["steps"]):
        _gate(f"flow sampler step {i + 1}/{g['steps']}", out["all_latent"][i], ref16["all_latent"][i].float(),
              g["all_latent"][i], 8e-3)  # the gate that binds is the 1.5x-of-reference-fp16 one inside _gate
    assert torch.equal(out["latent"], out["all_latent"][-1])


def test_config1_end_to_end_vs_reference_golden(full_unet, full_vae):
    """configs[0]: single 8-frame 256x256 clip, DDIM-20, text-cfg 7.5 / video-cfg 1.5, then the per-frame decode —
    the reference's InferenceIP2PVideo.__call__ + decode_latent_to_image run in fp32 on the CPU is the golden."""
    from insv2v_b200.inference import InferenceIP2PVideo
    from insv2v_b200.pipeline import InsV2VPipeline
    m, sd = full_unet
    vae, vsd = full_vae
    O = _O()
    g = golden("c1_e2e.pt")
    s = g["seeds"]
    lat, cond = seeded((1, 8, 4, 32, 32), s["lat"]).cuda(), seeded((1, 8, 4, 32, 32), s["cond"]).cuda()
    tc, tu = seeded((1, 77, 768), s["tc"]).cuda(), seeded((1, 77, 768), s["tu"]).cuda()
    pipe = InferenceIP2PVideo(m, scheduler="ddim", num_ddim_steps=g["steps"])
    out = pipe(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, text_cfg=g["text_cfg"], img_cfg=g["img_cfg"])

    def unet_fn(x, t, c):
        return O.unet3d_forward(sd, O.UNET_CONFIG_FULL, x, t, c).float()
    with torch.no_grad(), _RefFp16() as Oh:
        ref16 = Oh.sample_ip2p_video(unet_fn, lat, tc, tu, cond, g["text_cfg"], g["img_cfg"], g["steps"],
                                     return_all=True)
        fr16 = Oh.decode_latent_to_image(vsd, O.VAE_CONFIG_FULL, ref16["latent"]).float()
    for i in (0, 4, 9, 19):
        _gate(f"config1 DDIM-20 latent after step {i + 1}", out["all_latent"][i], ref16["all_latent"][i].float(),
              g["all_latent"][i], 2e-2)
    frames = InsV2VPipeline(m, vae).decode(out["latent"])
    _gate("config1 decoded frames", frames, fr16, g["frames"].float(), 2e-2)


def test_ddpm_rescale_start_time_vs_reference_golden():
    """scheduler='ddpm' (insv2v_run_loveu_tgve.py:64-74), guidance_rescale, start_time and the all_latent / all_pred
    lists against the reference's InferenceIP2PVideo / ...OpticalFlow on the micro UNet (fp32 CPU goldens)."""
    from insv2v_b200.inference import InferenceIP2PVideo, InferenceIP2PVideoOpticalFlow
    from insv2v_b200.pipeline import ddpm_timesteps
    from insv2v_b200.unet import UNet3DConditionModel
    O = _O()
    cfg = O.UNET_CONFIG_MICRO
    m = UNet3DConditionModel(**cfg)
    m.load_state_dict(O.seeded_state_dict(schema("unet_micro"), seed=100), strict=True)
    m = m.cuda().eval()
    g = golden("sampler_micro_ddpm.pt")
    s = g["seeds"]
    cd = cfg["cross_attention_dim"]
    steps = g["steps"]
    lat, cond = seeded((1, 6, 4, 16, 16), s["lat"]).cuda(), seeded((1, 6, 4, 16, 16), s["cond"]).cuda()
    tc, tu = seeded((1, 77, cd), s["tc"]).cuda(), seeded((1, 77, cd), s["tu"]).cuda()
    lref = seeded((1, 2, 4, 16, 16), s["lref"]).cuda()
    flows = [seeded((2, 2, 128, 128), s["flow0"] + q, 6.0).cuda() for q in range(4)]
    assert ddpm_timesteps(steps) == g["ddpm_timesteps_4"] and ddpm_timesteps(20) == g["ddpm_timesteps_20"]

    def noise_for(start):
        # DDPMScheduler.step draws torch.randn(model_output.shape) once per step with t > 0; the golden run seeded
        # the global CPU generator with noise_seed right before the loop
        torch.manual_seed(g["noise_seed"])
        k = sum(1 for t in g["ddpm_timesteps_4"][start:] if t > 0)
        return torch.stack([torch.randn(1, 6, 4, 16, 16) for _ in range(k)])

    def check(name, out, gold):
        for key in ("all_latent", "all_pred"):
            assert len(out[key]) == len(gold[key])
        e = err_stats(out["latent"], gold["latent"])
        ep = err_stats(out["all_pred"][-1], gold["all_pred"][-1])
        e0 = err_stats(out["all_latent"][0], gold["all_latent"][0])
        print(f"[{name}] latent rel_l2={e['rel_l2']:.3e} first-step rel_l2={e0['rel_l2']:.3e} "
              f"last pred rel_l2={ep['rel_l2']:.3e}")
        assert e0["rel_l2"] <= 1e-2 and e["rel_l2"] <= 3e-2 and ep["rel_l2"] <= 3e-2

    kw = dict(latent=lat, text_cond=tc, text_uncond=tu, img_cond=cond, text_cfg=g["text_cfg"], img_cfg=g["img_cfg"])
    pipe = InferenceIP2PVideo(m, scheduler="ddpm", num_ddim_steps=steps)
    assert [int(t) for t in pipe.scheduler.timesteps] == g["ddpm_timesteps_4"]
    den = pipe.pipe.denoise
    check("ddpm first clip", den(lat, tc, tu, cond, text_cfg=7.5, img_cfg=1.5, return_all=True, noise=noise_for(0)),
          g["ddpm_first"])
    check("ddpm rescale 0.7, start_time 1",
          den(lat, tc, tu, cond, text_cfg=7.5, img_cfg=1.5, return_all=True, noise=noise_for(1), guidance_rescale=0.7,
              start_time=1), g["ddpm_rescale_start1"])
    check("ddpm second clip (mean)",
          den(lat, tc, tu, cond, text_cfg=7.5, img_cfg=1.5, return_all=True, noise=noise_for(0), latent_ref=lref,
              noise_correct_step=0.5), g["ddpm_second_mean"])
    check("ddpm second clip (flow, rescale 0.3)",
          den(lat, tc, tu, cond, text_cfg=7.5, img_cfg=1.5, return_all=True, noise=noise_for(0), latent_ref=lref,
              noise_correct_step=0.5, flows=flows, guidance_rescale=0.3), g["ddpm_second_flow"])
    pd = InferenceIP2PVideo(m, scheduler="ddim", num_ddim_steps=steps)
    check("ddim rescale 0.5, start_time 2", pd(**kw, guidance_rescale=0.5, start_time=2), g["ddim_rescale_start2"])
    # the class API draws its own noise (device generator, as diffusers does): runs, finite, right structure
    out = pipe(**kw)
    assert torch.isfinite(out["latent"]).all() and len(out["all_latent"]) == steps
    pf = InferenceIP2PVideoOpticalFlow(m, scheduler="ddpm", num_ddim_steps=steps, flow_estimator=object())
    pf.obtain_flow_batched = lambda a, b: flows
    out = pf.second_clip_forward(**kw, latent_ref=lref, ref_images=torch.zeros(1, 2, 3, 8, 8),
                                 query_images=torch.zeros(1, 4, 3, 8, 8), noise_correct_step=0.5)
    assert torch.isfinite(out["latent"]).all() and len(out["all_pred"]) == steps
