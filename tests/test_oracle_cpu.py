"""CPU suite, part 1: the oracle against the golden vectors minted from the reference's own modules
(oracle/pin_against_reference.py). Runs anywhere (no GPU, no /root/reference)."""
import pytest
import torch

from oracle import insv2v_oracle as O
from tests.helpers import golden, schema, seeded

TOL = 2e-5


def _close(a, b, tol=TOL):
    assert a.shape == b.shape
    assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("tag,case", [("micro", "a"), ("micro", "b"), ("tiny", "a")])
def test_oracle_unet_matches_reference_golden(tag, case):
    cfg = {"micro": O.UNET_CONFIG_MICRO, "tiny": O.UNET_CONFIG_TINY}[tag]
    g = golden(f"unet_{tag}_{case}.pt")
    sd = O.seeded_state_dict(schema(f"unet_{tag}"), seed=g["weight_seed"])
    x = seeded(g["shape"], g["x_seed"])
    ctx = seeded((g["shape"][0], 77, cfg["cross_attention_dim"]), g["ctx_seed"])
    with torch.no_grad():
        y = O.unet3d_forward(sd, cfg, x, torch.tensor(g["t"]), ctx, video_start_index=g["vsi"])
    _close(y, g["out"])


def test_oracle_pe_guard_matches_reference():
    cfg = O.UNET_CONFIG_MICRO
    sd = O.seeded_state_dict(schema("unet_micro"), seed=100)
    with pytest.raises(ValueError):
        O.unet3d_forward(sd, cfg, seeded((1, 8, 4, 16, 16), 1), torch.tensor([1]), seeded((1, 77, 64), 2),
                         video_start_index=30)


def test_oracle_vae_matches_reference_golden():
    g = golden("vae_tiny.pt")
    sd = O.seeded_state_dict(schema("vae_tiny"), seed=g["weight_seed"])
    with torch.no_grad():
        y = O.vae_decode(sd, O.VAE_CONFIG_TINY, seeded(g["z_shape"], g["z_seed"]))
    _close(y, g["out"])


def test_oracle_flow_matches_reference_golden():
    g = golden("flow.pt")
    s = g["seeds"]
    img, flow, big = seeded((4, 4, 32, 48), s["img"]), seeded((4, 2, 32, 48), s["flow"], 6.0), \
        seeded((4, 2, 256, 384), s["big"], 5.0)
    _close(O.warp_image(img, flow), g["warp"], 1e-6)
    _close(O.resize_flow(big, (32, 48)), g["resize"], 1e-6)
    _close(O.resize_flow(big[:, :, :100, :90].contiguous(), (37, 53)), g["resize_general"], 1e-6)
    # exact /8 case: each output is the mean of the 2x2 source pixels at rows/cols (8i+3, 8i+4), times the scale
    manual = big[:, :, 3::8, :][:, :, :, 3::8] + big[:, :, 4::8, :][:, :, :, 3::8] \
        + big[:, :, 3::8, :][:, :, :, 4::8] + big[:, :, 4::8, :][:, :, :, 4::8]
    _close(manual / 4 / 8, g["resize"], 1e-5)


def test_oracle_sampler_matches_reference_golden():
    cfg = O.UNET_CONFIG_MICRO
    g = golden("sampler_micro.pt")
    s = g["seeds"]
    sd = O.seeded_state_dict(schema("unet_micro"), seed=g["weight_seed"])
    cd = cfg["cross_attention_dim"]
    lat, cond = seeded((1, 6, 4, 16, 16), s["lat"]), seeded((1, 6, 4, 16, 16), s["cond"])
    tc, tu, lref = seeded((1, 77, cd), s["tc"]), seeded((1, 77, cd), s["tu"]), seeded((1, 2, 4, 16, 16), s["lref"])
    flows = [seeded((2, 2, 128, 128), s["flow0"] + q, 6.0) for q in range(4)]
    fn = lambda x, t, c: O.unet3d_forward(sd, cfg, x, t, c)  # noqa: E731
    with torch.no_grad():
        a = O.sample_ip2p_video(fn, lat, tc, tu, cond, g["text_cfg"], g["img_cfg"], g["steps"])
        b = O.sample_ip2p_video(fn, lat, tc, tu, cond, g["text_cfg"], g["img_cfg"], g["steps"], latent_ref=lref,
                                noise_correct_step=g["noise_correct_step"])
        c = O.sample_ip2p_video(fn, lat, tc, tu, cond, g["text_cfg"], g["img_cfg"], g["steps"], latent_ref=lref,
                                noise_correct_step=g["noise_correct_step"], flows=flows)
    _close(a, g["first"], 1e-4)
    _close(b, g["second_mean"], 1e-4)
    _close(c, g["second_flow"], 1e-4)


def test_ddim_schedule_matches_reference():
    import json
    import os
    from tests.helpers import GOLD
    meta = json.load(open(os.path.join(GOLD, "meta.json")))
    assert O.ddim_timesteps(50) == meta["ddim_timesteps_50"] and O.ddim_timesteps(50)[0] == 981
    assert O.ddim_timesteps(3) == meta["ddim_timesteps_3"]
    from insv2v_b200 import pipeline as P
    assert P.ddim_timesteps(50) == meta["ddim_timesteps_50"]
    assert torch.equal(P.alphas_cumprod(), O.alphas_cumprod())


# ------------------------------------------------------------------------------------------------ RAFT optical flow
def test_raft_oracle_is_pinned_to_torchvision_raft_large():
    """The reference's RAFTFlow wraps torchvision.models.optical_flow.raft_large (misc_utils/flow_utils.py:155-159);
    the restatement in oracle/raft_oracle.py must reproduce that module bit for bit, in the mode the reference runs it
    (train: batch-statistics BatchNorm) and in eval mode, and its schema must be the module's state dict."""
    from torchvision.models.optical_flow import raft_large
    from oracle import raft_oracle as R
    sd = R.raft_seeded_state_dict(5)
    tv = raft_large(weights=None)
    assert {k: tuple(v.shape) for k, v in tv.state_dict().items()} == R.raft_schema()
    tv.load_state_dict(sd)
    g = torch.Generator().manual_seed(6)
    img1, img2 = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1, torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        for training in (True, False):
            tv.train(training)
            ref = tv(img1, img2, num_flow_updates=4)
            tv.load_state_dict(sd)  # train-mode forward updates the running statistics
            mine = R.raft_forward(sd, img1, img2, num_flow_updates=4, bn_training=training, return_all=True)
            assert len(ref) == len(mine) == 4
            for a, b in zip(ref, mine):
                assert torch.equal(a, b)


def test_raft_golden_is_the_oracles_output():
    from oracle import raft_oracle as R
    g = golden("raft_small.pt")
    sd = R.raft_seeded_state_dict(g["seed_w"])
    with torch.no_grad():
        flow = R.raft_flow(sd, g["img1"].float() / 255, g["img2"].float() / 255)
    _close(flow, g["flow"], 1e-5)


# ---- goldens minted at the benchmarked sizes / for the entry script's scheduler (oracle/pin_full_size.py) ----------
@pytest.mark.parametrize("tag", ["tiny", "full"])
def test_oracle_vae_encoder_matches_reference_golden(tag):
    cfg = {"tiny": O.VAE_CONFIG_TINY, "full": O.VAE_CONFIG_FULL}[tag]
    g = golden(f"vae_encoder_{tag}.pt")
    sd = O.seeded_state_dict(schema(f"vae_encoder_{tag}"), seed=g["weight_seed"])
    with torch.no_grad():
        m = O.vae_encode_moments(sd, cfg, seeded(g["x_shape"], g["x_seed"]))
    _close(m, g["moments"])
    torch.manual_seed(3)
    z = O.vae_encode(sd, cfg, seeded(g["x_shape"], g["x_seed"]))
    torch.manual_seed(3)
    mean, logvar = m.chunk(2, dim=1)
    assert torch.allclose(z, mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * torch.randn(mean.shape), atol=1e-6)


def test_oracle_vae_full_config_matches_reference_golden():
    g = golden("vae_full_decode.pt")
    sd = O.seeded_state_dict(schema("vae_full"), seed=g["weight_seed"])
    with torch.no_grad():
        y = O.vae_decode(sd, O.VAE_CONFIG_FULL, seeded(g["z_shape"], g["z_seed"])[:1])
    _close(y, g["out"][:1])


def test_oracle_ddpm_rescale_start_time_match_reference_golden():
    cfg = O.UNET_CONFIG_MICRO
    g = golden("sampler_micro_ddpm.pt")
    s = g["seeds"]
    sd = O.seeded_state_dict(schema("unet_micro"), seed=g["weight_seed"])
    cd = cfg["cross_attention_dim"]
    lat, cond = seeded((1, 6, 4, 16, 16), s["lat"]), seeded((1, 6, 4, 16, 16), s["cond"])
    tc, tu = seeded((1, 77, cd), s["tc"]), seeded((1, 77, cd), s["tu"])
    lref = seeded((1, 2, 4, 16, 16), s["lref"])
    flows = [seeded((2, 2, 128, 128), s["flow0"] + q, 6.0) for q in range(4)]
    assert O.ddpm_timesteps(4) == g["ddpm_timesteps_4"] and O.ddpm_timesteps(20) == g["ddpm_timesteps_20"]

    def run(**kw):
        torch.manual_seed(g["noise_seed"])
        with torch.no_grad():
            return O.sample_ip2p_video(lambda x, t, c: O.unet3d_forward(sd, cfg, x, t, c), lat, tc, tu, cond,
                                       g["text_cfg"], g["img_cfg"], g["steps"], return_all=True, **kw)
    for name, kw in (("ddpm_first", dict(scheduler="ddpm")),
                     ("ddpm_rescale_start1", dict(scheduler="ddpm", guidance_rescale=0.7, start_time=1)),
                     ("ddpm_second_mean", dict(scheduler="ddpm", latent_ref=lref, noise_correct_step=0.5)),
                     ("ddpm_second_flow", dict(scheduler="ddpm", latent_ref=lref, noise_correct_step=0.5, flows=flows,
                                               guidance_rescale=0.3)),
                     ("ddim_rescale_start2", dict(scheduler="ddim", guidance_rescale=0.5, start_time=2))):
        out = run(**kw)
        assert len(out["all_latent"]) == len(g[name]["all_latent"])
        _close(out["latent"], g[name]["latent"], 1e-4)
        _close(out["all_pred"][-1], g[name]["all_pred"][-1], 1e-4)
        _close(out["all_latent"][0], g[name]["all_latent"][0], 1e-4)
