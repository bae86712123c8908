"""TEST INFRASTRUCTURE — CPU oracle for the InsV2V denoising hot path.  NOT part of the product.

A plain-PyTorch fp32 restatement of the reference algorithm for the path BASELINE.json names: the 3-D video UNet forward,
the KL-VAE decode, the optical-flow warp and the sampling loop that drives them. Every function cites the reference
file:line it follows (paths relative to /root/reference, commit 6a51b48). The arithmetic of diffusers 0.21.4 (Attention,
FeedForward/GEGLU, Timesteps, TimestepEmbedding, DDIM/DDPM schedulers — pinned by THIRD-PARTY:31, not vendored in the
reference) is restated from its published algorithm.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned against the reference's
OWN modules executed in the build container by oracle/pin_against_reference.py (which imports /root/reference unchanged
through oracle/shim) — that script asserts agreement to ~1e-6 and writes tests/golden/*.pt. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.

Weights are a flat {name: tensor} state dict with the reference's key names; the network
structure is driven by the same config dict the reference's YAML provides.
"""
import math

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------------------------
# configs (configs/instruct_v2v_inference.yaml:23-89) and the seeded state-dict generator shared by tests and bench
# ------------------------------------------------------------------------------------------------------------------
UNET_CONFIG_FULL = dict(
    in_channels=8, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    flip_sin_to_cos=True, freq_shift=0, use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8),
    motion_module_mid_block=False, motion_module_decoder_only=False,
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                              temporal_attention_dim_div=1),
)
# same topology, 1/5 width: small enough that the CPU oracle runs in seconds
UNET_CONFIG_TINY = dict(UNET_CONFIG_FULL, block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)
# one layer per block, for the quickest end-to-end checks
UNET_CONFIG_MICRO = dict(UNET_CONFIG_FULL, block_out_channels=(64, 64, 128, 128), layers_per_block=1,
                         cross_attention_dim=64)

VAE_CONFIG_FULL = dict(embed_dim=4, ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3,
                                                  ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=(),
                                                  dropout=0.0))
VAE_CONFIG_TINY = dict(embed_dim=4, ddconfig=dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3,
                                                  ch=64, ch_mult=(1, 2, 2, 2), num_res_blocks=1, attn_resolutions=(),
                                                  dropout=0.0))


def sinusoidal_pe(max_len, d_model):
    """PositionalEncoding buffer, motion_module.py:228-234."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def seeded_state_dict(schema, seed):
    """Deterministic weights for a {name: shape} schema: the same call gives the same tensors in the build container
    (reference side) and on the GPU box (product side). Matrices ~ N(0, 1/fan_in), norm scales ~ 1 + 0.1 N, biases
    ~ 0.1 N; `pe` buffers keep their closed form. The motion modules' zero-initialised proj_out (motion_module.py:68)
    is randomised too — otherwise the whole temporal path would contribute exactly 0 (SURVEY.md §4 trap a)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(schema):
        shape = tuple(schema[name])
        if name.endswith("pos_encoder.pe"):
            sd[name] = sinusoidal_pe(shape[1], shape[2])
        elif len(shape) == 1:
            r = torch.randn(shape, generator=g) * 0.1
            is_norm_scale = name.endswith(".weight")  # 1-D weights are always norm scales in these networks
            sd[name] = 1.0 + r if is_norm_scale else r
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[name] = torch.randn(shape, generator=g) * fan_in ** -0.5
    return sd


# ------------------------------------------------------------------------------------------------------------------
# diffusers 0.21.4 pieces (restated)
# ------------------------------------------------------------------------------------------------------------------
def timestep_sinusoid(timesteps, dim, flip_sin_to_cos=True, freq_shift=0.0):
    """diffusers get_timestep_embedding (models/embeddings.py), called through Timesteps at unet.py:358."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def linear(sd, pfx, x):
    return F.linear(x, sd[pfx + ".weight"], sd.get(pfx + ".bias"))


def attention_core(q, k, v, heads):
    """softmax(q k^T / sqrt(d)) v per head — diffusers AttnProcessor2_0 / xformers memory_efficient_attention."""
    b, sq, c = q.shape
    d = c // heads
    qh = q.reshape(b, sq, heads, d).transpose(1, 2)
    kh = k.reshape(b, -1, heads, d).transpose(1, 2)
    vh = v.reshape(b, -1, heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * (d ** -0.5), dim=-1)
    return (p @ vh).transpose(1, 2).reshape(b, sq, c)


def attention(sd, pfx, x, ctx, heads):
    """diffusers Attention.forward: to_q / to_k / to_v (no bias), core, to_out.0 (bias)."""
    ctx = x if ctx is None else ctx
    q, k, v = linear(sd, pfx + ".to_q", x), linear(sd, pfx + ".to_k", ctx), linear(sd, pfx + ".to_v", ctx)
    return linear(sd, pfx + ".to_out.0", attention_core(q, k, v, heads))


def feed_forward(sd, pfx, x):
    """diffusers FeedForward(activation_fn='geglu'): net.0.proj -> hidden * gelu_erf(gate) -> net.2."""
    hidden, gate = linear(sd, pfx + ".net.0.proj", x).chunk(2, dim=-1)
    return linear(sd, pfx + ".net.2", hidden * F.gelu(gate))


# ------------------------------------------------------------------------------------------------------------------
# modules/video_unet_temporal
# ------------------------------------------------------------------------------------------------------------------
def inflated_conv(sd, pfx, x, stride=1, padding=1):
    """InflatedConv3d.forward, resnet.py:10-18: Conv2d on every frame."""
    b, c, f, h, w = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w), sd[pfx + ".weight"], sd[pfx + ".bias"],
                 stride=stride, padding=padding)
    return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def resnet_block3d(sd, pfx, x, temb, eps, groups):
    """ResnetBlock3D.forward, resnet.py:174-204. GroupNorm acts on the 5-D tensor: statistics span f*h*w jointly."""
    h = F.group_norm(x, groups, sd[pfx + ".norm1.weight"], sd[pfx + ".norm1.bias"], eps)
    h = inflated_conv(sd, pfx + ".conv1", F.silu(h))
    h = h + linear(sd, pfx + ".time_emb_proj", F.silu(temb))[:, :, None, None, None]
    h = F.group_norm(h, groups, sd[pfx + ".norm2.weight"], sd[pfx + ".norm2.bias"], eps)
    h = inflated_conv(sd, pfx + ".conv2", F.silu(h))
    if pfx + ".conv_shortcut.weight" in sd:
        x = inflated_conv(sd, pfx + ".conv_shortcut", x, padding=0)
    return x + h  # output_scale_factor = 1.0


def transformer3d(sd, pfx, x, ctx, heads, groups):
    """Transformer3DModel.forward (attention.py:91-138) + BasicTransformerBlock.forward (:233-270)."""
    b, c, f, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    ctx_f = ctx.repeat_interleave(f, dim=0)  # 'b n c -> (b f) n c', attention.py:96
    res = xf
    hs = F.group_norm(xf, groups, sd[pfx + ".norm.weight"], sd[pfx + ".norm.bias"], 1e-6)
    hs = F.conv2d(hs, sd[pfx + ".proj_in.weight"], sd[pfx + ".proj_in.bias"])
    hs = hs.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    tb = pfx + ".transformer_blocks.0"
    n = F.layer_norm(hs, (c,), sd[tb + ".norm1.weight"], sd[tb + ".norm1.bias"])
    hs = attention(sd, tb + ".attn1", n, None, heads) + hs
    n = F.layer_norm(hs, (c,), sd[tb + ".norm2.weight"], sd[tb + ".norm2.bias"])
    hs = attention(sd, tb + ".attn2", n, ctx_f, heads) + hs
    n = F.layer_norm(hs, (c,), sd[tb + ".norm3.weight"], sd[tb + ".norm3.bias"])
    hs = feed_forward(sd, tb + ".ff", n) + hs
    hs = hs.reshape(b * f, h, w, c).permute(0, 3, 1, 2)
    hs = F.conv2d(hs, sd[pfx + ".proj_out.weight"], sd[pfx + ".proj_out.bias"]) + res
    return hs.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def motion_module(sd, pfx, x, heads, groups, n_attn_blocks, video_start_index=0):
    """VanillaTemporalModule -> TemporalTransformer3DModel.forward (motion_module.py:128-152),
    TemporalTransformerBlock.forward (:204-217), VersatileAttention.forward (:270-336), PositionalEncoding (:236-242)."""
    p = pfx + ".temporal_transformer"
    b, c, f, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    res = xf
    hs = F.group_norm(xf, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    hs = hs.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    hs = linear(sd, p + ".proj_in", hs)
    tb = p + ".transformer_blocks.0"
    d = h * w
    for i in range(n_attn_blocks):
        n = F.layer_norm(hs, (c,), sd[f"{tb}.norms.{i}.weight"], sd[f"{tb}.norms.{i}.bias"])
        t = n.reshape(b, f, d, c).permute(0, 2, 1, 3).reshape(b * d, f, c)  # '(b f) d c -> (b d) f c'
        pe = sd[f"{tb}.attention_blocks.{i}.pos_encoder.pe"]
        start = video_start_index
        if start + f > pe.shape[1]:
            start = start - pe.shape[1]
        if start < 0:
            raise ValueError(f"start_index must be non-negative, but got {start}")
        t = t + pe[:, start:start + f]
        t = attention(sd, f"{tb}.attention_blocks.{i}", t, None, heads)
        t = t.reshape(b, d, f, c).permute(0, 2, 1, 3).reshape(b * f, d, c)
        hs = t + hs
    n = F.layer_norm(hs, (c,), sd[tb + ".ff_norm.weight"], sd[tb + ".ff_norm.bias"])
    hs = feed_forward(sd, tb + ".ff", n) + hs
    hs = linear(sd, p + ".proj_out", hs)
    hs = hs.reshape(b * f, h, w, c).permute(0, 3, 1, 2) + res
    return hs.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def unet3d_forward(sd, cfg, sample, timestep, encoder_hidden_states, video_start_index=0):
    """UNet3DConditionModel.forward, unet.py:296-434 (blocks: unet_blocks.py:229-236, 352-362, 447-456, 557-589,
    655-678). sample [b, c_in, f, h, w] fp32, timestep [b] or scalar, encoder_hidden_states [b, 77, ctx]."""
    boc = cfg["block_out_channels"]
    groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    heads = cfg["attention_head_dim"]
    heads = (heads,) * len(boc) if isinstance(heads, int) else tuple(heads)
    mm = cfg["motion_module_kwargs"]
    mheads, n_attn = mm["num_attention_heads"], len(mm["attention_block_types"])
    lpb = cfg["layers_per_block"]
    b = sample.shape[0]
    t = torch.as_tensor(timestep)
    t = t[None] if t.dim() == 0 else t
    t = t.expand(b)
    default_up = 2 ** (len(boc) - 1)
    forward_upsample_size = any(s % default_up != 0 for s in sample.shape[-2:])  # unet.py:329-331

    temb = timestep_sinusoid(t, boc[0], cfg.get("flip_sin_to_cos", True), cfg.get("freq_shift", 0))
    temb = linear(sd, "time_embedding.linear_2", F.silu(linear(sd, "time_embedding.linear_1", temb)))

    def has_motion(res, down):
        if not cfg.get("use_motion_module", True) or res not in cfg["motion_module_resolutions"]:
            return False
        return not (down and cfg.get("motion_module_decoder_only", False))

    x = inflated_conv(sd, "conv_in", sample)
    skips = [x]
    for i, btype in enumerate(cfg["down_block_types"]):
        p = f"down_blocks.{i}"
        for j in range(lpb):
            x = resnet_block3d(sd, f"{p}.resnets.{j}", x, temb, eps, groups)
            if btype == "CrossAttnDownBlock3D":
                x = transformer3d(sd, f"{p}.attentions.{j}", x, encoder_hidden_states, heads[i], groups)
            if has_motion(2 ** i, True):
                x = motion_module(sd, f"{p}.motion_modules.{j}", x, mheads, groups, n_attn, video_start_index)
            skips.append(x)
        if i != len(boc) - 1:
            x = inflated_conv(sd, f"{p}.downsamplers.0.conv", x, stride=2, padding=1)
            skips.append(x)

    x = resnet_block3d(sd, "mid_block.resnets.0", x, temb, eps, groups)
    x = transformer3d(sd, "mid_block.attentions.0", x, encoder_hidden_states, heads[-1], groups)
    if cfg.get("use_motion_module", True) and cfg.get("motion_module_mid_block", False):
        x = motion_module(sd, "mid_block.motion_modules.0", x, mheads, groups, n_attn, video_start_index)
    x = resnet_block3d(sd, "mid_block.resnets.1", x, temb, eps, groups)

    rheads = tuple(reversed(heads))
    for i, btype in enumerate(cfg["up_block_types"]):
        p = f"up_blocks.{i}"
        res_level = 2 ** (len(boc) - 1 - i)
        is_final = i == len(boc) - 1
        upsample_size = None
        for j in range(lpb + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block3d(sd, f"{p}.resnets.{j}", x, temb, eps, groups)
            if btype == "CrossAttnUpBlock3D":
                x = transformer3d(sd, f"{p}.attentions.{j}", x, encoder_hidden_states, rheads[i], groups)
            if has_motion(res_level, False):
                x = motion_module(sd, f"{p}.motion_modules.{j}", x, mheads, groups, n_attn, video_start_index)
        if not is_final:
            if forward_upsample_size:
                upsample_size = skips[-1].shape[2:]  # unet.py:409-410
            if upsample_size is None:
                x = F.interpolate(x, scale_factor=[1.0, 2.0, 2.0], mode="nearest")  # resnet.py:59
            else:
                x = F.interpolate(x, size=tuple(upsample_size), mode="nearest")
            x = inflated_conv(sd, f"{p}.upsamplers.0.conv", x)

    x = F.silu(F.group_norm(x, groups, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], eps))
    return inflated_conv(sd, "conv_out", x)


# ------------------------------------------------------------------------------------------------------------------
# modules/kl_autoencoder + modules/vqvae/model.py (decode only)
# ------------------------------------------------------------------------------------------------------------------
def _vae_norm(sd, pfx, x):
    return F.group_norm(x, 32, sd[pfx + ".weight"], sd[pfx + ".bias"], 1e-6)  # Normalize, vqvae/model.py:31-32


def _conv(sd, pfx, x, padding):
    return F.conv2d(x, sd[pfx + ".weight"], sd[pfx + ".bias"], padding=padding)


def vae_resnet(sd, pfx, x):
    """ResnetBlock.forward (temb is None), vqvae/model.py:116-136; swish = x*sigmoid(x) (:25-27)."""
    h = _conv(sd, pfx + ".conv1", F.silu(_vae_norm(sd, pfx + ".norm1", x)), 1)
    h = _conv(sd, pfx + ".conv2", F.silu(_vae_norm(sd, pfx + ".norm2", h)), 1)
    if pfx + ".nin_shortcut.weight" in sd:
        x = _conv(sd, pfx + ".nin_shortcut", x, 0)
    return x + h


def vae_attn(sd, pfx, x):
    """AttnBlock.forward, vqvae/model.py:173-197: single head, d = c, explicit softmax over keys."""
    h_ = _vae_norm(sd, pfx + ".norm", x)
    q, k, v = _conv(sd, pfx + ".q", h_, 0), _conv(sd, pfx + ".k", h_, 0), _conv(sd, pfx + ".v", h_, 0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    h_ = torch.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, pfx + ".proj_out", h_, 0)


def vae_decode(sd, cfg, z):
    """AutoencoderKL.decode (autoencoder.py:97-100) -> Decoder.forward (vqvae/model.py:378-411). z [n, 4, h, w]."""
    dd = cfg["ddconfig"]
    nres = len(dd["ch_mult"])
    h = _conv(sd, "post_quant_conv", z, 0)
    h = _conv(sd, "decoder.conv_in", h, 1)
    h = vae_resnet(sd, "decoder.mid.block_1", h)
    h = vae_attn(sd, "decoder.mid.attn_1", h)
    h = vae_resnet(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(nres)):
        for blk in range(dd["num_res_blocks"] + 1):
            h = vae_resnet(sd, f"decoder.up.{lvl}.block.{blk}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # Upsample.forward :46-52
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", h, 1)
    h = F.silu(_vae_norm(sd, "decoder.norm_out", h))
    return _conv(sd, "decoder.conv_out", h, 1)


def decode_latent_to_image(sd, cfg, latents, scale_factor=0.18215):
    """InstructP2PVideoTrainer.decode_latent_to_image (instruct_p2p_video.py:66-79): latents [b, f, 4, h, w] ->
    [b, f, 3, 8h, 8w], one frame at a time, latent divided by scale_factor first (diffusion.py:247-249)."""
    b, f = latents.shape[:2]
    frames = [vae_decode(sd, cfg, latents[:, i] / scale_factor) for i in range(f)]
    return torch.stack(frames, dim=1)


def vae_encode_moments(sd, cfg, x):
    """AutoencoderKL.encode up to the moments (autoencoder.py:89-91): Encoder.forward (vqvae/model.py:275-302; attn
    lists are empty for attn_resolutions=()), Downsample with the asymmetric (0,1,0,1) zero pad (:67-71), quant_conv.
    x [n, 3, H, W] -> moments [n, 2*embed_dim, H/8, W/8] (mean | logvar)."""
    dd = cfg["ddconfig"]
    nres = len(dd["ch_mult"])
    h = _conv(sd, "encoder.conv_in", x, 1)
    for lvl in range(nres):
        for blk in range(dd["num_res_blocks"]):
            h = vae_resnet(sd, f"encoder.down.{lvl}.block.{blk}", h)
        if lvl != nres - 1:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"encoder.down.{lvl}.downsample.conv.weight"],
                         sd[f"encoder.down.{lvl}.downsample.conv.bias"], stride=2)
    h = vae_resnet(sd, "encoder.mid.block_1", h)
    h = vae_attn(sd, "encoder.mid.attn_1", h)
    h = vae_resnet(sd, "encoder.mid.block_2", h)
    h = _conv(sd, "encoder.conv_out", F.silu(_vae_norm(sd, "encoder.norm_out", h)), 1)
    return _conv(sd, "quant_conv", h, 0)


def vae_encode(sd, cfg, x, noise=None):
    """AutoencoderKL.encode (autoencoder.py:89-95) = DiagonalGaussianDistribution(moments).sample() (:10-24): logvar
    clamped to [-30, 20], mean + exp(0.5 logvar) * randn drawn on the CPU default generator and moved to the device."""
    mean, logvar = torch.chunk(vae_encode_moments(sd, cfg, x), 2, dim=1)
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    noise = torch.randn(mean.shape) if noise is None else noise
    return mean + std * noise.to(mean.device)


# ------------------------------------------------------------------------------------------------------------------
# misc_utils/flow_utils.py
# ------------------------------------------------------------------------------------------------------------------
def warp_image(image, flow, mode="bilinear"):
    """flow_utils.py:25-57."""
    if image.dim() == 3:
        image = image.unsqueeze(0)
    if flow.dim() == 3:
        flow = flow.unsqueeze(0)
    assert image.shape[0] == flow.shape[0] and image.shape[2:] == flow.shape[2:]
    n, _, h, w = image.shape
    ys, xs = torch.meshgrid(torch.arange(h, device=image.device), torch.arange(w, device=image.device), indexing="ij")
    grid = torch.stack([xs, ys], dim=-1).to(torch.float32)[None].repeat(n, 1, 1, 1)
    grid = grid + flow.permute(0, 2, 3, 1)
    gx = 2 * (grid[..., 0] / (w - 1) - 0.5)
    gy = 2 * (grid[..., 1] / (h - 1) - 0.5)
    return F.grid_sample(image, torch.stack([gx, gy], dim=-1), mode=mode, align_corners=True)


def resize_flow(flow, size):
    """flow_utils.py:59-86."""
    H, W = size
    h, w = flow.shape[2:]
    scaled = flow.clone()
    scaled[:, 0] *= W / w
    scaled[:, 1] *= H / h
    return F.interpolate(scaled, size=(H, W), mode="bilinear", align_corners=False)


# ------------------------------------------------------------------------------------------------------------------
# pl_trainer/inference/inference.py (sampling loops) + diffusers schedulers (restated)
# ------------------------------------------------------------------------------------------------------------------
def alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000):
    """scaled_linear schedule, inference.py:31,44-49 -> diffusers: linspace(sqrt(b0), sqrt(b1), n)**2."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_steps, n_train=1000, steps_offset=1):
    """DDIMScheduler.set_timesteps, timestep_spacing='leading', steps_offset=1 (inference.py:37)."""
    ratio = n_train // num_steps
    return [int(i * ratio + steps_offset) for i in reversed(range(num_steps))]


def ddpm_timesteps(num_steps, n_train=1000):
    """DDPMScheduler.set_timesteps, timestep_spacing='leading' (no steps_offset): 950, 900 ... 0 for 20 steps."""
    ratio = n_train // num_steps
    return [int(i * ratio) for i in reversed(range(num_steps))]


def ddim_step(ac, eps, t, x, num_steps, n_train=1000):
    """DDIMScheduler.step with eta=0, epsilon prediction, clip_sample=False, set_alpha_to_one=False."""
    prev_t = t - n_train // num_steps
    a_t = ac[t]
    a_prev = ac[prev_t] if prev_t >= 0 else ac[0]
    x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps, x0


def ddpm_coefficients(ac, t, num_steps, n_train=1000):
    """The scalars of DDPMScheduler.step (diffusers 0.21.4, epsilon prediction, variance_type='fixed_small',
    clip_sample=False): returns (sqrt(1-a_t), sqrt(a_t), x0 coefficient, sample coefficient, sigma). The posterior
    mean is c0*x0 + cs*x_t (DDPM eq. 7) and sigma^2 = clamp((1-a_prev)/(1-a_t) * beta_t, 1e-20) for t > 0, else 0."""
    prev_t = t - n_train // num_steps
    a_t = ac[t]
    a_prev = ac[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    cur_alpha = a_t / a_prev
    cur_beta = 1 - cur_alpha
    c0 = (a_prev ** 0.5 * cur_beta) / (1 - a_t)
    cs = cur_alpha ** 0.5 * (1 - a_prev) / (1 - a_t)
    var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)
    sigma = var ** 0.5 if t > 0 else torch.tensor(0.0)
    return (1 - a_t) ** 0.5, a_t ** 0.5, c0, cs, sigma


def ddpm_step(ac, eps, t, x, num_steps, n_train=1000, generator=None):
    """DDPMScheduler.step: the variance noise is torch.randn(eps.shape) on the (global, unless given) CPU generator."""
    sb, sa, c0, cs, sigma = ddpm_coefficients(ac, t, num_steps, n_train)
    x0 = (x - sb * eps) / sa
    prev = c0 * x0 + cs * x
    if t > 0:
        prev = prev + sigma * torch.randn(eps.shape, generator=generator, dtype=eps.dtype).to(eps.device)
    return prev, x0


def cfg_combine(e1, e2, e3, text_cfg, img_cfg):
    """inference.py:198-203."""
    return e1 + img_cfg * (e2 - e1) + text_cfg * (e3 - e2)


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """inference.py:13-24: unbiased std over all but the batch dim; NOTE the reference passes noise_pred1 (the fully
    unconditional branch) as `noise_pred_text` (:205-206), which is mirrored by the callers below."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


def sample_ip2p_video(unet_fn, latent, text_cond, text_uncond, img_cond, text_cfg=7.5, img_cfg=1.2, num_steps=20,
                      latent_ref=None, noise_correct_step=1.0, flows=None, scheduler="ddim", start_time=0,
                      guidance_rescale=0.0, return_all=False, generator=None):
    """InferenceIP2PVideo.__call__ (inference.py:163-219), .second_clip_forward (:221-289) when latent_ref is given,
    and InferenceIP2PVideoOpticalFlow.second_clip_forward (:314-398) when flows [Q][R,2,H,W] are given too.
    unet_fn(x [3, 8, f, h, w], t LongTensor[3], ctx [3, 77, c]) -> eps [3, 4, f, h, w]. scheduler 'ddim' | 'ddpm'
    (inference.py:35-49). The loop index i restarts at 0 when start_time > 0 (enumerate over the sliced timesteps,
    :181,240), so the noise-correction window counts from the first executed step, as in the reference."""
    ac = alphas_cumprod()
    latent = latent.clone()
    steps = ddim_timesteps(num_steps) if scheduler == "ddim" else ddpm_timesteps(num_steps)
    all_latent, all_pred = [], []
    for i, t in enumerate(steps[start_time:]):
        l1 = torch.cat([latent, torch.zeros_like(img_cond)], dim=2)
        l2 = torch.cat([latent, img_cond], dim=2)
        x = torch.cat([l1, l2, l2.clone()], dim=0).permute(0, 2, 1, 3, 4)  # 'b f c h w -> b c f h w'
        ctx = torch.cat([text_uncond, text_uncond, text_cond], dim=0)
        eps = unet_fn(x, torch.full((3,), t, dtype=torch.long, device=x.device), ctx).permute(0, 2, 1, 3, 4)
        e1, e2, e3 = eps.chunk(3, dim=0)
        eps = cfg_combine(e1, e2, e3, text_cfg, img_cfg)
        if guidance_rescale > 0:
            eps = rescale_noise_cfg(eps, e1, guidance_rescale)
        if latent_ref is not None and noise_correct_step * num_steps > i:
            r = latent_ref.shape[1]
            a_t = ac[t]
            noise_ref = (latent[:, :r] - a_t ** 0.5 * latent_ref) / (1 - a_t) ** 0.5
            delta = noise_ref - eps[:, :r]
            eps = eps.clone()
            eps[:, :r] = eps[:, :r] + delta
            if flows is None:
                eps[:, r:] = eps[:, r:] + delta.mean(dim=1, keepdim=True)  # inference.py:275-277
            else:
                for q, flow in zip(range(r, eps.shape[1]), flows):  # inference.py:374-386
                    fl = resize_flow(flow, delta.shape[3:])
                    warped = warp_image(delta[0], fl)
                    mask = warp_image(torch.ones_like(delta[0])[:, :1], fl)
                    msum = mask[None].sum(dim=1, keepdim=True)
                    corr = torch.where(msum > 0.5, warped[None].sum(dim=1, keepdim=True) / msum,
                                       torch.zeros_like(msum))
                    eps[:, q:q + 1] += torch.where(msum > 0.5, corr, torch.zeros_like(corr))
        if scheduler == "ddim":
            latent, x0 = ddim_step(ac, eps, t, latent, num_steps)
        else:
            latent, x0 = ddpm_step(ac, eps, t, latent, num_steps, generator=generator)
        all_latent.append(latent)
        all_pred.append(x0)
    if return_all:
        return {"latent": latent, "all_latent": all_latent, "all_pred": all_pred}
    return latent
