"""One launch of every hot-path kernel family at small / moderate shapes: the command run under
`compute-sanitizer --tool memcheck` and `--tool racecheck` (tools/gpu_sanitize.sh). Only C-ABI launches and torch
allocations happen here, so the sanitizer's slow-down stays bounded."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def h16(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).half()


def f32(*shape, scale=1.0):
    return torch.randn(*shape, device=dev, generator=g) * scale


done = []


def run(name, fn):
    out = fn()
    torch.cuda.synchronize()
    outs = out if isinstance(out, (tuple, list)) else [out]
    assert all(torch.isfinite(o.float()).all() for o in outs if torch.is_tensor(o)), name
    done.append(name)


# ---- GEMM family -----------------------------------------------------------------------------------------------------
def lin(rows, k, n, res=True, bias=True, **kw):
    x, w = h16(rows, k), ops.pack_linear(h16(n, k, scale=k ** -0.5))
    return ops.linear(x, w, bias=h16(n) if bias else None, residual=h16(rows, n) if res else None, **kw)


run("pair160 residual 4000x320x320", lambda: lin(4000, 320, 320))
run("pair160 no-residual 3000x640x1920", lambda: lin(3000, 640, 1920, res=False, bias=False))
run("pair160 ghost tile 1152x1280x1280", lambda: lin(1152, 1280, 1280))
run("v2 pair 256-wide 2000x2560x512", lambda: lin(2000, 2560, 512))
run("v2 DS 128-wide 3000x256x128", lambda: lin(3000, 256, 128))
run("pair160 activation-stationary 19000x320x960 (>= 74 M pairs)", lambda: lin(19000, 320, 960, res=False, bias=False))
run("pair160 one slab / six stages + residual 3000x1280x640", lambda: lin(3000, 1280, 640))
run("320-wide pair tiles 2880x2560x1280 + residual (register residual, two slabs per group)", lambda: lin(2880, 2560, 1280))
run("single tile 3x320x1280", lambda: lin(3, 320, 1280, res=False))
run("ragged N 130x72x40", lambda: lin(130, 72, 40))


def geglu(rows, c):
    w, b = ops.pack_geglu(h16(8 * c, c, scale=c ** -0.5), h16(8 * c, scale=0.1))
    return ops.linear(h16(rows, c), w, bias=b, geglu=True)


run("GEGLU 2000x320", lambda: geglu(2000, 320))
run("GEGLU 300x1280", lambda: geglu(300, 1280))
run("GEGLU 6000x320 (several tiles per cluster: staged bias, two output slabs)", lambda: geglu(6000, 320))
run("GEGLU activation-stationary 19000x320", lambda: geglu(19000, 320))


def conv(n, ci, co, h, w, res=False, rowbias=False, f32out=False):
    x, wt = h16(n * h * w, ci), ops.pack_conv3x3(h16(co, ci, 3, 3, scale=(9 * ci) ** -0.5))
    kw = {}
    if rowbias:
        kw = dict(rowbias=h16(2, co), rowbias_group=(n // 2) * h * w)
    return ops.conv3x3(x, wt, n, h, w, bias=h16(co), residual=h16(n * h * w, co) if res else None, out_f32=f32out, **kw)


run("halo conv 16x8 box, 160-wide 6x320x320 32x48 +temb +res", lambda: conv(6, 320, 320, 32, 48, res=True, rowbias=True))
run("halo conv 8x16 box 4x640x640 16x24", lambda: conv(4, 640, 640, 16, 24))
run("halo conv 128-wide VAE 2x128x128 64x96", lambda: conv(2, 128, 128, 64, 96))
run("per-tap conv 8x12 level 16x1280x1280", lambda: conv(16, 1280, 1280, 8, 12))
run("per-tap conv 8x12 level, 320-wide tiles 30x320x1280 +temb +res", lambda: conv(30, 320, 1280, 8, 12, res=True, rowbias=True))
run("split-K conv 4x6 level 48x1280x1280", lambda: conv(48, 1280, 1280, 4, 6, res=True))
run("fp32 head conv 320->4", lambda: conv(2, 320, 4, 32, 48, f32out=True))
run("stride-2 conv", lambda: ops.conv3x3_s2(h16(4 * 16 * 24, 320), ops.pack_conv3x3_im2col(h16(320, 320, 3, 3, scale=0.02)),
                                            4, 16, 24, bias=h16(320))[0])

# ---- attention ---------------------------------------------------------------------------------------------------------
def attn(n, s, skv, heads, d, kv_div=1):
    c = heads * d
    q = h16(n * s, c)
    kv = h16((n // kv_div) * skv, 2 * c)
    return ops.attention(q, kv[:, :c], kv[:, c:], n_batch=n, s_q=s, s_kv=skv, heads=heads, d=d, q_ld=c, kv_ld=2 * c,
                         kv_div=kv_div)


run("attention persistent pairs S=1536 d=40", lambda: attn(2, 1536, 1536, 8, 40))
run("attention cross 77 keys d=40 (kv_div)", lambda: attn(4, 1536, 77, 8, 40, kv_div=2))
run("attention one-tile d=80 S=384", lambda: attn(4, 384, 384, 8, 80))
run("attention one-tile d=160 S=96", lambda: attn(4, 96, 96, 8, 160))
run("attention ragged S=100 d=40", lambda: attn(3, 100, 100, 8, 40))
run("attention persistent one-tile d=80 S=384 (576 items)", lambda: attn(24, 384, 384, 8, 80))
run("attention persistent one-tile d=160 77 keys (kv_div)", lambda: attn(48, 96, 77, 8, 160, kv_div=16))
for fr in (16, 24, 64):
    run(f"temporal attention F={fr}", lambda fr=fr: ops.temporal_attention(h16(2 * fr * 24, 3 * 320), 2, fr, 24, 320, 8))

# ---- norms -------------------------------------------------------------------------------------------------------------
run("groupnorm two-kernel form (more batch groups than SMs)",
    lambda: ops.groupnorm(h16(160 * 24, 320), h16(320), h16(320), 160, 24, 32, 1, 1e-5, True))
run("groupnorm one-kernel form, two sources", lambda: ops.groupnorm2(h16(8 * 96, 640), h16(8 * 96, 320), h16(960), h16(960),
                                                                     8, 96, 32, 4, 1e-5, True))
run("groupnorm 5-D + SiLU", lambda: ops.groupnorm(h16(2 * 4 * 384, 320), h16(320), h16(320), 8, 384, 32, 4, 1e-5, True))
run("groupnorm per frame", lambda: ops.groupnorm(h16(6 * 96, 1280), h16(1280), h16(1280), 6, 96, 32, 1, 1e-6, False))
run("layernorm C=320", lambda: ops.layernorm(h16(5000, 320), h16(320), h16(320)))
run("layernorm C=1280 + pe", lambda: ops.layernorm(h16(2 * 16 * 24, 1280), h16(1280), h16(1280), pe=f32(32, 1280),
                                                  rows_per_frame=24, frames=16, pe_start=3))
run("softmax rows", lambda: ops.softmax_rows(f32(300, 1536), 0.5))

# ---- data movement, warp, sampler ----------------------------------------------------------------------------------------
run("upsample", lambda: ops.upsample_nearest(h16(4 * 8 * 12, 640), 4, 8, 12)[0])
run("concat", lambda: ops.concat_channels(h16(1000, 640), h16(1000, 320)))
run("layout", lambda: ops.frames_to_ncfhw(ops.ncfhw_to_frames(f32(2, 5, 3, 6, 7), 8), 2, 5, 3, 6, 7))
run("timestep embedding", lambda: ops.timestep_embedding(torch.tensor([981.0, 1.0], device=dev), 320))
run("warp_image", lambda: ops.warp_image_f32(f32(4, 4, 32, 48), f32(4, 2, 32, 48, scale=6.0)))
run("resize_flow", lambda: ops.resize_flow_f32(f32(4, 2, 256, 384), 32, 48))


def sampler(mode):
    F_, C, h, w, R, Q = 6, 4, 16, 24, 2, 4
    hw, n = h * w, F_ * C * h * w
    table = torch.zeros(4, ops.SAMPLER_ROW, device=dev)
    table[:, 0], table[:, 1], table[:, 2], table[:, 3], table[:, 5] = 500.0, 0.8, 0.6, 0.9, 0.43
    table[:, 6], table[:, 7], table[:, 8], table[:, 9], table[:, 10] = 0.1, 1.0, 7.5, 1.5, 0.3
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    lat2, cond, eps_cfg = f32(2, n), f32(n), torch.empty(n, device=dev)
    x = torch.empty(3 * F_ * hw, 8, device=dev, dtype=torch.float16)
    t = torch.empty(3, device=dev)
    partials = torch.zeros(ops.sampler_partials(F_, hw), 4, dtype=torch.float64, device=dev)
    ops.sampler_begin(table, state, lat2, cond, x, t, F_, C, hw, 8)
    ops.sampler_combine(table, state, f32(3 * F_ * hw, 4), eps_cfg, partials, F_, C, hw)
    ops.sampler_update(table, state, lat2, eps_cfg, partials, mode, f32(R * C * hw) if mode else None,
                       f32(Q * R * 2 * hw, scale=3.0) if mode == 2 else None, f32(2, n), torch.empty(2, n, device=dev),
                       torch.empty(2, n, device=dev), F_, C, R if mode else 0, Q if mode == 2 else 0, h, w)
    assert int(state[0]) == 1
    return lat2


for mode in (0, 1, 2):
    run(f"sampler step mode {mode}", lambda mode=mode: sampler(mode))
print(f"ok: {len(done)} kernel families launched")
for d_ in done:
    print("  ", d_)
