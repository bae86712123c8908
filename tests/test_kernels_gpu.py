"""Per-kernel parity on the GPU: every C-ABI entry point against a plain PyTorch fp32 statement of the same op
(the op-level definitions are the ones oracle/insv2v_oracle.py uses; tolerances are written next to each check).
Run with `pytest -m gpu`."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-4  # BASELINE.json north_star: rtol=1e-3 / atol=1e-4 (fp16 outputs)


@pytest.fixture(scope="module", autouse=True)
def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from insv2v_b200 import lib
    lib.load()
    yield


def _ops():
    from insv2v_b200 import ops
    return ops


def report(name, got, ref, rtol=RTOL, atol=ATOL, frac_ok=1.0):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).float().mean().item()
    print(f"[{name}] max_abs={err.max().item():.3e} max_ref={ref.abs().max().item():.3e} "
          f"viol_frac={bad:.3e} mean_abs={err.mean().item():.3e}")
    assert torch.isfinite(got).all(), f"{name}: non-finite output"
    assert bad <= 1.0 - frac_ok, f"{name}: {bad:.3e} of elements outside rtol={rtol} atol={atol}"


def h16(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).half()


# ------------------------------------------------------------------------------------------------ GEMM / linear
@pytest.mark.parametrize("rows,k,n", [(128, 64, 128), (300, 320, 320), (1000, 1280, 640), (77 * 3, 768, 2560),
                                      (3, 320, 1280), (4608, 2560, 1280), (130, 72, 40),
                                      # short-K residual GEMMs: double staging slab (K <= 640, N % 160 == 0)
                                      (73728, 320, 320), (40000, 320, 960), (18432, 640, 640), (1000, 640, 320),
                                      (50000, 256, 128)])
def test_linear(rows, k, n):
    ops = _ops()
    x = h16(rows, k, seed=1)
    w = h16(n, k, scale=k ** -0.5, seed=2)
    b = h16(n, seed=3)
    res = h16(rows, n, seed=4)
    out = ops.linear(x, ops.pack_linear(w), bias=b, residual=res)
    ref = x.float() @ w.float().t() + b.float() + res.float()
    report(f"linear {rows}x{k}x{n}", out, ref)


@pytest.mark.parametrize("rows,k,n", [(73728, 320, 320), (40000, 320, 960), (50000, 256, 128), (18432, 640, 1920),
                                      (300, 320, 320)])
def test_linear_no_residual(rows, k, n):
    """Same GEMMs without the residual term: the weight-stationary mode (short K, many row tiles) and the plain pair
    kernel."""
    ops = _ops()
    x = h16(rows, k, seed=1)
    w = h16(n, k, scale=k ** -0.5, seed=2)
    b = h16(n, seed=3)
    out = ops.linear(x, ops.pack_linear(w), bias=b)
    ref = x.float() @ w.float().t() + b.float()
    report(f"linear (no residual) {rows}x{k}x{n}", out, ref)


@pytest.mark.parametrize("rows,k,n,bias,res,relu", [
    (4608, 1280, 3840, False, False, False),   # QKV projection, no bias: slab_free recycling path
    (1152, 1280, 1280, True, True, False),     # 9 row tiles: odd pair count (ghost tile), residual prefetch of 2 slabs
    (333, 320, 640, True, False, True),        # ragged last tile + ReLU
    (20000, 640, 1920, False, True, False),    # residual without bias
    (256, 64, 160, True, True, False),         # exactly one tile pair, one K block
    (129, 1280, 320, True, True, False),       # second CTA of the pair holds a single valid row
])
def test_linear_pair160_variants(rows, k, n, bias, res, relu):
    """The v3 pair kernel (16-warp epilogue + store warp, csrc/gemm_tc.cu gemm_tc_pair160_kernel): every combination of
    bias / residual / ReLU, ragged and ghost tiles, and operands that are column slices of wider buffers."""
    ops = _ops()
    xw = h16(rows, k + 64, seed=1)
    x = xw[:, 32:32 + k]                              # a_ld = k + 64
    w = h16(n, k, scale=k ** -0.5, seed=2)
    b = h16(n, seed=3) if bias else None
    rw = h16(rows, n + 16, seed=4)
    r = rw[:, 8:8 + n] if res else None               # res_ld = n + 16
    ow = torch.zeros(rows, n + 24, device="cuda", dtype=torch.float16)
    out = ops.gemm(x, ops.pack_linear(w), n_img=1, h=1, w=rows, c=k, bias=b, residual=r, relu=relu, out=ow[:, 16:16 + n])
    ref = x.float() @ w.float().t()
    if bias:
        ref = ref + b.float()
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.clamp_min(0)
    report(f"pair160 {rows}x{k}x{n} bias={bias} res={res} relu={relu}", out, ref)
    assert float(ow[:, :16].abs().max()) == 0 and float(ow[:, 16 + n:].abs().max()) == 0  # nothing outside the window
    # same call twice in a row on one stream (barrier phases / slab recycling across launches)
    out2 = ops.gemm(x, ops.pack_linear(w), n_img=1, h=1, w=rows, c=k, bias=b, residual=r, relu=relu)
    assert torch.equal(out2, out.contiguous())


@pytest.mark.parametrize("rows,k,n,bias,res,relu", [
    (73728, 320, 960, False, False, False),        # the QKV projection of the 32x48 level
    (128 * 301 + 50, 320, 960, True, True, False),  # ragged last M tile, odd number of M tiles (ghost tile of the last pair)
    (40000, 320, 480, True, False, True), (30000, 200, 480, False, True, False),  # three N tiles; K tail (200 = 3 * 64 + 8)
    (40000, 256, 1600, True, True, False)])
def test_linear_activation_stationary(rows, k, n, bias, res, relu):
    """Short-K pair kernel, activation-stationary mode (K <= 320, >= 3 N tiles, >= 74 M pairs): a cluster keeps the 2 x 128
    activation rows of an M pair in shared memory and walks all its N tiles; only weight tiles stream."""
    from insv2v_b200 import lib
    ops = _ops()
    x = h16(rows, k, seed=1)
    w = h16(n, k, scale=k ** -0.5, seed=2)
    b = h16(n, seed=3) if bias else None
    r = h16(rows, n, seed=4) if res else None
    out = ops.gemm(x, ops.pack_linear(w), n_img=1, h=1, w=rows, c=k, bias=b, residual=r, relu=relu)
    assert lib.load().ivv_debug_last_gemm_as() == 1
    ref = x.float() @ w.float().t()
    if bias:
        ref = ref + b.float()
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.clamp_min(0)
    report(f"AS {rows}x{k}x{n} bias={bias} res={res} relu={relu}", out, ref)
    out2 = ops.gemm(x, ops.pack_linear(w), n_img=1, h=1, w=rows, c=k, bias=b, residual=r, relu=relu)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("rows,c,n,frames,hw", [(4608, 320, 960, 0, 0), (20000, 320, 960, 0, 0),
                                              (3 * 16 * 512, 320, 960, 16, 512),  # activation-stationary consumers (1000, 640, 640, 0, 0), (1152, 1280, 3840, 0, 0),
                                              (2 * 5 * 384, 320, 960, 5, 384), (3 * 16 * 128, 640, 1920, 16, 128)])
def test_layernorm_folded_into_gemms(rows, c, n, frames, hw):
    """K8 fold: the producer GEMM emits per-row partial statistics of its output y, the consumer GEMM computes
    LayerNorm(y) (+ pe[frame]) W^T from the RAW y (ops.pack_ln_linear). Against fp32 torch; the unfused product path
    (LayerNorm kernel -> fp16 -> GEMM) is measured beside it and must not be more accurate by more than 10 %."""
    ops = _ops()
    assert ops.ln_fold_ok(rows, c, c) and ops.ln_fold_ok(rows, c, n)
    x0 = h16(rows, c, seed=1)
    wp = h16(c, c, scale=c ** -0.5, seed=2)
    bp = h16(c, seed=3)
    res = h16(rows, c, seed=4) * 2 + 0.75          # non-zero row means: the mean * wsum term matters
    gamma, beta = (1 + 0.2 * h16(c, seed=5)).float(), 0.3 * h16(c, seed=6).float()
    w = h16(n, c, scale=c ** -0.5, seed=7)
    pe = torch.randn(32, c, device="cuda") * 0.5 if frames else None
    # producer: y = x0 Wp^T + b + res, with row statistics
    stats = ops.row_stats(rows, c, x0.device)
    y = ops.linear(x0, ops.pack_linear(wp), bias=bp, residual=res, row_stats_out=stats)
    y32 = x0.float() @ wp.float().t() + bp.float() + res.float()
    s_ref = torch.stack([y32.sum(1), (y32 * y32).sum(1)], dim=1)
    s_got = stats.sum(dim=1)
    assert torch.allclose(s_got, s_ref, rtol=2e-4, atol=1e-2), (s_got[:2], s_ref[:2])
    # consumer
    wfold, bfold, wsum, table = ops.pack_ln_linear(gamma, beta, w, None, pe)
    kw = {}
    if frames:
        kw = dict(rowbias=table[3:3 + frames], rowbias_group=hw, rowbias_mod=frames)
    out = ops.linear(y, wfold, bias=bfold, ln=(stats, wsum, 1e-5), **kw)
    # truth: LayerNorm of the fp16 tensor the consumer actually reads, in fp32
    ln = F.layer_norm(y.float(), (c,), gamma, beta, 1e-5)
    if frames:
        fidx = (torch.arange(rows, device="cuda") // hw) % frames
        ln = ln + pe[3 + fidx]
    ref = ln @ w.float().t()
    report(f"LN fold {rows}x{c}->{n}", out, ref, rtol=2e-3, atol=2e-3)
    # the unfused path of the product on the same inputs
    g16, b16 = gamma.half(), beta.half()
    nrm = ops.layernorm(y, g16, b16, eps=1e-5, **(dict(pe=pe, rows_per_frame=hw, frames=frames, pe_start=3) if frames else {}))
    unf = ops.linear(nrm, ops.pack_linear(w))
    e_f = (out.float() - ref).norm() / ref.norm()
    e_u = (unf.float() - ref).norm() / ref.norm()
    print(f"  rel-L2 vs fp32: folded {e_f:.3e}, LayerNorm kernel + GEMM {e_u:.3e}")
    assert e_f <= 1.1 * e_u + 1e-5


@pytest.mark.parametrize("rows,c,mult,bias", [(500, 320, 8, True), (40000, 320, 8, True), (40000, 320, 8, False),
                                              (9100, 640, 8, True), (2400, 1280, 4, True),
                                              # activation-stationary mode (K <= 320, >= 74 M pairs): ragged and ghost tiles,
                                              # the full-size call, a K tail (72 = 64 + 8), three N tiles
                                              (128 * 301 + 50, 320, 8, True), (73728, 320, 8, True), (30000, 72, 32, True),
                                              (25000, 192, 4, False)])
def test_linear_geglu(rows, c, mult, bias):
    """GEGLU epilogue (hidden * gelu(gate), tile-interleaved weights). The larger cases give every cluster a run of
    tiles: the bias slices staged one tile ahead in shared memory and the two alternating output slabs (K <= 320) are
    only exercised across tile boundaries."""
    ops = _ops()
    x = h16(rows, c, seed=1)
    w = h16(mult * c, c, scale=c ** -0.5, seed=2)
    b = h16(mult * c, scale=0.5, seed=3) if bias else None
    wp, bp = ops.pack_geglu(w, b if bias else torch.zeros(mult * c, device="cuda", dtype=torch.float16))
    out = ops.linear(x, wp, bias=bp if bias else None, geglu=True)
    from insv2v_b200 import lib
    assert lib.load().ivv_debug_last_gemm_as() == int(c <= 320 and rows > 146 * 128 and mult * c >= 768)
    assert torch.equal(out, ops.linear(x, wp, bias=bp if bias else None, geglu=True))
    y = x.float() @ w.float().t()
    if bias:
        y = y + b.float()
    hid, gate = y.chunk(2, dim=-1)
    ref = hid * F.gelu(gate)
    assert out.shape == (rows, mult * c // 2)
    report(f"geglu {rows}x{c}->{mult * c // 2} bias={bias}", out, ref)


def test_linear_out_f32_and_rowbias():
    ops = _ops()
    rows, k, n = 2 * 4 * 24, 128, 72
    x = h16(rows, k, seed=1)
    w = h16(n, k, scale=k ** -0.5, seed=2)
    rb = h16(2, n, seed=5)
    out = ops.gemm(x, ops.pack_linear(w), n_img=1, h=1, w=rows, c=k, rowbias=rb, rowbias_group=rows // 2,
                   out_f32=True)
    ref = x.float() @ w.float().t() + rb.float().repeat_interleave(rows // 2, dim=0)
    assert out.dtype == torch.float32
    report("linear f32+rowbias", out, ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ conv
def _frames(x_nchw):
    n, c, h, w = x_nchw.shape
    return x_nchw.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def _nchw(fr, n, h, w):
    return fr.reshape(n, h, w, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize("n,ci,co,h,w", [(2, 64, 64, 16, 16), (6, 320, 320, 32, 48), (4, 8, 320, 32, 48),
                                         (5, 640, 1280, 8, 12), (48, 1280, 1280, 4, 6), (3, 128, 8, 12, 20),
                                         (2, 192, 320, 5, 7),
                                         # halo kernel (one activation box per filter column): 8x16 box with 160-wide
                                         # tiles, 256- and 128-wide tiles, ragged tiles in both directions, an odd
                                         # number of M tiles (ghost tile of the CTA pair), channel tail (ci % 64 != 0)
                                         (3, 640, 640, 16, 24), (2, 128, 512, 32, 48), (2, 64, 128, 24, 40),
                                         (1, 96, 320, 20, 44), (3, 200, 256, 8, 16)])
def test_conv3x3(n, ci, co, h, w):
    ops = _ops()
    x = h16(n, ci, h, w, seed=1)
    wt = h16(co, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    b = h16(co, seed=3)
    out = ops.conv3x3(_frames(x), ops.pack_conv3x3(wt), n, h, w, bias=b)
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=1)  # fp64: no reference-side rounding
    report(f"conv3x3 n{n} {ci}->{co} {h}x{w}", _nchw(out, n, h, w), ref)


@pytest.mark.parametrize("h,w", [(8, 12), (16, 24), (32, 48)])  # the two larger ones take the halo kernel
def test_conv3x3_fused_temb_residual(h, w):
    ops = _ops()
    b_, f, ci, co = 2, 4, 128, 192
    n = b_ * f
    x = h16(n, ci, h, w, seed=1)
    wt = h16(co, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    bias = h16(co, seed=3)
    temb = h16(b_, co, seed=4)
    res = h16(n, co, h, w, seed=5)
    out = ops.conv3x3(_frames(x), ops.pack_conv3x3(wt), n, h, w, bias=bias, rowbias=temb, rowbias_group=f * h * w,
                      residual=_frames(res))
    ref = F.conv2d(x.float(), wt.float(), bias.float(), padding=1)
    ref = ref + temb.float().repeat_interleave(f, dim=0)[:, :, None, None] + res.float()
    report("conv3x3+temb+res", _nchw(out, n, h, w), ref)


@pytest.mark.parametrize("n,c,co,h,w", [(4, 320, 320, 32, 48), (3, 64, 128, 9, 13)])
def test_conv3x3_stride2(n, c, co, h, w):
    ops = _ops()
    x = h16(n, c, h, w, seed=1)
    wt = h16(co, c, 3, 3, scale=(9 * c) ** -0.5, seed=2)
    b = h16(co, seed=3)
    out, ho, wo = ops.conv3x3_s2(_frames(x), ops.pack_conv3x3_im2col(wt), n, h, w, bias=b)
    ref = F.conv2d(x.float(), wt.float(), b.float(), stride=2, padding=1)
    assert (ho, wo) == tuple(ref.shape[-2:])
    report("conv3x3 s2", _nchw(out, n, ho, wo), ref)


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("b,f,c,h,w,fpg,silu", [(3, 16, 320, 32, 48, 16, True), (2, 4, 640, 8, 12, 1, False),
                                                (1, 2, 2560, 4, 6, 2, True), (2, 3, 128, 5, 7, 3, True),
                                                (1, 4, 512, 16, 16, 1, True),
                                                # the UNet's small levels at full size: the one-kernel form (whole tensor
                                                # in shared memory, csrc/norm.cu gn_fused_kernel)
                                                (3, 16, 1280, 8, 12, 16, True), (3, 16, 1280, 8, 12, 1, True),
                                                (3, 16, 640, 16, 24, 1, True), (3, 16, 640, 16, 24, 16, False),
                                                (3, 16, 1280, 4, 6, 16, True), (3, 8, 1280, 4, 4, 1, True)])
def test_groupnorm(b, f, c, h, w, fpg, silu):
    ops = _ops()
    x = (h16(b * f, c, h, w, seed=1).float() * 1.5 + 0.3).half()
    g = h16(c, seed=2)
    be = h16(c, seed=3)
    eps = 1e-5
    out = ops.groupnorm(_frames(x), g, be, b * f, h * w, 32, fpg, eps, silu)
    # frames_per_group frames share statistics: reshape to [groups_of_frames, c, fpg, h, w]
    x5 = x.float().reshape(b * f // fpg, fpg, c, h, w).permute(0, 2, 1, 3, 4)
    ref = F.group_norm(x5, 32, g.float(), be.float(), eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    report(f"groupnorm c{c} fpg{fpg}", _nchw(out, b * f, h, w), ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("b,f,c1,c2,h,w,fpg", [(2, 4, 320, 320, 8, 12, 4), (1, 6, 1280, 640, 4, 6, 6), (3, 2, 64, 128, 16, 16, 1),
                                                (1, 16, 640, 320, 16, 24, 16), (3, 16, 1280, 1280, 8, 12, 16),
                                                (3, 16, 1280, 1280, 4, 6, 16)])
def test_groupnorm_two_sources_equals_concat(b, f, c1, c2, h, w, fpg):
    """ivv_groupnorm2 reads the skip concatenation [x1 | x2] in place (unet_blocks.py:561,659 + resnet.py:177): it must
    give BIT-identical results to ivv_groupnorm on the materialised concatenation, and match torch."""
    ops = _ops()
    x1 = (h16(b * f * h * w, c1, seed=1).float() * 1.5 + 0.3).half()
    x2 = (h16(b * f * h * w, c2, seed=2).float() * 0.7 - 0.2).half()
    c = c1 + c2
    g, be = h16(c, seed=3), h16(c, seed=4)
    out2 = ops.groupnorm2(x1, x2, g, be, b * f, h * w, 32, fpg, 1e-5, True)
    cat = ops.concat_channels(x1, x2)
    assert torch.equal(cat, torch.cat([x1, x2], dim=1))
    out1 = ops.groupnorm(cat, g, be, b * f, h * w, 32, fpg, 1e-5, True)
    assert torch.equal(out2, out1), "two-source GroupNorm differs from GroupNorm of the concatenation"
    x5 = cat.float().reshape(b * f // fpg, fpg, h, w, c).permute(0, 4, 1, 2, 3)
    ref = F.silu(F.group_norm(x5, 32, g.float(), be.float(), 1e-5)).permute(0, 2, 3, 4, 1).reshape(-1, c)
    report(f"groupnorm2 {c1}+{c2} fpg{fpg}", out2, ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("rows,c", [(1000, 320), (77, 640), (4608, 1280)])
def test_layernorm(rows, c):
    ops = _ops()
    x = h16(rows, c, seed=1)
    g, be = h16(c, seed=2), h16(c, seed=3)
    out = ops.layernorm(x, g, be)
    ref = F.layer_norm(x.float(), (c,), g.float(), be.float(), 1e-5)
    report(f"layernorm {rows}x{c}", out, ref, rtol=1e-3, atol=1e-3)


def test_layernorm_pe():
    ops = _ops()
    clips, frames, hw, c = 2, 8, 24, 320
    x = h16(clips * frames * hw, c, seed=1)
    g, be = h16(c, seed=2), h16(c, seed=3)
    pe = torch.randn(32, c, device="cuda")
    out = ops.layernorm(x, g, be, pe=pe, rows_per_frame=hw, frames=frames, pe_start=3)
    ref = F.layer_norm(x.float(), (c,), g.float(), be.float(), 1e-5).reshape(clips, frames, hw, c)
    ref = ref + pe[3:3 + frames][None, :, None, :]
    report("layernorm+pe", out, ref.reshape(-1, c), rtol=1e-3, atol=1e-3)


# ------------------------------------------------------------------------------------------------ attention
def _sdpa_ref(q, k, v, heads):
    # q [n, sq, h*d], k/v [n, skv, h*d] fp32
    n, sq, c = q.shape
    d = c // heads
    qh = q.reshape(n, sq, heads, d).transpose(1, 2)
    kh = k.reshape(n, -1, heads, d).transpose(1, 2)
    vh = v.reshape(n, -1, heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(n, sq, c)


@pytest.mark.parametrize("n,s,heads,d", [(2, 128, 2, 64), (3, 1536, 8, 40), (4, 384, 8, 80), (6, 96, 8, 160),
                                         (5, 24, 8, 160), (2, 200, 4, 40)])
def test_self_attention(n, s, heads, d):
    ops = _ops()
    c = heads * d
    qkv = h16(n * s, 3 * c, seed=1)
    out = ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], n_batch=n, s_q=s, s_kv=s, heads=heads, d=d,
                        q_ld=3 * c, kv_ld=3 * c)
    q, k, v = (t.float().reshape(n, s, c) for t in qkv.chunk(3, dim=-1))
    ref = _sdpa_ref(q, k, v, heads)
    # P is rounded to fp16 before the PV product (as in every fp16 tensor-core flash attention, incl. the reference's
    # xformers/SDPA path): allow 2e-3 relative + 5e-4 absolute against the fp32 truth.
    report(f"self-attn n{n} s{s} h{heads} d{d}", out.reshape(n, s, c), ref, rtol=2e-3, atol=5e-4)


@pytest.mark.parametrize("s_q,s_kv,d,gain", [(1536, 1536, 40, 3.0), (640, 900, 40, 4.0), (300, 1100, 24, 2.0),
                                             (256, 257, 56, 1.0), (384, 384, 80, 3.0)])
def test_self_attention_peaky(s_q, s_kv, d, gain):
    """Large score ranges (|q.k|*scale up to ~50): the running maximum jumps, so the lazy O rescale, the masked last
    block, the ghost tile of an odd CTA pair and the FMA-pipe exp2 (arguments down to -125) are all exercised."""
    ops = _ops()
    n, heads = 2, 4
    c = heads * d
    q = h16(n * s_q, c, scale=gain, seed=1)
    kv = h16(n * s_kv, 2 * c, scale=gain, seed=2)
    out = ops.attention(q, kv[:, :c], kv[:, c:], n_batch=n, s_q=s_q, s_kv=s_kv, heads=heads, d=d, q_ld=c, kv_ld=2 * c)
    k, v = (t.float().reshape(n, s_kv, c) for t in kv.chunk(2, dim=-1))
    ref = _sdpa_ref(q.float().reshape(n, s_q, c), k, v, heads)
    report(f"peaky self-attn sq{s_q} skv{s_kv} d{d}", out.reshape(n, s_q, c), ref, rtol=2e-3, atol=5e-4 * gain)


@pytest.mark.parametrize("n,s_q,s_kv,heads,d,kv_div,gain", [
    (48, 384, 384, 8, 80, 1, 1.0),     # UNet level 1 self-attention: 1 152 items of 3 key blocks
    (48, 96, 96, 8, 160, 1, 1.0),      # level 2: one block per item, one-stage K/V ring
    (48, 384, 77, 8, 80, 16, 1.0),     # cross-attention: masked single block, K/V shared by 16 frames
    (48, 96, 77, 8, 160, 16, 1.0),
    (40, 200, 300, 8, 80, 1, 2.0),     # ragged query tile, masked last block, jumping row maxima (O rescale)
    (30, 130, 129, 8, 160, 1, 2.0),    # two query tiles / two key blocks with one valid row / key in the second
    (150, 24, 24, 8, 160, 1, 1.0),     # 4x6 level shape, more items than SMs
])
def test_attention_persistent_one_tile(n, s_q, s_kv, heads, d, kv_div, gain):
    """d = 80 / 160 with more (query tile, head, frame) items than SMs: the persistent one-tile kernel
    (csrc/attention_tc.cu attention_persist1_kernel; Q / O double buffering and barrier phases across items)."""
    ops = _ops()
    c = heads * d
    q = h16(n * s_q, c, scale=gain, seed=1)
    kv = h16(n // kv_div * s_kv, 2 * c, scale=gain, seed=2)
    out = ops.attention(q, kv[:, :c], kv[:, c:], n_batch=n, s_q=s_q, s_kv=s_kv, heads=heads, d=d, q_ld=c, kv_ld=2 * c,
                        kv_div=kv_div)
    k, v = (t.float().reshape(n // kv_div, s_kv, c).repeat_interleave(kv_div, dim=0) for t in kv.chunk(2, dim=-1))
    ref = _sdpa_ref(q.float().reshape(n, s_q, c), k, v, heads)
    report(f"persistent one-tile attn n{n} sq{s_q} skv{s_kv} d{d}", out.reshape(n, s_q, c), ref, rtol=2e-3, atol=5e-4 * gain)
    out2 = ops.attention(q, kv[:, :c], kv[:, c:], n_batch=n, s_q=s_q, s_kv=s_kv, heads=heads, d=d, q_ld=c, kv_ld=2 * c,
                         kv_div=kv_div)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("clips,frames,s,heads,d", [(3, 4, 384, 8, 80), (2, 3, 1536, 8, 40), (2, 2, 24, 8, 160)])
def test_cross_attention(clips, frames, s, heads, d):
    ops = _ops()
    c = heads * d
    n = clips * frames
    q = h16(n * s, c, seed=1)
    kv = h16(clips * 77, 2 * c, seed=2)
    out = ops.attention(q, kv[:, :c], kv[:, c:], n_batch=n, s_q=s, s_kv=77, heads=heads, d=d, q_ld=c, kv_ld=2 * c,
                        kv_div=frames)
    k, v = (t.float().reshape(clips, 77, c).repeat_interleave(frames, dim=0) for t in kv.chunk(2, dim=-1))
    ref = _sdpa_ref(q.float().reshape(n, s, c), k, v, heads)
    report(f"cross-attn s{s} d{d}", out.reshape(n, s, c), ref, rtol=2e-3, atol=5e-4)


@pytest.mark.parametrize("clips,frames,hw,heads,d", [(3, 16, 96, 8, 40), (2, 16, 24, 8, 160), (1, 5, 35, 8, 80),
                                                     (1, 32, 6, 8, 160), (1, 24, 12, 8, 40),
                                                     # long-form capture of configs[4]: 64 frames per UNet call
                                                     (1, 64, 10, 8, 40), (2, 48, 6, 8, 80), (1, 64, 4, 8, 160)])
def test_temporal_attention(clips, frames, hw, heads, d):
    ops = _ops()
    c = heads * d
    qkv = h16(clips * frames * hw, 3 * c, seed=1)
    out = ops.temporal_attention(qkv, clips, frames, hw, c, heads)
    t = qkv.float().reshape(clips, frames, hw, 3 * c).permute(0, 2, 1, 3).reshape(clips * hw, frames, 3 * c)
    q, k, v = t.chunk(3, dim=-1)
    ref = _sdpa_ref(q, k, v, heads).reshape(clips, hw, frames, c).permute(0, 2, 1, 3).reshape(-1, c)
    report(f"temporal-attn f{frames} d{d}", out, ref)


def test_softmax_rows():
    ops = _ops()
    x = h16(300, 1536, scale=3.0, seed=1)
    out = ops.softmax_rows(x, 0.5)
    report("softmax_rows", out, torch.softmax(x.float() * 0.5, dim=-1), rtol=1e-3, atol=1e-6)


# ------------------------------------------------------------------------------------------------ glue
def test_layout_roundtrip_and_glue():
    ops = _ops()
    x = torch.randn(2, 5, 3, 6, 7, device="cuda")
    fr = ops.ncfhw_to_frames(x, 8)
    assert fr.shape == (2 * 3 * 42, 8)
    ref = x.permute(0, 2, 3, 4, 1).reshape(-1, 5)
    assert torch.equal(fr[:, :5], ref.half()) and (fr[:, 5:] == 0).all()
    back = ops.frames_to_ncfhw(fr, 2, 5, 3, 6, 7)
    assert torch.equal(back, x.half().float())
    a, b = h16(100, 64, seed=1), h16(100, 128, seed=2)
    assert torch.equal(ops.concat_channels(a, b), torch.cat([a, b], dim=1))
    img = h16(3, 64, 5, 7, seed=3)
    up, ho, wo = ops.upsample_nearest(_frames(img), 3, 5, 7)
    assert torch.equal(_nchw(up, 3, ho, wo), F.interpolate(img.float(), scale_factor=2.0, mode="nearest").half())
    up2, ho, wo = ops.upsample_nearest(_frames(img), 3, 5, 7, 9, 15)
    assert torch.equal(_nchw(up2, 3, 9, 15), F.interpolate(img.float(), size=(9, 15), mode="nearest").half())
    # rows longer than a CTA (several vector columns per thread) and a ragged last row chunk
    a2, b2 = h16(77, 1280, seed=4), h16(77, 1536, seed=5)
    assert torch.equal(ops.concat_channels(a2, b2), torch.cat([a2, b2], dim=1))
    img2 = h16(2, 2304, 3, 4, seed=6)
    up3, ho, wo = ops.upsample_nearest(_frames(img2), 2, 3, 4)
    assert torch.equal(_nchw(up3, 2, ho, wo), F.interpolate(img2.float(), scale_factor=2.0, mode="nearest").half())
    report("silu", ops.silu(a), F.silu(a.float()))
    report("scale", ops.scale(a, 1 / 0.18215), a.float() / 0.18215)


def test_timestep_embedding():
    ops = _ops()
    t = torch.tensor([981.0, 1.0, 500.0], device="cuda")
    out = ops.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    arg = t[:, None] * freqs[None]
    ref = torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)
    report("timestep", out, ref, rtol=1e-3, atol=1e-3)


# ------------------------------------------------------------------------------------------------ flow warp
def _warp_ref(image, flow):
    n, c, h, w = image.shape
    ys, xs = torch.meshgrid(torch.arange(h, device=image.device), torch.arange(w, device=image.device), indexing="ij")
    grid = torch.stack([xs, ys], dim=-1).float()[None].repeat(n, 1, 1, 1) + flow.permute(0, 2, 3, 1)
    grid[..., 0] = 2 * (grid[..., 0] / (w - 1) - 0.5)
    grid[..., 1] = 2 * (grid[..., 1] / (h - 1) - 0.5)
    return F.grid_sample(image, grid, mode="bilinear", align_corners=True)


def test_warp_and_resize_flow():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.randn(4, 4, 32, 48, device="cuda", generator=g)
    flow = torch.randn(4, 2, 32, 48, device="cuda", generator=g) * 6
    report("warp_image", ops.warp_image_f32(img, flow), _warp_ref(img, flow), rtol=1e-4, atol=1e-4)
    big = torch.randn(4, 2, 256, 384, device="cuda", generator=g) * 5
    scaled = big.clone()
    scaled[:, 0] *= 48 / 384
    scaled[:, 1] *= 32 / 256
    ref = F.interpolate(scaled, size=(32, 48), mode="bilinear", align_corners=False)
    report("resize_flow /8", ops.resize_flow_f32(big, 32, 48), ref, rtol=1e-5, atol=1e-5)
    ref2 = F.interpolate(big[:, :, :100, :90] * 1.0, size=(37, 53), mode="bilinear", align_corners=False)
    b2 = big[:, :, :100, :90].contiguous()
    got2 = ops.resize_flow_f32(b2, 37, 53)
    ref2[:, 0] *= 1.0
    s = b2.clone()
    s[:, 0] *= 53 / 90
    s[:, 1] *= 37 / 100
    report("resize_flow general", got2, F.interpolate(s, size=(37, 53), mode="bilinear", align_corners=False),
           rtol=1e-5, atol=1e-5)


def test_flow_noise_correction():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    Q, R, C, h, w = 12, 4, 4, 32, 48
    delta = torch.randn(R, C, h, w, device="cuda", generator=g)
    flow = torch.randn(Q, R, 2, h, w, device="cuda", generator=g) * 8
    eps = torch.randn(Q, C, h, w, device="cuda", generator=g)
    ref = eps.clone()
    for q in range(Q):
        wd = _warp_ref(delta, flow[q])
        m = _warp_ref(torch.ones_like(delta[:, :1]), flow[q])
        msum = m.sum(dim=0, keepdim=True)
        corr = torch.where(msum > 0.5, wd.sum(dim=0, keepdim=True) / msum, torch.zeros_like(msum))
        ref[q:q + 1] += torch.where(msum > 0.5, corr, torch.zeros_like(corr))
    got = ops.flow_noise_correction_(eps.clone(), delta, flow)
    report("flow_noise_correction", got, ref, rtol=1e-4, atol=1e-4, frac_ok=0.9999)


def test_cfg_ddim_step():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    eps3 = torch.randn(3, 1000, device="cuda", generator=g)
    lat = torch.randn(1000, device="cuda", generator=g)
    at, ap = 0.31, 0.42
    e = eps3[0] + 1.5 * (eps3[1] - eps3[0]) + 7.5 * (eps3[2] - eps3[1])
    x0 = (lat - (1 - at) ** 0.5 * e) / at ** 0.5
    ref = ap ** 0.5 * x0 + (1 - ap) ** 0.5 * e
    got = ops.cfg_ddim_step_(eps3, lat.clone(), 7.5, 1.5, at, ap)
    report("cfg_ddim", got, ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ kernel variants
_VARIANT_SNIPPET = r"""
import sys, torch
sys.path.insert(0, %r)
from insv2v_b200 import ops
torch.manual_seed(0)
dev = "cuda"
# conv 3x3 + bias + residual (persistent GEMM variants)
n, ci, co, h, w = 6, 320, 320, 32, 48
x = torch.randn(n, ci, h, w, device=dev).half(); wt = (torch.randn(co, ci, 3, 3, device=dev) * (9 * ci) ** -0.5).half()
b = torch.randn(co, device=dev).half(); res = torch.randn(n, co, h, w, device=dev).half()
fr = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
out = ops.conv3x3(fr(x), ops.pack_conv3x3(wt), n, h, w, bias=b, residual=fr(res))
ref = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=1) + res.double()
err = (out.reshape(n, h, w, co).permute(0, 3, 1, 2).double() - ref).abs()
assert (err <= 1e-4 + 1e-3 * ref.abs()).all(), float(err.max())
# short-K linear + residual (double staging slab / weight-stationary / single-slab variants)
rows, c = 128 * 37 + 50, 320
xl = torch.randn(rows, c, device=dev).half(); wl = (torch.randn(c, c, device=dev) * c ** -0.5).half()
bl = torch.randn(c, device=dev).half(); rl = torch.randn(rows, c, device=dev).half()
o = ops.linear(xl, ops.pack_linear(wl), bias=bl, residual=rl)
refl = xl.double() @ wl.double().t() + bl.double() + rl.double()
assert ((o.double() - refl).abs() <= 1e-4 + 1e-3 * refl.abs()).all()
# odd number of M tiles (ghost tile of a CTA pair) and GEGLU
rows, c = 128 * 5, 320
xl = torch.randn(rows, c, device=dev).half(); wl = (torch.randn(8 * c, c, device=dev) * c ** -0.5).half()
bl = (torch.randn(8 * c, device=dev) * 0.1).half()
wp, bp = ops.pack_geglu(wl, bl)
o = ops.linear(xl, wp, bias=bp, geglu=True)
y = xl.double() @ wl.double().t() + bl.double(); hid, gate = y.chunk(2, dim=-1)
refg = hid * torch.nn.functional.gelu(gate)
assert ((o.double() - refg).abs() <= 1e-4 + 1e-3 * refg.abs()).all()
# self-attention S=640, d=40
nb, s, heads, d = 2, 640, 8, 40
qkv = torch.randn(nb * s, 3 * heads * d, device=dev).half(); C = heads * d
a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n_batch=nb, s_q=s, s_kv=s, heads=heads, d=d, q_ld=3 * C, kv_ld=3 * C)
q, k, v = (t.float().reshape(nb, s, heads, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
ra = (torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v).transpose(1, 2).reshape(nb * s, C)
assert ((a.float() - ra).abs() <= 5e-4 + 2e-3 * ra.abs()).all()
print("variant ok")
"""


@pytest.mark.parametrize("env", [{"IVV_PAIR": "0"}, {"IVV_PAIR": "0", "IVV_CLUSTER": "2"},
                                 {"IVV_ATTN_PAIR": "0", "IVV_ATTN_TWO_TILE": "1"}, {"IVV_ATTN_PAIR": "0"},
                                 {"IVV_ATTN_PAIR": "0", "IVV_ATTN_QK_FIRST": "0"}, {"IVV_ATTN_MODE": "0"},
                                 {"IVV_ATTN_MODE": "1"}, {"IVV_ATTN_MODE": "2"}, {"IVV_ATTN_POLY": "1"},
                                 {"IVV_ATTN_PAIR_SHORT": "0"},
                                 {"IVV_ATTN_MODE": "0", "IVV_ATTN_POLY": "1"},
                                 {"IVV_FORCE_BN": "128"}, {"IVV_FORCE_BN": "256"}, {"IVV_HALO": "0"}, {"IVV_DS": "0"},
                                 {"IVV_DS": "0", "IVV_NO_WS": "1"}])
def test_kernel_variants(env):
    """The opt-in / fallback code paths (single-CTA GEMM, multicast clusters, two-tile attention, other tile widths)
    stay correct: same checks in a subprocess with the tuning environment variables set."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _VARIANT_SNIPPET % root], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("n,ci,co,h,w", [(48, 1280, 1280, 4, 6), (48, 2560, 1280, 4, 6), (3, 640, 320, 8, 12)])
def test_conv3x3_splitk(n, ci, co, h, w):
    """Few-row / long-K convolutions go through the split-K path (fp32 partial planes + deterministic reduce) with the
    same fused terms as the direct epilogue."""
    ops = _ops()
    assert ops._splitk_policy(n * h * w, 9 * ci, co) > 1
    x = h16(n, ci, h, w, seed=1)
    wt = h16(co, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    b, res = h16(co, seed=3), h16(n, co, h, w, seed=5)
    temb = h16(3, co, seed=4)
    f = n // 3
    out = ops.conv3x3(_frames(x), ops.pack_conv3x3(wt), n, h, w, bias=b, rowbias=temb, rowbias_group=f * h * w,
                      residual=_frames(res))
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=1) + res.double() \
        + temb.double().repeat_interleave(f, dim=0)[:, :, None, None]
    report(f"conv3x3 split-K {ci}->{co} {h}x{w}", _nchw(out, n, h, w), ref)
    out2 = ops.conv3x3(_frames(x), ops.pack_conv3x3(wt), n, h, w, bias=b, rowbias=temb, rowbias_group=f * h * w,
                       residual=_frames(res))
    assert torch.equal(out, out2)  # fixed-order reduction: bit-reproducible


@pytest.mark.parametrize("n,ci,co,h,w,fused", [(48, 1280, 1280, 8, 12, True), (48, 640, 1280, 8, 12, False),
                                               (48, 320, 1280, 8, 12, True)])
def test_conv3x3_wide_tiles(n, ci, co, h, w, fused):
    """The N = 1280 convolutions of the 8x12 level run as 320-wide tiles (two N = 160 MMAs per k-step on one
    accumulator): 72 tiles = one round of the 74 clusters. Same fused terms as every other epilogue."""
    from insv2v_b200 import lib
    ops = _ops()
    x = h16(n, ci, h, w, seed=1)
    wt = h16(co, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    b, res, temb = h16(co, seed=3), h16(n, co, h, w, seed=5), h16(3, co, seed=4)
    f = n // 3
    kw = dict(rowbias=temb, rowbias_group=f * h * w, residual=_frames(res)) if fused else {}
    out = ops.conv3x3(_frames(x), ops.pack_conv3x3(wt), n, h, w, bias=b, **kw)
    assert lib.load().ivv_debug_last_gemm_tile() == 320
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=1)
    if fused:
        ref = ref + res.double() + temb.double().repeat_interleave(f, dim=0)[:, :, None, None]
    report(f"conv3x3 wide {ci}->{co} {h}x{w}", _nchw(out, n, h, w), ref)
    out2 = ops.conv3x3(_frames(x), ops.pack_conv3x3(wt), n, h, w, bias=b, **kw)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("rows,k,n,res", [(4608, 5120, 1280, True), (18432, 2560, 640, True), (4608, 2560, 1280, False),
                                          (128 * 35 + 50, 2624, 1280, True)])
def test_linear_wide_tiles(rows, k, n, res):
    """FF out-projections (4608 x 5120 -> 1280, 18432 x 2560 -> 640, + residual) and the 1x1 shortcuts of the 8x12 level:
    320-wide tiles whenever they save rounds of the 74 clusters. Residual read straight from global memory one chunk
    ahead; ragged last tile (36 M tiles, the last one with 50 rows)."""
    from insv2v_b200 import lib
    ops = _ops()
    x, w, b = h16(rows, k, seed=1), h16(n, k, scale=k ** -0.5, seed=2), h16(n, seed=3)
    r = h16(rows, n, seed=4) if res else None
    out = ops.linear(x, ops.pack_linear(w), bias=b, residual=r)
    assert lib.load().ivv_debug_last_gemm_tile() == 320
    ref = x.double() @ w.double().t() + b.double() + (r.double() if res else 0)
    report(f"linear wide {rows}x{k}x{n}", out, ref)
    assert torch.equal(out, ops.linear(x, ops.pack_linear(w), bias=b, residual=r))


_WIDE_SNIPPET = r"""
import sys, torch
sys.path.insert(0, %r)
from insv2v_b200 import ops, lib
torch.manual_seed(0)
dev = "cuda"
fr = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
# several tiles per cluster (the single accumulator's barriers flip every tile), ragged last M tile, odd number of M tiles
for rows, k, n, with_res in [(128 * 157 + 50, 2560, 640, True), (128 * 301, 2624, 320, False), (300, 2560, 960, True)]:
    x = torch.randn(rows, k, device=dev).half(); w = (torch.randn(n, k, device=dev) * k ** -0.5).half()
    b = torch.randn(n, device=dev).half(); r = torch.randn(rows, n, device=dev).half() if with_res else None
    o = ops.linear(x, ops.pack_linear(w), bias=b, residual=r)
    assert lib.load().ivv_debug_last_gemm_tile() == 320, (rows, k, n)
    ref = x.double() @ w.double().t() + b.double() + (r.double() if with_res else 0)
    err = (o.double() - ref).abs()
    assert (err <= 1e-4 + 1e-3 * ref.abs()).all(), (rows, k, n, float(err.max()))
    assert torch.equal(o, ops.linear(x, ops.pack_linear(w), bias=b, residual=r))
# per-tap convolution with a ragged box (5 frames of 8x12: 3.75 M tiles) and a channel tail (ci %% 64 != 0)
n, ci, co, h, w = 5, 328, 640, 8, 12
x = torch.randn(n, ci, h, w, device=dev).half(); wt = (torch.randn(co, ci, 3, 3, device=dev) * (9 * ci) ** -0.5).half()
b = torch.randn(co, device=dev).half(); res = torch.randn(n, co, h, w, device=dev).half()
out = ops.conv3x3(fr(x), ops.pack_conv3x3(wt), n, h, w, bias=b, residual=fr(res))
assert lib.load().ivv_debug_last_gemm_tile() == 320
ref = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=1) + res.double()
err = (out.reshape(n, h, w, co).permute(0, 3, 1, 2).double() - ref).abs()
assert (err <= 1e-4 + 1e-3 * ref.abs()).all(), float(err.max())
print("wide ok")
"""


def test_wide_tiles_forced():
    """IVV_FORCE_BN=320 takes the 320-wide tile wherever it is legal: tile lists longer than the cluster count, ragged
    and ghost tiles, a per-tap convolution with a clipped box."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _WIDE_SNIPPET % root], env=dict(os.environ, IVV_FORCE_BN="320"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "wide ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ------------------------------------------------------------------------------------------------ frame I/O
def test_frame_io_matches_the_reference_transform(tmp_path):
    """video_io: uint8 frames -> [-1, 1] tensors bit-identical to the reference's cv2.cvtColor + ToTensor + Normalize
    (dataset/loveu_tgve_dataset.py:13-16,50-52); tensors -> uint8 identical to `x / 2 + 0.5`, `* 255`, astype(uint8)
    (misc_utils/image_utils.py:130,233-235); the file loader against the same pipeline run with cv2 + torchvision."""
    import cv2
    import numpy as np
    from torchvision import transforms
    from insv2v_b200 import video_io
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(5, 48, 64, 3), dtype=np.uint8)
    frames[0, 0, :4] = [[0, 0, 0], [255, 255, 255], [1, 127, 128], [254, 128, 127]]
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    ref = torch.stack([tf(cv2.cvtColor(f, cv2.COLOR_BGR2RGB)) for f in frames])
    got = video_io.frames_u8_to_tensor(frames)
    assert got.shape == (5, 3, 48, 64) and torch.equal(got.cpu(), ref), "uint8 -> [-1,1] is not bit-identical"
    # way back: the reference's numpy arithmetic
    x = (torch.rand(4, 3, 40, 56) * 2 - 1)
    x[0, :, 0, :3] = torch.tensor([[-1.0, 1.0, 0.0]] * 3)
    want = ((x.numpy().transpose(0, 2, 3, 1) / 2 + 0.5) * 255).astype(np.uint8)
    u8 = video_io.tensor_to_frames_u8(x.cuda()).cpu().numpy()
    assert np.array_equal(u8, want), "[-1,1] -> uint8 differs from the reference arithmetic"
    u8h = video_io.tensor_to_frames_u8(x.half().cuda().unsqueeze(0)).cpu().numpy()
    assert np.abs(u8h.astype(int) - want.astype(int)).max() <= 1
    # file round trip: encode a short clip, load it with the product and with the reference's pipeline
    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 10, (64, 48))
    if not wr.isOpened():
        pytest.skip("no video encoder available in this OpenCV build")
    for f in frames:
        wr.write(f)
    wr.release()
    cap, ref_frames = cv2.VideoCapture(path), []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        ref_frames.append(tf(cv2.cvtColor(cv2.resize(f, (32, 40)), cv2.COLOR_BGR2RGB)))
    cap.release()
    got = video_io.load_video_frames(path, (32, 40))
    assert len(ref_frames) == 5 and torch.equal(got.cpu(), torch.stack(ref_frames))
    video_io.save_tensor_to_gif(got.unsqueeze(0), str(tmp_path / "out" / "a.gif"), fps=5)
    video_io.save_tensor_to_images(got.unsqueeze(0), str(tmp_path / "jpg"))
    from PIL import Image
    gif = Image.open(str(tmp_path / "out" / "a.gif"))
    assert gif.n_frames == 5 and gif.size == (32, 40) and len(list((tmp_path / "jpg").iterdir())) == 5
