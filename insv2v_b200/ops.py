"""Tensor-level wrappers over the C ABI (include/ivv.h). PyTorch is plumbing only: it owns device memory and the
stream; every operation below is one call into libivv_b200.so. No CPU or eager-PyTorch fallback exists."""
import ctypes

import torch

from . import lib as _lib

F16 = torch.float16


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _chk16(t, name):
    if t is None:
        return
    if not t.is_cuda or t.dtype != F16 or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA fp16 tensor, got {t.dtype} {t.device} "
                         f"contiguous={t.is_contiguous()}")


def _chk16v(t, name):
    """Like _chk16 but also accepts a column-slice view of a row-major buffer (unit inner stride, 16-byte aligned)."""
    if t is None:
        return
    if not t.is_cuda or t.dtype != F16 or t.stride(-1) != 1 or t.data_ptr() % 16 != 0:
        raise ValueError(f"{name}: expected a CUDA fp16 tensor with unit inner stride and 16-byte alignment, got "
                         f"{t.dtype} {t.device} strides={t.stride()}")
    if t.dim() != 2 and not t.is_contiguous():
        raise ValueError(f"{name}: only 2-D tensors may be strided views")


def _count():
    _lib.LAUNCH_COUNT += 1


# ---- optional per-call device timing (tools/profile_ops.py); off in production -------------------------------
class Prof:
    enabled = False
    records = []  # (key, flops, bytes, start_event, end_event)

    @classmethod
    def begin(cls):
        if not cls.enabled:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    @classmethod
    def end(cls, e0, key, flops=0.0, nbytes=0.0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        cls.records.append((key, flops, nbytes, e0, e1))

    @classmethod
    def report(cls):
        torch.cuda.synchronize()
        agg = {}
        for key, fl, nb, e0, e1 in cls.records:
            a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += fl
            a[3] += nb
        cls.records = []
        return agg


# default allocator; UNet/VAE runners swap in an arena so that CUDA-graph replays see stable addresses
class Alloc:
    fn = staticmethod(lambda shape, dtype, device: torch.empty(shape, dtype=dtype, device=device))


def empty(shape, dtype, device):
    return Alloc.fn(tuple(shape), dtype, device)


# ----------------------------------------------------------------------------------------------------------------
# weight packing (done once at load time)
# ----------------------------------------------------------------------------------------------------------------
def _pad_last(t, mult=8):
    k = t.shape[-1]
    kp = (k + mult - 1) // mult * mult
    if kp == k:
        return t.contiguous()
    out = t.new_zeros(*t.shape[:-1], kp)
    out[..., :k] = t
    return out


def pack_linear(w):
    """nn.Linear / 1x1-conv weight [out, in(,1,1)] -> fp16 [1, out, in_pad8] (K-major rows, as tcgen05 wants B)."""
    w = w.reshape(w.shape[0], -1)
    return _pad_last(w.to(F16)).unsqueeze(0).contiguous()


def pack_conv3x3(w):
    """Conv2d weight [co, ci, 3, 3] -> fp16 [9, co, ci_pad8] (tap-major: tap = ky*3+kx)."""
    co, ci = w.shape[0], w.shape[1]
    return _pad_last(w.to(F16).permute(2, 3, 0, 1).reshape(9, co, ci)).contiguous()


def pack_conv3x3_im2col(w):
    """Conv2d weight [co, ci, 3, 3] -> fp16 [1, co, 9*ci] matching ivv_im2col_s2's column order (tap*ci + c)."""
    co, ci = w.shape[0], w.shape[1]
    return w.to(F16).permute(0, 2, 3, 1).reshape(1, co, 9 * ci).contiguous()


def pack_geglu(w, b):
    """GEGLU proj weight [2*inner, c] (rows: hidden | gate) -> 256-row tiles [128 hidden | 128 gate] so that one
    tcgen05 accumulator tile holds both halves of the same 128 output columns; bias permuted identically."""
    inner = w.shape[0] // 2
    if inner % 128 != 0:
        raise ValueError(f"GEGLU inner dim {inner} must be a multiple of 128")
    idx = torch.arange(inner, device=w.device).reshape(inner // 128, 128)
    perm = torch.cat([idx, idx + inner], dim=1).reshape(-1)
    return pack_linear(w[perm]), b[perm].to(F16).contiguous()


# ----------------------------------------------------------------------------------------------------------------
# GEMM / conv
# ----------------------------------------------------------------------------------------------------------------
SPLITK = True  # module switch (tests / tuning)


def _splitk_policy(rows, k_total, n_out):
    """Number of K splits: only when the output has too few 128x128 tiles to fill the SMs and K is long."""
    if rows > 2304 or k_total < 4096 or n_out % 8 != 0:
        return 1
    tiles = ((rows + 127) // 128) * ((n_out + 127) // 128)
    return max(1, min(296 // tiles, k_total // 1024, 8))


def ln_fold_ok(rows, k, n_out):
    """True if ivv_gemm serves the linear layer [rows, k] -> [rows, n_out] with the kernel that can emit row statistics
    (row_stats_out=) and apply a folded LayerNorm (ln=)."""
    return bool(_lib.load().ivv_gemm_ln_fold_ok(rows, k, n_out))


def pack_ln_linear(gamma, beta, w, b=None, pe=None):
    """LayerNorm folded into the Linear that follows it: LN(x) W^T + b = rstd (x W'^T - mean wsum) + b' with
    W' = W diag(gamma), wsum[n] = sum_k W'[n][k] (of the fp16-rounded W', which is what the tensor core multiplies),
    b' = b + W beta. pe [L, k] (temporal positional encoding added AFTER the norm, motion_module.py:236-242,287) becomes
    a per-frame bias table pe W^T [L, n]. Returns (packed W' fp16 [1, n, k], b' fp16 [n], wsum fp16 [n], table or None)."""
    wf, g32, b32 = w.detach().float(), gamma.detach().float(), beta.detach().float()
    wp = (wf * g32[None, :]).to(F16)
    wsum = wp.float().sum(dim=1).to(F16).contiguous()
    bias = wf @ b32
    if b is not None:
        bias = bias + b.detach().float()
    table = (pe.detach().float() @ wf.t()).to(F16).contiguous() if pe is not None else None
    return pack_linear(wp), bias.to(F16).contiguous(), wsum, table


def gemm(a, wgt, *, n_img, h, w, c, n_out=None, taps=1, a_ld=None, bias=None, rowbias=None, rowbias_group=0,
         residual=None, geglu=False, out=None, out_f32=False, _splits=1, tap_hw=None, relu=False, rowbias_mod=0,
         row_stats_out=None, ln=None):
    """D = conv/linear(A, W) with fused epilogue; A rows are pixels [n_img*h*w, a_ld], W packed by pack_*().
    a / out / residual may be column-slice views of wider row-major buffers (their row stride is passed as the
    leading dimension). tap_hw = (kh, kw): odd stride-1 'same' window other than 1x1 / 3x3 (taps = kh*kw).
    row_stats_out: fp32 [rows, n_out // 40, 2] receiving per-row partial (sum, sum of squares) of the results;
    ln = (stats [rows, parts, 2] fp32, wsum fp16 [n_out], eps): A holds UN-normalised rows whose LayerNorm is applied
    in the epilogue (weights from pack_ln_linear). Both need ln_fold_ok(rows, c, n_out)."""
    _chk16v(a, "a"), _chk16(wgt, "wgt"), _chk16(bias, "bias"), _chk16(rowbias, "rowbias"), _chk16v(residual, "residual")
    if wgt.dim() != 3 or wgt.shape[0] != taps:
        raise ValueError(f"weight must be [taps={taps}, n_out, k_pad], got {tuple(wgt.shape)}")
    n_out = wgt.shape[1] if n_out is None else n_out
    a_ld = a.stride(-2) if a_ld is None else a_ld
    rows = n_img * h * w
    cols = n_out // 2 if geglu else n_out
    # split-K for the few-row / long-K convolutions of the 4x6 and 8x12 levels: too few output tiles for 148 SMs
    splits = _splitk_policy(rows, c * taps, n_out) if (SPLITK and not geglu and not out_f32 and not relu and
                                                        row_stats_out is None and ln is None and
                                                        tap_hw is None and a.is_contiguous() and
                                                        (residual is None or residual.is_contiguous()) and
                                                        (out is None or (out.dtype == F16 and out.is_contiguous()))
                                                        ) else 1
    if splits > 1:
        partial = empty((splits, rows, n_out), torch.float32, a.device)
        gemm(a, wgt, n_img=n_img, h=h, w=w, c=c, n_out=n_out, taps=taps, a_ld=a_ld, out=partial, _splits=splits)
        if out is None:
            out = empty((rows, n_out), F16, a.device)
        _lib.check(_lib.load().ivv_splitk_reduce(
            _p(partial), splits, rows, n_out, n_out, _p(bias), _p(rowbias), rowbias_group if rowbias is not None else 1,
            rowbias.shape[-1] if rowbias is not None else 0, _p(residual),
            residual.shape[-1] if residual is not None else 0, _p(out), out.shape[-1], _s()), "ivv_splitk_reduce")
        _count()
        return out
    if out is None:
        out = empty((rows, cols), torch.float32 if out_f32 else F16, a.device)
    args = _lib.GemmArgs()
    args.a, args.n_img, args.h, args.w, args.c, args.a_ld = a.data_ptr(), n_img, h, w, c, a_ld
    args.wgt, args.n_out, args.w_ld = wgt.data_ptr(), n_out, wgt.shape[2]
    args.taps, args.geglu = taps, int(geglu)
    if out.stride(-1) != 1:
        raise ValueError("gemm: out must have unit inner stride")
    args.d, args.d_ld, args.out_f32 = out.data_ptr(), out.stride(-2), int(out.dtype == torch.float32)
    args.splits = _splits
    args.relu = int(relu)
    if tap_hw is not None:
        args.tap_h, args.tap_w = tap_hw
    args.bias = bias.data_ptr() if bias is not None else None
    if rowbias is not None:
        args.rowbias, args.rowbias_group, args.rowbias_ld = rowbias.data_ptr(), rowbias_group, rowbias.shape[-1]
        args.rowbias_mod = rowbias_mod
    if row_stats_out is not None:
        if (row_stats_out.dtype != torch.float32 or not row_stats_out.is_contiguous()
                or row_stats_out.numel() != rows * (n_out // 40) * 2):
            raise ValueError(f"row_stats_out must be contiguous fp32 [rows={rows}, {n_out // 40}, 2]")
        args.row_stats_out = row_stats_out.data_ptr()
    if ln is not None:
        stats, wsum, eps = ln
        _chk16(wsum, "ln wsum")
        if stats.dtype != torch.float32 or not stats.is_contiguous() or stats.dim() != 3 or stats.shape[0] != rows \
                or stats.shape[2] != 2 or wsum.numel() != n_out:
            raise ValueError(f"ln: stats must be fp32 [rows={rows}, parts, 2] and wsum fp16 [{n_out}]")
        args.ln_stats, args.ln_wsum, args.ln_parts, args.ln_eps = stats.data_ptr(), wsum.data_ptr(), stats.shape[1], eps
    if residual is not None:
        args.residual, args.res_ld = residual.data_ptr(), residual.stride(-2)
    e0 = Prof.begin()
    _lib.check(_lib.load().ivv_gemm(ctypes.byref(args), _s()), "ivv_gemm")
    if e0 is not None:
        Prof.end(e0, ("gemm", rows, c * taps, n_out, "geglu" if geglu else "", "res" if residual is not None else ""),
                 2.0 * rows * c * taps * n_out,
                 2.0 * (rows * c + rows * cols * (2 if residual is not None else 1) + n_out * c * taps))
    _count()
    return out


def linear(x, wgt, bias=None, residual=None, geglu=False, out=None, **kw):
    """x [rows, k] -> [rows, n_out] (nn.Linear semantics); kw: row_stats_out / ln / rowbias* of gemm()."""
    rows, k = x.shape
    return gemm(x, wgt, n_img=1, h=1, w=rows, c=k, bias=bias, residual=residual, geglu=geglu, out=out, **kw)


def row_stats(rows, n_out, device):
    """Buffer for gemm(row_stats_out=): per row, one (sum, sum of squares) pair per 40 output columns."""
    return empty((rows, n_out // 40, 2), torch.float32, device)


def conv3x3(x, wgt, n_img, h, w, bias=None, rowbias=None, rowbias_group=0, residual=None, out=None, out_f32=False):
    """x frames [n_img*h*w, c] -> [n_img*h*w, co]; stride 1, zero pad 1 (InflatedConv3d, resnet.py:10-18)."""
    return gemm(x, wgt, n_img=n_img, h=h, w=w, c=x.shape[-1], taps=9, bias=bias, rowbias=rowbias,
                rowbias_group=rowbias_group, residual=residual, out=out, out_f32=out_f32)


def conv3x3_s2(x, wgt_im2col, n_img, h, w, bias=None, pad=1):
    """stride-2 3x3 conv: im2col gather then one tcgen05 GEMM. pad=1: Downsample3D (resnet.py:99-107);
    pad=0: VAE encoder Downsample (zero pad right/bottom only, vqvae/model.py:67-71)."""
    _chk16(x, "x")
    c = x.shape[-1]
    ho, wo = (h - 2 + pad) // 2 + 1, (w - 2 + pad) // 2 + 1
    cols = empty((n_img * ho * wo, 9 * c), F16, x.device)
    _lib.check(_lib.load().ivv_im2col_s2(_p(x), _p(cols), n_img, h, w, c, ho, wo, pad, _s()), "ivv_im2col_s2")
    _count()
    return gemm(cols, wgt_im2col, n_img=1, h=1, w=n_img * ho * wo, c=9 * c, bias=bias), ho, wo


# ----------------------------------------------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------------------------------------------
_ws_cache = {}


class WS:
    """Norm-statistics workspace policy. Eager calls share one zero-initialised buffer per (device, stream). A CUDA
    graph must own its workspace (two graphs replayed on different streams would otherwise race on the ticket
    counters): graph owners measure `high_water` over an eager warm-up pass, allocate their own zero-filled buffer
    OUTSIDE capture and install it as `override` while capturing."""
    override = None
    high_water = 0


def _gn_ws(nbytes, device):
    WS.high_water = max(WS.high_water, nbytes)
    if WS.override is not None:
        if WS.override.numel() < nbytes or WS.override.device != device:
            raise RuntimeError(f"graph-owned norm workspace too small ({WS.override.numel()} < {nbytes} bytes)")
        return WS.override
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        # zero-filled once: the kernels return the ticket counters at its head to zero after every call (ivv.h K6/K7)
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def groupnorm(x, gamma, beta, n_img, hw, groups, frames_per_group, eps, silu, out=None):
    """x frames [n_img*hw, c]; statistics span frames_per_group consecutive frames (5-D GroupNorm of
    ResnetBlock3D, resnet.py:177) or one frame (frames_per_group=1: attention.py:101)."""
    _chk16(x, "x"), _chk16(gamma, "gamma"), _chk16(beta, "beta")
    c = x.shape[-1]
    if out is None:
        out = empty(x.shape, F16, x.device)
    L = _lib.load()
    need = L.ivv_groupnorm_ws_bytes(n_img, groups, frames_per_group)
    ws = _gn_ws(need, x.device)
    e0 = Prof.begin()
    _lib.check(L.ivv_groupnorm(_p(x), _p(out), _p(gamma), _p(beta), n_img, hw, c, groups, frames_per_group,
                               float(eps), int(silu), _p(ws), ws.numel(), _s()), "ivv_groupnorm")
    Prof.end(e0, ("groupnorm", n_img * hw, c, frames_per_group), 0.0, 6.0 * n_img * hw * c)
    _lib.LAUNCH_COUNT += 1 if L.ivv_groupnorm_is_fused(n_img, hw, c, groups, frames_per_group) else 2
    return out


def groupnorm2(x1, x2, gamma, beta, n_img, hw, groups, frames_per_group, eps, silu, out=None):
    """GroupNorm(+SiLU) of the channel concatenation [x1 | x2] read in place (no torch.cat / concat kernel)."""
    _chk16(x1, "x1"), _chk16(x2, "x2"), _chk16(gamma, "gamma"), _chk16(beta, "beta")
    c1, c2 = x1.shape[-1], x2.shape[-1]
    if x1.shape[0] != x2.shape[0] or gamma.numel() != c1 + c2:
        raise ValueError(f"groupnorm2: x1 {tuple(x1.shape)}, x2 {tuple(x2.shape)}, gamma {gamma.numel()}")
    if out is None:
        out = empty((x1.shape[0], c1 + c2), F16, x1.device)
    L = _lib.load()
    ws = _gn_ws(L.ivv_groupnorm_ws_bytes(n_img, groups, frames_per_group), x1.device)
    e0 = Prof.begin()
    _lib.check(L.ivv_groupnorm2(_p(x1), c1, _p(x2), c2, _p(out), _p(gamma), _p(beta), n_img, hw, groups,
                                frames_per_group, float(eps), int(silu), _p(ws), ws.numel(), _s()), "ivv_groupnorm2")
    Prof.end(e0, ("groupnorm", n_img * hw, c1 + c2, frames_per_group), 0.0, 6.0 * n_img * hw * (c1 + c2))
    _lib.LAUNCH_COUNT += 1 if L.ivv_groupnorm_is_fused(n_img, hw, c1 + c2, groups, frames_per_group) else 2
    return out


def layernorm(x, gamma, beta, eps=1e-5, pe=None, rows_per_frame=0, frames=0, pe_start=0, out=None):
    _chk16(x, "x"), _chk16(gamma, "gamma"), _chk16(beta, "beta")
    rows, c = x.shape
    if out is None:
        out = empty(x.shape, F16, x.device)
    if pe is not None and (pe.dtype != torch.float32 or not pe.is_contiguous()):
        raise ValueError("pe must be contiguous fp32")
    e0 = Prof.begin()
    _lib.check(_lib.load().ivv_layernorm(_p(x), _p(out), _p(gamma), _p(beta), rows, c, float(eps), _p(pe),
                                         rows_per_frame, frames, pe_start, _s()), "ivv_layernorm")
    Prof.end(e0, ("layernorm", rows, c), 0.0, 4.0 * rows * c)
    _count()
    return out


# ----------------------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------------------
def attention(q, k, v, *, n_batch, s_q, s_kv, heads, d, q_ld, kv_ld, kv_div=1, scale=None, out=None):
    """softmax(q k^T * scale) v per (batch, head). q/k/v may be column slices of wider row-major buffers
    (pass the buffer's row stride as q_ld / kv_ld)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.dtype != F16:
            raise ValueError(f"{n}: expected CUDA fp16")
    scale = d ** -0.5 if scale is None else scale
    if out is None:
        out = empty((n_batch * s_q, heads * d), F16, q.device)
    e0 = Prof.begin()
    _lib.check(_lib.load().ivv_attention(_p(q), q_ld, _p(k), _p(v), kv_ld, _p(out), out.shape[-1], n_batch, s_q, s_kv,
                                         kv_div, heads, d, float(scale), _s()), "ivv_attention")
    Prof.end(e0, ("attention", n_batch, s_q, s_kv, heads, d), 4.0 * n_batch * heads * s_q * s_kv * d,
             2.0 * n_batch * heads * d * (2 * s_q + 2 * s_kv / kv_div))
    _count()
    return out


def temporal_attention(qkv, clips, frames, hw, c, heads, scale=None, out=None):
    _chk16(qkv, "qkv")
    scale = (c // heads) ** -0.5 if scale is None else scale
    if out is None:
        out = empty((clips * frames * hw, c), F16, qkv.device)
    e0 = Prof.begin()
    _lib.check(_lib.load().ivv_temporal_attention(_p(qkv), _p(out), clips, frames, hw, c, heads, float(scale), _s()),
               "ivv_temporal_attention")
    Prof.end(e0, ("temporal_attention", clips * frames * hw, c), 4.0 * clips * hw * frames * frames * c,
             8.0 * clips * frames * hw * c)
    _count()
    return out


def softmax_rows(x, scale, out=None):
    if not x.is_cuda or x.dtype not in (F16, torch.float32) or not x.is_contiguous():
        raise ValueError("softmax_rows: expected contiguous CUDA fp16/fp32")
    rows, cols = x.shape
    if out is None:
        out = empty(x.shape, F16, x.device)
    _lib.check(_lib.load().ivv_softmax_rows(_p(x), int(x.dtype == torch.float32), _p(out), rows, cols, float(scale),
                                            _s()), "ivv_softmax_rows")
    _count()
    return out


def channelnorm(x, n_img, hw, imgs_per_group, gamma=None, beta=None, eps=1e-5, relu=False, residual=None, out=None):
    """InstanceNorm2d (imgs_per_group=1) / batch-statistics BatchNorm2d (imgs_per_group=n_img) over frames
    [n_img*hw, c], + ReLU, + relu(residual + y) (torchvision raft.py ResidualBlock)."""
    _chk16(x, "x"), _chk16(gamma, "gamma"), _chk16(beta, "beta"), _chk16(residual, "residual")
    c = x.shape[-1]
    if out is None:
        out = empty(x.shape, F16, x.device)
    L = _lib.load()
    ws = _gn_ws(L.ivv_channelnorm_ws_bytes(n_img, c, imgs_per_group), x.device)
    _lib.check(L.ivv_channelnorm(_p(x), _p(out), _p(gamma), _p(beta), n_img, hw, c, imgs_per_group, float(eps),
                                 int(relu), _p(residual), _p(ws), ws.numel(), _s()), "ivv_channelnorm")
    _lib.LAUNCH_COUNT += 2
    return out


def im2col(x, n_img, h, w, kh, kw, stride, pad_h, pad_w):
    """frames [n_img*h*w, c] -> [n_img*ho*wo, kh*kw*c] (column = tap*c + ci), zero padded."""
    _chk16(x, "x")
    c = x.shape[-1]
    ho, wo = (h + 2 * pad_h - kh) // stride + 1, (w + 2 * pad_w - kw) // stride + 1
    out = empty((n_img * ho * wo, kh * kw * c), F16, x.device)
    _lib.check(_lib.load().ivv_im2col(_p(x), _p(out), n_img, h, w, c, kh, kw, stride, pad_h, pad_w, ho, wo, _s()),
               "ivv_im2col")
    _count()
    return out, ho, wo


def pack_conv_im2col(w, c_pad=None):
    """Conv2d weight [co, ci, kh, kw] -> fp16 [1, co, kh*kw*c_pad] in ivv_im2col's column order (tap*c_pad + ci)."""
    co, ci, kh, kw = w.shape
    c_pad = ci if c_pad is None else c_pad
    t = w.new_zeros(co, kh, kw, c_pad)
    t[..., :ci] = w.permute(0, 2, 3, 1)
    return t.to(F16).reshape(1, co, kh * kw * c_pad).contiguous()


def pack_conv_taps(w, c_pad=None, co_pad=None):
    """Conv2d weight [co, ci, kh, kw] -> fp16 [kh*kw, co_pad, c_pad8] (tap = ky*kw + kx) for the implicit-tap GEMM."""
    co, ci, kh, kw = w.shape
    c_pad = (ci + 7) // 8 * 8 if c_pad is None else c_pad
    co_pad = co if co_pad is None else co_pad
    t = w.new_zeros(kh * kw, co_pad, c_pad)
    t[:, :co, :ci] = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    return t.to(F16).contiguous()


# ----------------------------------------------------------------------------------------------------------------
# data movement / elementwise
# ----------------------------------------------------------------------------------------------------------------
def upsample_nearest(x, n_img, h, w, ho=None, wo=None):
    _chk16(x, "x")
    c = x.shape[-1]
    ho, wo = (2 * h if ho is None else ho), (2 * w if wo is None else wo)
    out = empty((n_img * ho * wo, c), F16, x.device)
    _lib.check(_lib.load().ivv_upsample_nearest(_p(x), _p(out), n_img, h, w, c, ho, wo, _s()), "ivv_upsample_nearest")
    _count()
    return out, ho, wo


def concat_channels(a, b):
    _chk16(a, "a"), _chk16(b, "b")
    rows = a.shape[0]
    out = empty((rows, a.shape[1] + b.shape[1]), F16, a.device)
    _lib.check(_lib.load().ivv_concat_channels(_p(a), a.shape[1], _p(b), b.shape[1], _p(out), rows, _s()),
               "ivv_concat_channels")
    _count()
    return out


def ncfhw_to_frames(x, c_pad):
    """[b, c, f, h, w] fp32/fp16 -> frames [b*f*h*w, c_pad] fp16 (the reference's `(b f)` fold made physical once)."""
    if x.dtype not in (torch.float32, F16) or not x.is_cuda:
        raise ValueError("ncfhw_to_frames: expected CUDA fp32/fp16")
    x = x.contiguous()
    b, c, f, h, w = x.shape
    out = empty((b * f * h * w, c_pad), F16, x.device)
    _lib.check(_lib.load().ivv_ncfhw_to_frames(_p(x), int(x.dtype == torch.float32), _p(out), b, c, f, h * w, c_pad,
                                               _s()), "ivv_ncfhw_to_frames")
    _count()
    return out


def frames_to_ncfhw(x, b, c, f, h, w, out_dtype=torch.float32):
    if x.dtype not in (torch.float32, F16) or not x.is_contiguous():
        raise ValueError("frames_to_ncfhw: expected contiguous fp32/fp16")
    out = torch.empty((b, c, f, h, w), dtype=out_dtype, device=x.device)
    _lib.check(_lib.load().ivv_frames_to_ncfhw(_p(x), int(x.dtype == torch.float32), x.shape[-1], _p(out),
                                               int(out_dtype == torch.float32), b, c, f, h * w, _s()),
               "ivv_frames_to_ncfhw")
    _count()
    return out


def timestep_embedding(t, dim, flip_sin_to_cos=True, freq_shift=0.0):
    t = t.to(torch.float32).contiguous()
    out = empty((t.shape[0], dim), F16, t.device)
    _lib.check(_lib.load().ivv_timestep_embedding(_p(t), _p(out), t.shape[0], dim, int(flip_sin_to_cos),
                                                  float(freq_shift), _s()), "ivv_timestep_embedding")
    _count()
    return out


def silu(x):
    _chk16(x, "x")
    out = empty(x.shape, F16, x.device)
    _lib.check(_lib.load().ivv_silu(_p(x), _p(out), x.numel(), _s()), "ivv_silu")
    _count()
    return out


def scale(x, a, b=0.0):
    _chk16(x, "x")
    out = empty(x.shape, F16, x.device)
    _lib.check(_lib.load().ivv_scale(_p(x), _p(out), x.numel(), float(a), float(b), _s()), "ivv_scale")
    _count()
    return out


# ----------------------------------------------------------------------------------------------------------------
# flow warp / sampler step
# ----------------------------------------------------------------------------------------------------------------
def _chk32(t, name):
    if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA fp32 tensor")


def warp_image_f32(image, flow):
    _chk32(image, "image"), _chk32(flow, "flow")
    n, c, h, w = image.shape
    out = torch.empty_like(image)
    _lib.check(_lib.load().ivv_warp_image(_p(image), _p(flow), _p(out), n, c, h, w, _s()), "ivv_warp_image")
    _count()
    return out


def resize_flow_f32(flow, ho, wo):
    _chk32(flow, "flow")
    n, _, h, w = flow.shape
    out = torch.empty((n, 2, ho, wo), dtype=torch.float32, device=flow.device)
    _lib.check(_lib.load().ivv_resize_flow(_p(flow), _p(out), n, h, w, ho, wo, _s()), "ivv_resize_flow")
    _count()
    return out


def flow_noise_correction_(eps, delta_ref, flow_lat):
    """eps [Q, C, h, w] += masked mean over R reference frames of warp(delta_ref[r], flow_lat[q, r]) — in place."""
    _chk32(eps, "eps"), _chk32(delta_ref, "delta_ref"), _chk32(flow_lat, "flow_lat")
    q, c, h, w = eps.shape
    r = delta_ref.shape[0]
    if tuple(delta_ref.shape) != (r, c, h, w) or tuple(flow_lat.shape) != (q, r, 2, h, w):
        raise ValueError(f"flow_noise_correction_: eps {tuple(eps.shape)} needs delta_ref [R,{c},{h},{w}] and flows "
                         f"[{q},R,2,{h},{w}]; got {tuple(delta_ref.shape)} and {tuple(flow_lat.shape)}")
    _lib.check(_lib.load().ivv_flow_noise_correction(_p(delta_ref), _p(flow_lat), _p(eps), q, r, c, h, w, _s()),
               "ivv_flow_noise_correction")
    _count()
    return eps


def cfg_ddim_step_(eps3, latent, text_cfg, img_cfg, alpha_t, alpha_prev, eps_out=None):
    _chk32(eps3, "eps3"), _chk32(latent, "latent")
    n = latent.numel()
    _lib.check(_lib.load().ivv_cfg_ddim_step(_p(eps3), _p(latent), _p(eps_out), n, float(text_cfg), float(img_cfg),
                                             float(alpha_t), float(alpha_prev), _s()), "ivv_cfg_ddim_step")
    _count()
    return latent


# ----------------------------------------------------------------------------------------------------------------
# fused sampling step (csrc/sampler.cu): begin -> UNet -> combine -> update, all parameters read from device memory
# ----------------------------------------------------------------------------------------------------------------
SAMPLER_ROW = 16  # IVV_SAMPLER_ROW


def sampler_partials(frames, hw):
    return int(_lib.load().ivv_sampler_partials(frames, hw))


def sampler_begin(table, state, lat2, cond, x_frames, t_out, frames, c, hw, c_pad):
    _chk32(table, "table"), _chk32(lat2, "lat2"), _chk32(cond, "cond"), _chk32(t_out, "t_out"), _chk16(x_frames, "x")
    if x_frames.shape != (3 * frames * hw, c_pad) or lat2.numel() != 2 * frames * c * hw or cond.numel() != frames * c * hw:
        raise ValueError("sampler_begin: buffer shapes do not match (frames, c, hw, c_pad)")
    _lib.check(_lib.load().ivv_sampler_begin(_p(table), _p(state), _p(lat2), _p(cond), _p(x_frames), _p(t_out), frames,
                                             c, hw, c_pad, _s()), "ivv_sampler_begin")
    _count()


def sampler_combine(table, state, eps3, eps_cfg, partials, frames, c, hw):
    _chk32(eps3, "eps3"), _chk32(eps_cfg, "eps_cfg")
    if eps3.shape[0] != 3 * frames * hw or eps3.shape[1] < c or eps_cfg.numel() != frames * c * hw:
        raise ValueError("sampler_combine: buffer shapes do not match (frames, c, hw)")
    if partials.dtype != torch.float64 or partials.numel() < 4 * sampler_partials(frames, hw):
        raise ValueError("sampler_combine: partials must be float64 [ivv_sampler_partials, 4]")
    _lib.check(_lib.load().ivv_sampler_combine(_p(table), _p(state), _p(eps3), eps3.stride(0), _p(eps_cfg), _p(partials),
                                               frames, c, hw, _s()), "ivv_sampler_combine")
    _count()


def sampler_update(table, state, lat2, eps_cfg, partials, mode, latent_ref, flows_lat, noise, hist_lat, hist_pred,
                   frames, c, r, q, h, w):
    n = frames * c * h * w
    for t, name, numel in ((latent_ref, "latent_ref", r * c * h * w), (flows_lat, "flows_lat", q * r * 2 * h * w)):
        if t is not None:
            _chk32(t, name)
            if t.numel() != numel:
                raise ValueError(f"sampler_update: {name} has {t.numel()} elements, expected {numel}")
    for t, name in ((noise, "noise"), (hist_lat, "hist_lat"), (hist_pred, "hist_pred")):
        if t is not None:
            _chk32(t, name)
            if t.numel() % n != 0:
                raise ValueError(f"sampler_update: {name} must be [rows, {n}]")
    _lib.check(_lib.load().ivv_sampler_update(_p(table), _p(state), _p(lat2), _p(eps_cfg), _p(partials), mode,
                                              _p(latent_ref), _p(flows_lat), _p(noise), _p(hist_lat), _p(hist_pred),
                                              frames, c, r, q, h, w, _s()), "ivv_sampler_update")
    _count()
