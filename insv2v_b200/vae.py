"""Drop-in for the reference's `modules.kl_autoencoder.autoencoder.AutoencoderKL` (autoencoder.py:50-100) over
`modules/vqvae/model.py` Encoder/Decoder: same constructor (`ddconfig`, `lossconfig`, `embed_dim`), same
`decode(z)` / `encode(x)` contract, same state-dict keys (`encoder.*`, `decoder.*`, `quant_conv.*`,
`post_quant_conv.*`; tests/golden/schema_vae_*.json). The module tree only holds parameters; the arithmetic is a
sequence of C-ABI calls (tcgen05 implicit-GEMM convolutions, GroupNorm+swish kernels) over channels-last fp16 images,
with all frames of a clip batched into one launch per layer (the reference decodes frame by frame,
instruct_p2p_video.py:72-76; per-image GroupNorm makes batching exact)."""
import torch
from torch import nn

from . import ops
from .unet import _Holder, _h, weights_stamp

F16 = torch.float16


def _norm(c):
    return nn.GroupNorm(32, c, eps=1e-6)  # Normalize, vqvae/model.py:31-32


class _ResBlock(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = _norm(cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = _norm(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)


class _AttnBlock(_Holder):
    def __init__(self, c):
        super().__init__()
        self.norm = _norm(c)
        self.q = nn.Conv2d(c, c, 1)
        self.k = nn.Conv2d(c, c, 1)
        self.v = nn.Conv2d(c, c, 1)
        self.proj_out = nn.Conv2d(c, c, 1)


class _ConvOnly(_Holder):
    def __init__(self, c, stride, padding):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=padding)


class _Level(_Holder):
    pass


class _Mid(_Holder):
    def __init__(self, c):
        super().__init__()
        self.block_1 = _ResBlock(c, c)
        self.attn_1 = _AttnBlock(c)
        self.block_2 = _ResBlock(c, c)


class _Decoder(_Holder):
    """Parameter tree of Decoder (vqvae/model.py:305-376); `up[0]` is the highest resolution, as in the reference."""

    def __init__(self, *, ch, out_ch, ch_mult, num_res_blocks, attn_resolutions, in_channels, resolution, z_channels,
                 **ignored):
        super().__init__()
        if len(attn_resolutions) != 0:
            raise NotImplementedError("attn_resolutions must be empty (InsV2V config)")
        nres = len(ch_mult)
        block_in = ch * ch_mult[nres - 1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = _Mid(block_in)
        ups = []
        for lvl in reversed(range(nres)):
            block_out = ch * ch_mult[lvl]
            up = _Level()
            blocks = []
            for _ in range(num_res_blocks + 1):
                blocks.append(_ResBlock(block_in, block_out))
                block_in = block_out
            up.block = nn.ModuleList(blocks)
            up.attn = nn.ModuleList()
            if lvl != 0:
                up.upsample = _ConvOnly(block_in, 1, 1)
            ups.insert(0, up)
        self.up = nn.ModuleList(ups)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)
        self.num_resolutions, self.num_res_blocks = nres, num_res_blocks


class _Encoder(_Holder):
    """Parameter tree of Encoder (vqvae/model.py:211-273)."""

    def __init__(self, *, ch, out_ch, ch_mult, num_res_blocks, attn_resolutions, in_channels, resolution, z_channels,
                 double_z=True, **ignored):
        super().__init__()
        nres = len(ch_mult)
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_ch_mult = (1,) + tuple(ch_mult)
        downs = []
        block_in = ch
        for lvl in range(nres):
            block_in = ch * in_ch_mult[lvl]
            block_out = ch * ch_mult[lvl]
            down = _Level()
            blocks = []
            for _ in range(num_res_blocks):
                blocks.append(_ResBlock(block_in, block_out))
                block_in = block_out
            down.block = nn.ModuleList(blocks)
            down.attn = nn.ModuleList()
            if lvl != nres - 1:
                down.downsample = _ConvOnly(block_in, 2, 0)
            downs.append(down)
        self.down = nn.ModuleList(downs)
        self.mid = _Mid(block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, padding=1)
        self.num_resolutions, self.num_res_blocks = nres, num_res_blocks


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None):
        super().__init__()
        dd = dict(ddconfig)
        dd["ch_mult"] = tuple(dd["ch_mult"])
        dd["attn_resolutions"] = tuple(dd.get("attn_resolutions", ()))
        assert dd["double_z"]
        self.image_key = image_key
        self.encoder = _Encoder(**dd)
        self.decoder = _Decoder(**dd)
        self.loss = nn.Identity()  # lossconfig: torch.nn.Identity (configs/instruct_v2v_inference.yaml:88-89)
        self.quant_conv = nn.Conv2d(2 * dd["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, dd["z_channels"], 1)
        self.embed_dim = embed_dim
        self._packed = None
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.invalidate())
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=list(ignore_keys))

    def init_from_ckpt(self, path, ignore_keys=()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        self.load_state_dict(sd, strict=False)

    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    def invalidate(self):
        """Drop the packed weights (automatic: load-state-dict post hook + weight stamp, see unet.weights_stamp)."""
        self._packed = None

    def packed(self, dev):
        stamp = weights_stamp(self)
        P = self._packed
        if P is None or P["device"] != dev or P["stamp"] != stamp:
            P = self._pack(dev)
            P["stamp"] = stamp
        return P

    # ---- packing ---------------------------------------------------------------------------------------------
    def _pack(self, dev):
        def conv3(c, pad_in=None, pad_out=None):
            w, b = c.weight.detach().to(dev), c.bias.detach().to(dev)
            if pad_in is not None and w.shape[1] < pad_in:
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_in - w.shape[1], 3, 3)], 1)
            return ops.pack_conv3x3(w), _h(b, dev)

        def lin(c, pad_out=None, pad_in=None):
            w, b = c.weight.detach().to(dev).reshape(c.weight.shape[0], -1), c.bias.detach().to(dev)
            if pad_in is not None and w.shape[1] < pad_in:
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_in - w.shape[1])], 1)
            if pad_out is not None and w.shape[0] < pad_out:
                w = torch.cat([w, w.new_zeros(pad_out - w.shape[0], w.shape[1])], 0)
                b = torch.cat([b, b.new_zeros(pad_out - b.shape[0])], 0)
            return ops.pack_linear(w), _h(b, dev)

        def gn(n):
            return _h(n.weight, dev), _h(n.bias, dev)

        def res(r):
            return dict(n1=gn(r.norm1), c1=conv3(r.conv1), n2=gn(r.norm2), c2=conv3(r.conv2),
                        sc=lin(r.nin_shortcut) if hasattr(r, "nin_shortcut") else None)

        def attn(a):
            return dict(norm=gn(a.norm), q=lin(a.q), k=lin(a.k), v=lin(a.v), out=lin(a.proj_out), c=a.q.weight.shape[0])

        d, e = self.decoder, self.encoder
        P = dict(device=dev)
        P["pq"] = lin(self.post_quant_conv, pad_out=8, pad_in=8)
        P["d_in"] = conv3(d.conv_in, pad_in=8)
        P["d_mid"] = (res(d.mid.block_1), attn(d.mid.attn_1), res(d.mid.block_2))
        P["d_up"] = [dict(blocks=[res(b) for b in up.block], up=conv3(up.upsample.conv) if hasattr(up, "upsample")
                          else None) for up in d.up]
        P["d_nout"] = gn(d.norm_out)
        P["d_out"] = conv3(d.conv_out)
        P["d_out_ch"] = d.conv_out.weight.shape[0]
        P["e_in"] = conv3(e.conv_in, pad_in=8)
        P["e_down"] = [dict(blocks=[res(b) for b in dn.block],
                            down=(ops.pack_conv3x3_im2col(dn.downsample.conv.weight.detach().to(dev)),
                                  _h(dn.downsample.conv.bias, dev)) if hasattr(dn, "downsample") else None)
                       for dn in e.down]
        P["e_mid"] = (res(e.mid.block_1), attn(e.mid.attn_1), res(e.mid.block_2))
        P["e_nout"] = gn(e.norm_out)
        P["e_out"] = conv3(e.conv_out)
        P["q"] = lin(self.quant_conv)
        self._packed = P
        return P

    # ---- building blocks on frames [n*h*w, c] ------------------------------------------------------------------
    @staticmethod
    def _res(x, p, n, h, w):
        t = ops.groupnorm(x, *p["n1"], n, h * w, 32, 1, 1e-6, True)
        t = ops.conv3x3(t, p["c1"][0], n, h, w, bias=p["c1"][1])
        t = ops.groupnorm(t, *p["n2"], n, h * w, 32, 1, 1e-6, True)
        r = x if p["sc"] is None else ops.linear(x, p["sc"][0], bias=p["sc"][1])
        return ops.conv3x3(t, p["c2"][0], n, h, w, bias=p["c2"][1], residual=r)

    @staticmethod
    def _attn(x, p, n, h, w):
        """AttnBlock (vqvae/model.py:173-197): one head of width c. Scores are materialised per image in fp32
        (0.5 % of the decode FLOPs); V is produced transposed (V^T = Wv X^T) so that P V is a K-major tcgen05 GEMM,
        and its bias is added after the product (softmax rows sum to 1)."""
        c, s = p["c"], h * w
        hn = ops.groupnorm(x, *p["norm"], n, s, 32, 1, 1e-6, False)
        q = ops.linear(hn, p["q"][0], bias=p["q"][1])
        k = ops.linear(hn, p["k"][0], bias=p["k"][1])
        o = ops.empty((n * s, c), F16, x.device)
        for i in range(n):
            hi, qi, ki = hn[i * s:(i + 1) * s], q[i * s:(i + 1) * s], k[i * s:(i + 1) * s]
            vt = ops.gemm(p["v"][0][0], hi.unsqueeze(0), n_img=1, h=1, w=c, c=c)          # [c, s] = Wv hn^T
            sc = ops.gemm(qi, ki.unsqueeze(0), n_img=1, h=1, w=s, c=c, out_f32=True)      # [s, s] fp32
            pr = ops.softmax_rows(sc, c ** -0.5)
            ops.gemm(pr, vt.unsqueeze(0), n_img=1, h=1, w=s, c=s, bias=p["v"][1], out=o[i * s:(i + 1) * s])
        return ops.linear(o, p["out"][0], bias=p["out"][1], residual=x)

    # ---- API ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z):
        """z [n, z_channels, h, w] -> image [n, out_ch, 8h, 8w] (dtype follows z)."""
        if not z.is_cuda:
            raise RuntimeError("insv2v_b200.AutoencoderKL runs only on CUDA (sm_100a); there is no CPU path")
        P = self.packed(z.device)
        n, zc, h, w = z.shape
        x = ops.ncfhw_to_frames(z.reshape(n, zc, 1, h, w), 8)
        x = ops.linear(x, P["pq"][0], bias=P["pq"][1])
        x = ops.conv3x3(x, P["d_in"][0], n, h, w, bias=P["d_in"][1])
        x = self._res(x, P["d_mid"][0], n, h, w)
        x = self._attn(x, P["d_mid"][1], n, h, w)
        x = self._res(x, P["d_mid"][2], n, h, w)
        for lvl in reversed(range(len(P["d_up"]))):
            for bp in P["d_up"][lvl]["blocks"]:
                x = self._res(x, bp, n, h, w)
            if P["d_up"][lvl]["up"] is not None:
                x, h, w = ops.upsample_nearest(x, n, h, w)
                x = ops.conv3x3(x, P["d_up"][lvl]["up"][0], n, h, w, bias=P["d_up"][lvl]["up"][1])
        x = ops.groupnorm(x, *P["d_nout"], n, h * w, 32, 1, 1e-6, True)
        oc = P["d_out_ch"]
        out = ops.empty((n * h * w, 8), torch.float32, z.device)
        ops.conv3x3(x, P["d_out"][0], n, h, w, bias=P["d_out"][1], out=out)
        img = ops.frames_to_ncfhw(out, n, oc, 1, h, w, torch.float32).reshape(n, oc, h, w)
        return img if z.dtype == torch.float32 else img.to(z.dtype)

    @torch.no_grad()
    def encode_moments(self, x):
        """Encoder + quant_conv: image [n, 3, H, W] -> moments [n, 2*embed_dim, H/8, W/8] (mean | logvar), fp32."""
        if not x.is_cuda:
            raise RuntimeError("insv2v_b200.AutoencoderKL runs only on CUDA (sm_100a); there is no CPU path")
        P = self.packed(x.device)
        n, c, h, w = x.shape
        t = ops.ncfhw_to_frames(x.reshape(n, c, 1, h, w), 8)
        t = ops.conv3x3(t, P["e_in"][0], n, h, w, bias=P["e_in"][1])
        for lvl in P["e_down"]:
            for bp in lvl["blocks"]:
                t = self._res(t, bp, n, h, w)
            if lvl["down"] is not None:
                t, h, w = ops.conv3x3_s2(t, lvl["down"][0], n, h, w, bias=lvl["down"][1], pad=0)
        t = self._res(t, P["e_mid"][0], n, h, w)
        t = self._attn(t, P["e_mid"][1], n, h, w)
        t = self._res(t, P["e_mid"][2], n, h, w)
        t = ops.groupnorm(t, *P["e_nout"], n, h * w, 32, 1, 1e-6, True)
        t = ops.conv3x3(t, P["e_out"][0], n, h, w, bias=P["e_out"][1])
        mc = P["q"][0].shape[1]
        m = ops.gemm(t, P["q"][0], n_img=1, h=1, w=n * h * w, c=t.shape[-1], bias=P["q"][1], out_f32=True)
        return ops.frames_to_ncfhw(m, n, mc, 1, h, w, torch.float32).reshape(n, mc, h, w)

    @torch.no_grad()
    def encode(self, x):
        """autoencoder.py:89-95: samples the diagonal Gaussian posterior, logvar clamped to [-30, 20]. As in the
        reference (DiagonalGaussianDistribution.sample, :22) the noise is drawn on the CPU default generator and moved
        to the device, so a seeded run reproduces the reference's latents."""
        mean, logvar = torch.chunk(self.encode_moments(x), 2, dim=1)
        std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
        z = mean + std * torch.randn(mean.shape).to(device=mean.device)
        return z if x.dtype == torch.float32 else z.to(x.dtype)

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("training forward (reconstruction + posterior) is outside the inference hot path")
