"""Drop-in for the reference's `misc_utils.flow_utils.RAFTFlow` (flow_utils.py:134-189): optical flow between two
frame batches with torchvision's `raft_large` architecture, on the sm_100a kernels of libivv_b200.so.

`RAFTFlow.model` is a parameter tree with torchvision's own key names (`feature_encoder.*`, `context_encoder.*`,
`update_block.*`, `mask_predictor.*` — `Raft_Large_Weights` state dicts load unchanged); the arithmetic is a sequence
of C-ABI calls over channels-last fp16 maps at 1/2, 1/4 and 1/8 resolution:

  * every convolution is a tcgen05 GEMM (`ivv_gemm`): implicit 3x3 / 1x5 / 5x1 taps, `ivv_im2col` + GEMM for the 7x7 and
    the stride-2 ones; ReLU, bias and the GRU's `W_q [r*h]` + `W_q x` split ride in the GEMM epilogue;
  * InstanceNorm / BatchNorm (+ReLU, + the residual join) are one two-pass `ivv_channelnorm`;
  * the all-pairs correlation is a GEMM per image pair with fp32 output, pooled into the 4-level pyramid; the 9x9x4
    look-up, the GRU gates, the coordinate update and the convex 8x up-sampling are small fused kernels (raft.cu).

Reference behaviour kept on purpose: `RAFTFlow` is never switched to eval mode by the reference
(pl_trainer/inference/inference.py:294), so BatchNorm normalises with batch statistics; `.eval()` gives the usual
running-statistics behaviour (folded into the convolutions). fp16 storage with fp32 accumulation; the hidden state,
the coordinates and the correlation pyramid stay fp32.
"""
import ctypes

import torch
from torch import nn

from . import lib as _lib
from . import ops

F16 = torch.float16

_CFG = dict(encoder_layers=(64, 64, 96, 128, 256), encoder_strides=(2, 1, 2, 2), corr_levels=4, corr_radius=4,
            corr_layers=(256, 192), flow_layers=(128, 64), motion_out=128, hidden=128, gru_kernels=((1, 5), (5, 1)),
            flow_head_hidden=256, mask_hidden=256, mask_multiplier=0.25)


class _Node(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the arithmetic runs in insv2v_b200.raft._Engine")


def _register(root, key, shape):
    """Create the nested holders for a dotted state-dict key and register the tensor (buffer for BN statistics)."""
    *path, leaf = key.split(".")
    mod = root
    for p in path:
        if not hasattr(mod, p):
            mod.add_module(p, _Node())
        mod = getattr(mod, p)
    if leaf in ("running_mean", "num_batches_tracked"):
        mod.register_buffer(leaf, torch.zeros(shape, dtype=torch.long if leaf == "num_batches_tracked" else None))
    elif leaf == "running_var":
        mod.register_buffer(leaf, torch.ones(shape))
    else:
        mod.register_parameter(leaf, nn.Parameter(torch.zeros(shape), requires_grad=False))


def _schema(cfg=_CFG):
    """State-dict layout of torchvision raft_large (models/optical_flow/raft.py:121-151, 176-198, 216-220, 277-281,
    317-321): Conv2dNormActivation = Sequential(conv `.0`, norm `.1`, relu)."""
    L, sch = cfg["encoder_layers"], {}

    def conv(name, co, ci, kh, kw):
        sch[name + ".weight"] = (co, ci, kh, kw)
        sch[name + ".bias"] = (co,)

    for enc, batch in (("feature_encoder", False), ("context_encoder", True)):
        def cna(name, co, ci, k):
            conv(f"{enc}.{name}.0", co, ci, k, k)
            if batch:
                for leaf, shp in (("weight", (co,)), ("bias", (co,)), ("running_mean", (co,)), ("running_var", (co,)),
                                  ("num_batches_tracked", ())):
                    sch[f"{enc}.{name}.1.{leaf}"] = shp
        cna("convnormrelu", L[0], 3, 7)
        cin = L[0]
        for li, (cout, stride) in enumerate(zip(L[1:4], cfg["encoder_strides"][1:]), start=1):
            for bi in range(2):
                cna(f"layer{li}.{bi}.convnormrelu1", cout, cin if bi == 0 else cout, 3)
                cna(f"layer{li}.{bi}.convnormrelu2", cout, cout, 3)
                if bi == 0 and stride != 1:
                    cna(f"layer{li}.{bi}.downsample", cout, cin, 1)
            cin = cout
        conv(f"{enc}.conv", L[4], L[3], 1, 1)
    ccorr = cfg["corr_levels"] * (2 * cfg["corr_radius"] + 1) ** 2
    me = "update_block.motion_encoder"
    conv(me + ".convcorr1.0", cfg["corr_layers"][0], ccorr, 1, 1)
    conv(me + ".convcorr2.0", cfg["corr_layers"][1], cfg["corr_layers"][0], 3, 3)
    conv(me + ".convflow1.0", cfg["flow_layers"][0], 2, 7, 7)
    conv(me + ".convflow2.0", cfg["flow_layers"][1], cfg["flow_layers"][0], 3, 3)
    conv(me + ".conv.0", cfg["motion_out"] - 2, cfg["corr_layers"][1] + cfg["flow_layers"][1], 3, 3)
    hid = cfg["hidden"]
    gin = hid + cfg["motion_out"] + (L[4] - hid)
    for gi, (kh, kw) in enumerate(cfg["gru_kernels"], start=1):
        for g in ("convz", "convr", "convq"):
            conv(f"update_block.recurrent_block.convgru{gi}.{g}", hid, gin, kh, kw)
    conv("update_block.flow_head.conv1", cfg["flow_head_hidden"], hid, 3, 3)
    conv("update_block.flow_head.conv2", 2, cfg["flow_head_hidden"], 3, 3)
    conv("mask_predictor.convrelu.0", cfg["mask_hidden"], hid, 3, 3)
    conv("mask_predictor.conv", 8 * 8 * 9, cfg["mask_hidden"], 1, 1)
    return sch


def _h(t, dev):
    return t.detach().to(device=dev, dtype=F16).contiguous()


class _Unit:
    """One Conv2dNormActivation: packed conv weight/bias, the norm kind and (train-mode BN) its affine parameters."""
    __slots__ = ("kind", "w", "b", "gamma", "beta", "norm", "co")


class _Engine:
    def __init__(self, model, device, training):
        self.device, self.training = device, training
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in model.state_dict().items()}
        cfg = self.cfg = _CFG
        self.enc = {}
        for enc, batch in (("feature_encoder", False), ("context_encoder", True)):
            E = dict(stem=self._unit(sd, f"{enc}.convnormrelu", "im2col7", batch), blocks=[])
            for li, stride in zip((1, 2, 3), cfg["encoder_strides"][1:]):
                for bi in range(2):
                    pfx = f"{enc}.layer{li}.{bi}"
                    s2 = bi == 0 and stride != 1
                    E["blocks"].append(dict(
                        s2=s2,
                        c1=self._unit(sd, pfx + ".convnormrelu1", "im2col3" if s2 else "taps", batch),
                        c2=self._unit(sd, pfx + ".convnormrelu2", "taps", batch),
                        ds=self._unit(sd, pfx + ".downsample", "lin", batch) if s2 else None))
            E["conv"] = (ops.pack_linear(sd[f"{enc}.conv.weight"]), _h(sd[f"{enc}.conv.bias"], device))
            self.enc[enc] = E
        me = "update_block.motion_encoder"
        hid = cfg["hidden"]

        def taps(name, co_pad=None):
            w, b = sd[name + ".weight"], sd[name + ".bias"]
            if co_pad is not None and co_pad > b.shape[0]:
                b = torch.cat([b, b.new_zeros(co_pad - b.shape[0])])
            return ops.pack_conv_taps(w, co_pad=co_pad), _h(b, device)

        self.convcorr1 = (ops.pack_linear(sd[me + ".convcorr1.0.weight"]), _h(sd[me + ".convcorr1.0.bias"], device))
        self.convcorr2 = taps(me + ".convcorr2.0")
        self.convflow1 = (ops.pack_conv_im2col(sd[me + ".convflow1.0.weight"], c_pad=8),
                          _h(sd[me + ".convflow1.0.bias"], device))
        self.convflow2 = taps(me + ".convflow2.0")
        self.convmotion = taps(me + ".conv.0", co_pad=cfg["motion_out"])
        self.gru = []
        for gi, (kh, kw) in enumerate(cfg["gru_kernels"], start=1):
            p = f"update_block.recurrent_block.convgru{gi}"
            wq = sd[p + ".convq.weight"].clone()
            wq_h = wq[:, :hid].contiguous()
            wq[:, :hid] = 0  # the r*h part of convq runs as its own GEMM once r is known
            w3 = torch.cat([sd[p + ".convz.weight"], sd[p + ".convr.weight"], wq], dim=0)
            b3 = torch.cat([sd[p + ".convz.bias"], sd[p + ".convr.bias"], sd[p + ".convq.bias"]])
            self.gru.append(dict(hw=(kh, kw), w3=ops.pack_conv_taps(w3), b3=_h(b3, device),
                                 wqh=ops.pack_conv_taps(wq_h)))
        self.fh1 = taps("update_block.flow_head.conv1")
        self.fh2 = taps("update_block.flow_head.conv2", co_pad=8)
        self.mk1 = taps("mask_predictor.convrelu.0")
        mm = cfg["mask_multiplier"]
        self.mk2 = (ops.pack_linear(sd["mask_predictor.conv.weight"] * mm), _h(sd["mask_predictor.conv.bias"] * mm, device))

    # ---- packing of one conv(+norm) unit ---------------------------------------------------------------------
    def _unit(self, sd, pfx, kind, batch):
        u = _Unit()
        w, b = sd[pfx + ".0.weight"], sd[pfx + ".0.bias"]
        u.kind, u.co = kind, w.shape[0]
        u.gamma = u.beta = None
        if not batch:
            u.norm = "instance"
        elif self.training:
            u.norm = "batch"
            u.gamma, u.beta = _h(sd[pfx + ".1.weight"], self.device), _h(sd[pfx + ".1.bias"], self.device)
        else:  # eval-mode BatchNorm is a per-channel affine: fold it into the convolution
            u.norm = "folded"
            s = sd[pfx + ".1.weight"] / torch.sqrt(sd[pfx + ".1.running_var"] + 1e-5)
            w = w * s.view(-1, 1, 1, 1)
            b = (b - sd[pfx + ".1.running_mean"]) * s + sd[pfx + ".1.bias"]
        if kind == "im2col7":
            u.w = ops.pack_conv_im2col(w, c_pad=8)
        elif kind == "im2col3":
            u.w = ops.pack_conv_im2col(w)
        elif kind == "taps":
            u.w = ops.pack_conv_taps(w)
        else:
            u.w = ops.pack_linear(w)
        u.b = _h(b, self.device)
        return u

    # ---- conv -> norm -> relu (-> residual join) -----------------------------------------------------------------
    def _cna(self, a, u, n_imgs, hw, relu, residual=None, **gemm_geom):
        L = _lib.load()
        folded = u.norm == "folded"
        y = ops.gemm(a, u.w, bias=u.b, relu=folded and relu, **gemm_geom)
        if folded:
            if residual is not None:
                _lib.check(L.ivv_add_relu(ops._p(y), ops._p(residual), ops._p(y), y.numel(), ops._s()), "ivv_add_relu")
                ops._count()
            return y
        return ops.channelnorm(y, n_imgs, hw, n_imgs if u.norm == "batch" else 1, u.gamma, u.beta, 1e-5, relu, residual,
                               out=y)

    def _encoder(self, x, n, H, W, E):
        cols, h, w = ops.im2col(x, n, H, W, 7, 7, 2, 3, 3)
        y = self._cna(cols, E["stem"], n, h * w, True, n_img=1, h=1, w=n * h * w, c=cols.shape[1])
        for blk in E["blocks"]:
            if blk["s2"]:
                c = y.shape[1]
                cols, h, w = ops.im2col(y, n, h, w, 3, 3, 2, 1, 1)
                rows = n * h * w
                t = self._cna(cols, blk["c1"], n, h * w, True, n_img=1, h=1, w=rows, c=9 * c)
                # the 1x1 stride-2 projection reads the centre tap of the same im2col matrix
                skip = self._cna(cols[:, 4 * c:5 * c], blk["ds"], n, h * w, False, n_img=1, h=1, w=rows, c=c)
            else:
                t = self._cna(y, blk["c1"], n, h * w, True, n_img=n, h=h, w=w, c=y.shape[1], taps=9)
                skip = y
            y = self._cna(t, blk["c2"], n, h * w, True, residual=skip, n_img=n, h=h, w=w, c=t.shape[1], taps=9)
        return ops.linear(y, E["conv"][0], bias=E["conv"][1]), h, w

    # ---- the whole estimator ---------------------------------------------------------------------------------------
    def run(self, img1, img2, size, num_flow_updates=12):
        L, dev, cfg = _lib.load(), self.device, self.cfg
        B, _, Hs, Ws = img1.shape
        H, W = size
        hid, cm = cfg["hidden"], cfg["motion_out"]
        x = ops.empty((2 * B * H * W, 8), F16, dev)
        for i, img in enumerate((img1, img2)):
            _lib.check(L.ivv_raft_prep_images(ops._p(img), ctypes.c_void_p(x.data_ptr() + i * B * H * W * 16), B, Hs, Ws,
                                              H, W, ops._s()), "ivv_raft_prep_images")
            ops._count()
        fmaps, h, w = self._encoder(x, 2 * B, H, W, self.enc["feature_encoder"])
        ctx, _, _ = self._encoder(x[:B * H * W], B, H, W, self.enc["context_encoder"])
        hw, rows = h * w, B * h * w
        cf = fmaps.shape[1]

        # all-pairs correlation (un-normalised; 1/sqrt(c) is applied by the look-up) and its average-pooled pyramid
        pyr = [ops.empty((rows, (h >> l) * (w >> l)), torch.float32, dev) for l in range(cfg["corr_levels"])]
        for b in range(B):
            ops.gemm(fmaps[b * hw:(b + 1) * hw], fmaps[(B + b) * hw:(B + b + 1) * hw].unsqueeze(0), n_img=1, h=1, w=hw,
                     c=cf, out=pyr[0][b * hw:(b + 1) * hw])
        for l in range(1, cfg["corr_levels"]):
            _lib.check(L.ivv_avgpool2_f32(ops._p(pyr[l - 1]), ops._p(pyr[l]), rows, h >> (l - 1), w >> (l - 1), ops._s()),
                       "ivv_avgpool2_f32")
            ops._count()
        pyr_ptrs = (ctypes.c_void_p * len(pyr))(*[p.data_ptr() for p in pyr])

        # recurrent state: hx = [h | context | motion features (incl. the 2 flow channels)], fp32 master of h
        gin = hid + (ctx.shape[1] - hid) + cm
        hx = ops.empty((rows, gin), F16, dev)
        h32 = ops.empty((rows, hid), torch.float32, dev)
        _lib.check(L.ivv_raft_init_state(ops._p(ctx), ctx.shape[1], ops._p(h32), ops._p(hx), gin, rows, hid,
                                         ctx.shape[1] - hid, ops._s()), "ivv_raft_init_state")
        ops._count()
        ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        coords1 = torch.stack([xs, ys], dim=-1).float()[None].repeat(B, 1, 1, 1).contiguous()  # [B, h, w, (x, y)]
        flow8 = ops.empty((rows, 8), F16, dev)
        ncorr = cfg["corr_levels"] * (2 * cfg["corr_radius"] + 1) ** 2
        corr_feat = torch.zeros((rows, (ncorr + 7) // 8 * 8), dtype=F16, device=dev)  # K padding must be zero
        cfbuf = ops.empty((rows, cfg["corr_layers"][1] + cfg["flow_layers"][1]), F16, dev)
        zrq = ops.empty((rows, 3 * hid), F16, dev)
        rh = ops.empty((rows, hid), F16, dev)
        qpre = ops.empty((rows, hid), F16, dev)
        delta = ops.empty((rows, 8), torch.float32, dev)
        motion = hx[:, gin - cm:]
        flow_slot = ctypes.c_void_p(hx.data_ptr() + (gin - 2) * 2)

        def update_coords(with_delta, slot):
            _lib.check(L.ivv_raft_update_coords(ops._p(delta) if with_delta else None, 8, ops._p(coords1), ops._p(flow8),
                                                slot, gin, B, h, w, ops._s()), "ivv_raft_update_coords")
            ops._count()

        update_coords(False, None)
        g = dict(n_img=B, h=h, w=w)
        for _ in range(num_flow_updates):
            _lib.check(L.ivv_corr_lookup(pyr_ptrs, len(pyr), ops._p(coords1), ops._p(corr_feat), corr_feat.shape[1], B, h,
                                         w, cfg["corr_radius"], float(cf) ** -0.5, ops._s()), "ivv_corr_lookup")
            ops._count()
            # motion encoder (raft.py:200-211)
            t = ops.gemm(corr_feat, self.convcorr1[0], bias=self.convcorr1[1], relu=True, c=corr_feat.shape[1], **g)
            ops.gemm(t, self.convcorr2[0], bias=self.convcorr2[1], relu=True, c=t.shape[1], taps=9,
                     out=cfbuf[:, :cfg["corr_layers"][1]], **g)
            cols, _, _ = ops.im2col(flow8, B, h, w, 7, 7, 1, 3, 3)
            t = ops.gemm(cols, self.convflow1[0], bias=self.convflow1[1], relu=True, c=cols.shape[1], **g)
            ops.gemm(t, self.convflow2[0], bias=self.convflow2[1], relu=True, c=t.shape[1], taps=9,
                     out=cfbuf[:, cfg["corr_layers"][1]:], **g)
            ops.gemm(cfbuf, self.convmotion[0], bias=self.convmotion[1], relu=True, c=cfbuf.shape[1], taps=9,
                     out=motion, **g)
            update_coords(False, flow_slot)  # the two flow channels behind the 126 conv outputs
            # separable ConvGRU (raft.py:222-229, 265-269)
            for G in self.gru:
                kh, kw = G["hw"]
                ops.gemm(hx, G["w3"], bias=G["b3"], c=gin, taps=kh * kw, tap_hw=(kh, kw), out=zrq, **g)
                _lib.check(L.ivv_gru_gate_r(ops._p(zrq), 3 * hid, ops._p(h32), ops._p(rh), rows, hid, ops._s()),
                           "ivv_gru_gate_r")
                ops.gemm(rh, G["wqh"], c=hid, taps=kh * kw, tap_hw=(kh, kw), residual=zrq[:, 2 * hid:], out=qpre, **g)
                _lib.check(L.ivv_gru_update(ops._p(zrq), 3 * hid, ops._p(qpre), ops._p(h32), ops._p(hx), gin, rows, hid,
                                            ops._s()), "ivv_gru_update")
                _lib.LAUNCH_COUNT += 2
            # flow head (raft.py:283-284) and coordinate update
            t = ops.gemm(hx[:, :hid], self.fh1[0], bias=self.fh1[1], relu=True, c=hid, taps=9, **g)
            ops.gemm(t, self.fh2[0], bias=self.fh2[1], c=t.shape[1], taps=9, out=delta, **g)
            update_coords(True, None)
        # only the last prediction is used by the reference (flow_utils.py:186): mask + convex upsampling once
        t = ops.gemm(hx[:, :hid], self.mk1[0], bias=self.mk1[1], relu=True, c=hid, taps=9, **g)
        mask = ops.gemm(t, self.mk2[0], bias=self.mk2[1], c=t.shape[1], **g)
        out = ops.empty((B, 2, 8 * h, 8 * w), torch.float32, dev)
        _lib.check(L.ivv_convex_upsample(ops._p(mask), mask.shape[1], ops._p(coords1), ops._p(out), B, h, w, ops._s()),
                   "ivv_convex_upsample")
        ops._count()
        return out


class _Graph:
    """CUDA-graph capture of one estimator call (340 launches at 12 iterations): static image buffers, one
    cudaGraphLaunch per call."""

    def __init__(self, eng, img1, img2, size):
        self.a, self.b = img1.clone(), img2.clone()
        stream = torch.cuda.Stream(device=img1.device)
        stream.wait_stream(torch.cuda.current_stream())
        ops.WS.high_water = 0
        with torch.cuda.stream(stream):
            eng.run(self.a, self.b, size)  # warm-up: lazy kernel attribute setup happens outside the capture
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        # graph-owned norm-statistics workspace (ticket counters), zero-filled outside capture (see ops.WS)
        self.ws = torch.zeros(max(ops.WS.high_water, 1 << 16), dtype=torch.uint8, device=img1.device)
        n0 = _lib.LAUNCH_COUNT
        self.graph = torch.cuda.CUDAGraph()
        ops.WS.override = self.ws
        try:
            with torch.cuda.graph(self.graph):
                self.out = eng.run(self.a, self.b, size)
        finally:
            ops.WS.override = None
        self.n_launches = _lib.LAUNCH_COUNT - n0

    def replay(self, img1, img2):
        self.a.copy_(img1)
        self.b.copy_(img2)
        self.graph.replay()
        _lib.LAUNCH_COUNT += self.n_launches
        return self.out.clone()


class RAFTFlow(nn.Module):
    """Same call contract as the reference class: `flow = RAFTFlow()(img1, img2, img_size=None)` with images
    [B, 3, H, W] in [0, 1]; `flow` [B, 2, H, W] warps img2 onto img1 (`warp_image(img2, flow)`).

    The reference constructor downloads `Raft_Large_Weights.DEFAULT`; there is no network here, so weights are
    loaded by the caller: `RAFTFlow(weights=<state dict of raft_large>)` or `.model.load_state_dict(...)` /
    `.load_state_dict(...)` (keys `model.*`, exactly the reference module's)."""

    def __init__(self, *args, weights=None):
        super().__init__(*args)
        self.model = _Node()
        for key, shape in _schema().items():
            _register(self.model, key, shape)
        self._engine = None
        self._graphs = {}
        self.use_cuda_graph = True
        if weights is not None:
            self.model.load_state_dict(weights)

    def _apply(self, fn, recurse=True):
        self._engine = None
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def train(self, mode=True):
        self._engine = None
        return super().train(mode)

    @torch.no_grad()
    def forward(self, img1, img2, img_size=None):
        if not img1.is_cuda:
            raise RuntimeError("insv2v_b200.RAFTFlow runs only on CUDA (sm_100a); there is no CPU path")
        if img1.shape != img2.shape or img1.dim() != 4 or img1.shape[1] != 3:
            raise ValueError(f"expected two [B, 3, H, W] batches of equal shape, got {tuple(img1.shape)} and "
                             f"{tuple(img2.shape)}")
        original = tuple(img1.shape[2:])
        size = original if img_size is None else (int(img_size[0]), int(img_size[1]))
        hh, ww = size
        if not ((hh % 8 == 0) and (ww % 8 == 0)):  # torchvision RAFT.forward, raft.py:488-489
            raise ValueError(f"input image H and W should be divisible by 8, instead got {hh} (h) and {ww} (w)")
        if min(hh, ww) // 8 < 16:  # CorrBlock.build_pyramid, raft.py:371-379
            raise ValueError("Feature maps are too small to be down-sampled by the correlation pyramid. H and W of "
                             f"feature maps should be at least 16; got: {(hh // 8, ww // 8)}.")
        dev = img1.device
        eng = self._engine
        stamp = sum(t._version for t in self.model.state_dict(keep_vars=True).values())  # in-place weight updates
        if eng is not None and eng.stamp != stamp:
            eng = None
        if eng is None or eng.device != dev or eng.training != self.training:
            # the packed engine is rebuilt whenever weights, device or train/eval mode change
            eng = self._engine = _Engine(self.model, dev, self.training)
            eng.stamp = stamp
            self._graphs = {}
        img1, img2 = img1.float().contiguous(), img2.float().contiguous()
        if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            key = (tuple(img1.shape), size)
            gr = self._graphs.get(key)
            if gr is None:
                if len(self._graphs) >= 4:
                    self._graphs.clear()
                gr = self._graphs[key] = _Graph(eng, img1, img2, size)
            flow = gr.replay(img1, img2)
        else:
            flow = eng.run(img1, img2, size)
        if img_size is not None:
            flow = ops.resize_flow_f32(flow, original[0], original[1])
        return flow
