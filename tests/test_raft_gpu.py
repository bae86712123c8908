"""RAFT optical flow on the GPU (`insv2v_b200.raft.RAFTFlow`, the drop-in for misc_utils/flow_utils.py:134-189):
per-kernel parity against plain PyTorch statements of the same op, and whole-estimator parity against
oracle/raft_oracle.py (itself pinned bit-for-bit against torchvision's raft_large, tests/test_oracle_cpu.py) and the
committed golden vectors. Run with `pytest -m gpu`."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import err_stats, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL, ATOL = 1e-3, 1e-4  # BASELINE.json north_star tolerance for fp16 outputs


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


def _lib():
    from insv2v_b200 import lib
    return lib


def h16(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV, torch.float16)


def frames(x):  # [n, c, h, w] -> [n*h*w, c]
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous()


def nchw(fr, n, h, w):
    return fr.reshape(n, h, w, -1).permute(0, 3, 1, 2)


def report(name, got, ref, rtol=RTOL, atol=ATOL):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    bad = (err > atol + rtol * ref.abs()).float().mean().item()
    print(f"[{name}] max_abs={err.max().item():.3e} max_ref={ref.abs().max().item():.3e} viol_frac={bad:.3e}")
    assert torch.isfinite(got).all(), f"{name}: non-finite output"
    assert bad == 0.0, f"{name}: {bad:.3e} of elements outside rtol={rtol} atol={atol}"


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("n,c,h,w,k,stride,pad", [(3, 8, 32, 40, 7, 2, 3), (2, 64, 17, 23, 3, 2, 1), (2, 8, 16, 24, 7, 1, 3)])
def test_im2col(n, c, h, w, k, stride, pad):
    ops = _ops()
    x = h16(n, c, h, w, seed=1)
    cols, ho, wo = ops.im2col(frames(x), n, h, w, k, k, stride, pad, pad)
    ref = F.unfold(x.float(), k, padding=pad, stride=stride)  # [n, c*k*k, L] with channel-major columns
    ref = ref.reshape(n, c, k * k, ho * wo).permute(0, 3, 2, 1).reshape(n * ho * wo, k * k * c)
    assert torch.equal(cols.float(), ref)  # pure data movement: bit exact


@pytest.mark.parametrize("kh,kw", [(1, 5), (5, 1), (3, 3)])
def test_conv_taps_relu_and_strided_views(kh, kw):
    """The GRU's separable convolutions as implicit taps; ReLU epilogue; input, output and residual given as column
    slices of wider buffers."""
    ops = _ops()
    n, ci, co, h, w = 4, 128, 128, 32, 48
    big_in = h16(n * h * w, 384, seed=1)
    x = big_in[:, 128:256]
    wt = h16(co, ci, kh, kw, scale=(kh * kw * ci) ** -0.5, seed=2)
    b = h16(co, seed=3, scale=0.1)
    big_res = h16(n * h * w, 384, seed=4)
    big_out = torch.zeros(n * h * w, 256, device=DEV, dtype=torch.float16)
    ops.gemm(x, ops.pack_conv_taps(wt), n_img=n, h=h, w=w, c=ci, taps=kh * kw, tap_hw=(kh, kw), bias=b, relu=True,
             residual=big_res[:, 256:], out=big_out[:, 64:192])
    xin = nchw(x, n, h, w).double()
    ref = F.conv2d(xin, wt.double(), b.double(), padding=((kh - 1) // 2, (kw - 1) // 2))
    ref = F.relu(ref + nchw(big_res[:, 256:], n, h, w).double())
    report(f"conv {kh}x{kw}+res+relu", nchw(big_out[:, 64:192], n, h, w), ref)
    assert float(big_out[:, :64].abs().max()) == 0.0 and float(big_out[:, 192:].abs().max()) == 0.0


def test_conv_taps_f32_out_small_n():
    """flow head: 3x3, 256 -> 2 channels (padded to 8), fp32 output."""
    ops = _ops()
    n, ci, h, w = 2, 256, 16, 20
    x = h16(n, ci, h, w, seed=1)
    wt = h16(2, ci, 3, 3, scale=(9 * ci) ** -0.5, seed=2)
    b = h16(2, seed=3)
    out = torch.empty(n * h * w, 8, device=DEV, dtype=torch.float32)
    ops.gemm(frames(x), ops.pack_conv_taps(wt, co_pad=8), n_img=n, h=h, w=w, c=ci, taps=9,
             bias=torch.cat([b, b.new_zeros(6)]), out=out)
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=1)
    report("flow head", nchw(out[:, :2], n, h, w), ref, rtol=1e-4, atol=1e-5)
    assert float(out[:, 2:].abs().max()) == 0.0


@pytest.mark.parametrize("n,c,h,w,ipg,relu,res,affine", [(8, 64, 32, 40, 1, True, False, False),
                                                        (4, 96, 16, 20, 4, True, True, True),
                                                        (2, 128, 8, 12, 2, False, False, True),
                                                        (4, 256, 16, 24, 1, True, True, False)])
def test_channelnorm(n, c, h, w, ipg, relu, res, affine):
    ops = _ops()
    x = h16(n, c, h, w, seed=1, scale=2.0) + 0.5
    gamma = (1 + 0.2 * h16(c, seed=2)) if affine else None
    beta = 0.3 * h16(c, seed=3) if affine else None
    r = h16(n, c, h, w, seed=4) if res else None
    y = ops.channelnorm(frames(x), n, h * w, ipg, gamma, beta, 1e-5, relu, frames(r) if res else None)
    xf = x.float().reshape(n // ipg, ipg, c, h, w)
    mean = xf.mean(dim=(1, 3, 4), keepdim=True)
    var = xf.var(dim=(1, 3, 4), unbiased=False, keepdim=True)
    ref = ((xf - mean) / torch.sqrt(var + 1e-5)).reshape(n, c, h, w)
    if affine:
        ref = ref * gamma.float().view(1, c, 1, 1) + beta.float().view(1, c, 1, 1)
    if relu:
        ref = F.relu(ref)
    if res:
        ref = F.relu(ref + r.float())
    report(f"channelnorm c{c} ipg{ipg}", nchw(y, n, h, w), ref, rtol=2e-3, atol=2e-3)


def _pyramid(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    lvl0 = torch.randn(b * h * w, 1, h, w, generator=g)
    pyr = [lvl0]
    for _ in range(3):
        pyr.append(F.avg_pool2d(pyr[-1], 2, 2))
    return pyr


def test_avgpool_and_corr_lookup():
    from oracle import raft_oracle as ro
    L, ops = _lib().load(), _ops()
    b, h, w = 2, 16, 20
    pyr = _pyramid(b, h, w, 5)
    rows = b * h * w
    dev_pyr = [pyr[0].reshape(rows, -1).to(DEV).contiguous()]
    for l in range(1, 4):
        nxt = torch.empty(rows, (h >> l) * (w >> l), device=DEV)
        _lib().check(L.ivv_avgpool2_f32(ops._p(dev_pyr[-1]), ops._p(nxt), rows, h >> (l - 1), w >> (l - 1), ops._s()), "pool")
        torch.testing.assert_close(nxt.cpu(), pyr[l].reshape(rows, -1), rtol=1e-6, atol=1e-6)
        dev_pyr.append(nxt)
    g = torch.Generator().manual_seed(6)
    coords = ro.coords_grid(b, h, w) + 3.0 * torch.randn(b, 2, h, w, generator=g)  # some windows leave the map
    ref = ro.index_pyramid(pyr, coords, 4)  # [b, 324, h, w]; oracle has no 1/sqrt(c) here (it is in build_pyramid)
    out = torch.zeros(rows, 328, device=DEV, dtype=torch.float16)
    ptrs = (ctypes.c_void_p * 4)(*[p.data_ptr() for p in dev_pyr])
    c_dev = coords.permute(0, 2, 3, 1).contiguous().to(DEV)
    _lib().check(L.ivv_corr_lookup(ptrs, 4, ops._p(c_dev), ops._p(out), 328, b, h, w, 4, 0.5, ops._s()), "lookup")
    got = out[:, :324].float().reshape(b, h, w, 324).permute(0, 3, 1, 2)
    report("corr lookup", got, 0.5 * ref, rtol=1e-3, atol=1e-3)
    assert float(out[:, 324:].abs().max()) == 0.0


def test_gru_gates_state_and_coords():
    L, ops = _lib().load(), _ops()
    rows, hid = 4 * 16 * 20, 128
    zrq = h16(rows, 3 * hid, seed=1)
    qpre = h16(rows, hid, seed=2)
    h32 = torch.tanh(torch.randn(rows, hid, generator=torch.Generator().manual_seed(3))).to(DEV)
    hx = torch.zeros(rows, 384, device=DEV, dtype=torch.float16)
    rh = torch.empty(rows, hid, device=DEV, dtype=torch.float16)
    _lib().check(L.ivv_gru_gate_r(ops._p(zrq), 3 * hid, ops._p(h32), ops._p(rh), rows, hid, ops._s()), "gate_r")
    report("gru r*h", rh, torch.sigmoid(zrq[:, hid:2 * hid].float()) * h32)
    z = torch.sigmoid(zrq[:, :hid].float())
    ref_h = (1 - z) * h32 + z * torch.tanh(qpre.float())
    _lib().check(L.ivv_gru_update(ops._p(zrq), 3 * hid, ops._p(qpre), ops._p(h32), ops._p(hx), 384, rows, hid, ops._s()),
                 "gru_update")
    torch.testing.assert_close(h32, ref_h, rtol=1e-5, atol=1e-5)
    report("gru h fp16", hx[:, :hid], ref_h)
    assert float(hx[:, hid:].abs().max()) == 0.0
    # init state
    ctx = h16(rows, 256, seed=4)
    _lib().check(L.ivv_raft_init_state(ops._p(ctx), 256, ops._p(h32), ops._p(hx), 384, rows, 128, 128, ops._s()), "init")
    torch.testing.assert_close(h32, torch.tanh(ctx[:, :128].float()), rtol=1e-5, atol=1e-6)
    report("init ctx", hx[:, 128:256], F.relu(ctx[:, 128:].float()))
    # coordinate update
    b, h, w = 4, 16, 20
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    grid = torch.stack([xs, ys], -1).float()[None].repeat(b, 1, 1, 1).to(DEV)
    coords1 = (grid + torch.randn(b, h, w, 2, generator=torch.Generator().manual_seed(5)).to(DEV)).contiguous()
    delta = torch.randn(rows, 8, generator=torch.Generator().manual_seed(6)).to(DEV)
    ref_c = coords1 + delta[:, :2].reshape(b, h, w, 2)
    flow8 = torch.empty(rows, 8, device=DEV, dtype=torch.float16)
    slot = ctypes.c_void_p(hx.data_ptr() + 382 * 2)
    _lib().check(L.ivv_raft_update_coords(ops._p(delta), 8, ops._p(coords1), ops._p(flow8), slot, 384, b, h, w, ops._s()),
                 "coords")
    torch.testing.assert_close(coords1, ref_c, rtol=0, atol=1e-6)
    flow = (ref_c - grid).reshape(rows, 2)
    report("flow8", flow8[:, :2], flow)
    report("flow slot", hx[:, 382:384], flow)
    assert float(flow8[:, 2:].abs().max()) == 0.0


def test_convex_upsample():
    from oracle import raft_oracle as ro
    L, ops = _lib().load(), _ops()
    b, h, w = 2, 16, 20
    mask = h16(b * h * w, 576, seed=1, scale=2.0)
    flow = 4.0 * torch.randn(b, 2, h, w, generator=torch.Generator().manual_seed(2))
    coords1 = (ro.coords_grid(b, h, w) + flow).permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.empty(b, 2, 8 * h, 8 * w, device=DEV)
    _lib().check(L.ivv_convex_upsample(ops._p(mask), 576, ops._p(coords1), ops._p(out), b, h, w, ops._s()), "upsample")
    m = mask.float().cpu().reshape(b, h, w, 576).permute(0, 3, 1, 2)
    ref = ro.upsample_flow(flow, m)
    report("convex upsample", out.cpu(), ref, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ estimator
def _oracle():
    from oracle import raft_oracle as ro
    return ro


def _model(seed, train=True):
    from insv2v_b200.raft import RAFTFlow
    m = RAFTFlow(weights=_oracle().raft_seeded_state_dict(seed)).to(DEV)
    return m.train(train)


def test_raft_encoders_match_oracle():
    """Feature (InstanceNorm) and context (batch-statistics BatchNorm) encoders on the golden inputs: 13 fp16
    convolution + norm layers against the fp32 oracle, 0.5 % of the feature maps' L2 norm."""
    from insv2v_b200.raft import _Engine
    ro, L, ops = _oracle(), _lib().load(), _ops()
    g = golden("raft_small.pt")
    sd = ro.raft_seeded_state_dict(g["seed_w"])
    m = _model(g["seed_w"])
    eng = _Engine(m.model, torch.device(DEV, torch.cuda.current_device()), True)
    imgs = torch.cat([g["img1"], g["img2"]]).float().div(255)
    n, _, H, W = imgs.shape
    x = torch.empty(n * H * W, 8, device=DEV, dtype=torch.float16)
    _lib().check(L.ivv_raft_prep_images(ops._p(imgs.to(DEV).contiguous()), ops._p(x), n, H, W, H, W, ops._s()), "prep")
    ref_in = (imgs - 0.5) / 0.5
    report("prep", nchw(x[:, :3], n, H, W), ref_in)
    for enc, batch, nimg in (("feature_encoder", False, n), ("context_encoder", True, n // 2)):
        fm, h, w = eng._encoder(x[:nimg * H * W], nimg, H, W, eng.enc[enc])
        with torch.no_grad():
            ref = ro.feature_encoder(sd, enc, ref_in[:nimg], batch, True)
        st = err_stats(nchw(fm, nimg, h, w), ref)
        print(f"[raft {enc}]", st)
        assert (h, w) == (H // 8, W // 8) and st["rel_l2"] < 5e-3


def test_raft_golden_flow():
    """Committed vectors (oracle == torchvision raft_large bit for bit, oracle/gen_raft_golden.py). The estimator is 12
    recurrent iterations of fp16 convolutions against an fp32 reference: tolerance is 1 % of the flow's L2 norm and
    0.1 px worst case on flows of ~6 px mean / 13 px max."""
    g = golden("raft_small.pt")
    m = _model(g["seed_w"])
    img1, img2 = g["img1"].float().div(255).to(DEV), g["img2"].float().div(255).to(DEV)
    flow = m(img1, img2)
    st = err_stats(flow, g["flow"])
    print("[raft golden]", st)
    assert flow.shape == g["flow"].shape and flow.dtype == torch.float32
    assert st["rel_l2"] < 1e-2 and st["max_abs"] < 0.1
    flow_r = m(img1, img2, img_size=(128, 128))
    st = err_stats(flow_r, g["flow_resized_128"])
    print("[raft golden, img_size]", st)
    assert flow_r.shape == g["flow_resized_128"].shape
    assert st["rel_l2"] < 1e-2 and st["max_abs"] < 0.1


def test_raft_eval_mode_matches_oracle():
    ro = _oracle()
    sd = ro.raft_seeded_state_dict(21)
    m = _model(21, train=False)
    g = torch.Generator().manual_seed(22)
    img1, img2 = torch.rand(1, 3, 128, 128, generator=g), torch.rand(1, 3, 128, 128, generator=g)
    ref = ro.raft_flow(sd, img1, img2, bn_training=False)
    st = err_stats(m(img1.to(DEV), img2.to(DEV)), ref)
    print("[raft eval]", st)
    assert st["rel_l2"] < 1e-2 and st["max_abs"] < 0.1


def test_raft_api_contract():
    from insv2v_b200.raft import RAFTFlow
    from torchvision.models.optical_flow import raft_large
    m = RAFTFlow()
    assert list(m.state_dict().keys()) == ["model." + k for k in raft_large(weights=None).state_dict().keys()]
    with pytest.raises(RuntimeError, match="only on CUDA"):
        m(torch.zeros(1, 3, 128, 128), torch.zeros(1, 3, 128, 128))
    m = m.to(DEV)
    with pytest.raises(ValueError, match="divisible by 8"):
        m(torch.zeros(1, 3, 130, 128, device=DEV), torch.zeros(1, 3, 130, 128, device=DEV))
    with pytest.raises(ValueError, match="too small"):
        m(torch.zeros(1, 3, 64, 128, device=DEV), torch.zeros(1, 3, 64, 128, device=DEV))


def test_raft_full_size_batch_consistency():
    """Config-3 shape (4 reference frames, 256x384 px): the reference repeats the query frame over the
    batch (inference.py:306-308), so every BatchNorm statistic equals the single-image one and each pair's flow must
    equal the flow of that pair computed alone; warping with it must be finite."""
    from insv2v_b200.flow_utils import warp_image
    m = _model(31)
    g = torch.Generator().manual_seed(32)
    query = torch.rand(1, 3, 256, 384, generator=g).to(DEV)
    refs = torch.rand(4, 3, 256, 384, generator=g).to(DEV)
    flow = m(query.repeat(4, 1, 1, 1), refs)
    assert flow.shape == (4, 2, 256, 384) and torch.isfinite(flow).all()
    one = m(query, refs[2:3])
    st = err_stats(flow[2:3], one)
    print("[raft batch consistency]", st)
    assert st["rel_l2"] < 2e-3
    assert torch.isfinite(warp_image(refs, flow)).all()
