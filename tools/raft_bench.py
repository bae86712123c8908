"""RAFT optical flow at the config-3 shape (4 reference frames of 256x384 px against one query frame):
device time per call of insv2v_b200.raft.RAFTFlow against torchvision's raft_large on the same GPU (fp32, train mode —
what the reference's RAFTFlow runs, misc_utils/flow_utils.py:155-189), plus the agreement of the two flows."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from insv2v_b200 import lib  # noqa: E402
from insv2v_b200.raft import RAFTFlow  # noqa: E402
from oracle import raft_oracle as ro  # noqa: E402  (seeded weights only)


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    from torchvision.models.optical_flow import raft_large
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = ro.raft_seeded_state_dict(41)
    g = torch.Generator().manual_seed(42)
    B, H, W = 4, 256, 384
    query = torch.rand(1, 3, H, W, generator=g).to(dev).repeat(B, 1, 1, 1)
    refs = torch.rand(B, 3, H, W, generator=g).to(dev)
    ours = RAFTFlow(weights=sd).to(dev)
    ours.use_cuda_graph = False
    n0 = lib.LAUNCH_COUNT
    ours(query, refs)
    launches = lib.LAUNCH_COUNT - n0
    ms_eager, flow_eager = timed(lambda: ours(query, refs))
    ours.use_cuda_graph = True
    ms_ours, flow = timed(lambda: ours(query, refs))
    assert torch.equal(flow, flow_eager), "graph replay differs from the eager launch sequence"
    tv = raft_large(weights=None).to(dev)
    tv.load_state_dict(sd)
    tv.train()
    with torch.no_grad():
        ms_tv, flow_tv = timed(lambda: tv((query - 0.5) / 0.5, (refs - 0.5) / 0.5)[-1])
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_tv_tf32, _ = timed(lambda: tv((query - 0.5) / 0.5, (refs - 0.5) / 0.5)[-1])
        with torch.autocast("cuda", dtype=torch.float16):
            ms_tv_amp, _ = timed(lambda: tv((query - 0.5) / 0.5, (refs - 0.5) / 0.5)[-1])
    rel = float((flow - flow_tv).norm() / flow_tv.norm())
    print(json.dumps(dict(shape=[B, 3, H, W], ms_per_call_ours=round(ms_ours, 3), ms_per_call_ours_no_graph=round(ms_eager, 3),
                          launches_per_call=launches,
                          ms_per_call_torchvision_fp32=round(ms_tv, 3), ms_per_call_torchvision_tf32=round(ms_tv_tf32, 3),
                          ms_per_call_torchvision_fp16_autocast=round(ms_tv_amp, 3),
                          rel_l2_vs_torchvision_fp32=rel, max_abs_px=float((flow - flow_tv).abs().max()),
                          flow_mean_abs_px=float(flow_tv.abs().mean()))))


if __name__ == "__main__":
    main()
