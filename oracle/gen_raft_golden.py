"""Generate tests/golden/raft_small.pt: seeded inputs and the flow of the RAFT oracle, after asserting that the oracle
reproduces torchvision's own raft_large module (the third-party network the reference's RAFTFlow wraps,
misc_utils/flow_utils.py:155-159) bit for bit on those inputs. Run in the build container:
    python oracle/gen_raft_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import raft_oracle as ro  # noqa: E402

SEED_W, SEED_X = 11, 12


def inputs(b, h, w, seed=SEED_X):
    """Two smooth random images in [0, 1] quantised to 8 bits (so the fixture stores them exactly as uint8)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(b, 3, h // 8, w // 8, generator=g)
    img1 = torch.nn.functional.interpolate(base, size=(h, w), mode="bilinear", align_corners=False)
    img1 = (img1 + 0.15 * torch.rand(b, 3, h, w, generator=g)).clamp(0, 1)
    img2 = torch.roll(img1, shifts=(2, -3), dims=(2, 3))
    img2 = (img2 + 0.05 * torch.rand(b, 3, h, w, generator=g)).clamp(0, 1)
    q = lambda t: (t * 255).round().to(torch.uint8)
    return q(img1), q(img2)


def main():
    from torchvision.models.optical_flow import raft_large
    sd = ro.raft_seeded_state_dict(SEED_W)
    img1_u8, img2_u8 = inputs(2, 128, 160)
    img1, img2 = img1_u8.float() / 255, img2_u8.float() / 255
    with torch.no_grad():
        flow = ro.raft_flow(sd, img1, img2)
        tv = raft_large(weights=None)
        tv.load_state_dict(sd)
        tv.train()  # the reference never calls .eval() on RAFTFlow (inference.py:294)
        ref = tv((img1 - 0.5) / 0.5, (img2 - 0.5) / 0.5)[-1]
        assert torch.equal(ref, flow), float((ref - flow).abs().max())
        flow_resized = ro.raft_flow(sd, img1, img2, img_size=(128, 128))
    out = dict(seed_w=SEED_W, img1=img1_u8, img2=img2_u8, flow=flow, flow_resized_128=flow_resized)
    path = os.path.join(ROOT, "tests", "golden", "raft_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; |flow| mean", float(flow.abs().mean()), "max", float(flow.abs().max()))


if __name__ == "__main__":
    main()
