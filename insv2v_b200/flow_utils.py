"""Drop-in for `misc_utils.flow_utils.warp_image` / `resize_flow` (flow_utils.py:25-86) on the sm_100a gather kernels,
plus the fused per-step motion-compensated noise correction that replaces the 12-iteration Python loop of
`InferenceIP2PVideoOpticalFlow.second_clip_forward` (pl_trainer/inference/inference.py:374-386)."""
import torch

from . import ops
from .raft import RAFTFlow  # noqa: F401  (drop-in for misc_utils.flow_utils.RAFTFlow, flow_utils.py:134-189)


def _f32c(t):
    return t.to(torch.float32).contiguous()


def warp_image(image, flow, mode="bilinear"):
    """Warp `image` [N,C,H,W] (or [C,H,W]) with `flow` [N,2,H,W] (or [2,H,W]): bilinear, zeros padding,
    align_corners=True, sampling position = pixel + flow."""
    if mode != "bilinear":
        raise NotImplementedError("only mode='bilinear' (the reference's only use) is implemented")
    if image.dim() == 3:
        image = image.unsqueeze(0)
    if flow.dim() == 3:
        flow = flow.unsqueeze(0)
    if image.device != flow.device:
        flow = flow.to(image.device)
    assert image.shape[0] == flow.shape[0], \
        f"Batch size of image and flow must be the same. Got {image.shape[0]} and {flow.shape[0]}."
    assert image.shape[2:] == flow.shape[2:], \
        f"Height and width of image and flow must be the same. Got {image.shape[2:]} and {flow.shape[2:]}."
    if not image.is_cuda:
        raise RuntimeError("insv2v_b200.flow_utils.warp_image runs only on CUDA; there is no CPU path")
    out = ops.warp_image_f32(_f32c(image), _f32c(flow))
    return out if image.dtype == torch.float32 else out.to(image.dtype)


def resize_flow(flow, size):
    """Resize `flow` [B,2,h,w] to (H,W): u,v scaled by W/w, H/h, then bilinear (align_corners=False). The input is
    not modified."""
    if not flow.is_cuda:
        raise RuntimeError("insv2v_b200.flow_utils.resize_flow runs only on CUDA; there is no CPU path")
    H, W = size
    out = ops.resize_flow_f32(_f32c(flow), int(H), int(W))
    return out if flow.dtype == torch.float32 else out.to(flow.dtype)


def flow_noise_correction_(noise_pred_query, delta_noise_ref, flows_latent):
    """In place: for every query frame q, noise_pred_query[q] += masked mean over reference frames r of
    warp(delta_noise_ref[r], flows_latent[q, r]) where the summed warped ones-mask exceeds 0.5.
    noise_pred_query [Q,C,h,w] fp32 contiguous, delta_noise_ref [R,C,h,w], flows_latent [Q,R,2,h,w] (already resized
    to latent resolution with resize_flow — the flows do not change between denoising steps)."""
    return ops.flow_noise_correction_(noise_pred_query, _f32c(delta_noise_ref), _f32c(flows_latent))
