"""GPU-resident InsV2V editing loop: the reference's `InferenceIP2PVideo.__call__` / `.second_clip_forward` and
`InferenceIP2PVideoOpticalFlow.second_clip_forward` (pl_trainer/inference/inference.py:163-398) with the DDIM scheduler
of diffusers 0.21.4, restated so that a whole denoising step is: one UNet launch (3 CFG branches in one batch, as in the
reference, :183-194) + one fused CFG-combine/DDIM kernel (+ one fused flow-compensation kernel). The reference's own
sampler classes also run unchanged on `insv2v_b200.unet.UNet3DConditionModel`; this module is the fused variant used by
bench.py and the clip-parallel runner.

Scheduler tables stay on the host as Python floats (the reference indexes `alphas_cumprod` with `int(t)` on the CPU too,
inference.py:182,271)."""
import torch

from . import ops
from .flow_utils import resize_flow


def alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000):
    """scaled_linear schedule of the reference (inference.py:31; diffusers: linspace(sqrt(b0), sqrt(b1), n)**2)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_steps, n_train=1000, steps_offset=1):
    """DDIMScheduler.set_timesteps with timestep_spacing='leading', steps_offset=1 (inference.py:37)."""
    ratio = n_train // num_steps
    return [int(i * ratio + steps_offset) for i in reversed(range(num_steps))]


class InsV2VPipeline:
    def __init__(self, unet, vae=None, num_ddim_steps=20, scale_factor=0.18215, beta_start=0.00085, beta_end=0.012,
                 flow_estimator=None):
        self.unet, self.vae = unet, vae
        self.flow_estimator = flow_estimator  # insv2v_b200.raft.RAFTFlow (InferenceIP2PVideoOpticalFlow, inference.py:294)
        self.num_ddim_steps = num_ddim_steps
        self.scale_factor = scale_factor
        self.ac = alphas_cumprod(beta_start, beta_end).tolist()
        self.timesteps = ddim_timesteps(num_ddim_steps)

    @torch.no_grad()
    def obtain_flows(self, ref_images, query_images):
        """obtain_flow_batched (inference.py:303-311): for every query frame, the RAFT flow from the query (repeated
        over the batch) to each of the R reference frames. ref_images [R, 3, H, W], query_images [Q, 3, H, W] in
        [0, 1]. Returns Q tensors [R, 2, H, W] - the `flows=` argument of denoise()."""
        if self.flow_estimator is None:
            raise RuntimeError("no flow_estimator: build the pipeline with flow_estimator=insv2v_b200.raft.RAFTFlow(...)")
        r = ref_images.shape[0]
        return [self.flow_estimator(q.unsqueeze(0).repeat(r, 1, 1, 1), ref_images) for q in query_images]

    @torch.no_grad()
    def denoise(self, latent, text_cond, text_uncond, img_cond, text_cfg=7.5, img_cfg=1.2, latent_ref=None,
                noise_correct_step=1.0, flows=None, ref_images=None, query_images=None):
        """latent, img_cond [1, F, 4, h, w]; text_* [1, 77, C]; latent_ref [1, R, 4, h, w] (chained clips);
        flows: list of Q tensors [R, 2, H, W] at pixel resolution (optical-flow variant), or ref_images
        [1, R, 3, H, W] + query_images [1, Q, 3, H, W] as in second_clip_forward (inference.py:313-345), from which
        the flows are estimated with RAFT first. Returns the final latent."""
        if latent.shape[0] != 1:
            raise ValueError("one clip per call (shard clips across GPUs with insv2v_b200.parallel)")
        if flows is None and ref_images is not None:
            assert ref_images.shape[0] == 1, 'only support batch size 1'
            flows = self.obtain_flows(ref_images[0], query_images[0])
        dev = latent.device
        _, f, c, h, w = latent.shape
        n = latent.numel()
        lat = latent.to(torch.float32).contiguous().clone()
        # static UNet input [3, 8, F, h, w]: latent channels rewritten every step, condition channels once
        x_in = torch.zeros(3, 2 * c, f, h, w, device=dev, dtype=torch.float32)
        cond = img_cond[0].to(torch.float32).permute(1, 0, 2, 3)  # [4, F, h, w]
        x_in[1, c:] = cond
        x_in[2, c:] = cond
        ctx = torch.cat([text_uncond, text_uncond, text_cond], dim=0).contiguous()
        t_dev = torch.empty(3, device=dev, dtype=torch.float32)
        n_train = len(self.ac)
        step_ratio = n_train // self.num_ddim_steps
        flows_lat = None
        if flows is not None:
            # flows are constant over the denoising steps: resize once (the reference redoes it every step)
            flows_lat = torch.stack([resize_flow(fl.to(dev), (h, w)) for fl in flows], dim=0).contiguous()
        eps = torch.empty_like(lat)
        for i, t in enumerate(self.timesteps):
            x_in[:, :c] = lat[0].permute(1, 0, 2, 3)
            t_dev.fill_(float(t))
            eps3 = self.unet(x_in, t_dev, encoder_hidden_states=ctx).sample  # [3, 4, F, h, w] fp32
            # 'b c f h w -> b f c h w' view for the combine kernel: it is elementwise, so permute latent instead
            eps3 = eps3.permute(0, 2, 1, 3, 4).contiguous()  # [3, F, 4, h, w]
            a_t = self.ac[t]
            prev_t = t - step_ratio
            a_prev = self.ac[prev_t] if prev_t >= 0 else self.ac[0]
            correct = latent_ref is not None and noise_correct_step * self.num_ddim_steps > i
            if not correct:
                ops.cfg_ddim_step_(eps3.reshape(3, n), lat.reshape(n), text_cfg, img_cfg, a_t, a_prev)
                continue
            # chained clip: overwrite eps on the reference frames from the known clean latents and propagate
            # the correction to the new frames (inference.py:270-277 / 367-386)
            r = latent_ref.shape[1]
            ops.cfg_ddim_step_(eps3.reshape(3, n), lat.clone().reshape(n), text_cfg, img_cfg, a_t, a_prev,
                               eps_out=eps.reshape(n))  # eps_out = combined eps; the latent update is redone below
            noise_ref = (lat[:, :r] - (a_t ** 0.5) * latent_ref.to(torch.float32)) / ((1 - a_t) ** 0.5)
            delta = noise_ref - eps[:, :r]
            eps[:, :r] += delta
            if flows_lat is None:
                eps[:, r:] += delta.mean(dim=1, keepdim=True)
            else:
                q = eps.shape[1] - r
                ops.flow_noise_correction_(eps[0, r:], delta[0].contiguous(), flows_lat[:q])
            x0 = (lat - ((1 - a_t) ** 0.5) * eps) / (a_t ** 0.5)
            lat = (a_prev ** 0.5) * x0 + ((1 - a_prev) ** 0.5) * eps
        return lat.to(latent.dtype)

    @torch.no_grad()
    def decode(self, latents):
        """latents [1, F, 4, h, w] -> frames [1, F, 3, 8h, 8w] (instruct_p2p_video.py:66-79 with all frames in one
        batch; latent / scale_factor first, diffusion.py:247-249)."""
        b, f = latents.shape[:2]
        z = (latents.reshape(b * f, *latents.shape[2:]).to(torch.float32) / self.scale_factor)
        img = self.vae.decode(z)
        return img.reshape(b, f, *img.shape[1:])

    @torch.no_grad()
    def edit_clip(self, latent, text_cond, text_uncond, img_cond, **kw):
        return self.decode(self.denoise(latent, text_cond, text_uncond, img_cond, **kw))
