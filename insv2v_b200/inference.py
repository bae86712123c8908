"""Drop-in for the reference's sampler classes `pl_trainer.inference.inference.InferenceIP2PVideo` and
`InferenceIP2PVideoOpticalFlow` (inference.py:159-398): same constructor (`unet, scheduler='ddim'|'ddpm', beta_start,
beta_end, beta_schedule, num_ddim_steps, guidance_scale`), same `__call__` / `second_clip_forward` keyword arguments
(`start_time`, `guidance_rescale`, `noise_correct_step`, `ref_images` / `query_images`), same
`{'latent', 'all_latent', 'all_pred'}` return, so `insv2v_run_loveu_tgve.py:64-78,123-161` runs on them unchanged.
The loop itself is `insv2v_b200.pipeline.InsV2VPipeline.denoise`: one CUDA graph per step, nothing in between.

(The reference's own classes also run unchanged on `insv2v_b200.unet.UNet3DConditionModel`; these are the fused
variant. `Inference.__call__` / `InferenceIP2PEditRef`, which drive a different UNet signature (`context={'text': ..}`)
for the ModelScope generator, are outside the InsV2V path and not provided.)"""
import torch

from .pipeline import InsV2VPipeline, alphas_cumprod, scheduler_timesteps


class _SchedulerView:
    """The scheduler attributes callers read: `.timesteps` (LongTensor), `.alphas_cumprod`, `.config`."""

    def __init__(self, name, ac, timesteps, n_train):
        self.name = name
        self.alphas_cumprod = ac
        self.timesteps = torch.tensor(timesteps, dtype=torch.long)
        self.num_inference_steps = len(timesteps)
        self.init_noise_sigma = 1.0
        self.config = dict(num_train_timesteps=n_train, clip_sample=False, prediction_type="epsilon",
                           steps_offset=1 if name == "ddim" else 0)


class Inference:
    def __init__(self, unet, scheduler='ddim', beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 num_ddim_steps=20, guidance_scale=5):
        if scheduler not in ("ddim", "ddpm"):
            raise NotImplementedError()  # inference.py:42-43
        if beta_schedule != "scaled_linear":
            raise NotImplementedError(f"beta_schedule {beta_schedule!r}: the reference only uses 'scaled_linear'")
        self.unet = unet
        self.num_ddim_steps = num_ddim_steps
        self.guidance_scale = guidance_scale
        self.pipe = InsV2VPipeline(unet, None, num_ddim_steps=num_ddim_steps, beta_start=beta_start, beta_end=beta_end,
                                   scheduler=scheduler)
        ac = alphas_cumprod(beta_start, beta_end)
        self.scheduler = _SchedulerView(scheduler, ac, scheduler_timesteps(scheduler, num_ddim_steps, ac.shape[0]),
                                        ac.shape[0])

    def __call__(self, *a, **k):
        raise NotImplementedError("Inference.__call__ drives the ModelScope UNet signature (context={'text': ...}); "
                                  "use InferenceIP2PVideo for InsV2V")


class InferenceIP2PVideo(Inference):
    def zeros(self, x):
        return torch.zeros_like(x)

    @torch.no_grad()
    def __call__(self, latent, text_cond, text_uncond, img_cond, text_cfg=7.5, img_cfg=1.2, start_time: int = 0,
                 guidance_rescale: float = 0.0):
        return self.pipe.denoise(latent, text_cond, text_uncond, img_cond, text_cfg=text_cfg, img_cfg=img_cfg,
                                 start_time=start_time, guidance_rescale=guidance_rescale, return_all=True)

    @torch.no_grad()
    def second_clip_forward(self, latent, text_cond, text_uncond, img_cond, latent_ref, noise_correct_step: float = 1.,
                            text_cfg=7.5, img_cfg=1.2, start_time: int = 0, guidance_rescale: float = 0.0):
        return self.pipe.denoise(latent, text_cond, text_uncond, img_cond, text_cfg=text_cfg, img_cfg=img_cfg,
                                 latent_ref=latent_ref, noise_correct_step=noise_correct_step, start_time=start_time,
                                 guidance_rescale=guidance_rescale, return_all=True)


class InferenceIP2PVideoOpticalFlow(InferenceIP2PVideo):
    """inference.py:291-398. The reference constructor builds `RAFTFlow().cuda()`, which downloads torchvision's
    weights; offline, pass `flow_estimator=insv2v_b200.raft.RAFTFlow(weights=...)` (default: the same architecture with
    its initial parameters)."""

    def __init__(self, *args, flow_estimator=None, **kwargs):
        super().__init__(*args, **kwargs)
        if flow_estimator is None:
            from .raft import RAFTFlow
            flow_estimator = RAFTFlow().cuda()
        self.flow_estimator = flow_estimator
        self.pipe.flow_estimator = flow_estimator

    def obtain_flow_batched(self, ref_images, query_images):
        """inference.py:303-311, returning the flows themselves ([Q] x [R, 2, H, W]) instead of closures."""
        return self.pipe.obtain_flows(ref_images, query_images)

    @torch.no_grad()
    def second_clip_forward(self, latent, text_cond, text_uncond, img_cond, latent_ref, ref_images, query_images,
                            noise_correct_step: float = 1., text_cfg=7.5, img_cfg=1.2, start_time: int = 0,
                            guidance_rescale: float = 0.0):
        assert ref_images.shape[0] == 1, 'only support batch size 1'
        flows = self.obtain_flow_batched(ref_images[0], query_images[0])
        return self.pipe.denoise(latent, text_cond, text_uncond, img_cond, text_cfg=text_cfg, img_cfg=img_cfg,
                                 latent_ref=latent_ref, noise_correct_step=noise_correct_step, flows=flows,
                                 start_time=start_time, guidance_rescale=guidance_rescale, return_all=True)
