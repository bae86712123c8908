"""GPU-resident InsV2V editing loop: the reference's `InferenceIP2PVideo.__call__` / `.second_clip_forward` and
`InferenceIP2PVideoOpticalFlow.second_clip_forward` (pl_trainer/inference/inference.py:163-398) with the DDIM / DDPM
schedulers of diffusers 0.21.4, restated so that a whole denoising step is ONE CUDA graph:

    ivv_sampler_begin  (3-branch UNet input from the latent, inference.py:183-189)
    UNet3D forward     (insv2v_b200.unet._Engine.forward_body, ~1200 kernels)
    ivv_sampler_combine (CFG combine, :198-203, + statistics for rescale_noise_cfg, :13-24)
    ivv_sampler_update (reference-frame noise correction :262-277 / :367-386, scheduler.step)

and the sampling loop is N launches of that graph with nothing in between: the per-step scalars live in a device
table indexed by a device-side step counter (csrc/sampler.cu). `insv2v_b200.inference` wraps this in classes with the
reference's names and call signatures; the reference's own sampler classes also run unchanged on
`insv2v_b200.unet.UNet3DConditionModel`.

The scheduler tables are built on the host in fp32 torch arithmetic, the way the reference's schedulers compute them.
"""
import torch

from . import lib as _lib
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


def ddpm_timesteps(num_steps, n_train=1000):
    """DDPMScheduler.set_timesteps, timestep_spacing='leading' (inference.py:38-40): 950, 900 ... 0 for 20 steps."""
    ratio = n_train // num_steps
    return [int(i * ratio) for i in reversed(range(num_steps))]


def scheduler_timesteps(scheduler, num_steps, n_train=1000):
    if scheduler == "ddim":
        return ddim_timesteps(num_steps, n_train)
    if scheduler == "ddpm":
        return ddpm_timesteps(num_steps, n_train)
    raise NotImplementedError(f"scheduler {scheduler!r} (the reference supports 'ddim' and 'ddpm', inference.py:35-43)")


def sampler_table(scheduler, timesteps, ac, num_steps, text_cfg, img_cfg, guidance_rescale=0.0, n_correct=0):
    """One row of IVV_SAMPLER_ROW floats per executed step (layout: csrc/sampler.cu). `ac` fp32 alphas_cumprod tensor;
    the first `n_correct` rows carry the noise-correction flag. prev = c_x0*x0 + c_xt*x_t + c_eps*eps + sigma*noise:
      DDIM (eta 0, set_alpha_to_one=False): c_x0 = sqrt(a_prev), c_eps = sqrt(1 - a_prev), a_prev = ac[0] past the end;
      DDPM (fixed_small, a_prev = 1 past the end): posterior mean coefficients (DDPM eq. 7), sigma^2 =
      clamp((1-a_prev)/(1-a_t) * beta_t, 1e-20) for t > 0."""
    n_train = ac.shape[0]
    rows = torch.zeros(len(timesteps), ops.SAMPLER_ROW, dtype=torch.float32)
    noise_row = 0
    one = torch.tensor(1.0)
    for i, t in enumerate(timesteps):
        prev_t = t - n_train // num_steps
        a_t = ac[t]
        if scheduler == "ddim":
            a_prev = ac[prev_t] if prev_t >= 0 else ac[0]
            c_x0, c_xt, c_eps, sigma = a_prev ** 0.5, torch.tensor(0.0), (1 - a_prev) ** 0.5, torch.tensor(0.0)
        else:
            a_prev = ac[prev_t] if prev_t >= 0 else one
            cur_alpha = a_t / a_prev
            cur_beta = 1 - cur_alpha
            c_x0 = (a_prev ** 0.5 * cur_beta) / (1 - a_t)
            c_xt = cur_alpha ** 0.5 * (1 - a_prev) / (1 - a_t)
            c_eps = torch.tensor(0.0)
            sigma = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20) ** 0.5 if t > 0 else torch.tensor(0.0)
        rows[i, 0] = float(t)
        rows[i, 1], rows[i, 2] = a_t ** 0.5, (1 - a_t) ** 0.5
        rows[i, 3], rows[i, 4], rows[i, 5], rows[i, 6] = c_x0, c_xt, c_eps, sigma
        rows[i, 7] = 1.0 if i < n_correct else 0.0
        rows[i, 8], rows[i, 9], rows[i, 10] = float(text_cfg), float(img_cfg), float(guidance_rescale)
        rows[i, 11] = float(noise_row)
        if scheduler == "ddpm" and t > 0:
            noise_row += 1
    return rows, noise_row


class _StepGraph:
    """Static buffers + the captured graph of one sampling step for a fixed problem shape."""
    TABLE_ROWS = 1000

    def __init__(self, eng, f, c, h, w, r, q, mode, ctx, pe_start, n_noise, n_hist):
        dev = eng.device
        hw = h * w
        n = f * c * hw
        self.eng, self.shape, self.mode, self.r, self.q = eng, (f, c, h, w), mode, r, q
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)  # noqa: E731
        self.table = z(self.TABLE_ROWS, ops.SAMPLER_ROW)
        self.state = z(4, dt=torch.int32)
        self.lat2, self.cond, self.eps_cfg = z(2, n), z(n), z(n)
        self.partials = z(ops.sampler_partials(f, hw), 4, dt=torch.float64)
        self.latent_ref = z(r * c * hw) if mode else None
        self.flows_lat = z(q * r * 2 * hw) if mode == 2 else None
        self.noise = z(n_noise, n) if n_noise else None
        self.hist_lat = z(n_hist, n) if n_hist else None
        self.hist_pred = z(n_hist, n) if n_hist else None
        self.x = z(3 * f * hw, eng.cin_pad, dt=torch.float16)
        self.t = z(3)
        self.c_in = ctx.clone()
        self.ctx_ref, self.ctx_ver = None, -1
        # warm-up (eager, side stream): lazy kernel attribute set-up must happen outside capture
        self.table[0, 1:3] = 1.0  # sqrt(a_t), sqrt(1 - a_t): anything finite
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream())
        ops.WS.high_water = 0
        with torch.cuda.stream(stream):
            self.ctx_kv = eng._context_kv(self.c_in)
            self._step(pe_start)
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        self.ws = torch.zeros(max(ops.WS.high_water, 1 << 16), dtype=torch.uint8, device=dev)  # see ops.WS
        self.state.zero_()
        n0 = _lib.LAUNCH_COUNT
        self.graph = torch.cuda.CUDAGraph()
        ops.WS.override = self.ws
        try:
            with torch.cuda.graph(self.graph):
                self._step(pe_start)
        finally:
            ops.WS.override = None
        self.n_launches = _lib.LAUNCH_COUNT - n0

    def _step(self, pe_start):
        f, c, h, w = self.shape
        hw = h * w
        ops.sampler_begin(self.table, self.state, self.lat2, self.cond, self.x, self.t, f, c, hw, self.eng.cin_pad)
        eps3, ho, wo = self.eng.forward_body(self.x, self.t, 3, f, h, w, pe_start, self.ctx_kv, self.c_in.shape[1])
        assert (ho, wo) == (h, w) and eps3.shape[1] >= c
        ops.sampler_combine(self.table, self.state, eps3, self.eps_cfg, self.partials, f, c, hw)
        ops.sampler_update(self.table, self.state, self.lat2, self.eps_cfg, self.partials, self.mode, self.latent_ref,
                           self.flows_lat, self.noise, self.hist_lat, self.hist_pred, f, c, self.r, self.q, h, w)

    def load(self, latent, cond, ctx, table, latent_ref=None, flows_lat=None, noise=None):
        """Copy one clip's inputs into the static buffers (stream-ordered, no synchronisation)."""
        k = table.shape[0]
        if k > self.TABLE_ROWS:
            raise ValueError(f"{k} sampling steps exceed the table capacity {self.TABLE_ROWS}")
        self.table[:k].copy_(table, non_blocking=True)
        self.state.zero_()
        self.lat2[0].copy_(latent.reshape(-1))
        self.cond.copy_(cond.reshape(-1))
        if self.latent_ref is not None:
            self.latent_ref.copy_(latent_ref.reshape(-1))
        if self.flows_lat is not None:
            self.flows_lat.copy_(flows_lat.reshape(-1))
        if noise is not None:
            self.noise[:noise.shape[0]].copy_(noise.reshape(noise.shape[0], -1))
        if ctx is not self.ctx_ref or ctx._version != self.ctx_ver:
            self.c_in.copy_(ctx)
            self.eng._context_kv(self.c_in, out=self.ctx_kv)
            self.ctx_ref, self.ctx_ver = ctx, ctx._version
        self.k = k

    def run(self):
        for _ in range(self.k):
            self.graph.replay()
        _lib.LAUNCH_COUNT += self.n_launches * self.k
        return self.lat2[self.k & 1]


class InsV2VPipeline:
    def __init__(self, unet, vae=None, num_ddim_steps=20, scale_factor=0.18215, beta_start=0.00085, beta_end=0.012,
                 flow_estimator=None, scheduler="ddim"):
        self.unet, self.vae = unet, vae
        self.flow_estimator = flow_estimator  # insv2v_b200.raft.RAFTFlow (InferenceIP2PVideoOpticalFlow, inference.py:294)
        self.num_ddim_steps = num_ddim_steps
        self.scale_factor = scale_factor
        self.scheduler = scheduler
        self.ac = alphas_cumprod(beta_start, beta_end)
        self.timesteps = scheduler_timesteps(scheduler, num_ddim_steps, self.ac.shape[0])
        self._graphs = {}
        self._graphs_engine = None

    @torch.no_grad()
    def obtain_flows(self, ref_images, query_images):
        """obtain_flow_batched (inference.py:303-311): for every query frame, the RAFT flow from the query (repeated
        over the batch) to each of the R reference frames. ref_images [R, 3, H, W], query_images [Q, 3, H, W] in
        [0, 1]. Returns Q tensors [R, 2, H, W] - the `flows=` argument of denoise()."""
        if self.flow_estimator is None:
            raise RuntimeError("no flow_estimator: build the pipeline with flow_estimator=insv2v_b200.raft.RAFTFlow(...)")
        r = ref_images.shape[0]
        return [self.flow_estimator(q.unsqueeze(0).repeat(r, 1, 1, 1), ref_images) for q in query_images]

    def _graph(self, eng, key, make):
        if self._graphs_engine is not eng:  # weights re-packed: every captured graph is stale
            self._graphs, self._graphs_engine = {}, eng
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 6:
                self._graphs.clear()
            g = self._graphs[key] = make()
        return g

    @torch.no_grad()
    def denoise(self, latent, text_cond, text_uncond, img_cond, text_cfg=7.5, img_cfg=1.2, latent_ref=None,
                noise_correct_step=1.0, flows=None, ref_images=None, query_images=None, start_time=0,
                guidance_rescale=0.0, return_all=False, noise=None, video_start_index=0):
        """latent, img_cond [1, F, 4, h, w]; text_* [1, 77, C]; latent_ref [1, R, 4, h, w] (chained clips);
        flows: list of Q tensors [R, 2, H, W] at pixel resolution (optical-flow variant), or ref_images
        [1, R, 3, H, W] + query_images [1, Q, 3, H, W] as in second_clip_forward (inference.py:313-345), from which
        the flows are estimated with RAFT first. start_time / guidance_rescale as in the reference (:171-172).
        noise: DDPM variance noise [k, 1, F, 4, h, w], one row per executed step with t > 0, in step order; default:
        torch.randn on the latent's device per step (what diffusers' randn_tensor does for generator=None).
        Returns the final latent, or the reference's {'latent', 'all_latent', 'all_pred'} dict when return_all."""
        if latent.shape[0] != 1:
            raise ValueError("one clip per call (shard clips across GPUs with insv2v_b200.parallel)")
        if not latent.is_cuda:
            raise RuntimeError("InsV2VPipeline runs only on CUDA (sm_100a); there is no CPU path")
        if flows is None and ref_images is not None:
            assert ref_images.shape[0] == 1, 'only support batch size 1'
            flows = self.obtain_flows(ref_images[0], query_images[0])
        dev = latent.device
        _, f, c, h, w = latent.shape
        steps = self.timesteps[start_time:]
        k = len(steps)
        # the reference enumerates the SLICED timesteps (inference.py:181,240): the correction window counts from the
        # first executed step
        r = latent_ref.shape[1] if latent_ref is not None else 0
        n_correct = sum(1 for i in range(k) if noise_correct_step * self.num_ddim_steps > i) if r else 0
        mode, q, flows_lat = 0, 0, None
        if r:
            if not 0 < r < f:
                raise ValueError(f"latent_ref has {r} frames; need 0 < R < F = {f}")
            mode = 1
            if flows is not None:
                mode = 2
                flows = list(flows)[:f - r]  # zip(range(R, F), warp_funcs), inference.py:374
                q = len(flows)
                # flows are constant over the denoising steps: resize once (the reference redoes it every step, :297)
                flows_lat = torch.stack([resize_flow(fl.to(dev), (h, w)) for fl in flows], dim=0).contiguous()
                if tuple(flows_lat.shape) != (q, r, 2, h, w):
                    raise ValueError(f"flows must be Q tensors [R={r}, 2, H, W]; got {tuple(flows_lat.shape)} at latent size")
        table, n_noise = sampler_table(self.scheduler, steps, self.ac, self.num_ddim_steps, text_cfg, img_cfg,
                                       guidance_rescale, n_correct)
        if n_noise:
            if noise is None:
                noise = torch.stack([torch.randn(latent.shape, device=dev, dtype=torch.float32) for _ in range(n_noise)])
            elif noise.shape[0] != n_noise:
                raise ValueError(f"noise must have one row per step with t > 0 ({n_noise}), got {noise.shape[0]}")
        ctx = torch.cat([text_uncond, text_uncond, text_cond], dim=0).contiguous()
        eng = self.unet.engine(dev)
        pe_start = eng.pe_start_for(int(video_start_index), f)
        n_hist = k if return_all else 0
        key = (f, c, h, w, r, q, mode, tuple(ctx.shape), ctx.dtype, pe_start, n_noise, n_hist)
        g = self._graph(eng, key, lambda: _StepGraph(eng, f, c, h, w, r, q, mode, ctx, pe_start, n_noise, n_hist))
        g.load(latent.to(torch.float32), img_cond.to(torch.float32), ctx, table,
               None if latent_ref is None else latent_ref.to(torch.float32), flows_lat,
               None if noise is None else noise.to(device=dev, dtype=torch.float32))
        lat = g.run().reshape(latent.shape).to(latent.dtype, copy=True)  # copy: g.run() returns a static buffer
        if not return_all:
            return lat
        shp = tuple(latent.shape)
        return {"latent": lat,
                "all_latent": [g.hist_lat[i].reshape(shp).to(latent.dtype, copy=True) for i in range(k)],
                "all_pred": [g.hist_pred[i].reshape(shp).to(latent.dtype, copy=True) for i in range(k)]}

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
