/*
 * ivv.h — C ABI of libivv_b200.so: the sm_100a kernels behind the InsV2V denoising hot path.
 *
 * The reference (amazon-science/instruct-video-to-video) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below replaces the library call that the cited reference line dispatches to.
 * Contract: the caller (PyTorch) owns every buffer; the library never allocates device memory, keeps no
 * pointer after return, enqueues on the caller's stream, never synchronises, and is CUDA-graph capturable.
 * Return value: 0 = ok, non-zero = error; the message is available from ivv_last_error() (thread-local).
 *
 * Layout vocabulary: activations are channels-last fp16 "frames": [n_img, h, w, c] with n_img = clips*frames
 * (so the reference's `(b f)` fold, resnet.py:14, is a free view); "tokens" [rows, c] is the same memory.
 */
#ifndef IVV_H_
#define IVV_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ivv_stream_t; /* cudaStream_t */

#define IVV_ABI_VERSION 6

int ivv_abi_version(void);
const char* ivv_last_error(void);

/* ---- K1/K2/K11: tcgen05 implicit-GEMM convolution / linear --------------------------------------------------
 * Replaces InflatedConv3d.forward (modules/video_unet_temporal/resnet.py:10-18), nn.Linear inside diffusers
 * Attention/FeedForward (called at attention.py:241-259, motion_module.py:104,126,289-301,327), nn.Conv2d in the
 * VAE decoder (modules/vqvae/model.py:86-136,386-408).
 *   D[pix, co] = sum_{tap, ci} A[pix + tap_offset, ci] * Wt[tap][co][ci]  (+bias[co]) (+rowbias[pix/group][co])
 *                (+residual[pix, co]);  GEGLU: D[pix, j] = (acc_h + b_h) * gelu_erf(acc_g + b_g)
 * taps = 1 (Linear / 1x1 conv; h = n_img = 1, w = rows is allowed) or 9 (3x3, stride 1, zero pad 1); with
 * tap_h x tap_w set, any odd stride-1 "same" window (RAFT's 1x5 / 5x1 GRU convolutions, torchvision raft.py:216-229):
 * taps = tap_h * tap_w, tap t reads pixel (y + t / tap_w - tap_h / 2, x + t % tap_w - tap_w / 2).                    */
typedef struct {
  const void* a;      /* fp16 [n_img, h, w, a_ld] (first c channels used)                                     */
  int64_t n_img, h, w, c, a_ld;
  const void* wgt;    /* fp16 [taps][n_out][w_ld]  (w_ld >= c, multiple of 8, zero padded)                     */
  int64_t n_out, w_ld;
  int32_t taps;       /* 1, 9, or tap_h * tap_w                                                                 */
  int32_t geglu;      /* 1: weight rows are tile-interleaved [128 hidden | 128 gate]; output width n_out/2      */
  void* d;            /* fp16 (or fp32 when out_f32) [n_img*h*w, d_ld]                                         */
  int64_t d_ld;
  int32_t out_f32;
  int32_t splits;     /* split-K: >1 -> d is fp32 [splits, rows, d_ld] partial sums (no epilogue terms);        */
                      /* finish with ivv_splitk_reduce. 0/1 = off                                                */
  const void* bias;     /* fp16 [n_out] or NULL (GEGLU: same tile-interleaved order as wgt rows)               */
  const void* rowbias;  /* fp16 [groups, rowbias_ld] or NULL: added to every pixel of group pix/rowbias_group  */
  int64_t rowbias_group, rowbias_ld;
  const void* residual; /* fp16 [n_img*h*w, res_ld] or NULL                                                    */
  int64_t res_ld;
  int32_t tap_h, tap_w; /* 0, 0: taps = 1 -> 1x1, taps = 9 -> 3x3; otherwise taps must equal tap_h * tap_w (both odd) */
  int32_t relu;         /* 1: D = max(D, 0) after bias / rowbias / residual (not with GEGLU or split-K)           */
  int32_t rowbias_mod;  /* > 0: rowbias row = (pix / rowbias_group) % rowbias_mod (per-frame tables that repeat per clip) */
  /* ---- LayerNorm folded into the GEMMs around it (K8; attention.py:241-256, motion_module.py:204-217,236-242) ----
   * LN(x) W^T = rstd (x W'^T - mean * wsum) + (W beta) with W' = W diag(gamma), wsum[n] = sum_k W'[n][k]: the consumer
   * GEMM runs on the RAW rows x and applies the per-row statistics in its epilogue; the statistics come from the GEMM
   * that PRODUCED x, which writes per-row partial (sum, sum of squares) of its fp32 results. Both sides are served by
   * the short-K pair kernel only: ivv_gemm_ln_fold_ok() tells whether a shape takes it.
   * row_stats_out: fp32 [rows][n_out / 40][2] or NULL (producer side).
   * ln_stats: fp32 [rows][ln_parts][2] or NULL (consumer side), ln_wsum: fp16 [n_out]; mean / variance over c columns. */
  void* row_stats_out;
  const void* ln_stats;
  const void* ln_wsum;
  int32_t ln_parts;
  float ln_eps;
} ivv_gemm_args;
/* 1 if ivv_gemm serves a linear layer [rows, k] -> [rows, n_out] with the kernel that supports row_stats_out / ln_stats */
int ivv_gemm_ln_fold_ok(int64_t rows, int64_t k, int64_t n_out);
int ivv_gemm(const ivv_gemm_args* args, ivv_stream_t stream);

/* split-K finish: out[r, c] = sum_z partial[z][r][c] (+bias[c]) (+rowbias[r/group][c]) (+residual[r][c]) -> fp16.   */
int ivv_splitk_reduce(const float* partial, int32_t splits, int64_t rows, int64_t n, int64_t p_ld, const void* bias,
                      const void* rowbias, int64_t rowbias_group, int64_t rowbias_ld, const void* residual,
                      int64_t res_ld, void* out, int64_t out_ld, ivv_stream_t stream);

/* im2col for the stride-2 3x3 convolutions: out[n,ho,wo, tap*c + ci] = x[n, 2ho+ky-pad, 2wo+kx-pad, ci].
 * pad = 1: Downsample3D (resnet.py:99-107, symmetric pad 1);  pad = 0: the VAE encoder's Downsample
 * (vqvae/model.py:55-71: zero pad (0,1,0,1) then stride-2 conv) — the missing bottom/right taps read as zero.      */
int ivv_im2col_s2(const void* x, void* out, int64_t n_img, int64_t h, int64_t w, int64_t c, int64_t ho, int64_t wo,
                  int32_t pad, ivv_stream_t stream);

/* ---- K6/K7: GroupNorm (+SiLU) ------------------------------------------------------------------------------
 * Replaces torch.nn.GroupNorm + F.silu at resnet.py:177-178,188-194, unet.py:427-428 (5-D: statistics span
 * frames_per_group = F frames) and attention.py:101, motion_module.py:136, vqvae/model.py:31-32 (per frame: 1).
 * x,y: fp16 [n_img, hw, c]; stats workspace: ivv_groupnorm_ws_bytes() bytes, 16-byte aligned. Its first 4096 bytes (the
 * CTA ticket counters) must be ZERO the first time the buffer is used; every call returns them to zero, so one
 * zero-filled buffer serves any number of calls on a stream (no memset in front of each norm).                    */
int ivv_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int64_t n_img, int64_t hw, int64_t c,
                  int32_t groups, int64_t frames_per_group, float eps, int32_t silu, void* stats_ws,
                  size_t stats_ws_bytes, ivv_stream_t stream);
size_t ivv_groupnorm_ws_bytes(int64_t n_img, int32_t groups, int64_t frames_per_group);
/* 1 when ivv_groupnorm / ivv_groupnorm2 on this shape run as ONE kernel (the whole tensor held in shared memory between
 * statistics and normalisation, one CTA per SM), 0 for the statistics + apply pair: launch accounting of the caller. */
int ivv_groupnorm_is_fused(int64_t n_img, int64_t hw, int64_t c, int32_t groups, int64_t frames_per_group);
/* Same over the channel concatenation [x1 | x2] (x1: [n_img, hw, c1], x2: [n_img, hw, c2]) WITHOUT materialising it:
 * the `torch.cat([hidden_states, res_hidden_states], dim=1)` of the up blocks (unet_blocks.py:561,659) feeding
 * ResnetBlock3D.norm1 (resnet.py:177). y: fp16 [n_img, hw, c1 + c2].                                                 */
int ivv_groupnorm2(const void* x1, int64_t c1, const void* x2, int64_t c2, void* y, const void* gamma,
                   const void* beta, int64_t n_img, int64_t hw, int32_t groups, int64_t frames_per_group, float eps,
                   int32_t silu, void* stats_ws, size_t stats_ws_bytes, ivv_stream_t stream);

/* ---- K8: LayerNorm (+ temporal positional encoding) ---------------------------------------------------------
 * Replaces nn.LayerNorm at attention.py:168,185,191 / motion_module.py:195,201 and, when pe != NULL, the
 * PositionalEncoding add of motion_module.py:236-242 (pe fp32 [pe_len, c], frame = (row / rows_per_frame) % frames,
 * pe row = pe_start + frame).                                                                                  */
int ivv_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t rows, int64_t c, float eps,
                  const float* pe, int64_t rows_per_frame, int64_t frames, int64_t pe_start, ivv_stream_t stream);

/* ---- K3/K4: fused flash attention (tcgen05 + TMEM + TMA) ----------------------------------------------------
 * Replaces diffusers Attention core / xformers.memory_efficient_attention called at attention.py:241-256.
 * q: fp16 rows [n_batch*s_q, q_ld], head hd at columns [hd*d, hd*d+d); k, v likewise with kv batch = batch/kv_div
 * (kv_div = frames for cross-attention: the per-clip context is shared by all frames, attention.py:96).
 * o: fp16 [n_batch*s_q, o_ld].  softmax(q k^T * scale) v, fp32 statistics.                                     */
int ivv_attention(const void* q, int64_t q_ld, const void* k, const void* v, int64_t kv_ld, void* o, int64_t o_ld,
                  int64_t n_batch, int64_t s_q, int64_t s_kv, int64_t kv_div, int32_t heads, int32_t d, float scale,
                  ivv_stream_t stream);

/* ---- K5: temporal self-attention over frames (HBM-bound, L = frames <= 64) ----------------------------------
 * Replaces VersatileAttention.forward core, motion_module.py:275,303-334, without either transpose copy:
 * qkv fp16 [clips*frames*hw, 3c] (q | k | v); sequences run over the frame index at fixed (clip, pixel).        */
int ivv_temporal_attention(const void* qkv, void* o, int64_t clips, int64_t frames, int64_t hw, int64_t c,
                           int32_t heads, float scale, ivv_stream_t stream);

/* ---- row softmax for the materialised VAE attention (vqvae/model.py:183-186): x fp32 or fp16 -> y fp16 ------- */
int ivv_softmax_rows(const void* x, int32_t x_is_f32, void* y, int64_t rows, int64_t cols, float scale,
                     ivv_stream_t stream);

/* ---- K9 and layout glue -------------------------------------------------------------------------------------*/
/* nearest 2x upsample of frames (resnet.py:59-61, vqvae/model.py:48): [n,h,w,c] -> [n,2h,2w,c] (or to ho,wo)     */
int ivv_upsample_nearest(const void* x, void* y, int64_t n_img, int64_t h, int64_t w, int64_t c, int64_t ho,
                         int64_t wo, ivv_stream_t stream);
/* channel concat (unet_blocks.py:561,659): y[rows, ca+cb] = [a | b]                                            */
int ivv_concat_channels(const void* a, int64_t ca, const void* b, int64_t cb, void* y, int64_t rows,
                        ivv_stream_t stream);
/* [b, c, f, h, w] (fp32 or fp16) -> frames [b*f, h, w, c_pad] fp16 (channels >= c zero)                          */
int ivv_ncfhw_to_frames(const void* x, int32_t x_is_f32, void* y, int64_t b, int64_t c, int64_t f, int64_t hw,
                        int64_t c_pad, ivv_stream_t stream);
/* frames [b*f, hw, ld] (fp16 or fp32, first c channels) -> [b, c, f, h, w] fp32 or fp16                          */
int ivv_frames_to_ncfhw(const void* x, int32_t x_is_f32, int64_t ld, void* y, int32_t y_is_f32, int64_t b,
                        int64_t c, int64_t f, int64_t hw, ivv_stream_t stream);
/* sinusoidal timestep projection (diffusers Timesteps, called at unet.py:358): t[n] -> fp16 [n, dim] (cos | sin)  */
int ivv_timestep_embedding(const float* t, void* y, int64_t n, int32_t dim, int32_t flip_sin_to_cos,
                           float freq_shift, ivv_stream_t stream);
/* y = silu(x) elementwise, fp16                                                                                */
int ivv_silu(const void* x, void* y, int64_t n, ivv_stream_t stream);
/* y = a * x + b (fp16 in, fp16 out) — latent scaling (diffusion.py:247-249)                                      */
int ivv_scale(const void* x, void* y, int64_t n, float a, float b, ivv_stream_t stream);

/* ---- K12: optical-flow motion compensation -------------------------------------------------------------------
 * ivv_warp_image: misc_utils/flow_utils.py:25-57 (bilinear, zeros padding, align_corners=True), NCHW fp32.
 * ivv_resize_flow: flow_utils.py:59-86 (scale u,v then bilinear align_corners=False), NCHW fp32.
 * ivv_flow_noise_correction: the whole per-step loop pl_trainer/inference/inference.py:374-386 in one launch:
 *   for each query frame q: eps[q] += where(sum_r mask > 0.5, sum_r warp(delta[r], flow[q,r]) / sum_r mask, 0).   */
int ivv_warp_image(const float* image, const float* flow, float* out, int64_t n, int64_t c, int64_t h, int64_t w,
                   ivv_stream_t stream);
int ivv_resize_flow(const float* flow, float* out, int64_t n, int64_t h, int64_t w, int64_t ho, int64_t wo,
                    ivv_stream_t stream);
int ivv_flow_noise_correction(const float* delta_ref, const float* flow_lat, float* eps, int64_t q, int64_t r,
                              int64_t c, int64_t h, int64_t w, ivv_stream_t stream);

/* ---- K13: classifier-free-guidance combine + DDIM update (inference.py:198-210) ------------------------------
 * eps3: fp32 [3, n] (branches uncond | image | text+image); latent fp32 [n] updated in place.                   */
int ivv_cfg_ddim_step(const float* eps3, float* latent, float* eps_out, int64_t n, float text_cfg, float img_cfg,
                      float alpha_prod_t, float alpha_prod_prev, ivv_stream_t stream);

/* ---- frame I/O ------------------------------------------------------------------------
 * ivv_frames_u8_to_f32: decoded frames uint8 [n, h, w, 3] -> fp32 [n, 3, h, w] in [-1, 1] = cv2.cvtColor(BGR2RGB)
 * (swap_rb = 1) + ToTensor + Normalize(0.5, 0.5) of dataset/loveu_tgve_dataset.py:13-16,50-52, bit-identical.
 * ivv_frames_to_u8: frames fp32 / fp16 [n, 3, h, w] in [-1, 1] -> uint8 [n, h, w, 3] = `x / 2 + 0.5`, `* 255`,
 * astype(uint8) of misc_utils/image_utils.py:130,233-241 (clamped to [0, 255]).                                       */
int ivv_frames_u8_to_f32(const void* x, float* y, int64_t n, int64_t hw, int32_t swap_rb, ivv_stream_t stream);
int ivv_frames_to_u8(const void* x, int32_t x_is_f32, void* y, int64_t n, int64_t hw, ivv_stream_t stream);

/* ---- K13b: a whole sampling step around the UNet, graph-capturable without per-step host arguments ------------------
 * The reference's loops (pl_trainer/inference/inference.py:163-219, 221-289, 313-398) do, per step: assemble the
 * 3-branch UNet input, call the UNet, CFG-combine (:198-203), optional rescale_noise_cfg (:13-24, 205-206), optional
 * reference-frame noise correction (mean :262-277 / optical flow :367-386), scheduler.step (diffusers 0.21.4 DDIM
 * eta=0 or DDPM fixed_small). Here the host folds each step's scheduler scalars into one row of `table`
 * (IVV_SAMPLER_ROW floats: t, sqrt(a_t), sqrt(1-a_t), c_x0, c_xt, c_eps, sigma, correct?, text_cfg, img_cfg,
 * guidance_rescale, noise_row: prev = c_x0*x0 + c_xt*x_t + c_eps*eps + sigma*noise[noise_row]) once per clip; the
 * kernels read the row selected by state[0], which ivv_sampler_update advances, so begin -> UNet -> combine -> update
 * captured once is every step.  state: int32[4], zeroed by the caller before the first step.
 * lat2: fp32 [2][F][C][hw] ping-pong latent (step s reads half s&1, writes half (s+1)&1); cond, eps_cfg: [F][C][hw].  */
#define IVV_SAMPLER_ROW 16
/* x_frames fp16 [3*F*hw, c_pad] = [latent | 0 or cond] for the branches (uncond, image, text+image); t_out[0..2] = t   */
int ivv_sampler_begin(const float* table, const int32_t* state, const float* lat2, const float* cond, void* x_frames,
                      float* t_out, int64_t frames, int64_t c, int64_t hw, int64_t c_pad, ivv_stream_t stream);
/* eps3: UNet output frames fp32 [3*F*hw, eps_ld]; partials: double [ivv_sampler_partials(F, hw)][4]                   */
int64_t ivv_sampler_partials(int64_t frames, int64_t hw);
int ivv_sampler_combine(const float* table, const int32_t* state, const float* eps3, int64_t eps_ld, float* eps_cfg,
                        double* partials, int64_t frames, int64_t c, int64_t hw, ivv_stream_t stream);
/* mode 0: plain step; 1: mean correction from latent_ref [R][C][hw]; 2: flow correction, flows_lat [Q][R][2][hw] at
 * latent resolution (query frame R+q, q < Q). noise: fp32 [rows][F*C*hw] or NULL (DDPM variance noise, drawn by the
 * caller); hist_lat / hist_pred: fp32 [n_steps][F*C*hw] or NULL (the reference's all_latent / all_pred lists).        */
int ivv_sampler_update(const float* table, int32_t* state, float* lat2, const float* eps_cfg, const double* partials,
                       int32_t mode, const float* latent_ref, const float* flows_lat, const float* noise,
                       float* hist_lat, float* hist_pred, int64_t frames, int64_t c, int64_t r, int64_t q, int64_t h,
                       int64_t w, ivv_stream_t stream);

/* ---- RAFT optical flow -------------------------------------------------------------------
 * The reference's RAFTFlow (misc_utils/flow_utils.py:134-189) wraps torchvision.models.optical_flow.raft_large
 * (torchvision 0.26, models/optical_flow/raft.py; file:line below refer to it). Convolutions run on ivv_gemm; the
 * entry points here are the remaining pieces, all channels-last.                                                  */
/* generic im2col for the strided / 7x7 convolutions (raft.py:136-138 stem, :191 convflow1, stride-2 blocks):
 * out[(n,oy,ox), (ky*kw+kx)*c + ci] = x[n, oy*stride+ky-pad_h, ox*stride+kx-pad_w, ci], zero outside; c % 8 == 0.   */
int ivv_im2col(const void* x, void* out, int64_t n_img, int64_t h, int64_t w, int64_t c, int32_t kh, int32_t kw,
               int32_t stride, int32_t pad_h, int32_t pad_w, int64_t ho, int64_t wo, ivv_stream_t stream);
/* InstanceNorm2d (imgs_per_group = 1) / BatchNorm2d with BATCH statistics (imgs_per_group = n_img: the reference
 * never calls .eval() on RAFTFlow, inference.py:294) of the encoders' Conv2dNormActivation (raft.py:38-59), biased
 * variance, optional affine (gamma/beta may be NULL), optional ReLU, optional residual join
 * y = relu(residual + y) (ResidualBlock.forward, raft.py:63-71). x, y, residual: fp16 [n_img, hw, c].             */
int ivv_channelnorm(const void* x, void* y, const void* gamma, const void* beta, int64_t n_img, int64_t hw, int64_t c,
                    int64_t imgs_per_group, float eps, int32_t relu, const void* residual, void* stats_ws,
                    size_t stats_ws_bytes, ivv_stream_t stream);
size_t ivv_channelnorm_ws_bytes(int64_t n_img, int64_t c, int64_t imgs_per_group);
/* y = relu(a + b), fp16, n % 8 == 0: the ResidualBlock join when BatchNorm is folded (eval mode)                    */
int ivv_add_relu(const void* a, const void* b, void* y, int64_t n, ivv_stream_t stream);
/* images fp32 [n, 3, hs, ws] -> fp16 [n, h, w, 8] (channels 3..7 zero): TF.resize(antialias=False) when the size
 * differs (flow_utils.py:180-182) and the OpticalFlow preset's (v - 0.5) / 0.5 (flow_utils.py:184).                */
int ivv_raft_prep_images(const float* img, void* out, int64_t n, int64_t hs, int64_t ws, int64_t h, int64_t w,
                         ivv_stream_t stream);
/* F.avg_pool2d(2, 2) over the last two dims of fp32 [n, h, w] -> [n, h/2, w/2] (CorrBlock.build_pyramid :388-390)  */
int ivv_avgpool2_f32(const float* x, float* y, int64_t n, int64_t h, int64_t w, ivv_stream_t stream);
/* CorrBlock.index_pyramid (raft.py:393-421). pyramid[l]: fp32 [n_pairs*h*w, h>>l, w>>l] (un-normalised dot
 * products; scale = 1/sqrt(channels) is applied here), coords fp32 [n_pairs*h*w, 2] (x, y).
 * out fp16 [rows, out_ld], channel l*(2r+1)^2 + i*(2r+1) + j = corr_l(x/2^l + i - r, y/2^l + j - r).              */
int ivv_corr_lookup(const float* const* pyramid, int32_t levels, const float* coords, void* out, int64_t out_ld,
                    int64_t n_pairs, int64_t h, int64_t w, int32_t radius, float scale, ivv_stream_t stream);
/* hidden = tanh(ctx[:, :hidden]) (fp32 master h32 + fp16 into hx[:, :hidden]); hx[:, hidden:hidden+context] =
 * relu(ctx[:, hidden:]) (raft.py:512-514)                                                                        */
int ivv_raft_init_state(const void* ctx, int64_t ctx_ld, float* h32, void* hx, int64_t hx_ld, int64_t rows,
                        int32_t hidden, int32_t context, ivv_stream_t stream);
/* ConvGRU gates (raft.py:222-229). zrq fp16 [rows, zrq_ld] = [z_pre | r_pre | ...] pre-activations:
 * ivv_gru_gate_r: rh = sigmoid(r_pre) * h;  ivv_gru_update: h = (1 - z) h + z tanh(q_pre), z = sigmoid(z_pre).    */
int ivv_gru_gate_r(const void* zrq, int64_t zrq_ld, const float* h32, void* rh, int64_t rows, int32_t hidden,
                   ivv_stream_t stream);
int ivv_gru_update(const void* zrq, int64_t zrq_ld, const void* q_pre, float* h32, void* hx, int64_t hx_ld,
                   int64_t rows, int32_t hidden, ivv_stream_t stream);
/* coords1 += delta[:, 0:2] (delta may be NULL) (raft.py:527); flow = coords1 - grid as fp16 into flow8 [rows, 8]
 * (input of the motion encoder's 7x7 conv) and, if flow_slot != NULL, into flow_slot[row*ld + 0..1] (raft.py:211).  */
int ivv_raft_update_coords(const float* delta, int64_t delta_ld, float* coords1, void* flow8, void* flow_slot,
                           int64_t flow_slot_ld, int64_t n_pairs, int64_t h, int64_t w, ivv_stream_t stream);
/* convex 8x upsampling (torchvision _utils.upsample_flow): mask fp16 [rows, mask_ld >= 576] -> fp32 [n, 2, 8h, 8w]  */
int ivv_convex_upsample(const void* mask, int64_t mask_ld, const float* coords1, float* out, int64_t n_pairs,
                        int64_t h, int64_t w, ivv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IVV_H_ */
