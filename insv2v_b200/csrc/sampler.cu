// The denoising step around the UNet as three small kernels, so that ONE CUDA graph (begin -> UNet3D -> combine ->
// update) is a whole step of the reference's sampling loops and the loop itself is N graph launches with nothing in
// between. Reference: pl_trainer/inference/inference.py:163-219 (__call__), :221-289 (second_clip_forward, mean
// correction), :313-398 (optical-flow correction), :13-24 (rescale_noise_cfg); scheduler arithmetic: diffusers 0.21.4
// DDIMScheduler.step (eta 0) / DDPMScheduler.step (fixed_small), folded by the host into one row of scalars per step.
//
// State lives in caller-owned device buffers:
//   table  fp32 [n_steps][IVV_SAMPLER_ROW]   one row per step (layout below), written once per clip by the host
//   state  int32 [4]                          state[0] = index of the next row; advanced by the update kernel
//   lat2   fp32 [2][F][C][hw]                 ping-pong latent: step s reads half (s & 1), writes half ((s + 1) & 1)
// Every kernel reads the step index from `state`, so the captured graph is identical for every step.
#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

// row layout (floats): must match insv2v_b200/pipeline.py::sampler_table
enum { R_T = 0, R_SQRT_AT, R_SQRT_1MAT, R_C_X0, R_C_XT, R_C_EPS, R_SIGMA, R_CORRECT, R_TEXT_CFG, R_IMG_CFG, R_RESCALE,
       R_NOISE_ROW };

struct Bilin2 {
  int x0, y0;
  float w00, w01, w10, w11;
};

// identical to warp.cu::bilinear_zeros (flow_utils.py:43-52 + grid_sample align_corners=True, zeros padding)
__device__ __forceinline__ Bilin2 bilin_zeros(float px, float py, int w, int h) {
  const float gx = 2.f * (px / (float)(w - 1) - 0.5f);
  const float gy = 2.f * (py / (float)(h - 1) - 0.5f);
  const float ix = ((gx + 1.f) / 2.f) * (float)(w - 1);
  const float iy = ((gy + 1.f) / 2.f) * (float)(h - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  Bilin2 b;
  b.x0 = (int)fx;
  b.y0 = (int)fy;
  const float tx = ix - fx, ty = iy - fy;
  const bool x0in = b.x0 >= 0 && b.x0 < w, x1in = b.x0 + 1 >= 0 && b.x0 + 1 < w;
  const bool y0in = b.y0 >= 0 && b.y0 < h, y1in = b.y0 + 1 >= 0 && b.y0 + 1 < h;
  b.w00 = (x0in && y0in) ? (1.f - tx) * (1.f - ty) : 0.f;
  b.w01 = (x1in && y0in) ? tx * (1.f - ty) : 0.f;
  b.w10 = (x0in && y1in) ? (1.f - tx) * ty : 0.f;
  b.w11 = (x1in && y1in) ? tx * ty : 0.f;
  return b;
}

// ---- begin: UNet input frames [3*F*hw, c_pad] fp16 = [latent | (branch 0: zeros, branches 1,2: condition)] ---------
// (inference.py:183-189: latent1 = [latent, 0], latent2 = latent3 = [latent, img_cond]; 'b f c h w -> b c f h w' is a
// no-op here because the UNet works on channels-last frames), and t_out[0..2] = timestep of this step.
__global__ void sampler_begin_kernel(const float* __restrict__ table, const int* __restrict__ state,
                                     const float* __restrict__ lat2, const float* __restrict__ cond,
                                     __half* __restrict__ x, float* __restrict__ t_out, int F, int C, long long hw,
                                     int c_pad) {
  griddep_sync();
  const int step = state[0];
  const float* lat = lat2 + (long long)(step & 1) * F * C * hw;
  if (blockIdx.x == 0 && threadIdx.x < 3) t_out[threadIdx.x] = table[(long long)step * IVV_SAMPLER_ROW + R_T];
  const long long rows = 3LL * F * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / (F * hw));
    const long long fp = i % (F * hw);
    const int f = (int)(fp / hw);
    const long long p = fp % hw;
    __half* dst = x + i * c_pad;
    for (int c = 0; c < c_pad; ++c) {
      float v = 0.f;
      if (c < C) v = lat[((long long)f * C + c) * hw + p];
      else if (c < 2 * C && b > 0) v = cond[((long long)f * C + (c - C)) * hw + p];
      dst[c] = __float2half_rn(v);
    }
  }
}

// ---- combine: eps_cfg = e1 + img_cfg (e2 - e1) + text_cfg (e3 - e2) (inference.py:198-203) ------------------------
// eps3: UNet output frames fp32 [3*F*hw, eps_ld] (first C channels). eps_cfg: fp32 [F][C][hw]. Per-CTA partial sums
// (sum, sum of squares of eps_cfg and of branch 1) for rescale_noise_cfg go to partials [gridDim.x][4] (double).
__global__ void sampler_combine_kernel(const float* __restrict__ table, const int* __restrict__ state,
                                       const float* __restrict__ eps3, long long eps_ld, float* __restrict__ eps_cfg,
                                       double* __restrict__ partials, int F, int C, long long hw) {
  griddep_sync();
  const int step = state[0];
  const float* row = table + (long long)step * IVV_SAMPLER_ROW;
  const float text_cfg = row[R_TEXT_CFG], img_cfg = row[R_IMG_CFG];
  const long long fhw = (long long)F * hw;
  double s_c = 0.0, q_c = 0.0, s_1 = 0.0, q_1 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < fhw; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / hw);
    const long long p = i % hw;
    for (int c = 0; c < C; ++c) {
      const float e1 = eps3[i * eps_ld + c], e2 = eps3[(fhw + i) * eps_ld + c], e3 = eps3[(2 * fhw + i) * eps_ld + c];
      const float e = e1 + img_cfg * (e2 - e1) + text_cfg * (e3 - e2);
      eps_cfg[((long long)f * C + c) * hw + p] = e;
      s_c += e, q_c += (double)e * e, s_1 += e1, q_1 += (double)e1 * e1;
    }
  }
  __shared__ double red[4][32];
  double v[4] = {s_c, q_c, s_1, q_1};
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) red[k][warp] = v[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) t += red[threadIdx.x][wi];  // fixed order: deterministic
    partials[(long long)blockIdx.x * 4 + threadIdx.x] = t;
  }
}

// ---- update: (rescale) -> reference-frame noise correction -> scheduler step ----------------------------------------
// mode 0: no correction; 1: mean over the R reference frames (inference.py:270-277); 2: optical-flow warp
// (inference.py:367-386; flows_lat [Q][R][2][hw] already at latent resolution). The correction is applied only on steps
// whose table row has correct != 0. delta(r, c, p) = noise_ref - eps is recomputed where it is needed from the
// read-only half of lat2, so nothing written by this kernel is read by it.
template <int MODE>
__global__ void sampler_update_kernel(const float* __restrict__ table, int* __restrict__ state,
                                      float* __restrict__ lat2, const float* __restrict__ eps_cfg,
                                      const double* __restrict__ partials, int n_partials,
                                      const float* __restrict__ latent_ref, const float* __restrict__ flows_lat,
                                      const float* __restrict__ noise, float* __restrict__ hist_lat,
                                      float* __restrict__ hist_pred, int F, int C, int R, int Q, int h, int w) {
  griddep_sync();
  const long long hw = (long long)h * w;
  const long long n = (long long)F * C * hw;
  const int step = state[0];
  const float* row = table + (long long)step * IVV_SAMPLER_ROW;
  const float sa = row[R_SQRT_AT], sb = row[R_SQRT_1MAT];
  const float c_x0 = row[R_C_X0], c_xt = row[R_C_XT], c_eps = row[R_C_EPS], sigma = row[R_SIGMA];
  const bool correct = MODE != 0 && row[R_CORRECT] != 0.f;
  const float rescale = row[R_RESCALE];
  const float* src = lat2 + (long long)(step & 1) * n;
  float* dst = lat2 + (long long)((step + 1) & 1) * n;
  const float* nz = (sigma != 0.f && noise != nullptr) ? noise + (long long)row[R_NOISE_ROW] * n : nullptr;

  // rescale_noise_cfg (inference.py:13-24): eps' = eps * (g * std_1 / std_cfg + 1 - g); unbiased std over F*C*hw
  __shared__ float mix_s;
  if (threadIdx.x == 0) {
    float mix = 1.f;
    if (rescale > 0.f) {
      double s_c = 0.0, q_c = 0.0, s_1 = 0.0, q_1 = 0.0;
      for (int i = 0; i < n_partials; ++i) {  // fixed order, every CTA the same result
        s_c += partials[i * 4 + 0], q_c += partials[i * 4 + 1], s_1 += partials[i * 4 + 2], q_1 += partials[i * 4 + 3];
      }
      const double dn = (double)n;
      const double var_c = (q_c - s_c * s_c / dn) / (dn - 1.0), var_1 = (q_1 - s_1 * s_1 / dn) / (dn - 1.0);
      const float std_c = (float)sqrt(var_c > 0.0 ? var_c : 0.0), std_1 = (float)sqrt(var_1 > 0.0 ? var_1 : 0.0);
      mix = rescale * (std_1 / std_c) + (1.f - rescale);
    }
    mix_s = mix;
  }
  __syncthreads();
  const float mix = mix_s;
  const float* lref = latent_ref;

  auto delta_at = [&](int r, int c, long long p) -> float {  // noise_ref - eps on reference frame r
    const long long o = ((long long)r * C + c) * hw + p;
    const float e = eps_cfg[o] * mix;
    const float nr = (src[o] - sa * lref[o]) / sb;
    return nr - e;
  };

  const long long fhw = (long long)F * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < fhw; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / hw);
    const long long p = i % hw;
    float add[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) add[c] = 0.f;
    if (correct) {
      if (f < R) {
        for (int c = 0; c < C; ++c) add[c] = delta_at(f, c, p);
      } else if (MODE == 1) {
        for (int c = 0; c < C; ++c) {
          float s = 0.f;
          for (int r = 0; r < R; ++r) s += delta_at(r, c, p);
          add[c] = s / (float)R;  // delta_noise_ref.mean(dim=1)
        }
      } else if (MODE == 2 && f - R < Q) {  // zip(range(R, F), warp_funcs): only the first Q query frames
        const int q = f - R;
        const int y = (int)(p / w), x = (int)(p % w);
        float msum = 0.f;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        for (int r = 0; r < R; ++r) {
          const float* fl = flows_lat + ((long long)(q * R + r) * 2) * hw;
          const Bilin2 b = bilin_zeros((float)x + fl[p], (float)y + fl[hw + p], w, h);
          msum += b.w00 + b.w01 + b.w10 + b.w11;
          for (int c = 0; c < C; ++c) {
            float v = 0.f;
            const long long p00 = (long long)b.y0 * w + b.x0;
            if (b.w00 != 0.f) v += b.w00 * delta_at(r, c, p00);
            if (b.w01 != 0.f) v += b.w01 * delta_at(r, c, p00 + 1);
            if (b.w10 != 0.f) v += b.w10 * delta_at(r, c, p00 + w);
            if (b.w11 != 0.f) v += b.w11 * delta_at(r, c, p00 + w + 1);
            acc[c] += v;
          }
        }
        if (msum > 0.5f)
          for (int c = 0; c < C; ++c) add[c] = acc[c] / msum;
      }
    }
    for (int c = 0; c < C; ++c) {
      const long long o = ((long long)f * C + c) * hw + p;
      const float e = eps_cfg[o] * mix + add[c];
      const float xt = src[o];
      const float x0 = (xt - sb * e) / sa;
      float prev = c_x0 * x0 + c_xt * xt + c_eps * e;
      if (nz) prev += sigma * nz[o];
      dst[o] = prev;
      if (hist_lat) hist_lat[(long long)step * n + o] = prev;
      if (hist_pred) hist_pred[(long long)step * n + o] = x0;
    }
  }
  // advance the step index once every CTA has read it: last CTA to finish (ticket in state[1], self-resetting)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicInc(reinterpret_cast<unsigned*>(state + 1), gridDim.x - 1);
    if (done == gridDim.x - 1) state[0] = step + 1;
  }
}

static inline unsigned sgrid(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  if (b > 148LL * 4) b = 148LL * 4;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace ivv

using namespace ivv;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" int ivv_sampler_begin(const float* table, const int32_t* state, const float* lat2, const float* cond,
                                 void* x_frames, float* t_out, int64_t frames, int64_t c, int64_t hw, int64_t c_pad,
                                 ivv_stream_t stream_) {
  IVV_REQUIRE(table && state && lat2 && cond && x_frames && t_out, "ivv_sampler_begin: null pointer");
  IVV_REQUIRE(frames > 0 && c > 0 && hw > 0 && c_pad >= 2 * c, "ivv_sampler_begin: bad shape (c_pad %lld < 2*c %lld?)",
              (long long)c_pad, (long long)(2 * c));
  sampler_begin_kernel<<<sgrid(3 * frames * hw, 256), 256, 0, STREAM>>>(table, state, lat2, cond,
                                                                      reinterpret_cast<__half*>(x_frames), t_out,
                                                                      (int)frames, (int)c, hw, (int)c_pad);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t ivv_sampler_partials(int64_t frames, int64_t hw) { return (int64_t)sgrid(frames * hw, 256); }

extern "C" int ivv_sampler_combine(const float* table, const int32_t* state, const float* eps3, int64_t eps_ld,
                                   float* eps_cfg, double* partials, int64_t frames, int64_t c, int64_t hw,
                                   ivv_stream_t stream_) {
  IVV_REQUIRE(table && state && eps3 && eps_cfg && partials, "ivv_sampler_combine: null pointer");
  IVV_REQUIRE(frames > 0 && c > 0 && c <= 8 && hw > 0 && eps_ld >= c, "ivv_sampler_combine: bad shape");
  sampler_combine_kernel<<<sgrid(frames * hw, 256), 256, 0, STREAM>>>(table, state, eps3, eps_ld, eps_cfg, partials,
                                                                    (int)frames, (int)c, hw);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_sampler_update(const float* table, int32_t* state, float* lat2, const float* eps_cfg,
                                  const double* partials, int32_t mode, const float* latent_ref,
                                  const float* flows_lat, const float* noise, float* hist_lat, float* hist_pred,
                                  int64_t frames, int64_t c, int64_t r, int64_t q, int64_t h, int64_t w,
                                  ivv_stream_t stream_) {
  IVV_REQUIRE(table && state && lat2 && eps_cfg && partials, "ivv_sampler_update: null pointer");
  IVV_REQUIRE(frames > 0 && c > 0 && c <= 8 && h > 0 && w > 0, "ivv_sampler_update: bad shape");
  IVV_REQUIRE(mode >= 0 && mode <= 2, "ivv_sampler_update: mode %d not in {0,1,2}", mode);
  IVV_REQUIRE(mode == 0 || (latent_ref && r > 0 && r < frames), "ivv_sampler_update: correction needs 0 < R < frames");
  IVV_REQUIRE(mode != 2 || (flows_lat && q > 0 && q <= frames - r), "ivv_sampler_update: mode 2 needs 0 < Q <= F - R flows");
  const long long hw = h * w;
  const unsigned grid = sgrid(frames * hw, 128);
  const int np = (int)sgrid(frames * hw, 256);
#define LAUNCH(M)                                                                                                       \
  sampler_update_kernel<M><<<grid, 128, 0, STREAM>>>(table, state, lat2, eps_cfg, partials, np, latent_ref, flows_lat,  \
                                                     noise, hist_lat, hist_pred, (int)frames, (int)c, (int)r, (int)q,   \
                                                     (int)h, (int)w)
  if (mode == 0) LAUNCH(0);
  else if (mode == 1) LAUNCH(1);
  else LAUNCH(2);
#undef LAUNCH
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}
