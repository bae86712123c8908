// Data-movement and elementwise glue kernels (HBM-bound, 16-byte vectors where the layout allows).
// Reference call sites: ivv.h (K9, K13 and layout glue).
#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

// The three copy kernels below run one CTA per output pixel row (n, oy) and walk it with walk_rows(): a thread keeps its
// 16-byte column and strides over the pixels of the row, so the only per-vector work is the copy itself (the flat
// i % V, i / V ... decomposition cost eight 64-bit divisions per vector and ran at 1.1 TB/s).

// out[n, ho, wo, tap*c + ci] = x[n, 2*ho + ky - pad, 2*wo + kx - pad, ci]  (zero outside), tap = ky*3+kx
__global__ void im2col_s2_kernel(const __half* __restrict__ x, __half* __restrict__ out, long long n_img, int h, int w,
                                 int c, int ho, int wo, int pad) {
  griddep_sync();
  const int V = c / 8;
  const long long n = blockIdx.x / ho;
  const int oy = blockIdx.x - (int)n * ho;
  const __half* ximg = x + n * h * w * c;
  __half* orow = out + (long long)blockIdx.x * wo * 9 * c;
  walk_rows(9 * V, wo, 9LL * c, 0, [&](int ox, int j, long long g, int) {
    const int tap = j / V;  // loop-invariant per column
    const int vec = j - tap * V;
    const int ky = tap / 3;
    const int iy = 2 * oy + ky - pad;
    const int ix = 2 * ox + (tap - 3 * ky) - pad;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w)
      v = *reinterpret_cast<const uint4*>(ximg + ((long long)iy * w + ix) * c + vec * 8);
    *reinterpret_cast<uint4*>(orow + g) = v;
  });
}

__global__ void upsample_nearest_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long n_img, int h,
                                        int w, int c, int ho, int wo) {
  griddep_sync();
  const long long n = blockIdx.x / ho;
  const int oy = blockIdx.x - (int)n * ho;
  // PyTorch 'nearest': src = floor(dst * in / out)
  const int iy = min((int)(((long long)oy * h) / ho), h - 1);
  const __half* xrow = x + ((n * h + iy) * w) * c;
  __half* yrow = y + (long long)blockIdx.x * wo * c;
  const bool twice = wo == 2 * w;
  walk_rows(c / 8, wo, c, 0, [&](int ox, int j, long long g, int) {
    const int ix = twice ? (ox >> 1) : min((int)(((long long)ox * w) / wo), w - 1);
    *reinterpret_cast<uint4*>(yrow + g) = *reinterpret_cast<const uint4*>(xrow + (long long)ix * c + j * 8);
  });
}

constexpr int kConcatRows = 32;  // rows per CTA
__global__ void concat_channels_kernel(const __half* __restrict__ a, int ca, const __half* __restrict__ b, int cb,
                                       __half* __restrict__ y, long long rows) {
  griddep_sync();
  const int Va = ca / 8, V = (ca + cb) / 8;
  const long long r0 = (long long)blockIdx.x * kConcatRows;
  const int nr = (int)min((long long)kConcatRows, rows - r0);
  const __half* a0 = a + r0 * ca;
  const __half* b0 = b + r0 * cb;
  __half* y0 = y + r0 * (ca + cb);
  walk_rows(V, nr, ca + cb, 0, [&](int r, int j, long long g, int) {
    const uint4 v = j < Va ? *reinterpret_cast<const uint4*>(a0 + (long long)r * ca + j * 8)
                           : *reinterpret_cast<const uint4*>(b0 + (long long)r * cb + (j - Va) * 8);
    *reinterpret_cast<uint4*>(y0 + g) = v;
  });
}

template <typename TIn>
__global__ void ncfhw_to_frames_kernel(const TIn* __restrict__ x, __half* __restrict__ y, long long b, int c, int f,
                                       long long hw, int c_pad) {
  const long long total = b * f * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const long long bf = i / hw;
    const int fi = (int)(bf % f);
    const long long bi = bf / f;
    __half* yo = y + i * c_pad;
    for (int ch = 0; ch < c_pad; ++ch) {
      float v = 0.f;
      if (ch < c) v = (float)x[((bi * c + ch) * f + fi) * hw + pix];
      yo[ch] = __float2half_rn(v);
    }
  }
}

template <typename TIn, typename TOut>
__global__ void frames_to_ncfhw_kernel(const TIn* __restrict__ x, long long ld, TOut* __restrict__ y, long long b,
                                       int c, int f, long long hw) {
  const long long total = b * f * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const long long bf = i / hw;
    const int fi = (int)(bf % f);
    const long long bi = bf / f;
    const TIn* xi = x + i * ld;
    for (int ch = 0; ch < c; ++ch) y[((bi * c + ch) * f + fi) * hw + pix] = (TOut)(float)xi[ch];
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, __half* __restrict__ y, int n, int dim,
                                          int flip, float freq_shift) {
  const int half_dim = dim / 2;
  const int total = n * half_dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % half_dim;
    const int r = i / half_dim;
    const float freq = expf(-logf(10000.f) * (float)k / ((float)half_dim - freq_shift));
    const float arg = t[r] * freq;
    const float s = sinf(arg), co = cosf(arg);
    __half* yr = y + (long long)r * dim;
    if (flip) {
      yr[k] = __float2half_rn(co);
      yr[half_dim + k] = __float2half_rn(s);
    } else {
      yr[k] = __float2half_rn(s);
      yr[half_dim + k] = __float2half_rn(co);
    }
  }
}

__global__ void silu_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(silu_f(__half2float(x[i])));
}

__global__ void scale_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long n, float a, float b) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(__half2float(x[i]) * a + b);
}

// eps = e1 + img_cfg (e2 - e1) + text_cfg (e3 - e2);  DDIM (eta = 0, epsilon prediction, no clipping)
__global__ void cfg_ddim_kernel(const float* __restrict__ eps3, float* __restrict__ latent, float* __restrict__ eps_out,
                                long long n, float text_cfg, float img_cfg, float sqrt_at, float sqrt_1mat,
                                float sqrt_ap, float sqrt_1map) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float e1 = eps3[i], e2 = eps3[n + i], e3 = eps3[2 * n + i];
    const float e = e1 + img_cfg * (e2 - e1) + text_cfg * (e3 - e2);
    const float x = latent[i];
    const float x0 = (x - sqrt_1mat * e) / sqrt_at;
    latent[i] = sqrt_ap * x0 + sqrt_1map * e;
    if (eps_out) eps_out[i] = e;
  }
}

// split-K finish: 8 columns per thread, fp32 partial planes summed in a fixed order (deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long rows, int n, long long p_ld,
                                     const __half* __restrict__ bias, const __half* __restrict__ rowbias,
                                     long long rowbias_group, long long rowbias_ld, const __half* __restrict__ residual,
                                     long long res_ld, __half* __restrict__ out, long long out_ld) {
  griddep_sync();
  const int V = n / 8;
  const long long total = rows * V;
  const long long plane = rows * p_ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int vec = (int)(i % V);
    const long long r = i / V;
    const int c = vec * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int z = 0; z < splits; ++z) {
      const float4 a0 = *reinterpret_cast<const float4*>(partial + z * plane + r * p_ld + c);
      const float4 a1 = *reinterpret_cast<const float4*>(partial + z * plane + r * p_ld + c + 4);
      acc[0] += a0.x, acc[1] += a0.y, acc[2] += a0.z, acc[3] += a0.w;
      acc[4] += a1.x, acc[5] += a1.y, acc[6] += a1.z, acc[7] += a1.w;
    }
    auto add8 = [&](const __half* src) {
      const uint4 u = *reinterpret_cast<const uint4*>(src);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    };
    if (bias) add8(bias + c);
    if (rowbias) add8(rowbias + (r / rowbias_group) * rowbias_ld + c);
    if (residual) add8(residual + r * res_ld + c);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + r * out_ld + c) = o;
  }
}

// ---- frame I/O -------------------------------------------------------------------------
// decoded video frames, uint8 [n, h, w, 3] (cv2: BGR) -> fp32 [n, 3, h, w] in [-1, 1]: cv2.cvtColor(BGR2RGB) +
// transforms.ToTensor() (x / 255) + Normalize(0.5, 0.5) ((t - 0.5) / 0.5) of dataset/loveu_tgve_dataset.py:13-16,50-52,
// with the same operation order (two IEEE divisions, one subtraction), so the result is bit-identical.
__global__ void u8hwc_to_f32chw_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, long long n, long long hw,
                                       int swap_rb) {
  const long long total = n * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw, pix = i % hw;
    const uint8_t* src = x + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = __fdiv_rn((float)src[swap_rb ? 2 - c : c], 255.f);
      y[(img * 3 + c) * hw + pix] = __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
    }
  }
}

// edited frames fp32 / fp16 [n, 3, h, w] in [-1, 1] -> uint8 [n, h, w, 3]: `x / 2 + 0.5`, `* 255`, astype(uint8)
// (misc_utils/image_utils.py:130,233-235: float32 arithmetic in that order, truncation; values outside [0, 255] are
// clamped here, where numpy's cast would wrap - callers clip to [-1, 1] first, insv2v_run_loveu_tgve.py:165)
template <typename T>
__global__ void chw_to_u8hwc_kernel(const T* __restrict__ x, uint8_t* __restrict__ y, long long n, long long hw) {
  const long long total = n * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw, pix = i % hw;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = (float)x[(img * 3 + c) * hw + pix];
      const float u = __fmul_rn(__fadd_rn(__fmul_rn(v, 0.5f), 0.5f), 255.f);
      y[i * 3 + c] = (uint8_t)fminf(fmaxf(u, 0.f), 255.f);
    }
  }
}

static inline unsigned grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace ivv

using namespace ivv;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" int ivv_splitk_reduce(const float* partial, int32_t splits, int64_t rows, int64_t n, int64_t p_ld,
                                 const void* bias, const void* rowbias, int64_t rowbias_group, int64_t rowbias_ld,
                                 const void* residual, int64_t res_ld, void* out, int64_t out_ld,
                                 ivv_stream_t stream_) {
  IVV_REQUIRE(partial && out && splits >= 1 && rows > 0 && n > 0, "ivv_splitk_reduce: bad arguments");
  IVV_REQUIRE(n % 8 == 0 && p_ld % 4 == 0 && out_ld % 8 == 0, "ivv_splitk_reduce: n, out_ld must be multiples of 8");
  IVV_REQUIRE(!rowbias || (rowbias_group > 0 && rowbias_ld % 8 == 0), "ivv_splitk_reduce: bad rowbias geometry");
  IVV_REQUIRE(!residual || res_ld % 8 == 0, "ivv_splitk_reduce: res_ld must be a multiple of 8");
  const long long total = rows * (n / 8);
  IVV_CHECK_CUDA(launch_pdl(splitk_reduce_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM, partial, (int)splits,
                            rows, (int)n, p_ld, reinterpret_cast<const __half*>(bias),
                            reinterpret_cast<const __half*>(rowbias), rowbias_group > 0 ? rowbias_group : 1, rowbias_ld,
                            reinterpret_cast<const __half*>(residual), res_ld, reinterpret_cast<__half*>(out), out_ld));
  return 0;
}

extern "C" int ivv_im2col_s2(const void* x, void* out, int64_t n_img, int64_t h, int64_t w, int64_t c, int64_t ho,
                             int64_t wo, int32_t pad, ivv_stream_t stream_) {
  IVV_REQUIRE(x && out && n_img > 0 && c % 8 == 0, "ivv_im2col_s2: bad arguments (c must be a multiple of 8)");
  IVV_REQUIRE(pad == 0 || pad == 1, "ivv_im2col_s2: pad must be 0 (VAE Downsample) or 1 (Downsample3D)");
  IVV_REQUIRE(ho == (h - 2 + pad) / 2 + 1 && wo == (w - 2 + pad) / 2 + 1,
              "ivv_im2col_s2: output size must be (h-2+pad)/2+1");
  IVV_REQUIRE(n_img * ho < (1LL << 31) && h * w * c < (1LL << 40), "ivv_im2col_s2: tensor too large");
  IVV_CHECK_CUDA(launch_pdl(im2col_s2_kernel, dim3((unsigned)(n_img * ho)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(out), n_img, (int)h, (int)w,
                            (int)c, (int)ho, (int)wo, (int)pad));
  return 0;
}

extern "C" int ivv_upsample_nearest(const void* x, void* y, int64_t n_img, int64_t h, int64_t w, int64_t c, int64_t ho,
                                    int64_t wo, ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && n_img > 0 && c % 8 == 0, "ivv_upsample_nearest: bad arguments (c must be a multiple of 8)");
  IVV_REQUIRE(h > 0 && w > 0 && ho > 0 && wo > 0 && n_img * ho < (1LL << 31), "ivv_upsample_nearest: bad extents");
  IVV_CHECK_CUDA(launch_pdl(upsample_nearest_kernel, dim3((unsigned)(n_img * ho)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), n_img, (int)h, (int)w,
                            (int)c, (int)ho, (int)wo));
  return 0;
}

extern "C" int ivv_concat_channels(const void* a, int64_t ca, const void* b, int64_t cb, void* y, int64_t rows,
                                   ivv_stream_t stream_) {
  IVV_REQUIRE(a && b && y && rows > 0 && ca % 8 == 0 && cb % 8 == 0, "ivv_concat_channels: bad arguments");
  const long long ctas = (rows + ivv::kConcatRows - 1) / ivv::kConcatRows;
  IVV_REQUIRE(ctas < (1LL << 31), "ivv_concat_channels: too many rows");
  IVV_CHECK_CUDA(launch_pdl(concat_channels_kernel, dim3((unsigned)ctas), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(a), (int)ca, reinterpret_cast<const __half*>(b), (int)cb,
                            reinterpret_cast<__half*>(y), rows));
  return 0;
}

extern "C" int ivv_ncfhw_to_frames(const void* x, int32_t x_is_f32, void* y, int64_t b, int64_t c, int64_t f,
                                   int64_t hw, int64_t c_pad, ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && b > 0 && c > 0 && f > 0 && hw > 0 && c_pad >= c, "ivv_ncfhw_to_frames: bad arguments");
  const long long total = b * f * hw;
  if (x_is_f32)
    ncfhw_to_frames_kernel<float><<<grid_for(total, 256), 256, 0, STREAM>>>(
        reinterpret_cast<const float*>(x), reinterpret_cast<__half*>(y), b, (int)c, (int)f, hw, (int)c_pad);
  else
    ncfhw_to_frames_kernel<__half><<<grid_for(total, 256), 256, 0, STREAM>>>(
        reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), b, (int)c, (int)f, hw, (int)c_pad);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_frames_to_ncfhw(const void* x, int32_t x_is_f32, int64_t ld, void* y, int32_t y_is_f32, int64_t b,
                                   int64_t c, int64_t f, int64_t hw, ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && b > 0 && c > 0 && f > 0 && hw > 0 && ld >= c, "ivv_frames_to_ncfhw: bad arguments");
  const long long total = b * f * hw;
  const unsigned g = grid_for(total, 256);
  if (x_is_f32 && y_is_f32)
    frames_to_ncfhw_kernel<float, float><<<g, 256, 0, STREAM>>>(reinterpret_cast<const float*>(x), ld,
                                                                reinterpret_cast<float*>(y), b, (int)c, (int)f, hw);
  else if (x_is_f32)
    frames_to_ncfhw_kernel<float, __half><<<g, 256, 0, STREAM>>>(reinterpret_cast<const float*>(x), ld,
                                                                 reinterpret_cast<__half*>(y), b, (int)c, (int)f, hw);
  else if (y_is_f32)
    frames_to_ncfhw_kernel<__half, float><<<g, 256, 0, STREAM>>>(reinterpret_cast<const __half*>(x), ld,
                                                                 reinterpret_cast<float*>(y), b, (int)c, (int)f, hw);
  else
    frames_to_ncfhw_kernel<__half, __half><<<g, 256, 0, STREAM>>>(reinterpret_cast<const __half*>(x), ld,
                                                                  reinterpret_cast<__half*>(y), b, (int)c, (int)f, hw);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_timestep_embedding(const float* t, void* y, int64_t n, int32_t dim, int32_t flip_sin_to_cos,
                                      float freq_shift, ivv_stream_t stream_) {
  IVV_REQUIRE(t && y && n > 0 && dim > 0 && dim % 2 == 0, "ivv_timestep_embedding: bad arguments");
  timestep_embedding_kernel<<<grid_for(n * dim / 2, 128), 128, 0, STREAM>>>(t, reinterpret_cast<__half*>(y), (int)n,
                                                                            dim, flip_sin_to_cos, freq_shift);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_silu(const void* x, void* y, int64_t n, ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && n > 0, "ivv_silu: bad arguments");
  silu_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y),
                                                    n);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_scale(const void* x, void* y, int64_t n, float a, float b, ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && n > 0, "ivv_scale: bad arguments");
  scale_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y),
                                                     n, a, b);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_cfg_ddim_step(const float* eps3, float* latent, float* eps_out, int64_t n, float text_cfg,
                                 float img_cfg, float alpha_prod_t, float alpha_prod_prev, ivv_stream_t stream_) {
  IVV_REQUIRE(eps3 && latent && n > 0, "ivv_cfg_ddim_step: bad arguments");
  IVV_REQUIRE(alpha_prod_t > 0.f && alpha_prod_t <= 1.f && alpha_prod_prev > 0.f && alpha_prod_prev <= 1.f,
              "ivv_cfg_ddim_step: alphas must be in (0, 1]");
  cfg_ddim_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(eps3, latent, eps_out, n, text_cfg, img_cfg,
                                                        sqrtf(alpha_prod_t), sqrtf(1.f - alpha_prod_t),
                                                        sqrtf(alpha_prod_prev), sqrtf(1.f - alpha_prod_prev));
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_frames_u8_to_f32(const void* x, float* y, int64_t n, int64_t hw, int32_t swap_rb,
                                    ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && n > 0 && hw > 0, "ivv_frames_u8_to_f32: bad arguments");
  u8hwc_to_f32chw_kernel<<<grid_for(n * hw, 256), 256, 0, STREAM>>>(reinterpret_cast<const uint8_t*>(x), y, n, hw,
                                                                  swap_rb);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_frames_to_u8(const void* x, int32_t x_is_f32, void* y, int64_t n, int64_t hw, ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && n > 0 && hw > 0, "ivv_frames_to_u8: bad arguments");
  if (x_is_f32)
    chw_to_u8hwc_kernel<float><<<grid_for(n * hw, 256), 256, 0, STREAM>>>(reinterpret_cast<const float*>(x),
                                                                          reinterpret_cast<uint8_t*>(y), n, hw);
  else
    chw_to_u8hwc_kernel<__half><<<grid_for(n * hw, 256), 256, 0, STREAM>>>(reinterpret_cast<const __half*>(x),
                                                                           reinterpret_cast<uint8_t*>(y), n, hw);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}
