// Optical-flow motion compensation: bilinear gather kernels (coalesced along x; fp32 NCHW like the reference).
// Reference: misc_utils/flow_utils.py:25-86 and the per-step correction loop pl_trainer/inference/inference.py:374-386.
#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

struct Bilin {
  int x0, y0;
  float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1); zero when the corner is outside the image
};

// sampling position exactly as the reference builds it: grid = pixel + flow, normalised to [-1,1] with (size-1)
// (flow_utils.py:43-52), then F.grid_sample(align_corners=True) un-normalises with ((g + 1) / 2) * (size - 1).
__device__ __forceinline__ Bilin bilinear_zeros(float px, float py, int w, int h) {
  const float gx = 2.f * (px / (float)(w - 1) - 0.5f);
  const float gy = 2.f * (py / (float)(h - 1) - 0.5f);
  const float ix = ((gx + 1.f) / 2.f) * (float)(w - 1);
  const float iy = ((gy + 1.f) / 2.f) * (float)(h - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  Bilin b;
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

__device__ __forceinline__ float gather4(const float* __restrict__ img, const Bilin& b, int w, int h) {
  float v = 0.f;
  const int x0 = b.x0, y0 = b.y0;
  if (b.w00 != 0.f) v += b.w00 * img[(long long)y0 * w + x0];
  if (b.w01 != 0.f) v += b.w01 * img[(long long)y0 * w + x0 + 1];
  if (b.w10 != 0.f) v += b.w10 * img[(long long)(y0 + 1) * w + x0];
  if (b.w11 != 0.f) v += b.w11 * img[(long long)(y0 + 1) * w + x0 + 1];
  return v;
}

__global__ void warp_image_kernel(const float* __restrict__ image, const float* __restrict__ flow,
                                  float* __restrict__ out, long long n, int c, int h, int w) {
  const long long hw = (long long)h * w;
  const long long total = n * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long ni = i / hw;
    const int y = (int)((i % hw) / w), x = (int)(i % w);
    const float u = flow[(ni * 2 + 0) * hw + (long long)y * w + x];
    const float v = flow[(ni * 2 + 1) * hw + (long long)y * w + x];
    const Bilin b = bilinear_zeros((float)x + u, (float)y + v, w, h);
    for (int ch = 0; ch < c; ++ch)
      out[(ni * c + ch) * hw + (long long)y * w + x] = gather4(image + (ni * c + ch) * hw, b, w, h);
  }
}

// F.interpolate(mode='bilinear', align_corners=False) source index (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int* i0, int* i1, float* l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  *i0 = (int)s;
  *i1 = *i0 + ((*i0 < in_size - 1) ? 1 : 0);
  *l1 = s - (float)*i0;
}

__global__ void resize_flow_kernel(const float* __restrict__ flow, float* __restrict__ out, long long n, int h, int w,
                                   int ho, int wo) {
  const long long total = n * 2 * ho * wo;
  const float sy = (float)h / (float)ho, sx = (float)w / (float)wo;
  const float mul_x = (float)((double)wo / (double)w), mul_y = (float)((double)ho / (double)h);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo);
    const int oy = (int)((i / wo) % ho);
    const long long nc = i / ((long long)wo * ho);
    const int ch = (int)(nc % 2);
    const float mul = ch == 0 ? mul_x : mul_y;
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(oy, sy, h, &y0, &y1, &ly);
    src_index(ox, sx, w, &x0, &x1, &lx);
    const float* f = flow + nc * (long long)h * w;
    const float v00 = f[(long long)y0 * w + x0] * mul, v01 = f[(long long)y0 * w + x1] * mul;
    const float v10 = f[(long long)y1 * w + x0] * mul, v11 = f[(long long)y1 * w + x1] * mul;
    out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

// eps[q, c, y, x] += where(msum > 0.5, (sum_r warp(delta[r], flow[q, r])) / msum, 0), msum = sum_r warp(1, flow[q, r])
__global__ void flow_noise_correction_kernel(const float* __restrict__ delta, const float* __restrict__ flow,
                                             float* __restrict__ eps, int Q, int R, int C, int h, int w) {
  const long long hw = (long long)h * w;
  const long long total = (long long)Q * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i / hw);
    const long long pix = i % hw;
    const int y = (int)(pix / w), x = (int)(pix % w);
    float msum = 0.f;
    float acc[8];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) acc[ch] = 0.f;
    for (int r = 0; r < R; ++r) {
      const float* fl = flow + ((long long)(q * R + r) * 2) * hw;
      const Bilin b = bilinear_zeros((float)x + fl[pix], (float)y + fl[hw + pix], w, h);
      msum += b.w00 + b.w01 + b.w10 + b.w11;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        if (ch < C) acc[ch] += gather4(delta + ((long long)r * C + ch) * hw, b, w, h);
    }
    if (msum > 0.5f) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        if (ch < C) eps[((long long)q * C + ch) * hw + pix] += acc[ch] / msum;
    }
  }
}

static inline unsigned wgrid(long long total) {
  long long b = (total + 127) / 128;
  if (b > 148LL * 16) b = 148LL * 16;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace ivv

using namespace ivv;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" int ivv_warp_image(const float* image, const float* flow, float* out, int64_t n, int64_t c, int64_t h,
                              int64_t w, ivv_stream_t stream_) {
  IVV_REQUIRE(image && flow && out && n > 0 && c > 0 && h > 0 && w > 0, "ivv_warp_image: bad arguments");
  warp_image_kernel<<<wgrid(n * h * w), 128, 0, STREAM>>>(image, flow, out, n, (int)c, (int)h, (int)w);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_resize_flow(const float* flow, float* out, int64_t n, int64_t h, int64_t w, int64_t ho, int64_t wo,
                               ivv_stream_t stream_) {
  IVV_REQUIRE(flow && out && n > 0 && h > 0 && w > 0 && ho > 0 && wo > 0, "ivv_resize_flow: bad arguments");
  resize_flow_kernel<<<wgrid(n * 2 * ho * wo), 128, 0, STREAM>>>(flow, out, n, (int)h, (int)w, (int)ho, (int)wo);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_flow_noise_correction(const float* delta_ref, const float* flow_lat, float* eps, int64_t q,
                                         int64_t r, int64_t c, int64_t h, int64_t w, ivv_stream_t stream_) {
  IVV_REQUIRE(delta_ref && flow_lat && eps && q > 0 && r > 0 && h > 0 && w > 0, "ivv_flow_noise_correction: bad args");
  IVV_REQUIRE(c > 0 && c <= 8, "ivv_flow_noise_correction: c (%lld) must be in 1..8", (long long)c);
  flow_noise_correction_kernel<<<wgrid(q * h * w), 128, 0, STREAM>>>(delta_ref, flow_lat, eps, (int)q, (int)r, (int)c,
                                                                     (int)h, (int)w);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}
