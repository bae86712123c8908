// RAFT optical flow (torchvision.models.optical_flow.raft_large, used by the reference through
// misc_utils/flow_utils.py:134-189): the pieces that are not convolutions. Convolutions run on ivv_gemm (implicit
// 3x3 / 1x5 / 5x1 taps, or ivv_im2col + GEMM for the strided and 7x7 ones); norms on ivv_channelnorm.
// Everything here is launch-latency / HBM bound glue at 1/8 resolution; layouts are channels-last.
#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

static inline unsigned rgrid(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ---------------------------------------------------------------------------------------------------------------
// generic im2col: out[(n,oy,ox), (ky*kw+kx)*c + ci] = x[n, oy*stride+ky-pad_h, ox*stride+kx-pad_w, ci] (zero outside)
// ---------------------------------------------------------------------------------------------------------------
__global__ void im2col_kernel(const __half* __restrict__ x, __half* __restrict__ out, long long n_img, int h, int w,
                              int c, int kh, int kw, int stride, int pad_h, int pad_w, int ho, int wo) {
  griddep_sync();
  const int V = c / 8;
  const int taps = kh * kw;
  const long long total = n_img * ho * wo * taps * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    const int tap = (int)((i / V) % taps);
    const long long pix = i / ((long long)V * taps);
    const int ox = (int)(pix % wo);
    const int oy = (int)((pix / wo) % ho);
    const long long n = pix / ((long long)wo * ho);
    const int iy = oy * stride + tap / kw - pad_h;
    const int ix = ox * stride + tap % kw - pad_w;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w)
      u = *reinterpret_cast<const uint4*>(x + ((n * h + iy) * w + ix) * c + v * 8);
    *reinterpret_cast<uint4*>(out + pix * ((long long)taps * c) + (long long)tap * c + v * 8) = u;
  }
}

// F.interpolate(mode='bilinear', align_corners=False) source index (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void src_index_r(int dst, float scale, int in_size, int* i0, int* i1, float* l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  *i0 = (int)s;
  *i1 = *i0 + ((*i0 < in_size - 1) ? 1 : 0);
  *l1 = s - (float)*i0;
}

// images fp32 [n, 3, hs, ws] -> fp16 [n, h, w, 8]: optional TF.resize(antialias=False) (flow_utils.py:180-182), then the
// OpticalFlow preset transform 2 v - 1 (torchvision transforms/_presets.py OpticalFlow.forward); channels 3..7 = 0
__global__ void raft_prep_kernel(const float* __restrict__ img, __half* __restrict__ out, long long n, int hs, int ws,
                                 int h, int w) {
  griddep_sync();
  const long long total = n * h * w;
  const float sy = (float)hs / (float)h, sx = (float)ws / (float)w;
  const bool resize = (hs != h) || (ws != w);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const long long ni = i / ((long long)w * h);
    float v[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* f = img + (ni * 3 + ch) * (long long)hs * ws;
      if (!resize) {
        v[ch] = f[(long long)y * ws + x];
      } else {
        int y0, y1, x0, x1;
        float ly, lx;
        src_index_r(y, sy, hs, &y0, &y1, &ly);
        src_index_r(x, sx, ws, &x0, &x1, &lx);
        v[ch] = (1.f - ly) * ((1.f - lx) * f[(long long)y0 * ws + x0] + lx * f[(long long)y0 * ws + x1]) +
                ly * ((1.f - lx) * f[(long long)y1 * ws + x0] + lx * f[(long long)y1 * ws + x1]);
      }
      v[ch] = (v[ch] - 0.5f) / 0.5f;
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
    oh[0] = __floats2half2_rn(v[0], v[1]);
    oh[1] = __floats2half2_rn(v[2], 0.f);
    oh[2] = __floats2half2_rn(0.f, 0.f);
    oh[3] = oh[2];
    *reinterpret_cast<uint4*>(out + i * 8) = o;
  }
}

// F.avg_pool2d(kernel 2, stride 2) on the last two dims of fp32 [n, h, w] (CorrBlock.build_pyramid, raft.py:388-390)
__global__ void avgpool2_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int h, int w) {
  griddep_sync();
  const int ho = h / 2, wo = w / 2;
  const long long total = n * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo), oy = (int)((i / wo) % ho);
    const long long ni = i / ((long long)wo * ho);
    const float* p = x + (ni * h + 2 * oy) * w + 2 * ox;
    y[i] = (p[0] + p[1] + p[w] + p[w + 1]) * 0.25f;
  }
}

// CorrBlock.index_pyramid (raft.py:393-421): for level l the centroid is coords / 2^l; the (2r+1)^2 window is sampled
// bilinearly (grid_sample align_corners=True, zero padding). Channel order: level, then i (offset added to X), then j
// (offset added to Y) - torchvision's meshgrid(di, dj, "ij") stacked as (x, y).
struct PyrPtrs {
  const float* lvl[4];
};
__global__ void corr_lookup_kernel(PyrPtrs pyr, const float* __restrict__ coords, __half* __restrict__ out,
                                   long long rows, int hw, int h, int w, int levels, int radius, float scale,
                                   int out_ld) {
  griddep_sync();
  const int side = 2 * radius + 1;
  const int per_level = side * side;
  const int per_row = levels * per_level;
  const long long total = rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % per_row);
    const long long row = i / per_row;
    const int l = ch / per_level;
    const int k = ch - l * per_level;
    const int ii = k / side, jj = k - ii * side;
    const int hl = h >> l, wl = w >> l;
    const float inv = 1.f / (float)(1 << l);
    const float cx = coords[row * 2] * inv + (float)(ii - radius);
    const float cy = coords[row * 2 + 1] * inv + (float)(jj - radius);
    // torchvision _utils.grid_sample normalises with (size - 1) and F.grid_sample(align_corners=True) undoes it
    const float gx = 2.f * cx / (float)(wl - 1) - 1.f;
    const float gy = hl > 1 ? 2.f * cy / (float)(hl - 1) - 1.f : cy;
    const float ix = ((gx + 1.f) * 0.5f) * (float)(wl - 1);
    const float iy = ((gy + 1.f) * 0.5f) * (float)(hl - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float tx = ix - fx, ty = iy - fy;
    const float* img = pyr.lvl[l] + row * (long long)(hl * wl);
    float v = 0.f;
    const bool x0in = x0 >= 0 && x0 < wl, x1in = x0 + 1 >= 0 && x0 + 1 < wl;
    const bool y0in = y0 >= 0 && y0 < hl, y1in = y0 + 1 >= 0 && y0 + 1 < hl;
    if (x0in && y0in) v += (1.f - tx) * (1.f - ty) * img[y0 * wl + x0];
    if (x1in && y0in) v += tx * (1.f - ty) * img[y0 * wl + x0 + 1];
    if (x0in && y1in) v += (1.f - tx) * ty * img[(y0 + 1) * wl + x0];
    if (x1in && y1in) v += tx * ty * img[(y0 + 1) * wl + x0 + 1];
    out[row * out_ld + ch] = __float2half_rn(v * scale);
  }
  (void)hw;
}

// hidden = tanh(ctx[:, :hid]), context = relu(ctx[:, hid:]) (raft.py:512-514) written straight into the GRU input
// buffer hx = [h | context | motion features]; an fp32 master copy of h is kept for the recurrent update
__global__ void raft_init_state_kernel(const __half* __restrict__ ctx, int ctx_ld, float* __restrict__ h32,
                                       __half* __restrict__ hx, int hx_ld, long long rows, int hid, int cctx) {
  griddep_sync();
  const int per_row = hid + cctx;
  const long long total = rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % per_row);
    const long long row = i / per_row;
    const float v = __half2float(ctx[row * ctx_ld + ch]);
    if (ch < hid) {
      const float t = tanhf(v);
      h32[row * hid + ch] = t;
      hx[row * hx_ld + ch] = __float2half_rn(t);
    } else {
      hx[row * hx_ld + ch] = __float2half_rn(fmaxf(v, 0.f));
    }
  }
}

// y = relu(a + b) (eval-mode ResidualBlock join, raft.py:71), 8 halfs per thread
__global__ void add_relu_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ y,
                                long long n8) {
  griddep_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 ua = reinterpret_cast<const uint4*>(a)[i], ub = reinterpret_cast<const uint4*>(b)[i];
    const __half2* ha = reinterpret_cast<const __half2*>(&ua);
    const __half2* hb = reinterpret_cast<const __half2*>(&ub);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __half22float2(ha[j]), fb = __half22float2(hb[j]);
      oh[j] = __floats2half2_rn(fmaxf(fa.x + fb.x, 0.f), fmaxf(fa.y + fb.y, 0.f));
    }
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// ConvGRU (raft.py:222-229): rh = sigmoid(r_pre) * h. zrq = [z_pre | r_pre | qx_pre] fp16, ld = zrq_ld.
__global__ void gru_gate_r_kernel(const __half* __restrict__ zrq, int zrq_ld, const float* __restrict__ h32,
                                  __half* __restrict__ rh, long long rows, int hid) {
  griddep_sync();
  const long long total = rows * hid;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % hid);
    const long long row = i / hid;
    const float r = sigmoid_f(__half2float(zrq[row * zrq_ld + hid + ch]));
    rh[i] = __float2half_rn(r * h32[i]);
  }
}

// h' = (1 - z) h + z tanh(q_pre), z = sigmoid(z_pre); fp32 master updated in place, fp16 copy into hx[:, :hid]
__global__ void gru_update_kernel(const __half* __restrict__ zrq, int zrq_ld, const __half* __restrict__ q_pre,
                                  float* __restrict__ h32, __half* __restrict__ hx, int hx_ld, long long rows,
                                  int hid) {
  griddep_sync();
  const long long total = rows * hid;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % hid);
    const long long row = i / hid;
    const float z = sigmoid_f(__half2float(zrq[row * zrq_ld + ch]));
    const float q = tanhf(__half2float(q_pre[i]));
    const float hn = (1.f - z) * h32[i] + z * q;
    h32[i] = hn;
    hx[row * hx_ld + ch] = __float2half_rn(hn);
  }
}

// coords1 += delta (raft.py:527); flow = coords1 - coords0 goes, as fp16, to the 8-channel input of the motion
// encoder's 7x7 convolution and to the two flow channels appended to the motion features (raft.py:211)
__global__ void raft_update_coords_kernel(const float* __restrict__ delta, int delta_ld, float* __restrict__ coords1,
                                          __half* __restrict__ flow8, __half* __restrict__ flow_slot,
                                          int flow_slot_ld, long long rows, int hw, int w) {
  griddep_sync();
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < rows;
       row += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(row % hw);
    const float x0 = (float)(p % w), y0 = (float)(p / w);
    float cx = coords1[row * 2], cy = coords1[row * 2 + 1];
    if (delta) {
      cx += delta[row * delta_ld];
      cy += delta[row * delta_ld + 1];
      coords1[row * 2] = cx;
      coords1[row * 2 + 1] = cy;
    }
    const __half fx = __float2half_rn(cx - x0), fy = __float2half_rn(cy - y0);
    uint4 o = make_uint4(0, 0, 0, 0);
    __half* oh = reinterpret_cast<__half*>(&o);
    oh[0] = fx;
    oh[1] = fy;
    *reinterpret_cast<uint4*>(flow8 + row * 8) = o;
    if (flow_slot) {
      flow_slot[row * flow_slot_ld] = fx;
      flow_slot[row * flow_slot_ld + 1] = fy;
    }
  }
}

// upsample_flow with a learnt convex mask (torchvision _utils.py upsample_flow): mask fp16 [rows, 9*64] with channel
// k*64 + i*8 + j; out fp32 [b, 2, 8h, 8w] = sum_k softmax_k(mask) * 8 * flow[neighbour k] (3x3, zero padded)
__global__ void convex_upsample_kernel(const __half* __restrict__ mask, int mask_ld, const float* __restrict__ coords1,
                                       float* __restrict__ out, long long b, int h, int w) {
  griddep_sync();
  const long long total = b * h * w * 64;
  const int hw = h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int sub = (int)(i & 63);
    const int si = sub >> 3, sj = sub & 7;
    const long long row = i >> 6;
    const int p = (int)(row % hw);
    const long long bi = row / hw;
    const int y = p / w, x = p % w;
    float m[9], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      m[k] = __half2float(mask[row * mask_ld + k * 64 + sub]);
      mx = fmaxf(mx, m[k]);
    }
    float den = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float e = __expf(m[k] - mx);
      den += e;
      const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const long long r2 = bi * hw + (long long)yy * w + xx;
        ax += e * 8.f * (coords1[r2 * 2] - (float)xx);
        ay += e * 8.f * (coords1[r2 * 2 + 1] - (float)yy);
      }
    }
    const long long oy = 8LL * y + si, ox = 8LL * x + sj;
    const long long plane = 64LL * hw;
    out[(bi * 2 + 0) * plane + oy * (8LL * w) + ox] = ax / den;
    out[(bi * 2 + 1) * plane + oy * (8LL * w) + ox] = ay / den;
  }
}

}  // namespace ivv

using namespace ivv;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" int ivv_im2col(const void* x, void* out, int64_t n_img, int64_t h, int64_t w, int64_t c, int32_t kh,
                          int32_t kw, int32_t stride, int32_t pad_h, int32_t pad_w, int64_t ho, int64_t wo,
                          ivv_stream_t stream_) {
  IVV_REQUIRE(x && out && n_img > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0,
              "ivv_im2col: bad arguments (c must be a multiple of 8)");
  IVV_REQUIRE(kh > 0 && kw > 0 && stride > 0 && pad_h >= 0 && pad_w >= 0, "ivv_im2col: bad window");
  IVV_REQUIRE(ho == (h + 2 * pad_h - kh) / stride + 1 && wo == (w + 2 * pad_w - kw) / stride + 1,
              "ivv_im2col: output size must be (size + 2 pad - k) / stride + 1");
  const long long total = n_img * ho * wo * kh * kw * (c / 8);
  IVV_CHECK_CUDA(launch_pdl(im2col_kernel, dim3(rgrid(total, 256)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(out), (long long)n_img,
                            (int)h, (int)w, (int)c, (int)kh, (int)kw, (int)stride, (int)pad_h, (int)pad_w, (int)ho,
                            (int)wo));
  return 0;
}

extern "C" int ivv_add_relu(const void* a, const void* b, void* y, int64_t n, ivv_stream_t stream_) {
  IVV_REQUIRE(a && b && y && n > 0 && n % 8 == 0, "ivv_add_relu: n must be a positive multiple of 8");
  IVV_CHECK_CUDA(launch_pdl(add_relu_kernel, dim3(rgrid(n / 8, 256)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(a), reinterpret_cast<const __half*>(b),
                            reinterpret_cast<__half*>(y), (long long)(n / 8)));
  return 0;
}

extern "C" int ivv_raft_prep_images(const float* img, void* out, int64_t n, int64_t hs, int64_t ws, int64_t h,
                                    int64_t w, ivv_stream_t stream_) {
  IVV_REQUIRE(img && out && n > 0 && hs > 0 && ws > 0 && h > 0 && w > 0, "ivv_raft_prep_images: bad arguments");
  IVV_CHECK_CUDA(launch_pdl(raft_prep_kernel, dim3(rgrid(n * h * w, 256)), dim3(256), 0, STREAM, img,
                            reinterpret_cast<__half*>(out), (long long)n, (int)hs, (int)ws, (int)h, (int)w));
  return 0;
}

extern "C" int ivv_avgpool2_f32(const float* x, float* y, int64_t n, int64_t h, int64_t w, ivv_stream_t stream_) {
  IVV_REQUIRE(x && y && n > 0 && h >= 2 && w >= 2, "ivv_avgpool2_f32: bad arguments");
  IVV_CHECK_CUDA(launch_pdl(avgpool2_kernel, dim3(rgrid(n * (h / 2) * (w / 2), 256)), dim3(256), 0, STREAM, x, y,
                            (long long)n, (int)h, (int)w));
  return 0;
}

extern "C" int ivv_corr_lookup(const float* const* pyramid, int32_t levels, const float* coords, void* out,
                               int64_t out_ld, int64_t n_pairs, int64_t h, int64_t w, int32_t radius, float scale,
                               ivv_stream_t stream_) {
  IVV_REQUIRE(pyramid && coords && out && n_pairs > 0 && h > 0 && w > 0, "ivv_corr_lookup: bad arguments");
  IVV_REQUIRE(levels >= 1 && levels <= 4 && radius >= 1 && radius <= 7, "ivv_corr_lookup: levels in 1..4, radius 1..7");
  IVV_REQUIRE((h >> (levels - 1)) >= 2 && (w >> (levels - 1)) >= 2,
              "ivv_corr_lookup: feature map %lldx%lld too small for %d pyramid levels", (long long)h, (long long)w,
              levels);
  const int per_row = levels * (2 * radius + 1) * (2 * radius + 1);
  IVV_REQUIRE(out_ld >= per_row, "ivv_corr_lookup: out_ld (%lld) < %d channels", (long long)out_ld, per_row);
  PyrPtrs pp{};
  for (int l = 0; l < levels; ++l) {
    IVV_REQUIRE(pyramid[l] != nullptr, "ivv_corr_lookup: null pyramid level %d", l);
    pp.lvl[l] = pyramid[l];
  }
  const long long rows = n_pairs * h * w;
  IVV_CHECK_CUDA(launch_pdl(corr_lookup_kernel, dim3(rgrid(rows * per_row, 256)), dim3(256), 0, STREAM, pp, coords,
                            reinterpret_cast<__half*>(out), rows, (int)(h * w), (int)h, (int)w, (int)levels,
                            (int)radius, scale, (int)out_ld));
  return 0;
}

extern "C" int ivv_raft_init_state(const void* ctx, int64_t ctx_ld, float* h32, void* hx, int64_t hx_ld, int64_t rows,
                                   int32_t hidden, int32_t context, ivv_stream_t stream_) {
  IVV_REQUIRE(ctx && h32 && hx && rows > 0 && hidden > 0 && context > 0, "ivv_raft_init_state: bad arguments");
  IVV_REQUIRE(ctx_ld >= hidden + context && hx_ld >= hidden + context, "ivv_raft_init_state: leading dims too small");
  IVV_CHECK_CUDA(launch_pdl(raft_init_state_kernel, dim3(rgrid(rows * (hidden + context), 256)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(ctx), (int)ctx_ld, h32, reinterpret_cast<__half*>(hx),
                            (int)hx_ld, (long long)rows, (int)hidden, (int)context));
  return 0;
}

extern "C" int ivv_gru_gate_r(const void* zrq, int64_t zrq_ld, const float* h32, void* rh, int64_t rows,
                              int32_t hidden, ivv_stream_t stream_) {
  IVV_REQUIRE(zrq && h32 && rh && rows > 0 && hidden > 0 && zrq_ld >= 2 * hidden, "ivv_gru_gate_r: bad arguments");
  IVV_CHECK_CUDA(launch_pdl(gru_gate_r_kernel, dim3(rgrid(rows * hidden, 256)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(zrq), (int)zrq_ld, h32, reinterpret_cast<__half*>(rh),
                            (long long)rows, (int)hidden));
  return 0;
}

extern "C" int ivv_gru_update(const void* zrq, int64_t zrq_ld, const void* q_pre, float* h32, void* hx, int64_t hx_ld,
                              int64_t rows, int32_t hidden, ivv_stream_t stream_) {
  IVV_REQUIRE(zrq && q_pre && h32 && hx && rows > 0 && hidden > 0 && zrq_ld >= hidden && hx_ld >= hidden,
              "ivv_gru_update: bad arguments");
  IVV_CHECK_CUDA(launch_pdl(gru_update_kernel, dim3(rgrid(rows * hidden, 256)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(zrq), (int)zrq_ld, reinterpret_cast<const __half*>(q_pre),
                            h32, reinterpret_cast<__half*>(hx), (int)hx_ld, (long long)rows, (int)hidden));
  return 0;
}

extern "C" int ivv_raft_update_coords(const float* delta, int64_t delta_ld, float* coords1, void* flow8,
                                      void* flow_slot, int64_t flow_slot_ld, int64_t n_pairs, int64_t h, int64_t w,
                                      ivv_stream_t stream_) {
  IVV_REQUIRE(coords1 && flow8 && n_pairs > 0 && h > 0 && w > 0, "ivv_raft_update_coords: bad arguments");
  IVV_REQUIRE(!delta || delta_ld >= 2, "ivv_raft_update_coords: delta_ld must be >= 2");
  IVV_REQUIRE(!flow_slot || flow_slot_ld >= 2, "ivv_raft_update_coords: flow_slot_ld must be >= 2");
  const long long rows = n_pairs * h * w;
  IVV_CHECK_CUDA(launch_pdl(raft_update_coords_kernel, dim3(rgrid(rows, 128)), dim3(128), 0, STREAM, delta,
                            (int)delta_ld, coords1, reinterpret_cast<__half*>(flow8),
                            reinterpret_cast<__half*>(flow_slot), (int)flow_slot_ld, rows, (int)(h * w), (int)w));
  return 0;
}

extern "C" int ivv_convex_upsample(const void* mask, int64_t mask_ld, const float* coords1, float* out,
                                   int64_t n_pairs, int64_t h, int64_t w, ivv_stream_t stream_) {
  IVV_REQUIRE(mask && coords1 && out && n_pairs > 0 && h > 0 && w > 0 && mask_ld >= 576,
              "ivv_convex_upsample: bad arguments (mask needs 9*8*8 channels)");
  IVV_CHECK_CUDA(launch_pdl(convex_upsample_kernel, dim3(rgrid(n_pairs * h * w * 64, 256)), dim3(256), 0, STREAM,
                            reinterpret_cast<const __half*>(mask), (int)mask_ld, coords1, out, (long long)n_pairs,
                            (int)h, (int)w));
  return 0;
}
