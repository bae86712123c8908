// GroupNorm(+SiLU) and LayerNorm(+positional encoding) for channels-last fp16 activations. HBM-bound kernels:
// 16-byte vector loads, fp32 per-thread partials, double-precision cross-CTA merge (statistics match the fp32
// reference to ~1e-7). Reference call sites: ivv.h (K6/K7/K8).
#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

// ------------------------------------------------------------------------------------------------
// GroupNorm pass 1: per (batch-group, channel-group) sum and sum of squares
//   x: [n_bg, rows_per_bg, C]; thread owns one 8-channel vector column and strides over rows.
// ------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __half* __restrict__ x, double* __restrict__ stats, long long rows_per_bg, int C,
                                int groups, long long rows_per_cta, int V, int R) {
  __shared__ float s_acc[64 * 2];  // groups <= 64
  const int bg = blockIdx.y;
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(rows_per_bg, row_begin + rows_per_cta);
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int vec = threadIdx.x % V;
  const int rsub = threadIdx.x / V;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  if (rsub < R) {
    const __half* base = x + ((long long)bg * rows_per_bg) * C + vec * 8;
    for (long long r = row_begin + rsub; r < row_end; r += R) {
      const uint4 u = *reinterpret_cast<const uint4*>(base + r * C);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x;
        ss[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y;
        ss[2 * j + 1] += f.y * f.y;
      }
    }
    const int cpg = C / groups;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (vec * 8 + j) / cpg;
      atomicAdd(&s_acc[2 * g], s[j]);
      atomicAdd(&s_acc[2 * g + 1], ss[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x)
    atomicAdd(&stats[(long long)bg * groups * 2 + i], (double)s_acc[i]);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm pass 2: y = (x - mean) * rstd * gamma + beta, optional SiLU
// ------------------------------------------------------------------------------------------------
__global__ void gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                                const __half* __restrict__ beta, const double* __restrict__ stats,
                                long long rows_per_bg, int C, int groups, float eps, int silu, long long rows_per_cta) {
  extern __shared__ float s_ab[];  // a[C], b[C]
  float* s_a = s_ab;
  float* s_b = s_ab + C;
  const int bg = blockIdx.y;
  const int cpg = C / groups;
  const double inv_n = 1.0 / ((double)rows_per_bg * cpg);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double sum = stats[((long long)bg * groups + g) * 2];
    const double sq = stats[((long long)bg * groups + g) * 2 + 1];
    const double mean = sum * inv_n;
    double var = sq * inv_n - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = rstd * __half2float(gamma[c]);
    s_a[c] = a;
    s_b[c] = __half2float(beta[c]) - (float)mean * a;
  }
  __syncthreads();
  const int V = C / 8;
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(rows_per_bg, row_begin + rows_per_cta);
  const long long total = (row_end - row_begin) * V;
  const __half* xb = x + ((long long)bg * rows_per_bg + row_begin) * C;
  __half* yb = y + ((long long)bg * rows_per_bg + row_begin) * C;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int vec = (int)(i % V);
    const uint4 u = *reinterpret_cast<const uint4*>(xb + i * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      const int c = vec * 8 + 2 * j;
      float v0 = f.x * s_a[c] + s_b[c];
      float v1 = f.y * s_a[c + 1] + s_b[c + 1];
      if (silu) {
        v0 = silu_f(v0);
        v1 = silu_f(v1);
      }
      oh[j] = __floats2half2_rn(v0, v1);
    }
    *reinterpret_cast<uint4*>(yb + i * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (two-pass variance), optional + pe[frame]
// ------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                 const __half* __restrict__ gamma, const __half* __restrict__ beta, long long rows,
                                 int C, float eps, const float* __restrict__ pe, long long rows_per_frame,
                                 long long frames, long long pe_start) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int V = C / 8;
  float v[MAXV][8];
  float sum = 0.f;
  const __half* xr = x + row * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vec = lane + i * 32;
    if (vec < V) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + vec * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[i][2 * j] = f.x;
        v[i][2 * j + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vec = lane + i * 32;
    if (vec < V) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
  const float* per = nullptr;
  if (pe) {
    const long long frame = (row / rows_per_frame) % frames;
    per = pe + (pe_start + frame) * C;
  }
  __half* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vec = lane + i * 32;
    if (vec < V) {
      const uint4 ug = *reinterpret_cast<const uint4*>(gamma + vec * 8);
      const uint4 ub = *reinterpret_cast<const uint4*>(beta + vec * 8);
      const __half2* hg = reinterpret_cast<const __half2*>(&ug);
      const __half2* hb = reinterpret_cast<const __half2*>(&ub);
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 g = __half22float2(hg[j]);
        const float2 b = __half22float2(hb[j]);
        float o0 = (v[i][2 * j] - mean) * rstd * g.x + b.x;
        float o1 = (v[i][2 * j + 1] - mean) * rstd * g.y + b.y;
        if (per) {
          o0 += per[vec * 8 + 2 * j];
          o1 += per[vec * 8 + 2 * j + 1];
        }
        oh[j] = __floats2half2_rn(o0, o1);
      }
      *reinterpret_cast<uint4*>(yr + vec * 8) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// row softmax (VAE mid attention scores): y = softmax(x * scale) per row, fp16 in/out, fp32 math
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void softmax_rows_kernel(const TIn* __restrict__ x, __half* __restrict__ y, long long rows, int cols,
                                    float scale) {
  const long long row = blockIdx.x;
  if (row >= rows) return;
  __shared__ float s_red[32];
  const TIn* xr = x + row * cols;
  __half* yr = y + row * cols;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) m = fmaxf(m, (float)xr[i] * scale);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, s_red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) s += __expf((float)xr[i] * scale - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += s_red[i];
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < cols; i += blockDim.x)
    yr[i] = __float2half_rn(__expf((float)xr[i] * scale - m) * inv);
}

}  // namespace ivv

extern "C" size_t ivv_groupnorm_ws_bytes(int64_t n_img, int32_t groups, int64_t frames_per_group) {
  if (frames_per_group <= 0) return 0;
  return (size_t)(n_img / frames_per_group) * groups * 2 * sizeof(double);
}

extern "C" int ivv_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int64_t n_img, int64_t hw,
                             int64_t c, int32_t groups, int64_t frames_per_group, float eps, int32_t silu,
                             void* stats_ws, size_t stats_ws_bytes, ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(x && y && gamma && beta && stats_ws, "ivv_groupnorm: null pointer");
  IVV_REQUIRE(n_img > 0 && hw > 0 && c > 0, "ivv_groupnorm: empty input");
  IVV_REQUIRE(frames_per_group > 0 && n_img % frames_per_group == 0,
              "ivv_groupnorm: n_img (%lld) must be a multiple of frames_per_group (%lld)", (long long)n_img,
              (long long)frames_per_group);
  IVV_REQUIRE(groups > 0 && groups <= 64 && c % groups == 0, "ivv_groupnorm: bad groups %d for c=%lld", groups,
              (long long)c);
  IVV_REQUIRE(c % 8 == 0 && c <= 8192, "ivv_groupnorm: c (%lld) must be a multiple of 8 and <= 8192", (long long)c);
  const size_t need = ivv_groupnorm_ws_bytes(n_img, groups, frames_per_group);
  IVV_REQUIRE(stats_ws_bytes >= need, "ivv_groupnorm: workspace too small (%zu < %zu)", stats_ws_bytes, need);
  const long long n_bg = n_img / frames_per_group;
  const long long rows_per_bg = frames_per_group * hw;
  IVV_REQUIRE(n_bg <= 65535, "ivv_groupnorm: too many batch groups");
  IVV_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, need, stream));
  const int V = (int)(c / 8);
  const int R = V >= 256 ? 1 : 256 / V;
  const int threads = V * R;
  // ~4 CTAs per SM in total, but at least 8 row sweeps per CTA
  long long chunks = (148 * 4 + n_bg - 1) / n_bg;
  long long rows_per_cta = (rows_per_bg + chunks - 1) / chunks;
  if (rows_per_cta < 8LL * R) rows_per_cta = 8LL * R;
  chunks = (rows_per_bg + rows_per_cta - 1) / rows_per_cta;
  {
    dim3 grid((unsigned)chunks, (unsigned)n_bg);
    gn_stats_kernel<<<grid, threads, 0, stream>>>(reinterpret_cast<const __half*>(x),
                                                  reinterpret_cast<double*>(stats_ws), rows_per_bg, (int)c, groups,
                                                  rows_per_cta, V, R);
    IVV_CHECK_CUDA(cudaGetLastError());
  }
  {
    long long chunks2 = (148 * 8 + n_bg - 1) / n_bg;
    long long rpc = (rows_per_bg + chunks2 - 1) / chunks2;
    if (rpc < 16) rpc = 16;
    chunks2 = (rows_per_bg + rpc - 1) / rpc;
    dim3 grid((unsigned)chunks2, (unsigned)n_bg);
    const size_t smem = (size_t)c * 2 * sizeof(float);
    gn_apply_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y),
                                                 reinterpret_cast<const __half*>(gamma),
                                                 reinterpret_cast<const __half*>(beta),
                                                 reinterpret_cast<const double*>(stats_ws), rows_per_bg, (int)c, groups,
                                                 eps, silu, rpc);
    IVV_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int ivv_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t rows, int64_t c,
                             float eps, const float* pe, int64_t rows_per_frame, int64_t frames, int64_t pe_start,
                             ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(x && y && gamma && beta, "ivv_layernorm: null pointer");
  IVV_REQUIRE(rows > 0 && c > 0 && c % 8 == 0 && c <= 2048, "ivv_layernorm: c (%lld) must be a multiple of 8, <= 2048",
              (long long)c);
  IVV_REQUIRE(!pe || (rows_per_frame > 0 && frames > 0), "ivv_layernorm: pe given without frame geometry");
  const int warps = 8;
  const long long blocks = (rows + warps - 1) / warps;
  const __half* xx = reinterpret_cast<const __half*>(x);
  __half* yy = reinterpret_cast<__half*>(y);
  const __half* g = reinterpret_cast<const __half*>(gamma);
  const __half* b = reinterpret_cast<const __half*>(beta);
  if (c <= 256 * 2)
    layernorm_kernel<2><<<(unsigned)blocks, warps * 32, 0, stream>>>(xx, yy, g, b, rows, (int)c, eps, pe,
                                                                     rows_per_frame, frames, pe_start);
  else if (c <= 256 * 5)
    layernorm_kernel<5><<<(unsigned)blocks, warps * 32, 0, stream>>>(xx, yy, g, b, rows, (int)c, eps, pe,
                                                                     rows_per_frame, frames, pe_start);
  else
    layernorm_kernel<8><<<(unsigned)blocks, warps * 32, 0, stream>>>(xx, yy, g, b, rows, (int)c, eps, pe,
                                                                     rows_per_frame, frames, pe_start);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_softmax_rows(const void* x, int32_t x_is_f32, void* y, int64_t rows, int64_t cols, float scale,
                                ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(x && y && rows > 0 && cols > 0, "ivv_softmax_rows: bad arguments");
  if (x_is_f32)
    softmax_rows_kernel<float><<<(unsigned)rows, 256, 0, stream>>>(reinterpret_cast<const float*>(x),
                                                                   reinterpret_cast<__half*>(y), rows, (int)cols, scale);
  else
    softmax_rows_kernel<__half><<<(unsigned)rows, 256, 0, stream>>>(reinterpret_cast<const __half*>(x),
                                                                    reinterpret_cast<__half*>(y), rows, (int)cols, scale);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}
