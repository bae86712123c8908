// Temporal self-attention over the frame axis (AnimateDiff motion module), L = frames <= 32.
// HBM-bound: 0.1 % of the UNet FLOPs but 4 full activation passes. The reference transposes (b f) d c -> (b d) f c
// and back around SDPA (motion_module.py:275,334); here the sequence is gathered with a frame stride straight from the
// fused q|k|v projection and written back in token order, so neither transpose copy exists.
// One CTA per (clip, pixel, head group); one warp per head; scores and softmax in fp32 registers.
#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

constexpr int kTPad = 8;  // halfs of row padding: makes 16-byte reads of consecutive frames conflict-free

__global__ void temporal_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ o, int frames, long long hw,
                                     int C, int heads, int heads_per_cta, float scale) {
  extern __shared__ __align__(16) uint8_t smem_t[];
  __half* s = reinterpret_cast<__half*>(smem_t);
  const int d = C / heads;
  const int seg = heads_per_cta * d;        // halfs per q (or k, v) segment in smem
  const int row_ld = 3 * seg + kTPad;       // smem row stride
  const long long bp = blockIdx.x;          // clip * hw + pixel
  const long long clip = bp / hw, pix = bp % hw;
  const int head0 = blockIdx.y * heads_per_cta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- gather: frames x 3 segments of seg halfs ----
  const int vec_per_seg = seg / 8;
  const int total_vec = frames * 3 * vec_per_seg;
  for (int i = threadIdx.x; i < total_vec; i += blockDim.x) {
    const int v = i % vec_per_seg;
    const int part = (i / vec_per_seg) % 3;
    const int f = i / (3 * vec_per_seg);
    const long long row = (clip * frames + f) * hw + pix;
    const uint4 u = *reinterpret_cast<const uint4*>(qkv + row * 3 * C + (long long)part * C + head0 * d + v * 8);
    *reinterpret_cast<uint4*>(s + f * row_ld + part * seg + v * 8) = u;
  }
  __syncthreads();
  if (warp >= heads_per_cta) return;
  const int qoff = warp * d, koff = seg + warp * d, voff = 2 * seg + warp * d;
  const int dv = d / 8;

  for (int i = 0; i < frames; ++i) {
    // score of (query i, key = lane)
    float sc = -INFINITY;
    if (lane < frames) {
      float acc = 0.f;
      const __half* qi = s + i * row_ld + qoff;
      const __half* kj = s + lane * row_ld + koff;
      for (int v = 0; v < dv; ++v) {
        const uint4 uq = *reinterpret_cast<const uint4*>(qi + v * 8);
        const uint4 uk = *reinterpret_cast<const uint4*>(kj + v * 8);
        const __half2* hq = reinterpret_cast<const __half2*>(&uq);
        const __half2* hk = reinterpret_cast<const __half2*>(&uk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __half22float2(hq[j]);
          const float2 b = __half22float2(hk[j]);
          acc += a.x * b.x + a.y * b.y;
        }
      }
      sc = acc * scale;
    }
    const float m = warp_max(sc);
    float p = lane < frames ? __expf(sc - m) : 0.f;
    const float denom = warp_sum(p);
    p /= denom;
    // out_i[dd] = sum_j p_j v_j[dd]; lane handles dd = lane, lane + 32, ...
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    for (int j = 0; j < frames; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const __half* vj = s + j * row_ld + voff;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int dd = lane + t * 32;
        if (dd < d) acc[t] += pj * __half2float(vj[dd]);
      }
    }
    const long long row = (clip * frames + i) * hw + pix;
    __half* orow = o + row * C + (long long)(head0 + warp) * d;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int dd = lane + t * 32;
      if (dd < d) orow[dd] = __float2half_rn(acc[t]);
    }
  }
}

}  // namespace ivv

extern "C" int ivv_temporal_attention(const void* qkv, void* o, int64_t clips, int64_t frames, int64_t hw, int64_t c,
                                      int32_t heads, float scale, ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(qkv && o && clips > 0 && frames > 0 && hw > 0 && c > 0 && heads > 0, "ivv_temporal_attention: bad args");
  IVV_REQUIRE(frames <= 32, "ivv_temporal_attention: frames (%lld) must be <= 32", (long long)frames);
  IVV_REQUIRE(c % heads == 0 && (c / heads) % 8 == 0 && (c / heads) <= 256,
              "ivv_temporal_attention: head dim %lld must be a multiple of 8 and <= 256", (long long)(c / heads));
  const int d = (int)(c / heads);
  int hpc = heads;
  auto smem_for = [&](int h) { return (size_t)frames * (3 * h * d + kTPad) * sizeof(__half); };
  while (hpc > 1 && (smem_for(hpc) > 96 * 1024 || heads % hpc != 0)) --hpc;
  const size_t smem = smem_for(hpc);
  IVV_REQUIRE(smem <= 200 * 1024, "ivv_temporal_attention: tile does not fit shared memory");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  const long long bp = clips * hw;
  IVV_REQUIRE(bp < (1LL << 31), "ivv_temporal_attention: too many sequences");
  dim3 grid((unsigned)bp, (unsigned)(heads / hpc));
  int threads = hpc * 32;
  if (threads < 128) threads = 128;
  temporal_attn_kernel<<<grid, threads, smem, stream>>>(reinterpret_cast<const __half*>(qkv),
                                                        reinterpret_cast<__half*>(o), (int)frames, hw, (int)c, heads,
                                                        hpc, scale);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}
