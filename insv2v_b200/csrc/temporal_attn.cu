// Temporal self-attention over the frame axis (AnimateDiff motion module), L = frames <= 32.
// HBM-bound: 0.1 % of the UNet FLOPs but 4 full activation passes. The reference transposes (b f) d c -> (b d) f c
// and back around SDPA (motion_module.py:275,334); here the sequence is gathered with a frame stride straight from the
// fused q|k|v projection and written back in token order, so neither transpose copy exists.
// One CTA per (clip, pixel, head group); one THREAD per (head, query frame): its 32 scores live in registers, K/V rows
// are broadcast reads from shared memory, the output row overwrites the thread's own Q slot and leaves the CTA as
// coalesced 16-byte stores.
#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

constexpr int kTPad = 8;       // halfs of row padding: 16-byte reads of consecutive frames hit distinct banks
constexpr int kMaxFrames = 32;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

template <int F_MAX>
__global__ void temporal_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ o, int frames, long long hw,
                                     int C, int heads, int heads_per_cta, float scale) {
  extern __shared__ __align__(16) uint8_t smem_t[];
  __half* s = reinterpret_cast<__half*>(smem_t);
  const int d = C / heads;
  const int seg = heads_per_cta * d;   // halfs per q (or k, v) segment of one frame
  const int row_ld = 3 * seg + kTPad;  // smem row stride (halfs)
  const long long bp = blockIdx.x;     // clip * hw + pixel
  const long long clip = bp / hw, pix = bp % hw;
  const int head0 = blockIdx.y * heads_per_cta;

  // ---- gather: frames x 3 segments of seg halfs (16-byte vectors, coalesced per frame row) ----
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

  const int hl = threadIdx.x / frames;  // local head
  const int qi = threadIdx.x % frames;  // query frame
  if (hl < heads_per_cta) {
    const int dv = d / 8;
    __half* qrow = s + qi * row_ld + hl * d;
    const __half* kbase = s + seg + hl * d;
    const __half* vbase = s + 2 * seg + hl * d;
    float sc[F_MAX];
#pragma unroll
    for (int j = 0; j < F_MAX; ++j) sc[j] = 0.f;
    for (int v = 0; v < dv; ++v) {
      float qf[8];
      unpack8(*reinterpret_cast<const uint4*>(qrow + v * 8), qf);
#pragma unroll
      for (int j = 0; j < F_MAX; ++j) {
        if (j < frames) {
          float kf[8];
          unpack8(*reinterpret_cast<const uint4*>(kbase + j * row_ld + v * 8), kf);
#pragma unroll
          for (int t = 0; t < 8; ++t) sc[j] = fmaf(qf[t], kf[t], sc[j]);
        }
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < F_MAX; ++j)
      if (j < frames) m = fmaxf(m, sc[j] * scale);
    float denom = 0.f;
#pragma unroll
    for (int j = 0; j < F_MAX; ++j) {
      sc[j] = j < frames ? __expf(sc[j] * scale - m) : 0.f;
      denom += sc[j];
    }
    const float inv = 1.f / denom;
    for (int v = 0; v < dv; ++v) {
      float acc[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = 0.f;
#pragma unroll
      for (int j = 0; j < F_MAX; ++j) {
        if (j < frames) {
          float vf[8];
          unpack8(*reinterpret_cast<const uint4*>(vbase + j * row_ld + v * 8), vf);
#pragma unroll
          for (int t = 0; t < 8; ++t) acc[t] = fmaf(sc[j], vf[t], acc[t]);
        }
      }
      uint4 out;
      __half2* oh = reinterpret_cast<__half2*>(&out);
#pragma unroll
      for (int t = 0; t < 4; ++t) oh[t] = __floats2half2_rn(acc[2 * t] * inv, acc[2 * t + 1] * inv);
      // the output row takes the place of this thread's own (already consumed) Q chunk
      *reinterpret_cast<uint4*>(qrow + v * 8) = out;
    }
  }
  __syncthreads();
  // ---- coalesced write-out of the seg-wide output rows ----
  const int total_out = frames * vec_per_seg;
  for (int i = threadIdx.x; i < total_out; i += blockDim.x) {
    const int v = i % vec_per_seg;
    const int f = i / vec_per_seg;
    const long long row = (clip * frames + f) * hw + pix;
    *reinterpret_cast<uint4*>(o + row * C + head0 * d + v * 8) = *reinterpret_cast<const uint4*>(s + f * row_ld + v * 8);
  }
}

}  // namespace ivv

extern "C" int ivv_temporal_attention(const void* qkv, void* o, int64_t clips, int64_t frames, int64_t hw, int64_t c,
                                      int32_t heads, float scale, ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(qkv && o && clips > 0 && frames > 0 && hw > 0 && c > 0 && heads > 0, "ivv_temporal_attention: bad args");
  IVV_REQUIRE(frames <= kMaxFrames, "ivv_temporal_attention: frames (%lld) must be <= 32", (long long)frames);
  IVV_REQUIRE(c % heads == 0 && (c / heads) % 8 == 0 && (c / heads) <= 256,
              "ivv_temporal_attention: head dim %lld must be a multiple of 8 and <= 256", (long long)(c / heads));
  const int d = (int)(c / heads);
  int hpc = heads;
  auto smem_for = [&](int h) { return (size_t)frames * (3 * h * d + kTPad) * sizeof(__half); };
  while (hpc > 1 && (smem_for(hpc) > 96 * 1024 || heads % hpc != 0 || hpc * frames > 1024)) --hpc;
  const size_t smem = smem_for(hpc);
  IVV_REQUIRE(smem <= 200 * 1024, "ivv_temporal_attention: tile does not fit shared memory");
  static bool configured = false;
  if (!configured) {
    IVV_CHECK_CUDA(
        cudaFuncSetAttribute(temporal_attn_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    IVV_CHECK_CUDA(
        cudaFuncSetAttribute(temporal_attn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const long long bp = clips * hw;
  IVV_REQUIRE(bp < (1LL << 31), "ivv_temporal_attention: too many sequences");
  dim3 grid((unsigned)bp, (unsigned)(heads / hpc));
  int threads = (int)(hpc * frames);
  threads = (threads + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  const __half* in = reinterpret_cast<const __half*>(qkv);
  __half* out = reinterpret_cast<__half*>(o);
  if (frames <= 16)
    temporal_attn_kernel<16><<<grid, threads, smem, stream>>>(in, out, (int)frames, hw, (int)c, heads, hpc, scale);
  else
    temporal_attn_kernel<32><<<grid, threads, smem, stream>>>(in, out, (int)frames, hw, (int)c, heads, hpc, scale);
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}
