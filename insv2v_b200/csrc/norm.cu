// GroupNorm(+SiLU) and LayerNorm(+positional encoding) for channels-last fp16 activations. HBM-bound kernels:
// 16-byte vector loads, fp32 per-thread partials, deterministic double-precision cross-CTA merge (statistics match
// the fp32 reference to ~1e-7). Reference call sites: ivv.h (K6/K7/K8).
#include <cstdlib>

#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

// ------------------------------------------------------------------------------------------------
// GroupNorm pass 1: per (batch-group, channel-group) mean / rstd.
//   x: [n_bg, rows_per_bg, C]; a thread owns one 8-channel vector column and strides over rows (4 loads in flight).
//   Deterministic: per-CTA partials are reduced in a fixed order inside the CTA, written to the workspace, and the
//   LAST CTA of each batch-group (atomic ticket) sums them in chunk order in double precision.
// workspace: [counters u32 x n_bg (>= 4 KB)] [final float2 x n_bg x G] [partials float2 x n_bg x max_chunks x G]
// The ticket counters wrap back to zero by themselves (atomicInc with the chunk count as the limit), so a workspace
// that was zero when first used stays usable call after call without a memset node in front of every norm.
// ------------------------------------------------------------------------------------------------
#ifndef IVV_GN_LOADS_STATS
#define IVV_GN_LOADS_STATS 8
#endif
#ifndef IVV_GN_LOADS_APPLY
#define IVV_GN_LOADS_APPLY 4  // 8 costs 108 registers (two CTAs per SM)
#endif
// 16-byte loads a thread keeps in flight (tuning: tools/build_variant.sh x -DIVV_GN_LOADS_APPLY=8)
constexpr int kGnLoads = IVV_GN_LOADS_STATS, kGnLoadsApply = IVV_GN_LOADS_APPLY;
struct GnWs {
  unsigned int* counters;
  float2* final_;   // (mean, rstd)
  float2* partial;  // (sum, sumsq)
  int max_chunks;
};
#ifndef IVV_GN_STATS_CPS
#define IVV_GN_STATS_CPS 4  // statistics CTAs per SM (tuning: tools/build_variant.sh x -DIVV_GN_STATS_CPS=2)
#endif
#ifndef IVV_GN_APPLY_CPS
#define IVV_GN_APPLY_CPS 8
#endif
__host__ __device__ inline int gn_max_chunks(long long n_bg) { return (int)(148 * IVV_GN_STATS_CPS / n_bg) + 2; }
constexpr long long kGnSelfCleanBytes = 4096;  // counter region the caller zero-fills once (n_bg <= 1024)
__host__ __device__ inline size_t gn_counter_bytes(long long n_bg) {
  const long long b = (n_bg * 4 + 255) / 256 * 256;
  return (size_t)(b < kGnSelfCleanBytes ? kGnSelfCleanBytes : b);
}

// In-CTA reduction of the per-thread partials s_part[R][C] (sum, sum of squares) to one pair per group, in a fixed order
// and spread over the CTA: every channel first sums its R row lanes (s_ch[C]), then a group is summed by one warp
// (lanes stride over the group's channels, shuffle tree) or, for narrow groups, by one thread. (One thread per group
// walking R x C/groups shared-memory values was a serial tail of up to ~3 000 clk per CTA.) Ends with a __syncthreads
// after the s_ch pass only; the caller orders the writes of `emit`.
template <class Emit>
__device__ __forceinline__ void gn_reduce_cta(const float2* s_part, float2* s_ch, int R, int C, int groups, Emit&& emit) {
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < R; ++rr) {
      const float2 v = s_part[rr * C + c];
      a += v.x;
      b += v.y;
    }
    s_ch[c] = make_float2(a, b);
  }
  __syncthreads();
  if (cpg >= 16) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (int)blockDim.x >> 5;  // full warps only (the block size need not be a multiple of 32)
    if (warp < nwarps) {
      for (int g = warp; g < groups; g += nwarps) {
        float a = 0.f, b = 0.f;
        for (int c = lane; c < cpg; c += 32) {
          const float2 v = s_ch[g * cpg + c];
          a += v.x;
          b += v.y;
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) emit(g, a, b);
      }
    }
  } else {
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      float a = 0.f, b = 0.f;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        a += s_ch[c].x;
        b += s_ch[c].y;
      }
      emit(g, a, b);
    }
  }
}

// Tail shared by the statistics kernels: per-CTA partials (fixed-order reduction of s_part[R][C]) -> workspace, ticket, and
// the merge by the last CTA of the batch group. Must be reached by every thread of the CTA.
__device__ __forceinline__ void gn_stats_finish(float* s_part, bool* s_last_p, const GnWs& ws, int bg, int chunks,
                                                long long rows_per_bg, int C, int groups, int R, float eps) {
  bool& s_last = *s_last_p;
  const int cpg = C / groups;
  __syncthreads();
  gn_reduce_cta(reinterpret_cast<const float2*>(s_part), reinterpret_cast<float2*>(s_part) + (size_t)R * C, R, C, groups,
                [&](int g, float a, float b) {
                  ws.partial[((long long)bg * ws.max_chunks + blockIdx.x) * groups + g] = make_float2(a, b);
                });
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(&ws.counters[bg], (unsigned int)chunks - 1);  // wraps to 0 on the last ticket
    s_last = (t == (unsigned int)chunks - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // Final merge by the last CTA, still in a fixed order (bit-reproducible), but spread over the whole CTA: L lanes per
  // group each sum every L-th chunk, then one thread per group adds the L lane sums in lane order. A single thread per
  // group walking ~200 chunks (5-D norms: 3 batch groups x 199 chunks) was a ~15 us serial tail of L2 round trips.
  const int L = (int)blockDim.x / groups;
  const double inv_n = 1.0 / ((double)rows_per_bg * cpg);
  auto finish = [&](int g, double a, double b) {
    const double mean = a * inv_n;
    double var = b * inv_n - mean * mean;
    if (var < 0) var = 0;
    ws.final_[(long long)bg * groups + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  };
  if (L >= 2) {
    double2* s_lane = reinterpret_cast<double2*>(s_part);  // [L][groups]; the row partials are no longer needed
    const int g = threadIdx.x % groups, l = threadIdx.x / groups;
    if (l < L) {
      double a = 0.0, b = 0.0;
#pragma unroll 4
      for (int ch = l; ch < chunks; ch += L) {
        const float2 pv = __ldcg(&ws.partial[((long long)bg * ws.max_chunks + ch) * groups + g]);
        a += (double)pv.x;
        b += (double)pv.y;
      }
      s_lane[l * groups + g] = make_double2(a, b);
    }
    __syncthreads();
    if (threadIdx.x < groups) {
      double a = 0.0, b = 0.0;
      for (int ll = 0; ll < L; ++ll) {
        a += s_lane[ll * groups + threadIdx.x].x;
        b += s_lane[ll * groups + threadIdx.x].y;
      }
      finish(threadIdx.x, a, b);
    }
    return;
  }
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
#pragma unroll 4
    for (int ch = 0; ch < chunks; ++ch) {
      const float2 pv = __ldcg(&ws.partial[((long long)bg * ws.max_chunks + ch) * groups + g]);
      a += (double)pv.x;
      b += (double)pv.y;
    }
    finish(g, a, b);
  }
}

// Two-source form (x2 != nullptr): channels [0, C1) come from x [.., C1], channels [C1, C) from x2 [.., C - C1] - the
// skip concatenation of the up blocks (unet_blocks.py:561,659) read in place instead of being materialised.
__global__ void gn_stats_kernel(const __half* __restrict__ x, const __half* __restrict__ x2, int C1, GnWs ws,
                                long long rows_per_bg, int C, int groups, long long rows_per_cta, int V, int R,
                                float eps) {
  extern __shared__ __align__(16) float s_part[];  // [R][C][2] | [C][2] (gn_reduce_cta)
  __shared__ bool s_last;
  griddep_sync();
  const int bg = blockIdx.y;
  const int chunks = gridDim.x;
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(rows_per_bg, row_begin + rows_per_cta);
  const int vec = threadIdx.x % V;
  const int rsub = threadIdx.x / V;
  const int cpg = C / groups;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  if (rsub < R) {
    const bool second = x2 != nullptr && vec * 8 >= C1;
    const long long ld = x2 == nullptr ? C : (second ? C - C1 : C1);
    const __half* base = (second ? x2 + (vec * 8 - C1) : x + vec * 8) + ((long long)bg * rows_per_bg) * ld;
    long long r = row_begin + rsub;
    for (; r + (long long)(kGnLoads - 1) * R < row_end; r += (long long)kGnLoads * R) {
      uint4 u[kGnLoads];
#pragma unroll
      for (int k = 0; k < kGnLoads; ++k) u[k] = *reinterpret_cast<const uint4*>(base + (r + (long long)k * R) * ld);
#pragma unroll
      for (int k = 0; k < kGnLoads; ++k) {
        const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          s[2 * j] += f.x;
          ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
          s[2 * j + 1] += f.y;
          ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
        }
      }
    }
    for (; r < row_end; r += R) {
      const uint4 u = *reinterpret_cast<const uint4*>(base + r * ld);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x;
        ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
        s[2 * j + 1] += f.y;
        ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_part[((rsub * C) + vec * 8 + j) * 2] = s[j];
      s_part[((rsub * C) + vec * 8 + j) * 2 + 1] = ss[j];
    }
  }
  gn_stats_finish(s_part, &s_last, ws, bg, chunks, rows_per_bg, C, groups, R, eps);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm pass 2: y = (x - mean) * rstd * gamma + beta, optional SiLU. Same thread->column mapping as pass 1:
// the 8 scale/shift pairs of a thread are loop invariants held in registers.
// ------------------------------------------------------------------------------------------------
__global__ void gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                                const __half* __restrict__ beta, const float2* __restrict__ final_,
                                long long rows_per_bg, int C, int groups, int act, long long rows_per_cta, int V,
                                int R, const __half* __restrict__ residual, const __half* __restrict__ x2, int C1) {
  griddep_sync();
  const int bg = blockIdx.y;
  const int vec = threadIdx.x % V;
  const int rsub = threadIdx.x / V;
  if (rsub >= R) return;
  const int cpg = C / groups;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vec * 8 + j;
    const float2 mr = final_[(long long)bg * groups + c / cpg];
    a[j] = mr.y * (gamma ? __half2float(gamma[c]) : 1.f);
    b[j] = (beta ? __half2float(beta[c]) : 0.f) - mr.x * a[j];
  }
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(rows_per_bg, row_begin + rows_per_cta);
  const bool second = x2 != nullptr && vec * 8 >= C1;
  const long long ld = x2 == nullptr ? C : (second ? C - C1 : C1);
  const __half* xb = (second ? x2 + (vec * 8 - C1) : x + vec * 8) + ((long long)bg * rows_per_bg) * ld;
  __half* yb = y + ((long long)bg * rows_per_bg) * C + vec * 8;
  // act: 0 none, 1 SiLU, 2 ReLU; residual (channel norm only): y = relu(residual + act(norm(x)))
  const __half* rbase = residual ? residual + ((long long)bg * rows_per_bg) * C + vec * 8 : nullptr;
  auto xform = [&](const uint4& u, long long r) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint4 o, ru = make_uint4(0, 0, 0, 0);
    if (rbase) ru = *reinterpret_cast<const uint4*>(rbase + r * C);
    const __half2* rh = reinterpret_cast<const __half2*>(&ru);
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      float v0 = fmaf(f.x, a[2 * j], b[2 * j]);
      float v1 = fmaf(f.y, a[2 * j + 1], b[2 * j + 1]);
      if (act == 1) {
        v0 = silu_f(v0);
        v1 = silu_f(v1);
      } else if (act == 2) {
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
      }
      if (rbase) {
        const float2 rf = __half22float2(rh[j]);
        v0 = fmaxf(v0 + rf.x, 0.f);
        v1 = fmaxf(v1 + rf.y, 0.f);
      }
      oh[j] = __floats2half2_rn(v0, v1);
    }
    return o;
  };
  long long r = row_begin + rsub;
  for (; r + (long long)(kGnLoadsApply - 1) * R < row_end; r += (long long)kGnLoadsApply * R) {
    uint4 u[kGnLoadsApply];
#pragma unroll
    for (int k = 0; k < kGnLoadsApply; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (r + (long long)k * R) * ld);
#pragma unroll
    for (int k = 0; k < kGnLoadsApply; ++k)
      *reinterpret_cast<uint4*>(yb + (r + (long long)k * R) * C) = xform(u[k], r + (long long)k * R);
  }
  for (; r < row_end; r += R)
    *reinterpret_cast<uint4*>(yb + r * C) = xform(*reinterpret_cast<const uint4*>(xb + r * ld), r);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm in ONE kernel, for tensors that fit in the shared memory of the chip (one CTA per SM, <= kGnFusedSmem each):
// the two-kernel form above is launch- and dependency-latency bound on the small levels of the UNet (4608 x 1280:
// 2 x 9.5 us for 11.8 MB, 18432 x 640: 2 x 14 us, profiles/r02_graph_timeline_final.txt). Here a CTA reads its rows ONCE,
// keeps them in shared memory, publishes its per-group partial sums (same workspace and ticket counter as
// gn_stats_kernel), WAITS until every CTA of its batch group has done so, merges the partials itself (every CTA in the
// same fixed order, double precision: identical statistics everywhere, no second hop through a "last CTA") and
// normalises out of shared memory.
// The wait: the ticket counter wraps to zero on the last arrival (atomicInc with limit chunks - 1), and a CTA that has
// arrived sees a non-zero counter until then - so "counter == 0 after my own arrival" means "all partials published",
// and the counter is back at zero for the next launch without a reset. All CTAs of the grid are co-resident by
// construction (grid <= SM count, one CTA per SM; a programmatic dependent launch only starts once every CTA here has
// started), so the spin cannot deadlock.
// ------------------------------------------------------------------------------------------------
constexpr int kGnFusedSmem = 208 * 1024;
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(512, 1)
gn_fused_kernel(const __half* __restrict__ x, const __half* __restrict__ x2, int C1, __half* __restrict__ y,
                const __half* __restrict__ gamma, const __half* __restrict__ beta, GnWs ws, long long rows_per_bg, int C,
                int groups, int rows_per_cta, int V, int R, float eps, int act) {
  extern __shared__ __align__(16) uint8_t gsm[];  // [rows_per_cta][V] uint4 | [R][C] float2 | [C] float2
  __shared__ float2 s_final[64];
  uint4* tile = reinterpret_cast<uint4*>(gsm);
  float2* s_part = reinterpret_cast<float2*>(gsm + (size_t)rows_per_cta * C * 2);
  float2* s_ch = s_part + (size_t)R * C;
  griddep_sync();
  const int bg = blockIdx.y;
  const int chunks = gridDim.x;
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  const int nrows = (int)(min(rows_per_bg, row_begin + rows_per_cta) - row_begin);
  const int vec = threadIdx.x % V;
  const int rsub = threadIdx.x / V;
  const int cpg = C / groups;
  const bool second = x2 != nullptr && vec * 8 >= C1;
  const long long ld = x2 == nullptr ? C : (second ? C - C1 : C1);
  const __half* base = (second ? x2 + (vec * 8 - C1) : x + vec * 8) + ((long long)bg * rows_per_bg + row_begin) * ld;
  // ---- pass 1: global -> shared memory + per-thread partial sums (thread = one 8-channel column, every R-th row) ----
  if (rsub < R) {
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    auto acc = [&](const uint4& u) {
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x;
        ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
        s[2 * j + 1] += f.y;
        ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
      }
    };
    int r = rsub;
    for (; r + 7 * R < nrows; r += 8 * R) {  // eight 16-byte loads in flight per thread
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = *reinterpret_cast<const uint4*>(base + (long long)(r + k * R) * ld);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        tile[(r + k * R) * V + vec] = u[k];
        acc(u[k]);
      }
    }
    for (; r < nrows; r += R) {
      const uint4 u = *reinterpret_cast<const uint4*>(base + (long long)r * ld);
      tile[r * V + vec] = u;
      acc(u);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[rsub * C + vec * 8 + j] = make_float2(s[j], ss[j]);
  }
  __syncthreads();
  // ---- in-CTA reduction (fixed order, spread over the CTA) -> this CTA's partials ----
  gn_reduce_cta(s_part, s_ch, R, C, groups, [&](int g, float a, float b) {
    ws.partial[((long long)bg * ws.max_chunks + blockIdx.x) * groups + g] = make_float2(a, b);
  });
  __threadfence();
  __syncthreads();
  // ---- arrive, then wait for the batch group (see the note above) ----
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(&ws.counters[bg], (unsigned int)chunks - 1);  // wraps to 0 on the last arrival
    if (t != (unsigned int)chunks - 1) {
      while (ld_acquire_u32(&ws.counters[bg]) != 0u) __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
  // ---- every CTA merges the partials of its batch group: fixed order, double precision ----
  {
    const int L = max(1, (int)blockDim.x / groups);
    double2* s_lane = reinterpret_cast<double2*>(s_part);  // [L][groups]; the row partials are no longer needed
    const int g = threadIdx.x % groups, l = threadIdx.x / groups;
    if (l < L) {
      double a = 0.0, b = 0.0;
#pragma unroll 4
      for (int ch = l; ch < chunks; ch += L) {
        const float2 pv = __ldcg(&ws.partial[((long long)bg * ws.max_chunks + ch) * groups + g]);
        a += (double)pv.x;
        b += (double)pv.y;
      }
      s_lane[l * groups + g] = make_double2(a, b);
    }
    __syncthreads();
    if (threadIdx.x < groups) {
      const double inv_n = 1.0 / ((double)rows_per_bg * cpg);
      double a = 0.0, b = 0.0;
      for (int ll = 0; ll < L; ++ll) {
        a += s_lane[ll * groups + threadIdx.x].x;
        b += s_lane[ll * groups + threadIdx.x].y;
      }
      const double mean = a * inv_n;
      double var = b * inv_n - mean * mean;
      if (var < 0) var = 0;
      s_final[threadIdx.x] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    __syncthreads();
  }
  // ---- pass 2: normalise out of shared memory (each thread re-reads exactly what it wrote) ----
  if (rsub >= R) return;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vec * 8 + j;
    const float2 mr = s_final[c / cpg];
    a[j] = mr.y * __half2float(gamma[c]);
    b[j] = __half2float(beta[c]) - mr.x * a[j];
  }
  __half* yb = y + ((long long)bg * rows_per_bg + row_begin) * C + vec * 8;
  for (int r = rsub; r < nrows; r += R) {
    const uint4 u = tile[r * V + vec];
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      float v0 = fmaf(f.x, a[2 * j], b[2 * j]);
      float v1 = fmaf(f.y, a[2 * j + 1], b[2 * j + 1]);
      if (act == 1) {
        v0 = silu_f(v0);
        v1 = silu_f(v1);
      } else if (act == 2) {
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
      }
      oh[j] = __floats2half2_rn(v0, v1);
    }
    *reinterpret_cast<uint4*>(yb + (long long)r * C) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (two-pass variance), optional + pe[frame]
// ------------------------------------------------------------------------------------------------
template <int LPR, int MAXV>
__global__ void layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                 const __half* __restrict__ gamma, const __half* __restrict__ beta, long long rows,
                                 int C, float eps, const float* __restrict__ pe, long long rows_per_frame,
                                 long long frames, long long pe_start) {
  // LPR lanes cooperate on one row (32/LPR rows per warp): short rows (C = 320) keep every lane busy
  constexpr int RPW = 32 / LPR;
  griddep_sync();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool active = row < rows;
  const int V = C / 8;
  float v[MAXV][8];
  float sum = 0.f;
  const __half* xr = x + (active ? row : 0) * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vec = sub + i * LPR;
    if (vec < V && active) {
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
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vec = sub + i * LPR;
    if (vec < V && active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (!active) return;
  const float rstd = rsqrtf(sq / (float)C + eps);
  const float* per = nullptr;
  if (pe) {
    const long long frame = (row / rows_per_frame) % frames;
    per = pe + (pe_start + frame) * C;
  }
  __half* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vec = sub + i * LPR;
    if (vec < V) {
      const uint4 ug = *reinterpret_cast<const uint4*>(gamma + vec * 8);
      const uint4 ub = *reinterpret_cast<const uint4*>(beta + vec * 8);
      const __half2* hg = reinterpret_cast<const __half2*>(&ug);
      const __half2* hb = reinterpret_cast<const __half2*>(&ub);
      float pv[8];
      if (per) {
        const float4 p0 = *reinterpret_cast<const float4*>(per + vec * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(per + vec * 8 + 4);
        pv[0] = p0.x, pv[1] = p0.y, pv[2] = p0.z, pv[3] = p0.w, pv[4] = p1.x, pv[5] = p1.y, pv[6] = p1.z, pv[7] = p1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) pv[j] = 0.f;
      }
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 g = __half22float2(hg[j]);
        const float2 b = __half22float2(hb[j]);
        const float o0 = fmaf((v[i][2 * j] - mean) * rstd, g.x, b.x) + pv[2 * j];
        const float o1 = fmaf((v[i][2 * j + 1] - mean) * rstd, g.y, b.y) + pv[2 * j + 1];
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
  griddep_sync();
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
  if (frames_per_group <= 0 || groups <= 0) return 0;
  const long long n_bg = n_img / frames_per_group;
  if (n_bg <= 0) return 0;
  return ivv::gn_counter_bytes(n_bg) + (size_t)n_bg * groups * sizeof(float2) +
         (size_t)n_bg * ivv::gn_max_chunks(n_bg) * groups * sizeof(float2);
}

namespace ivv {
// IVV_GN_FUSED=0 keeps the two-kernel GroupNorm everywhere (tuning hook; read once)
static bool gn_fused_enabled() {
  static const bool on = [] {
    const char* v = getenv("IVV_GN_FUSED");
    return v == nullptr || atoi(v) != 0;
  }();
  return on;
}
static int gn_sm_count() {  // of the current device (cached per device ordinal)
  static int n[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (n[dev] == 0 && cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 0;
  return n[dev];
}
// one-kernel GroupNorm (gn_fused_kernel): at most one CTA per SM, each holding its rows in shared memory
struct GnFusedPlan {
  int chunks, rows_per_cta, R;
  size_t smem;
};
static bool gn_fused_plan(long long n_bg, long long rows_per_bg, long long c, int groups, GnFusedPlan* out) {
  if (!gn_fused_enabled() || groups > 64 || c % 8 != 0) return false;
  const int sms = gn_sm_count();
  const int V = (int)(c / 8);
  const int R = V >= 512 ? 1 : 512 / V;
  if (n_bg <= 0 || n_bg > sms || V * R > 512 || V * R < groups) return false;
  const long long chunks = sms / n_bg;  // one CTA per SM at most: the CTAs wait for each other
  const long long rpc = (rows_per_bg + chunks - 1) / chunks;
  const long long nch = (rows_per_bg + rpc - 1) / rpc;
  const size_t smem = (size_t)rpc * c * 2 + (size_t)(R + 1) * c * 2 * sizeof(float);
  if (smem > (size_t)kGnFusedSmem || rpc >= (1LL << 20)) return false;
  out->chunks = (int)nch;
  out->rows_per_cta = (int)rpc;
  out->R = R;
  out->smem = smem;
  return true;
}
static int groupnorm_impl(const void* x, void* y, const void* gamma, const void* beta, int64_t n_img, int64_t hw,
                          int64_t c, int32_t groups, int64_t frames_per_group, float eps, int32_t act,
                          const void* residual, void* stats_ws, size_t stats_ws_bytes, int max_groups,
                          cudaStream_t stream, const void* x2 = nullptr, int64_t c1 = 0, bool allow_fused = false) {
  IVV_REQUIRE(x2 == nullptr || (c1 > 0 && c1 < c && c1 % 8 == 0), "ivv_groupnorm2: c1 (%lld) must be a multiple of 8 in (0, c)",
              (long long)c1);
  IVV_REQUIRE(x && y && stats_ws, "ivv_groupnorm: null pointer");
  IVV_REQUIRE(n_img > 0 && hw > 0 && c > 0, "ivv_groupnorm: empty input");
  IVV_REQUIRE(frames_per_group > 0 && n_img % frames_per_group == 0,
              "ivv_groupnorm: n_img (%lld) must be a multiple of frames_per_group (%lld)", (long long)n_img,
              (long long)frames_per_group);
  IVV_REQUIRE(groups > 0 && groups <= max_groups && c % groups == 0, "ivv_groupnorm: bad groups %d for c=%lld", groups,
              (long long)c);
  IVV_REQUIRE(c % 8 == 0 && c <= 8192, "ivv_groupnorm: c (%lld) must be a multiple of 8 and <= 8192", (long long)c);
  const size_t need = ivv_groupnorm_ws_bytes(n_img, groups, frames_per_group);
  IVV_REQUIRE(stats_ws_bytes >= need, "ivv_groupnorm: workspace too small (%zu < %zu)", stats_ws_bytes, need);
  IVV_REQUIRE((reinterpret_cast<uintptr_t>(stats_ws) & 15) == 0, "ivv_groupnorm: workspace must be 16-byte aligned");
  const long long n_bg = n_img / frames_per_group;
  const long long rows_per_bg = frames_per_group * hw;
  IVV_REQUIRE(n_bg <= 65535, "ivv_groupnorm: too many batch groups");
  GnWs ws;
  ws.counters = reinterpret_cast<unsigned int*>(stats_ws);
  ws.final_ = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(stats_ws) + gn_counter_bytes(n_bg));
  ws.partial = ws.final_ + n_bg * groups;
  ws.max_chunks = gn_max_chunks(n_bg);
  // more batch groups than the self-cleaning region covers: counters beyond it may alias older final/partial data
  if (n_bg * 4 > kGnSelfCleanBytes) IVV_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, gn_counter_bytes(n_bg), stream));
  const int V = (int)(c / 8);
  // ---- one-kernel form: the whole tensor in the shared memory of the chip (see gn_fused_kernel) ----
  GnFusedPlan fp;
  if (allow_fused && residual == nullptr && gamma != nullptr && beta != nullptr && gn_fused_plan(n_bg, rows_per_bg, c, groups, &fp) &&
      fp.chunks <= ws.max_chunks) {
    static DeviceOnce configured;
    if (configured.first())
      IVV_CHECK_CUDA(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGnFusedSmem));
    IVV_CHECK_CUDA(launch_pdl(gn_fused_kernel, dim3((unsigned)fp.chunks, (unsigned)n_bg), dim3((unsigned)(V * fp.R)),
                              fp.smem, stream, reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(x2),
                              (int)c1, reinterpret_cast<__half*>(y), reinterpret_cast<const __half*>(gamma),
                              reinterpret_cast<const __half*>(beta), ws, rows_per_bg, (int)c, groups, fp.rows_per_cta, V,
                              fp.R, eps, (int)act));
    return 0;
  }
  const int R = V >= 256 ? 1 : 256 / V;
  const int threads = V * R;
  {
    // ~4 CTAs per SM in total, at least 8 row sweeps per CTA, never more chunks than the workspace holds
    long long chunks = (148 * IVV_GN_STATS_CPS + n_bg - 1) / n_bg;
    long long rows_per_cta = (rows_per_bg + chunks - 1) / chunks;
    if (rows_per_cta < 8LL * R) rows_per_cta = 8LL * R;
    chunks = (rows_per_bg + rows_per_cta - 1) / rows_per_cta;
    IVV_REQUIRE(chunks <= ws.max_chunks, "ivv_groupnorm: internal chunking error");
    dim3 grid((unsigned)chunks, (unsigned)n_bg);
    const size_t smem = (size_t)(R + 1) * c * 2 * sizeof(float);
    static DeviceOnce configured;  // per device (cudaFuncSetAttribute is a per-device attribute)
    if (smem > 48 * 1024 && configured.first())
      IVV_CHECK_CUDA(cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 144 * 1024));
    IVV_CHECK_CUDA(launch_pdl(gn_stats_kernel, grid, dim3(threads), smem, stream, reinterpret_cast<const __half*>(x),
                              reinterpret_cast<const __half*>(x2), (int)c1, ws, rows_per_bg, (int)c, groups,
                              rows_per_cta, V, R, eps));
  }
  {
    long long chunks2 = (148 * IVV_GN_APPLY_CPS + n_bg - 1) / n_bg;
    long long rpc = (rows_per_bg + chunks2 - 1) / chunks2;
    if (rpc < 4LL * R) rpc = 4LL * R;
    chunks2 = (rows_per_bg + rpc - 1) / rpc;
    dim3 grid((unsigned)chunks2, (unsigned)n_bg);
    IVV_CHECK_CUDA(launch_pdl(gn_apply_kernel, grid, dim3(threads), 0, stream, reinterpret_cast<const __half*>(x),
                              reinterpret_cast<__half*>(y), reinterpret_cast<const __half*>(gamma),
                              reinterpret_cast<const __half*>(beta), ws.final_, rows_per_bg, (int)c, groups, (int)act,
                              rpc, V, R, reinterpret_cast<const __half*>(residual),
                              reinterpret_cast<const __half*>(x2), (int)c1));
  }
  return 0;
}
}  // namespace ivv

// 1 if ivv_groupnorm / ivv_groupnorm2 on this shape run as ONE kernel (launch accounting of the host side)
extern "C" int ivv_groupnorm_is_fused(int64_t n_img, int64_t hw, int64_t c, int32_t groups, int64_t frames_per_group) {
  if (frames_per_group <= 0 || n_img <= 0 || n_img % frames_per_group != 0 || groups <= 0) return 0;
  const long long n_bg = n_img / frames_per_group;
  ivv::GnFusedPlan fp;
  return ivv::gn_fused_plan(n_bg, frames_per_group * hw, c, groups, &fp) && fp.chunks <= ivv::gn_max_chunks(n_bg) ? 1 : 0;
}

extern "C" int ivv_groupnorm2(const void* x1, int64_t c1, const void* x2, int64_t c2, void* y, const void* gamma,
                              const void* beta, int64_t n_img, int64_t hw, int32_t groups, int64_t frames_per_group,
                              float eps, int32_t silu, void* stats_ws, size_t stats_ws_bytes, ivv_stream_t stream_) {
  IVV_REQUIRE(gamma && beta && x2, "ivv_groupnorm2: null pointer");
  return ivv::groupnorm_impl(x1, y, gamma, beta, n_img, hw, c1 + c2, groups, frames_per_group, eps, silu ? 1 : 0,
                             nullptr, stats_ws, stats_ws_bytes, 64, reinterpret_cast<cudaStream_t>(stream_), x2, c1, true);
}

extern "C" int ivv_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int64_t n_img, int64_t hw,
                             int64_t c, int32_t groups, int64_t frames_per_group, float eps, int32_t silu,
                             void* stats_ws, size_t stats_ws_bytes, ivv_stream_t stream_) {
  IVV_REQUIRE(gamma && beta, "ivv_groupnorm: null pointer");
  return ivv::groupnorm_impl(x, y, gamma, beta, n_img, hw, c, groups, frames_per_group, eps, silu ? 1 : 0, nullptr,
                             stats_ws, stats_ws_bytes, 64, reinterpret_cast<cudaStream_t>(stream_), nullptr, 0, true);
}

// Per-channel normalisation over imgs_per_group images: InstanceNorm2d (imgs_per_group = 1) and BatchNorm2d with
// batch statistics (imgs_per_group = n_img; the reference never puts RAFTFlow in eval mode) of torchvision's RAFT
// encoders, + ReLU, + the residual join relu(residual + y) of ResidualBlock.forward (raft.py:63-71).
extern "C" size_t ivv_channelnorm_ws_bytes(int64_t n_img, int64_t c, int64_t imgs_per_group) {
  return ivv_groupnorm_ws_bytes(n_img, (int32_t)c, imgs_per_group);
}
extern "C" int ivv_channelnorm(const void* x, void* y, const void* gamma, const void* beta, int64_t n_img, int64_t hw,
                               int64_t c, int64_t imgs_per_group, float eps, int32_t relu, const void* residual,
                               void* stats_ws, size_t stats_ws_bytes, ivv_stream_t stream_) {
  IVV_REQUIRE(c <= 1024, "ivv_channelnorm: c (%lld) must be <= 1024", (long long)c);
  return ivv::groupnorm_impl(x, y, gamma, beta, n_img, hw, c, (int32_t)c, imgs_per_group, eps, relu ? 2 : 0, residual,
                             stats_ws, stats_ws_bytes, 1024, reinterpret_cast<cudaStream_t>(stream_));
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
  const __half* xx = reinterpret_cast<const __half*>(x);
  __half* yy = reinterpret_cast<__half*>(y);
  const __half* g = reinterpret_cast<const __half*>(gamma);
  const __half* b = reinterpret_cast<const __half*>(beta);
  IVV_REQUIRE(!pe || (reinterpret_cast<uintptr_t>(pe) & 15) == 0, "ivv_layernorm: pe must be 16-byte aligned");
  const int V = (int)(c / 8);
#define IVV_LN_LAUNCH(LPR, MAXV)                                                                                   \
  {                                                                                                                \
    const long long rows_per_block = (long long)warps * (32 / LPR);                                                \
    const long long blocks = (rows + rows_per_block - 1) / rows_per_block;                                         \
    IVV_CHECK_CUDA(launch_pdl(layernorm_kernel<LPR, MAXV>, dim3((unsigned)blocks), dim3(warps * 32), 0, stream,   \
                              xx, yy, g, b, rows, (int)c, eps, pe, rows_per_frame, frames, pe_start));            \
  }
  if (V <= 8 * 5) IVV_LN_LAUNCH(8, 5)
  else if (V <= 16 * 5) IVV_LN_LAUNCH(16, 5)
  else if (V <= 32 * 5) IVV_LN_LAUNCH(32, 5)
  else IVV_LN_LAUNCH(32, 8)
#undef IVV_LN_LAUNCH
  IVV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ivv_softmax_rows(const void* x, int32_t x_is_f32, void* y, int64_t rows, int64_t cols, float scale,
                                ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(x && y && rows > 0 && cols > 0, "ivv_softmax_rows: bad arguments");
  if (x_is_f32)
    IVV_CHECK_CUDA(launch_pdl(softmax_rows_kernel<float>, dim3((unsigned)rows), dim3(256), 0, stream,
                              reinterpret_cast<const float*>(x), reinterpret_cast<__half*>(y), rows, (int)cols, scale));
  else
    IVV_CHECK_CUDA(launch_pdl(softmax_rows_kernel<__half>, dim3((unsigned)rows), dim3(256), 0, stream,
                              reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), rows, (int)cols, scale));
  return 0;
}
