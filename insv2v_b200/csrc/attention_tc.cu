// Fused flash-attention forward for sm_100a: softmax(Q K^T * scale) V with tcgen05.mma, accumulators in TMEM,
// operands staged by TMA straight out of the projection buffers (no head_to_batch copies: the head is a coordinate of
// the 4-D tensor map (d, head, token, frame); head dims 40/80/160 are zero-padded to 64-wide chunks by TMA OOB fill).
// Spatial self-attention and cross-attention of the InsV2V UNet (reference call sites: ivv.h, K3/K4).
//
// One CTA = 128 query rows of one (frame, head). Warps 0-3: softmax (one row per thread, fp32 statistics, online
// rescale of the TMEM-resident O), warp 4: TMA producer, warp 5: MMA issuer + TMEM owner.
// Per 128-key block:  S = Q K^T (TMEM cols [0,128))  ->  P = exp2(S*sl - m*sl) as fp16 in swizzled smem
//                     ->  O += P V (TMEM cols [128, 128+dn)).
// TMEM use is 256 columns for d <= 111 so two CTAs share an SM and one CTA's softmax overlaps the other's MMAs.
#include <cstdlib>

#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

constexpr int kQ = 128;        // query rows per CTA
constexpr int kKV = 128;       // keys per block
constexpr int kSlab = 128 * 128;  // bytes of one [128 rows x 64 fp16] swizzled slab
constexpr int kAttnThreads = 192;

struct AttnParams {
  int s_q, s_kv, kv_div, d;
  float scale_log2;  // scale * log2(e)
  __half* o;
  long long o_ld;
  long long* trace;  // tuning only (ivv_debug_attn_trace): clock64 stamps of the first cluster's leader, [block][16]
  int qk_first;  // one-tile kernel: issue S(j+1) = Q K(j+1)^T before O += P(j) V(j) (shortens the softmax -> softmax chain)
};

template <int DC, int NS>
constexpr int attn_smem_bytes() {
  return (DC + 2 * NS * DC + 2) * kSlab + 256;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax 2^f
// (max relative error 7.5e-5, a third of the fp16 half-ulp of P), exponent added to the bit pattern.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the integer n lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float q = fmaf(5.517166492e-02f, f, 2.426111219e-01f);
  q = fmaf(q, f, 6.932609862e-01f);
  q = fmaf(q, f, 9.999280736e-01f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t ex2_poly_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(ex2_poly(lo), ex2_poly(hi));
  return *reinterpret_cast<uint32_t*>(&h);
}
// exp2 of two scores -> packed fp16 pair. (ex2.approx.f16x2 was tried: on sm_100a it is executed as two MUFU ops plus
// byte permutes — more instructions than two fp32 ex2 and one pack — so the fp32 form is used.)
__device__ __forceinline__ uint32_t ex2_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(ex2(lo), ex2(hi));
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------------
// The softmax / correction / output role of one 128-row query tile (4 warps, thread = query row). Shared by the
// one-tile kernel (two CTAs per SM) and the two-tile kernel (two softmax groups per CTA sharing every K/V load).
// ---------------------------------------------------------------------------------------------------------------
// PAIR: the tile belongs to one CTA of a cta_group::2 pair. The V stage of this CTA then holds only 32 of the 64 value
// columns (ones column at local column d % 32, in the CTA that owns it), and the "P is ready" signal is one arrive per
// warp on the LEADER's barrier (p_full_remote, a shared::cluster address). POLY: every fourth packed pair of
// exponentials is computed on the FMA pipe (ex2_poly) instead of the MUFU, which is the unit this kernel is bound by.
template <int DC, int NS, bool PAIR = false, int POLY = 0>
__device__ __forceinline__ void softmax_tile(const AttnParams& p, int r, uint32_t lane_off, uint32_t tmem_S,
                                             uint32_t tmem_O, uint8_t* sP, uint8_t* sV, uint64_t* s_full,
                                             uint64_t* p_full, uint64_t* pv_done, int q0, int head, int nb, int nblk,
                                             int dn, uint32_t p_full_remote = 0, int crank = 0, int g0 = 0,
                                             int* st_io = nullptr) {
  // g0 / st_io (persistent kernels): the tile's first key block is block g0 of the CTA's flat block sequence (barrier
  // phases continue across tiles), and the K/V ring stage is carried from tile to tile through *st_io
  const float sl = p.scale_log2;
  float m_run = -INFINITY;
  // The softmax denominator is never summed on the CUDA cores: a column of ones is written into the V tile at
  // column d (slab d/64, 16-byte chunk (d%64)/8, element d%8), so O[:, d] accumulates sum(P) in fp32 inside the
  // tensor core and follows every online rescale for free.
  const int one_col = PAIR ? (p.d & 31) : (p.d & 63);
  const int one_slab = PAIR ? 0 : (p.d >> 6), one_chunk = one_col >> 3, one_elem = p.d & 7;
  const bool write_ones = PAIR ? ((p.d >> 5) == crank) : true;
  int st = st_io != nullptr ? *st_io : 0;
  for (int j = 0; j < nblk; ++j) {
    const int valid = min(kKV, p.s_kv - j * kKV);
    const int nchunk = (valid + 31) / 32;
    mbar_wait(s_full, (g0 + j) & 1);
    tc_fence_after();
    // pass 1: row max (full blocks take the mask-free path: the softmax warps are instruction-issue bound)
    float mx = -INFINITY;
    if (valid == kKV) {
      // two TMEM loads in flight per wait: the softmax warps are bound by exposed tcgen05.ld latency otherwise
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        uint32_t va[32], vb[32];
        tmem_ld32(tmem_S + lane_off + c * 32, va);
        tmem_ld32(tmem_S + lane_off + c * 32 + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
      }
    } else {
      for (int c = 0; c < nchunk; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
    }
    // Lazy rescale: keep the running reference m_run while the new block maximum exceeds it by less than a factor
    // 2^2 (exp2 arguments stay below 2, where the fp16 argument grid is still 2^-10: P keeps ~fp16 accuracy; O and
    // the denominator column accumulate in fp32). O is then rescaled only when a row's maximum jumps, which after the
    // first blocks is rare, instead of on every block.
    float m_new = m_run;
    if ((mx - m_run) * sl > 2.f) m_new = mx;  // also taken on the first block (m_run = -inf)
    const float alpha = ex2((m_run - m_new) * sl);
    const float m_sl = m_new * sl;
    if (j > 0) {
      // previous P V must have retired before O is rescaled and P is overwritten
      mbar_wait(pv_done, (g0 + j - 1) & 1);
      tc_fence_after();
      if (!__all_sync(0xffffffffu, m_new == m_run)) {
        for (int c = 0; c < dn; c += 16) {
          uint32_t o[16];
          tmem_ld16(tmem_O + lane_off + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tmem_O + lane_off + c, o);
        }
        tmem_st_wait();
      }
    }
    // pass 2: P = exp2(S*sl - m*sl) -> fp16 pairs, K-major SW128 smem (row r, 16-byte chunk cc ^ (r & 7))
    if (valid == kKV) {
      // software pipeline: the load of chunk c+1 is in flight while chunk c is exponentiated and stored
      auto emit = [&](const uint32_t (&v)[32], int c) {
        uint8_t* slab = sP + (c >> 1) * kSlab + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float x0 = fmaf(__uint_as_float(v[g * 8 + 2 * t]), sl, -m_sl);
            const float x1 = fmaf(__uint_as_float(v[g * 8 + 2 * t + 1]), sl, -m_sl);
            pk[t] = (POLY != 0 && t == 3) ? ex2_poly_h2(x0, x1) : ex2_h2(x0, x1);
          }
          const int cc = (c & 1) * 4 + g;
          *reinterpret_cast<uint4*>(slab + ((cc ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      };
      uint32_t va[32], vb[32];
      tmem_ld32(tmem_S + lane_off, va);
      tmem_ld_wait();
      tmem_ld32(tmem_S + lane_off + 32, vb);
      emit(va, 0);
      tmem_ld_wait();
      tmem_ld32(tmem_S + lane_off + 64, va);
      emit(vb, 1);
      tmem_ld_wait();
      tmem_ld32(tmem_S + lane_off + 96, vb);
      emit(va, 2);
      tmem_ld_wait();
      emit(vb, 3);
    } else {
      for (int c = 0; c < nchunk; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c * 32, v);
        tmem_ld_wait();
        uint8_t* slab = sP + (c >> 1) * kSlab + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int i0 = g * 8 + 2 * t;
            const float x0 = (c * 32 + i0 < valid) ? fmaf(__uint_as_float(v[i0]), sl, -m_sl) : -INFINITY;
            const float x1 = (c * 32 + i0 + 1 < valid) ? fmaf(__uint_as_float(v[i0 + 1]), sl, -m_sl) : -INFINITY;
            pk[t] = ex2_h2(x0, x1);
          }
          const int cc = (c & 1) * 4 + g;  // 16-byte chunk index inside the 128-byte row
          *reinterpret_cast<uint4*>(slab + ((cc ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    m_run = m_new;
    if (write_ones) {  // ones column of this V stage, key row r
      uint8_t* vrow = sV + (st * DC + one_slab) * kSlab + r * 128;
      *reinterpret_cast<__half*>(vrow + ((one_chunk ^ (r & 7)) << 4) + one_elem * 2) = __float2half_rn(1.f);
    }
    if (++st == NS) st = 0;
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    if constexpr (PAIR) {
      __syncwarp();
      if ((r & 31) == 0) mbar_arrive_cluster(p_full_remote);
    } else {
      mbar_arrive(p_full);
    }
  }
  // ---- epilogue: O / l -> global ----
  if (st_io != nullptr) *st_io = st;
  mbar_wait(pv_done, (g0 + nblk - 1) & 1);
  tc_fence_after();
  float inv_l;
  {  // the denominator is the accumulator column d (one-column load: no register array to index at run time)
    const uint32_t lbits = tmem_ld1(tmem_O + lane_off + p.d);
    tmem_ld_wait();
    inv_l = 1.f / __uint_as_float(lbits);
  }
  const int row = q0 + r;
  __half* orow = p.o + (static_cast<long long>(nb) * p.s_q + row) * p.o_ld + static_cast<long long>(head) * p.d;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);
  for (int c = 0; c < p.d; c += 16) {
    uint32_t o[16];
    tmem_ld16(tmem_O + lane_off + c, o);
    tmem_ld_wait();
    if (row < p.s_q) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = c + g * 8;
        if (col + 8 <= p.d && vec_ok) {
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __half2 h = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * t]) * inv_l,
                                          __uint_as_float(o[g * 8 + 2 * t + 1]) * inv_l);
            pk[t] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(orow + col) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)  // unrolled: a runtime index would move o[] to local memory
            if (col + i < p.d) orow[col + i] = __float2half_rn(__uint_as_float(o[g * 8 + i]) * inv_l);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass softmax role (pair kernel). The two-pass role above reads every S block from tensor memory twice (row
// maximum, then exponentials) and tcgen05.ld is the unit it saturates: 2 x 64 KB per 128x128 block. Here a block is
// read ONCE: the exponentials are taken against the running reference m_run while the block maximum is tracked on the
// side, the packed fp16 probabilities stay in registers (64 per thread), and only when some row's maximum rose by more
// than 2^8 over its reference (first block, or a rare jump) does the warp redo the block the two-pass way. Because P
// waits in registers, the S buffer is handed back ("s_free") before the previous P V has even retired, so the tensor
// core computes S(j+1) under the tail of softmax(j) and the wait for P V(j-1) is normally already satisfied.
// ---------------------------------------------------------------------------------------------------------------
// DBG (timing experiments only, results are garbage): 1 = no MUFU (exp2 replaced by an FMA), 2 = no P stores / proxy
// fence, 3 = no tcgen05.ld of S (constants), 4 = no max tracking
template <int NS, int POLY, int DBG = 0>
__device__ __forceinline__ void softmax_tile_sp(const AttnParams& p, int r, uint32_t lane_off, uint32_t tmem_S,
                                                uint32_t tmem_O, uint8_t* sP, uint8_t* sV, uint64_t* s_full,
                                                uint64_t* pv_done, uint32_t s_free_remote, uint32_t p_full_remote,
                                                int q0, int head, int nb, int nblk, int dn, int crank) {
  const float sl = p.scale_log2;
  float m_run = -INFINITY;
  const int one_col = p.d & 31, one_chunk = one_col >> 3, one_elem = p.d & 7;
  const bool write_ones = (p.d >> 5) == crank;
  const bool lane0 = (r & 31) == 0;
  int st = 0;
  uint32_t pk[64];  // P row of this thread, fp16 pairs (keys 2i, 2i+1)

  auto exp_chunk = [&](const uint32_t (&v)[32], int c, float m_sl) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x0 = fmaf(__uint_as_float(v[2 * i]), sl, -m_sl);
      const float x1 = fmaf(__uint_as_float(v[2 * i + 1]), sl, -m_sl);
      if constexpr (DBG == 1) {
        __half2 h = __floats2half2_rn(fmaf(x0, 0.5f, 1.f), fmaf(x1, 0.5f, 1.f));
        pk[c * 16 + i] = *reinterpret_cast<uint32_t*>(&h);
      } else {
        pk[c * 16 + i] = (POLY != 0 && (i & 3) == 3) ? ex2_poly_h2(x0, x1) : ex2_h2(x0, x1);
      }
    }
  };
  auto max_chunk = [&](const uint32_t (&v)[32], float mx) {
    if constexpr (DBG == 4) return fmaxf(mx, __uint_as_float(v[0]));
#pragma unroll
    for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
    return mx;
  };

  for (int j = 0; j < nblk; ++j) {
    const int valid = min(kKV, p.s_kv - j * kKV);
    mbar_wait(s_full, j & 1);
    tc_fence_after();
    float m_new = m_run;
    bool redo = true;
    if (valid == kKV && j > 0) {
      // ---- optimistic single pass against m_run ----
      const float m_sl = m_run * sl;
      float mx = -INFINITY;
      uint32_t va[32], vb[32];
      auto ld = [&](int c, uint32_t (&v)[32]) {
        if constexpr (DBG == 3) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(m_run - 0.01f * (i + c));
        } else {
          tmem_ld32(tmem_S + lane_off + c * 32, v);
        }
      };
      ld(0, va);
      tmem_ld_wait();
      ld(1, vb);
      mx = max_chunk(va, mx);
      exp_chunk(va, 0, m_sl);
      tmem_ld_wait();
      ld(2, va);
      mx = max_chunk(vb, mx);
      exp_chunk(vb, 1, m_sl);
      tmem_ld_wait();
      ld(3, vb);
      mx = max_chunk(va, mx);
      exp_chunk(va, 2, m_sl);
      tmem_ld_wait();
      mx = max_chunk(vb, mx);
      redo = __any_sync(0xffffffffu, (mx - m_run) * sl > 8.f);
      if (!redo) {
        tc_fence_before();
        __syncwarp();
        if (lane0) mbar_arrive_cluster(s_free_remote);  // S(j) is in registers: the tensor core may overwrite it
        exp_chunk(vb, 3, m_sl);
      }
    }
    if (redo) {
      // ---- two-pass: first block, masked last block, or a row maximum that jumped ----
      const int nchunk = (valid + 31) / 32;
      float mx = -INFINITY;
      for (int c = 0; c < nchunk; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      if ((mx - m_run) * sl > 2.f) m_new = mx;  // also taken on the first block (m_run = -inf)
      const float m_sl = m_new * sl;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunk) {
          uint32_t v[32];
          tmem_ld32(tmem_S + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k0 = c * 32 + 2 * i;
            const float x0 = (k0 < valid) ? fmaf(__uint_as_float(v[2 * i]), sl, -m_sl) : -INFINITY;
            const float x1 = (k0 + 1 < valid) ? fmaf(__uint_as_float(v[2 * i + 1]), sl, -m_sl) : -INFINITY;
            pk[c * 16 + i] = ex2_h2(x0, x1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[c * 16 + i] = 0u;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive_cluster(s_free_remote);
    }
    // ---- P V(j-1) must have retired before O is rescaled and the P buffer is overwritten ----
    if (j > 0) {
      mbar_wait(pv_done, (j - 1) & 1);
      tc_fence_after();
      if (!__all_sync(0xffffffffu, m_new == m_run)) {
        const float alpha = ex2((m_run - m_new) * sl);
        for (int c = 0; c < dn; c += 16) {
          uint32_t o[16];
          tmem_ld16(tmem_O + lane_off + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tmem_O + lane_off + c, o);
        }
        tmem_st_wait();
      }
    }
    m_run = m_new;
    // ---- P -> K-major SW128 smem (row r, 16-byte chunk cc ^ (r & 7) of slab cc / 8) ----
    if constexpr (DBG == 2) {
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < 64; ++i) acc ^= pk[i];
      if (acc == 0x12345678u) *reinterpret_cast<uint32_t*>(sP + r * 128) = acc;  // keep pk alive
    }
#pragma unroll
    for (int cc = 0; DBG != 2 && cc < 16; ++cc) {
      uint8_t* slab = sP + (cc >> 3) * kSlab + r * 128;
      *reinterpret_cast<uint4*>(slab + (((cc & 7) ^ (r & 7)) << 4)) =
          make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
    }
    if (write_ones) {  // ones column of this V stage, key row r
      uint8_t* vrow = sV + st * kSlab + r * 128;
      *reinterpret_cast<__half*>(vrow + ((one_chunk ^ (r & 7)) << 4) + one_elem * 2) = __float2half_rn(1.f);
    }
    if (++st == NS) st = 0;
    if constexpr (DBG != 2) fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncwarp();
    if (lane0) mbar_arrive_cluster(p_full_remote);
  }
  // ---- epilogue: O / l -> global ----
  mbar_wait(pv_done, (nblk - 1) & 1);
  tc_fence_after();
  float inv_l;
  {  // the denominator is the accumulator column d (one-column load: no register array to index at run time)
    const uint32_t lbits = tmem_ld1(tmem_O + lane_off + p.d);
    tmem_ld_wait();
    inv_l = 1.f / __uint_as_float(lbits);
  }
  const int row = q0 + r;
  __half* orow = p.o + (static_cast<long long>(nb) * p.s_q + row) * p.o_ld + static_cast<long long>(head) * p.d;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);
  for (int c = 0; c < p.d; c += 16) {
    uint32_t o[16];
    tmem_ld16(tmem_O + lane_off + c, o);
    tmem_ld_wait();
    if (row < p.s_q) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = c + g * 8;
        if (col + 8 <= p.d && vec_ok) {
          uint32_t q4[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __half2 h = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * t]) * inv_l,
                                          __uint_as_float(o[g * 8 + 2 * t + 1]) * inv_l);
            q4[t] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(orow + col) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)  // unrolled: a runtime index would move o[] to local memory
            if (col + i < p.d) orow[col + i] = __float2half_rn(__uint_as_float(o[g * 8 + i]) * inv_l);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass softmax role, eight warps per tile: warps w and w + 4 share the 32 query rows of TMEM lane quarter w % 4
// and split the 128 keys of a block in halves (thread = (row, half)). Four softmax warps per scheduler instead of two
// is what hides the MUFU / tcgen05.ld / fence latencies of the one-row-per-thread role (ncu: 48 % MUFU, warps 70 %
// of the time in fixed-latency waits), and 64 keys per thread keep the P row in 32 registers. The two threads of a row
// exchange their half-row maxima through shared memory (double-buffered by block parity) and one 64-thread named
// barrier per block, so both always agree on the reference m and on the (rare) two-pass redo.
// ---------------------------------------------------------------------------------------------------------------
template <int NS, int POLY>
__device__ __forceinline__ void softmax_tile_sp2(const AttnParams& p, int warp, int lane, uint32_t tmem_S,
                                                 uint32_t tmem_O, uint8_t* sP, uint8_t* sV, float* mxbuf,
                                                 uint64_t* s_full, uint64_t* pv_done, uint32_t s_free_remote,
                                                 uint32_t p_full_remote, int q0, int head, int nb, int nblk,
                                                 int crank) {
  const int quarter = warp & 3, hlf = warp >> 2;
  const int r = quarter * 32 + lane;                                   // query row = TMEM lane
  const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
  const uint32_t tS = tmem_S + lane_off + hlf * 64;                    // this thread's 64 score columns
  const float sl = p.scale_log2;
  float m_run = -INFINITY;
  const int one_col = p.d & 31, one_chunk = one_col >> 3, one_elem = p.d & 7;
  const bool write_ones = ((p.d >> 5) == crank) && hlf == 0;
  const bool lane0 = lane == 0;
  int st = 0;
  uint32_t pk[32];  // this thread's 64 probabilities, fp16 pairs

  auto exp_chunk = [&](const uint32_t (&v)[32], int c, float m_sl) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x0 = fmaf(__uint_as_float(v[2 * i]), sl, -m_sl);
      const float x1 = fmaf(__uint_as_float(v[2 * i + 1]), sl, -m_sl);
      pk[c * 16 + i] = (POLY != 0 && (i & 3) == 3) ? ex2_poly_h2(x0, x1) : ex2_h2(x0, x1);
    }
  };
  auto max_chunk = [&](const uint32_t (&v)[32]) {  // four independent chains
    float a = -INFINITY, b = -INFINITY, c = -INFINITY, d = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      a = fmaxf(a, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
      b = fmaxf(b, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
      c = fmaxf(c, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
      d = fmaxf(d, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
    }
    return fmaxf(fmaxf(a, b), fmaxf(c, d));
  };

  for (int j = 0; j < nblk; ++j) {
    const int valid = min(kKV, p.s_kv - j * kKV);
    const int hvalid = max(0, min(64, valid - hlf * 64));  // valid keys of this thread's half
    float* mxw = mxbuf + (j & 1) * 256;
    mbar_wait(s_full, j & 1);
    tc_fence_after();
    float m_new = m_run;
    bool redo = true;
    float mx_half;
    if (valid == kKV && j > 0) {
      // ---- optimistic single pass against m_run ----
      const float m_sl = m_run * sl;
      uint32_t va[32], vb[32];
      tmem_ld32(tS, va);
      tmem_ld32(tS + 32, vb);
      tmem_ld_wait();
      mx_half = fmaxf(max_chunk(va), max_chunk(vb));
      mxw[hlf * 128 + r] = mx_half;
      exp_chunk(va, 0, m_sl);
      exp_chunk(vb, 1, m_sl);
      named_bar_sync(1 + quarter, 64);
      const float mx = fmaxf(mx_half, mxw[(hlf ^ 1) * 128 + r]);
      redo = __any_sync(0xffffffffu, (mx - m_run) * sl > 8.f);
      if (redo && (mx - m_run) * sl > 2.f) m_new = mx;
    } else {
      // ---- first / masked block: the row maximum is needed before any exponential ----
      mx_half = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c * 32 < hvalid) {
          uint32_t v[32];
          tmem_ld32(tS + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < hvalid) mx_half = fmaxf(mx_half, __uint_as_float(v[i]));
        }
      }
      mxw[hlf * 128 + r] = mx_half;
      named_bar_sync(1 + quarter, 64);
      const float mx = fmaxf(mx_half, mxw[(hlf ^ 1) * 128 + r]);
      if ((mx - m_run) * sl > 2.f) m_new = mx;  // also taken on the first block (m_run = -inf)
    }
    if (redo) {
      const float m_sl = m_new * sl;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c * 32 < hvalid) {
          uint32_t v[32];
          tmem_ld32(tS + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k0 = c * 32 + 2 * i;
            const float x0 = (k0 < hvalid) ? fmaf(__uint_as_float(v[2 * i]), sl, -m_sl) : -INFINITY;
            const float x1 = (k0 + 1 < hvalid) ? fmaf(__uint_as_float(v[2 * i + 1]), sl, -m_sl) : -INFINITY;
            pk[c * 16 + i] = ex2_h2(x0, x1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[c * 16 + i] = 0u;
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane0) mbar_arrive_cluster(s_free_remote);  // S(j) has been consumed: the tensor core may overwrite it
    // ---- P V(j-1) must have retired before O is rescaled and the P buffer is overwritten ----
    if (j > 0) {
      mbar_wait(pv_done, (j - 1) & 1);
      tc_fence_after();
      if (!__all_sync(0xffffffffu, m_new == m_run)) {
        const float alpha = ex2((m_run - m_new) * sl);
#pragma unroll
        for (int c = 0; c < 32; c += 16) {  // this half's 32 of the 64 accumulator columns
          uint32_t o[16];
          tmem_ld16(tmem_O + lane_off + hlf * 32 + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tmem_O + lane_off + hlf * 32 + c, o);
        }
        tmem_st_wait();
      }
    }
    m_run = m_new;
    // ---- P -> K-major SW128 smem: slab hlf, row r, 16-byte chunk cc ^ (r & 7) ----
    {
      uint8_t* slab = sP + hlf * kSlab + r * 128;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc)
        *reinterpret_cast<uint4*>(slab + ((cc ^ (r & 7)) << 4)) =
            make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
    }
    if (write_ones) {  // ones column of this V stage, key row r
      uint8_t* vrow = sV + st * kSlab + r * 128;
      *reinterpret_cast<__half*>(vrow + ((one_chunk ^ (r & 7)) << 4) + one_elem * 2) = __float2half_rn(1.f);
    }
    if (++st == NS) st = 0;
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncwarp();
    if (lane0) mbar_arrive_cluster(p_full_remote);
  }
  // ---- epilogue: O / l -> global; each half writes 32 of the value columns ----
  mbar_wait(pv_done, (nblk - 1) & 1);
  tc_fence_after();
  float inv_l;
  {  // the denominator is the accumulator column d (one-column load: no register array to index at run time)
    const uint32_t lbits = tmem_ld1(tmem_O + lane_off + p.d);
    tmem_ld_wait();
    inv_l = 1.f / __uint_as_float(lbits);
  }
  const int row = q0 + r;
  __half* orow = p.o + (static_cast<long long>(nb) * p.s_q + row) * p.o_ld + static_cast<long long>(head) * p.d;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);
  for (int c = hlf * 32; c < min(p.d, hlf * 32 + 32); c += 16) {
    uint32_t o[16];
    tmem_ld16(tmem_O + lane_off + c, o);
    tmem_ld_wait();
    if (row < p.s_q) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = c + g * 8;
        if (col + 8 <= p.d && vec_ok) {
          uint32_t q4[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __half2 h = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * t]) * inv_l,
                                          __uint_as_float(o[g * 8 + 2 * t + 1]) * inv_l);
            q4[t] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(orow + col) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)  // unrolled: a runtime index would move o[] to local memory
            if (col + i < p.d) orow[col + i] = __float2half_rn(__uint_as_float(o[g * 8 + i]) * inv_l);
        }
      }
    }
  }
}

template <int DC, int NS>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  constexpr uint32_t kTmemCols = DC <= 2 ? 256u : 512u;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B slabs need 1024-byte alignment
  uint8_t* sQ = smem;                          // DC slabs
  uint8_t* sK = sQ + DC * kSlab;               // NS x DC slabs
  uint8_t* sV = sK + NS * DC * kSlab;          // NS x DC slabs
  uint8_t* sP = sV + NS * DC * kSlab;          // 2 slabs (keys 0-63, 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kSlab);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + NS;
  uint64_t* s_full = kv_empty + NS;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ;
  const int head = blockIdx.y;
  const int nb = blockIdx.z;
  const int nkb = nb / p.kv_div;
  const int nblk = (p.s_kv + kKV - 1) / kKV;
  const int dk16 = (p.d + 15) / 16;         // K-steps of S = Q K^T
  const int dn = (p.d + 1 + 15) / 16 * 16;  // N of O = P [V | 1]: d value columns + the ones column at index d

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  griddep_sync();

  if (warp == 4) {
    if (role_elect()) {
      // ===== TMA producer =====
      mbar_expect_tx(q_full, DC * kSlab);
      for (int dc = 0; dc < DC; ++dc) tma_load_4d(sQ + dc * kSlab, &tmQ, q_full, dc * 64, head, q0, nb);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * DC * kSlab);
        for (int dc = 0; dc < DC; ++dc) {
          tma_load_4d(sK + (st * DC + dc) * kSlab, &tmK, &kv_full[st], dc * 64, head, j * kKV, nkb);
          tma_load_4d(sV + (st * DC + dc) * kSlab, &tmV, &kv_full[st], dc * 64, head, j * kKV, nkb);
        }
        if (++st == NS) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    if (role_elect()) {
      // ===== MMA issuer =====
      const uint32_t idesc_pv = umma_idesc_f16(128, dn, 0, 1);  // B (= [V | 1]) is MN-major
      mbar_wait(q_full, 0);
      int st = 0;
      uint32_t ph = 0;
      auto n16_of = [&](int j) { return (min(kKV, p.s_kv - j * kKV) + 15) / 16; };
      auto issue_qk = [&](int j, int st_, uint32_t ph_) {  // S = Q K(j)^T
        const uint32_t idesc_qk = umma_idesc_f16(128, n16_of(j) * 16, 0, 0);
        mbar_wait(&kv_full[st_], ph_);
        tc_fence_after();
        for (int ks = 0; ks < dk16; ++ks) {
          const int dc = ks >> 2, kin = ks & 3;
          const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ + dc * kSlab)) + 2 * kin;
          const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK + (st_ * DC + dc) * kSlab)) + 2 * kin;
          umma_f16_ss(tmem_S, qd, kd, idesc_qk, ks != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      };
      // qk_first (needs NS >= 2): the softmax warps have released S(j) when they signal P(j), so S(j+1) is issued
      // BEFORE O += P(j) V(j); the tile's serial chain softmax(j) -> QK(j+1) -> softmax(j+1) no longer contains PV.
      const bool qk_first = NS >= 2 && p.qk_first != 0;
      if (qk_first) issue_qk(0, 0, 0);
      for (int j = 0; j < nblk; ++j) {
        const int n16 = n16_of(j);  // QK^T N (keys) and PV K-steps, in units of 16
        if (!qk_first) issue_qk(j, st, ph);
        // O += P V
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        if (qk_first && j + 1 < nblk) issue_qk(j + 1, st + 1 == NS ? 0 : st + 1, st + 1 == NS ? ph ^ 1 : ph);
        for (int kk = 0; kk < n16; ++kk) {
          const uint64_t pd = umma_desc_kmajor_sw128(smem_u32(sP + (kk >> 2) * kSlab)) + 2 * (kk & 3);
          // V slab: rows = keys (128 B each), 64-wide d atoms kSlab apart; 16 keys = 2048 B
          const uint64_t vd = umma_desc_mnmajor_sw128(smem_u32(sV + st * DC * kSlab) + kk * 2048, kSlab);
          umma_f16_ss(tmem_O, pd, vd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        if (++st == NS) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ===== softmax / correction / output: thread = query row =====
    softmax_tile<DC, NS>(p, warp * 32 + lane, static_cast<uint32_t>(warp * 32) << 16, tmem_S, tmem_O, sP, sV, s_full,
                         p_full, pv_done, q0, head, nb, nblk, dn);
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent form of the one-tile kernel, for the head dims it serves (d = 80 / 160: DC = 2 / 3) where a (frame, head)
// has only 1 - 3 key blocks: with one 128-row tile per CTA every CTA pays tensor-memory allocation, barrier set-up, the
// full TMA latency of its first Q / K / V loads and the output tail for ~4 us of work (S = 384, d = 80: 1 152 CTAs,
// 77 us for a 14 us floor). Here one CTA per SM walks a flat list of (query tile, head, frame) items: Q and the O
// accumulator are double-buffered, the K/V ring, the MMA queue and every barrier phase run across item boundaries, so
// the next item's loads and its first S = Q K^T overlap the current item's softmax tail and output.
// Roles as in attention_tc_kernel (warps 0-3 softmax / output via softmax_tile, warp 4 TMA, warp 5 MMA).
// ---------------------------------------------------------------------------------------------------------------
template <int DC, int NS>
constexpr int attnp1_smem_bytes() {
  return (2 * DC + 2 * NS * DC + 2) * kSlab + 256;
}
struct AttnItems {
  int n_qt, heads, n_items;  // items = n_qt * heads * n_batch, query tile fastest
};

template <int DC, int NS>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_persist1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p,
                          const __grid_constant__ AttnItems pi) {
  constexpr uint32_t kTmemCols = 512u;
  constexpr uint32_t kOStride = DC <= 2 ? 128u : 192u;  // columns per O accumulator (dn <= 128 / 176)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                          // 2 x DC slabs: query tile of the current / next item
  uint8_t* sK = sQ + 2 * DC * kSlab;           // NS x DC slabs
  uint8_t* sV = sK + NS * DC * kSlab;          // NS x DC slabs
  uint8_t* sP = sV + NS * DC * kSlab;          // 2 slabs (keys 0-63, 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kSlab);
  uint64_t* q_full = bars;              // [2]
  uint64_t* q_empty = q_full + 2;       // [2] the last QK of the item has read this Q buffer
  uint64_t* kv_full = q_empty + 2;      // [NS]
  uint64_t* kv_empty = kv_full + NS;    // [NS]
  uint64_t* s_full = kv_empty + NS;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_done = p_full + 1;
  uint64_t* o_empty = pv_done + 1;      // [2] the output role has read this O accumulator
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = (int)blockIdx.x, step = (int)gridDim.x;
  const int nblk = (p.s_kv + kKV - 1) / kKV;
  const int dk16 = (p.d + 15) / 16;         // K-steps of S = Q K^T
  const int dn = (p.d + 1 + 15) / 16 * 16;  // N of O = P [V | 1]
  const int n_my = first < pi.n_items ? (pi.n_items - first + step - 1) / step : 0;
  auto decode = [&](int item, int& q0, int& head, int& nb) {
    q0 = (item % pi.n_qt) * kQ;
    head = (item / pi.n_qt) % pi.heads;
    nb = item / (pi.n_qt * pi.heads);
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
      mbar_init(&o_empty[b], 128);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  griddep_sync();

  if (warp == 4) {
    if (role_elect()) {
      // ===== TMA producer =====
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_my; ++it) {
        int q0, head, nb;
        decode(first + it * step, q0, head, nb);
        const int nkb = nb / p.kv_div;
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], DC * kSlab);
        for (int dc = 0; dc < DC; ++dc) tma_load_4d(sQ + (qb * DC + dc) * kSlab, &tmQ, &q_full[qb], dc * 64, head, q0, nb);
        for (int j = 0; j < nblk; ++j) {
          mbar_wait(&kv_empty[st], ph ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * DC * kSlab);
          for (int dc = 0; dc < DC; ++dc) {
            tma_load_4d(sK + (st * DC + dc) * kSlab, &tmK, &kv_full[st], dc * 64, head, j * kKV, nkb);
            tma_load_4d(sV + (st * DC + dc) * kSlab, &tmV, &kv_full[st], dc * 64, head, j * kKV, nkb);
          }
          if (++st == NS) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 5) {
    if (role_elect()) {
      // ===== MMA issuer: flat block sequence g = it * nblk + j; S(g+1) is issued as soon as the softmax warps have
      // consumed S(g) (they signal it with P(g)), i.e. before O += P(g) V(g) - across item boundaries too =====
      const uint32_t idesc_pv = umma_idesc_f16(128, dn, 0, 1);  // B (= [V | 1]) is MN-major
      const int total = n_my * nblk;
      auto n16_of = [&](int j) { return (min(kKV, p.s_kv - j * kKV) + 15) / 16; };
      int st_qk = 0, it_qk = 0, j_qk = 0;
      uint32_t ph_qk = 0;
      auto issue_qk = [&]() {  // S = Q(it_qk) K(j_qk)^T
        const int qb = it_qk & 1;
        if (j_qk == 0) {
          mbar_wait(&q_full[qb], (it_qk >> 1) & 1);
          tc_fence_after();
        }
        const uint32_t idesc_qk = umma_idesc_f16(128, n16_of(j_qk) * 16, 0, 0);
        mbar_wait(&kv_full[st_qk], ph_qk);
        tc_fence_after();
        for (int ks = 0; ks < dk16; ++ks) {
          const int dc = ks >> 2, kin = ks & 3;
          const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ + (qb * DC + dc) * kSlab)) + 2 * kin;
          const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK + (st_qk * DC + dc) * kSlab)) + 2 * kin;
          umma_f16_ss(tmem_S, qd, kd, idesc_qk, ks != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        if (++j_qk == nblk) {
          umma_commit(&q_empty[qb]);  // Q buffer may be refilled
          j_qk = 0;
          ++it_qk;
        }
        if (++st_qk == NS) {
          st_qk = 0;
          ph_qk ^= 1;
        }
      };
      if (total > 0) issue_qk();
      int st = 0, it = 0, j = 0;
      for (int g = 0; g < total; ++g) {
        const int ob = it & 1;
        if (j == 0 && it >= 2) {  // the output role must have drained this accumulator (item it - 2)
          mbar_wait(&o_empty[ob], ((it >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(p_full, g & 1);  // P(g) written, S(g) consumed
        tc_fence_after();
        // with a one-stage ring the next block's K/V can only be loaded once P V(g) has released the stage: QK after PV
        if (NS >= 2 && g + 1 < total) issue_qk();
        const int n16 = n16_of(j);
        const uint32_t tO = tmem_O + ob * kOStride;
        for (int kk = 0; kk < n16; ++kk) {
          const uint64_t pd = umma_desc_kmajor_sw128(smem_u32(sP + (kk >> 2) * kSlab)) + 2 * (kk & 3);
          const uint64_t vd = umma_desc_mnmajor_sw128(smem_u32(sV + st * DC * kSlab) + kk * 2048, kSlab);
          umma_f16_ss(tO, pd, vd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        if (NS < 2 && g + 1 < total) issue_qk();
        if (++st == NS) st = 0;
        if (++j == nblk) {
          j = 0;
          ++it;
        }
      }
    }
  } else {
    // ===== softmax / correction / output: thread = query row; one softmax_tile per item with flat barrier phases =====
    int st = 0;
    for (int it = 0; it < n_my; ++it) {
      int q0, head, nb;
      decode(first + it * step, q0, head, nb);
      const int ob = it & 1;
      softmax_tile<DC, NS>(p, warp * 32 + lane, static_cast<uint32_t>(warp * 32) << 16, tmem_S, tmem_O + ob * kOStride, sP,
                           sV, s_full, p_full, pv_done, q0, head, nb, nblk, dn, 0u, 0, it * nblk, &st);
      tc_fence_before();
      mbar_arrive(&o_empty[ob]);  // this thread's row of the accumulator has been read
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-tile kernel (d <= 63, many key blocks): one CTA owns TWO 128-row query tiles A and B of the same (frame, head).
// Every K/V stage is loaded once and used by both tiles, the key/value ring is NS deep (the one-tile kernel was bound by
// the ~1.5 us TMA latency of its 2-deep ring), and the single MMA thread interleaves
//   QK_A(j+1) | PV_A(j) | QK_B(j+1) | PV_B(j)
// so that the tensor core works on one tile while the four softmax warps of the other tile exponentiate.
// TMEM: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).  Warps 0-3 softmax A, 4-7 softmax B, 8 TMA, 9 MMA.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAttn2Threads = 320;
template <int NS>
constexpr int attn2_smem_bytes() {
  return (2 + 2 * NS + 4) * kSlab + 256;
}

template <int NS>
__global__ void __launch_bounds__(kAttn2Threads, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  constexpr int DC = 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                       // 2 slabs: tile A, tile B
  uint8_t* sK = sQ + 2 * kSlab;             // NS slabs
  uint8_t* sV = sK + NS * kSlab;            // NS slabs
  uint8_t* sP = sV + NS * kSlab;            // 2 x 2 slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kSlab);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + NS;
  uint64_t* s_full = kv_empty + NS;  // [2]
  uint64_t* p_full = s_full + 2;     // [2]
  uint64_t* pv_done = p_full + 2;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * kQ;
  const int head = blockIdx.y;
  const int nb = blockIdx.z;
  const int nkb = nb / p.kv_div;
  const int nblk = (p.s_kv + kKV - 1) / kKV;
  const int dk16 = (p.d + 15) / 16;
  const int dn = (p.d + 1 + 15) / 16 * 16;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&pv_done[g], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_sync();

  if (warp == 8) {
    if (role_elect()) {
      // ===== TMA producer =====
      mbar_expect_tx(q_full, 2 * kSlab);
      tma_load_4d(sQ, &tmQ, q_full, 0, head, q0, nb);
      tma_load_4d(sQ + kSlab, &tmQ, q_full, 0, head, q0 + kQ, nb);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * kSlab);
        tma_load_4d(sK + st * kSlab, &tmK, &kv_full[st], 0, head, j * kKV, nkb);
        tma_load_4d(sV + st * kSlab, &tmV, &kv_full[st], 0, head, j * kKV, nkb);
        if (++st == NS) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 9) {
    if (role_elect()) {
      // ===== MMA issuer =====
      const uint32_t idesc_pv = umma_idesc_f16(128, dn, 0, 1);
      auto issue_qk = [&](int g, int st, int n16) {
        const uint32_t idesc_qk = umma_idesc_f16(128, n16 * 16, 0, 0);
        for (int ks = 0; ks < dk16; ++ks) {
          const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ + g * kSlab)) + 2 * ks;
          const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK + st * kSlab)) + 2 * ks;
          umma_f16_ss(tmem_base + g * 128, qd, kd, idesc_qk, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int st, int n16, int j) {
        for (int kk = 0; kk < n16; ++kk) {
          const uint64_t pd = umma_desc_kmajor_sw128(smem_u32(sP + (2 * g + (kk >> 2)) * kSlab)) + 2 * (kk & 3);
          const uint64_t vd = umma_desc_mnmajor_sw128(smem_u32(sV + st * kSlab) + kk * 2048, kSlab);
          umma_f16_ss(tmem_base + 256 + g * 64, pd, vd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[g]);
      };
      auto n16_of = [&](int j) { return (min(kKV, p.s_kv - j * kKV) + 15) / 16; };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0, n16_of(0));
      issue_qk(1, 0, n16_of(0));
      int st = 0;
      uint32_t ph = 0;  // stage / phase of block j
      for (int j = 0; j < nblk; ++j) {
        int st1 = st + 1;
        uint32_t ph1 = ph;
        if (st1 == NS) {
          st1 = 0;
          ph1 ^= 1;
        }
        const bool more = j + 1 < nblk;
        // tile A: O_A += P_A V(j), then S_A = Q_A K(j+1)^T (softmax A has finished reading S_A(j) before p_full)
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
        issue_pv(0, st, n16_of(j), j);
        if (more) {
          mbar_wait(&kv_full[st1], ph1);
          tc_fence_after();
          issue_qk(0, st1, n16_of(j + 1));
        }
        // tile B
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        issue_pv(1, st, n16_of(j), j);
        umma_commit(&kv_empty[st]);  // both tiles are done with K(j) and V(j)
        if (more) issue_qk(1, st1, n16_of(j + 1));
        st = st1;
        ph = ph1;
      }
    }
  } else {
    const int g = warp >> 2;  // softmax group = query tile
    softmax_tile<DC, NS>(p, (warp & 3) * 32 + lane, static_cast<uint32_t>((warp & 3) * 32) << 16, tmem_base + g * 128,
                         tmem_base + 256 + g * 64, sP + 2 * g * kSlab, sV, &s_full[g], &p_full[g], &pv_done[g],
                         q0 + g * kQ, head, nb, nblk, dn);
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair kernel (d <= 62): a 2-CTA cluster owns two consecutive 128-row query tiles of one (frame, head) and runs
// every contraction as ONE tcgen05.mma.cta_group::2 (M = 256). Why: at d = 40 a cta_group::1 MMA never costs fewer than
// 76-96 clk however small N is (tools/mma_microbench.cu), so the 3 + 8 instructions of one 128x128 block keep the tensor
// pipe busy for ~1000 clk - as long as the MUFU needs for the 16 384 exponentials. With M = 256 the per-instruction
// floor is 45 clk for twice the rows: ~550 clk per block and SM, which takes the tensor pipe off the critical path.
//   S (both tiles) = [Q_0; Q_1] K^T : each CTA stages its own Q tile and HALF of the key block (64 keys, N = 128)
//   O (both tiles) += [P_0; P_1] [V | 1] : each CTA stages 32 of the 64 value columns (MN-major B, N = 64)
// The leader (even CTA) issues; TMA bytes of both CTAs are counted on the leader's barriers, tcgen05.commit is
// multicast to both, and "P ready" is one remote arrive per softmax warp. S(j+1) is issued before P(j) V(j).
// Two clusters share an SM pair (256 TMEM columns, 96 KB shared memory per CTA).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kKHalf = kSlab / 2;  // bytes of [64 keys x 64 fp16]
template <int NS>
constexpr int attnp_smem_bytes() {
  return kSlab + NS * kKHalf + NS * kSlab + 2 * kSlab + 256;
}

// MODE 0: two-pass softmax role (4 warps); 1: single-pass, one row per thread (4 warps); 2: single-pass, eight warps
template <int NS>
constexpr int attnp_smem_total() {
  return attnp_smem_bytes<NS>() + 2048;  // + half-row maxima exchange buffer of MODE 2
}

template <int NS, int POLY, int MODE, int DBG = 0>
__global__ void __launch_bounds__(MODE == 2 ? 320 : 192, 2)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  constexpr int DC = 1;
  constexpr int SW = MODE == 2 ? 8 : 4;  // softmax warps; warp SW = TMA producer, warp SW + 1 = MMA issuer
  constexpr bool SP = MODE != 0;
  constexpr uint16_t kMask = 3;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                     // 1 slab: this CTA's query tile
  uint8_t* sK = sQ + kSlab;               // NS half slabs: this CTA's 64 keys of the block
  uint8_t* sV = sK + NS * kKHalf;         // NS slabs: 128 keys x this CTA's value columns [32*rank, 32*rank + 64)
  uint8_t* sP = sV + NS * kSlab;          // 2 slabs (keys 0-63, 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kSlab);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + NS;
  uint64_t* s_full = kv_empty + NS;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_done = p_full + 1;
  uint64_t* s_free = pv_done + 1;  // single-pass softmax: S(j) has been read by every softmax warp of the pair
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_free + 1);
  float* mxbuf = reinterpret_cast<float*>(sP + 2 * kSlab + 256);  // [2 parities][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int q0 = blockIdx.x * kQ;
  const int head = blockIdx.y;
  const int nb = blockIdx.z;
  const int nkb = nb / p.kv_div;
  const int nblk = (p.s_kv + kKV - 1) / kKV;
  const int dk16 = (p.d + 15) / 16;
  constexpr int dn = 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);   // leader: expects the bytes of both CTAs
      mbar_init(&kv_empty[s], 1);  // multicast commit
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 2 * SW);  // one arrive per softmax warp of both CTAs (used in the leader only)
    mbar_init(pv_done, 1);
    mbar_init(s_free, 2 * SW);
    fence_mbar_init();
  }
  if (warp == SW + 1) tmem_alloc_2sm<256>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  griddep_sync();

  if (warp == SW) {
    if (role_elect()) {
      // ===== TMA producer (both CTAs; bytes are counted on the leader's barriers) =====
      const uint32_t lead_q = mapa_shared(smem_u32(q_full), 0);
      if (crank == 0) mbar_expect_tx(q_full, 2 * kSlab);
      tma_load_4d_2sm(sQ, &tmQ, lead_q, 0, head, q0, nb);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[st], ph ^ 1);
        const uint32_t lead_kv = mapa_shared(smem_u32(&kv_full[st]), 0);
        if (DBG == 5 && j >= NS) {  // timing experiment: no K/V reloads (stale data), the ring never waits for TMA
          if (crank == 0) mbar_arrive(&kv_full[st]);
          if (++st == NS) {
            st = 0;
            ph ^= 1;
          }
          continue;
        }
        if (crank == 0) mbar_expect_tx(&kv_full[st], 2 * (kKHalf + kSlab));
        tma_load_4d_2sm(sK + st * kKHalf, &tmK, lead_kv, 0, head, j * kKV + crank * 64, nkb);
        tma_load_4d_2sm(sV + st * kSlab, &tmV, lead_kv, crank * 32, head, j * kKV, nkb);
        if (++st == NS) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == SW + 1) {
    if (crank == 0 && role_elect()) {
      // ===== MMA issuer (leader CTA, for both) =====
      constexpr uint32_t idesc_qk = umma_idesc_f16(256, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(256, dn, 0, 1);  // B (= [V | 1]) is MN-major
      auto issue_qk = [&](int st_, uint32_t ph_) {
        mbar_wait(&kv_full[st_], ph_);
        tc_fence_after();
        for (int ks = 0; ks < dk16; ++ks) {
          const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ)) + 2 * ks;
          const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK + st_ * kKHalf)) + 2 * ks;
          if (DBG != 7) umma_f16_ss_2sm(tmem_S, qd, kd, idesc_qk, ks != 0 ? 1u : 0u);
        }
        umma_commit_2sm(s_full, kMask);
      };
      mbar_wait(q_full, 0);
      issue_qk(0, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        const int n16 = (min(kKV, p.s_kv - j * kKV) + 15) / 16;
        if constexpr (SP) {
          if (j + 1 < nblk) {  // S(j) has left tensor memory: S(j+1) goes first, under the tail of softmax(j)
            mbar_wait(s_free, j & 1);
            tc_fence_after();
            issue_qk(st + 1 == NS ? 0 : st + 1, st + 1 == NS ? ph ^ 1 : ph);
          }
          mbar_wait(p_full, j & 1);
          tc_fence_after();
        } else {
          mbar_wait(p_full, j & 1);
          tc_fence_after();
          if (j + 1 < nblk) issue_qk(st + 1 == NS ? 0 : st + 1, st + 1 == NS ? ph ^ 1 : ph);
        }
        for (int kk = 0; kk < n16; ++kk) {
          const uint64_t pd = umma_desc_kmajor_sw128(smem_u32(sP + (kk >> 2) * kSlab)) + 2 * (kk & 3);
          const uint64_t vd = umma_desc_mnmajor_sw128(smem_u32(sV + st * kSlab) + kk * 2048, kSlab);
          if (DBG != 6 || kk == 0) umma_f16_ss_2sm(tmem_O, pd, vd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit_2sm(&kv_empty[st], kMask);
        umma_commit_2sm(pv_done, kMask);
        if (++st == NS) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    const uint32_t s_free_lead = mapa_shared(smem_u32(s_free), 0), p_full_lead = mapa_shared(smem_u32(p_full), 0);
    if constexpr (MODE == 2)
      softmax_tile_sp2<NS, POLY>(p, warp, lane, tmem_S, tmem_O, sP, sV, mxbuf, s_full, pv_done, s_free_lead,
                                 p_full_lead, q0, head, nb, nblk, crank);
    else if constexpr (MODE == 1)
      softmax_tile_sp<NS, POLY, DBG>(p, warp * 32 + lane, static_cast<uint32_t>(warp * 32) << 16, tmem_S, tmem_O, sP, sV,
                                s_full, pv_done, s_free_lead, p_full_lead, q0, head, nb, nblk, dn, crank);
    else
      softmax_tile<DC, NS, true, POLY>(p, warp * 32 + lane, static_cast<uint32_t>(warp * 32) << 16, tmem_S, tmem_O, sP,
                                       sV, s_full, p_full, pv_done, q0, head, nb, nblk, dn, p_full_lead, crank);
  }

  tc_fence_before();
  // neither CTA may exit while the leader's MMAs still read its shared memory or signal its barriers
  cluster_sync_all();
  if (warp == SW + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm<256>(tmem_base);
  }
}

template <int NS, int POLY, int MODE, int DBG = 0>
static int launch_attn_pair(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& ap,
                            dim3 grid, cudaStream_t stream) {
  constexpr int smem = attnp_smem_total<NS>();
  static_assert(2 * smem <= 227 * 1024, "pair attention: two CTAs per SM must fit");
  auto kern = attention_pair_kernel<NS, POLY, MODE, DBG>;
  static DeviceOnce configured;
  if (configured.first()) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(MODE == 2 ? 320 : 192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  IVV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, ap));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent CTA-pair kernel. ncu + knock-out timings of the per-tile pair kernel (tools/attn_bench.py, IVV_ATTN_DBG)
// showed that none of the softmax work (MUFU, P stores, S loads, max) moves its 440 us: the time goes to what every
// 12-block CTA pays once - tensor-memory allocation, two cluster barriers, the first Q/K/V loads with their full TMA
// latency, the output tail - and to the 2-deep K/V ring draining at every tile boundary. Here 2 x 148 CTAs (148
// pairs, two per SM pair) stay resident and walk a flat list of (query-tile pair, head, frame) items: the K/V ring,
// the S hand-off and the MMA queue run across item boundaries, Q and the O accumulator are double-buffered so the
// next item's loads and first MMAs overlap the current item's output, and allocation / cluster syncs happen once.
// Softmax role = the single-pass one (softmax_tile_sp), restated over the flat block sequence.
// ---------------------------------------------------------------------------------------------------------------
struct AttnPersist {
  int n_qpairs, heads, n_items;  // items = n_qpairs * heads * n_batch
};

template <int NS>
constexpr int attnpp_smem_bytes() {
  return 2 * kSlab + NS * kKHalf + NS * kSlab + 2 * kSlab + 256;
}

template <int NS, int POLY>
__global__ void __launch_bounds__(kAttnThreads, attnpp_smem_bytes<NS>() <= 113 * 1024 ? 2 : 1)
attention_pair_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                              const __grid_constant__ AttnParams p, const __grid_constant__ AttnPersist pp) {
  constexpr uint16_t kMask = 3;
  constexpr int SW = 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                     // 2 slabs: query tile of the current / next item
  uint8_t* sK = sQ + 2 * kSlab;           // NS half slabs
  uint8_t* sV = sK + NS * kKHalf;         // NS slabs
  uint8_t* sP = sV + NS * kSlab;          // 2 slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kSlab);
  uint64_t* q_full = bars;             // [2]
  uint64_t* q_empty = q_full + 2;      // [2] last QK of the item has read this Q buffer (multicast commit)
  uint64_t* kv_full = q_empty + 2;     // [NS]
  uint64_t* kv_empty = kv_full + NS;   // [NS]
  uint64_t* s_full = kv_empty + NS;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;
  uint64_t* pv_done = p_full + 1;
  uint64_t* o_empty = pv_done + 1;     // [2] the output role has read this O buffer (leader only)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int first = (int)cluster_id_x(), step = (int)num_clusters_x();
  const int nblk = (p.s_kv + kKV - 1) / kKV;
  const int dk16 = (p.d + 15) / 16;
  constexpr int dn = 64;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int g_, int slot) {
    if (tracing && g_ < 64) p.trace[g_ * 16 + slot] = clock64();
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
      mbar_init(&o_empty[b], 2 * SW);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 2 * SW);
    mbar_init(p_full, 2 * SW);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == SW + 1) tmem_alloc_2sm<256>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;         // [0, 128)
  const uint32_t tmem_O = tmem_base + 128;   // two 64-column accumulators
  griddep_sync();

  if (warp == SW) {
    if (role_elect()) {
      // ===== TMA producer =====
      int st = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int item = first; item < pp.n_items; item += step, ++it) {
        const int qp = item % pp.n_qpairs, head = (item / pp.n_qpairs) % pp.heads, nb = item / (pp.n_qpairs * pp.heads);
        const int q0 = (2 * qp + crank) * kQ, nkb = nb / p.kv_div;
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        if (crank == 0) mbar_expect_tx(&q_full[qb], 2 * kSlab);
        tma_load_4d_2sm(sQ + qb * kSlab, &tmQ, mapa_shared(smem_u32(&q_full[qb]), 0), 0, head, q0, nb);
        for (int j = 0; j < nblk; ++j) {
          mbar_wait(&kv_empty[st], ph ^ 1);
          stamp(it * nblk + j, 8);
          const uint32_t lead_kv = mapa_shared(smem_u32(&kv_full[st]), 0);
          if (crank == 0) mbar_expect_tx(&kv_full[st], 2 * (kKHalf + kSlab));
          tma_load_4d_2sm(sK + st * kKHalf, &tmK, lead_kv, 0, head, j * kKV + crank * 64, nkb);
          tma_load_4d_2sm(sV + st * kSlab, &tmV, lead_kv, crank * 32, head, j * kKV, nkb);
          if (++st == NS) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == SW + 1) {
    if (crank == 0 && role_elect()) {
      // ===== MMA issuer (leader CTA, for both). Flat block sequence g = it * nblk + j =====
      // One thread issues everything, and what it executes between two tcgen05.mma is serial scalar code: a clock64
      // trace showed 110 clk per MMA (integer divisions, descriptor arithmetic, rolled loops) against the 45 clk the
      // instruction needs - the issue loop, not the tensor pipe, set the pace. Hence: descriptors precomputed, block
      // coordinates carried incrementally, fixed-trip unrolled loops for full blocks.
      constexpr uint32_t idesc_qk = umma_idesc_f16(256, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(256, dn, 0, 1);
      const int n_my = first < pp.n_items ? (pp.n_items - first + step - 1) / step : 0;
      const int total = n_my * nblk;
      const uint64_t qd0 = umma_desc_kmajor_sw128(smem_u32(sQ));         // + 1024 (16-byte units) for Q buffer 1
      const uint64_t kd0 = umma_desc_kmajor_sw128(smem_u32(sK));         // + 512 per ring stage
      const uint64_t pd0 = umma_desc_kmajor_sw128(smem_u32(sP));         // + 1024 for keys 64-127
      const uint64_t vd0 = umma_desc_mnmajor_sw128(smem_u32(sV), kSlab); // + 1024 per ring stage, + 128 per 16 keys
      const int n16_last = (p.s_kv - (nblk - 1) * kKV + 15) / 16;
      int st_qk = 0, st_pv = 0;       // ring stage of the next QK / next PV
      uint32_t ph_qk = 0;
      int it_qk = 0, j_qk = 0;        // (item, block) of the next QK
      auto issue_qk = [&]() {         // S = Q(it_qk) K(j_qk)^T
        const int qb = it_qk & 1;
        if (j_qk == 0) {
          mbar_wait(&q_full[qb], (it_qk >> 1) & 1);
          tc_fence_after();
        }
        mbar_wait(&kv_full[st_qk], ph_qk);
        tc_fence_after();
        stamp(it_qk * nblk + j_qk, 5);
        const uint64_t qd = qd0 + static_cast<uint64_t>(qb * (kSlab >> 4));
        const uint64_t kd = kd0 + static_cast<uint64_t>(st_qk * (kKHalf >> 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks < dk16) umma_f16_ss_2sm(tmem_S, qd + 2 * ks, kd + 2 * ks, idesc_qk, ks != 0 ? 1u : 0u);
        umma_commit_2sm(s_full, kMask);
        if (++j_qk == nblk) {
          umma_commit_2sm(&q_empty[qb], kMask);  // Q buffer may be refilled
          j_qk = 0;
          ++it_qk;
        }
        if (++st_qk == NS) {
          st_qk = 0;
          ph_qk ^= 1;
        }
      };
      if (total > 0) issue_qk();
      int it = 0, j = 0;
      for (int g = 0; g < total; ++g) {
        const int ob = it & 1;
        if (g + 1 < total) {  // S(g) has left tensor memory: S(g+1) goes first, under the tail of softmax(g)
          mbar_wait(s_free, g & 1);
          tc_fence_after();
          stamp(g, 4);
          issue_qk();
        }
        if (j == 0 && it >= 2) {  // the output role must have drained this accumulator (item it - 2)
          mbar_wait(&o_empty[ob], ((it >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(p_full, g & 1);
        tc_fence_after();
        stamp(g, 6);
        const uint32_t tO = tmem_O + ob * 64;
        const uint64_t vd = vd0 + static_cast<uint64_t>(st_pv * (kSlab >> 4));
        if (j + 1 < nblk || n16_last == 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16_ss_2sm(tO, pd0 + (kk >> 2) * (kSlab >> 4) + 2 * (kk & 3), vd + 128 * kk, idesc_pv,
                            (kk != 0 || j != 0) ? 1u : 0u);
        } else {
          for (int kk = 0; kk < n16_last; ++kk)
            umma_f16_ss_2sm(tO, pd0 + (kk >> 2) * (kSlab >> 4) + 2 * (kk & 3), vd + 128 * kk, idesc_pv,
                            (kk != 0 || j != 0) ? 1u : 0u);
        }
        umma_commit_2sm(&kv_empty[st_pv], kMask);
        umma_commit_2sm(pv_done, kMask);
        stamp(g, 7);
        if (++st_pv == NS) st_pv = 0;
        if (++j == nblk) {
          j = 0;
          ++it;
        }
      }
    }
  } else {
    // ===== softmax / correction / output: thread = query row; single pass over S (see softmax_tile_sp) =====
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t s_free_lead = mapa_shared(smem_u32(s_free), 0), p_full_lead = mapa_shared(smem_u32(p_full), 0);
    const float sl = p.scale_log2;
    const int one_col = p.d & 31, one_chunk = one_col >> 3, one_elem = p.d & 7;
    const bool write_ones = (p.d >> 5) == crank;
    const bool lane0 = lane == 0;
    const bool store_issuer = role_elect();  // one lane per warp; only warp 0's issues the output stores
    int st = 0;
    int g = 0;  // flat block counter
    int it = 0;
    uint32_t pk[64];
    auto exp_chunk = [&](const uint32_t (&v)[32], int c, float m_sl) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x0 = fmaf(__uint_as_float(v[2 * i]), sl, -m_sl);
        const float x1 = fmaf(__uint_as_float(v[2 * i + 1]), sl, -m_sl);
        pk[c * 16 + i] = (POLY != 0 && (i & 3) == 3) ? ex2_poly_h2(x0, x1) : ex2_h2(x0, x1);
      }
    };
    auto max_chunk = [&](const uint32_t (&v)[32], float mx) {  // four independent chains
      float a = -INFINITY, b = -INFINITY, c = -INFINITY, d = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        a = fmaxf(a, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        b = fmaxf(b, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
        c = fmaxf(c, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
        d = fmaxf(d, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
      }
      return fmaxf(mx, fmaxf(fmaxf(a, b), fmaxf(c, d)));
    };
    // item coordinates are advanced incrementally (three runtime div/mod per item cost ~1 000 clk on this path)
    int qp = first % pp.n_qpairs, head = (first / pp.n_qpairs) % pp.heads, nb = first / (pp.n_qpairs * pp.heads);
    const int d_qp = step % pp.n_qpairs, d_head = (step / pp.n_qpairs) % pp.heads, d_nb = step / (pp.n_qpairs * pp.heads);
    for (int item = first; item < pp.n_items; item += step, ++it) {
      const int q0 = (2 * qp + crank) * kQ;
      const uint32_t tO = tmem_O + (it & 1) * 64 + lane_off;
      float m_run = -INFINITY;
      for (int j = 0; j < nblk; ++j, ++g) {
        const int valid = min(kKV, p.s_kv - j * kKV);
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        if (r == 0) stamp(g, 0);
        float m_new = m_run;
        bool redo = true;
        if (valid == kKV && j > 0) {
          // ---- optimistic single pass against m_run ----
          const float m_sl = m_run * sl;
          float mx = -INFINITY;
          uint32_t va[32], vb[32];
          tmem_ld32(tmem_S + lane_off, va);
          tmem_ld_wait();
          tmem_ld32(tmem_S + lane_off + 32, vb);
          mx = max_chunk(va, mx);
          exp_chunk(va, 0, m_sl);
          tmem_ld_wait();
          tmem_ld32(tmem_S + lane_off + 64, va);
          mx = max_chunk(vb, mx);
          exp_chunk(vb, 1, m_sl);
          tmem_ld_wait();
          tmem_ld32(tmem_S + lane_off + 96, vb);
          mx = max_chunk(va, mx);
          exp_chunk(va, 2, m_sl);
          tmem_ld_wait();
          mx = max_chunk(vb, mx);
          redo = __any_sync(0xffffffffu, (mx - m_run) * sl > 8.f);
          if (!redo) {
            tc_fence_before();
            __syncwarp();
            if (r == 0) stamp(g, 1);
            if (lane0) mbar_arrive_cluster(s_free_lead);  // S(g) is in registers
            exp_chunk(vb, 3, m_sl);
          }
        }
        if (redo) {
          // ---- two-pass: first block of an item, masked last block, or a row maximum that jumped ----
          if (valid == kKV) {
            // full block: two loads in flight per wait
            float mx = -INFINITY;
            uint32_t va[32], vb[32];
#pragma unroll
            for (int c = 0; c < 4; c += 2) {
              tmem_ld32(tmem_S + lane_off + c * 32, va);
              tmem_ld32(tmem_S + lane_off + c * 32 + 32, vb);
              tmem_ld_wait();
              mx = max_chunk(va, mx);
              mx = max_chunk(vb, mx);
            }
            if ((mx - m_run) * sl > 2.f) m_new = mx;  // also taken on the first block (m_run = -inf)
            const float m_sl = m_new * sl;
#pragma unroll
            for (int c = 0; c < 4; c += 2) {
              tmem_ld32(tmem_S + lane_off + c * 32, va);
              tmem_ld32(tmem_S + lane_off + c * 32 + 32, vb);
              tmem_ld_wait();
              exp_chunk(va, c, m_sl);
              exp_chunk(vb, c + 1, m_sl);
            }
          } else {
            const int nchunk = (valid + 31) / 32;
            float mx = -INFINITY;
            for (int c = 0; c < nchunk; ++c) {
              uint32_t v[32];
              tmem_ld32(tmem_S + lane_off + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            if ((mx - m_run) * sl > 2.f) m_new = mx;
            const float m_sl = m_new * sl;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < nchunk) {
                uint32_t v[32];
                tmem_ld32(tmem_S + lane_off + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int k0 = c * 32 + 2 * i;
                  const float x0 = (k0 < valid) ? fmaf(__uint_as_float(v[2 * i]), sl, -m_sl) : -INFINITY;
                  const float x1 = (k0 + 1 < valid) ? fmaf(__uint_as_float(v[2 * i + 1]), sl, -m_sl) : -INFINITY;
                  pk[c * 16 + i] = ex2_h2(x0, x1);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[c * 16 + i] = 0u;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane0) mbar_arrive_cluster(s_free_lead);
        }
        // ---- P V(g-1) must have retired before O is rescaled and the P buffer is overwritten (at j = 0 the
        //      output step of the previous item has already waited for it) ----
        if (j > 0) {
          if (r == 0) stamp(g, 9);
          mbar_wait(pv_done, (g - 1) & 1);
          tc_fence_after();
          if (r == 0) stamp(g, 2);
          if (!__all_sync(0xffffffffu, m_new == m_run)) {
            const float alpha = ex2((m_run - m_new) * sl);
            for (int c = 0; c < dn; c += 16) {
              uint32_t o[16];
              tmem_ld16(tO + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tO + c, o);
            }
            tmem_st_wait();
          }
        }
        m_run = m_new;
        if (j == 0 && it > 0) {  // the output tile of the previous item sits in P slab 0 until its TMA store has read it
          if (warp == 0 && store_issuer) bulk_wait_group_read<0>();
          named_bar_sync(1, 128);
        }
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          uint8_t* slab = sP + (cc >> 3) * kSlab + r * 128;
          *reinterpret_cast<uint4*>(slab + (((cc & 7) ^ (r & 7)) << 4)) =
              make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
        }
        if (write_ones) {  // ones column of this V stage, key row r
          uint8_t* vrow = sV + st * kSlab + r * 128;
          *reinterpret_cast<__half*>(vrow + ((one_chunk ^ (r & 7)) << 4) + one_elem * 2) = __float2half_rn(1.f);
        }
        if (++st == NS) st = 0;
        if (r == 0) stamp(g, 10);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (r == 0) stamp(g, 3);
        if (lane0) mbar_arrive_cluster(p_full_lead);
      }
      // ---- output of this item: O / l -> global, then hand the accumulator back ----
      if (r == 0) stamp(g - 1, 11);
      mbar_wait(pv_done, (g - 1) & 1);
      tc_fence_after();
      if (r == 0) stamp(g - 1, 12);
      // The normalised tile goes to global memory as ONE TMA store out of P slab 0 (free until the next item's first
      // P): per-row 16-byte STGs (128 rows x 640-byte stride) clogged the LSU queue in front of the next mbarrier
      // polls and cost 1 700 clk per item in the clock64 trace. Rows >= s_q and columns >= d are clipped by the map.
      uint32_t o[64];
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        uint32_t t16[16];
        tmem_ld16(tO + c, t16);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c + i] = t16[i];
      }
      const uint32_t lbits = tmem_ld1(tO + p.d);  // denominator = accumulator column d
      tmem_ld_wait();
      const float inv_l = __fdividef(1.f, __uint_as_float(lbits));
      {
        uint8_t* srow = sP + r * 128;
#pragma unroll
        for (int c8 = 0; c8 < 7; ++c8) {
          if (c8 * 8 < p.d) {
            uint32_t q4[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __half2 h = __floats2half2_rn(__uint_as_float(o[c8 * 8 + 2 * t]) * inv_l,
                                            __uint_as_float(o[c8 * 8 + 2 * t + 1]) * inv_l);
              q4[t] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(srow + ((c8 ^ (r & 7)) << 4)) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, 128);
      if (warp == 0 && store_issuer) {
        tma_store_4d(&tmO, sP, 0, head, q0, nb);
        bulk_commit_group();
      }
      tc_fence_before();
      __syncwarp();
      if (r == 0) stamp(g - 1, 13);
      if (lane0) mbar_arrive_cluster(mapa_shared(smem_u32(&o_empty[it & 1]), 0));
      qp += d_qp;
      head += d_head;
      nb += d_nb;
      if (qp >= pp.n_qpairs) {
        qp -= pp.n_qpairs;
        ++head;
      }
      if (head >= pp.heads) {
        head -= pp.heads;
        ++nb;
      }
    }
    if (warp == 0 && store_issuer) bulk_wait_group<0>();  // output stores complete before the CTA may exit
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == SW + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm<256>(tmem_base);
  }
}

template <int NS, int POLY>
static int launch_attn_pair_persist(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                    const CUtensorMap& to, const AttnParams& ap, const AttnPersist& pp,
                                    cudaStream_t stream) {
  constexpr int smem = attnpp_smem_bytes<NS>();
  constexpr int kPerSm = smem <= 113 * 1024 ? 2 : 1;
  static_assert(smem <= 227 * 1024, "persistent pair attention exceeds shared memory");
  auto kern = attention_pair_persist_kernel<NS, POLY>;
  static DeviceOnce configured;
  static int n_sm = 148;
  if (configured.first()) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev = 0;
    IVV_CHECK_CUDA(cudaGetDevice(&dev));
    IVV_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  int clusters = n_sm / 2 * kPerSm;  // kPerSm CTAs per SM
  if (clusters > pp.n_items) clusters = pp.n_items;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  IVV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, to, ap, pp));
  return 0;
}

template <int NS>
static int launch_attn2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& ap,
                        dim3 grid, cudaStream_t stream) {
  constexpr int smem = attn2_smem_bytes<NS>();
  static_assert(smem <= 227 * 1024, "two-tile attention exceeds shared memory");
  static DeviceOnce configured;
  if (configured.first()) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  IVV_CHECK_CUDA(launch_pdl(attention_tc2_kernel<NS>, grid, dim3(kAttn2Threads), smem, stream, tq, tk, tv, ap));
  return 0;
}

template <int DC, int NS>
static int launch_attn_persist1(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& ap,
                                const AttnItems& pi, cudaStream_t stream) {
  constexpr int smem = attnp1_smem_bytes<DC, NS>();
  static_assert(smem <= 227 * 1024, "persistent one-tile attention exceeds shared memory");
  auto kern = attention_persist1_kernel<DC, NS>;
  static DeviceOnce configured;
  static int n_sm = 148;
  if (configured.first()) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev = 0;
    IVV_CHECK_CUDA(cudaGetDevice(&dev));
    IVV_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const int ctas = pi.n_items < n_sm ? pi.n_items : n_sm;
  IVV_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)ctas), dim3(kAttnThreads), smem, stream, tq, tk, tv, ap, pi));
  return 0;
}

template <int DC, int NS>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& ap,
                       dim3 grid, cudaStream_t stream) {
  constexpr int smem = attn_smem_bytes<DC, NS>();
  static DeviceOnce configured;
  if (configured.first()) {
    IVV_CHECK_CUDA(
        cudaFuncSetAttribute(attention_tc_kernel<DC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  IVV_CHECK_CUDA(launch_pdl(attention_tc_kernel<DC, NS>, grid, dim3(kAttnThreads), smem, stream, tq, tk, tv, ap));
  return 0;
}

}  // namespace ivv

static long long* g_attn_trace = nullptr;
// tuning hook (not part of include/ivv.h): device buffer of 64 x 16 int64 that the persistent kernel fills with clock64
// stamps of its first CTA; nullptr switches tracing off
extern "C" void ivv_debug_attn_trace(void* buf) { g_attn_trace = reinterpret_cast<long long*>(buf); }

namespace ivv {
// Tuning switches of ivv_attention, read once per process. A non-tuning build (no -DIVV_TUNING) only has the default
// kernels: persistent CTA pairs for d <= 62, the one-tile kernel otherwise.
struct AttnEnv {
  int qk_first, mode, dbg;
  bool pair, pair_short, poly, ns6, two_tile, persist1;
};
static const AttnEnv& attn_env() {
  static const AttnEnv e = [] {
    auto geti = [](const char* name, int dflt) {
      const char* v = getenv(name);
      return v ? atoi(v) : dflt;
    };
    AttnEnv a{};
    a.qk_first = geti("IVV_ATTN_QK_FIRST", 1);
    a.pair = geti("IVV_ATTN_PAIR", 1) != 0;
    a.pair_short = geti("IVV_ATTN_PAIR_SHORT", 1) != 0;
    a.poly = geti("IVV_ATTN_POLY", 0) != 0;
    a.ns6 = geti("IVV_ATTN_NS", 2) == 6;
    a.persist1 = geti("IVV_ATTN_PERSIST1", 1) != 0;
#ifdef IVV_TUNING
    a.mode = geti("IVV_ATTN_MODE", 3);
    a.dbg = geti("IVV_ATTN_DBG", 0);
    a.two_tile = getenv("IVV_ATTN_TWO_TILE") != nullptr;
#else
    a.mode = 3;
#endif
    return a;
  }();
  return e;
}
}  // namespace ivv

extern "C" int ivv_attention(const void* q, int64_t q_ld, const void* k, const void* v, int64_t kv_ld, void* o,
                             int64_t o_ld, int64_t n_batch, int64_t s_q, int64_t s_kv, int64_t kv_div, int32_t heads,
                             int32_t d, float scale, ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(q && k && v && o, "ivv_attention: null pointer");
  IVV_REQUIRE(n_batch > 0 && s_q > 0 && s_kv > 0 && heads > 0 && kv_div > 0, "ivv_attention: empty problem");
  IVV_REQUIRE(n_batch % kv_div == 0, "ivv_attention: n_batch (%lld) not a multiple of kv_div (%lld)",
              (long long)n_batch, (long long)kv_div);
  IVV_REQUIRE(d % 8 == 0 && d >= 8 && d <= 184, "ivv_attention: head dim %d must be a multiple of 8 in [8, 184]", d);
  IVV_REQUIRE(q_ld % 8 == 0 && kv_ld % 8 == 0 && o_ld % 8 == 0, "ivv_attention: leading dims must be multiples of 8");
  IVV_REQUIRE(n_batch <= 65535 && heads <= 65535, "ivv_attention: grid too large");

  AttnParams ap{};
  ap.s_q = (int)s_q;
  ap.s_kv = (int)s_kv;
  ap.kv_div = (int)kv_div;
  ap.d = d;
  ap.scale_log2 = scale * 1.4426950408889634f;
  ap.o = reinterpret_cast<__half*>(o);
  ap.o_ld = o_ld;
  ap.trace = g_attn_trace;

  CUtensorMap tq, tk, tv;
  const uint32_t box[4] = {64, 1, 128, 1};
  {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)s_q, (uint64_t)n_batch};
    const uint64_t str[4] = {2, (uint64_t)d * 2, (uint64_t)q_ld * 2, (uint64_t)q_ld * 2 * s_q};
    if (int rc = make_tmap_f16(&tq, q, 4, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)s_kv, (uint64_t)(n_batch / kv_div)};
    const uint64_t str[4] = {2, (uint64_t)d * 2, (uint64_t)kv_ld * 2, (uint64_t)kv_ld * 2 * s_kv};
    if (int rc = make_tmap_f16(&tk, k, 4, dims, str, box, 128)) return rc;
    if (int rc = make_tmap_f16(&tv, v, 4, dims, str, box, 128)) return rc;
  }
  const int dc = (d + 1 + 63) / 64;  // 64-wide chunks holding the d value columns plus the ones column
  const AttnEnv& env = attn_env();  // tuning switches, read from the environment ONCE (no getenv on the launch path)
  ap.qk_first = env.qk_first;
  // CTA pairs (cta_group::2): default for d <= 62 with at least two query tiles and more than one key block;
  // IVV_ATTN_PAIR=0 falls back to the one-tile kernel, IVV_ATTN_POLY=0 keeps every exponential on the MUFU
  {
    // single-block problems (cross-attention, 77 keys) also go through the persistent pair kernel: its Q / O double
    // buffering overlaps one item's loads and output with the next (82 -> 55 us at S_q = 1536); IVV_ATTN_PAIR_SHORT=0
    // sends them back to the one-tile kernel
    const bool short_ok = s_kv > kKV || env.pair_short;
    const bool pair = dc == 1 && s_q > kQ && short_ok && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && env.pair;
    if (pair) {
      CUtensorMap tk64;
      const uint32_t box64[4] = {64, 1, 64, 1};
      const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)s_kv, (uint64_t)(n_batch / kv_div)};
      const uint64_t str[4] = {2, (uint64_t)d * 2, (uint64_t)kv_ld * 2, (uint64_t)kv_ld * 2 * s_kv};
      if (int rc = make_tmap_f16(&tk64, k, 4, dims, str, box64, 128)) return rc;
      const unsigned tiles = (unsigned)((s_q + kQ - 1) / kQ);
      const bool poly = env.poly;
      if (env.mode == 3) {  // persistent pairs (default; the only mode of a non-tuning build)
        AttnPersist pp{};
        pp.n_qpairs = (int)((tiles + 1) / 2);
        pp.heads = heads;
        pp.n_items = pp.n_qpairs * heads * (int)n_batch;
        CUtensorMap to;
        {
          const uint64_t odims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)s_q, (uint64_t)n_batch};
          const uint64_t ostr[4] = {2, (uint64_t)d * 2, (uint64_t)o_ld * 2, (uint64_t)o_ld * 2 * s_q};
          if (int rc = make_tmap_f16(&to, o, 4, odims, ostr, box, 128)) return rc;
        }
        if (env.ns6) return launch_attn_pair_persist<6, 0>(tq, tk64, tv, to, ap, pp, stream);
        return poly ? launch_attn_pair_persist<2, 1>(tq, tk64, tv, to, ap, pp, stream)
                    : launch_attn_pair_persist<2, 0>(tq, tk64, tv, to, ap, pp, stream);
      }
#ifdef IVV_TUNING  // retired variants, kept for A/B timing only (build with IVV_NVCC_EXTRA=-DIVV_TUNING)
      dim3 gridp((tiles + 1) / 2 * 2, (unsigned)heads, (unsigned)n_batch);
      switch (env.dbg) {  // timing experiments (tools/attn_bench.py); garbage results
        case 1: return launch_attn_pair<2, 0, 1, 1>(tq, tk64, tv, ap, gridp, stream);
        case 2: return launch_attn_pair<2, 0, 1, 2>(tq, tk64, tv, ap, gridp, stream);
        case 3: return launch_attn_pair<2, 0, 1, 3>(tq, tk64, tv, ap, gridp, stream);
        case 4: return launch_attn_pair<2, 0, 1, 4>(tq, tk64, tv, ap, gridp, stream);
        case 5: return launch_attn_pair<2, 0, 1, 5>(tq, tk64, tv, ap, gridp, stream);
        case 6: return launch_attn_pair<2, 0, 1, 6>(tq, tk64, tv, ap, gridp, stream);
        case 7: return launch_attn_pair<2, 0, 1, 7>(tq, tk64, tv, ap, gridp, stream);
        default: break;
      }
      switch (env.mode * 2 + (poly ? 1 : 0)) {
        case 0: return launch_attn_pair<2, 0, 0>(tq, tk64, tv, ap, gridp, stream);
        case 1: return launch_attn_pair<2, 1, 0>(tq, tk64, tv, ap, gridp, stream);
        case 2: return launch_attn_pair<2, 0, 1>(tq, tk64, tv, ap, gridp, stream);
        case 3: return launch_attn_pair<2, 1, 1>(tq, tk64, tv, ap, gridp, stream);
        case 4: return launch_attn_pair<2, 0, 2>(tq, tk64, tv, ap, gridp, stream);
        default: return launch_attn_pair<2, 1, 2>(tq, tk64, tv, ap, gridp, stream);
      }
#endif
    }
  }
#ifdef IVV_TUNING
  // the two-tile kernel (K/V loads shared by 256 queries, 4-deep ring) measured equal to the one-tile kernel at
  // S=1536, d=40 (488 vs 471 us): both are bound by the softmax warps, so it stays a tuning variant
  if (dc == 1 && s_kv > 2 * kKV && s_q > kQ && env.two_tile) {
    dim3 grid2((unsigned)((s_q + 2 * kQ - 1) / (2 * kQ)), (unsigned)heads, (unsigned)n_batch);
    return launch_attn2<4>(tq, tk, tv, ap, grid2, stream);
  }
#endif
  dim3 grid((unsigned)((s_q + kQ - 1) / kQ), (unsigned)heads, (unsigned)n_batch);
  if (dc == 1) return launch_attn<1, 2>(tq, tk, tv, ap, grid, stream);
  // d = 80 / 160: persistent one-tile kernel once there are more items than SMs to amortise over (in the graph of a
  // forward: S = 384 d = 80 76.9 -> 62.5 us, its 77-key cross-attention 43.9 -> 32.9 us, d = 160 18.7 / 18.3 -> 16.3 /
  // 15.4 us; profiles/r02_attention_persist1_ab.txt). IVV_ATTN_PERSIST1=0 goes back to one tile per CTA (tuning hook).
  const long long n_items = (long long)grid.x * heads * n_batch;
  if (env.persist1 && n_items > 148 && n_items < (1LL << 30)) {
    AttnItems pi{};
    pi.n_qt = (int)grid.x;
    pi.heads = heads;
    pi.n_items = (int)n_items;
    if (dc == 2) return launch_attn_persist1<2, 2>(tq, tk, tv, ap, pi, stream);
    return launch_attn_persist1<3, 1>(tq, tk, tv, ap, pi, stream);
  }
  if (dc == 2) return launch_attn<2, 2>(tq, tk, tv, ap, grid, stream);
  return launch_attn<3, 1>(tq, tk, tv, ap, grid, stream);
}
