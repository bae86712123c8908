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
template <int DC, int NS>
__device__ __forceinline__ void softmax_tile(const AttnParams& p, int r, uint32_t lane_off, uint32_t tmem_S,
                                             uint32_t tmem_O, uint8_t* sP, uint8_t* sV, uint64_t* s_full,
                                             uint64_t* p_full, uint64_t* pv_done, int q0, int head, int nb, int nblk,
                                             int dn) {
  const float sl = p.scale_log2;
  float m_run = -INFINITY;
  // The softmax denominator is never summed on the CUDA cores: a column of ones is written into the V tile at
  // column d (slab d/64, 16-byte chunk (d%64)/8, element d%8), so O[:, d] accumulates sum(P) in fp32 inside the
  // tensor core and follows every online rescale for free.
  const int one_slab = p.d >> 6, one_chunk = (p.d & 63) >> 3, one_elem = p.d & 7;
  int st = 0;
  for (int j = 0; j < nblk; ++j) {
    const int valid = min(kKV, p.s_kv - j * kKV);
    const int nchunk = (valid + 31) / 32;
    mbar_wait(s_full, j & 1);
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
      mbar_wait(pv_done, (j - 1) & 1);
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
          for (int t = 0; t < 4; ++t)
            pk[t] = ex2_h2(fmaf(__uint_as_float(v[g * 8 + 2 * t]), sl, -m_sl),
                           fmaf(__uint_as_float(v[g * 8 + 2 * t + 1]), sl, -m_sl));
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
    {  // ones column of this V stage, key row r
      uint8_t* vrow = sV + (st * DC + one_slab) * kSlab + r * 128;
      *reinterpret_cast<__half*>(vrow + ((one_chunk ^ (r & 7)) << 4) + one_elem * 2) = __float2half_rn(1.f);
    }
    if (++st == NS) st = 0;
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    mbar_arrive(p_full);
  }
  // ---- epilogue: O / l -> global ----
  mbar_wait(pv_done, (nblk - 1) & 1);
  tc_fence_after();
  float inv_l;
  {
    uint32_t o[16];
    tmem_ld16(tmem_O + lane_off + (p.d & ~15), o);
    tmem_ld_wait();
    float l = 1.f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i == (p.d & 15)) l = __uint_as_float(o[i]);
    inv_l = 1.f / l;
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
          for (int i = 0; i < 8; ++i)
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
    if (lane == 0) {
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
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc_pv = umma_idesc_f16(128, dn, 0, 1);  // B (= [V | 1]) is MN-major
      mbar_wait(q_full, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        const int valid = min(kKV, p.s_kv - j * kKV);
        const int n16 = (valid + 15) / 16;  // QK^T N (keys) and PV K-steps, in units of 16
        const uint32_t idesc_qk = umma_idesc_f16(128, n16 * 16, 0, 0);
        mbar_wait(&kv_full[st], ph);
        tc_fence_after();
        // S = Q K^T
        for (int ks = 0; ks < dk16; ++ks) {
          const int dc = ks >> 2, kin = ks & 3;
          const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ + dc * kSlab)) + 2 * kin;
          const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK + (st * DC + dc) * kSlab)) + 2 * kin;
          umma_f16_ss(tmem_S, qd, kd, idesc_qk, ks != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        // O += P V
        mbar_wait(p_full, j & 1);
        tc_fence_after();
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
    if (lane == 0) {
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
    if (lane == 0) {
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

template <int NS>
static int launch_attn2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& ap,
                        dim3 grid, cudaStream_t stream) {
  constexpr int smem = attn2_smem_bytes<NS>();
  static_assert(smem <= 227 * 1024, "two-tile attention exceeds shared memory");
  static bool configured = false;
  if (!configured) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  IVV_CHECK_CUDA(launch_pdl(attention_tc2_kernel<NS>, grid, dim3(kAttn2Threads), smem, stream, tq, tk, tv, ap));
  return 0;
}

template <int DC, int NS>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& ap,
                       dim3 grid, cudaStream_t stream) {
  constexpr int smem = attn_smem_bytes<DC, NS>();
  static bool configured = false;
  if (!configured) {
    IVV_CHECK_CUDA(
        cudaFuncSetAttribute(attention_tc_kernel<DC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  IVV_CHECK_CUDA(launch_pdl(attention_tc_kernel<DC, NS>, grid, dim3(kAttnThreads), smem, stream, tq, tk, tv, ap));
  return 0;
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
  // the two-tile kernel (K/V loads shared by 256 queries, 4-deep ring) measured equal to the one-tile kernel at
  // S=1536, d=40 (488 vs 471 us): both are bound by the softmax warps, so it stays opt-in
  if (dc == 1 && s_kv > 2 * kKV && s_q > kQ && getenv("IVV_ATTN_TWO_TILE") != nullptr) {
    dim3 grid2((unsigned)((s_q + 2 * kQ - 1) / (2 * kQ)), (unsigned)heads, (unsigned)n_batch);
    return launch_attn2<4>(tq, tk, tv, ap, grid2, stream);
  }
  dim3 grid((unsigned)((s_q + kQ - 1) / kQ), (unsigned)heads, (unsigned)n_batch);
  if (dc == 1) return launch_attn<1, 2>(tq, tk, tv, ap, grid, stream);
  if (dc == 2) return launch_attn<2, 2>(tq, tk, tv, ap, grid, stream);
  return launch_attn<3, 1>(tq, tk, tv, ap, grid, stream);
}
