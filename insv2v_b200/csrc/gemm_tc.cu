// tcgen05 implicit-GEMM convolution / linear for sm_100a.
//
//   D[pix, co] = sum_{tap, ci} A[pix (+tap shift), ci] * W[tap][co][ci]     fp16 x fp16 -> fp32 (TMEM) -> fp16/fp32
//
// A is a channels-last activation tensor [n_img, h, w, c]. One CTA computes a 128-pixel x BLOCK_N tile. The 128 pixels
// are a (bw x bh x bn) box of the (w, h, n_img) index space, fetched per filter tap as ONE 4-D TMA box whose start
// coordinate is shifted by the tap offset: the zero padding of the 3x3 convolution is TMA out-of-bounds zero fill, and
// no im2col buffer ever exists. A Linear layer is the same kernel with taps = 1 and box (128, 1, 1).
// Warp roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM owner), warps 2..5 = epilogue
// (tcgen05.ld -> bias / temb row-bias / residual / GEGLU -> global).  Reference call sites: ivv.h (K1/K2/K11).
#include <cstdlib>

#include "../../include/ivv.h"
#include "common.cuh"

namespace ivv {

struct GemmKParams {
  int bw, bh, bn;                 // A box (pixels); bw*bh*bn == 128
  int tiles_w, tiles_h, tiles_g;  // tiles along w, h and image groups
  int W, H, NI;                   // output extent
  int kblocks;                    // ceil(c / 64) per tap
  int taps;                       // tap_h * tap_w
  int tap_w, tap_h;               // window geometry (stride 1, "same" zero padding): 1x1, 3x3, 1x5, 5x1, ...
  int relu;                       // epilogue: max(x, 0) after bias / rowbias / residual
  int n_out;                      // GEMM N (weight rows)
  int out_cols;                   // columns written (n_out, or n_out/2 for GEGLU)
  int geglu, out_f32;
  int splits;                     // split-K (v1 kernel): blockIdx.z owns a contiguous range of (tap, k-block) iterations
  long long split_stride;         // elements between the fp32 partial planes
  int ws_stages;                  // >0: weight-stationary mode (persistent kernel): the CTA's weight tile (all of K) is
                                  // loaded once and stays in shared memory; only A tiles stream through ws_stages slots
  int dbg_skip;                   // tuning only (IVV_DEBUG_SKIP): 1 = no MMA issue, 2 = no TMA loads, 3 = no TMA store,
                                  // 4 = no residual term, 5 = no epilogue work (results are garbage)
  long long* trace;               // tuning only (ivv_debug_gemm_trace): clock64 stamps of CTA 0, [tile][16]
  int halo_bytes;                 // HALO kernels: bytes of one (bw x (bh + 2)) activation box = (bh + 2) * bw * 128
  int as_prefetch;                // AS modes: prefetch the activation rows of the cluster's NEXT M pair into L2
  void* d;
  long long d_ld;
  const __half* bias;
  const __half* rowbias;
  long long rowbias_group, rowbias_ld;
  const __half* residual;
  long long res_ld;
  // pair160 kernel only
  int rowbias_mod;            // > 0: rowbias row = (pix / rowbias_group) % rowbias_mod
  float2* row_stats_out;      // [rows][n_out / 40]: per-row partial (sum, sum of squares) of the fp32 results
  const float2* ln_stats;     // [rows][ln_parts]: statistics of the A rows (LayerNorm folded into this GEMM)
  const __half* ln_wsum;      // [n_out]
  int ln_parts;
  float ln_eps, ln_inv_c;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kGemmThreads = 192;

template <int BN>
constexpr uint32_t tmem_cols_for() {
  return BN <= 32 ? 32u : BN <= 64 ? 64u : BN <= 128 ? 128u : 256u;
}

template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  return STAGES * (kABytes + BN * 128) + 256 /*barriers + tmem ptr*/;
}

__device__ __forceinline__ void store8(void* d, bool f32, long long off, const float (&v)[8], int nvalid, bool vec_ok) {
  if (f32) {
    float* p = reinterpret_cast<float*>(d) + off;
    if (nvalid == 8 && vec_ok) {
      reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      for (int j = 0; j < nvalid; ++j) p[j] = v[j];
    }
  } else {
    __half* p = reinterpret_cast<__half*>(d) + off;
    if (nvalid == 8 && vec_ok) {
      __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
      __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2);
      u.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(p) = u;
    } else {
      for (int j = 0; j < nvalid; ++j) p[j] = __float2half_rn(v[j]);
    }
  }
}

__device__ __forceinline__ void load8h(const __half* p, int nvalid, bool vec_ok, float (&v)[8]) {
  if (nvalid == 8 && vec_ok) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < nvalid ? __half2float(p[j]) : 0.f;
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmKParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  constexpr int kStageBytes = kABytes + BN * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntile = blockIdx.x;
  const int mtile = blockIdx.y;
  // tile -> pixel box origin
  const int tw = mtile % p.tiles_w;
  const int th = (mtile / p.tiles_w) % p.tiles_h;
  const int tg = mtile / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.bw, h0 = th * p.bh;
      const int n0 = mtile < p.tiles_w * p.tiles_h * p.tiles_g ? tg * p.bn : p.NI;  // ghost tile of an odd cluster: fully out of bounds
  const int all_it = p.taps * p.kblocks;
  const int per_split = (all_it + p.splits - 1) / p.splits;
  const int it_begin = blockIdx.z * per_split;
  const int total_it = min(all_it, it_begin + per_split) - it_begin;  // iterations of this split (>= 1 by construction)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<tmem_cols_for<BN>()>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_sync();  // PDL: everything above overlapped the previous kernel's tail

  // Single-thread roles are entered through elect.sync on a converged warp, not `lane == 0`: ptxas then knows exactly
  // one thread runs the branch and keeps descriptors / barrier addresses in uniform registers. With a divergent
  // `lane == 0` every tcgen05.mma / TMA / commit was wrapped in an ELECT + 5x R2UR + BRA.U.ANY serialisation loop
  // (~15 scalar instructions per MMA), which paced the main loop instead of the tensor pipe.
  if (warp == 0) {
    if (role_elect()) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total_it; ++it) {
        const int tap = (it_begin + it) / p.kblocks;
        const int kb = (it_begin + it) - tap * p.kblocks;
        int dy = 0, dx = 0;
        if (p.taps > 1) {
          dy = tap / p.tap_w - (p.tap_h >> 1);
          dx = tap % p.tap_w - (p.tap_w >> 1);
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_4d(sa, &tmA, &full_bar[stage], kb * kBlockK, w0 + dx, h0 + dy, n0);
        tma_load_3d(sb, &tmB, &full_bar[stage], kb * kBlockK, ntile * BN, tap);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (role_elect()) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total_it; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint64_t adesc = umma_desc_kmajor_sw128(sa);
        const uint64_t bdesc = umma_desc_kmajor_sw128(sa + kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 fp16 = 32 B along K inside the 128-B swizzle atom: +2 in the (addr >> 4) field
          umma_f16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4) .. +32 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int wi = r % p.bw;
    const int hi = (r / p.bw) % p.bh;
    const int ni = r / (p.bw * p.bh);
    const int w = w0 + wi, h = h0 + hi, n = n0 + ni;
    const bool valid = (w < p.W) && (h < p.H) && (n < p.NI);
    const long long pix = (static_cast<long long>(n) * p.H + h) * p.W + w;
    void* dptr = p.splits > 1 ? static_cast<void*>(reinterpret_cast<float*>(p.d) + blockIdx.z * p.split_stride) : p.d;
    const bool d_vec = (p.d_ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(dptr) & 31) == 0);
    const bool r_vec = p.residual && (p.res_ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    const __half* rb = p.rowbias ? p.rowbias + (pix / p.rowbias_group) * p.rowbias_ld : nullptr;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    if (!p.geglu) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t acc[32];
        tmem_ld32(taddr + c, acc);
        tmem_ld_wait();
        const int col0 = ntile * BN + c;
        if (valid && col0 < p.n_out) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            const int nv = min(8, p.n_out - col);
            if (nv <= 0) break;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nv) v[j] += __half2float(p.bias[col + j]);
            }
            if (rb) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nv) v[j] += __half2float(rb[col + j]);
            }
            if (p.residual) {
              float rv[8];
              load8h(p.residual + pix * p.res_ld + col, nv, r_vec, rv);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += rv[j];
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            store8(dptr, p.out_f32 != 0, pix * p.d_ld + col, v, nv, d_vec);
          }
        }
      }
    } else {
      // GEGLU: tile columns [0, BN/2) = hidden, [BN/2, BN) = gate of the same output columns
      constexpr int HALF = BN / 2;
#pragma unroll 1
      for (int c = 0; c < HALF; c += 32) {
        uint32_t hacc[32], gacc[32];
        tmem_ld32(taddr + c, hacc);
        tmem_ld32(taddr + HALF + c, gacc);
        tmem_ld_wait();
        const int bcol0 = ntile * BN + c;      // position of hidden bias in the interleaved bias vector
        const int ocol0 = ntile * HALF + c;    // output column
        if (valid && ocol0 < p.out_cols) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int ocol = ocol0 + g * 8;
            const int nv = min(8, p.out_cols - ocol);
            if (nv <= 0) break;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float hv = __uint_as_float(hacc[g * 8 + j]);
              float gv = __uint_as_float(gacc[g * 8 + j]);
              if (p.bias && j < nv) {
                hv += __half2float(p.bias[bcol0 + g * 8 + j]);
                gv += __half2float(p.bias[bcol0 + HALF + g * 8 + j]);
              }
              v[j] = hv * gelu_erf_f(gv);
            }
            store8(dptr, p.out_f32 != 0, pix * p.d_ld + ocol, v, nv, d_vec);
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<tmem_cols_for<BN>()>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v2: persistent kernel. One CTA per SM walks the tile list; the fp32 accumulator is double-buffered in TMEM so the
// epilogue of tile i (8 warps: TMEM -> regs -> bias/temb/residual/GEGLU -> fp16 -> swizzled smem -> TMA store, fully
// coalesced and clipped by the tensor map) overlaps the TMA/MMA main loop of tile i+1. Used for fp16 outputs whose row
// stride is a multiple of 16 bytes (everything in the UNet/VAE except the 4- and 3-channel fp32 heads).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPersistThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two groups of 4)

template <int BN>
constexpr uint32_t acc_stride_for() {
  return BN <= 32 ? 32u : BN <= 64 ? 64u : BN <= 128 ? 128u : 256u;
}
// HALO (3x3 convolutions, pair mode): one stage holds a (bw x (bh + 2)) activation box -- at most kHaloRows pixel rows of
// 64 channels -- and the three weight tiles of the filter column it serves (dy = -1, 0, +1).
constexpr int kHaloRows = 160;
constexpr int kHaloBytes = kHaloRows * 128;  // 20 KB, a multiple of the 1024-B swizzle atom
// GEGLU epilogue: hidden | gate bias slices of the tile in shared memory: [2 tiles][2 groups][2 chunks][32 + 32] fp16
constexpr int kGegluBiasBytes = 2 * 2 * 2 * 64 * 2;
template <int BN, int STAGES, int CW, bool GEGLU, bool TILEWIDE, bool TWO = false, bool HALO = false, bool DS = false>
constexpr int persist_smem_bytes() {
  // TILEWIDE: the whole fp16 output tile is staged (DS: two such slabs); otherwise one CW-wide slab per epilogue group
  return STAGES * (HALO ? kHaloBytes + 3 * (BN / 2) * 128 : kABytes + (TWO ? BN / 2 : BN) * 128) +
         (TILEWIDE ? (DS ? 2 : 1) * kBlockM * (GEGLU ? BN / 2 : BN) * 2 : (BN > 256 ? 4 : 2) * kBlockM * CW * 2) + 256 +
         (GEGLU ? kGegluBiasBytes : 0);
}

template <int BN, int STAGES, int CW, bool GEGLU, bool TILEWIDE, int CS, bool TWO, bool HALO = false, bool DS = false,
          bool AS = false>
__global__ void __launch_bounds__(kPersistThreads, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
                          const __grid_constant__ GemmKParams p, int n_tiles, int total_tiles) {
  // CS > 1: a cluster of CS CTAs works on CS consecutive M tiles of the same N tile; every CTA fetches 1/CS of the
  // weight tile and TMA-multicasts it to all of them (opt-in, measured neutral: the same bytes still enter every SM).
  // total_tiles then counts super tiles (M-tile groups x N tiles) and the loop strides by clusters.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // TWO: the pair runs ONE tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own 128 A rows and only HALF of the
  // weight tile, which is what lifts the ~60 B/clk per-SM operand-ingest limit of the single-CTA kernel.
  // HALO (3x3, stride 1, box inside one frame, bw % 8 == 0): the three taps of one filter COLUMN read the same pixels
  // shifted by whole box rows, so ONE (bw x (bh + 2)) box per (dx, 64-channel block) serves all three: the MMA of tap
  // dy starts (dy + 1) * bw rows into the box (a multiple of the 8-row swizzle atom, so the descriptor stays canonical).
  // Activation bytes per tile drop from 9 x 16 KB to 3 x <= 20 KB per channel block: the main loops fill shared memory
  // at a near-constant ~40 B/clk per SM, so time follows staged bytes, not tensor-pipe work.
  static_assert(!TWO || CS == 2, "pair MMA needs a 2-CTA cluster");
  static_assert(!HALO || (TWO && !GEGLU), "halo mode is built on the pair kernel");
  // DS (tile-wide staging, residual GEMMs with short main loops): TWO staging slabs used alternately. With one slab the
  // residual tile of tile i+1 can only be requested after the store of tile i has drained it, so every tile pays a full
  // HBM round trip in its epilogue -- with K <= 640 that chain (not the main loop) set the pace. With two, the residual
  // of tile i+1 is requested when the epilogue of tile i STARTS and has a whole tile period to land.
  static_assert(!DS || TILEWIDE, "double staging belongs to the tile-wide epilogues");
  static_assert(!GEGLU || (TILEWIDE && CW == 32 && BN == 256), "GEGLU epilogue: tile-wide staging, 32-column chunks");
  // AS (activation-stationary, pair mode, at most STAGES K blocks per tile: the K = 320 GEGLU GEMM of the 32x48 level): as
  // in gemm_tc_pair160_kernel<.., AS>, a cluster takes ALL N tiles of an M pair; the activation parts of the stages hold
  // the pair's rows for the whole M pair (slice it in stage it), only the weight parts cycle. With K = 320 a tile needs
  // the whole ring, so in the plain mode every tile pays a load round trip; here only weights (L2-resident) are in flight.
  static_assert(!AS || (TWO && !HALO && BN <= 256), "activation-stationary mode is built on the plain pair kernel");
  // WIDE (BN = 320, pair mode): the tile is two 160-wide halves that share the activation tile -- two N = 160 MMAs per
  // k-step into ONE 320-column accumulator (2 x 320 fp32 columns do not fit the 512 of TMEM, so the accumulator is not
  // double-buffered: the epilogue of a tile is exposed, which only pays for long main loops). The N = 1280 layers of the
  // 8x12 level (36 M tiles) are 72 such tiles = ONE wave of the 74 clusters; as 160- / 256-wide tiles they are 144 / 90 = two
  // rounds, and fetch the activation rows 8 / 5 times instead of 4.
  constexpr bool kWide = BN > 256;
  static_assert(!kWide || (BN == 320 && TWO && !HALO && !GEGLU && !TILEWIDE && !DS && CW == 32),
                "320-wide tiles: pair mode, ring staging, 32-column chunks");
  constexpr int kBRows = TWO ? BN / 2 : BN;
  constexpr int kBTileBytes = kBRows * 128;
  constexpr int kAOff = HALO ? kHaloBytes : kABytes;  // offset of the weight tile(s) inside a stage
  constexpr int kStageBytes = kAOff + (HALO ? 3 : 1) * kBTileBytes;
  constexpr uint16_t kMask = (1u << CS) - 1;
  const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int tile_first = CS > 1 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int tile_step = CS > 1 ? (int)num_clusters_x() : (int)gridDim.x;
  // tiles of this CTA / cluster, in order: round robin over the M-major tile list; AS: the M pairs tile_first,
  // tile_first + tile_step, ... with all their N tiles
  const int as_groups = total_tiles / n_tiles;
  const int n_local = AS ? (tile_first < as_groups ? ((as_groups - tile_first + tile_step - 1) / tile_step) * n_tiles : 0)
                         : (tile_first < total_tiles ? (total_tiles - tile_first + tile_step - 1) / tile_step : 0);
  auto tile_of = [&](int l) -> int {
    if constexpr (AS) {
      const int j = l / n_tiles;
      return (tile_first + j * tile_step) * n_tiles + (l - j * n_tiles);
    } else {
      return tile_first + l * tile_step;
    }
  };
  constexpr int kChunkBytes = kBlockM * CW * 2;  // one [128 rows x CW fp16] swizzled slab per column chunk
  constexpr int kStagingBytes = TILEWIDE ? kBlockM * (GEGLU ? BN / 2 : BN) * 2 : (kWide ? 4 : 2) * kChunkBytes;
  constexpr uint32_t kAccStride = kWide ? 0u : acc_stride_for<BN>();
  constexpr uint32_t kTmemCols = kWide ? 512u : 2 * kAccStride;
  // accumulator buffer and barrier parity of the CTA's local-th tile (WIDE: one buffer, its barriers flip every tile)
  auto acc_buf = [](int local_) { return kWide ? 0 : (local_ & 1); };
  auto acc_phase = [](int local_) { return static_cast<uint32_t>(kWide ? (local_ & 1) : ((local_ >> 1) & 1)); };
  uint8_t* staging = smem + STAGES * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (DS ? 2 : 1) * kStagingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint64_t* res_bar = tmem_empty_bar + 2;        // [2 groups][2 slabs] residual tile landed (single slab: [g] only)
  uint64_t* b_full = res_bar + 4;                // weight-stationary mode: resident weight tile landed
  uint64_t* a_full = b_full + 1;                 // AS [STAGES]: resident activation slice landed (leader counts both CTAs)
  uint64_t* a_empty = a_full + STAGES;           // AS [STAGES]: the last N tile of the M pair has read the slice
  static_assert(!AS || (4 * STAGES + 9) * 8 + 4 <= 256, "barrier block");
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(AS ? a_empty + STAGES : b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int its_per_tile = (HALO ? 3 : p.taps) * p.kblocks;  // HALO: one iteration = (filter column, channel block)
  // weight-stationary layout of the same stage region: [its_per_tile x (BN x 128 B) resident B][ws_stages x 16 KB A ring]
  const bool ws = !TWO && CS == 1 && p.ws_stages > 0;
  uint8_t* b_res = smem;
  uint8_t* a_ring = smem + its_per_tile * (BN * 128);
  const int nstages = ws ? p.ws_stages : STAGES;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                // pair mode: the leader expects the bytes of both CTAs
      mbar_init(&empty_bar[s], TWO ? 1 : CS);    // multicast mode: every CTA must have consumed the stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], TWO ? 16 : 8);  // one arrive per epilogue warp (of both CTAs in pair mode)
      mbar_init(&res_bar[b], 1);
      mbar_init(&res_bar[2 + b], 1);
    }
    mbar_init(b_full, 1);
    if constexpr (AS) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&a_full[s], 1);
        mbar_init(&a_empty[s], 1);
      }
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (TWO) tmem_alloc_2sm<kTmemCols>(tmem_ptr); else tmem_alloc<kTmemCols>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (CS > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_sync();  // PDL: everything above overlapped the previous kernel's tail
  // tuning trace (tools/gemm_trace.py): slots 0-1 producer (first / last load of the tile issued), 2-3 MMA thread
  // (accumulator free, tile committed), 4-11 epilogue group 0 (top, store drained, accumulator seen, residual seen,
  // last chunk done, group barrier, store issued; first chunk only: 11 bias requested, 12 accumulator in registers,
  // 13 bias added, 14 result written to the slab)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int t_, int slot) {
    if (tracing && t_ < 32) p.trace[t_ * 16 + slot] = clock64();
  };

  if (warp == 0) {
    if (role_elect()) {  // elect.sync, not lane == 0: see gemm_tc_kernel
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      if (ws) {  // gridDim.x is a multiple of n_tiles, so every tile of this CTA has the same N tile
        const int ntile0 = blockIdx.x % n_tiles;
        mbar_expect_tx(b_full, its_per_tile * BN * 128);
        for (int it = 0; it < its_per_tile; ++it) {
          const int tap = it / p.kblocks;
          tma_load_3d(b_res + it * (BN * 128), &tmB, b_full, (it - tap * p.kblocks) * kBlockK, ntile0 * BN, tap);
        }
      }
      for (int plocal = 0; plocal < n_local; ++plocal) {
        const int tile = tile_of(plocal);
        const int ntile = tile % n_tiles;
        const int mtile = (tile / n_tiles) * CS + crank;
        const int tw = mtile % p.tiles_w;
        const int th = (mtile / p.tiles_w) % p.tiles_h;
        const int tg = mtile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh;
      const int n0 = mtile < p.tiles_w * p.tiles_h * p.tiles_g ? tg * p.bn : p.NI;  // ghost tile of an odd cluster: fully out of bounds
        for (int it = 0; it < its_per_tile; ++it) {
          if (it == 0 || it == its_per_tile - 1) stamp(plocal, it == 0 ? 0 : 1);
          const int tap = it / p.kblocks;
          const int kb = it - tap * p.kblocks;
          int dy = 0, dx = 0;
          if (p.taps > 1) {
            dy = tap / p.tap_w - (p.tap_h >> 1);
            dx = tap % p.tap_w - (p.tap_w >> 1);
          }
          if constexpr (AS) {
            if (ntile == 0) {  // first N tile of an M pair: (re)load the resident activation slice of this K block
              const uint32_t jpar = (uint32_t)(plocal / n_tiles) & 1u;
              mbar_wait(&a_empty[it], jpar ^ 1u);
              const uint32_t lead_a = mapa_shared(smem_u32(&a_full[it]), 0);
              if (crank == 0) mbar_expect_tx(&a_full[it], 2 * kABytes);
              tma_load_4d_2sm(smem + it * kStageBytes, &tmA, lead_a, kb * kBlockK, w0 + dx, h0 + dy, n0);
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          if constexpr (HALO) {
            const int dxi = it / p.kblocks;  // filter column 0..2 (dx = dxi - 1)
            const int kc = (it - dxi * p.kblocks) * kBlockK;
            const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * (p.halo_bytes + 3 * kBTileBytes));
            tma_load_4d_2sm(sa, &tmA, lead_bar, kc, w0 + dxi - 1, h0 - 1, n0);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi)
              tma_load_3d_2sm(sa + kAOff + dyi * kBTileBytes, &tmB, lead_bar, kc, ntile * BN + crank * kBRows,
                              dyi * 3 + dxi);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if constexpr (TWO) {
            // bytes of BOTH CTAs are counted on the leader's barrier. The peer needs no arrive of its own: it can only
            // refill a stage after the leader's MMA released it, i.e. after the leader's barrier finished that phase.
            const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_expect_tx(&full_bar[stage], AS ? 2 * kBTileBytes : 2 * kStageBytes);
            if constexpr (!AS) tma_load_4d_2sm(sa, &tmA, lead_bar, kb * kBlockK, w0 + dx, h0 + dy, n0);
            if constexpr (kWide) {
              // weight rows of this CTA: its quarter of each 160-wide half (tmB's box is BN / 4 rows), so that the pair's
              // MMA h sees the rows [h * 160, h * 160 + 160) of the tile in order
              tma_load_3d_2sm(sa + kABytes, &tmB, lead_bar, kb * kBlockK, ntile * BN + crank * (BN / 4), tap);
              tma_load_3d_2sm(sa + kABytes + (BN / 4) * 128, &tmB, lead_bar, kb * kBlockK,
                              ntile * BN + BN / 2 + crank * (BN / 4), tap);
            } else {
              tma_load_3d_2sm(sa + kABytes, &tmB, lead_bar, kb * kBlockK, ntile * BN + crank * kBRows, tap);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if (ws) {
            mbar_expect_tx(&full_bar[stage], kABytes);
            tma_load_4d(a_ring + stage * kABytes, &tmA, &full_bar[stage], kb * kBlockK, w0 + dx, h0 + dy, n0);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if (p.dbg_skip == 2) {
            mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_4d(sa, &tmA, &full_bar[stage], kb * kBlockK, w0 + dx, h0 + dy, n0);
          if constexpr (CS > 1)
            tma_load_3d_multicast(sa + kABytes + crank * (BN / CS) * 128, &tmB, &full_bar[stage], kb * kBlockK,
                                  ntile * BN + crank * (BN / CS), tap, kMask);
          else
            tma_load_3d(sa + kABytes, &tmB, &full_bar[stage], kb * kBlockK, ntile * BN, tap);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (AS) {
          // the M switch is the one place where this mode waits for DRAM: pull the NEXT M pair's rows into L2 meanwhile
          if (ntile == 0 && p.as_prefetch && plocal + n_tiles < n_local) {
            const int mt1 = (tile_of(plocal + n_tiles) / n_tiles) * CS + crank;
            const int w1 = (mt1 % p.tiles_w) * p.bw, h1 = ((mt1 / p.tiles_w) % p.tiles_h) * p.bh;
            const int n1 = mt1 < p.tiles_w * p.tiles_h * p.tiles_g ? (mt1 / (p.tiles_w * p.tiles_h)) * p.bn : p.NI;
            for (int kb = 0; kb < p.kblocks; ++kb) tma_prefetch_l2_4d(&tmA, kb * kBlockK, w1, h1, n1);
          }
        }
      }
    }
  } else if (warp == 1) {
    if ((!TWO || crank == 0) && role_elect()) {
      // ===== MMA issuer (pair mode: the leader CTA issues for both) =====
      constexpr uint32_t idesc = umma_idesc_f16(TWO ? 2 * kBlockM : kBlockM, kWide ? BN / 2 : BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      if (ws) mbar_wait(b_full, 0);
      for (int local = 0; local < n_local; ++local) {
        const int as_j = AS ? local / n_tiles : 0;          // AS: M pair index of this cluster, N tile inside it
        const int as_nt = AS ? local - as_j * n_tiles : 0;
        const int buf = acc_buf(local);
        mbar_wait(&tmem_empty_bar[buf], acc_phase(local) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        stamp(local, 2);
        const uint32_t tacc = tmem_base + buf * kAccStride;
        for (int it = 0; it < its_per_tile; ++it) {
          if constexpr (AS) {
            if (as_nt == 0) mbar_wait(&a_full[it], (uint32_t)(as_j & 1));  // resident activation slice of this M pair
          }
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = ws ? smem_u32(a_ring + stage * kABytes) : smem_u32(smem + stage * kStageBytes);
          if constexpr (HALO) {
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const uint64_t adesc = umma_desc_kmajor_sw128(sa + dyi * p.bw * 128);
              const uint64_t bdesc = umma_desc_kmajor_sw128(sa + kAOff + dyi * kBTileBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16_ss_2sm(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | dyi | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], kMask);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          const uint64_t adesc = umma_desc_kmajor_sw128(AS ? smem_u32(smem + it * kStageBytes) : sa);
          const uint64_t bdesc = umma_desc_kmajor_sw128(ws ? smem_u32(b_res + it * (BN * 128)) : sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if (p.dbg_skip == 1 && (it | k) != 0) continue;
            if constexpr (kWide) {
              const uint64_t bdesc_hi = umma_desc_kmajor_sw128(sa + kABytes + (BN / 4) * 128);
              umma_f16_ss_2sm(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
              umma_f16_ss_2sm(tacc + BN / 2, adesc + 2 * k, bdesc_hi + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
            } else if constexpr (TWO) {
              umma_f16_ss_2sm(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
            } else {
              umma_f16_ss(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (TWO) umma_commit_2sm(&empty_bar[stage], kMask);
          else if constexpr (CS > 1) umma_commit_multicast(&empty_bar[stage], kMask);
          else umma_commit(&empty_bar[stage]);
          if constexpr (AS) {
            if (as_nt == n_tiles - 1) umma_commit_2sm(&a_empty[it], kMask);  // last reader of the slice: it may be replaced
          }
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (TWO) umma_commit_2sm(&tmem_full_bar[buf], kMask); else umma_commit(&tmem_full_bar[buf]);
        stamp(local, 3);
      }
    }
  } else {
    // ===== epilogue: group g = (warp-2)/4 handles column chunks g, g+2, ...; warp%4 = TMEM lane quarter =====
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool issuer = (warp == 2 + 4 * g) && (lane == 0);
    const int wi = r % p.bw;
    const int hi = (r / p.bw) % p.bh;
    const int ni = r / (p.bw * p.bh);
    constexpr int OUT_W = GEGLU ? BN / 2 : BN;  // output columns per tile
    constexpr int NCHUNK = OUT_W / CW;
    constexpr int VPR = CW / 8;                 // 16-byte vectors per staging row
    // accumulator hand-back: in pair mode the MMA issuer lives in the leader CTA
    auto release_acc = [&](int buf_) {
      if constexpr (TWO) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[buf_]), 0));
      else mbar_arrive(&tmem_empty_bar[buf_]);
    };
    const bool has_res = !GEGLU && p.residual != nullptr && p.dbg_skip != 4;  // dbg_skip 4: residual term dropped
    const int my_chunks = (NCHUNK - g + 1) / 2;  // chunks g, g+2, ...
    uint32_t res_phase = 0;
    // DS: request the residual tile of `tile_` into staging slab `slab_` (barrier index 2 * slab_ + g)
    auto request_res = [&](int tile_, int slab_) {
      const int ntile_ = tile_ % n_tiles;
      const int mtile_ = (tile_ / n_tiles) * CS + crank;
      const int w0_ = (mtile_ % p.tiles_w) * p.bw, h0_ = ((mtile_ / p.tiles_w) % p.tiles_h) * p.bh;
      const int n0_ = mtile_ < p.tiles_w * p.tiles_h * p.tiles_g ? (mtile_ / (p.tiles_w * p.tiles_h)) * p.bn : p.NI;
      mbar_expect_tx(&res_bar[2 * slab_ + g], my_chunks * kChunkBytes);
      for (int chunk = g; chunk < NCHUNK; chunk += 2)
        tma_load_4d(staging + slab_ * kStagingBytes + chunk * kChunkBytes, &tmR, &res_bar[2 * slab_ + g],
                    ntile_ * OUT_W + chunk * CW, w0_, h0_, n0_);
    };
    if constexpr (DS) {
      if (issuer && has_res && my_chunks > 0 && n_local > 0) request_res(tile_of(0), 0);
    }
    // GEGLU: the hidden | gate bias slices of a tile's chunks are staged in shared memory one tile ahead (threads r < 16 of
    // each group: one 16-byte vector each, requested at the top of the previous tile and written at its end, so the L2
    // round trip never sits in front of the arithmetic; the eight per-chunk global loads of the old form did).
    // Layout: [tile parity][group][chunk of the group][hidden 32 | gate 32] fp16.
    __half* const gb_base = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(full_bar) + 256);
    auto gbias_fetch = [&](int tile_) -> uint4 {
      const int ntile_ = tile_ % n_tiles;
      const int col = ntile_ * BN + ((r >> 2) & 1) * (BN / 2) + (g + 2 * (r >> 3)) * CW + (r & 3) * 8;
      if (p.bias == nullptr) return make_uint4(0u, 0u, 0u, 0u);
      if ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) return __ldg(reinterpret_cast<const uint4*>(p.bias + col));
      union { uint4 u; __half h[8]; } t;
#pragma unroll
      for (int j = 0; j < 8; ++j) t.h[j] = p.bias[col + j];
      return t.u;
    };
    auto gbias_slot = [&](int buf_) -> uint4* {
      return reinterpret_cast<uint4*>(gb_base + ((buf_ * 2 + g) * 2 + (r >> 3)) * 64) + (r & 7);
    };
    if constexpr (GEGLU) {
      if (r < 16 && n_local > 0) *gbias_slot(0) = gbias_fetch(tile_of(0));  // visible after the tile's top barrier
    }
    // WIDE: the epilogue runs in the open (one accumulator), so it is kept short: two 32-column slabs per group used
    // alternately (only the store issued two chunks ago must have drained), and the residual comes straight from global
    // memory into registers, one chunk ahead -- the first chunk's while the main loop is still running. (Ring staging
    // with a TMA-fetched residual paid a store drain plus a TMA round trip per chunk: ~14 000 clk per tile.)
    uint32_t wchunk = 0;  // chunks this group has written so far (slab parity)
    for (int local = 0; local < n_local; ++local) {
      const int tile = tile_of(local);
      uint8_t* const slab = staging + (DS ? (local & 1) * kStagingBytes : 0);
      uint4 gb_next = make_uint4(0u, 0u, 0u, 0u);
      const bool gb_pre = GEGLU && r < 16 && local + 1 < n_local;
      if constexpr (GEGLU) {
        if (gb_pre) gb_next = gbias_fetch(tile_of(local + 1));
      }
      const int ntile = tile % n_tiles;
      const int mtile = (tile / n_tiles) * CS + crank;
      const int tw = mtile % p.tiles_w;
      const int th = (mtile / p.tiles_w) % p.tiles_h;
      const int tg = mtile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.bw, h0 = th * p.bh;
      const int n0 = mtile < p.tiles_w * p.tiles_h * p.tiles_g ? tg * p.bn : p.NI;  // ghost tile of an odd cluster: fully out of bounds
      const int w = w0 + wi, h = h0 + hi, n = n0 + ni;
      const bool valid = (w < p.W) && (h < p.H) && (n < p.NI);
      const long long pix = (static_cast<long long>(n) * p.H + h) * p.W + w;
      const __half* rb = (p.rowbias && valid) ? p.rowbias + (pix / p.rowbias_group) * p.rowbias_ld : nullptr;
      const int buf = acc_buf(local);
      const bool etr = issuer && g == 0;
      if (etr) stamp(local, 4);
      if constexpr (TILEWIDE) {
        // While the main loop of this tile is still running: make sure the stores of the previous tile have drained
        // the staging slabs, then let TMA drop the residual tile straight into them (coalesced, no registers).
        if (issuer) {
          if constexpr (DS) {
            // residual GEMM: this tile's slab was refilled by the residual fetch, which already waited for its store;
            // the drain of the PREVIOUS tile's store (it sits behind the producer's loads in the TMA queue: ~1 100 clk
            // in the trace) is waited for after the first chunk, where it has long finished
            // (GEGLU, two slabs and no residual: only the store issued two tiles ago must have left this slab)
            if constexpr (GEGLU) bulk_wait_group_read<1>();
            else if (!has_res) bulk_wait_group_read<0>();
          } else {
            bulk_wait_group_read<0>();
            if (has_res && my_chunks > 0) {
              mbar_expect_tx(&res_bar[g], my_chunks * kChunkBytes);
              for (int chunk = g; chunk < NCHUNK; chunk += 2)
                tma_load_4d(staging + chunk * kChunkBytes, &tmR, &res_bar[g], ntile * OUT_W + chunk * CW, w0, h0, n0);
            }
          }
        }
        if (!has_res) named_bar_sync(1 + g, 128);
      }
      if (etr) stamp(local, 5);
      uint4 rcur[kWide ? VPR : 1], rnxt[kWide ? VPR : 1];
      const __half* rrow = nullptr;  // WIDE: this thread's residual row (nullptr: none / row outside the tensor)
      if constexpr (kWide) {
        if (has_res && valid) rrow = p.residual + pix * p.res_ld + ntile * OUT_W;
#pragma unroll
        for (int cc = 0; cc < VPR; ++cc)
          rcur[cc] = rrow != nullptr && g < NCHUNK ? __ldg(reinterpret_cast<const uint4*>(rrow + g * CW) + cc)
                                                   : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&tmem_full_bar[buf], acc_phase(local));
      tc_fence_after();
      if (etr) stamp(local, 6);
      if constexpr (TILEWIDE) {
        if (has_res && my_chunks > 0) {
          if constexpr (DS) {
            mbar_wait(&res_bar[2 * (local & 1) + g], (local >> 1) & 1);
          } else {
            mbar_wait(&res_bar[g], res_phase);
            res_phase ^= 1;
          }
        }
      }
      if (etr) stamp(local, 7);
      const uint32_t taddr = tmem_base + buf * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      if (p.dbg_skip == 5) {  // tuning only: no epilogue work at all (accumulator handed straight back, nothing stored)
        tc_fence_before();
        if (lane == 0) release_acc(buf);
        continue;
      }
#pragma unroll 1
      for (int chunk = g; chunk < NCHUNK; chunk += 2) {
        uint8_t* stg = TILEWIDE ? slab + chunk * kChunkBytes
                                : staging + (kWide ? 2 * g + (int)(wchunk & 1u) : g) * kChunkBytes;
        const int c0 = chunk * CW;             // column inside the tile's output window
        const int ocol0 = ntile * OUT_W + c0;  // global output column
        if constexpr (kWide) {
          ++wchunk;
          if (issuer) bulk_wait_group_read<1>();  // the store issued two chunks ago has left this slab
#pragma unroll
          for (int cc = 0; cc < VPR; ++cc)  // residual of this group's next chunk: in flight under this chunk's work
            rnxt[cc] = rrow != nullptr && chunk + 2 < NCHUNK ? __ldg(reinterpret_cast<const uint4*>(rrow + c0 + 2 * CW) + cc)
                                                             : make_uint4(0u, 0u, 0u, 0u);
          named_bar_sync(1 + g, 128);
        } else if constexpr (!TILEWIDE) {
          // ring mode: one slab per group; wait for its previous store, then fetch this chunk's residual into it
          if (issuer) {
            bulk_wait_group_read<0>();
            if (has_res) {
              mbar_expect_tx(&res_bar[g], kChunkBytes);
              tma_load_4d(stg, &tmR, &res_bar[g], ocol0, w0, h0, n0);
            }
          }
          if (has_res) {
            mbar_wait(&res_bar[g], res_phase);
            res_phase ^= 1;
          } else {
            named_bar_sync(1 + g, 128);
          }
        }
        float v[CW];
        if constexpr (!GEGLU) {
          // Bias / per-clip row bias of this chunk as whole 16-byte vectors, requested BEFORE the accumulator is read so
          // the round trips overlap the TMEM load. (The per-8-column load8h() form compiled into four branchy regions
          // whose loads could not be hoisted: four serialized L1/L2 round trips, ~1 100 clk per 32-column chunk in the
          // clock64 trace -- the epilogue, not the main loop, paced every K <= 640 GEMM.)
          const bool whole = ocol0 + CW <= p.n_out;
          const bool bias_fast = whole && p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
          const bool rb_fast = whole && rb != nullptr && (reinterpret_cast<uintptr_t>(rb) & 15) == 0;
          uint4 bvec[VPR], rvec[VPR];
          if (bias_fast) {
#pragma unroll
            for (int cc = 0; cc < VPR; ++cc) bvec[cc] = __ldg(reinterpret_cast<const uint4*>(p.bias + ocol0) + cc);
          }
          if (rb_fast) {
#pragma unroll
            for (int cc = 0; cc < VPR; ++cc) rvec[cc] = __ldg(reinterpret_cast<const uint4*>(rb + ocol0) + cc);
          }
          if (etr && chunk == g) stamp(local, 11);
#pragma unroll
          for (int s = 0; s < CW; s += 32) {
            uint32_t acc[32];
            tmem_ld32(taddr + c0 + s, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[s + j] = __uint_as_float(acc[j]);
          }
          if (etr && chunk == g) stamp(local, 12);
          if (chunk + 2 >= NCHUNK) {  // last chunk of this warp: the accumulator can be handed back to the MMA warp
            tc_fence_before();
            if (lane == 0) release_acc(buf);
          }
          auto add_vec = [&](const uint4 (&src)[VPR]) {
#pragma unroll
            for (int cc = 0; cc < VPR; ++cc) {
              const __half2* hh = reinterpret_cast<const __half2*>(&src[cc]);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __half22float2(hh[t]);
                v[cc * 8 + 2 * t] += f.x;
                v[cc * 8 + 2 * t + 1] += f.y;
              }
            }
          };
          if (bias_fast) add_vec(bvec);
          if (rb_fast) add_vec(rvec);
          if ((p.bias != nullptr && !bias_fast) || (rb != nullptr && !rb_fast)) {  // ragged last tile / unaligned vectors
            const bool bias_slow = p.bias != nullptr && !bias_fast, rb_slow = rb != nullptr && !rb_fast;
#pragma unroll
            for (int j = 0; j < CW; ++j) {  // static indices: v must stay in registers
              const int col = ocol0 + j;
              if (col < p.n_out) {
                if (bias_slow) v[j] += __half2float(p.bias[col]);
                if (rb_slow) v[j] += __half2float(rb[col]);
              }
            }
          }
        } else {
          constexpr int HALF = BN / 2;
#pragma unroll
          for (int s = 0; s < CW; s += 32) {
            uint32_t hacc[32], gacc[32];
            tmem_ld32(taddr + c0 + s, hacc);
            tmem_ld32(taddr + HALF + c0 + s, gacc);
            tmem_ld_wait();
            // interleaved bias order: [128 hidden | 128 gate] per tile; this chunk's slices were staged by gbias_fetch
            const uint4* bsm = reinterpret_cast<const uint4*>(gb_base + ((buf * 2 + g) * 2 + ((chunk - g) >> 1)) * 64);
#pragma unroll
            for (int j8 = 0; j8 < 32; j8 += 8) {
              float hb[8], gb[8];
              {
                const uint4 hv4 = bsm[j8 >> 3], gv4 = bsm[4 + (j8 >> 3)];  // same address in every lane: broadcast
                const __half2* hh = reinterpret_cast<const __half2*>(&hv4);
                const __half2* gh = reinterpret_cast<const __half2*>(&gv4);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 hf = __half22float2(hh[t]), gf = __half22float2(gh[t]);
                  hb[2 * t] = hf.x;
                  hb[2 * t + 1] = hf.y;
                  gb[2 * t] = gf.x;
                  gb[2 * t + 1] = gf.y;
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float hv = __uint_as_float(hacc[j8 + j]) + hb[j];
                const float gv = __uint_as_float(gacc[j8 + j]) + gb[j];
                v[s + j8 + j] = hv * gelu_fast_f(gv);
              }
            }
          }
          if (chunk + 2 >= NCHUNK) {
            tc_fence_before();
            if (lane == 0) release_acc(buf);
          }
        }
        if (etr && chunk == g) stamp(local, 13);
        // ---- own row of the slab: add the TMA-fetched residual, then overwrite it with the fp16 result ----
        // CW=64: 128-byte rows, SWIZZLE_128B (chunk ^ (row & 7)); CW=32: 64-byte rows, SWIZZLE_64B (chunk ^ ((row>>1)&3))
        uint8_t* srow = stg + r * (CW * 2);
        // CW = 32: the four residual vectors of the row are read up front. Read and write of a slot go through the same
        // shared-memory array, so in the per-slot form below the compiler keeps LDS(cc+1) behind STS(cc) and the pass
        // becomes four dependent LDS -> FADD -> STS groups (~490 clk per chunk in the clock64 trace).
        uint4 rpre[CW == 32 ? VPR : 1];
        if constexpr (kWide) {
#pragma unroll
          for (int cc = 0; cc < VPR; ++cc) {
            rpre[cc] = rcur[cc];
            rcur[cc] = rnxt[cc];
          }
        } else if constexpr (CW == 32) {
          if (has_res) {
#pragma unroll
            for (int cc = 0; cc < VPR; ++cc)
              rpre[cc] = *reinterpret_cast<const uint4*>(srow + ((cc ^ ((r >> 1) & 3)) << 4));
          }
        }
#pragma unroll
        for (int cc = 0; cc < VPR; ++cc) {
          const int sw = (CW == 64) ? (cc ^ (r & 7)) : (cc ^ ((r >> 1) & 3));
          uint4* slot = reinterpret_cast<uint4*>(srow + (sw << 4));
          if (has_res) {
            const uint4 u = (CW == 32) ? rpre[CW == 32 ? cc : 0] : *slot;
            const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __half22float2(hh[t]);
              v[cc * 8 + 2 * t] += f.x;
              v[cc * 8 + 2 * t + 1] += f.y;
            }
          }
          if (!GEGLU && p.relu) {
#pragma unroll
            for (int t = 0; t < 8; ++t) v[cc * 8 + t] = fmaxf(v[cc * 8 + t], 0.f);
          }
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __half2 hh = __floats2half2_rn(v[cc * 8 + 2 * t], v[cc * 8 + 2 * t + 1]);
            pk[t] = *reinterpret_cast<uint32_t*>(&hh);
          }
          *slot = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        if (etr && chunk == g) stamp(local, 14);
        if constexpr (TILEWIDE && DS) {
          // first chunk done: the previous tile's store has drained the other slab by now -> fetch the NEXT tile's
          // residual into it; it has the rest of this tile (two chunks, barrier, store, loop turn) to land
          if (chunk == g && issuer && has_res) {
            bulk_wait_group_read<0>();
            if (local + 1 < n_local) request_res(tile_of(local + 1), (local + 1) & 1);
          }
        }
        if constexpr (!TILEWIDE) {
          fence_proxy_async_smem();
          named_bar_sync(1 + g, 128);
          if (issuer) {
            tma_store_4d(&tmD, stg, ocol0, w0, h0, n0);
            bulk_commit_group();
          }
        }
      }
      if constexpr (TILEWIDE) {
        if (etr) stamp(local, 8);
        fence_proxy_async_smem();
        named_bar_sync(1 + g, 128);
        if (etr) stamp(local, 9);
        if (issuer && p.dbg_skip != 3) {
          for (int chunk = g; chunk < NCHUNK; chunk += 2)
            tma_store_4d(&tmD, slab + chunk * kChunkBytes, ntile * OUT_W + chunk * CW, w0, h0, n0);
          bulk_commit_group();
        }
        if (etr) stamp(local, 10);
      }
      if constexpr (GEGLU) {
        // the other parity's slices were last read in the previous tile, which every thread of the group left before this
        // thread passed this tile's top barrier; the next tile's top barrier publishes the write
        if (gb_pre) *gbias_slot((local + 1) & 1) = gb_next;
      }
      if (NCHUNK == 1 && g == 1) {  // this group had no chunk: still release the accumulator
        tc_fence_before();
        if (lane == 0) release_acc(buf);
      }
    }
    if (issuer) bulk_wait_group<0>();
  }

  tc_fence_before();
  // no CTA may exit while a peer can still multicast into its shared memory or signal its barriers
  if constexpr (CS > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_2sm<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v3: pair kernel with a 16-warp epilogue and a dedicated store warp, for the GEMMs whose epilogue sets the pace
// (K <= 1280 linears: attention / temporal / feed-forward out-projections, QKV, 1x1 projections; 160-wide tiles).
//
// What the clock64 trace of the v2 epilogue showed on the 73728x320->320 residual GEMM (profiles/r01_gemm_epilogue_
// trace.txt): a 128x160 tile took 5 600 clk of epilogue against 1 800 clk of main loop, and most of it was not arithmetic:
// the thread that issues the TMA stores also waited ~1 100 clk for the previous store to drain before it could request
// the next residual tile, ~500 clk to issue three stores, ~400 clk of tile-coordinate divisions, ~300 clk for the bias
// round trip, and group 0 carried three of the five column chunks. Here:
//   * warp 2 is a store warp: it alone issues TMA stores, waits for their drain and requests the residual tile of the
//     tile after next; the 16 epilogue warps never touch the TMA queue;
//   * 16 epilogue warps (4 per TMEM lane quarter) own 40 columns each, read them with ONE pass (tcgen05.ld x32 + x8) and
//     hand the accumulator back immediately, so all TMEM reads of a tile (1 280 clk at 64 B/clk) are in flight at once and
//     the next main loop starts under the arithmetic;
//   * the bias vector lives in shared memory (fp16, read as broadcast 16-byte vectors); bias and residual are added with
//     the mixed-precision add (FHADD: fp32 + fp16 in one instruction, no conversion);
//   * epilogue threads need no tile coordinates at all (TMA clips the store and zero-fills the residual): the N tile
//     index is advanced by increments.
// Roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 store / residual, warp 3 idle, warps 4..19 epilogue.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kP2Threads = 640;
constexpr int kP2BN = 160;
constexpr int kP2Chunks = 5;                                // 32-column TMA boxes per tile (SWIZZLE_64B, 64-byte rows)
constexpr int kP2ChunkBytes = kBlockM * 32 * 2;             // 8 KB
constexpr int kP2SlabBytes = kP2Chunks * kP2ChunkBytes;     // 40 KB: one fp16 output tile
constexpr int kP2BiasBytes = 8192;                          // n_out <= 4096
constexpr int kP2BTileBytes = (kP2BN / 2) * 128;            // 10 KB: this CTA's half of a 160 x 64 weight tile
constexpr int kP2StageBytes = kABytes + kP2BTileBytes;      // 16 KB activations + 10 KB half weight tile
constexpr int kP2WsKBlocks = 5;                             // weight-stationary mode: K <= 320
constexpr int kP2RowStatBytes = 2 * kBlockM * 8;            // folded LayerNorm: (rstd, -mean * rstd) per row, 2 buffers
constexpr int kP2TileVecBytes = 2 * 3 * kP2BN * 2;          // folded LayerNorm: bias | row bias | wsum slices of the tile
template <int STAGES, bool WS, int SLABS = 2, int KB = 1>
constexpr int pair160_smem_bytes() {
  return (WS ? STAGES * kABytes + kP2WsKBlocks * kP2BTileBytes : STAGES * KB * kP2StageBytes) + SLABS * kP2SlabBytes +
         kP2BiasBytes + kP2RowStatBytes + kP2TileVecBytes + 384;
}

// WS (weight-stationary, K <= 320): these GEMMs are bound by what the SM can pull through TMA (knock-outs in
// profiles/r02_gemm_pair160_knockouts.txt: time follows bytes moved per tile: 133 KB operands + 40 KB residual + 40 KB
// store; removing the MMAs changes nothing). The weight half-tile of a CTA (all of K: <= 50 KB) therefore stays resident
// in shared memory and only activations stream: 82 KB instead of 133 KB of operands per tile. To keep one N tile per
// cluster for as long as possible, the (N tile, M pair) units are walked N-major and each cluster takes a contiguous
// range; the weights are reloaded only where a range crosses into the next N tile.
//
// CL = 4 (two pairs per cluster, A multicast): every GEMM of this family that does not sit on the DRAM floor sits on the
// L2 -> SM throughput cap instead (~6 300 B/clk for the whole chip = ~43 B/clk per SM; ncu: lts__throughput 38-43 % on
// all of them, 601 MB of L2 traffic in 96 400 clk on the 73728x320->960 QKV projection, of which only 189 MB are
// algorithmic): a 256x160 pair tile re-reads its activation rows once per N tile. Here the two pairs of a 4-CTA cluster
// take the SAME 256 rows and two ADJACENT N tiles; each CTA fetches half of the 128-row activation box it shares with
// its counterpart in the other pair and TMA-multicasts it to both, so the activation bytes leaving L2 halve
// (operand bytes per cluster and K block: 2 x (32 + 20) KB -> 32 + 40 KB). A stage may be refilled when the MMAs of BOTH
// pairs have read it (empty barrier count 2, commits multicast to all four CTAs). Linear layers with an even number of
// N tiles only; 4-CTA clusters fit 132 of the 148 SMs (33 clusters: ivv_debug_cl4_clusters()).
// RESULT (B200, profiles/r02_linear_ab_cl4_geglu.txt): correct, and 8-15 % SLOWER on every shape it applies to
// (QKV 73728x320->960 59.3 -> 65.1 us, 18432x640->1920 50.9 -> 56.2, 4608x1280->3840 43.9 -> 50.7): 24 % fewer bytes
// out of L2 buy nothing, the time follows the number of SMs at work (132 / 148). So the L2 -> SM throughput is NOT what
// bounds these GEMMs; what a CTA pair can keep in flight is (the ring-latency law). Opt-in.
//
// AS (activation-stationary, K <= 320, >= 3 N tiles): with K = 320 a tile needs all five stages of the ring, so stage s of
// tile i + 1 can only be requested when the MMAs of stage s of tile i have retired, and every tile pays a full load round
// trip (~4 000 clk under load against 1 600 clk of MMAs: the ring-latency law). The weight-stationary mode above does not
// change that (the activation stream still needs the whole ring per tile), which is why it measured equal. Here a cluster
// takes ALL N tiles of an M pair: the pair's 2 x 128 activation rows (5 x 16 KB per CTA, the activation parts of the five
// stages) are loaded once per M pair and stay put, and only the weight half-tiles (10 KB per K block, L2-resident) cycle
// through the five weight parts of the stages. Operand bytes per tile drop from 2 x 130 KB to 2 x 50 KB, the activation
// is read exactly once (no reliance on L2 for the re-reads), and the weight ring holds a whole tile ahead.
//
// SLABS = 1 (K >= 640): these GEMMs are paced by bytes in flight / load round trip (the ring-latency law), and a tile lasts
// two or four ring revolutions while a slab is only occupied for ~3 250 clk (written, stored, drained): ONE output slab is
// enough and the 40 KB it frees hold a sixth pipeline stage (+20 % bytes in flight). The slab-state barriers keep
// alternating with the tile parity (slab_full / slab_free / res_full [local & 1]); only the memory is shared, so the slab
// of tile i + 1 (residual fetch or "free") is released when the store of tile i has drained, not of tile i - 1.
//
// KB = 2 (K a multiple of 128): a ring slot holds TWO 64-channel K blocks (two swizzle atoms per operand row, loaded by
// two TMA boxes each) behind ONE full / empty barrier pair: half as many barrier round trips, waits, commits and
// producer wake-ups per tile for the same bytes in flight (3 slots x 52 KB = 6 x 26 KB).
template <int STAGES, bool WS, int CL = 2, bool AS = false, int SLABS = 2, int KB = 1>
__global__ void __launch_bounds__(kP2Threads, 1)
gemm_tc_pair160_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
                       const __grid_constant__ GemmKParams p, int n_tiles, int total_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  static_assert(CL == 2 || (CL == 4 && !WS), "cluster of one pair, or of two pairs sharing the activation rows");
  static_assert(!AS || (!WS && CL == 2), "activation-stationary mode: plain pair clusters");
  static_assert(SLABS == 2 || (SLABS == 1 && !WS && !AS && CL == 2), "single output slab: plain pair clusters");
  static_assert(STAGES <= 6, "barrier block holds six stages");
  static_assert(KB == 1 || (KB == 2 && !WS && !AS && CL == 2), "two K blocks per ring slot: plain pair clusters");
  constexpr bool C4 = CL == 4;
  constexpr uint32_t kAccStride = 256;  // TMEM columns per accumulator buffer
  constexpr int kOperandBytes = WS ? STAGES * kABytes + kP2WsKBlocks * kP2BTileBytes : STAGES * KB * kP2StageBytes;
  uint8_t* b_res = smem + STAGES * kABytes;  // WS only: [k block][80 weight rows x 128 B]
  uint8_t* slabs = smem + kOperandBytes;
  __half* sbias = reinterpret_cast<__half*>(slabs + SLABS * kP2SlabBytes);
  float2* smr = reinterpret_cast<float2*>(slabs + SLABS * kP2SlabBytes + kP2BiasBytes);  // [2][128] (rstd, -mean * rstd)
  __half* tvec = reinterpret_cast<__half*>(slabs + SLABS * kP2SlabBytes + kP2BiasBytes + kP2RowStatBytes);  // [2][3][160]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slabs + SLABS * kP2SlabBytes + kP2BiasBytes + kP2RowStatBytes +
                                                   kP2TileVecBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint64_t* slab_full = tmem_empty_bar + 2;      // [2] all 16 epilogue warps have written the slab
  uint64_t* slab_free = slab_full + 2;           // [2] the slab's store has been read out (GEMMs without residual)
  uint64_t* res_full = slab_free + 2;            // [2] the residual tile has landed in the slab
  uint64_t* b_full = res_full + 2;               // WS: resident weights landed (leader counts both CTAs' bytes)
  uint64_t* b_free = b_full + 1;                 // WS: every MMA that read the resident weights has completed
  uint64_t* stat_full = b_free + 1;              // [2] folded LayerNorm: the row statistics of the tile are in smr
  uint64_t* a_full = stat_full + 2;              // AS [5]: resident activation slice landed (leader counts both CTAs' bytes)
  uint64_t* a_empty = a_full + 5;                // AS [5]: the last N tile of the M pair has read the slice
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_empty + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cl_rank = (int)cluster_ctarank();
  const int crank = C4 ? (cl_rank & 1) : cl_rank;                    // rank inside the MMA pair (0 = leader)
  const int pair_id = C4 ? (cl_rank >> 1) : 0;                       // CL = 4: which of the cluster's two pairs
  const uint32_t lead_rank = C4 ? (uint32_t)(cl_rank & ~1) : 0u;     // cluster rank of this pair's leader
  const uint16_t kMask = C4 ? (uint16_t)(3u << (2 * pair_id)) : (uint16_t)3;  // the two CTAs of this pair
  const uint16_t kAllMask = C4 ? (uint16_t)0xF : kMask;              // every CTA that fills stages this pair reads
  const int n_clusters = (int)num_clusters_x();
  const int cid = (int)cluster_id_x();
  const int groups = total_tiles / n_tiles;  // M-tile pairs
  // units of this pair: WS -> contiguous range of the N-major order; else round robin over the M-major order (CL = 4:
  // the cluster takes units 2s and 2s + 1 -- same rows, adjacent N tiles, n_tiles is even -- one per pair)
  // AS: the cluster's units are numbered locally: unit u = (its j-th M pair = cid + j * n_clusters, N tile u % n_tiles)
  const int as_pairs = AS && cid < groups ? (groups - cid + n_clusters - 1) / n_clusters : 0;
  const int u_begin = WS ? (int)((long long)cid * total_tiles / n_clusters) : AS ? 0 : C4 ? 2 * cid + pair_id : cid;
  const int u_end = WS ? (int)((long long)(cid + 1) * total_tiles / n_clusters) : AS ? as_pairs * n_tiles : total_tiles;
  const int u_step = (WS || AS) ? 1 : C4 ? 2 * n_clusters : n_clusters;
  const int its_per_tile = p.taps * p.kblocks;
  const bool has_res = p.residual != nullptr && p.dbg_skip != 4;  // dbg_skip 4 (tuning): residual term dropped
  // tuning trace (tools/gemm_trace.py): CTA 0, per tile: 0-1 producer (first / last load issued), 2-3 MMA thread
  // (accumulator free, tile committed), 4-8 epilogue warp 4 (top, accumulator seen, in registers, slab ready, written),
  // 9-12 store warp (slab full seen, stores issued, drained, next residual requested)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int t_, int slot) {
    if (tracing && t_ < 32) p.trace[t_ * 16 + slot] = clock64();
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader expects the bytes of both CTAs
      mbar_init(&empty_bar[s], C4 ? 2 : 1);  // CL = 4: the MMAs of both pairs read (part of) what this CTA loads
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 32);  // one arrive per epilogue warp of both CTAs
      mbar_init(&slab_full[b], 16);
      mbar_init(&slab_free[b], 1);
      mbar_init(&res_full[b], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_free, 1);
    mbar_init(&stat_full[0], 1);
    mbar_init(&stat_full[1], 1);
    for (int s = 0; s < 5; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<2 * kAccStride>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_sync();  // PDL: everything above overlapped the previous kernel's tail

  // unit -> (N tile, M pair); single-thread roles only (the epilogue warps advance the N tile by increments)
  auto decode = [&](int u, int& ntile, int& mg) {
    if (WS) {
      ntile = u / groups;
      mg = u - ntile * groups;
    } else if (AS) {
      const int j = u / n_tiles;
      ntile = u - j * n_tiles;
      mg = cid + j * n_clusters;
    } else {
      ntile = u % n_tiles;
      mg = u / n_tiles;
    }
  };
  // pixel-box origin of this CTA's 128 rows of M pair mg
  auto origin = [&](int mg, int& w0, int& h0, int& n0) {
    const int mtile = mg * 2 + crank;
    if (p.tiles_h == 1 && p.tiles_g == 1) {  // linear layers: a strip of rows
      w0 = mtile * p.bw;
      h0 = 0;
      n0 = mtile < p.tiles_w ? 0 : p.NI;
    } else {
      const int tw = mtile % p.tiles_w;
      const int th = (mtile / p.tiles_w) % p.tiles_h;
      const int tg = mtile / (p.tiles_w * p.tiles_h);
      w0 = tw * p.bw;
      h0 = th * p.bh;
      n0 = mtile < p.tiles_w * p.tiles_h * p.tiles_g ? tg * p.bn : p.NI;  // ghost tile of an odd pair: out of bounds
    }
  };

  if (warp == 0) {
    if (role_elect()) {
      // ===== TMA producer: own 128 activation rows + own half of the weight tile; bytes counted on the leader =====
      int stage = 0;
      uint32_t phase = 0;
      int plocal = 0;
      int cur_nt = -1;
      uint32_t bfree_par = 0;
      for (int u = u_begin; u < u_end; u += u_step, ++plocal) {
        int ntile, mg, w0, h0, n0;
        decode(u, ntile, mg);
        origin(mg, w0, h0, n0);
        if constexpr (WS) {
          if (ntile != cur_nt) {
            if (cur_nt >= 0) {  // the MMAs of the previous N tile must be done with the resident weights
              mbar_wait(b_free, bfree_par);
              bfree_par ^= 1;
            }
            const uint32_t lead_b = mapa_shared(smem_u32(b_full), lead_rank);
            if (crank == 0) mbar_expect_tx(b_full, 2 * its_per_tile * kP2BTileBytes);
            for (int it = 0; it < its_per_tile; ++it)
              tma_load_3d_2sm(b_res + it * kP2BTileBytes, &tmB, lead_b, it * kBlockK, ntile * kP2BN + crank * (kP2BN / 2), 0);
            cur_nt = ntile;
          }
        }
        for (int it = 0; it < its_per_tile; it += KB) {
          if (it == 0 || it >= its_per_tile - KB) stamp(plocal, it == 0 ? 0 : 1);
          const int tap = it / p.kblocks;
          const int kb = it - tap * p.kblocks;
          int dy = 0, dx = 0;
          if (p.taps > 1) {
            dy = tap / p.tap_w - (p.tap_h >> 1);
            dx = tap % p.tap_w - (p.tap_w >> 1);
          }
          if constexpr (AS) {
            if (ntile == 0) {  // first N tile of an M pair: (re)load the resident activation slice of this K block
              const int j = u / n_tiles;
              mbar_wait(&a_empty[it], (uint32_t)(j & 1) ^ 1u);
              const uint32_t lead_a = mapa_shared(smem_u32(&a_full[it]), lead_rank);
              if (crank == 0) mbar_expect_tx(&a_full[it], 2 * kABytes);
              tma_load_4d_2sm(smem + it * kP2StageBytes, &tmA, lead_a, kb * kBlockK, w0 + dx, h0 + dy, n0);
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), lead_rank);
          if constexpr (AS) {
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * kP2BTileBytes);
            tma_load_3d_2sm(smem + stage * kP2StageBytes + kABytes, &tmB, lead_bar, kb * kBlockK,
                            ntile * kP2BN + crank * (kP2BN / 2), tap);
          } else if constexpr (WS) {
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * kABytes);
            tma_load_4d_2sm(smem + stage * kABytes, &tmA, lead_bar, kb * kBlockK, w0 + dx, h0 + dy, n0);
          } else if constexpr (KB == 2) {
            // (linear layers, K % 128 == 0: the host guarantees taps == 1 and an even number of K blocks)
            uint8_t* sa = smem + stage * (2 * kP2StageBytes);
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 4 * kP2StageBytes);
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              tma_load_4d_2sm(sa + sub * kP2StageBytes, &tmA, lead_bar, (kb + sub) * kBlockK, w0, h0, n0);
              tma_load_3d_2sm(sa + sub * kP2StageBytes + kABytes, &tmB, lead_bar, (kb + sub) * kBlockK,
                              ntile * kP2BN + crank * (kP2BN / 2), 0);
            }
          } else {
            uint8_t* sa = smem + stage * kP2StageBytes;
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * kP2StageBytes);
            if constexpr (C4) {
              // linear layers only (w0 = row index): tmA's box is HALF the 128 rows; this CTA's half goes to itself and
              // to the CTA of the same pair rank in the other pair, whose own half arrives the same way
              tma_load_4d_2sm_multicast(sa + pair_id * (kABytes / 2), &tmA, &full_bar[stage], kb * kBlockK,
                                        w0 + pair_id * (kBlockM / 2), h0, n0, (uint16_t)(5u << crank));
            } else {
              tma_load_4d_2sm(sa, &tmA, lead_bar, kb * kBlockK, w0 + dx, h0 + dy, n0);
            }
            tma_load_3d_2sm(sa + kABytes, &tmB, lead_bar, kb * kBlockK, ntile * kP2BN + crank * (kP2BN / 2), tap);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (AS) {
          // the M switch is the one place where this mode waits for DRAM (trace: ~5 900 clk against ~3 100 per tile): pull the
          // NEXT M pair's rows into L2 while this one is being worked on
          if (ntile == 0 && p.as_prefetch && mg + n_clusters < groups) {
            int w1, h1, n1;
            origin(mg + n_clusters, w1, h1, n1);
            for (int kb = 0; kb < p.kblocks; ++kb) tma_prefetch_l2_4d(&tmA, kb * kBlockK, w1, h1, n1);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0 && role_elect()) {
      // ===== MMA issuer: one tcgen05.mma.cta_group::2 (M = 256, N = 160, K = 16) per 32 bytes of K =====
      constexpr uint32_t idesc = umma_idesc_f16(2 * kBlockM, kP2BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      int cur_nt = -1;
      uint32_t bfull_par = 0;
      for (int u = u_begin; u < u_end; u += u_step, ++local) {
        const int buf = local & 1;
        if constexpr (WS) {
          int ntile, mg;
          decode(u, ntile, mg);
          if (ntile != cur_nt) {
            mbar_wait(b_full, bfull_par);
            bfull_par ^= 1;
            cur_nt = ntile;
          }
        }
        const int as_j = AS ? u / n_tiles : 0;           // AS: M pair index of this cluster, N tile inside it
        const int as_nt = AS ? u - as_j * n_tiles : 0;
        mbar_wait(&tmem_empty_bar[buf], ((local >> 1) & 1) ^ 1);  // both CTAs' epilogue warps have read this buffer
        tc_fence_after();
        stamp(local, 2);
        const uint32_t tacc = tmem_base + buf * kAccStride;
        for (int it = 0; it < its_per_tile; it += KB) {
          if constexpr (AS) {
            if (as_nt == 0) mbar_wait(&a_full[it], (uint32_t)(as_j & 1));  // resident activation slice of this M pair
          }
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0) stamp(local, 13);                      // first / last operand stage of the tile seen
          if (it >= its_per_tile - KB) stamp(local, 14);
#pragma unroll
          for (int sub = 0; sub < KB; ++sub) {
            const uint32_t sa = smem_u32(smem + (stage * KB + sub) * (WS ? kABytes : kP2StageBytes));
            const uint64_t adesc = umma_desc_kmajor_sw128(AS ? smem_u32(smem + it * kP2StageBytes) : sa);
            const uint64_t bdesc = umma_desc_kmajor_sw128(WS ? smem_u32(b_res + it * kP2BTileBytes) : sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              if (p.dbg_skip == 1 && (it | sub | k) != 0) continue;  // tuning only: one MMA per tile
              umma_f16_ss_2sm(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | sub | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2sm(&empty_bar[stage], kAllMask);
          if constexpr (AS) {
            if (as_nt == n_tiles - 1) umma_commit_2sm(&a_empty[it], kMask);  // last reader of the slice: it may be replaced
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full_bar[buf], kMask);
        if constexpr (WS) {
          // last unit of this N tile: once these MMAs retire, both producers may overwrite the resident weights
          const int un = u + u_step;
          if (un < u_end && un / groups != cur_nt) umma_commit_2sm(b_free, kMask);
        }
        stamp(local, 3);
      }
    }
  } else if (warp == 2) {
    // ===== store warp: slab -> global (TMA store), then recycle the slab: fetch the residual tile of the tile after
    // next into it (residual GEMMs) or declare it free =====
    const bool lead = role_elect();
    // s_: tile parity (which res_full barrier); the slab is s_ with two slabs, the only one otherwise
    auto request_res = [&](int u_, int s_) {
      int ntile_, mg_, w0_, h0_, n0_;
      decode(u_, ntile_, mg_);
      origin(mg_, w0_, h0_, n0_);
      mbar_expect_tx(&res_full[s_], kP2SlabBytes);
      for (int chunk = 0; chunk < kP2Chunks; ++chunk)
        tma_load_4d(slabs + (SLABS == 2 ? s_ : 0) * kP2SlabBytes + chunk * kP2ChunkBytes, &tmR, &res_full[s_],
                    ntile_ * kP2BN + chunk * 32, w0_, h0_, n0_);
    };
    if (lead && has_res) {
      if (u_begin < u_end) request_res(u_begin, 0);
      if (SLABS == 2 && u_begin + u_step < u_end) request_res(u_begin + u_step, 1);
    }
    int local = 0;
    for (int u = u_begin; u < u_end; u += u_step, ++local) {
      const int s = local & 1;
      mbar_wait(&slab_full[s], (local >> 1) & 1);
      if (lead) {
        int ntile, mg, w0, h0, n0;
        decode(u, ntile, mg);
        origin(mg, w0, h0, n0);
        stamp(local, 9);
        if (p.dbg_skip != 3) {
          for (int chunk = 0; chunk < kP2Chunks; ++chunk)
            tma_store_4d(&tmD, slabs + (SLABS == 2 ? s : 0) * kP2SlabBytes + chunk * kP2ChunkBytes,
                         ntile * kP2BN + chunk * 32, w0, h0, n0);
          bulk_commit_group();
          stamp(local, 10);
          bulk_wait_group_read<0>();  // only this thread waits for the drain
        }
        stamp(local, 11);
        // the drained slab goes to the tile after next (two slabs) or to the next tile (one slab: the other parity)
        const int nxt = u + SLABS * u_step;
        const int s_nxt = SLABS == 2 ? s : s ^ 1;
        if (has_res) {
          if (nxt < u_end) request_res(nxt, s_nxt);
        } else {
          mbar_arrive(&slab_free[s_nxt]);
        }
        stamp(local, 12);
      }
      __syncwarp();
    }
    if (lead) bulk_wait_group<0>();
  } else if (warp == 3) {
    // ===== statistics warp (folded LayerNorm only): turns the partial row sums the producer GEMM left in global memory
    // into (rstd, -mean * rstd) per row of the tile and stages the tile's bias / row-bias / wsum slices, up to two tiles
    // ahead of the epilogue. (Done by the store warp between its stores and drains this took ~5 000 clk per tile and
    // paced the kernel: profiles/r02_gemm_lnfold_trace.txt) =====
    if (p.ln_stats != nullptr) {
      const bool lead = role_elect();
      // statistics of the 128 rows of unit u_ -> smr[b_]; the partial sums are added in a fixed order (deterministic)
      auto prepare_stats = [&](int u_, int b_) {
        int ntile_, mg_, w0_, h0_, n0_;
        decode(u_, ntile_, mg_);
        origin(mg_, w0_, h0_, n0_);
        // lane l owns rows l, l + 32, l + 64, l + 96. ln_parts is a multiple of 4 (one pair per 40 producer columns,
        // 160-wide producer tiles): 16-byte loads, sixteen in flight at a time. (A dependent load -> add chain here cost an
        // L2 round trip per partial and made the consumer GEMM wait for its statistics: 54 -> 116 us on the QKV GEMM.)
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        const int n4 = p.ln_parts >> 1;
        const bool tile_ok = n0_ == 0;
        for (int i0 = 0; i0 < n4; i0 += 4) {
          float4 t[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long row = (long long)w0_ + lane + 32 * j;  // linear layers only: w0 is the row index
            const float4* src = reinterpret_cast<const float4*>(p.ln_stats + row * p.ln_parts);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              t[j][i] = (tile_ok && row < p.W && i0 + i < n4) ? __ldg(src + i0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // fixed order: deterministic
              s1[j] += t[j][i].x;
              s2[j] += t[j][i].y;
              s1[j] += t[j][i].z;
              s2[j] += t[j][i].w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float mean = s1[j] * p.ln_inv_c;
          const float var = fmaxf(s2[j] * p.ln_inv_c - mean * mean, 0.f);
          const float rstd = rsqrtf(var + p.ln_eps);
          smr[b_ * kBlockM + lane + 32 * j] = make_float2(rstd, -mean * rstd);
        }
        // bias, per-frame row bias and wsum slices of this tile -> shared memory. (Read straight from global memory in the
        // epilogue they cost an L2 round trip each - there is next to no L1 beside 225 KB of shared memory - and the
        // 40-column pass took 3 000-4 700 clk instead of 1 600: profiles/r02_gemm_lnfold_trace.txt)
        if (lane < kP2BN / 8) {
          const int col0 = ntile_ * kP2BN + lane * 8;
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          uint4* dst = reinterpret_cast<uint4*>(tvec + b_ * 3 * kP2BN) + lane;
          dst[0] = p.bias != nullptr ? __ldg(reinterpret_cast<const uint4*>(p.bias + col0)) : z;
          uint4 rv = z;
          if (p.rowbias != nullptr && n0_ == 0) {  // the host guarantees one table row per tile (group % 128 == 0)
            long long g_ = (long long)w0_ / p.rowbias_group;
            if (p.rowbias_mod > 0) g_ %= p.rowbias_mod;
            rv = __ldg(reinterpret_cast<const uint4*>(p.rowbias + g_ * p.rowbias_ld + col0));
          }
          dst[kP2BN / 8] = rv;
          dst[2 * (kP2BN / 8)] = __ldg(reinterpret_cast<const uint4*>(p.ln_wsum + col0));
        }
        __syncwarp();
        if (lead) mbar_arrive(&stat_full[b_]);
      };
      int local = 0;
      for (int u = u_begin; u < u_end; u += u_step, ++local) {
        const int b = local & 1;
        // buffer b was last read by the epilogue of unit local - 2, which arrives on slab_full[b] when it is done
        if (local >= 2) mbar_wait(&slab_full[b], ((local - 2) >> 1) & 1);
        prepare_stats(u, b);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp e = warp - 4; TMEM lane quarter q = warp & 3 (rows 32q .. 32q+31, thread = row), column
    // part = e >> 2 owns columns [40 part, 40 part + 40) of the tile = 16-byte vectors [5 part, 5 part + 5) =====
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const bool has_bias = p.bias != nullptr;
    const bool staged = p.ln_stats != nullptr;  // bias / row bias / wsum slices come from the store warp's per-tile staging
    if (has_bias && !staged) {
      const int nvec = p.n_out >> 3;
      for (int i = threadIdx.x - 128; i < nvec; i += 512)
        reinterpret_cast<uint4*>(sbias)[i] = __ldg(reinterpret_cast<const uint4*>(p.bias) + i);
    }
    named_bar_sync(1, 512);
    // own row of the slab: chunk c = vector / 4 at c * 8 KB, 64-byte rows, SWIZZLE_64B: 16-byte slot (cc ^ ((r >> 1) & 3))
    const uint32_t row_off = (uint32_t)r * 64u;
    const uint32_t sw = (uint32_t)(r >> 1) & 3u;
    uint32_t voff[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      const uint32_t vec = (uint32_t)part * 5u + (uint32_t)v;
      voff[v] = (vec >> 2) * (uint32_t)kP2ChunkBytes + row_off + (((vec & 3u) ^ sw) << 4);
    }
    // N tile of the current unit, advanced by increments (no per-tile division on this path)
    int ntile, mg;
    decode(u_begin < u_end ? u_begin : 0, ntile, mg);
    const int step_n = u_step % n_tiles, step_m = u_step / n_tiles;
    const uint32_t empty_addr0 = mapa_shared(smem_u32(&tmem_empty_bar[0]), lead_rank);
    const uint32_t empty_addr1 = mapa_shared(smem_u32(&tmem_empty_bar[1]), lead_rank);
    int local = 0;
    for (int u = u_begin; u < u_end; u += u_step, ++local) {
      const int buf = local & 1;
      const uint32_t ph = (local >> 1) & 1;
      uint8_t* slab = slabs + (SLABS == 2 ? buf : 0) * kP2SlabBytes;
      uint4 bv[5];
      if (has_bias && !staged) {
        const uint4* bsrc = reinterpret_cast<const uint4*>(sbias + ntile * kP2BN + part * 40);
#pragma unroll
        for (int v = 0; v < 5; ++v) bv[v] = bsrc[v];
      }
      // first row of this CTA's 128 (linear layers: the row index itself; only the row-indexed options use it)
      const int cur_ntile = ntile;
      const long long cur_row0 = (long long)(mg * 2 + crank) * kBlockM;
      const bool cur_valid = mg * 2 + crank < p.tiles_w;
      if constexpr (WS) {
        if (++mg == groups) {
          mg = 0;
          ++ntile;
        }
      } else if constexpr (AS) {
        if (++ntile == n_tiles) {
          ntile = 0;
          mg += n_clusters;
        }
      } else {
        ntile += step_n;
        mg += step_m;
        if (ntile >= n_tiles) {
          ntile -= n_tiles;
          ++mg;
        }
      }
      const bool etr = tracing && warp == 4 && lane == 0;
      if (etr) stamp(local, 4);
      mbar_wait(&tmem_full_bar[buf], ph);
      tc_fence_after();
      if (etr) stamp(local, 5);
      const uint32_t taddr = tmem_base + buf * kAccStride + (static_cast<uint32_t>(q * 32) << 16) + part * 40;
      uint32_t a0[32], a1[8];
      tmem_ld32(taddr, a0);
      tmem_ld8(taddr + 32, a1);
      tmem_ld_wait();
      if (etr) stamp(local, 6);
      tc_fence_before();
      if (lane == 0) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0);  // accumulator handed back right away
      if (has_res) mbar_wait(&res_full[buf], ph);   // residual landed (the fetch was issued after the slab's last store drained)
      else mbar_wait(&slab_free[buf], SLABS == 1 && buf ? ph : ph ^ 1);  // the slab's previous store has been read out
      // (one slab: slab_free[1] first fires after tile 0's store, slab_free[0] after tile 1's -- tile 0 itself waits for nothing)
      if (etr) stamp(local, 7);
      // folded LayerNorm (this GEMM's A rows are the un-normalised x): per-row (rstd, -mean * rstd) from the store warp
      float rstd = 1.f, nmr = 0.f;
      const uint4* tv = reinterpret_cast<const uint4*>(tvec + buf * 3 * kP2BN + part * 40);  // + 20: row bias, + 40: wsum
      if (staged) {
        mbar_wait(&stat_full[buf], ph);
        const float2 mr = smr[buf * kBlockM + r];
        rstd = mr.x;
        nmr = mr.y;
      }
      float st_s = 0.f, st_q = 0.f;
      if (p.dbg_skip != 5) {  // dbg_skip 5 (tuning): no arithmetic, nothing written
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v < 4 ? a0[v * 8 + j] : a1[j]);
          uint4* slot = reinterpret_cast<uint4*>(slab + voff[v]);
          if (staged) {
            // LN(x) W^T = rstd * (x W'^T) + (-mean * rstd) * wsum; the beta term is part of the bias
            const uint4 wv = tv[2 * (kP2BN / 8) + v];
            const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 wf = __half22float2(wh[t]);
              x[2 * t] = fmaf(rstd, x[2 * t], nmr * wf.x);
              x[2 * t + 1] = fmaf(rstd, x[2 * t + 1], nmr * wf.y);
            }
            if (has_bias) add_h8(x, tv[v]);
            if (p.rowbias != nullptr) add_h8(x, tv[kP2BN / 8 + v]);
          } else if (has_bias) {
            add_h8(x, bv[v]);
          }
          if (has_res) {
            const uint4 rr = *slot;
            add_h8(x, rr);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (p.row_stats_out != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              st_s += x[j];
              st_q = fmaf(x[j], x[j], st_q);
            }
          }
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __half2 hh = __floats2half2_rn(x[2 * t], x[2 * t + 1]);
            pk[t] = *reinterpret_cast<uint32_t*>(&hh);
          }
          *slot = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      // statistics of this row over the thread's 40 columns, for the LayerNorm folded into the next GEMM
      if (p.row_stats_out != nullptr && cur_valid && cur_row0 + r < p.W)
        p.row_stats_out[(cur_row0 + r) * (long long)(p.n_out / 40) + cur_ntile * 4 + part] = make_float2(st_s, st_q);
      if (etr) stamp(local, 8);
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store of the store warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&slab_full[buf]);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit while its peer can still signal its barriers or read its TMEM half
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm<2 * kAccStride>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int pow2_floor_div(long long x, int cap) {
  int b = 1;
  while (b * 2 <= cap && (x % (b * 2)) == 0) b *= 2;
  return b;
}
static int pow2_ceil(long long x) {
  int b = 1;
  while (b < x) b *= 2;
  return b;
}

// choose the (bw, bh, bn) pixel box with bw*bh*bn = 128 that minimises the number of M tiles
// halo mode needs the box inside one frame, whole swizzle atoms per box row and (bh + 2) * bw <= kHaloRows
static bool halo_box_ok(int bw, int bh, int bn) { return bn == 1 && (bw % 8) == 0 && (bh + 2) * bw <= kHaloRows; }

static void choose_box(long long W, long long H, long long NI, bool want_halo, int* bw, int* bh, int* bn) {
  long long best = -1;
  for (int cw = 1; cw <= 128; cw *= 2) {
    for (int ch = 1; cw * ch <= 128; ch *= 2) {
      const int cn = 128 / (cw * ch);
      const long long tiles = ((W + cw - 1) / cw) * ((H + ch - 1) / ch) * ((NI + cn - 1) / cn);
      // prefer fewer tiles; tie -> (3x3 convolutions) a box the halo kernel takes, then wider rows (longer contiguous
      // runs in memory)
      const bool halo_new = want_halo && halo_box_ok(cw, ch, cn);
      const bool halo_old = best >= 0 && want_halo && halo_box_ok(*bw, *bh, *bn);
      if (best < 0 || tiles < best || (tiles == best && (halo_new > halo_old || (halo_new == halo_old && cw > *bw)))) {
        best = tiles;
        *bw = cw;
        *bh = ch;
        *bn = cn;
      }
    }
  }
  (void)pow2_floor_div;
  (void)pow2_ceil;
}

static int sm_count() {  // of the current device (cached per device ordinal)
  static int n[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (n[dev] == 0 && cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
  return n[dev];
}

template <int BN, int STAGES, int CW, bool GEGLU, bool TILEWIDE, int CS, bool TWO = false, bool HALO = false,
          bool DS = false, bool AS = false>
static int launch_persistent_cs(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                                const CUtensorMap& tmR, const GemmKParams& kp, int m_tiles, int n_tiles,
                                cudaStream_t stream) {
  constexpr int smem = persist_smem_bytes<BN, STAGES, CW, GEGLU, TILEWIDE, TWO, HALO, DS>();
  static_assert(smem <= 227 * 1024, "persistent GEMM configuration exceeds shared memory");
  auto kern = gemm_tc_persistent_kernel<BN, STAGES, CW, GEGLU, TILEWIDE, CS, TWO, HALO, DS, AS>;
  static DeviceOnce configured;
  if (configured.first()) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const int groups = (m_tiles + CS - 1) / CS;
  const int total = groups * n_tiles;  // (super) tiles
  int clusters = sm_count() / CS;
  if (clusters > (AS ? groups : total)) clusters = AS ? groups : total;  // AS: a cluster owns whole M pairs
  if (kp.ws_stages > 0) clusters = (sm_count() / n_tiles) * n_tiles;  // weight-stationary: one N tile per CTA
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * CS));
  cfg.blockDim = dim3(kPersistThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  IVV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmD, tmR, kp, n_tiles, total));
  return 0;
}

template <int STAGES, bool WS, bool AS = false, int SLABS = 2, int KB = 1>
static int launch_pair160(const CUtensorMap& tmA, const CUtensorMap& tmB2, const CUtensorMap& tmD, const CUtensorMap& tmR,
                          const GemmKParams& kp, int m_tiles, int n_tiles, cudaStream_t stream) {
  constexpr int smem = pair160_smem_bytes<STAGES, WS, SLABS, KB>();
  static_assert(smem <= 227 * 1024, "pair160 configuration exceeds shared memory");
  auto kern = gemm_tc_pair160_kernel<STAGES, WS, 2, AS, SLABS, KB>;
  static DeviceOnce configured;
  if (configured.first()) IVV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int groups = (m_tiles + 1) / 2;
  const int total = groups * n_tiles;
  int clusters = sm_count() / 2;
  if (clusters > (AS ? groups : total)) clusters = AS ? groups : total;  // AS: a cluster owns whole M pairs
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cfg.blockDim = dim3(kP2Threads);
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
  IVV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB2, tmD, tmR, kp, n_tiles, total));
  return 0;
}

// 4-CTA clusters (two pairs, activation rows multicast): how many fit on this device at once (4-CTA clusters leave some
// SMs of a GPC unused: 33 clusters = 132 of 148 SMs is what the B200 places). Asked once per device; 0 = unavailable.
template <int STAGES>
static int pair160_cl4_clusters() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] != 0) return cached[dev] > 0 ? cached[dev] : 0;
  constexpr int smem = pair160_smem_bytes<STAGES, false>();
  auto kern = gemm_tc_pair160_kernel<STAGES, false, 4>;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(4 * (sm_count() / 4)));
    cfg.blockDim = dim3(kP2Threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
  }
  (void)cudaGetLastError();
  cached[dev] = n > 0 ? n : -1;
  return n > 0 ? n : 0;
}

// tmA here has a HALF box (64 rows): see the CL = 4 note at the kernel
template <int STAGES>
static int launch_pair160_cl4(const CUtensorMap& tmA, const CUtensorMap& tmB2, const CUtensorMap& tmD,
                              const CUtensorMap& tmR, const GemmKParams& kp, int m_tiles, int n_tiles, int max_clusters,
                              cudaStream_t stream) {
  constexpr int smem = pair160_smem_bytes<STAGES, false>();
  auto kern = gemm_tc_pair160_kernel<STAGES, false, 4>;
  const int groups = (m_tiles + 1) / 2;
  const int total = groups * n_tiles;  // even: n_tiles is
  int clusters = max_clusters;
  if (clusters > total / 2) clusters = total / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * 4));
  cfg.blockDim = dim3(kP2Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 4;
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
  IVV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB2, tmD, tmR, kp, n_tiles, total));
  return 0;
}

// cluster size 2 (weight tile multicast) whenever the M tiles pair up; IVV_CLUSTER=1 disables it (tuning hook)
template <int BN, int STAGES, int CW, bool GEGLU, bool TILEWIDE>
static int launch_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB2, const CUtensorMap& tmB1,
                             const CUtensorMap& tmD, const CUtensorMap& tmR, const GemmKParams& kp, int m_tiles,
                             int n_tiles, int cs, cudaStream_t stream) {
  if (cs == 2)
    return launch_persistent_cs<BN, STAGES, CW, GEGLU, TILEWIDE, 2>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
  return launch_persistent_cs<BN, STAGES, CW, GEGLU, TILEWIDE, 1>(tmA, tmB1, tmD, tmR, kp, m_tiles, n_tiles, stream);
}

template <int BN, int STAGES>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& kp, int m_tiles, int n_tiles,
                  cudaStream_t stream) {
  constexpr int smem = gemm_smem_bytes<BN, STAGES>();
  static DeviceOnce configured;
  if (configured.first()) {
    IVV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  dim3 grid(n_tiles, m_tiles, kp.splits);
  IVV_CHECK_CUDA(launch_pdl(gemm_tc_kernel<BN, STAGES>, grid, dim3(kGemmThreads), smem, stream, tmA, tmB, kp));
  return 0;
}

}  // namespace ivv

namespace ivv {
// shapes the v3 pair kernel (gemm_tc_pair160_kernel) takes: short K, N a multiple of its 160-wide tile, >= 2 row tiles
constexpr int kCl4MinTiles = 148;  // below one wave of pair tiles nothing is throughput-bound
static bool pair160_shape_ok(long long rows, long long k_total, long long n_out) {
  return k_total <= 1280 && (n_out % 160) == 0 && n_out <= 4096 && rows > kBlockM;
}
}  // namespace ivv

namespace ivv {
// Tuning switches of ivv_gemm, read from the environment ONCE per process (getenv is not on the launch path).
// -1 = unset. IVV_HALO / IVV_DS / IVV_EPI2 / IVV_PAIR: 0 disables; IVV_CLUSTER=2, IVV_FORCE_BN=32|64|128|160|256,
// IVV_NO_WS=1, IVV_DEBUG_SKIP=1..5 (knock-outs, results are garbage).
struct GemmEnv {
  int halo, ds, epi2, pair, cluster, force_bn, no_ws, ws, dbg_skip, geglu_ds, cl4, cl4_min, wide, wide_k, as, as_pf, slab1, kb2;
};
static const GemmEnv& gemm_env() {
  static const GemmEnv e = [] {
    auto geti = [](const char* name) {
      const char* v = getenv(name);
      return v ? atoi(v) : -1;
    };
    GemmEnv g{};
    g.halo = geti("IVV_HALO");
    g.ds = geti("IVV_DS");
    g.epi2 = geti("IVV_EPI2");
    g.pair = geti("IVV_PAIR");
    g.cluster = geti("IVV_CLUSTER");
    g.force_bn = geti("IVV_FORCE_BN");
    g.no_ws = geti("IVV_NO_WS");
    g.ws = geti("IVV_WS");
    g.dbg_skip = geti("IVV_DEBUG_SKIP");
    g.geglu_ds = geti("IVV_GEGLU_DS");
    g.cl4 = geti("IVV_CL4");
    g.cl4_min = geti("IVV_CL4_MIN");
    g.wide = geti("IVV_WIDE");
    g.wide_k = geti("IVV_WIDE_K");
    g.as = geti("IVV_AS");
    g.as_pf = geti("IVV_AS_PF");
    g.slab1 = geti("IVV_SLAB1");
    g.kb2 = geti("IVV_KB2");
    return g;
  }();
  return e;
}
}  // namespace ivv

extern "C" int ivv_gemm_ln_fold_ok(int64_t rows, int64_t k, int64_t n_out) {
  const ivv::GemmEnv& env = ivv::gemm_env();
  return ivv::pair160_shape_ok(rows, k, n_out) && (k % 8) == 0 && env.epi2 != 0 && env.pair < 0 && env.cluster < 0 &&
                 env.force_bn < 0
             ? 1
             : 0;
}


// tuning / test hook (not in ivv.h): the pixel box ivv_gemm picks for a [n_img, h, w] activation, and whether the halo
// kernel accepts it (box inside one frame, whole swizzle atoms per box row, (bh + 2) * bw <= 160)
extern "C" int ivv_debug_conv_box(int64_t w, int64_t h, int64_t n_img, int32_t want_halo, int32_t* bw, int32_t* bh,
                                  int32_t* bn) {
  int a = 0, b = 0, c = 0;
  ivv::choose_box(w, h, n_img, want_halo != 0, &a, &b, &c);
  *bw = a;
  *bh = b;
  *bn = c;
  return ivv::halo_box_ok(a, b, c) ? 1 : 0;
}

// tuning hook: how many 4-CTA clusters of the short-K pair kernel the current device runs at once (0 = unavailable)
extern "C" int ivv_debug_cl4_clusters() { return ivv::pair160_cl4_clusters<5>(); }

// tuning / test hook (not in ivv.h): tile width (BLOCK_N) the calling thread's last ivv_gemm chose
static thread_local int g_last_bn = 0, g_last_as = 0;
// test hook (not in ivv.h): while set, the calling thread's ivv_gemm validates its arguments, chooses the pixel box and the
// tile width (ivv_debug_last_gemm_tile) and returns before it touches the device -- the host logic, testable without a GPU
static thread_local bool g_plan_only = false;
extern "C" void ivv_debug_gemm_plan_only(int on) { g_plan_only = on != 0; }
extern "C" int ivv_debug_last_gemm_tile() { return g_last_bn; }
// 1 if the calling thread's last ivv_gemm ran the activation-stationary mode of the short-K pair kernel
extern "C" int ivv_debug_last_gemm_as() { return g_last_as; }

static long long* g_gemm_trace = nullptr;
// tuning only: clock64 trace of CTA 0 of the next persistent-kernel launches into buf ([32 tiles][16] int64); NULL = off
extern "C" void ivv_debug_gemm_trace(void* buf) { g_gemm_trace = reinterpret_cast<long long*>(buf); }

extern "C" int ivv_gemm(const ivv_gemm_args* a, ivv_stream_t stream_) {
  using namespace ivv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  IVV_REQUIRE(a != nullptr, "ivv_gemm: null args");
  int tap_h = a->tap_h, tap_w = a->tap_w;
  if (tap_h == 0 && tap_w == 0) {
    IVV_REQUIRE(a->taps == 1 || a->taps == 9, "ivv_gemm: taps must be 1 or 9 (or tap_h x tap_w given), got %d", a->taps);
    tap_h = tap_w = a->taps == 9 ? 3 : 1;
  }
  IVV_REQUIRE(tap_h > 0 && tap_w > 0 && (tap_h & 1) && (tap_w & 1) && tap_h <= 15 && tap_w <= 15 &&
                  a->taps == tap_h * tap_w,
              "ivv_gemm: bad tap window %d x %d for taps = %d", tap_h, tap_w, a->taps);
  IVV_REQUIRE(!(a->relu && (a->geglu || a->splits > 1)), "ivv_gemm: relu cannot be combined with GEGLU or split-K");
  IVV_REQUIRE(a->a && a->wgt && a->d, "ivv_gemm: null tensor pointer");
  IVV_REQUIRE(a->n_img > 0 && a->h > 0 && a->w > 0 && a->c > 0 && a->n_out > 0, "ivv_gemm: empty problem");
  IVV_REQUIRE(a->a_ld % 8 == 0 && a->w_ld % 8 == 0, "ivv_gemm: a_ld (%lld) and w_ld (%lld) must be multiples of 8",
              (long long)a->a_ld, (long long)a->w_ld);
  IVV_REQUIRE(a->c <= a->a_ld && a->c <= a->w_ld, "ivv_gemm: c exceeds a leading dimension");
  IVV_REQUIRE(!a->geglu || (a->n_out % 256 == 0), "ivv_gemm: GEGLU needs n_out %% 256 == 0");
  IVV_REQUIRE(!(a->geglu && (a->rowbias || a->residual)), "ivv_gemm: GEGLU epilogue takes bias only");

  GemmKParams kp{};
  const GemmEnv& env = gemm_env();
  // 3x3 convolutions on fp16 outputs may take the halo kernel (one activation box per filter column); IVV_HALO=0 disables
  const bool want_halo = a->taps == 9 && tap_h == 3 && tap_w == 3 && !a->geglu && !a->out_f32 && a->splits <= 1 &&
                         env.halo != 0;
  choose_box(a->w, a->h, a->n_img, want_halo, &kp.bw, &kp.bh, &kp.bn);
  kp.halo_bytes = (kp.bh + 2) * kp.bw * 128;
  kp.trace = g_gemm_trace;
  kp.as_prefetch = env.as_pf != 0;  // IVV_AS_PF=0: no L2 prefetch of the next M pair in the activation-stationary modes
  kp.tiles_w = (int)((a->w + kp.bw - 1) / kp.bw);
  kp.tiles_h = (int)((a->h + kp.bh - 1) / kp.bh);
  kp.tiles_g = (int)((a->n_img + kp.bn - 1) / kp.bn);
  kp.W = (int)a->w;
  kp.H = (int)a->h;
  kp.NI = (int)a->n_img;
  kp.kblocks = (int)((a->c + kBlockK - 1) / kBlockK);
  kp.taps = a->taps;
  kp.tap_w = tap_w;
  kp.tap_h = tap_h;
  kp.relu = a->relu;
  kp.n_out = (int)a->n_out;
  kp.geglu = a->geglu;
  kp.out_cols = a->geglu ? (int)(a->n_out / 2) : (int)a->n_out;
  kp.out_f32 = a->out_f32;
  if (env.dbg_skip >= 0) kp.dbg_skip = env.dbg_skip;
  kp.splits = a->splits > 1 ? a->splits : 1;
  kp.split_stride = (long long)a->n_img * a->h * a->w * a->d_ld;
  if (kp.splits > 1) {
    IVV_REQUIRE(a->out_f32 && !a->geglu && !a->bias && !a->rowbias && !a->residual,
                "ivv_gemm: split-K writes raw fp32 partial sums (no epilogue terms, out_f32 = 1)");
    IVV_REQUIRE(kp.splits <= kp.taps * kp.kblocks && kp.splits <= 64, "ivv_gemm: too many splits (%d)", kp.splits);
    // every split must own at least one iteration
    const int all_it = kp.taps * kp.kblocks, per = (all_it + kp.splits - 1) / kp.splits;
    IVV_REQUIRE((kp.splits - 1) * per < all_it, "ivv_gemm: splits (%d) leave an empty K range", kp.splits);
  }
  kp.d = a->d;
  kp.d_ld = a->d_ld;
  kp.bias = reinterpret_cast<const __half*>(a->bias);
  kp.rowbias = reinterpret_cast<const __half*>(a->rowbias);
  kp.rowbias_group = a->rowbias_group > 0 ? a->rowbias_group : 1;
  kp.rowbias_ld = a->rowbias_ld;
  kp.residual = reinterpret_cast<const __half*>(a->residual);
  kp.res_ld = a->res_ld;
  const long long m_tiles_ll = (long long)kp.tiles_w * kp.tiles_h * kp.tiles_g;
  IVV_REQUIRE(m_tiles_ll <= 65535, "ivv_gemm: too many M tiles (%lld)", m_tiles_ll);
  const int m_tiles = (int)m_tiles_ll;
  const bool res_ok = a->residual == nullptr ||
                      ((a->res_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0);
  const bool persistent_ok = !a->out_f32 && (a->d_ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(a->d) & 15) == 0) && res_ok;
  // the halo kernel is a pair-mode persistent kernel with 128/160/256-wide tiles
  const bool halo = want_halo && halo_box_ok(kp.bw, kp.bh, kp.bn) && persistent_ok && m_tiles >= 2 && a->n_out >= 96 &&
                    env.pair != 0 && env.cluster < 0 && env.force_bn < 0;

  // ---- tile-N choice: least padding first, then enough CTAs to fill 148 SMs ----
  int bn_sel;
  if (a->geglu) {
    bn_sel = 256;
  } else {
    // Tile-width choice from a measured cost model (tools/tile_sweep.py): these GEMMs are bound by operand delivery
    // from L2 (~12 TB/s), so time ~ padding x bytes fetched per output (1/BN + 1/128) / wave efficiency on 148 SMs.
    const int cands[5] = {256, 160, 128, 64, 32};
    double best = 1e30;
    bn_sel = 128;
    // halo mode: a pair fetches 2 * (bh + 2) * bw activation rows per three taps instead of 2 * 128 per tap
    const double a_rows = halo ? 2.0 * (kp.bh + 2) * kp.bw / 3.0 : 256.0;
    for (int i = 0; i < (halo ? 3 : 5); ++i) {
      const long long nt = (a->n_out + cands[i] - 1) / cands[i];
      const double waste = (double)nt * cands[i] / (double)a->n_out;
      const long long tiles = nt * m_tiles * (a->splits > 1 ? a->splits : 1);
      const long long waves = (tiles + sm_count() - 1) / sm_count();
      const double eff = (double)tiles / (double)(waves * sm_count());
      const double cost = waste * (a_rows / 256.0 / cands[i] + (m_tiles >= 2 ? 1.0 / 256.0 : 1.0 / 128.0)) / eff;
      if (cost < best * 0.999) {
        best = cost;
        bn_sel = cands[i];
      }
    }
  }
  if (!a->geglu) {  // tuning hook (tools/tile_sweep.py): IVV_FORCE_BN=32|64|128|160|256
    if (env.force_bn >= 0) {
      const int v = env.force_bn;
      if (v == 32 || v == 64 || v == 128 || v == 160 || v == 256) bn_sel = v;
    }
  }
  // Residual GEMMs with short main loops (the attention / temporal out-projections, K = 320 and 640): pair kernel with
  // two staging slabs and 160-wide tiles, so the residual fetch of the next tile overlaps this tile's epilogue.
  // IVV_DS=0 disables (tuning hook).
  const bool ds0 = a->residual != nullptr && !halo && !a->geglu && a->taps == 1 && a->c <= 640 && (a->n_out % 160) == 0 &&
                   persistent_ok && m_tiles >= 2 && env.ds != 0 && env.pair != 0 && env.cluster < 0 && env.force_bn < 0;
  // v3 pair kernel (16-warp epilogue + store warp, 160-wide tiles): every short-K GEMM whose N is a multiple of 160 and
  // that needs no per-row bias. IVV_EPI2=0 falls back to the v2 kernels (tuning hook).
  const bool is_linear = a->taps == 1 && a->h == 1 && a->n_img == 1;
  // the pair kernel takes a row-bias table only together with a folded LayerNorm (the store warp stages one table row per
  // tile), so every 128-row tile must lie inside one group
  const bool rowbias_ok = a->rowbias == nullptr ||
                          (is_linear && a->ln_stats != nullptr && (a->rowbias_group % kBlockM) == 0 &&
                           (a->rowbias_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(a->rowbias) & 15) == 0);
  const bool pair160_0 = pair160_shape_ok(a->n_img * a->h * a->w, (long long)a->c * a->taps, a->n_out) && !halo &&
                         !a->geglu && rowbias_ok && persistent_ok && a->splits <= 1 && m_tiles >= 2 &&
                         (a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0) &&
                         env.epi2 != 0 && env.pair < 0 && env.cluster < 0 && env.force_bn < 0;
  const bool wants_fold = a->row_stats_out != nullptr || a->ln_stats != nullptr || a->rowbias_mod > 0;
  // 320-wide tiles (pair kernel, two N = 160 MMAs per k-step, one accumulator): taken when the tile list then needs fewer
  // rounds of the 74 clusters. Cost model in clocks per (tap, k-block) of a tile, calibrated on B200
  // (profiles/r02_gemm_wide_ab.txt): 950 for the 320-wide tile, 840 / 772 for 256 / 160 (700 in the short-K pair kernel);
  // the 320-wide tile also pays its epilogue in the open (one accumulator): ~4 000 clk.
  // The N = 1280 convolutions and projections of the 8x12 level: 72 tiles = one round instead of 144 = two; the FF
  // out-projection of the 16x24 level: two rounds instead of three (measured equal). From K = 2560: with K = 1280 the wide tile
  // LOSES against the short-K pair kernel (73728x1280->320 72.2 -> 75.1 us, 4608x1280->1280 20.2 -> 22.5 us,
  // profiles/r02_gemm_wide_ab.txt): twenty k-blocks do not amortise an epilogue in the open. IVV_WIDE=0 disables,
  // IVV_WIDE_K=<k> sets the shortest K, IVV_FORCE_BN=320 takes it wherever it is legal (tuning hooks).
  bool use_wide = false;
  if (!a->geglu && !halo && !wants_fold && persistent_ok && res_ok && m_tiles >= 2 && (a->n_out % 320) == 0 &&
      (long long)a->c * a->taps >= (env.force_bn == 320 || env.wide_k <= 0 ? 2560 : env.wide_k) && a->splits <= 1 &&
      env.pair < 0 && env.cluster < 0 && env.no_ws < 0 && (env.force_bn < 0 || env.force_bn == 320) && env.wide != 0) {
    const long long clusters = sm_count() / 2, pairs = (m_tiles + 1) / 2;
    const double its = (double)kp.taps * kp.kblocks;
    const bool alt160 = ds0 || pair160_0;
    auto cost = [&](int bn) {
      const long long nt = (a->n_out + bn - 1) / bn;
      const long long rounds = (pairs * nt + clusters - 1) / clusters;
      const double per_it = bn == 320 ? 950.0 : bn == 256 ? 840.0 : bn == 160 ? (alt160 ? 700.0 : 772.0) : 700.0;
      return (double)rounds * (its * per_it + (bn == 320 ? 4000.0 : 0.0));
    };
    use_wide = env.force_bn == 320 || cost(320) < 0.95 * cost(alt160 ? 160 : bn_sel);
  }
  const bool ds = ds0 && !use_wide, pair160 = pair160_0 && !use_wide;
  IVV_REQUIRE(!wants_fold || (pair160 && is_linear),
              "ivv_gemm: row_stats_out / ln_stats / rowbias_mod need a linear layer served by the short-K pair kernel "
              "(ivv_gemm_ln_fold_ok(rows, k, n_out)); got rows=%lld k=%lld n_out=%lld taps=%d",
              (long long)(a->n_img * a->h * a->w), (long long)a->c, (long long)a->n_out, a->taps);
  if (a->ln_stats != nullptr) {
    IVV_REQUIRE(a->ln_wsum != nullptr && a->ln_parts > 0 && a->ln_parts <= 64 && a->ln_eps > 0.f &&
                    (reinterpret_cast<uintptr_t>(a->ln_wsum) & 15) == 0,
                "ivv_gemm: ln_stats needs ln_wsum (16-byte aligned), 0 < ln_parts <= 64 and ln_eps > 0");
  }
  kp.rowbias_mod = a->rowbias_mod;
  kp.row_stats_out = reinterpret_cast<float2*>(a->row_stats_out);
  kp.ln_stats = reinterpret_cast<const float2*>(a->ln_stats);
  kp.ln_wsum = reinterpret_cast<const __half*>(a->ln_wsum);
  kp.ln_parts = a->ln_parts;
  kp.ln_eps = a->ln_eps;
  kp.ln_inv_c = 1.f / (float)a->c;
  if (ds || pair160) bn_sel = 160;
  if (use_wide) bn_sel = 320;
  const int n_tiles = (int)((a->n_out + bn_sel - 1) / bn_sel);
  g_last_bn = bn_sel;
  g_last_as = 0;
  if (g_plan_only) return 0;

  // ---- tensor maps ----
  CUtensorMap tmA, tmB, tmB2;
  {
    const uint64_t dims[4] = {(uint64_t)a->c, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
    const uint64_t strides[4] = {2, (uint64_t)a->a_ld * 2, (uint64_t)a->a_ld * 2 * a->w,
                                 (uint64_t)a->a_ld * 2 * a->w * a->h};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kp.bw, (uint32_t)(halo ? kp.bh + 2 : kp.bh), (uint32_t)kp.bn};
    if (int rc = make_tmap_f16(&tmA, a->a, 4, dims, strides, box, 128)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->c, (uint64_t)a->n_out, (uint64_t)a->taps};
    const uint64_t strides[3] = {2, (uint64_t)a->w_ld * 2, (uint64_t)a->w_ld * 2 * a->n_out};
    const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)(bn_sel > 256 ? 256 : bn_sel), 1};  // (unused by the 320-wide tiles)
    if (int rc = make_tmap_f16(&tmB, a->wgt, 3, dims, strides, box, 128)) return rc;
    // half tile per CTA of a 2-CTA cluster (320-wide tiles: a quarter per load, two loads per CTA)
    const uint32_t box2[3] = {(uint32_t)kBlockK, (uint32_t)(bn_sel == 320 ? 80 : bn_sel / 2), 1};
    if (int rc = make_tmap_f16(&tmB2, a->wgt, 3, dims, strides, box2, 128)) return rc;
  }

  const bool persistent = persistent_ok && (long long)m_tiles * n_tiles < (1LL << 30);
  if (persistent) {
    const int cw = (a->geglu || pair160) ? 32 : halo ? (bn_sel == 128 ? 64 : 32) : (bn_sel == 256 || bn_sel == 128) ? 64 : 32;
    CUtensorMap tmD, tmR;
    const uint32_t box[4] = {(uint32_t)cw, (uint32_t)kp.bw, (uint32_t)kp.bh, (uint32_t)kp.bn};
    {
      const uint64_t dims[4] = {(uint64_t)kp.out_cols, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
      const uint64_t strides[4] = {2, (uint64_t)a->d_ld * 2, (uint64_t)a->d_ld * 2 * a->w,
                                   (uint64_t)a->d_ld * 2 * a->w * a->h};
      if (int rc = make_tmap_f16(&tmD, a->d, 4, dims, strides, box, cw == 64 ? 128 : 64)) return rc;
    }
    if (a->residual) {
      const uint64_t dims[4] = {(uint64_t)kp.out_cols, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
      const uint64_t strides[4] = {2, (uint64_t)a->res_ld * 2, (uint64_t)a->res_ld * 2 * a->w,
                                   (uint64_t)a->res_ld * 2 * a->w * a->h};
      if (int rc = make_tmap_f16(&tmR, a->residual, 4, dims, strides, box, cw == 64 ? 128 : 64)) return rc;
    } else {
      tmR = tmD;
    }
    const long long k_total = (long long)a->c * a->taps;
    // 2-CTA clusters with a TMA-multicast weight tile: measured neutral (tools/profile_ops.py with IVV_CLUSTER=2) — the
    // limit is the ~60 B/clk each SM can ingest, which multicast does not reduce — so it is opt-in.
    int cs = 1;
    if (env.cluster >= 0) cs = (env.cluster == 2 && m_tiles >= 2) ? 2 : 1;
    // Weight-stationary mode for short-K GEMMs with many M tiles (the K = 320 linears of the 32x48 level): the weight
    // tile of a CTA (all of K) stays in shared memory, only activations stream -> 2.25x less operand ingest per tile.
    {
      const int stages = bn_sel == 256 ? 3 : bn_sel == 160 ? 5 : 6;  // stage counts of the single-CTA configs below
      const long long region = (long long)stages * (kABytes + bn_sel * 128);
      const long long b_res = (long long)kp.taps * kp.kblocks * bn_sel * 128;
      const long long a_slots = (region - b_res) / kABytes;
      const bool ws_ok = !a->geglu && bn_sel >= 128 && a_slots >= 4 && n_tiles <= sm_count() / 2 &&
                         (long long)m_tiles * n_tiles >= 4LL * sm_count() && env.no_ws < 0;
      kp.ws_stages = ws_ok ? (int)(a_slots < stages ? a_slots : stages) : 0;
    }
    // CTA pairs (tcgen05.mma.cta_group::2, M = 256): default whenever there are at least two M tiles
    bool pair = m_tiles >= 2 && kp.ws_stages == 0;
    if (env.pair >= 0) pair = pair && env.pair != 0;
    if (pair160) {
      kp.ws_stages = 0;
      // two pairs per cluster sharing their activation rows (see the CL = 4 note at the kernel): linear layers with an
      // even number of N tiles and at least cl4_min 256x160 tiles. Measured SLOWER (profiles/r02_linear_ab_cl4_geglu.txt):
      // opt-in with IVV_CL4=1, IVV_CL4_MIN=<tiles> moves the bar.
      if (is_linear && (n_tiles % 2) == 0 && kp.bw == kBlockM && env.cl4 == 1 &&
          (long long)((m_tiles + 1) / 2) * n_tiles >= (env.cl4_min >= 0 ? env.cl4_min : kCl4MinTiles)) {
        const int max_clusters = pair160_cl4_clusters<5>();
        if (max_clusters >= 8) {
          CUtensorMap tmAh;
          const uint64_t dims[4] = {(uint64_t)a->c, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
          const uint64_t strides[4] = {2, (uint64_t)a->a_ld * 2, (uint64_t)a->a_ld * 2 * a->w,
                                       (uint64_t)a->a_ld * 2 * a->w * a->h};
          const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(kBlockM / 2), 1u, 1u};
          if (int rc = make_tmap_f16(&tmAh, a->a, 4, dims, strides, box, 128)) return rc;
          return launch_pair160_cl4<5>(tmAh, tmB2, tmD, tmR, kp, m_tiles, n_tiles, max_clusters, stream);
        }
      }
      // Weight-stationary mode (K <= 320, linear): built and measured NEUTRAL on the residual GEMMs (28.7 vs 28.0 us) and
      // SLOWER on the 960-wide QKV projection (65.8 vs 53.3 us under ncu: the N-major walk re-reads the 47 MB activation
      // six times with a reuse distance larger than L2 keeps, 129 MB instead of 48 MB of DRAM reads). These kernels are
      // bound by the latency of a ring revolution (~4 000 clk per 5-stage ring under load), not by operand bytes
      // (profiles/r02_gemm_pair160_ncu.txt), so it is opt-in: IVV_WS=1.
      if (a->taps == 1 && kp.kblocks <= kP2WsKBlocks && env.ws == 1)
        return launch_pair160<5, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
      // Activation-stationary mode (see the kernel): K <= 320 with at least three N tiles and enough M pairs for every
      // cluster (the 320 -> 960 QKV projections of the 32x48 level). IVV_AS=0 disables, IVV_AS=2 takes it from two N tiles.
      if (a->taps == 1 && kp.kblocks <= 5 && n_tiles >= (env.as == 2 ? 2 : 3) && (m_tiles + 1) / 2 >= sm_count() / 2 &&
          env.as != 0) {
        g_last_as = 1;
        return launch_pair160<5, false, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
      }
      // K >= 640: one output slab and a sixth pipeline stage (see the kernel). IVV_SLAB1=0 disables.
      // ... and, when K is a multiple of 128, two K blocks per ring slot: half the barrier round trips per tile
      // (18432x640->1920 49.5 -> 46.4 us, 4608x1280->3840 43.5 -> 39.7, profiles/r02_gemm_kb2_ab.txt). IVV_KB2=0 disables.
      if (kp.taps == 1 && kp.kblocks >= 10 && (kp.kblocks % 2) == 0 && (a->c % 128) == 0 && env.slab1 != 0 && env.kb2 != 0)
        return launch_pair160<3, false, false, 1, 2>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
      if (kp.taps * kp.kblocks >= 10 && env.slab1 != 0)
        return launch_pair160<6, false, false, 1>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
      return launch_pair160<5, false>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
    }
    if (ds) {
      kp.ws_stages = 0;
      return launch_persistent_cs<160, 5, 32, false, true, 2, true, false, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
    }
    if (halo) {
      // stage = 20 KB activation box + 3 half weight tiles; ring staging where the main loop (>= 15 iterations of 12
      // MMAs) hides the epilogue anyway
      kp.ws_stages = 0;
      switch (bn_sel) {
        case 256:
          return launch_persistent_cs<256, 3, 32, false, false, 2, true, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
        case 160:
          return launch_persistent_cs<160, 4, 32, false, false, 2, true, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
        default:
          return launch_persistent_cs<128, 4, 64, false, true, 2, true, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
      }
    }
    if (pair && cs == 1) {
#define IVV_PAIR_LAUNCH(BN_, ST_, CW_, GG_, TW_) \
  return launch_persistent_cs<BN_, ST_, CW_, GG_, TW_, 2, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream)
      if (a->geglu) {
        // short main loops (K <= 320: the 32x48 level): the epilogue sets the pace, so spend one pipeline stage on a second
        // output slab -- the store of tile i drains while tile i+1 is written. IVV_GEGLU_DS=0 disables, =2 takes it for
        // every K (tuning hooks).
        // ... and keep the activation rows of an M pair resident while the cluster walks its N tiles (AS, see the kernel):
        // the 320 -> 2560 GEGLU projection of the 32x48 level. IVV_AS=0 disables.
        if (env.geglu_ds != 0 && kp.taps * kp.kblocks <= 5 && n_tiles >= 3 && (m_tiles + 1) / 2 >= sm_count() / 2 &&
            env.as != 0) {
          g_last_as = 1;
          return launch_persistent_cs<256, 5, 32, true, true, 2, true, false, true, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
        }
        if (env.geglu_ds != 0 && (kp.taps * kp.kblocks <= 5 || env.geglu_ds == 2))
          return launch_persistent_cs<256, 5, 32, true, true, 2, true, false, true>(tmA, tmB2, tmD, tmR, kp, m_tiles, n_tiles, stream);
        IVV_PAIR_LAUNCH(256, 6, 32, true, true);
      }
      switch (bn_sel) {
        case 320: IVV_PAIR_LAUNCH(320, 5, 32, false, false);
        case 256:
          if (k_total >= 2560) IVV_PAIR_LAUNCH(256, 6, 64, false, false);
          IVV_PAIR_LAUNCH(256, 5, 64, false, true);
        case 160: IVV_PAIR_LAUNCH(160, 7, 32, false, true);
        case 128: IVV_PAIR_LAUNCH(128, 8, 64, false, true);
        case 64: IVV_PAIR_LAUNCH(64, 8, 32, false, true);
        default: IVV_PAIR_LAUNCH(32, 8, 32, false, true);
      }
#undef IVV_PAIR_LAUNCH
    }
    IVV_REQUIRE(bn_sel != 320, "ivv_gemm: internal: 320-wide tile chosen outside the pair kernel");
    if (a->geglu) return launch_persistent<256, 4, 32, true, true>(tmA, tmB2, tmB, tmD, tmR, kp, m_tiles, n_tiles, cs, stream);
    switch (bn_sel) {
      case 256:
        // long main loops hide the epilogue: spend the shared memory on a 4th pipeline stage instead of a tile-wide slab
        if (k_total >= 2560)
          return launch_persistent<256, 4, 64, false, false>(tmA, tmB2, tmB, tmD, tmR, kp, m_tiles, n_tiles, cs, stream);
        return launch_persistent<256, 3, 64, false, true>(tmA, tmB2, tmB, tmD, tmR, kp, m_tiles, n_tiles, cs, stream);
      case 160: return launch_persistent<160, 5, 32, false, true>(tmA, tmB2, tmB, tmD, tmR, kp, m_tiles, n_tiles, cs, stream);
      case 128: return launch_persistent<128, 6, 64, false, true>(tmA, tmB2, tmB, tmD, tmR, kp, m_tiles, n_tiles, cs, stream);
      case 64: return launch_persistent<64, 6, 32, false, true>(tmA, tmB2, tmB, tmD, tmR, kp, m_tiles, n_tiles, cs, stream);
      default: return launch_persistent<32, 6, 32, false, true>(tmA, tmB2, tmB, tmD, tmR, kp, m_tiles, n_tiles, cs, stream);
    }
  }
  switch (bn_sel) {
    case 256: return launch<256, 4>(tmA, tmB, kp, m_tiles, n_tiles, stream);
    case 160: return launch<160, 3>(tmA, tmB, kp, m_tiles, n_tiles, stream);
    case 128: return launch<128, 3>(tmA, tmB, kp, m_tiles, n_tiles, stream);
    case 64: return launch<64, 4>(tmA, tmB, kp, m_tiles, n_tiles, stream);
    default: return launch<32, 4>(tmA, tmB, kp, m_tiles, n_tiles, stream);
  }
}
