// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers.
// Everything here is inline PTX written for compute_100a; nothing is borrowed from CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ivv {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define IVV_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::ivv::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)
#define IVV_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::ivv::set_error(__VA_ARGS__);      \
      return 1;                           \
    }                                     \
  } while (0)

// host: encode a tiled TMA descriptor for an fp16 tensor (rank <= 5). dims/strides innermost-first,
// strides in BYTES for dims 1..rank-1 (dim 0 is contiguous). swizzle_bytes: 0 (none), 32, 64 or 128.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// host: "first launch of this kernel on the CURRENT device" latch. cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a
// per-device attribute, so a process-wide `static bool` would leave a second GPU of the same process unconfigured.
// Racy by design (two threads may both see "first": the attribute is then set twice, which is harmless).
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// host: programmatic dependent launch (PDL). Every kernel launched through launch_pdl() executes griddep_sync() before
// its first global-memory access, so kernel N+1's launch latency and prologue (barrier init, TMEM alloc, descriptor
// prefetch) overlap kernel N's tail; the chain stays ordered because each kernel only completes after its own wait.
// IVV_PDL=0 turns the launch attribute off (plain stream order) - the device-side wait is then a no-op.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// PDL, device side: block until the preceding kernel in the stream has completed and its writes are visible, then let
// the following kernel start launching. Must run before the first read or write of global memory.
__device__ __forceinline__ void griddep_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Entry into a single-thread role (TMA producer, MMA issuer) from a converged warp. elect.sync lets ptxas prove that one
// thread runs the branch, so descriptors and barrier addresses live in uniform registers and tcgen05.mma / TMA issue
// back to back; with `lane == 0` each of them sits in an ELECT + R2UR + BRA.U.ANY loop. -DIVV_ROLE_ELECT=0 restores the
// old form for A/B timing (tools/gpu_ab_lib.sh).
#ifndef IVV_ROLE_ELECT
#define IVV_ROLE_ELECT 1
#endif
__device__ __forceinline__ bool role_elect() {
#if IVV_ROLE_ELECT
  return elect_one();
#else
  return (threadIdx.x & 31) == 0;
#endif
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef IVV_MBAR_TEST_WAIT
#define IVV_MBAR_TEST_WAIT 0  // 1: spin on the non-blocking mbarrier.test_wait instead of try_wait (A/B of wake-up latency)
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if IVV_MBAR_TEST_WAIT
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
#endif
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug becomes a trapped launch (cudaErrorLaunchFailure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// ---- proxies / fences ----
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMA (tile mode) ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- clusters: rank / barrier / multicast TMA ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one L2 read, delivered to the same smem offset (and mbarrier) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_3d_multicast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                      int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], "
      "[%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}

// ---- CTA pairs (cta_group::2): the even CTA of a 2-CTA cluster leads ----
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// NB: no .release.cluster qualifier — ptxas turns that into MEMBAR + ERRBAR, which drains the thread's outstanding TMA
// traffic and serialises a producer loop (measured: 2.2x slower pair-mode GEMM)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA prefetch of a 4-D box into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// TMA loads of a CTA pair: data lands in the issuing CTA's shared memory, the bytes are counted on `bar_cluster_addr`
// (the leader's barrier)
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// CTA pairs inside a larger cluster (two pairs sharing an operand): ONE L2 read delivered to the same shared-memory
// offset of every CTA in cta_mask. The bytes are counted on the barrier at `bar`'s offset in the LEADER (even CTA) of each
// destination CTA's pair: the peer bit (bit 24) of the issuing CTA's own shared-window address is cleared, as the
// 2-SM multicast TMA atoms of CUTLASS do. Every mask used here holds CTAs of ONE parity (the issuer's).
__device__ __forceinline__ void tma_load_4d_2sm_multicast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                          int c2, int c3, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "h"(cta_mask)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs, M = 256] (+)= A[128 rows per CTA] * B[N/2 rows per CTA]; issued by the leader CTA only
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- TMA store (smem -> global, bulk async group) ----
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMEM allocation ----
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// ---- UMMA descriptors ----
// K-major operand, 128-byte swizzle, rows of 64 fp16 (=128 B); 8-row core-matrix groups are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);       // start address
  d |= static_cast<uint64_t>(1) << 16;                          // LBO (ignored for swizzled K-major; CUTLASS sets 1)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // SBO
  d |= static_cast<uint64_t>(1) << 46;                          // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                          // SWIZZLE_128B
  return d;
}
// MN-major operand (rows = K index, 64 contiguous MN elements per 128-B row), 128-byte swizzle.
// lbo_bytes = distance between consecutive 64-wide MN atoms; 8 K-rows are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// instruction descriptor: fp16 (fmt 0) or bf16 (fmt 1) inputs, fp32 accumulate, A K-major, B K-major or MN-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N, uint32_t fmt, uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// same, but the arrive is delivered to the barrier at this smem offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- TMEM loads: 32 lanes x 32 consecutive columns -> 32 registers per thread (thread = lane/row) ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// one column: thread = lane/row gets the single 32-bit value at column taddr
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Walk over a (row, vector) grid of a gather / scatter / copy -- nf rows (frames, pixels) of J 16-byte vectors -- without per-vector
// index arithmetic: a thread keeps its vector column j and strides over rows (R = blockDim / J rows per pass; when a
// row is longer than the CTA, several columns per thread), so the inner loop is one copy plus two pointer increments.
// fn(f, j, g, sm): g / sm = offsets in halfs into the global tensor (frame stride gs) and the shared tile (row stride ss).
// The straightforward i / J, i % J form with 64-bit row products spent ~75 instructions per vector and made the kernel
// temporal-attention and layout kernels issue-bound (1.1 - 2.4 TB/s).
template <class Fn>
__device__ __forceinline__ void walk_rows(int J, int nf, long long gs, int ss, Fn&& fn) {
  const int nt = blockDim.x, tid = threadIdx.x;
  int R = 1, f0 = 0, j0 = tid, jstep = nt;
  if (nt >= J) {
    R = nt / J;
    f0 = tid / J;
    j0 = tid - f0 * J;
    jstep = J;  // one column per thread
    if (f0 >= R) return;
  }
  const long long dg = R * gs;
  const int dsm = R * ss;
  for (int j = j0; j < J; j += jstep) {
    long long g = f0 * gs + j * 8;
    int sm = f0 * ss + j * 8;
#pragma unroll 4
    for (int f = f0; f < nf; f += R, g += dg, sm += dsm) fn(f, j, g, sm);
  }
}

// a += lo(packed), b += hi(packed): fp32 + fp16 in ONE instruction each (PTX add.f32.f16 -> SASS FHADD with an .H1
// operand selector), instead of a conversion (HADD2.F32) followed by FADD. Halves the FP instruction count of the GEMM
// epilogues' bias / residual adds.
#ifndef IVV_FHADD
#define IVV_FHADD 1  // -DIVV_FHADD=0: conversion + FADD instead (A/B timing of the FHADD rate, tools/build_variant.sh)
#endif
__device__ __forceinline__ void add_h2(float& a, float& b, uint32_t packed) {
#if !IVV_FHADD
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&packed));
  a += f.x;
  b += f.y;
  return;
#endif
  asm("{\n\t"
      ".reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %2;\n\t"
      "add.rn.f32.f16 %0, lo, %0;\n\t"
      "add.rn.f32.f16 %1, hi, %1;\n\t"
      "}"
      : "+f"(a), "+f"(b)
      : "r"(packed));
}
__device__ __forceinline__ void add_h8(float* v, const uint4& u) {
  add_h2(v[0], v[1], u.x);
  add_h2(v[2], v[3], u.y);
  add_h2(v[4], v[5], u.z);
  add_h2(v[6], v[7], u.w);
}

// ---- small math ----
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(x); explicit .ftz MUFU forms (no denormal-range fix-up code around the rcp / ex2)
__device__ __forceinline__ float silu_f(float x) { return x * rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
// erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below the fp16 output rounding): 1 rcp + 1 ex2 + 7 FMA
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_ftz(fmaf(0.3275911f, ax, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_ftz(ax * ax * -1.4426950408889634f);
  const float r = 1.f - poly * t * e;
  return copysignf(r, x);
}
// x * Phi(x) with the same erf, arranged for the GEGLU epilogue (the hot instruction stream of the K = 320 GEGLU GEMM,
// which is issue-bound in its epilogue): 0.5 * (x + |x| * erf(|x| / sqrt 2)), 2 MUFU + 12 FP instructions, explicit
// .ftz forms (the plain __fdividef / __expf carry denormal-range fix-ups: 8 more instructions per element).
__device__ __forceinline__ float gelu_fast_f(float x) {
  const float ax = fabsf(x);
  const float t = rcp_ftz(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_ftz((x * -0.72134752044448170f) * x);  // exp(-x^2 / 2)
  const float erf_abs = fmaf(-(poly * t), e, 1.f);
  return 0.5f * fmaf(ax, erf_abs, x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ivv
