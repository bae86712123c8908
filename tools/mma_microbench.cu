// tcgen05.mma instruction-rate microbenchmark (tuning tool, not part of libivv_b200.so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_microbench tools/mma_microbench.cu -lcuda
// One thread per CTA issues back-to-back MMAs on resident shared-memory operands (no TMA, no epilogue) and the CTA
// reports SM clocks per instruction; the host also reports the whole-chip rate from CUDA events. Variants:
//   ss   : A, B K-major SWIZZLE_128B in shared memory (the GEMM main loop)
//   ssmn : B MN-major (the P.V product of the attention kernel)
//   ts   : A read from tensor memory, B K-major in shared memory
//   2sm  : cta_group::2 pair, M = 256 (128 rows of A and N/2 rows of B per CTA)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../insv2v_b200/csrc/common.cuh"

namespace ivv {
void set_error(const char*, ...) {}
}  // namespace ivv
using namespace ivv;

constexpr int kStages = 4;
constexpr int kStageBytes = 48 * 1024;  // 16 KB A (128 x 64 fp16) + 32 KB B (256 x 64 fp16)

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0 ss, 1 ssmn, 2 ts
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int mode, int M, int N, int kblocks, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x38003800u, 0x3c003c00u, 0x34003400u);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(M, N, 0, mode == 1 ? 1 : 0);
    const long long t0 = clock64();
    int stage = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      const uint32_t sa = smem_u32(smem + stage * kStageBytes);
      const uint32_t sb = sa + 16 * 1024;
      const uint64_t adesc = umma_desc_kmajor_sw128(sa);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (mode == 0) {
          umma_f16_ss(tmem_base, adesc + 2 * k, umma_desc_kmajor_sw128(sb) + 2 * k, idesc, 1u);
        } else if (mode == 1) {
          // B: [16 K rows x 128 B] per 64-wide N atom, atoms 8 KB apart (2048 B per K step inside an atom)
          umma_f16_ss(tmem_base, adesc + 2 * k, umma_desc_mnmajor_sw128(sb + k * 2048, 8192), idesc, 1u);
        } else {
          umma_f16_ts(tmem_base, tmem_base + 256 + 8 * k, umma_desc_kmajor_sw128(sb) + 2 * k, idesc, 1u);
        }
      }
      if (++stage == kStages) stage = 0;
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_base);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    mma_bench_2sm_kernel(int N, int kblocks, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int crank = (int)cluster_ctarank();
  for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x38003800u, 0x3c003c00u, 0x34003400u);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc_2sm<512>(&tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    if (crank == 0) {
      const uint32_t idesc = umma_idesc_f16(256, N, 0, 0);
      int stage = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint32_t sb = sa + 16 * 1024;
        const uint64_t adesc = umma_desc_kmajor_sw128(sa), bdesc = umma_desc_kmajor_sw128(sb);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss_2sm(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        if (++stage == kStages) stage = 0;
      }
      umma_commit_2sm(&bar, 3);
    }
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc_2sm<512>(tmem_base);
}

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e = (x);                                                         \
    if (e != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

int main() {
  const int smem = kStages * kStageBytes;
  CK(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(mma_bench_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int kblocks = 2048;
  struct Cfg { const char* name; int mode, M, N; };
  std::vector<Cfg> cfgs;
  for (int n : {16, 32, 48, 64, 96, 128, 160, 192, 224, 256}) cfgs.push_back({"ss", 0, 128, n});
  for (int n : {64, 128, 256}) cfgs.push_back({"ss", 0, 64, n});
  for (int n : {48, 64, 128, 256}) cfgs.push_back({"ssmn", 1, 128, n});
  for (int n : {48, 64, 128, 256}) cfgs.push_back({"ts", 2, 128, n});
  for (int n : {64, 128, 160, 256}) cfgs.push_back({"2sm", 3, 256, n});
  for (int grid : {2, 148}) {
    printf("--- grid %d CTAs\n", grid);
    for (const Cfg& c : cfgs) {
      float best_ms = 1e30f;
      long long cyc = 0;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        if (c.mode == 3) mma_bench_2sm_kernel<<<grid, 128, smem>>>(c.N, kblocks, d_cycles);
        else mma_bench_kernel<<<grid, 128, smem>>>(c.mode, c.M, c.N, kblocks, d_cycles);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
        CK(cudaMemcpy(&cyc, d_cycles, 8, cudaMemcpyDeviceToHost));
      }
      const double n_mma = (double)kblocks * 4;
      const int issuers = c.mode == 3 ? grid / 2 : grid;
      const double flops = 2.0 * c.M * c.N * 16 * n_mma * issuers;
      printf("%-5s M=%3d N=%3d: %7.1f clk/MMA (nominal %5.1f)  chip %8.1f TFLOP/s  (%.3f ms)\n", c.name, c.M, c.N,
             cyc / n_mma, (c.mode == 3 ? 128 : c.M) * c.N * 16 / 4096.0, flops / (best_ms * 1e-3) / 1e12,
             best_ms);
    }
  }
  return 0;
}
