// TMA request-rate microbenchmark (tuning tool, not part of libivv_b200.so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tma_microbench tools/tma_microbench.cu -lcuda
// Question it answers: is the ~41 B/clk per SM the GEMM main loops see a BYTE limit (the chip-wide L2 cap of
// ~6300 B/clk shared by 148 SMs) or a per-SM REQUEST limit of the TMA unit (box rows per clock)? One thread per CTA
// keeps DEPTH 2-D tile loads (or stores) in flight on an L2-resident fp16 tensor [rows, 640]; boxes of 128 rows x 128 B
// are compared with 128 rows x 64 B (half the bytes, same number of rows) and 256 rows x 64 B (same bytes, twice the
// rows), on one SM and on all of them. If time follows rows, narrow boxes (the 32-column epilogue chunks of the
// 160-wide tiles) cost as much as full-width ones.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>

#include "../insv2v_b200/csrc/common.cuh"

namespace ivv {
void set_error(const char*, ...) {}
}  // namespace ivv
using namespace ivv;

constexpr int kDepth = 6;
constexpr int kSlotBytes = 32 * 1024;
constexpr int kPitchElems = 640;    // row pitch of the tensor (fp16 elements): the C = 640 activation layout
constexpr int kRegionRows = 512;    // rows per CTA region: 512 x 1280 B = 640 KB, x 148 CTAs = 95 MB < L2

__global__ void __launch_bounds__(128, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int box_cols, int box_rows, int iters, int store,
                unsigned long long* out_clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kDepth * kSlotBytes);
  const int box_bytes = box_cols * box_rows * 2;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    for (int s = 0; s < kDepth; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int row_base = blockIdx.x * kRegionRows;
    const int col_blocks = kPitchElems / box_cols;
    const int row_blocks = kRegionRows / box_rows;
    const int nboxes = col_blocks * row_blocks;
    // one untimed pass over the region first (L2 warm), then `iters` timed boxes; one continuous sequence so the
    // barrier phases simply keep alternating
    const int total = nboxes + iters;
    unsigned long long t0 = 0;
    for (int i = 0; i < total; ++i) {
      if (i == nboxes) t0 = clock64();
      const int slot = i % kDepth;
      const int b = i % nboxes;
      const int c0 = (b % col_blocks) * box_cols, r0 = row_base + (b / col_blocks) * box_rows;
      if (!store) {
        if (i >= kDepth) mbar_wait(&full[slot], ((i / kDepth) - 1) & 1);
        mbar_expect_tx(&full[slot], box_bytes);
        tma_load_2d(smem + slot * kSlotBytes, &tm, &full[slot], c0, r0);
      } else {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tm)),
                     "r"(smem_u32(smem + slot * kSlotBytes)), "r"(c0), "r"(r0)
                     : "memory");
        bulk_commit_group();
        bulk_wait_group_read<kDepth - 1>();
      }
    }
    if (!store) {
      for (int i = total - kDepth; i < total; ++i) mbar_wait(&full[i % kDepth], (i / kDepth) & 1);
    } else {
      bulk_wait_group<0>();
    }
    out_clk[blockIdx.x] = clock64() - t0;
  }
}

static CUtensorMap make_map(void* base, long long rows, int box_cols, int box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)kPitchElems, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)kPitchElems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(1);
  }
  return m;
}

int main() {
  cudaFree(0);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const long long rows = (long long)sms * kRegionRows;
  void* buf = nullptr;
  cudaMalloc(&buf, rows * kPitchElems * 2);
  cudaMemset(buf, 0, rows * kPitchElems * 2);
  unsigned long long* d_clk = nullptr;
  cudaMalloc(&d_clk, sms * sizeof(unsigned long long));
  const int smem = kDepth * kSlotBytes + 256;
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  struct Box { int cols, rows; };
  const Box boxes[] = {{64, 128}, {32, 128}, {32, 256}, {64, 64}, {64, 256}, {16, 256}};
  printf("%-6s %-5s %9s %9s | %12s %10s %10s\n", "op", "CTAs", "row bytes", "box rows", "clk per box", "B/clk/SM", "rows/clk");
  for (int store = 0; store < 2; ++store)
    for (int grid : {1, sms})
      for (const Box& b : boxes) {
        if (b.cols * b.rows * 2 > kSlotBytes) continue;
        CUtensorMap tm = make_map(buf, rows, b.cols, b.rows);
        tma_rate_kernel<<<grid, 128, smem>>>(tm, b.cols, b.rows, iters, store, d_clk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("kernel failed: %s\n", cudaGetErrorString(e));
          return 1;
        }
        std::vector<unsigned long long> clk(grid);
        cudaMemcpy(clk.data(), d_clk, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (auto c : clk) mean += (double)c;
        mean /= grid;
        const double per_box = mean / iters;
        printf("%-6s %-5d %9d %9d | %12.1f %10.1f %10.3f\n", store ? "store" : "load", grid, b.cols * 2, b.rows, per_box,
               b.cols * b.rows * 2 / per_box, b.rows / per_box);
      }
  return 0;
}
