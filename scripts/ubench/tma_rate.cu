// Micro-benchmark: per-SM TMA load ingest rate (no MMA): one CTA per SM streams boxes of a row-major
// fp32 matrix through an 8-slot smem ring; reports bytes/clk/SM and aggregate TB/s.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../arshamg_scrnaseq_wgan_b200/csrc/gemm_sm100.cuh"
using namespace wg;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int boxes_per_stage, int rows_per_box,
              int kblocks, int reuse_rows, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[8];
  const int STAGES = 200 * 1024 / (box_bytes * boxes_per_stage) > 8 ? 8 : 200 * 1024 / (box_bytes * boxes_per_stage);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(smem_u32(&full[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int row0 = (reuse_rows ? (blockIdx.x % reuse_rows) : blockIdx.x) * rows_per_box * boxes_per_stage;
    long long t0 = clock64();
    // prologue: fill the ring
    int issued = 0, waited = 0;
    uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (waited < kblocks) {
      while (issued < kblocks && issued - waited < STAGES) {
        const int s = issued % STAGES;
        mbar_expect_tx(smem_u32(&full[s]), box_bytes * boxes_per_stage);
        for (int b = 0; b < boxes_per_stage; ++b)
          tma_load_2d(base + (s * boxes_per_stage + b) * box_bytes, &map, smem_u32(&full[s]), issued * 32,
                      row0 + b * rows_per_box);
        ++issued;
      }
      const int s = waited % STAGES;
      mbar_wait(smem_u32(&full[s]), phase[s]);
      phase[s] ^= 1u;
      ++waited;
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int SMS = 148;
  long long* d; cudaMalloc(&d, SMS * 8);
  const size_t K = 8192;
  float* A; cudaMalloc(&A, (size_t)SMS * 256 * K * 4 + 4096);
  cudaMemset(A, 0, (size_t)SMS * 256 * K * 4);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  struct Cfg { const char* name; int rows_per_box; int boxes; CUtensorMapSwizzle sw; CUtensorMapL2promotion pr; int reuse; };
  Cfg cfgs[] = {
      {"box[128][128B] x1 sw128 HBM", 128, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 0},
      {"box[128][128B] x2 sw128 HBM", 128, 2, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 0},
      {"box[256][128B] x1 sw128 HBM", 256, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 0},
      {"box[256][128B] x1 sw128 pr256 HBM", 256, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 0},
      {"box[256][128B] x1 sw128 L2-resident(4 row groups)", 256, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 4},
      {"box[128][128B] x2 sw128 L2-resident(4 row groups)", 128, 2, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 4},
      {"box[32][128B] x8 atom32B L2-resident", 32, 8, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 4},
      {"box[256][128B] x1 noswizzle L2-resident", 256, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 4},
  };
  for (auto& c : cfgs) {
    CUtensorMap map;
    cuuint64_t dims[2] = {K, (cuuint64_t)SMS * 256};
    cuuint64_t strides[1] = {K * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)c.rows_per_box};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        c.sw, c.pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int kblocks = K / 32;
    const int box_bytes = c.rows_per_box * 128;
    for (int rep = 0; rep < 2; ++rep) {
      stream_kernel<<<SMS, 64, 220 * 1024>>>(map, box_bytes, c.boxes, c.rows_per_box, kblocks, c.reuse, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[148]; cudaMemcpy(h, d, SMS * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < SMS; ++i) mx = h[i] > mx ? h[i] : mx;
    double bytes = (double)kblocks * box_bytes * c.boxes;
    printf("%-55s %6.1f B/clk/SM  (%.2f TB/s aggregate at 1.9 GHz)\n", c.name, bytes / mx, bytes / mx * SMS * 1.9e9 / 1e12);
  }
  return 0;
}
