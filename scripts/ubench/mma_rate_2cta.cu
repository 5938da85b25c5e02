// Micro-benchmark: tcgen05.mma cta_group::2 (CTA pair, M=256) kind::tf32 issue rate, SS operands.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../arshamg_scrnaseq_wgan_b200/csrc/gemm_sm100.cuh"
using namespace wg;

__device__ __forceinline__ void umma_tf32_2cta(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
bench(int bn, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 1.0f;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tcgen05_before_sync(); __syncthreads(); cluster_sync_all(); tcgen05_after_sync();
  const uint32_t tmem = tslot;
  long long t0 = 0, t1 = 0;
  if (crank == 0 && threadIdx.x == 0) {
    // M = 256 across the pair, N = bn: A K-major (128 rows per CTA), B K-major (bn/2 rows per CTA)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t sa = base, sb = base + 16384;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int s = 0; s < 4; ++s)
        umma_tf32_2cta(tmem, make_smem_desc(sa + s * 32, 16, 1024, 2), make_smem_desc(sb + s * 32, 16, 1024, 2), idesc, 1u);
    }
    umma_commit_2cta_mc(smem_u32(&bar), 3);
  }
  if (threadIdx.x == 0) {
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (crank == 0) out[blockIdx.x / 2] = t1 - t0;
  }
  tcgen05_before_sync(); __syncthreads(); cluster_sync_all(); tcgen05_after_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int grid : {2, 148})
    for (int bn : {256, 224, 128}) {
      bench<<<grid, 128, 64 * 1024>>>(bn, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[74]; cudaMemcpy(h, d, (grid / 2) * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
      double cyc = (double)mx / (iters * 4);
      printf("tf32 cta_group::2 grid=%3d M=256 N=%3d: %.1f cycles/MMA -> %.0f MAC/clk/SM\n", grid, bn, cyc, 256.0 * bn * 8 / cyc / 2);
    }
  return 0;
}
