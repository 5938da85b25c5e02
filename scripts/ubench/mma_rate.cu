// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 / kind::f16 (SS mode, cta_group::1, M=128)
// for K-major vs MN-major operands. One CTA per SM, no loads, smem contents arbitrary.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../arshamg_scrnaseq_wgan_b200/csrc/gemm_sm100.cuh"
using namespace wg;

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(int a_mn, int b_mn, int bn, int iters, int f16, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 1.0f;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tcgen05_before_sync(); __syncthreads(); tcgen05_after_sync();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t fmt = f16 ? 1u : 2u;   // bf16 : tf32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = base, sb = base + 16384;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint64_t ad = a_mn ? make_smem_desc(sa + s * 1024, 4096, f16 ? 1024 : 512, f16 ? 2 : 1) : make_smem_desc(sa + s * 32, 16, 1024, 2);
        const uint64_t bd = b_mn ? make_smem_desc(sb + s * 1024, 4096, f16 ? 1024 : 512, f16 ? 2 : 1) : make_smem_desc(sb + s * 32, 16, 1024, 2);
        if (f16) umma_f16(tmem, ad, bd, idesc, 1u); else umma_tf32(tmem, ad, bd, idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tcgen05_before_sync(); __syncthreads(); tcgen05_after_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int f16 = 0; f16 < 2; ++f16)
    for (int grid : {1, 148})
      for (int bn : {256, 128})
        for (int cfg = 0; cfg < 3; ++cfg) {
          int a_mn = cfg == 2, b_mn = cfg >= 1;
          bench<<<grid, 128, 64 * 1024>>>(a_mn, b_mn, bn, iters, f16, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          double cyc = (double)mx / (iters * 4);
          int kk = f16 ? 16 : 8;
          printf("%s grid=%3d N=%3d A_%s B_%s: %.1f cycles/MMA -> %.0f MAC/clk/SM\n", f16 ? "bf16" : "tf32", grid, bn,
                 a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", cyc, 128.0 * bn * kk / cyc);
        }
  return 0;
}
