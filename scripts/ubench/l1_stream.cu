// Micro-benchmark: the LOAD pattern of the critic layer-1 forward product without the MMAs.  148 CTAs; CTA c streams
// a contiguous range of (128-row tile, K block) units of A [12288 x 6608] fp32 (HBM stream, 16 KB boxes [128][128 B])
// and, per K block, 16 KB of the L2-resident B [6605 x 200] (3-D MN-major map, 4 chunks), like one CTA of a pair.
// Variants: coupled ring (A and B share a stage, as the GEMM does) vs decoupled rings (deep A ring, shallow B ring),
// B on/off, grouped issue ("burst"), a consumer delay, the weight-gradient pattern, and a CTA-PAIR mode (cluster of 2,
// cta_group::2 loads completing on the leader's barrier).  Reports aggregate TB/s of A and the per-CTA cycle spread.
// Any argument: re-read the same buffer every launch (as scripts/profile_gemm.py does) instead of alternating two.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l1_stream l1_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <algorithm>
#include "../../arshamg_scrnaseq_wgan_b200/csrc/gemm_sm100.cuh"
using namespace wg;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct P {
  CUtensorMap a, b;
  int total_kb, row_tiles;
  int a_stages, b_stages;     // coupled: a_stages stages of (A + B); decoupled: separate rings
  int coupled, use_b;
  const float* lin;           // != nullptr: A is read as contiguous 16 KB pieces (cp.async.bulk), CTA-contiguous ranges
  int lin_interleave;         // 1: piece i goes to CTA i % grid (fine interleave) instead of contiguous ranges
  int wgrad;                  // 1: the layer-1 WEIGHT-GRADIENT pattern: A = x^T (MN-major 3-D map, box (32, 32 k, 4 chunks)), K = 12288 rows
  int burst;                  // A producer: wait for `burst` free slots, then issue that many K blocks back to back
  int consume_cycles;         // busy-wait per K block in the consumer (models the MMA time), 0 = none
  long long* out;             // [grid][3]: cycles, smid, units
};

__global__ void __launch_bounds__(96, 1) l1_stream_kernel(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[64];     // fullA[16] emptyA[16] fullB[16] emptyB[16]
  const uint32_t fullA = smem_u32(&bars[0]), emptyA = smem_u32(&bars[16]), fullB = smem_u32(&bars[32]), emptyB = smem_u32(&bars[48]);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) { mbar_init(fullA + 8 * s, 1); mbar_init(emptyA + 8 * s, 1); mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long U = (long long)p.row_tiles * p.total_kb;
  const long long u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
  const uint32_t a_base = base, b_base = base + (p.coupled ? 16384u : (uint32_t)p.a_stages * 16384u);
  const uint32_t stage_stride = (p.coupled && p.use_b) ? 32768u : 16384u;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 0 && threadIdx.x == 0) {            // A producer (and B when coupled)
    int s = 0; uint32_t ph = 0;
    for (long long u = u0; u < u1;) {
      const int nb = (int)((u1 - u < p.burst) ? (u1 - u) : p.burst);
      { int s2 = s; uint32_t ph2 = ph;
        for (int j = 0; j < nb; ++j) { mbar_wait(emptyA + 8 * s2, ph2 ^ 1u); if (++s2 == p.a_stages) { s2 = 0; ph2 ^= 1u; } } }
      for (int j = 0; j < nb; ++j, ++u) {
        const int tile = (int)(u / p.total_kb), kb = (int)(u % p.total_kb);
        mbar_expect_tx(fullA + 8 * s, (p.coupled && p.use_b) ? 32768 : 16384);
        if (p.lin != nullptr) {
          long long piece = p.lin_interleave ? ((u - u0) * gridDim.x + blockIdx.x) : u; if (piece >= 19824) piece = 19823;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(a_base + s * stage_stride), "l"(p.lin + piece * 4096), "r"(16384u), "r"(fullA + 8 * s) : "memory");
        } else if (p.wgrad) {
          tma_load_3d(a_base + s * stage_stride, &p.a, fullA + 8 * s, 0, kb * 32, tile * 4);
        } else
        tma_load_2d(a_base + s * stage_stride, &p.a, fullA + 8 * s, kb * 32, tile * 128);
        if (p.coupled && p.use_b) tma_load_3d(b_base + s * stage_stride, &p.b, fullA + 8 * s, 0, kb * 32, (blockIdx.x & 1) * 4);
        if (++s == p.a_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && threadIdx.x == 32 && !p.coupled && p.use_b) {     // B producer (decoupled)
    int s = 0; uint32_t ph = 0;
    for (long long u = u0; u < u1; ++u) {
      const int kb = (int)(u % p.total_kb);
      mbar_wait(emptyB + 8 * s, ph ^ 1u);
      mbar_expect_tx(fullB + 8 * s, 16384);
      tma_load_3d(b_base + s * 16384u, &p.b, fullB + 8 * s, 0, kb * 32, (blockIdx.x & 1) * 4);
      if (++s == p.b_stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 2 && threadIdx.x == 64) {    // consumer
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    for (long long u = u0; u < u1; ++u) {
      mbar_wait(fullA + 8 * sa, pa);
      if (!p.coupled && p.use_b) mbar_wait(fullB + 8 * sb, pb);
      if (p.consume_cycles) { long long t = clock64(); while (clock64() - t < p.consume_cycles) { } }
      mbar_arrive(emptyA + 8 * sa);
      if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
      if (!p.coupled && p.use_b) { mbar_arrive(emptyB + 8 * sb); if (++sb == p.b_stages) { sb = 0; pb ^= 1u; } }
    }
    uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.out[blockIdx.x * 3] = clock64() - t0; p.out[blockIdx.x * 3 + 1] = smid; p.out[blockIdx.x * 3 + 2] = u1 - u0;
  }
}


// Pair mode: the CTA-pair structure of gemm_tf32_2cta_kernel without the MMAs.  Cluster (2,1,1); pair q streams a
// contiguous range of (256-row tile, K block) units; CTA rank r loads its 128 rows of A (+ its half of the B tile)
// with cp.async.bulk.tensor...cta_group::2 completing on the LEADER's full barrier; the leader's consumer waits for
// both CTAs' bytes, optionally spins `consume_cycles`, and frees the slot in both CTAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(96, 1) l1_stream_pair_kernel(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[32];     // full[16] empty[16]
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[16]);
  const uint32_t crank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();
  const int npairs = gridDim.x / 2, pair = blockIdx.x / 2;
  const int tiles = p.row_tiles / 2;
  const long long U = (long long)tiles * p.total_kb;
  const long long u0 = U * pair / npairs, u1 = U * (pair + 1) / npairs;
  const uint32_t stage_bytes = p.use_b ? 32768u : 16384u;
  const uint32_t full_leader = mapa_cluster(full, 0);
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 0 && threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (long long u = u0; u < u1;) {
      const int nb = (int)((u1 - u < p.burst) ? (u1 - u) : p.burst);
      { int s2 = s; uint32_t ph2 = ph;
        for (int j = 0; j < nb; ++j) { mbar_wait(empty + 8 * s2, ph2 ^ 1u); if (++s2 == p.a_stages) { s2 = 0; ph2 ^= 1u; } } }
      for (int j = 0; j < nb; ++j, ++u) {
        const int tile = (int)(u / p.total_kb), kb = (int)(u % p.total_kb);
        if (crank == 0) mbar_expect_tx(full + 8 * s, 2 * stage_bytes);
        if (p.wgrad) tma_load_3d_2sm(base + s * stage_bytes, &p.a, full_leader + 8 * s, 0, kb * 32, tile * 8 + crank * 4);
        else         tma_load_2d_2sm(base + s * stage_bytes, &p.a, full_leader + 8 * s, kb * 32, tile * 256 + crank * 128);
        if (p.use_b) tma_load_3d_2sm(base + s * stage_bytes + 16384u, &p.b, full_leader + 8 * s, 0, kb * 32, crank * 4);
        if (++s == p.a_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 2 && threadIdx.x == 64 && crank == 0) {
    int s = 0; uint32_t ph = 0;
    for (long long u = u0; u < u1; ++u) {
      mbar_wait(full + 8 * s, ph);
      if (p.consume_cycles) { long long t = clock64(); while (clock64() - t < p.consume_cycles) { } }
      mbar_arrive(empty + 8 * s);
      mbar_arrive_cluster(empty + 8 * s, 1);
      if (++s == p.a_stages) { s = 0; ph ^= 1u; }
    }
    p.out[blockIdx.x * 3] = clock64() - t0; p.out[blockIdx.x * 3 + 1] = 0; p.out[blockIdx.x * 3 + 2] = u1 - u0;
    p.out[(blockIdx.x + 1) * 3] = clock64() - t0;
  }
  __syncthreads();
  cluster_sync_all();
}

int main(int argc, char** argv) {
  const bool same_buffer = argc > 1;      // any argument: re-read the SAME buffer every launch (like scripts/profile_gemm.py)
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int M = 12288, K = 6605, LD = 6608, N = 200;
  float *A, *B, *flush; long long* d;
  cudaMalloc(&A, (size_t)M * LD * 4 + 4096); cudaMemset(A, 0, (size_t)M * LD * 4);
  float* A2; cudaMalloc(&A2, (size_t)M * LD * 4 + 4096); cudaMemset(A2, 0, (size_t)M * LD * 4);
  cudaMalloc(&B, (size_t)K * N * 4 + 4096); cudaMemset(B, 0, (size_t)K * N * 4);
  cudaMalloc(&flush, 256 << 20); cudaMalloc(&d, 148 * 3 * 8);
  P p; memset(&p, 0, sizeof p);
  { cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t str[1] = {(cuuint64_t)LD * 4}; cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&p.a, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A %d\n", (int)r); return 1; } }
  { cuuint64_t dims[3] = {32, (cuuint64_t)K, (cuuint64_t)((N + 31) / 32)}; cuuint64_t str[2] = {(cuuint64_t)N * 4, 128}; cuuint32_t box[3] = {32, 32, 4}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.b, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, B, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B %d\n", (int)r); return 1; } }
  p.total_kb = (K + 31) / 32; p.row_tiles = M / 128; p.out = d;
  cudaFuncSetAttribute(l1_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  struct V { const char* name; int coupled, use_b, a_st, b_st, consume, burst; int promo; int f32; };
  V vs[] = {
    {"A tiles [128][128 B], 6 x 16 KB, one box per freed slot", 1, 0, 6, 0, 0, 1, 2, 0},
    {"A tiles, 12 x 16 KB, burst 1", 1, 0, 12, 0, 0, 1, 2, 0},
    {"A tiles, 12 x 16 KB, burst 2", 1, 0, 12, 0, 0, 2, 2, 0},
    {"A tiles, 12 x 16 KB, burst 3", 1, 0, 12, 0, 0, 3, 2, 0},
    {"A tiles, 12 x 16 KB, burst 4", 1, 0, 12, 0, 0, 4, 2, 0},
    {"A tiles, 12 x 16 KB, burst 4, L2 promotion none", 1, 0, 12, 0, 0, 4, 0, 0},
    {"A linear 16 KB pieces (cp.async.bulk), CTA-contiguous", 1, 0, 12, 0, 0, 1, 2, 2},
    {"coupled A+B, 6 x 32 KB burst 1 (the GEMM's ring)", 1, 1, 6, 0, 0, 1, 2, 0},
    {"coupled A+B, 6 x 32 KB burst 2", 1, 1, 6, 0, 0, 2, 2, 0},
    {"coupled A+B, 6 x 32 KB burst 1, consumer 512 clk", 1, 1, 6, 0, 512, 1, 2, 0},
    {"coupled A+B, 6 x 32 KB burst 2, consumer 512 clk", 1, 1, 6, 0, 512, 2, 2, 0},
    {"coupled A+B, 6 x 32 KB burst 3, consumer 512 clk", 1, 1, 6, 0, 512, 3, 2, 0},
    {"decoupled A 10 x 16 KB burst 4, B 3 x 16 KB", 0, 1, 10, 3, 0, 4, 2, 0},
    {"decoupled A 10 burst 4 / B 3, consumer 512 clk", 0, 1, 10, 3, 512, 4, 2, 0},
  };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (auto& v : vs) {
    p.coupled = v.coupled; p.use_b = v.use_b; p.burst = v.burst; p.lin = v.f32 >= 2 ? A : nullptr; p.lin_interleave = v.f32 == 3;
    { cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t str[1] = {(cuuint64_t)LD * 4}; cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
      CUtensorMapL2promotion pr[3] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
      encode(&p.a, v.f32 == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr[v.promo], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); } p.a_stages = v.a_st; p.b_stages = v.b_st; p.consume_cycles = v.consume;
    P p2 = p; p2.lin = p.lin ? A2 : nullptr;
    { cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t str[1] = {(cuuint64_t)LD * 4}; cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
      encode(&p2.a, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, A2, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      if (!same_buffer) l1_stream_kernel<<<148, 96, 225 * 1024>>>(p2);      // reads the other A buffer: evicts this one with CLEAN lines
      cudaEventRecord(e0);
      l1_stream_kernel<<<148, 96, 225 * 1024>>>(p);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
    }
    long long h[148 * 3]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    long long mn = 1LL << 60, mx = 0; double avg = 0;
    for (int i = 0; i < 148; ++i) { mn = std::min(mn, h[3 * i]); mx = std::max(mx, h[3 * i]); avg += h[3 * i] / 148.0; }
    printf("%-44s %7.1f us  A %.2f TB/s  | per-CTA cycles min %lld avg %.0f max %lld (per K block %.0f / %.0f / %.0f)\n", v.name, best * 1e3,
           (double)M * LD * 4 / best / 1e9, mn, avg, mx, mn / 134.3, avg / 134.3, mx / 134.3);
    if (false) {
      // per-SM-id cycles, sorted by smid: neighbouring ids share a TPC / GPC
      std::pair<long long, long long> s[148];
      for (int i = 0; i < 148; ++i) s[i] = {h[3 * i + 1], h[3 * i]};
      std::sort(s, s + 148);
      printf("   smid:kcycles ");
      for (int i = 0; i < 148; ++i) printf("%lld:%lld ", s[i].first, s[i].second / 1000);
      printf("\n");
    }
  }
  // ---- weight-gradient pattern: A = x^T [6605 m x 12288 k] (m contiguous), B = delta [12288 k x 200] ----
  {
    P w; memset(&w, 0, sizeof w);
    cuuint64_t dims[3] = {32, (cuuint64_t)M, (cuuint64_t)((K + 31) / 32)}; cuuint64_t str[2] = {(cuuint64_t)LD * 4, 128}; cuuint32_t box[3] = {32, 32, 4}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&w.a, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode wA %d\n", (int)r); return 1; }
    P w2 = w;
    encode(&w2.a, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, A2, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    float* D1; cudaMalloc(&D1, (size_t)M * N * 4 + 4096); cudaMemset(D1, 0, (size_t)M * N * 4);
    cuuint64_t bd[3] = {32, (cuuint64_t)M, (cuuint64_t)((N + 31) / 32)}; cuuint64_t bs[2] = {(cuuint64_t)N * 4, 128};
    encode(&w.b, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, D1, bd, bs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    w2.b = w.b;
    struct W { const char* name; int use_b, a_st, consume, burst; };
    W ws[] = {{"wgrad A only 12 x 16 KB burst 1", 0, 12, 0, 1}, {"wgrad A only 12 x 16 KB burst 2", 0, 12, 0, 2}, {"wgrad A only 12 x 16 KB burst 4", 0, 12, 0, 4},
              {"wgrad coupled A+B 6 x 32 KB burst 1", 1, 6, 0, 1}, {"wgrad coupled A+B 6 x 32 KB burst 2", 1, 6, 0, 2},
              {"wgrad coupled burst 1, consumer 512", 1, 6, 512, 1}, {"wgrad coupled burst 2, consumer 512", 1, 6, 512, 2}, {"wgrad coupled burst 3, consumer 512", 1, 6, 512, 3}};
    for (auto& v : ws) {
      for (P* q : {&w, &w2}) { q->wgrad = 1; q->coupled = 1; q->use_b = v.use_b; q->a_stages = v.a_st; q->burst = v.burst; q->consume_cycles = v.consume;
                               q->total_kb = M / 32; q->row_tiles = (K + 127) / 128; q->out = d; }
      float best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        if (!same_buffer) l1_stream_kernel<<<148, 96, 225 * 1024>>>(w2);
        cudaEventRecord(e0);
        l1_stream_kernel<<<148, 96, 225 * 1024>>>(w);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
      }
      long long h[148 * 3]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      long long mn = 1LL << 60, mx = 0; double avg = 0;
      for (int i = 0; i < 148; ++i) { mn = std::min(mn, h[3 * i]); mx = std::max(mx, h[3 * i]); avg += h[3 * i] / 148.0; }
      printf("%-44s %7.1f us  A %.2f TB/s  | per-CTA cycles min %lld avg %.0f max %lld\n", v.name, best * 1e3, (double)M * LD * 4 / best / 1e9, mn, avg, mx);
    }
  }
  // ---- pair mode (forward pattern and weight-gradient pattern) ----
  {
    cudaFuncSetAttribute(l1_stream_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    struct Q { const char* name; int wgrad, use_b, st, consume, burst; };
    Q qs[] = {{"PAIR fwd A only 12 x 16 KB burst 1", 0, 0, 12, 0, 1}, {"PAIR fwd A only 12 x 16 KB burst 2", 0, 0, 12, 0, 2}, {"PAIR fwd A only 12 x 16 KB burst 3", 0, 0, 12, 0, 3},
              {"PAIR fwd A+B 6 x 32 KB burst 1, consumer 512", 0, 1, 6, 512, 1}, {"PAIR fwd A+B 6 x 32 KB burst 2, consumer 512", 0, 1, 6, 512, 2}, {"PAIR fwd A+B 6 x 32 KB burst 3, consumer 512", 0, 1, 6, 512, 3},
              {"PAIR fwd A+B 6 x 32 KB burst 1, consumer 128", 0, 1, 6, 128, 1}, {"PAIR fwd A+B 6 x 32 KB burst 2, consumer 128", 0, 1, 6, 128, 2},
              {"PAIR wgrad A+B 6 x 32 KB burst 1, consumer 512", 1, 1, 6, 512, 1}, {"PAIR wgrad A+B 6 x 32 KB burst 2, consumer 512", 1, 1, 6, 512, 2}};
    for (auto& v : qs) {
      P q; memset(&q, 0, sizeof q);
      if (v.wgrad) {
        cuuint64_t dims[3] = {32, (cuuint64_t)M, (cuuint64_t)((K + 31) / 32)}; cuuint64_t str[2] = {(cuuint64_t)LD * 4, 128}; cuuint32_t box[3] = {32, 32, 4}; cuuint32_t es[3] = {1, 1, 1};
        encode(&q.a, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        q.total_kb = M / 32; q.row_tiles = 2 * ((K + 255) / 256);
      } else {
        q.a = p.a; q.total_kb = (K + 31) / 32; q.row_tiles = M / 128;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t str[1] = {(cuuint64_t)LD * 4}; cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
        encode(&q.a, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      q.b = p.b; q.wgrad = v.wgrad; q.use_b = v.use_b; q.a_stages = v.st; q.burst = v.burst; q.consume_cycles = v.consume; q.out = d;
      float best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        l1_stream_pair_kernel<<<148, 96, 225 * 1024>>>(q);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
      }
      printf("%-52s %7.1f us  A %.2f TB/s\n", v.name, best * 1e3, (double)M * LD * 4 / best / 1e9);
    }
  }
  return 0;
}
