// TF32 tensor-core GEMM for sm_100a: TMA-fed, tcgen05.mma with the accumulator in TMEM, persistent
// warp-specialised CTAs (1 TMA warp, 1 MMA warp, 8 epilogue warps), fused epilogue, TMA store.
//
// Two kernels share the structure:
//   gemm_tf32_2cta_kernel  CTA pairs (tcgen05.mma cta_group::2, cluster (2,1,1)): a pair owns a 256 x bn tile,
//                          each CTA holds its 128 rows of A, half of the B tile and its rows' accumulator; the
//                          leader issues the MMAs, every load of the pair completes on the leader's barrier.
//                          Default: full TF32 issue rate (N/2 cycles per UMMA_K step) and 32 KB stages.
//   gemm_tf32_kernel       one CTA per 128 x bn tile (cta_group::1, 43 + N/2 cycles per step); used when 256-row
//                          tiles would pad the row count >10 % more.  Optional B-tile cluster multicast and
//                          256-row (two-accumulator) tiles are kept as measured-and-rejected experiments.
//
// Computes, for every split z of the K range,
//      D[z][m, n] = epi( sum_{k in split z} A(m, k) * B(n, k) )
// where A and B are fp32 matrices in global memory read as TF32 by TMA (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32,
// round to nearest), each either "K-major" (K contiguous) or "MN-major" (M resp. N contiguous).  This covers
// the three dense-layer products of the reference graph with row-major activations [rows, features] and
// tf.layers.dense kernels [in, out]:
//      forward   Y  = X  * W      A = X  (K-major),  B = W  (MN-major)   [W:65,71,77,107,113,119]
//      dgrad     dX = dY * W^T    A = dY (K-major),  B = W  (K-major)    (tf.gradients, [W:154-155])
//      wgrad     dW = X^T * dY    A = X  (MN-major), B = dY (MN-major)
// Shared-memory layouts: K-major tiles are SWIZZLE_128B boxes [rows][32 k]; MN-major 32-bit tiles must be
// SWIZZLE_128B_BASE32B (TMA 128B_ATOM_32B) and come from a 3-D map (32 floats, K, MN/32) as one box
// [chunk][32 k][32 mn].  Ragged M / N / K are handled by TMA zero fill and store clipping.
// Long reductions with few output tiles run as stream-K in the CTA-pair kernel (one contiguous range of
// (tile, K block) units per pair; partial accumulators meet in a workspace, the tile's owner applies the fused
// epilogue).  The epilogue (applied when splits == 1; split-K partials, kept for raw-slab callers and the
// single-CTA kernel, are finished by finish_splitk_* in kernels.cuh with the same EpiParams) is
//      y = acc (+ bias[n]);  y = leaky_relu(y);  y *= slope(aux[m,n]);  y *= row_scale[m]
//      row_sumsq[tile][m] = sum_n y^2;   optional second output  y2 = eps[m] * x[m,n] + (1 - eps[m]) * y
// where slope(h) = h > 0 ? 1 : alpha recovers phi'(a) from the post-activation.  The [M, N] side
// operand (aux or x) is staged through shared memory by TMA.  Every kernel starts with
// griddepcontrol.launch_dependents and waits (griddepcontrol.wait) only after its barrier / TMEM prologue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wg {

#ifdef WGAN_DIAG_STAMPS      // diagnostic build: per-CTA phase time stamps of the pair kernel's last launch
__device__ unsigned long long g_diag_stamps[160][8];
__device__ __forceinline__ unsigned long long diag_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DIAG_STAMP(i) do { if (w.ew == 0 && lane == 0) g_diag_stamps[blockIdx.x][i] = diag_now(); } while (0)
#define g_diag_mode_store(mode, it) do { if (w.ew == 0 && lane == 0) g_diag_stamps[blockIdx.x][7] = (it == 0 ? 0ull : g_diag_stamps[blockIdx.x][7]) | (static_cast<unsigned long long>(mode + 1) << (8 * it)); } while (0)
#else
#define DIAG_STAMP(i) do { } while (0)
#define g_diag_mode_store(mode, it) do { } while (0)
#endif

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 32;           // 32 tf32 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 8;        // 32 B / sizeof(tf32)
// Epilogue shape (A/B-tested on the whole iteration: 4 warps x 2 boxes 361 it/s, 8 x 2 363, 4 x 1 355, 8 x 1 381)
#ifndef WGAN_EPI_WARPS
#define WGAN_EPI_WARPS 8
#endif
#ifndef WGAN_EPI_NBUF
#define WGAN_EPI_NBUF 1
#endif
constexpr int NUM_EPI_WARPS = WGAN_EPI_WARPS;       // 4 or 8 (8: two warps per TMEM lane quadrant alternate chunks)
constexpr int EPI_NBUF = WGAN_EPI_NBUF;             // staging boxes per epilogue warp (output ring and side ring)
constexpr int GEMM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int MAX_STAGES = 8;
constexpr int EPI_BOX_BYTES = 32 * 32 * 4;   // one [32 rows][32 cols] fp32 staging box
constexpr int EPI_BUFS = EPI_NBUF * NUM_EPI_WARPS;
constexpr int PREFETCH_DIST = 12;             // K blocks between the L2 prefetch and the smem load

struct EpiParams {
  const float* bias;        // [N] or nullptr
  const float* aux;         // [M, ld_aux] post-activation whose sign picks the LeakyReLU slope, or nullptr
  const float* row_scale;   // [M] or nullptr
  float* row_sumsq;         // [n_tiles * NUM_EPI_WARPS / 4, M] partial sums of y^2, or nullptr
  int ld_aux;
  int act;                  // 0 = identity, 1 = LeakyReLU(alpha)
  float alpha;
  int ld_blend;
  // second output (tma_d2): y2 = eps[m] * blend_x[m,n] + (1 - eps[m]) * y  -- the WGAN-GP interpolate
  // x_hat written by the generator-output GEMM itself; nullptr = off
  const float* blend_x;
  const float* blend_eps;
};

struct GemmParams {
  CUtensorMap tma_a;        // 64 B aligned by the driver type
  CUtensorMap tma_b;
  CUtensorMap tma_d;        // 3-D (N, M, splits) fp32, box (32, 32, 1), SWIZZLE_128B
  CUtensorMap tma_d2;       // same shape: second output of the blend epilogue (unused otherwise)
  CUtensorMap tma_side;     // 2-D (N, M) fp32, box (32, 32), SWIZZLE_128B: the epilogue's [M, N] side operand
                            // (slope-mask source or blend x) staged through smem by TMA when side_mode != 0
  int M, N, K;
  int bn;                   // N tile, multiple of 32, <= 256
  int splits;
  int kb_per_split;         // k-blocks (of BK) per split
  int m_tiles, n_tiles;
  int stages;
  int total_kb;
  int cluster;              // CTAs per cluster (1, 2 or 4) along M: they share the B tile by TMA multicast
  int m_groups;             // ceil(m_tiles / cluster)
  int prefetch_a;           // 1: TMA-prefetch A tiles into L2 PREFETCH_DIST K blocks ahead
  int prefetch_b;
  int side_mode;            // 0: side operand (if any) read with per-thread global loads; 1: TMA-staged
  int msub;                 // 128-row sub-tiles per CTA tile (1 or 2): 2 halves the B bytes per FLOP
  int nacc;                 // TMEM accumulator stages: 2 if 2 * msub * bn <= 512 columns, else 1
  // stream-K (CTA-pair kernel only): the (tile, K block) units of the whole product are cut into ONE contiguous
  // range per CTA pair, so every pair streams the same number of K blocks whatever the tile count.  A range that
  // starts inside a tile leaves its accumulator in `sk_ws` (+ flag); the pair that holds the tile's first K block
  // adds those partials in a fixed order and runs the fused epilogue: no split-K slabs, no finish pass.
  int streamk;              // 0 = tiles (x splits) round-robin, 1 = stream-K
  long long sk_units;       // m_tiles * n_tiles * total_kb
  float* sk_ws;             // partial accumulators [pair][cta 2][chunk 8][j 8][row 128] float4
  // unit range of pair q = [sk_b[q], sk_b[q + 1]): equal shares, except that a pair whose LAST segment is a partial
  // gets a few units less -- its dump (~4 us) then ends when the tile's owner is ready for it (engine.cu)
  int sk_b[80];
  unsigned int* sk_flags;   // [pair][cta 2][epilogue warp]: 1 = that warp's share of the partial is written
  EpiParams epi;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// Programmatic dependent launch: let the next grid be scheduled early / wait for the previous grid's memory.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// MN-major operands are described by a 3-D map (32 floats, K rows, MN/32 chunks) so that ONE box of
// (32, 32, chunks) fills the whole [chunk][k][32] smem tile: 4 KB boxes cap the TMA engine at ~31 B/clk/SM,
// 32 KB boxes reach ~51 B/clk/SM (scripts/ubench/tma_rate.cu).
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                               int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
// Warp-uniform leader election: the issuing warps keep their control flow (and therefore the TMA /
// UMMA descriptor arithmetic) uniform and predicate only the issue itself on the elected lane, so
// the operands live in uniform registers instead of being funnelled through R2UR per instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                               int c0, int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6}], [%2], %3;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.ld of 32 consecutive fp32 columns of this thread's TMEM lane: issue, then (later) wait.
// The wait lists v[] as in/out operands so the compiler cannot hoist any use of v above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),
                 "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]),
                 "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
// 32 consecutive floats: eight 16 B loads when the whole chunk is in range and aligned,
// guarded scalar loads (value `fill` out of range) otherwise.
__device__ __forceinline__ void load_chunk32(const float* __restrict__ base, int n0, int N, bool vec_ok,
                                             float fill, float (&out)[32]) {
  if (vec_ok && n0 + 32 <= N) {
    const float4* p4 = reinterpret_cast<const float4*>(base + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(p4 + j);
      out[4 * j] = t.x; out[4 * j + 1] = t.y; out[4 * j + 2] = t.z; out[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = (n0 + j < N) ? __ldg(base + n0 + j) : fill;
  }
}

// Shared-memory matrix descriptor (descriptor version 1 = Blackwell).
//   K-major : SWIZZLE_128B (layout type 2): rows of 128 B (32 tf32 of K), 8-row swizzle atoms
//             1024 B apart (SBO); LBO unused.
//   MN-major: 32-bit operands only support SWIZZLE_128B_BASE32B (layout type 1, TMA swizzle
//             128B_ATOM_32B: 32 B chunks XOR-ed with k-row % 4): a [32 k-rows][32 mn] box per 32 MN
//             elements; boxes 4096 B apart (LBO), the 4-k-row atoms inside a box 512 B apart (SBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;          // descriptor version (sm_100)
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// ---------------------------------------------------------------------------------------------
// Stream-K bookkeeping (CTA-pair kernel).  Units are (tile, K block) pairs in tile-major order; pair q owns
// units [q * U / P, (q + 1) * U / P).  Its first segment may start inside a tile (kb0 > 0): that accumulator is a
// PARTIAL, written to the pair's workspace slot.  A segment that starts a tile but does not finish it makes the
// pair the tile's OWNER: it waits for the partials of the following pairs (each of which produced it as its very
// first segment, long before), adds them in pair order and runs the fused epilogue.  Partial writers never wait
// before their partial is out, and pairs are numbered against the block index (see the kernel), so there is no
// cycle and no co-residency requirement.
// ---------------------------------------------------------------------------------------------
constexpr int SK_SLOT_FLOATS = 8 * 8 * 128 * 4;          // [chunk 8][j 8][row 128] float4 per CTA

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct Work {
  int mn;                 // m/n tile index (nt = mn % n_tiles, mt = mn / n_tiles)
  int z;                  // split slab of the output map (0 unless split-K)
  int kb0, kb1;           // K blocks [kb0, kb1)
  int sk_mode;            // 0 = complete accumulator, 1 = partial (write to workspace), 2 = owner (add partials)
  int c_first, c_end;     // sk_mode 2: contributing pairs [c_first, c_end)
};
struct Sched {
  int tile, step;         // round-robin over tiles x splits
  long long u, u_end;     // stream-K unit range of this pair
};
__device__ __forceinline__ long long sk_first_unit(const GemmParams& p, int pair, int /*npairs*/) {
  return p.sk_b[pair];
}
__device__ __forceinline__ void sched_init(const GemmParams& p, Sched& s, int pair, int npairs) {
  s.tile = pair; s.step = npairs;
  s.u = s.u_end = 0;
  if (p.streamk) { s.u = sk_first_unit(p, pair, npairs); s.u_end = sk_first_unit(p, pair + 1, npairs); }
}
// Next work item of this pair; identical (warp-uniform) in the producer, the MMA issuer and every epilogue warp.
__device__ __forceinline__ bool sched_next(const GemmParams& p, Sched& s, int pair, int npairs, Work& w) {
  const int mn_tiles = p.m_tiles * p.n_tiles;
  if (!p.streamk) {
    if (s.tile >= mn_tiles * p.splits) return false;
    w.mn = s.tile % mn_tiles;
    w.z = s.tile / mn_tiles;
    w.kb0 = w.z * p.kb_per_split;
    w.kb1 = min(w.kb0 + p.kb_per_split, p.total_kb);
    w.sk_mode = 0; w.c_first = w.c_end = 0;
    s.tile += s.step;
    return true;
  }
  if (s.u >= s.u_end) return false;
  const int KB = p.total_kb;
  w.mn = static_cast<int>(s.u / KB);
  w.z = 0;
  w.kb0 = static_cast<int>(s.u - static_cast<long long>(w.mn) * KB);
  const long long left = s.u_end - s.u;
  w.kb1 = (left < KB - w.kb0) ? w.kb0 + static_cast<int>(left) : KB;
  s.u += w.kb1 - w.kb0;
  w.c_first = w.c_end = 0;
  if (w.kb0 > 0) {
    w.sk_mode = 1;
  } else if (w.kb1 == KB) {
    w.sk_mode = 0;
  } else {
    w.sk_mode = 2;
    const long long tile_end = static_cast<long long>(w.mn + 1) * KB;
    w.c_first = pair + 1;
    int c = pair + 1;
    while (c < npairs && sk_first_unit(p, c, npairs) < tile_end) ++c;
    w.c_end = c;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Epilogue of one accumulator, shared by both kernels: this warp's 32 rows (TMEM lane quadrant q) x its share of
// the tile's 32-column chunks.  TMEM -> registers -> fused math -> swizzled smem box -> TMA store.
// ---------------------------------------------------------------------------------------------
struct EpiWarp {
  uint32_t tmem_base, epi_base, side_base, side_bar;
  int q, ew, cpart, lane;
  uint32_t nstore, side_issued, side_waited;       // ring positions of the staging boxes / TMA-staged side boxes
};

__device__ __forceinline__ void epilogue_tile(const GemmParams& p, EpiWarp& w, int nt, int z, int row0,
                                              uint32_t tcol0, int sk_mode, int sk_cta, int sk_pair, int c_first,
                                              int c_end) {
  constexpr int CPARTS = NUM_EPI_WARPS / 4;
  const EpiParams& e = p.epi;
  const int lane = w.lane, q = w.q, ew = w.ew, cpart = w.cpart;
  const bool fused = (p.splits == 1) && sk_mode != 1;
  const bool side = fused && p.side_mode != 0;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  float rs = 1.0f;
  if (fused && e.row_scale != nullptr && row_ok) rs = __ldg(e.row_scale + row);
  const bool bias_vec = (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0;
  const bool aux_vec = ((reinterpret_cast<uintptr_t>(e.aux) & 15) == 0) && ((e.ld_aux & 3) == 0);
  const bool blend = fused && e.blend_x != nullptr;
  const bool blend_vec = ((reinterpret_cast<uintptr_t>(e.blend_x) & 15) == 0) && ((e.ld_blend & 3) == 0);
  const float beps = (blend && row_ok) ? __ldg(e.blend_eps + row) : 0.0f;
  float sumsq = 0.0f;
  if (sk_mode == 2) {                                  // the contributors' partials (this warp's share) are written
    if (lane == 0) {
      for (int pr = c_first; pr < c_end; ++pr) {
        const unsigned int* f = p.sk_flags + (static_cast<size_t>(pr) * 2 + sk_cta) * NUM_EPI_WARPS + ew;
        while (ld_acquire_gpu(f) == 0u) { }
      }
    }
    __syncwarp();
  }
  if (row0 < p.M) {
    const int nchunks = p.bn / 32;
    if (side && nt * p.bn + cpart * 32 < p.N && cpart < nchunks) {   // side box of this warp's first chunk
      if (lane == 0) {
        const uint32_t sl = ew * EPI_NBUF + (w.side_issued % EPI_NBUF);
        mbar_expect_tx(w.side_bar + 8 * sl, EPI_BOX_BYTES);
        tma_load_2d(w.side_base + sl * EPI_BOX_BYTES, &p.tma_side, w.side_bar + 8 * sl, nt * p.bn + cpart * 32, row0);
      }
      ++w.side_issued;
    }
    // stream-K owner: the FIRST contributor's partial of the next chunk is fetched one chunk ahead (under the
    // activation + staged store of the current one); un / c_pref carry it across iterations
    float4 un[8];
    int c_pref = -1;
    auto sk_slot = [&](int pr, int cc) {
      return reinterpret_cast<const float4*>(p.sk_ws + (static_cast<size_t>(pr) * 2 + sk_cta) * SK_SLOT_FLOATS) +
             static_cast<size_t>(cc) * 8 * 128 + (32 * q + lane);
    };
    if (sk_mode == 2 && c_first < c_end && cpart < nchunks && nt * p.bn + cpart * 32 < p.N) {
      const float4* s0 = sk_slot(c_first, cpart);
#pragma unroll
      for (int j = 0; j < 8; ++j) un[j] = __ldcg(s0 + j * 128);
      c_pref = cpart;
    }
    for (int c = cpart; c < nchunks; c += CPARTS) {
      const int n0 = nt * p.bn + c * 32;
      if (n0 >= p.N) break;
      float sv[32];                              // TMA-staged side operand of this chunk
      uint32_t v[32];
      tmem_ld32_issue(w.tmem_base + (static_cast<uint32_t>(32 * q) << 16) + tcol0 + static_cast<uint32_t>(c * 32), v);
      if (sk_mode == 1) {
        // stream-K partial: raw accumulator to this CTA's workspace slot, 512 contiguous bytes per instruction
        tmem_ld32_wait(v);
        float4* dst = reinterpret_cast<float4*>(p.sk_ws + (static_cast<size_t>(sk_pair) * 2 + sk_cta) * SK_SLOT_FLOATS) +
                      static_cast<size_t>(c) * 8 * 128 + (32 * q + lane);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j * 128] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        continue;
      }
      if (fused) {
        // epilogue operands are fetched while the TMEM load is in flight
        float t[32];
        if (e.bias != nullptr) {
          load_chunk32(e.bias, n0, p.N, bias_vec, 0.0f, t);
          tmem_ld32_wait(v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + t[j]);
        } else {
          tmem_ld32_wait(v);
        }
        if (sk_mode == 2) {
          // add the other pairs' partial accumulators of this tile, in pair order (deterministic).  Two contributors'
          // loads are in flight together: this fix-up sits on the kernel's critical path (diagnostic builds: the
          // stream-K tail is 16 of the layer-1 products' 76 us) and each load is an L2 round trip.
          auto add8 = [&](const float4 (&u)[8]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + u[j].x);
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + u[j].y);
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + u[j].z);
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + u[j].w);
            }
          };
          int pr = c_first;
          if (c_pref == c) {                    // first contributor: already in registers
            if (pr + 1 < c_end) {               // second one in flight while the first is added
              float4 u1[8];
              const float4* s1 = sk_slot(pr + 1, c);
#pragma unroll
              for (int j = 0; j < 8; ++j) u1[j] = __ldcg(s1 + j * 128);
              add8(un);
              add8(u1);
              pr += 2;
            } else {
              add8(un);
              pr += 1;
            }
          }
          for (; pr + 1 < c_end; pr += 2) {
            float4 u0[8], u1[8];
            const float4* s0 = sk_slot(pr, c);
            const float4* s1 = sk_slot(pr + 1, c);
#pragma unroll
            for (int j = 0; j < 8; ++j) u0[j] = __ldcg(s0 + j * 128);   // written by another SM during this launch: L2 only
#pragma unroll
            for (int j = 0; j < 8; ++j) u1[j] = __ldcg(s1 + j * 128);
            add8(u0);
            add8(u1);
          }
          if (pr < c_end) {
            float4 u0[8];
            const float4* s0 = sk_slot(pr, c);
#pragma unroll
            for (int j = 0; j < 8; ++j) u0[j] = __ldcg(s0 + j * 128);
            add8(u0);
          }
          const int cn2 = c + CPARTS;           // next chunk's first contributor: in flight under this chunk's store
          if (c_first < c_end && cn2 < nchunks && nt * p.bn + cn2 * 32 < p.N) {
            const float4* s0 = sk_slot(c_first, cn2);
#pragma unroll
            for (int j = 0; j < 8; ++j) un[j] = __ldcg(s0 + j * 128);
            c_pref = cn2;
          }
        }
        if (e.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float y = __uint_as_float(v[j]);
            v[j] = __float_as_uint(fmaxf(y, e.alpha * y));
          }
        }
        if (side) {
          const uint32_t sl = ew * EPI_NBUF + (w.side_waited % EPI_NBUF);
          mbar_wait(w.side_bar + 8 * sl, (w.side_waited / EPI_NBUF) & 1);
          ++w.side_waited;
          const uint32_t sb = w.side_base + sl * EPI_BOX_BYTES + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t r0, r1, r2, r3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                         : "r"(sb + (static_cast<uint32_t>(j ^ (lane & 7)) << 4))
                         : "memory");
            sv[4 * j] = __uint_as_float(r0); sv[4 * j + 1] = __uint_as_float(r1);
            sv[4 * j + 2] = __uint_as_float(r2); sv[4 * j + 3] = __uint_as_float(r3);
          }
          // The refill below overwrites this box through the async proxy (TMA), which is not ordered after generic-
          // proxy loads that have merely been ISSUED: without the proxy fence a late ld.shared could see the next
          // chunk's data (measured: ~2e-4 of the boxes under shared-memory pressure, r02 xhat_probe).  The fence
          // completes every lane's loads; the warp barrier then hands the slot to the elected lane.
          fence_proxy_async_smem();
          __syncwarp();                          // every lane has its values: the ring slot is free again
          const int cn = c + CPARTS;             // prefetch the next chunk's box while this one is processed
          if (cn < nchunks && nt * p.bn + cn * 32 < p.N) {
            if (lane == 0) {
              const uint32_t sn = ew * EPI_NBUF + (w.side_issued % EPI_NBUF);
              mbar_expect_tx(w.side_bar + 8 * sn, EPI_BOX_BYTES);
              tma_load_2d(w.side_base + sn * EPI_BOX_BYTES, &p.tma_side, w.side_bar + 8 * sn, nt * p.bn + cn * 32, row0);
            }
            ++w.side_issued;
          }
        }
        if (e.aux != nullptr) {
          if (!side) {
            if (row_ok) load_chunk32(e.aux + static_cast<size_t>(row) * e.ld_aux, n0, p.N, aux_vec, 1.0f, sv);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sl = (row_ok && !(sv[j] > 0.0f)) ? e.alpha : 1.0f;
            v[j] = __float_as_uint(__uint_as_float(v[j]) * sl);
          }
        }
        if (e.row_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * rs);
        }
        if (e.row_sumsq != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float y = __uint_as_float(v[j]);
            if (n0 + j < p.N) sumsq = fmaf(y, y, sumsq);
          }
        }
      } else {
        tmem_ld32_wait(v);
      }
      const uint32_t buf = w.epi_base + static_cast<uint32_t>((ew * EPI_NBUF + (w.nstore % EPI_NBUF)) * EPI_BOX_BYTES);
      if (lane == 0) tma_store_wait_read<EPI_NBUF - 1>();   // the box used EPI_NBUF stores ago is free again
      __syncwarp();
      const uint32_t rbase = buf + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t a = rbase + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);   // SWIZZLE_128B
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]),
                     "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&p.tma_d, buf, n0, row0, z);
        tma_store_commit();
      }
      ++w.nstore;
      if (blend) {
        if (!side) {
          if (row_ok) load_chunk32(e.blend_x + static_cast<size_t>(row) * e.ld_blend, n0, p.N, blend_vec, 0.0f, sv);
        }
        const float om = 1.0f - beps;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = __float_as_uint(row_ok ? beps * sv[j] + om * __uint_as_float(v[j]) : 0.0f);
        const uint32_t buf2 = w.epi_base + static_cast<uint32_t>((ew * EPI_NBUF + (w.nstore % EPI_NBUF)) * EPI_BOX_BYTES);
        if (lane == 0) tma_store_wait_read<EPI_NBUF - 1>();
        __syncwarp();
        const uint32_t rb2 = buf2 + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a = rb2 + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]),
                       "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&p.tma_d2, buf2, n0, row0, z);
          tma_store_commit();
        }
        ++w.nstore;
      }
    }
  }
  if (fused && e.row_sumsq != nullptr && row_ok)
    e.row_sumsq[static_cast<size_t>(nt * CPARTS + cpart) * p.M + row] = sumsq;
  if (sk_mode == 1) {
    // publish this warp's share of the partial: every lane's stores, then the flag (release)
    __threadfence();
    __syncwarp();
    if (lane == 0)
      st_release_gpu(p.sk_flags + (static_cast<size_t>(sk_pair) * 2 + sk_cta) * NUM_EPI_WARPS + ew, 1u);
  } else if (sk_mode == 2) {
    // every lane has consumed the partials: rearm the flags for the next launch
    __syncwarp();
    if (lane == 0)
      for (int pr = c_first; pr < c_end; ++pr)
        p.sk_flags[(static_cast<size_t>(pr) * 2 + sk_cta) * NUM_EPI_WARPS + ew] = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// The kernel.  Dynamic smem (after 1024 B alignment):
//   [stages] x { A tile (128 * msub) x 128 B | B tile bn x 128 B }, 8 x 4 KB epilogue boxes,
//   full[8], empty[8], tmem_full[2], tmem_empty[2] mbarriers, tmem base pointer.
// ---------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = BM * 128u * static_cast<uint32_t>(p.msub);
  const int TM = BM * p.msub;                  // rows of one CTA tile
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * 128u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t epi_base = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;
  const uint32_t side_base = epi_base + EPI_BUFS * EPI_BOX_BYTES;             // 2 side boxes per epilogue warp
  const uint32_t bar_base = side_base + (p.side_mode ? EPI_BUFS * EPI_BOX_BYTES : 0);
  const uint32_t full_bar = bar_base;                       // 8 x 8 B
  const uint32_t empty_bar = bar_base + 64;                 // 8 x 8 B
  const uint32_t tfull_bar = bar_base + 128;                // 2 x 8 B
  const uint32_t tempty_bar = bar_base + 144;               // 2 x 8 B
  const uint32_t tmem_slot = bar_base + 160;                // 4 B
  const uint32_t side_bar = bar_base + 176;                 // EPI_BUFS x 8 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    prefetch_tmap(&p.tma_d);
    if (p.epi.blend_x != nullptr) prefetch_tmap(&p.tma_d2);
    if (p.side_mode) prefetch_tmap(&p.tma_side);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full_bar + 8 * s, 1);
        mbar_init(empty_bar + 8 * s, static_cast<uint32_t>(p.cluster));   // one commit per cluster CTA
      }
      for (int i = 0; i < EPI_BUFS; ++i) mbar_init(side_bar + 8 * i, 1);
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar + 8 * a, 1);
        mbar_init(tempty_bar + 8 * a, NUM_EPI_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    // 2 accumulators of bn <= 256 fp32 columns: allocate all 512 columns (1 CTA per SM).
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();       // peers' barriers are initialised before any multicast
  tcgen05_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();                                  // everything above overlapped the previous kernel's tail

  // Persistent schedule over "cluster tiles": the C CTAs of a cluster take C consecutive M tiles of
  // the same N tile and split, and walk the K blocks in lockstep.
  const int C = p.cluster;
  const uint32_t crank = (C > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1u);
  const int total_tiles = p.m_groups * p.n_tiles * p.splits;
  const int tile0 = blockIdx.x / C, tile_step = gridDim.x / C;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int nt = tile % p.n_tiles;
      const int mt = ((tile / p.n_tiles) % p.m_groups) * C + static_cast<int>(crank);
      const int z = tile / (p.n_tiles * p.m_groups);
      const int m0 = mt * TM, n0 = nt * p.bn;
      const int kb0 = z * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.total_kb);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar + 8 * stage, phase ^ 1u);
        const uint32_t fb = full_bar + 8 * stage;
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t sb = sa + a_bytes;
        const int k0 = kb * BK;
        if (elect_one()) {
          // HBM latency (~2 us under load) exceeds what the smem ring can cover for a streaming
          // operand: pull the tile PREFETCH_DIST K blocks ahead into L2 first
          const int kp = kb + PREFETCH_DIST;
          if (kp < kb1) {
            const int kp0 = kp * BK;
            if (p.prefetch_a) {
              if (!A_MN) {
                tma_prefetch_2d(&p.tma_a, kp0, m0);
              } else {
                tma_prefetch_3d(&p.tma_a, 0, kp0, m0 / 32);
              }
            }
            if (p.prefetch_b && crank == 0) {
              if (!B_MN) {
                for (int r = 0; r < C; ++r) tma_prefetch_2d(&p.tma_b, kp0, n0 + r * (p.bn / C));
              } else {
                tma_prefetch_3d(&p.tma_b, 0, kp0, n0 / 32);
              }
            }
          }
          mbar_expect_tx(fb, stage_bytes);
          if (!A_MN) {
            tma_load_2d(sa, &p.tma_a, fb, k0, m0);                       // box [128 * msub rows][32 k]
          } else {
            tma_load_3d(sa, &p.tma_a, fb, 0, k0, m0 / 32);               // box [TM/32 chunks][32 k][32 m]
          }
          if (C == 1) {
            if (!B_MN) {
              tma_load_2d(sb, &p.tma_b, fb, k0, n0);                     // box [bn rows][32 k]
            } else {
              tma_load_3d(sb, &p.tma_b, fb, 0, k0, n0 / 32);             // box [bn/32 chunks][32 k][32 n]
            }
          } else {
            // every CTA fetches 1/C of the shared B tile and multicasts it to the whole cluster
            if (!B_MN) {
              const int rows_per = p.bn / C;                             // tma_b box = [bn/C rows][32 k]
              tma_load_2d_mc(sb + crank * rows_per * 128, &p.tma_b, fb, k0, n0 + crank * rows_per, cmask);
            } else {
              // (cluster multicast is an opt-in experiment: B map box = bn/32/C chunks)
              const int cpc = p.bn / 32 / C;
              tma_load_3d_mc(sb + crank * cpc * 4096, &p.tma_b, fb, 0, k0, n0 / 32 + crank * cpc, cmask);
            }
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                           ((B_MN ? 1u : 0u) << 16) | (static_cast<uint32_t>(p.bn >> 3) << 17) |
                           (static_cast<uint32_t>(BM >> 4) << 24);
    // smem descriptors: only the 14-bit start-address field changes per stage / K step
    const uint64_t a_proto = A_MN ? make_smem_desc(0, 4096, 512, 1) : make_smem_desc(0, 16, 1024, 2);
    const uint64_t b_proto = B_MN ? make_smem_desc(0, 4096, 512, 1) : make_smem_desc(0, 16, 1024, 2);
    const uint32_t a_hi = static_cast<uint32_t>(a_proto >> 32), b_hi = static_cast<uint32_t>(b_proto >> 32);
    const uint32_t a_lo0 = static_cast<uint32_t>(a_proto) | ((smem_base >> 4) & 0x3FFFu);
    const uint32_t b_lo0 = static_cast<uint32_t>(b_proto) | (((smem_base + a_bytes) >> 4) & 0x3FFFu);
    constexpr uint32_t a_kstep = A_MN ? (1024u >> 4) : (32u >> 4);
    constexpr uint32_t b_kstep = B_MN ? (1024u >> 4) : (32u >> 4);
    const uint32_t stage_step = stage_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int z = tile / (p.n_tiles * p.m_groups);
      const int kb0 = z * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.total_kb);
      const int acc = (p.nacc == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.nacc == 2) ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1u);
      tcgen05_after_sync();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.msub * p.bn);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tcgen05_after_sync();
        const uint32_t a_lo = a_lo0 + stage * stage_step;
        const uint32_t b_lo = b_lo0 + stage * stage_step;
        if (elect_one()) {
          for (int sub = 0; sub < p.msub; ++sub) {       // both 128-row sub-tiles reuse the B tile in smem
            const uint32_t a_sub = a_lo + sub * ((BM * 128u) >> 4);
            const uint32_t d_sub = d_tmem + static_cast<uint32_t>(sub * p.bn);
#pragma unroll
            for (int s = 0; s < BK / UMMA_K; ++s)
              umma_tf32(d_sub, pack64(a_sub + s * a_kstep, a_hi), pack64(b_lo + s * b_kstep, b_hi), idesc,
                        (kb > kb0 || s > 0) ? 1u : 0u);
          }
          // frees the smem slot (in every CTA of the cluster: peers multicast into it) when these MMAs retire
          if (C == 1) umma_commit(empty_bar + 8 * stage);
          else umma_commit_mc(empty_bar + 8 * stage, cmask);
          if (kb == kb1 - 1) umma_commit(tfull_bar + 8 * acc);   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM -> regs -> smem -> TMA store) =====================
    EpiWarp w;
    w.tmem_base = tmem_base; w.epi_base = epi_base; w.side_base = side_base; w.side_bar = side_bar;
    w.q = warp & 3;                                // TMEM lane quadrant this warp may access
    w.ew = warp - 2;                               // epilogue warp index 0..NUM_EPI_WARPS-1
    w.cpart = w.ew >> 2;                           // which share of the tile's 32-column chunks this warp drains
    w.lane = lane;
    w.nstore = 0; w.side_issued = 0; w.side_waited = 0;
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int nt = tile % p.n_tiles;
      const int mt = ((tile / p.n_tiles) % p.m_groups) * C + static_cast<int>(crank);
      const int z = tile / (p.n_tiles * p.m_groups);
      const int acc = (p.nacc == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.nacc == 2) ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tcgen05_after_sync();
      for (int sub = 0; sub < p.msub; ++sub)
        epilogue_tile(p, w, nt, z, mt * TM + sub * BM + 32 * w.q, static_cast<uint32_t>((acc * p.msub + sub) * p.bn),
                      0, 0, 0, 0, 0);
      // all of this warp's TMEM reads for the tile are done: hand the accumulator back
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
    }
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_before_sync();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();       // no peer commit / multicast may target an exited CTA
  tcgen05_after_sync();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// ---- CTA-pair (cta_group::2) helpers ----
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(local_bar), "r"(cta)
      : "memory");
}
// data lands in this CTA's smem, the transaction bytes complete on `bar` (a shared::cluster address:
// the pair leader's full barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant: tcgen05.mma cta_group::2, 256 x bn pair tiles (full TF32 rate: N/2 cycles per
// UMMA_K step instead of 43 + N/2, and half the B bytes per CTA => 6 stages instead of 4).
// ---------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_2cta_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // CTA pair (cta_group::2): the pair computes a 256 x bn tile; each CTA holds its 128 rows of A, HALF of
  // the B tile (bn/2 columns) and the accumulator of its own 128 rows.  The leader (cluster rank 0) issues
  // the MMAs for both; all loads of the pair complete on the leader's full barrier.
  const uint32_t a_bytes = BM * 128u;
  const int TM = 2 * BM;                       // rows of one pair tile
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn / 2) * 128u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t epi_base = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;
  const uint32_t side_base = epi_base + EPI_BUFS * EPI_BOX_BYTES;             // 2 side boxes per epilogue warp
  const uint32_t bar_base = side_base + (p.side_mode ? EPI_BUFS * EPI_BOX_BYTES : 0);
  const uint32_t full_bar = bar_base;                       // 8 x 8 B
  const uint32_t empty_bar = bar_base + 64;                 // 8 x 8 B
  const uint32_t tfull_bar = bar_base + 128;                // 2 x 8 B
  const uint32_t tempty_bar = bar_base + 144;               // 2 x 8 B
  const uint32_t tmem_slot = bar_base + 160;                // 4 B
  const uint32_t side_bar = bar_base + 176;                 // EPI_BUFS x 8 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    prefetch_tmap(&p.tma_d);
    if (p.epi.blend_x != nullptr) prefetch_tmap(&p.tma_d2);
    if (p.side_mode) prefetch_tmap(&p.tma_side);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full_bar + 8 * s, 1);
        mbar_init(empty_bar + 8 * s, 1);
      }
      for (int i = 0; i < EPI_BUFS; ++i) mbar_init(side_bar + 8 * i, 1);
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar + 8 * a, 1);
        mbar_init(tempty_bar + 8 * a, 2 * NUM_EPI_WARPS);   // the epilogue warps of both CTAs of the pair
      }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  cluster_sync_all();                          // the peer's barriers / TMEM exist before anything targets them
  tcgen05_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();                                  // everything above overlapped the previous kernel's tail

  // Persistent schedule: pair tiles (x split-K slabs) round-robin, or one contiguous stream-K unit range per pair;
  // the three warp roles walk the same sched_next sequence.
  const uint32_t crank = cluster_ctarank();
  const uint16_t cmask = 3;
  // work items come from sched_next (tiles or stream-K).  Stream-K numbers the pairs AGAINST the block index: a
  // tile's owner waits for the partials of HIGHER pairs only, i.e. of blocks launched before it, so any prefix of
  // the grid that is resident can finish on its own -- no co-residency requirement, whatever else holds SMs.
  const int npairs = gridDim.x / 2;
  const int pair = p.streamk ? npairs - 1 - static_cast<int>(blockIdx.x / 2) : static_cast<int>(blockIdx.x / 2);
#ifdef WGAN_DIAG_STAMPS
  if (threadIdx.x == 64) g_diag_stamps[blockIdx.x][0] = diag_now();
#endif

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    const uint32_t full_leader = mapa_cluster(full_bar, 0);       // all loads signal the leader's barrier
    int stage = 0;
    uint32_t phase = 0;
    Sched sc;
    sched_init(p, sc, pair, npairs);
    Work wk;
    while (sched_next(p, sc, pair, npairs, wk)) {
      const int nt = wk.mn % p.n_tiles;
      const int mt = wk.mn / p.n_tiles;
      const int m0 = mt * TM + static_cast<int>(crank) * BM;      // this CTA's 128 rows
      const int n0 = nt * p.bn + static_cast<int>(crank) * (p.bn / 2);   // this CTA's half of the B tile
      const int kb0 = wk.kb0, kb1 = wk.kb1;
      for (int kb = kb0; kb < kb1; ++kb) {
#if defined(WGAN_PAIR_BURST) && WGAN_PAIR_BURST > 1
        // diagnostic / experimental build: issue K blocks in groups -- wait until the whole group's slots are free
        if ((kb - kb0) % WGAN_PAIR_BURST == 0) {
          int s2 = stage;
          uint32_t ph2 = phase;
          for (int j = 0; j < WGAN_PAIR_BURST && kb + j < kb1; ++j) {
            mbar_wait(empty_bar + 8 * s2, ph2 ^ 1u);
            if (++s2 == p.stages) { s2 = 0; ph2 ^= 1u; }
          }
        }
#endif
        mbar_wait(empty_bar + 8 * stage, phase ^ 1u);
        const uint32_t fb = full_leader + 8 * stage;
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t sb = sa + a_bytes;
        const int k0 = kb * BK;
        if (elect_one()) {
          // a streaming operand's HBM latency under load (2-3 us) is more than the 6-stage ring covers at one
          // stage per 512 MMA cycles: pull its tile into L2 PREFETCH_DIST K blocks ahead, the ring then sees L2 latency
          const int kp = kb + PREFETCH_DIST;
          if (kp < kb1) {
            if (p.prefetch_a) {
              if (!A_MN) tma_prefetch_2d(&p.tma_a, kp * BK, m0);
              else       tma_prefetch_3d(&p.tma_a, 0, kp * BK, m0 / 32);
            }
            if (p.prefetch_b) {
              if (!B_MN) tma_prefetch_2d(&p.tma_b, kp * BK, n0);
              else       tma_prefetch_3d(&p.tma_b, 0, kp * BK, n0 / 32);
            }
          }
#ifdef WGAN_DIAG_NO_B       // diagnostic build (scripts/diag_build.sh): stream A only -- results are garbage
          if (crank == 0) mbar_expect_tx(full_bar + 8 * stage, 2 * a_bytes);
#else
          if (crank == 0) mbar_expect_tx(full_bar + 8 * stage, 2 * stage_bytes);   // both CTAs' bytes
#endif
          if (!A_MN) {
            tma_load_2d_2sm(sa, &p.tma_a, fb, k0, m0);                   // box [128 rows][32 k]
          } else {
            tma_load_3d_2sm(sa, &p.tma_a, fb, 0, k0, m0 / 32);           // box [4 chunks][32 k][32 m]
          }
#ifndef WGAN_DIAG_NO_B
          if (!B_MN) {
            tma_load_2d_2sm(sb, &p.tma_b, fb, k0, n0);                   // box [bn/2 rows][32 k]
          } else {
            tma_load_3d_2sm(sb, &p.tma_b, fb, 0, k0, n0 / 32);           // box [bn/64 chunks][32 k][32 n]
          }
#endif
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA only (whole warp loops, one elected lane issues) =====
    if (crank == 0) {
    // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                           ((B_MN ? 1u : 0u) << 16) | (static_cast<uint32_t>(p.bn >> 3) << 17) |
                           (static_cast<uint32_t>((2 * BM) >> 4) << 24);
    // smem descriptors: only the 14-bit start-address field changes per stage / K step
    const uint64_t a_proto = A_MN ? make_smem_desc(0, 4096, 512, 1) : make_smem_desc(0, 16, 1024, 2);
    const uint64_t b_proto = B_MN ? make_smem_desc(0, 4096, 512, 1) : make_smem_desc(0, 16, 1024, 2);
    const uint32_t a_hi = static_cast<uint32_t>(a_proto >> 32), b_hi = static_cast<uint32_t>(b_proto >> 32);
    const uint32_t a_lo0 = static_cast<uint32_t>(a_proto) | ((smem_base >> 4) & 0x3FFFu);
    const uint32_t b_lo0 = static_cast<uint32_t>(b_proto) | (((smem_base + a_bytes) >> 4) & 0x3FFFu);
    constexpr uint32_t a_kstep = A_MN ? (1024u >> 4) : (32u >> 4);
    constexpr uint32_t b_kstep = B_MN ? (1024u >> 4) : (32u >> 4);
    const uint32_t stage_step = stage_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    Sched sc;
    sched_init(p, sc, pair, npairs);
    Work wk;
    for (; sched_next(p, sc, pair, npairs, wk); ++it) {
      const int kb0 = wk.kb0, kb1 = wk.kb1;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1u);
      tcgen05_after_sync();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.bn);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tcgen05_after_sync();
        const uint32_t a_lo = a_lo0 + stage * stage_step;
        const uint32_t b_lo = b_lo0 + stage * stage_step;
        if (elect_one()) {
#if defined(WGAN_DIAG_SW_CONSUMER)   // diagnostic build: no tensor-core work at all, the slot is freed by plain arrives
          mbar_arrive(empty_bar + 8 * stage);
          mbar_arrive_cluster(empty_bar + 8 * stage, 1);
          if (kb == kb1 - 1) { mbar_arrive(tfull_bar + 8 * acc); mbar_arrive_cluster(tfull_bar + 8 * acc, 1); }
        }
        if (false) {
#endif
#ifdef WGAN_DIAG_NO_MMA      // diagnostic build: one MMA per K block instead of four (keeps the commit semantics)
          umma_tf32_2cta(d_tmem, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc, (kb > kb0) ? 1u : 0u);
#else
#pragma unroll
          for (int s = 0; s < BK / UMMA_K; ++s)
            umma_tf32_2cta(d_tmem, pack64(a_lo + s * a_kstep, a_hi), pack64(b_lo + s * b_kstep, b_hi), idesc,
                           (kb > kb0 || s > 0) ? 1u : 0u);
#endif
          // frees the smem slot in BOTH CTAs of the pair when these MMAs retire
          umma_commit_2cta_mc(empty_bar + 8 * stage, cmask);
          if (kb == kb1 - 1) umma_commit_2cta_mc(tfull_bar + 8 * acc, cmask);   // accumulators complete -> both epilogues
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    }  // leader
    __syncwarp();
  } else {
    // ===================== epilogue warps (TMEM -> regs -> smem -> TMA store) =====================
    EpiWarp w;
    w.tmem_base = tmem_base; w.epi_base = epi_base; w.side_base = side_base; w.side_bar = side_bar;
    w.q = warp & 3;                                // TMEM lane quadrant this warp may access
    w.ew = warp - 2;                               // epilogue warp index 0..NUM_EPI_WARPS-1
    w.cpart = w.ew >> 2;                           // which share of the tile's 32-column chunks this warp drains
    w.lane = lane;
    w.nstore = 0; w.side_issued = 0; w.side_waited = 0;
    Sched sc;
    sched_init(p, sc, pair, npairs);
    Work wk;
    int it = 0;
    for (; sched_next(p, sc, pair, npairs, wk); ++it) {
      const int nt = wk.mn % p.n_tiles;
      const int mt = wk.mn / p.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (it == 0) DIAG_STAMP(1);                 // first segment: waiting for its accumulator starts
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tcgen05_after_sync();
      DIAG_STAMP(2 + 2 * (it > 0 ? 1 : 0));       // accumulator of segment `it` complete (2: first, 4: last)
      g_diag_mode_store(wk.sk_mode, it);
#ifndef WGAN_DIAG_NO_EPI     // diagnostic build: accumulators are dropped (no partials, no fix-up, no stores)
      epilogue_tile(p, w, nt, wk.z, mt * TM + static_cast<int>(crank) * BM + 32 * w.q, static_cast<uint32_t>(acc * p.bn),
                    wk.sk_mode, static_cast<int>(crank), pair, wk.c_first, wk.c_end);
#endif
      DIAG_STAMP(3 + 2 * (it > 0 ? 1 : 0));       // epilogue of segment `it` done (3: first, 5: last)
      // all of this warp's TMEM reads for the tile are done: hand the accumulator back to the leader's MMA warp
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar + 8 * acc, 0);
    }
    if (lane == 0) tma_store_wait_all();
    DIAG_STAMP(6);
  }

  tcgen05_before_sync();
  __syncthreads();
  cluster_sync_all();                          // no commit / remote arrive may target an exited CTA
  tcgen05_after_sync();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

}  // namespace wg
