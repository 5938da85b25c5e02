// Row-local fused chains of the critic's hidden width (H = disc_internal_size <= 256, H % 8 == 0): two chained
// tcgen05 products per 128-row block, the intermediate produced by the epilogue warps straight into shared memory
// in the K-major SWIZZLE_128B tile layout the second product reads (it never round-trips through HBM as an MMA
// operand), plus everything row-wise in between.  One CTA per 128 rows, cta_group::1, non-persistent.
//
//   MODE 0  critic middle  [W:113-123,137] + tf.gradients of it:
//           h2 = lrelu(h1 W2 + b2);  d = h2 . w3 + b3;  delta2 = coef(row) w3 . slope(h2);
//           delta1 = (delta2 W2^T) . slope(h1)
//           (replaces: layer-2 GEMM, critic_head_kernel, layer-2 dgrad GEMM)
//   MODE 1  Gram form of the penalty (SURVEY 8d shortcut ii), interpolate rows:
//           T = g1 K;  n^2 = T . g1;  c = lambda 2/B (n - 1) / n;  v1 = c T . slope(h1);  (c g1) -> gram_t;
//           v2 = (v1 W2) . slope(h2)
//           (replaces: g1 K GEMM, gp_coef_gram_kernel, v1 W2 GEMM)
//
// Shared memory: A tile [KB][128 rows][32 k] (KB = ceil(H/32) boxes of 16 KB; stage-1 operand, then overwritten
// chunk by chunk with the stage-2 operand -- the stage-2 MMAs of K block c start as soon as chunk c is written), a
// ring of B tiles (bn x 128 B each) through which the producer streams the KB blocks of the stage-1 weight and then
// the KB blocks of the stage-2 weight, and the two per-column vectors of MODE 0.  TMEM: stage-1 accumulator in
// columns [0, bn), stage-2 accumulator in [256, 256 + bn).
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2.. = epilogue: a thread owns one row and every
// MID_CP-th 32-column chunk of it; row-wise reductions are per-thread partials combined through shared memory.
// Outputs leave through the thread's own slot of the A tile: chunk c of a row first holds the stage-1 operand, then
// the first output (read back transposed by its warp and written with coalesced 128 B row stores), then the
// stage-2 operand (= second output, TMA store), then the stage-2 result (TMA store).  Measured on the way: 16 B
// stores / loads scattered over 32 rows per instruction run at ~1 per cycle per SM (7 K cycles per [128, 200]
// operand), and 4 KB TMA store boxes at ~31 B/clk per SM -- hence the split between the two paths, and slope
// masks travel as bit arrays instead of being re-read as floats.  The stage-2 operand is rounded to TF32
// (cvt.rna, what a TMA load of it would do) before it is written, so delta2 / v1 in global memory carry TF32
// precision; their only other consumers are TF32 products and one bias-gradient column sum.
#pragma once
#include "gemm_sm100.cuh"

namespace wg {

// Epilogue warps: MID_EPI_WARPS / 4 warps per TMEM lane quadrant share a row's 32-column chunks round-robin (the
// epilogue is a chain of dependent latencies per chunk, so it scales with the number of warps, not with ALU width)
#ifndef WGAN_MID_EPI_WARPS
#define WGAN_MID_EPI_WARPS 16
#endif
constexpr int MID_EPI_WARPS = WGAN_MID_EPI_WARPS;
constexpr int MID_CP = MID_EPI_WARPS / 4;
constexpr int MID_THREADS = 64 + 32 * MID_EPI_WARPS;
constexpr int MID_MAX_STAGES = 4;

struct MidParams {
  CUtensorMap tma_a;        // stage-1 A [rows, H] K-major, box [128 rows][32 k]
  CUtensorMap tma_b1;       // stage-1 B, MN-major 3-D map, box [bn/32][32 k][32 n]
  CUtensorMap tma_b2;       // stage-2 B: MODE 0 K-major 2-D (box [bn rows][32 k]); MODE 1 MN-major 3-D
  CUtensorMap tma_o2;       // MODE 0: delta2   MODE 1: v1 (over h1 rows): the stage-2 operand itself, fp32 [rows, H]
  CUtensorMap tma_o3;       // MODE 0: delta1   MODE 1: v2
  float* o1;                // MODE 0: h2       MODE 1: c g1 (gram_t)       [rows, ld], written with coalesced stores
  uint32_t* bits1;          // [rows][8] slope bits (h > 0) of h1 / h2 per 32-column chunk: written by MODE 0,
  uint32_t* bits2;          //           read by MODE 1 (its rows are a row range of the same stacked buffers)
  int rows, H, bn, kblocks, stages;
  int ld;                   // leading dimension of o1
  float alpha;
  // MODE 0
  const float* b2;
  const float* w3;
  const float* b3;
  float* dout;
  int seg;                  // rows per coefficient segment
  float c0, c1, c2;
  // MODE 1
  float two_lambda_over_B;
  float* crow;
  float* pen;
  float* norms;
};

__device__ __forceinline__ uint32_t to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// this thread's row (r = 32 q + lane) of one [128 rows][32 floats] SWIZZLE_128B box
__device__ __forceinline__ void smem_row_load(uint32_t box, int r, float (&out)[32]) {
  const uint32_t rb = box + static_cast<uint32_t>(r) * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t a, b, c, d;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "r"(rb + (static_cast<uint32_t>(j ^ (r & 7)) << 4))
                 : "memory");
    out[4 * j] = __uint_as_float(a); out[4 * j + 1] = __uint_as_float(b);
    out[4 * j + 2] = __uint_as_float(c); out[4 * j + 3] = __uint_as_float(d);
  }
}
__device__ __forceinline__ void smem_row_store(uint32_t box, int r, const uint32_t (&v)[32]) {
  const uint32_t rb = box + static_cast<uint32_t>(r) * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rb + (static_cast<uint32_t>(j ^ (r & 7)) << 4)),
                 "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                 : "memory");
}

// The warp's 32 rows x 32 columns of a SWIZZLE_128B box -> global memory, one coalesced 128 B row per instruction.
__device__ __forceinline__ void warp_box_to_global(uint32_t box, int q, int lane, float* dst, int ld, int row0,
                                                   int rows, int n0, int H) {
  const bool col_ok = n0 + lane < H;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int rr = 32 * q + i;
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];"
                 : "=r"(u)
                 : "r"(box + static_cast<uint32_t>(rr) * 128u + (static_cast<uint32_t>((lane >> 2) ^ (rr & 7)) << 4) +
                       static_cast<uint32_t>(lane & 3) * 4u)
                 : "memory");
    if (col_ok && row0 + rr < rows) dst[static_cast<size_t>(row0 + rr) * ld + n0 + lane] = __uint_as_float(u);
  }
}
__device__ __forceinline__ void smem_vec_load(uint32_t addr, float (&out)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t a, b, c, d;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr + 16u * j) : "memory");
    out[4 * j] = __uint_as_float(a); out[4 * j + 1] = __uint_as_float(b);
    out[4 * j + 2] = __uint_as_float(c); out[4 * j + 3] = __uint_as_float(d);
  }
}

template <int MODE>
__global__ void __launch_bounds__(MID_THREADS, 1) mid_chain_kernel(const __grid_constant__ MidParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int KB = p.kblocks;
  const uint32_t a_base = base;
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * 128u;
  const uint32_t ring = a_base + static_cast<uint32_t>(KB) * 16384u;
  const uint32_t vecs = ring + static_cast<uint32_t>(p.stages) * b_bytes;   // b2[256] | w3[256] (MODE 0)
  const uint32_t parts = vecs + 2048u;                                      // [MID_CP][128] row partials
  const uint32_t bar = parts + 2048u;
  const uint32_t bfull = bar, bempty = bar + 32, afull = bar + 64, acc1_bar = bar + 72, acc2_bar = bar + 80,
                 tmem_slot = bar + 88, a2full = bar + 96;                   // a2full: 8 x 8 B, one per K block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a); prefetch_tmap(&p.tma_b1); prefetch_tmap(&p.tma_b2);
    prefetch_tmap(&p.tma_o2); prefetch_tmap(&p.tma_o3);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < MID_MAX_STAGES; ++s) { mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1); }
      mbar_init(afull, 1);
      mbar_init(acc1_bar, 1);
      for (int c = 0; c < 8; ++c) mbar_init(a2full + 8 * c, 4);
      mbar_init(acc2_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();

  const int m0 = blockIdx.x * 128;
  constexpr bool B2_MN = (MODE == 1);

  if (warp == 0) {
    // ===================== producer =====================
    if (elect_one()) {
      mbar_expect_tx(afull, static_cast<uint32_t>(KB) * 16384u);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(a_base + kb * 16384u, &p.tma_a, afull, kb * BK, m0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < 2 * KB; ++i) {
      mbar_wait(bempty + 8 * stage, phase ^ 1u);
      if (elect_one()) {
        const uint32_t fb = bfull + 8 * stage, dst = ring + stage * b_bytes;
        mbar_expect_tx(fb, b_bytes);
        if (i < KB) {
          tma_load_3d(dst, &p.tma_b1, fb, 0, i * BK, 0);
        } else if (B2_MN) {
          tma_load_3d(dst, &p.tma_b2, fb, 0, (i - KB) * BK, 0);
        } else {
          tma_load_2d(dst, &p.tma_b2, fb, (i - KB) * BK, 0);
        }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_common = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(p.bn >> 3) << 17) |
                                  (static_cast<uint32_t>(BM >> 4) << 24);
    const uint32_t idesc1 = idesc_common | (1u << 16);                        // B MN-major
    const uint32_t idesc2 = idesc_common | ((B2_MN ? 1u : 0u) << 16);
    const uint64_t a_proto = make_smem_desc(0, 16, 1024, 2);
    const uint64_t bmn_proto = make_smem_desc(0, 4096, 512, 1), bk_proto = make_smem_desc(0, 16, 1024, 2);
    const uint32_t a_hi = static_cast<uint32_t>(a_proto >> 32);
    const uint32_t a_lo0 = static_cast<uint32_t>(a_proto) | ((a_base >> 4) & 0x3FFFu);
    int stage = 0;
    uint32_t phase = 0;
    mbar_wait(afull, 0);
    tcgen05_after_sync();
    for (int i = 0; i < 2 * KB; ++i) {
      const bool second = i >= KB;
      const int kb = second ? i - KB : i;
      if (second) mbar_wait(a2full + 8 * kb, 0);   // all four epilogue warps have written K block kb of the operand
      mbar_wait(bfull + 8 * stage, phase);
      tcgen05_after_sync();
      const bool bmn = second ? B2_MN : true;
      const uint64_t b_proto = bmn ? bmn_proto : bk_proto;
      const uint32_t b_hi = static_cast<uint32_t>(b_proto >> 32);
      const uint32_t b_lo = static_cast<uint32_t>(b_proto) | (((ring + stage * b_bytes) >> 4) & 0x3FFFu);
      const uint32_t b_kstep = bmn ? (1024u >> 4) : (32u >> 4);
      const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(kb) * (16384u >> 4);
      const int ksteps = min(BK / UMMA_K, (p.H - kb * BK + UMMA_K - 1) / UMMA_K);
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + (second ? 256u : 0u);
        for (int s = 0; s < ksteps; ++s)
          umma_tf32(d_tmem, pack64(a_lo + s * (32u >> 4), a_hi), pack64(b_lo + s * b_kstep, b_hi),
                    second ? idesc2 : idesc1, (kb > 0 || s > 0) ? 1u : 0u);
        umma_commit(bempty + 8 * stage);
        if (kb == KB - 1) umma_commit(second ? acc2_bar : acc1_bar);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ===================== epilogue: one thread per row =====================
    const int q = warp & 3;
    const int cpart = (warp - 2) >> 2;         // which share of the row's chunks this warp owns
    const int r = 32 * q + lane;               // row inside the block
    const int row = m0 + r;
    const bool row_ok = row < p.rows;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
    uint32_t mbits[8];                         // MODE 0: slope bits of h1 per 32-column chunk
#pragma unroll
    for (int c = 0; c < 8; ++c) mbits[c] = 0u;
    float cb = 0.0f;
    if (MODE == 0) {
      // per-column vectors once per CTA: b2 and w3 (zero beyond H) -> shared memory, read back as broadcasts
      const int t = threadIdx.x - 64;
      for (int n = t; n < 256; n += 32 * MID_EPI_WARPS) {
        const float bv = n < p.H ? __ldg(p.b2 + n) : 0.0f, wv = n < p.H ? __ldg(p.w3 + n) : 0.0f;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(vecs + 4u * n), "r"(__float_as_uint(bv)) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(vecs + 1024u + 4u * n), "r"(__float_as_uint(wv)) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * MID_EPI_WARPS) : "memory");
    }
    mbar_wait(acc1_bar, 0);
    tcgen05_after_sync();

    if (MODE == 0) {
      const int sgm = row / p.seg;
      const float coef = sgm == 0 ? p.c0 : (sgm == 1 ? p.c1 : p.c2);
      float dacc = 0.0f;
#pragma unroll 1
      for (int c = cpart; c < KB; c += MID_CP) {
        const int n0 = c * 32;
        uint32_t v[32];
        tmem_ld32_issue(lane_taddr + static_cast<uint32_t>(n0), v);
        float t[32], w[32];
        smem_vec_load(vecs + 4u * n0, t);
        smem_vec_load(vecs + 1024u + 4u * n0, w);
        // slope bits of h1 (this row of the stage-1 operand) before delta2 takes its place
        const uint32_t box = a_base + static_cast<uint32_t>(c) * 16384u;
        float hh[32];
        smem_row_load(box, r, hh);
        uint32_t bits = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) bits |= (hh[j] > 0.0f ? 1u : 0u) << j;
        mbits[c] = bits;
        tmem_ld32_wait(v);
        uint32_t d2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float y = __uint_as_float(v[j]) + t[j];
          y = fmaxf(y, p.alpha * y);
          v[j] = __float_as_uint(y);
          dacc = fmaf(y, w[j], dacc);          // w is 0 beyond H
          d2[j] = to_tf32_rna(coef * w[j] * ((y > 0.0f) ? 1.0f : p.alpha));
        }
        const uint32_t wbox = box + static_cast<uint32_t>(32 * q) * 128u;   // this warp's 32 rows of the box
        if (row_ok && p.bits1 != nullptr) {
          uint32_t b2bits = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) b2bits |= (__uint_as_float(v[j]) > 0.0f ? 1u : 0u) << j;
          p.bits1[static_cast<size_t>(row) * 8 + c] = bits;
          p.bits2[static_cast<size_t>(row) * 8 + c] = b2bits;
        }
        smem_row_store(box, r, v);             // h2 leaves through the slot h1 occupied
        __syncwarp();
        warp_box_to_global(box, q, lane, p.o1, p.ld, m0, p.rows, n0, p.H);
        __syncwarp();
        smem_row_store(box, r, d2);
        fence_proxy_async_smem();
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(a2full + 8 * c);         // K block c of the stage-2 operand: this warp's 32 rows are in place
          tma_store_3d(&p.tma_o2, wbox, n0, m0 + 32 * q, 0);
          tma_store_commit();
        }
      }
      // d = h2 . w3 + b3: combine the per-warp partials of the row in a fixed order
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(parts + 4u * (cpart * 128 + r)), "r"(__float_as_uint(dacc)) : "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(32 * MID_EPI_WARPS) : "memory");
      if (cpart == 0 && row_ok) {
        float d = 0.0f;
#pragma unroll
        for (int k = 0; k < MID_CP; ++k) {
          uint32_t u;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(parts + 4u * (k * 128 + r)) : "memory");
          d += __uint_as_float(u);
        }
        p.dout[row] = d + __ldg(p.b3);
      }
    } else {
      // pass A: n^2 = T . g1 (g1 = this row of the stage-1 operand)
      float nsq = 0.0f;
#pragma unroll 1
      for (int c = cpart; c < KB; c += MID_CP) {
        uint32_t v[32];
        tmem_ld32_issue(lane_taddr + static_cast<uint32_t>(c * 32), v);
        float g[32];
        smem_row_load(a_base + static_cast<uint32_t>(c) * 16384u, r, g);
        tmem_ld32_wait(v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < p.H) nsq = fmaf(__uint_as_float(v[j]), g[j], nsq);
      }
      // every warp of the row needs the full sum before pass B: partials through shared memory, fixed order
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(parts + 4u * (cpart * 128 + r)), "r"(__float_as_uint(nsq)) : "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(32 * MID_EPI_WARPS) : "memory");
      nsq = 0.0f;
#pragma unroll
      for (int k = 0; k < MID_CP; ++k) {
        uint32_t u;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(parts + 4u * (k * 128 + r)) : "memory");
        nsq += __uint_as_float(u);
      }
      const float n = sqrtf(fmaxf(nsq, 0.0f));
      cb = row_ok ? p.two_lambda_over_B * (n - 1.0f) / n : 0.0f;
      if (row_ok && cpart == 0) {
        p.crow[row] = cb;
        p.pen[row] = (n - 1.0f) * (n - 1.0f);
        p.norms[row] = n;
      }
      // pass B: v1 = c T . slope(h1) -> stage-2 operand (+ global), c g1 -> gram_t
#pragma unroll 1
      for (int c = cpart; c < KB; c += MID_CP) {
        const int n0 = c * 32;
        uint32_t v[32];
        tmem_ld32_issue(lane_taddr + static_cast<uint32_t>(n0), v);
        const uint32_t box = a_base + static_cast<uint32_t>(c) * 16384u;
        float g[32];
        smem_row_load(box, r, g);
        const uint32_t hb = row_ok ? __ldg(p.bits1 + static_cast<size_t>(row) * 8 + c) : 0xFFFFFFFFu;
        tmem_ld32_wait(v);
        uint32_t cg[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sl = ((hb >> j) & 1u) ? 1.0f : p.alpha;
          const float t = __uint_as_float(v[j]);
          v[j] = row_ok ? to_tf32_rna(cb * t * sl) : 0u;
          cg[j] = __float_as_uint(cb * g[j]);
        }
        const uint32_t wbox = box + static_cast<uint32_t>(32 * q) * 128u;
        smem_row_store(box, r, cg);            // c g1 leaves through the slot g1 occupied
        __syncwarp();
        warp_box_to_global(box, q, lane, p.o1, p.ld, m0, p.rows, n0, p.H);
        __syncwarp();
        smem_row_store(box, r, v);
        fence_proxy_async_smem();
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(a2full + 8 * c);
          tma_store_3d(&p.tma_o2, wbox, n0, m0 + 32 * q, 0);
          tma_store_commit();
        }
      }
    }
    mbar_wait(acc2_bar, 0);
    tcgen05_after_sync();
#pragma unroll 1
    for (int c = cpart; c < KB; c += MID_CP) {
      const int n0 = c * 32;
      uint32_t v[32];
      tmem_ld32_issue(lane_taddr + 256u + static_cast<uint32_t>(n0), v);
      const uint32_t sb = (MODE == 0) ? mbits[c]
                                      : (row_ok ? __ldg(p.bits2 + static_cast<size_t>(row) * 8 + c) : 0xFFFFFFFFu);
      tmem_ld32_wait(v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = __float_as_uint(__uint_as_float(v[j]) * (((sb >> j) & 1u) ? 1.0f : p.alpha));
      // the stage-2 MMAs are complete, so the operand slot is free once its own TMA store has read it
      const uint32_t box = a_base + static_cast<uint32_t>(c) * 16384u;
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      smem_row_store(box, r, v);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&p.tma_o3, box + static_cast<uint32_t>(32 * q) * 128u, n0, m0 + 32 * q, 0);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace wg
