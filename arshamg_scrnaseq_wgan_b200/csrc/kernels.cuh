// Non-GEMM device kernels of the WGAN(-GP) hot path: split-K finish, FP32 verification GEMM,
// critic head (layer-3 row dot + loss partials + delta/g2 construction), gradient-penalty
// coefficient, interpolate blend, column sums (bias grads), fused multi-tensor TF-Adam,
// Philox RNG (noise prior / epsilon / cell indices), row gather, latent-fit residual, and the data-parallel
// gradient all-reduce over NVLink peer memory.
// Reference semantics: losses [W:137-138], LeakyReLU [W:69..117], Adam [A:140-151,214-225],
// prior [W:91-94], minibatch gather [W:188-191].  Math per SURVEY Appendix A.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gemm_sm100.cuh"

namespace wg {

// ---------------------------------------------------------------------------------------------
// Split-K finish: out[m,n] = epi(sum_s part[s][m,n]).  Also the epilogue of the FP32
// verification GEMM (splits = 1, part == out allowed: same thread reads then writes).
// row_sumsq (if requested) is written as a single "tile" (n_tiles = 1): one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void finish_splitk_kernel(const float* __restrict__ part, size_t slab_stride, int splits,
                                     float* out, int ld, int M, int N, EpiParams e) {
  pdl_launch_dependents();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
       row += gridDim.x * warps_per_block) {
    const float rs = e.row_scale ? e.row_scale[row] : 1.0f;
    float sumsq = 0.0f;
    for (int n = lane; n < N; n += 32) {
      const size_t off = static_cast<size_t>(row) * ld + n;
      float y = 0.0f;
      for (int s = 0; s < splits; ++s) y += part[s * slab_stride + off];
      if (e.bias) y += e.bias[n];
      if (e.act == 1) y = fmaxf(y, e.alpha * y);
      if (e.aux) {
        const float h = e.aux[static_cast<size_t>(row) * e.ld_aux + n];
        y *= (h > 0.0f) ? 1.0f : e.alpha;
      }
      y *= rs;
      sumsq += y * y;
      out[off] = y;
    }
    if (e.row_sumsq) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
      if (lane == 0) e.row_sumsq[row] = sumsq;
    }
  }
}

// Same as finish_splitk_kernel without row_sumsq: one thread per 4 consecutive columns, so that
// small outputs with many splits (K = batch wgrads) still fill the machine.
__global__ void finish_splitk_vec4_kernel(const float* __restrict__ part, size_t slab_stride, int splits,
                                          float* out, int ld, int M, int N, EpiParams e) {
  pdl_launch_dependents();
  pdl_wait();
  const int n4 = (N + 3) >> 2;
  const size_t total = static_cast<size_t>(M) * n4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / n4), n = static_cast<int>(i % n4) * 4;
    const size_t off = static_cast<size_t>(row) * ld + n;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 t = *reinterpret_cast<const float4*>(part + s * slab_stride + off);
      y.x += t.x; y.y += t.y; y.z += t.z; y.w += t.w;
    }
    float v[4] = {y.x, y.y, y.z, y.w};
    const float rs = e.row_scale ? e.row_scale[row] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j >= N) { v[j] = 0.0f; continue; }
      if (e.bias) v[j] += e.bias[n + j];
      if (e.act == 1) v[j] = fmaxf(v[j], e.alpha * v[j]);
      if (e.aux) {
        const float h = e.aux[static_cast<size_t>(row) * e.ld_aux + n + j];
        v[j] *= (h > 0.0f) ? 1.0f : e.alpha;
      }
      v[j] *= rs;
    }
    if (n + 4 <= N) {
      *reinterpret_cast<float4*>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int j = 0; n + j < N; ++j) out[off + j] = v[j];
    }
  }
}

// Small outputs with many splits (K = batch weight gradients of the 200 x 200 layers): 8 warps of a
// block each sum every 8th slab of the same 32 float4 groups, then reduce through shared memory
// (fixed order: deterministic).
__global__ void __launch_bounds__(256)
finish_splitk_tall_kernel(const float* __restrict__ part, size_t slab_stride, int splits, float* out, int ld,
                          int M, int N, EpiParams e) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n4 = (N + 3) >> 2;
  const size_t total = static_cast<size_t>(M) * n4;
  const size_t i = static_cast<size_t>(blockIdx.x) * 32 + lane;
  float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
  int row = 0, n = 0;
  size_t off = 0;
  if (i < total) {
    row = static_cast<int>(i / n4);
    n = static_cast<int>(i % n4) * 4;
    off = static_cast<size_t>(row) * ld + n;
    for (int s = w; s < splits; s += 8) {
      const float4 t = *reinterpret_cast<const float4*>(part + s * slab_stride + off);
      y.x += t.x; y.y += t.y; y.z += t.z; y.w += t.w;
    }
  }
  red[w][lane] = y;
  __syncthreads();
  if (w == 0 && i < total) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 t = red[k][lane];
      y.x += t.x; y.y += t.y; y.z += t.z; y.w += t.w;
    }
    float v[4] = {y.x, y.y, y.z, y.w};
    const float rs = e.row_scale ? e.row_scale[row] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j >= N) { v[j] = 0.0f; continue; }
      if (e.bias) v[j] += e.bias[n + j];
      if (e.act == 1) v[j] = fmaxf(v[j], e.alpha * v[j]);
      if (e.aux) {
        const float h = e.aux[static_cast<size_t>(row) * e.ld_aux + n + j];
        v[j] *= (h > 0.0f) ? 1.0f : e.alpha;
      }
      v[j] *= rs;
    }
    if (n + 4 <= N) {
      *reinterpret_cast<float4*>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int j = 0; n + j < N; ++j) out[off + j] = v[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FP32 CUDA-core GEMM (verification mode only: separates algorithm bugs from TF32 rounding).
// C[m,n] = sum_k A(m,k) B(n,k), element (m,k) of A at A[m*sam + k*sak], same for B.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_fp32_kernel(const float* __restrict__ A, long sam, long sak, const float* __restrict__ B,
                 long sbn, long sbk, float* __restrict__ C, int ldc, int M, int N, int K) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int mm, kk;
      if (sak == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.0f;
    }
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int nn, kk;
      if (sbk == 1) { kk = i & 15; nn = i >> 4; } else { nn = i & 63; kk = i >> 6; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? B[n * sbn + k * sbk] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[static_cast<size_t>(m) * ldc + n] = acc[i][j];
    }
  }
}

// Deterministic sums for the loss scalars: out_j[0] = scale_j * sum(ptr_j[0..n_j)); a job with n == 0 writes 0.
// They ride along as extra blocks of the column-sum launch (see colsum_partial_kernel).
struct SumJob {
  const float* ptr;
  float* out;
  int n;
  float scale;
};
struct SumJobs {
  SumJob job[3];
};

// ---------------------------------------------------------------------------------------------
// Critic head.  One warp per row r of h2 [rows, ld]:
//   d[r]     = h2[r,:] . w3 + b3                                   [W:119-123]
//   dlt[r,j] = coef(r) * w3[j] * slope(h2[r,j])                    (delta2 / g2)
// Rows are split in three consecutive segments of `seg` rows with coefficients c0, c1, c2
// (fake: +1/B, interpolate: 1, real: -1/B; nseg in {1,2,3}).
// ---------------------------------------------------------------------------------------------
__global__ void critic_head_kernel(const float* __restrict__ h2, int ld, int rows, int H,
                                   const float* __restrict__ w3, const float* __restrict__ b3,
                                   float* __restrict__ d, float* __restrict__ dlt, int seg, float c0,
                                   float c1, float c2, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int s = r / seg;
    const float coef = s == 0 ? c0 : (s == 1 ? c1 : c2);
    const float* hr = h2 + static_cast<size_t>(r) * ld;
    float* dr = dlt + static_cast<size_t>(r) * ld;
    float acc = 0.0f;
    for (int j = lane; j < H; j += 32) {
      const float h = hr[j], w = w3[j];
      acc = fmaf(h, w, acc);
      dr[j] = coef * w * ((h > 0.0f) ? 1.0f : alpha);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) d[r] = acc + b3[0];
  }
}

// ---------------------------------------------------------------------------------------------
// Gradient-penalty coefficient.  One warp per interpolate row b:
//   n_b = sqrt(sum_t nsq[t][b]);  c_b = lambda * 2/B * (n_b - 1) / n_b;  g1[b,:] *= c_b
//   pen[b] = (n_b - 1)^2   (summed with scale lambda/B by the sum blocks of the next column-sum launch)
// ---------------------------------------------------------------------------------------------
__global__ void gp_coef_kernel(const float* __restrict__ nsq, int ntiles, int tile_stride, int rows,
                               float two_lambda_over_B, float* __restrict__ c, float* __restrict__ pen,
                               float* __restrict__ norms, float* __restrict__ g1, int ld, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    // the tiles' partial sums of squares (52 at the reference dims) are read by different lanes and combined by an
    // xor butterfly (commutative at every stage: all lanes hold the same bits) instead of one serial chain of
    // L2-latency loads per row
    float s = 0.0f;
    for (int t = lane; t < ntiles; t += 32) s += nsq[static_cast<size_t>(t) * tile_stride + r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float n = sqrtf(s);
    const float cb = two_lambda_over_B * (n - 1.0f) / n;
    if (lane == 0) {
      c[r] = cb;
      pen[r] = (n - 1.0f) * (n - 1.0f);
      norms[r] = n;
    }
    float* gr = g1 + static_cast<size_t>(r) * ld;
    for (int j = lane; j < H; j += 32) gr[j] *= cb;
  }
}

// ---------------------------------------------------------------------------------------------
// Gram form of the penalty (shortcuts i + ii; exact in real arithmetic).
//
// l1_gram_finish_kernel: critic layer 1 is linear before its activation, so the interpolate's
// pre-activation is the blend of the two that are computed anyway:
//   a_f = x~ W1, a_r = x W1 (raw products, `splits` slabs of [2B, ld], fake rows first)
//   h1[b]      = lrelu(a_f + b1)              h1[B + b] = lrelu(a_r + b1)
//   h1[2B + b] = lrelu(eps_b a_r + (1 - eps_b) a_f + b1)
// `part` may alias `h1` (splits == 1, products written in place): a thread reads its own two
// float4 groups before it writes.
// ---------------------------------------------------------------------------------------------
__global__ void l1_gram_finish_kernel(const float* part, size_t slab_stride, int splits, float* h1, int ld,
                                      int B, int N, const float* __restrict__ bias,
                                      const float* __restrict__ eps, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  const int n4 = (N + 3) >> 2;
  const size_t total = static_cast<size_t>(B) * n4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / n4), n = static_cast<int>(i % n4) * 4;
    const size_t off_f = static_cast<size_t>(b) * ld + n;
    const size_t off_r = off_f + static_cast<size_t>(B) * ld;
    const size_t off_h = off_r + static_cast<size_t>(B) * ld;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f), r = f;
    for (int s = 0; s < splits; ++s) {
      const float4 tf = *reinterpret_cast<const float4*>(part + s * slab_stride + off_f);
      const float4 tr = *reinterpret_cast<const float4*>(part + s * slab_stride + off_r);
      f.x += tf.x; f.y += tf.y; f.z += tf.z; f.w += tf.w;
      r.x += tr.x; r.y += tr.y; r.z += tr.z; r.w += tr.w;
    }
    const float e = eps[b], g = 1.0f - e;
    const float vf[4] = {f.x, f.y, f.z, f.w}, vr[4] = {r.x, r.y, r.z, r.w};
    float of[4], orr[4], oh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j >= N) { of[j] = orr[j] = oh[j] = 0.0f; continue; }
      const float bj = bias[n + j];
      const float yf = vf[j] + bj, yr = vr[j] + bj, yh = e * vr[j] + g * vf[j] + bj;
      of[j] = fmaxf(yf, alpha * yf);
      orr[j] = fmaxf(yr, alpha * yr);
      oh[j] = fmaxf(yh, alpha * yh);
    }
    *reinterpret_cast<float4*>(h1 + off_f) = make_float4(of[0], of[1], of[2], of[3]);
    *reinterpret_cast<float4*>(h1 + off_r) = make_float4(orr[0], orr[1], orr[2], orr[3]);
    *reinterpret_cast<float4*>(h1 + off_h) = make_float4(oh[0], oh[1], oh[2], oh[3]);
  }
}

// gp_coef_gram_kernel: with K = W1^T W1 and T = g1 K (both from the tensor-core GEMM),
//   n_b^2 = g1_b K g1_b^T = T_b . g1_b,   c_b = lambda 2/B (n_b - 1) / n_b,   pen[b] = (n_b - 1)^2
//   v1    = c (g1 K) . s1      (= c (gx W1) . s1 of the explicit chain), in place over the h1 rows
//   T    <- c g1               (left operand of S = (c g1)^T g1, the Gram factor of dW1 = W1 S)
// One warp per interpolate row.
__global__ void gp_coef_gram_kernel(float* __restrict__ T, const float* __restrict__ g1, float* __restrict__ h1hat,
                                    int ld, int rows, int H, float two_lambda_over_B, float* __restrict__ c,
                                    float* __restrict__ pen, float* __restrict__ norms, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    float* tr = T + static_cast<size_t>(r) * ld;
    const float* gr = g1 + static_cast<size_t>(r) * ld;
    float* hr = h1hat + static_cast<size_t>(r) * ld;
    float acc = 0.0f;
    for (int j = lane; j < H; j += 32) acc = fmaf(tr[j], gr[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float n = sqrtf(fmaxf(acc, 0.0f));
    const float cb = two_lambda_over_B * (n - 1.0f) / n;
    if (lane == 0) {
      c[r] = cb;
      pen[r] = (n - 1.0f) * (n - 1.0f);
      norms[r] = n;
    }
    for (int j = lane; j < H; j += 32) {
      const float t = tr[j], g = gr[j], hh = hr[j];
      hr[j] = cb * t * ((hh > 0.0f) ? 1.0f : alpha);
      tr[j] = cb * g;
    }
  }
}

// x_hat[b,:] = eps[b] * x[b,:] + (1 - eps[b]) * x_tilde[b,:]      (SURVEY A.4), float4 over padded rows
__global__ void blend_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ xt,
                             int ldt, const float* __restrict__ eps, float* __restrict__ xh, int ldh,
                             int rows, int cols4) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(rows) * cols4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols4), c = static_cast<int>(i % cols4) * 4;
    const float e = eps[r], f = 1.0f - e;
    const float4 a = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ldx + c);
    const float4 b = *reinterpret_cast<const float4*>(xt + static_cast<size_t>(r) * ldt + c);
    float4 o;
    o.x = e * a.x + f * b.x;
    o.y = e * a.y + f * b.y;
    o.z = e * a.z + f * b.z;
    o.w = e * a.w + f * b.w;
    *reinterpret_cast<float4*>(xh + static_cast<size_t>(r) * ldh + c) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// Column sums over up to three row segments (bias grads, d w3):
//   out[n] = sum_seg scale_seg * sum_r X_seg[r, n]
// Stage 1: 1-D grid, job j owns ceil(N_j/128) * RC blocks (32 float4 columns x 8 row lanes each) -> part[rc][n], followed
// by one block per loss-scalar sum.  Stage 2 sums the RC partials in a fixed order.
// ---------------------------------------------------------------------------------------------
struct ColsumSeg {
  const float* ptr;
  int rows;
  int ld;
  float scale;
};
struct ColsumJob {
  ColsumSeg seg[3];
  float* out;
  int nseg;
  int N;
};
struct ColsumArgs {
  ColsumJob job[3];       // blockIdx.z selects the job
  int part_stride;        // floats of partial storage per job
  int njobs;
  int RC;                 // row chunks per column strip
  int block0[4];          // first block of each job in the 1-D grid; block0[njobs] = first loss-sum block
  float* zero_out;        // optional: one float to clear (d b3 = sum(+1/B) + sum(-1/B) = 0)
};

__global__ void colsum_partial_kernel(ColsumArgs a, float* __restrict__ part, SumJobs sums, int nsums) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][33];
  __shared__ float4 red4[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int lin = blockIdx.x;
  if (lin >= a.block0[a.njobs]) {
    // trailing blocks: one loss-scalar sum each, in a fixed order (eight independent accumulators per thread so
    // that the loads overlap instead of forming one latency chain)
    {
      const SumJob jb = sums.job[lin - a.block0[a.njobs]];
      const int bd = blockDim.x;
      float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int i = threadIdx.x;
      for (; i + 7 * bd < jb.n; i += 8 * bd) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a8[u] += jb.ptr[i + u * bd];
      }
      for (; i < jb.n; i += bd) a8[0] += jb.ptr[i];
      float acc = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (tx == 0) red[0][ty] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        float v = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += red[0][i];
        jb.out[0] = v * jb.scale;
      }
      __syncthreads();
    }
    return;
  }
  const int j = (lin >= a.block0[1] ? 1 : 0) + (lin >= a.block0[2] ? 1 : 0);
  const ColsumJob& jb = a.job[j];
  const int strips = (jb.N + 127) >> 7, local = lin - a.block0[j];
  const int n = (local % strips) * 128 + tx * 4;          // this thread's 4 columns (rows are 16 B aligned)
  const int RC = a.RC, rc = local / strips;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < jb.N) {
    for (int s = 0; s < jb.nseg; ++s) {
      const ColsumSeg sg = jb.seg[s];
      const int per = (sg.rows + RC - 1) / RC;
      const int r0 = rc * per, r1 = min(r0 + per, sg.rows);
      // four independent chains: the row loads overlap instead of serialising on their latency
      const float* col = sg.ptr + n;
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      int r = r0 + ty;
      for (; r + 24 < r1; r += 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 v = *reinterpret_cast<const float4*>(col + static_cast<size_t>(r + 8 * u) * sg.ld);
          t[u].x += v.x; t[u].y += v.y; t[u].z += v.z; t[u].w += v.w;
        }
      }
      for (; r < r1; r += 8) {
        const float4 v = *reinterpret_cast<const float4*>(col + static_cast<size_t>(r) * sg.ld);
        t[0].x += v.x; t[0].y += v.y; t[0].z += v.z; t[0].w += v.w;
      }
      acc.x += sg.scale * ((t[0].x + t[1].x) + (t[2].x + t[3].x));
      acc.y += sg.scale * ((t[0].y + t[1].y) + (t[2].y + t[3].y));
      acc.z += sg.scale * ((t[0].z + t[1].z) + (t[2].z + t[3].z));
      acc.w += sg.scale * ((t[0].w + t[1].w) + (t[2].w + t[3].w));
    }
  }
  red4[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < jb.N) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = red4[i][tx];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    float* dst = part + static_cast<size_t>(j) * a.part_stride + static_cast<size_t>(rc) * jb.N + n;
    const float o[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (n + c < jb.N) dst[c] = o[c];
  }
}

__global__ void colsum_final_kernel(ColsumArgs a, const float* __restrict__ part, int RC) {
  pdl_launch_dependents();
  pdl_wait();
  const ColsumJob& jb = a.job[blockIdx.y];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (a.zero_out != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.zero_out[0] = 0.0f;
  if (n >= jb.N) return;
  const float* mine = part + static_cast<size_t>(blockIdx.y) * a.part_stride + n;
  float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;       // four independent chains: the partial loads overlap
  int r = 0;
  for (; r + 4 <= RC; r += 4) {
    t0 += mine[static_cast<size_t>(r) * jb.N];
    t1 += mine[static_cast<size_t>(r + 1) * jb.N];
    t2 += mine[static_cast<size_t>(r + 2) * jb.N];
    t3 += mine[static_cast<size_t>(r + 3) * jb.N];
  }
  for (; r < RC; ++r) t0 += mine[static_cast<size_t>(r) * jb.N];
  jb.out[n] = (t0 + t1) + (t2 + t3);
}

// ---------------------------------------------------------------------------------------------
// Fused multi-tensor TF-Adam over a flat arena (params/grads/m/v share one layout) [A:140-151]:
//   lr_t = lr * sqrt(1 - b2p) / (1 - b1p);  m += (g - m)(1 - b1);  v += (g*g - v)(1 - b2)
//   var -= (m * lr_t) / (sqrt(v) + eps)            (epsilon OUTSIDE the bias correction)
// Explicit _rn intrinsics: no FMA contraction, so the update is bit-identical to the FP32 oracle.
// The beta powers live on the device (pw[0] = b1p, pw[1] = b2p, pw[2] = block counter) and are advanced by the
// last block of the launch, once every block has read them and all tensors are updated [A:214-225].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_t,
                                         float omb1, float omb2, float eps) {
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
  v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, lr_t), __fadd_rn(__fsqrt_rn(v), eps)));
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n4, float* __restrict__ pw, float lr,
                            float beta1, float beta2, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const float b1p = pw[0], b2p = pw[1];
  const float lr_t = __fdiv_rn(__fmul_rn(lr, __fsqrt_rn(__fsub_rn(1.0f, b2p))), __fsub_rn(1.0f, b1p));
  const float omb1 = __fsub_rn(1.0f, beta1), omb2 = __fsub_rn(1.0f, beta2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr_t, omb1, omb2, eps);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr_t, omb1, omb2, eps);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr_t, omb1, omb2, eps);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr_t, omb1, omb2, eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // _finish [A:214-225]: the beta powers advance once, after every block has read them.  pw[2] is a
  // block counter (bit pattern of an unsigned int); the last block to arrive updates the powers and rearms it.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* counter = reinterpret_cast<unsigned int*>(pw + 2);
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      pw[0] = __fmul_rn(b1p, beta1);
      pw[1] = __fmul_rn(b2p, beta2);
      *counter = 0u;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (replaces the reference's unseeded global NumPy RNG [W:92-93,188]).
// counter = (idx_lo, idx_hi, call_id, stream), key = seed; identical to oracle/wgan_oracle.py.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__constant__ float c_poisson1_cdf[12];

// noise_prior: |Normal(0, sigma) + Poisson(1)|  [W:91-94]; element index = global_row*dim + col.
// In every RNG kernel the call id is `call_id + *call_id_dev` (call_id_dev may be null): the device
// counter lets a captured CUDA graph draw fresh numbers on every replay.
__global__ void rng_noise_prior_kernel(float* __restrict__ out, int ld, long row0, int rows, int dim,
                                       float sigma, uint64_t seed, uint32_t call_id, uint32_t stream,
                                       const uint32_t* __restrict__ call_id_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (call_id_dev != nullptr) call_id += *call_id_dev;
  const size_t total = static_cast<size_t>(rows) * dim;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / dim), c = static_cast<int>(i % dim);
    const uint64_t idx = static_cast<uint64_t>(row0 + r) * static_cast<uint64_t>(dim) + c;
    const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32),
                                               call_id, stream),
                                    make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const float u1 = __fmul_rn(__fadd_rn(static_cast<float>(rnd.x >> 8), 0.5f), 5.9604644775390625e-8f);
    const float u2 = __fmul_rn(__fadd_rn(static_cast<float>(rnd.y >> 8), 0.5f), 5.9604644775390625e-8f);
    const float u3 = __fmul_rn(static_cast<float>(rnd.z >> 8), 5.9604644775390625e-8f);
    const float normal = __fmul_rn(sqrtf(__fmul_rn(-2.0f, logf(u1))), cosf(__fmul_rn(6.2831854820251465f, u2)));
    float pois = 0.0f;
#pragma unroll
    for (int k = 0; k < 12; ++k) pois += (u3 >= c_poisson1_cdf[k]) ? 1.0f : 0.0f;
    out[static_cast<size_t>(r) * ld + c] = fabsf(__fadd_rn(__fmul_rn(sigma, normal), pois));
  }
}

// U[0,1) per global row (epsilon of the interpolates).
__global__ void rng_uniform_kernel(float* __restrict__ out, long row0, int rows, uint64_t seed,
                                   uint32_t call_id, uint32_t stream, const uint32_t* __restrict__ call_id_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (call_id_dev != nullptr) call_id += *call_id_dev;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const uint64_t idx = static_cast<uint64_t>(row0 + r);
  const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), call_id, stream),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  out[r] = __fmul_rn(static_cast<float>(rnd.x >> 8), 5.9604644775390625e-8f);
}

// Uniform-with-replacement cell indices in [0, n_cells)  [W:188].
__global__ void rng_indices_kernel(int64_t* __restrict__ out, long row0, int rows, uint32_t n_cells,
                                   uint64_t seed, uint32_t call_id, uint32_t stream,
                                   const uint32_t* __restrict__ call_id_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (call_id_dev != nullptr) call_id += *call_id_dev;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const uint64_t idx = static_cast<uint64_t>(row0 + r);
  const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), call_id, stream),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  out[r] = static_cast<int64_t>((static_cast<uint64_t>(rnd.x) * n_cells) >> 32);
}

__global__ void counter_add_kernel(uint32_t* counter, uint32_t inc) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) *counter += inc;
}

// Minibatch gather from the resident cells-major matrix: out[r,:] = data[idx[r],:]  [W:188-191]
__global__ void gather_rows_kernel(const float* __restrict__ data, int ldd, const int64_t* __restrict__ idx,
                                   float* __restrict__ out, int ldo, int rows, int cols4) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(rows) * cols4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols4), c = static_cast<int>(i % cols4) * 4;
    const int64_t src = idx[r];
    *reinterpret_cast<float4*>(out + static_cast<size_t>(r) * ldo + c) =
        *reinterpret_cast<const float4*>(data + static_cast<size_t>(src) * ldd + c);
  }
}

// One launch for everything a critic update draws [W:188-193]: block-per-row, row r of the minibatch gets
//   idx_r = uniform index (Philox stream 3)        -> x_out[r,:] = data[idx_r,:]   (and idx_out[r] if wanted)
//   z_out[r,:] = noise prior (stream 1),  eps_out[r] = U[0,1) (stream 2, optional)
// Same counters as the three separate RNG kernels + gather_rows_kernel, so the values are bit-identical.
__global__ void __launch_bounds__(256)
draw_minibatch_kernel(const float* __restrict__ data, int ldd, uint32_t n_cells, float* __restrict__ x_out, int ldo,
                      int cols4, float* __restrict__ z_out, int ldz, int dim, float sigma,
                      float* __restrict__ eps_out, int64_t* __restrict__ idx_out, long row0, int rows,
                      uint64_t seed, uint32_t call_id, const uint32_t* __restrict__ call_id_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (call_id_dev != nullptr) call_id += *call_id_dev;
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const uint64_t gr = static_cast<uint64_t>(row0 + r);
    const uint32_t glo = static_cast<uint32_t>(gr), ghi = static_cast<uint32_t>(gr >> 32);
    const uint4 ri = philox4x32_10(make_uint4(glo, ghi, call_id, 3u), key);            // block-uniform
    const int64_t src = static_cast<int64_t>((static_cast<uint64_t>(ri.x) * n_cells) >> 32);
    const float4* from = reinterpret_cast<const float4*>(data + static_cast<size_t>(src) * ldd);
    float4* to = reinterpret_cast<float4*>(x_out + static_cast<size_t>(r) * ldo);
    // four independent 16 B loads in flight per thread before the first store
    int c = threadIdx.x;
    const int bd = blockDim.x;
    for (; c + 3 * bd < cols4; c += 4 * bd) {
      const float4 v0 = __ldg(from + c), v1 = __ldg(from + c + bd), v2 = __ldg(from + c + 2 * bd),
                   v3 = __ldg(from + c + 3 * bd);
      to[c] = v0; to[c + bd] = v1; to[c + 2 * bd] = v2; to[c + 3 * bd] = v3;
    }
    for (; c < cols4; c += bd) to[c] = __ldg(from + c);
    if (threadIdx.x == 0) {
      if (idx_out != nullptr) idx_out[r] = src;
      if (eps_out != nullptr) {
        const uint4 re = philox4x32_10(make_uint4(glo, ghi, call_id, 2u), key);
        eps_out[r] = __fmul_rn(static_cast<float>(re.x >> 8), 5.9604644775390625e-8f);
      }
    }
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
      const uint64_t idx = gr * static_cast<uint64_t>(dim) + c;
      const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), call_id, 1u), key);
      const float u1 = __fmul_rn(__fadd_rn(static_cast<float>(rnd.x >> 8), 0.5f), 5.9604644775390625e-8f);
      const float u2 = __fmul_rn(__fadd_rn(static_cast<float>(rnd.y >> 8), 0.5f), 5.9604644775390625e-8f);
      const float u3 = __fmul_rn(static_cast<float>(rnd.z >> 8), 5.9604644775390625e-8f);
      const float normal = __fmul_rn(sqrtf(__fmul_rn(-2.0f, logf(u1))), cosf(__fmul_rn(6.2831854820251465f, u2)));
      float pois = 0.0f;
#pragma unroll
      for (int k = 0; k < 12; ++k) pois += (u3 >= c_poisson1_cdf[k]) ? 1.0f : 0.0f;
      z_out[static_cast<size_t>(r) * ldz + c] = fabsf(__fadd_rn(__fmul_rn(sigma, normal), pois));
    }
  }
}

// Same draw with the row copies done by the bulk-copy engine instead of the load/store units: thread 0 of a block
// moves whole padded rows global -> shared -> global with cp.async.bulk (two row buffers in flight per block) while
// the other threads generate the row's noise.  Bit-identical to draw_minibatch_kernel.
__global__ void __launch_bounds__(256)
draw_minibatch_bulk_kernel(const float* __restrict__ data, int ldd, uint32_t n_cells, float* __restrict__ x_out, int ldo,
                           uint32_t row_bytes, float* __restrict__ z_out, int ldz, int dim, float sigma,
                           float* __restrict__ eps_out, int64_t* __restrict__ idx_out, long row0, int rows,
                           uint64_t seed, uint32_t call_id, const uint32_t* __restrict__ call_id_dev) {
  extern __shared__ __align__(128) uint8_t draw_smem[];
  __shared__ __align__(8) uint64_t bars[2];
  pdl_launch_dependents();
  const uint32_t buf0 = smem_u32(draw_smem);
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  if (call_id_dev != nullptr) call_id += *call_id_dev;
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  int it = 0;
  for (int r = blockIdx.x; r < rows; r += gridDim.x, ++it) {
    const uint64_t gr = static_cast<uint64_t>(row0 + r);
    const uint32_t glo = static_cast<uint32_t>(gr), ghi = static_cast<uint32_t>(gr >> 32);
    if (threadIdx.x == 0) {
      const uint4 ri = philox4x32_10(make_uint4(glo, ghi, call_id, 3u), key);
      const int64_t src = static_cast<int64_t>((static_cast<uint64_t>(ri.x) * n_cells) >> 32);
      const int b = it & 1;
      const uint32_t buf = buf0 + static_cast<uint32_t>(b) * row_bytes, bar = bar0 + 8u * b;
      // the store that read this buffer two rows ago must be done with it
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      mbar_expect_tx(bar, row_bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(buf), "l"(data + static_cast<size_t>(src) * ldd), "r"(row_bytes), "r"(bar) : "memory");
      if (idx_out != nullptr) idx_out[r] = src;
      if (eps_out != nullptr) {
        const uint4 re = philox4x32_10(make_uint4(glo, ghi, call_id, 2u), key);
        eps_out[r] = __fmul_rn(static_cast<float>(re.x >> 8), 5.9604644775390625e-8f);
      }
      mbar_wait(bar, static_cast<uint32_t>(it >> 1) & 1u);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(x_out + static_cast<size_t>(r) * ldo), "r"(buf), "r"(row_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    } else {
      for (int c = threadIdx.x - 1; c < dim; c += blockDim.x - 1) {
        const uint64_t idx = gr * static_cast<uint64_t>(dim) + c;
        const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), call_id, 1u), key);
        const float u1 = __fmul_rn(__fadd_rn(static_cast<float>(rnd.x >> 8), 0.5f), 5.9604644775390625e-8f);
        const float u2 = __fmul_rn(__fadd_rn(static_cast<float>(rnd.y >> 8), 0.5f), 5.9604644775390625e-8f);
        const float u3 = __fmul_rn(static_cast<float>(rnd.z >> 8), 5.9604644775390625e-8f);
        const float normal = __fmul_rn(sqrtf(__fmul_rn(-2.0f, logf(u1))), cosf(__fmul_rn(6.2831854820251465f, u2)));
        float pois = 0.0f;
#pragma unroll
        for (int k = 0; k < 12; ++k) pois += (u3 >= c_poisson1_cdf[k]) ? 1.0f : 0.0f;
        z_out[static_cast<size_t>(r) * ldz + c] = fabsf(__fadd_rn(__fmul_rn(sigma, normal), pois));
      }
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Strided 2-D copy (any alignment): dst[r, c] = src[r, c].  Used instead of a memcpy node where the copy sits
// between two kernels of a step, so that the programmatic-launch chain is not broken.
__global__ void copy2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows,
                              int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(rows) * cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    dst[static_cast<size_t>(r) * ldd + c] = src[static_cast<size_t>(r) * lds + c];
  }
}

// ---------------------------------------------------------------------------------------------
// Data-parallel gradient exchange over NVLink peer memory: in-place SUM over `world` ranks of one
// gradient arena (+ the loss scalars in its tail), one launch per rank, no NCCL.
// Every rank maps every rank's window (gradient arenas + staging + flags, one allocation).  All traffic that
// crosses NVLink is POSTED WRITES (a pull design measured ~135 GB/s: 16-byte loads over NVLink are round trips).
// Rank r owns slice r of the arena; block b of rank r
//   1. pushes its part of every foreign slice k of its own arena into rank k's staging row r,
//   2. tells every peer "my contributions have landed" (arrive1[b][r] = epoch) and waits for the same from all,
//   3. adds, in rank order, the W contributions to its part of slice r (own arena + staging rows: local reads; every
//      replica therefore ends up with bit-identical sums) and pushes the sum into every rank's arena,
//   4. tells every peer "my sums have landed" (arrive2[b][r] = epoch) and waits for every peer's block b.
// When all blocks of a rank have left step 4 its whole arena holds the global sum and nobody writes to its
// window any more.  The epoch lives in the window (advanced by the kernel), so a CUDA-graph replay works.
// Flags are system-scope release/acquire; gradient values move as 16-byte relaxed.sys accesses.
// ---------------------------------------------------------------------------------------------
constexpr int DP_MAX_WORLD = 16;
constexpr int DP_MAX_BLOCKS = 148;
constexpr int DP_THREADS = 1024;        // block size (WGAN_B200_DP_THREADS lowers it for experiments)
constexpr int DP_FLAG_WORDS = DP_MAX_BLOCKS * (2 * DP_MAX_WORLD + 1);   // arrive1, arrive2 [block][rank]; epoch[block]

struct DpParams {
  float* peer[DP_MAX_WORLD];     // window base of every rank, as mapped in this process
  int rank, world;
  size_t grad_off;               // floats from the window base to this net's gradient arena
  size_t stage_off;              // floats from the window base to the staging rows [world][stage_stride]
  size_t stage_stride;           // floats per staging row (>= the largest slice)
  size_t flags_off;              // floats from the window base to the flag words
  size_t n4;                     // float4 groups in the arena (scalar tail included)
  int early_trigger;             // 1: release the next kernel's launch at once (one rank per GPU)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Gradient values: the staging rows are written by peers into THIS GPU's memory (home L2), so an L2 load (ld.cg)
// after the acquire sees them; pushes are plain 16-byte stores that the release + system fence of the rendezvous
// orders (WGAN_DP_SYS_ACCESS=1 builds use relaxed.sys accesses instead: measured slower, same results).
#ifndef WGAN_DP_SYS_ACCESS
#define WGAN_DP_SYS_ACCESS 0
#endif
__device__ __forceinline__ float4 dp_load_v4(const float4* p) {
#if WGAN_DP_SYS_ACCESS
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ void dp_store_v4(float4* p, float4 v) {
#if WGAN_DP_SYS_ACCESS
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#else
  *p = v;
#endif
}

// One cross-GPU rendezvous of block b: publish `epoch` in slot [b][rank] of every rank's `which` array (0 =
// arrive1, 1 = arrive2), then wait until every rank has published it here.  Called by all threads of the block.
__device__ __forceinline__ void dp_rendezvous(const DpParams& p, int which, unsigned int epoch) {
  __syncthreads();                                  // every thread's earlier accesses precede the release below
  const int t = threadIdx.x;
  if (t < p.world) {
    const size_t slot = static_cast<size_t>(which) * DP_MAX_BLOCKS * DP_MAX_WORLD + blockIdx.x * DP_MAX_WORLD;
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned int*>(p.peer[t] + p.flags_off) + slot + p.rank, epoch);
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(p.peer[p.rank] + p.flags_off) + slot + t;
    while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) { }
  }
  __syncthreads();
}

// this block's share [b0, b1) of the float4 range [lo, hi)
__device__ __forceinline__ void dp_block_part(size_t lo, size_t hi, size_t& b0, size_t& b1) {
  const size_t per = (hi - lo + gridDim.x - 1) / gridDim.x;
  b0 = lo + per * blockIdx.x;
  if (b0 > hi) b0 = hi;
  b1 = b0 + per < hi ? b0 + per : hi;
}

// Launched as a few fat blocks (16 x 1024 threads by default, engine.cu): it runs on a second stream beside the
// generator forward of the next update, and what matters there is how little it disturbs those kernels, not its own
// speed -- measured, the fewer SMs it touches the better (a 1024-thread block cannot share an SM with a persistent
// GEMM CTA, so those 16 SMs simply join the GEMM late), whereas one small block on every SM slowed every CTA a little
// and the whole persistent grid with it.
__global__ void __launch_bounds__(DP_THREADS, 1)
dp_allreduce_kernel(const DpParams p) {
  // early_trigger = 0 (several ranks on one GPU: tests): no griddepcontrol.launch_dependents before the end --
  // blocks of the next kernel parked on the SMs could keep a PEER's blocks (same GPU, other stream) from becoming
  // resident while this grid spins on them.  One rank per GPU: trigger at once, the successor's launch is hidden.
  if (p.early_trigger) pdl_launch_dependents();
  pdl_wait();
  __shared__ unsigned int s_epoch;
  unsigned int* my_flags = reinterpret_cast<unsigned int*>(p.peer[p.rank] + p.flags_off);
  unsigned int* epoch_word = my_flags + 2 * DP_MAX_BLOCKS * DP_MAX_WORLD + blockIdx.x;
  if (threadIdx.x == 0) s_epoch = *epoch_word + 1u;
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const int W = p.world, R = p.rank;
  const size_t nthr = blockDim.x;            // <= DP_THREADS (engine: WGAN_B200_DP_THREADS)
  float4* mine = reinterpret_cast<float4*>(p.peer[R] + p.grad_off);
  // 1. scatter: my contribution to every foreign slice -> its owner's staging row R
  for (int dk = 1; dk < W; ++dk) {
    const int k = (R + dk) % W;                    // start with the next rank: spreads the first writes over the links
    const size_t lo = p.n4 * k / W, hi = p.n4 * (k + 1) / W;
    size_t b0, b1;
    dp_block_part(lo, hi, b0, b1);
    float4* dst = reinterpret_cast<float4*>(p.peer[k] + p.stage_off + static_cast<size_t>(R) * p.stage_stride) - lo;
    size_t i = b0 + threadIdx.x;
    for (; i + nthr < b1; i += 2 * nthr) {
      const float4 v0 = mine[i], v1 = mine[i + nthr];
      dp_store_v4(dst + i, v0); dp_store_v4(dst + i + nthr, v1);
    }
    if (i < b1) dp_store_v4(dst + i, mine[i]);
  }
  dp_rendezvous(p, 0, epoch);
  // 3. reduce my part of slice R in rank order, broadcast the sum
  {
    const size_t lo = p.n4 * R / W, hi = p.n4 * (R + 1) / W;
    size_t b0, b1;
    dp_block_part(lo, hi, b0, b1);
    const float4* stage = reinterpret_cast<const float4*>(p.peer[R] + p.stage_off) - lo;
    const size_t row4 = p.stage_stride / 4;
    for (size_t i = b0 + threadIdx.x; i < b1; i += nthr) {
      float4 acc = (R == 0) ? mine[i] : dp_load_v4(stage + i);
      for (int k = 1; k < W; ++k) {
        const float4 v = (k == R) ? mine[i] : dp_load_v4(stage + static_cast<size_t>(k) * row4 + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      mine[i] = acc;
      for (int dk = 1; dk < W; ++dk)
        dp_store_v4(reinterpret_cast<float4*>(p.peer[(R + dk) % W] + p.grad_off) + i, acc);
    }
  }
  dp_rendezvous(p, 1, epoch);
  if (threadIdx.x == 0) *epoch_word = epoch;
  if (!p.early_trigger) pdl_launch_dependents();
}

// Latent fit (build-defined): r = G(z) - x; loss[b] = sum r^2; d3 = 2 r slope(G(z)).
// One block per row.
__global__ void latent_residual_kernel(const float* __restrict__ xt, int ldt, const float* __restrict__ x,
                                       int ldx, float* __restrict__ d3, int ldd, float* __restrict__ loss,
                                       int cols, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const int r = blockIdx.x;
  float acc = 0.0f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float g = xt[static_cast<size_t>(r) * ldt + c];
    const float d = g - x[static_cast<size_t>(r) * ldx + c];
    acc += d * d;
    d3[static_cast<size_t>(r) * ldd + c] = 2.0f * d * ((g > 0.0f) ? 1.0f : alpha);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) loss[r] = v;
  }
}

}  // namespace wg
