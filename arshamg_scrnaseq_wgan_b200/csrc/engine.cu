// Host engine + C ABI (include/wgan_b200.h) of the B200-native WGAN(-GP) hot path.
// One handle per GPU owns: flat parameter / gradient / Adam-m / Adam-v arenas per network (TF
// variable order [W:140-142]), the activation workspace, and the split-K workspace.  Every step
// is a fixed sequence of asynchronous launches on the caller's stream (CUDA-graph capturable:
// no host state changes per step; Adam's beta powers live on the device).
//
// Row layout of the critic-step activation buffers (Bl = local rows):
//     [ fake (x~ = G(z)) | interpolate (x^, later gx)  -- only if lambda_gp != 0 | real (x) ]
// so that the wgrad products over the stacked rows are single GEMMs with K = 2..3 Bl.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <algorithm>
#include <utility>
#include <string>
#include <vector>
#include <unistd.h>

#include "../../include/wgan_b200.h"
#include "gemm_sm100.cuh"
#include "kernels.cuh"
#include "mid_chain_sm100.cuh"

using namespace wg;

namespace {

thread_local std::string g_create_error;
std::atomic<long long> g_pipeline_launches{0};   // launches of the handle-less input-pipeline kernels

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

constexpr int kB200Sms = 148;      // grid sizing of the handle-less input-pipeline kernels (sm_100a only)
inline int round4(int x) { return (x + 3) & ~3; }
inline size_t round64(size_t x) { return (x + 63) & ~static_cast<size_t>(63); }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

struct TensorDesc {
  const char* name;
  int net, rows, cols, ld;
  size_t offset;
};

struct Part {          // one product of a split-K accumulated GEMM
  const float* A;
  int lda;
  const float* B;
  int ldb;
  int K;
};

}  // namespace

struct wgan_handle {
  wgan_config cfg;
  int Z, Hg, G, Hd, R;
  int Zl, Hgl, Gl, Hdl;             // padded leading dimensions
  int num_sms;
  EncodeTiledFn encode;
  std::string err;
  int64_t launches;

  TensorDesc td[12];
  size_t arena_n[2];                 // floats per net arena (incl. 4-float scalar tail)
  float* arena[2][4];                // [net][kind]
  float* powers[2];                  // device {beta1_power, beta2_power} per optimizer

  // workspace
  float *zbuf, *gh1, *gh2, *xcat, *h1, *h2, *d2, *d1, *dout, *v2, *gd2, *gd1, *dz;
  float *nsq, *crow, *pen, *norms, *colpart, *splitws, *sample_tmp;
  uint32_t *bits1, *bits2;           // [3R][8] slope bits of h1 / h2 (mid_chain_sm100.cuh)
  float *gram_k, *gram_s, *gram_t;   // Gram form of the penalty: W1^T W1, (c g1)^T g1 [Hd, Hd]; g1 K [R, Hd]
  int colpart_stride;
  size_t splitws_n;
  int nsq_tiles_max;
  int last_gx_rows;
  int gemm_cluster;                  // preferred cluster size of the tensor-core GEMM (1, 2 or 4)
  int gemm_msub;                     // 0 = auto, 1 / 2 = force 128- / 256-row CTA tiles
  int gemm_2cta;                     // 1 = CTA-pair (cta_group::2) kernel
  int sm_cap;                        // SMs the persistent GEMM grids may occupy (= num_sms, less inside wgan_*_prepare)
  int prepare_reserve;               // env WGAN_B200_PREPARE_SM_RESERVE: SMs left free for a concurrent collective
                                     // while the prepare calls run next to a gradient all-reduce (0 = off, default)
  // Prepared state.  wgan_critic_prepare / wgan_generator_prepare run the part of a step that does not depend on
  // the critic's parameters and hand back a token (= the workspace generation they left behind).  Every entry
  // point that writes the shared activation workspace advances `ws_gen`, so a token is honoured only while the
  // prepared activations are still intact; a stale token just means the work is redone.
  uint64_t ws_gen;
  uint64_t prep_token; int prep_rows;      // critic: token of the last critic_prepare (0 = none)
  uint64_t gprep_token; int gprep_rows;    // generator
  unsigned int* sk_flags;                  // stream-K partial-ready flags [pair][cta][epilogue warp] (gemm_sm100.cuh)
  bool attr_tc[4], attr_tc2[4], attr_mid[2];   // cudaFuncSetAttribute done for this handle's device, per instantiation
  int max_clusters[4][5];
  // data-parallel window: both gradient arenas + the cross-GPU flags in ONE allocation that peers map
  // (cudaIpc handle or, inside one process, the pointer itself); see dp_allreduce_kernel in kernels.cuh
  float* dp_win; size_t dp_win_bytes; size_t dp_grad_off[2]; size_t dp_flags_off;   // offsets in floats
  size_t dp_stage_off, dp_stage_stride;    // staging rows [dp_world][dp_stage_stride]: peers push their contributions here
  int dp_rank, dp_world, dp_blocks; bool dp_connected;
  bool dp_early_trigger;                   // every rank has its own GPU (wgan_dp_connect): see dp_allreduce_kernel
  cudaStream_t dp_stream; cudaEvent_t dp_fork, dp_join;   // wgan_generator_grads_exchange: exchange next to the backward pass
  // bias-gradient / loss column sums run on a second stream next to the weight-gradient GEMMs (they only read)
  cudaStream_t aux_stream; cudaEvent_t aux_fork, aux_join; bool aux_overlap;
  float* dp_peer[WGAN_DP_MAX_WORLD];       // window base of every rank as mapped here (own rank: dp_win)
  bool dp_peer_ipc[WGAN_DP_MAX_WORLD];     // opened with cudaIpcOpenMemHandle (closed in wgan_destroy)

  float* P(int t) { return arena[td[t].net][WGAN_KIND_PARAM] + td[t].offset; }
  float* Gr(int t) { return arena[td[t].net][WGAN_KIND_GRAD] + td[t].offset; }
  float* scalars(int net) { return arena[net][WGAN_KIND_GRAD] + arena_n[net] - 4; }
};

namespace {

enum { T_WG1 = 0, T_BG1, T_WG2, T_BG2, T_WG3, T_BG3, T_W1, T_B1, T_W2, T_B2, T_W3, T_B3 };

#define CK(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      char buf_[512];                                                                            \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),        \
               __FILE__, __LINE__);                                                              \
      h->err = buf_;                                                                             \
      return -1;                                                                                 \
    }                                                                                            \
  } while (0)

#define FAIL(...)                                  \
  do {                                             \
    char buf_[512];                                \
    snprintf(buf_, sizeof buf_, __VA_ARGS__);      \
    h->err = buf_;                                 \
    return -2;                                     \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int r_ = (expr);         \
    if (r_ != 0) return r_;  \
  } while (0)

inline int ew_grid(size_t n, int block, int sms) {
  size_t g = (n + block - 1) / block;
  size_t cap = static_cast<size_t>(sms) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// Every kernel is launched with programmatic stream serialisation (PDL): a dependent grid may be scheduled
// while its predecessor drains; each kernel executes griddepcontrol.wait before it touches global memory, so
// ordering is unchanged, only the launch gap (and, for the GEMMs, the barrier/TMEM prologue) overlaps.
const bool g_pdl = getenv("WGAN_B200_NO_PDL") == nullptr;

template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface in post_launch
}
#define KLAUNCH(kernel, grid, block, smem, stream, ...) \
  launch_pdl(kernel, dim3(grid), dim3(block), static_cast<size_t>(smem), stream, __VA_ARGS__)

int post_launch(wgan_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "launch of %s failed: %s", what, cudaGetErrorString(e));
    h->err = buf;
    return -3;
  }
  h->launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Tensor maps
// ------------------------------------------------------------------------------------------
int make_map_2d(wgan_handle* h, CUtensorMap* m, const float* ptr, uint64_t inner, uint64_t outer,
                uint64_t ld_floats, uint32_t box_inner, uint32_t box_outer, bool mn_major) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld_floats & 3) != 0)
    FAIL("TMA operand must be 16-byte aligned with ld %% 4 == 0 (ptr=%p ld=%llu)", (const void*)ptr,
         (unsigned long long)ld_floats);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_floats * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  // operands are read as TF32 (round-to-nearest on load); MN-major 32-bit operands must use the
  // 32-byte-atom flavour of the 128 B swizzle (see make_smem_desc in gemm_sm100.cuh)
  static const bool plain_f32 = getenv("WGAN_B200_TMA_F32") != nullptr;   // experiment: no TF32 rounding in TMA
  CUresult r = h->encode(m, plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2,
                         const_cast<float*>(ptr), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(2d) failed with %d", (int)r);
  return 0;
}

// MN-major operand [K rows, MN cols] (MN contiguous) as a 3-D map (32 floats, K, ceil(MN/32) chunks) with
// strides (ld, 32 floats): one box of (32, 32, chunks_per_box) lands as the [chunk][k][32] smem tile.  The last
// chunk of a row may read up to 124 B past column MN (its results are clipped by the store); every
// device buffer of the engine carries that slack.
int make_map_mn3d(wgan_handle* h, CUtensorMap* m, const float* ptr, uint64_t mn, uint64_t k, uint64_t ld_floats,
                  uint32_t chunks_per_box) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld_floats & 3) != 0)
    FAIL("TMA operand must be 16-byte aligned with ld %% 4 == 0 (ptr=%p ld=%llu)", (const void*)ptr,
         (unsigned long long)ld_floats);
  cuuint64_t dims[3] = {32, k, (mn + 31) / 32};
  cuuint64_t strides[2] = {ld_floats * 4, 128};
  cuuint32_t box[3] = {32, 32, chunks_per_box};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(3d MN-major) failed with %d", (int)r);
  return 0;
}

int make_map_out(wgan_handle* h, CUtensorMap* m, float* ptr, uint64_t N, uint64_t M, uint64_t S,
                 uint64_t ld_floats, uint64_t slab_floats) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld_floats & 3) != 0 || (slab_floats & 3) != 0)
    FAIL("TMA output must be 16-byte aligned with ld %% 4 == 0 (ptr=%p ld=%llu)", (void*)ptr,
         (unsigned long long)ld_floats);
  cuuint64_t dims[3] = {N, M, S};
  cuuint64_t strides[2] = {ld_floats * 4, (S > 1 ? slab_floats : ld_floats * M) * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(3d out) failed with %d", (int)r);
  return 0;
}

// N tile: minimise the tensor-pipe time of one M tile, n_tiles * (43 + bn/2) cycles per UMMA_K step
// (measured on B200: a cta_group::1 M=128 tcgen05.mma costs ~43 + N/2 cycles, scripts/ubench/mma_rate.cu);
// ties go to the wider tile (fewer re-reads of A).
int pick_bn(int N) {
  int best = 256;
  long best_cost = 1L << 60;
  for (int bn = 256; bn >= 32; bn -= 32) {
    const long cost = static_cast<long>(cdiv(N, bn)) * (86 + bn);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// N tile of the CTA-pair kernel: cta_group::2 runs at N/2 cycles per UMMA_K step with no fixed cost, so
// only padding and A re-reads matter; each CTA loads bn/2 columns, which must be whole 32-column boxes
// for an MN-major B (bn % 64 == 0).
int pick_bn_2cta(int N, bool b_mn, bool long_k = false) {
  const int step = b_mn ? 64 : 32;
  int best = 256;
  long best_cost = 1L << 60;
  for (int bn = 256; bn >= step; bn -= step) {
    // long reductions stream their operands: per K block a CTA ingests 16 KB of A + bn/2 * 128 B of B at ~51 B/clk
    // (scripts/ubench/tma_rate.cu) against 2 * bn MMA cycles, so narrow tiles that re-stream A more often lose
    // (N = 800: 4 x 256 instead of 5 x 192, BASELINE config 4)
    const long per_kb = long_k ? std::max<long>(2 * bn, (16384 + 64 * bn) * 10 / 512) : bn + 48;
    const long cost = static_cast<long>(cdiv(N, bn)) * per_kb;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

size_t gemm_smem_bytes(int bn, int msub, int* stages_out, bool side = false) {
  const size_t budget = 232448;                       // 227 KB opt-in maximum per CTA
  const size_t fixed = 1024 + (side ? 2 : 1) * EPI_BUFS * EPI_BOX_BYTES + 512;   // out ring (+ side ring) + barriers
  const size_t stage = static_cast<size_t>(BM) * msub * 128 + static_cast<size_t>(bn) * 128;
  int stages = static_cast<int>((budget - fixed) / stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  static const int cap = getenv("WGAN_B200_GEMM_STAGES") ? atoi(getenv("WGAN_B200_GEMM_STAGES")) : 0;   // experiments
  if (cap > 0 && stages > cap) stages = cap;
  *stages_out = stages;
  size_t need = fixed + stages * stage;
  if (need < 120 * 1024) need = 120 * 1024;           // force 1 CTA/SM: each CTA allocates all of TMEM
  return need;
}

int launch_finish(wgan_handle* h, const float* part, size_t slab, int splits, float* D, int ldd, int M, int N,
                  const EpiParams& epi, cudaStream_t s) {
  if (epi.row_sumsq == nullptr) {
    const size_t n = static_cast<size_t>(M) * ((N + 3) / 4);
    if (splits >= 8 && n <= static_cast<size_t>(h->num_sms) * 32 * 16) {
      KLAUNCH(finish_splitk_tall_kernel, static_cast<int>((n + 31) / 32), 256, 0, s, part, slab, splits, D, ldd, M, N, epi);
      return post_launch(h, "finish_splitk_tall_kernel");
    }
    KLAUNCH(finish_splitk_vec4_kernel, ew_grid(n, 256, h->num_sms), 256, 0, s, part, slab, splits, D, ldd, M, N, epi);
    return post_launch(h, "finish_splitk_vec4_kernel");
  }
  KLAUNCH(finish_splitk_kernel, ew_grid(static_cast<size_t>(M) * 32, 256, h->num_sms), 256, 0, s, part, slab, splits, D,
                                                                                            ldd, M, N, epi);
  return post_launch(h, "finish_splitk_kernel");
}

// Launches min(cluster_tiles, co-resident clusters) clusters of p.cluster persistent CTAs.
template <bool A_MN, bool B_MN>
int launch_tc(wgan_handle* h, const GemmParams& p, size_t smem, int cluster_tiles, cudaStream_t s) {
  // attribute / occupancy state is per handle (a handle is bound to one device and one host thread)
  constexpr int inst = (A_MN ? 2 : 0) + (B_MN ? 1 : 0);
  int* max_clusters = h->max_clusters[inst];
  if (!h->attr_tc[inst]) {
    CK(cudaFuncSetAttribute(gemm_tf32_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            232448));
    h->attr_tc[inst] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (max_clusters[p.cluster] == 0) {
    int n = 0;
    cfg.gridDim = dim3(h->num_sms / p.cluster * p.cluster);
    if (p.cluster == 1 || cudaOccupancyMaxActiveClusters(&n, gemm_tf32_kernel<A_MN, B_MN>, &cfg) != cudaSuccess ||
        n <= 0)
      n = h->num_sms / p.cluster;
    (void)cudaGetLastError();
    max_clusters[p.cluster] = n;
  }
  int clusters = cluster_tiles < max_clusters[p.cluster] ? cluster_tiles : max_clusters[p.cluster];
  if (clusters > h->sm_cap / p.cluster) clusters = h->sm_cap / p.cluster;
  cfg.gridDim = dim3(clusters * p.cluster);
  CK(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<A_MN, B_MN>, p));
  return post_launch(h, "gemm_tf32_kernel");
}

// CTA-pair kernel: cluster (2,1,1), at most num_sms/2 co-resident pairs.
template <bool A_MN, bool B_MN>
int launch_tc2(wgan_handle* h, const GemmParams& p, size_t smem, int pair_tiles, cudaStream_t s) {
  constexpr int inst = (A_MN ? 2 : 0) + (B_MN ? 1 : 0);
  if (!h->attr_tc2[inst]) {
    CK(cudaFuncSetAttribute(gemm_tf32_2cta_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            232448));
    h->attr_tc2[inst] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  const int max_pairs = h->sm_cap / 2;
  const int pairs = pair_tiles < max_pairs ? pair_tiles : max_pairs;
  cfg.gridDim = dim3(pairs * 2);
  CK(cudaLaunchKernelEx(&cfg, gemm_tf32_2cta_kernel<A_MN, B_MN>, p));
  return post_launch(h, "gemm_tf32_2cta_kernel");
}

// D[M,N] = epi( sum_parts A_i * B_i ).  a_mn/b_mn: operand majors (see gemm_sm100.cuh).
int run_gemm(wgan_handle* h, bool a_mn, bool b_mn, const Part* parts, int nparts, float* D, int ldd,
             int M, int N, const EpiParams& epi, int* sumsq_tiles, int force_splits, cudaStream_t s,
             float* blend_out = nullptr, int ws_slab0 = 0, int* raw_slabs = nullptr, bool raw_force_ws = false) {
  // raw_slabs != nullptr: leave the raw products for the caller's own finish pass (no epilogue, no finish
  // kernel).  They go to slabs [ws_slab0, ws_slab0 + *raw_slabs) of the split workspace, or -- single
  // unsplit product and !raw_force_ws -- straight into D (*raw_slabs = 0).
  const bool raw = raw_slabs != nullptr;
  if (M <= 0 || N <= 0) FAIL("gemm: empty output %dx%d", M, N);
  if (epi.blend_x != nullptr) force_splits = 1;       // the second output only exists in the fused epilogue
  for (int i = 0; i < nparts; ++i)
    if (parts[i].K <= 0) FAIL("gemm: empty K");
  const size_t slab = round64(static_cast<size_t>(M) * ldd);

  if (h->cfg.gemm_mode == 1) {
    // FP32 CUDA-core verification path: raw products into slabs, then the shared finish kernel.
    // raw products go through the workspace so that an epilogue operand aliasing D (in-place
    // slope mask) is still intact when the finish kernel reads it
    const bool raw_direct = raw && nparts == 1 && !raw_force_ws;
    const bool ws_fits = !raw_direct && slab * (ws_slab0 + nparts) <= h->splitws_n;
    if (!ws_fits && !raw_direct && (nparts > 1 || epi.aux == D || raw)) FAIL("gemm: split workspace too small");
    float* raw_out = ws_fits ? h->splitws + slab * ws_slab0 : D;
    for (int i = 0; i < nparts; ++i) {
      const Part& pt = parts[i];
      dim3 grid(cdiv(N, 64), cdiv(M, 64));
      KLAUNCH(gemm_fp32_kernel, grid, 256, 0, s, pt.A, a_mn ? 1 : pt.lda, a_mn ? pt.lda : 1, pt.B,
                                            b_mn ? 1 : pt.ldb, b_mn ? pt.ldb : 1,
                                            raw_out + slab * i, ldd, M, N, pt.K);
      TRY(post_launch(h, "gemm_fp32_kernel"));
    }
    if (raw) { *raw_slabs = raw_direct ? 0 : nparts; return 0; }
    TRY(launch_finish(h, raw_out, slab, nparts, D, ldd, M, N, epi, s));
    if (epi.blend_x != nullptr) {
      KLAUNCH(blend_kernel, ew_grid(static_cast<size_t>(M) * (ldd / 4), 256, h->num_sms), 256, 0, s, 
          epi.blend_x, epi.ld_blend, D, ldd, epi.blend_eps, blend_out, ldd, M, ldd / 4);
      TRY(post_launch(h, "blend_kernel"));
    }
    if (sumsq_tiles) *sumsq_tiles = 1;
    return 0;
  }

  // ---- tile configuration ----
  // two_cta: CTA pairs (tcgen05 cta_group::2) compute 256 x bn tiles, each CTA holding half of the B tile
  // (full TF32 MMA rate, 6 smem stages).  Otherwise one CTA per 128 x bn tile (optionally 2 sub-tiles).
  // The pair kernel is the default; a 128-row tiling is kept for outputs whose row count would be
  // padded >10 % more by 256-row tiles (e.g. M = 600: 768 vs 640 rows).
  const double waste2 = static_cast<double>(cdiv(M, 2 * BM)) * 2 * BM / M;
  const double waste1 = static_cast<double>(cdiv(M, BM)) * BM / M;
  const bool two_cta = h->gemm_2cta != 0 && waste2 <= 1.10 * waste1;
  static const int bn_cap = getenv("WGAN_B200_GEMM_BN") ? atoi(getenv("WGAN_B200_GEMM_BN")) : 0;          // experiments
  int kmax = 0;
  for (int i = 0; i < nparts; ++i) kmax = parts[i].K > kmax ? parts[i].K : kmax;
  static const bool longk_bn = getenv("WGAN_B200_NO_LONGK_BN") == nullptr;      // A/B switch
  // (row-major activations as A only: the weight gradient of the same wide layer measured 651 us with 5 x 192
  // tiles and 741 us with 4 x 256, the forward product 999 us and 739 us)
  int bn = two_cta ? pick_bn_2cta(N, b_mn, longk_bn && kmax >= 4096 && !a_mn) : pick_bn(N);
  if (!two_cta && bn_cap >= 32 && bn_cap % 32 == 0 && bn > bn_cap) bn = bn_cap;
  // a nearly empty second wave (e.g. [4096 x 600] x [600 x 600]: 5 x 16 = 80 tiles of 128 columns on 74 pairs): widen
  // the tile to the smallest width whose tiles fit one wave (192 -> 64 tiles)
  static const bool one_wave = getenv("WGAN_B200_NO_ONE_WAVE") == nullptr;
  if (one_wave && two_cta && nparts == 1) {
    const int pairs = h->sm_cap / 2, mt2 = cdiv(M, 2 * BM), step = b_mn ? 64 : 32;
    const int t0 = mt2 * cdiv(N, bn);
    if (t0 > pairs && 2 * t0 < 3 * pairs)
      for (int w = bn + step; w <= 256; w += step)
        if (mt2 * cdiv(N, w) <= pairs) { bn = w; break; }
  }
  // latency-bound outputs (the pair tiles would cover less than half of the SMs): halve the tile width, which
  // halves both the MMA time per K block and the epilogue of the single tile each CTA pair owns
  static const bool narrow_small = getenv("WGAN_B200_NO_NARROW_SMALL") == nullptr;
  // (not for long reductions over many rows: a second N tile would stream the A operand twice)
  if (narrow_small && two_cta && N > 128 && N <= bn && (kmax < 1024 || M <= 2 * BM) &&
      cdiv(M, 2 * BM) * cdiv(N, bn) * nparts * 2 <= h->num_sms / 2) {
    const int step = b_mn ? 64 : 32;
    bn = cdiv(cdiv(N, 2), step) * step;
  }
  const int n_tiles = cdiv(N, bn);
  const int ctas_per_tile = two_cta ? 2 : 1;
  // the epilogue's [M, N] side operand (slope-mask source / blend x) is staged by TMA when it is TMA-legal
  const float* side_ptr = epi.aux != nullptr ? epi.aux : epi.blend_x;
  const int side_ld = epi.aux != nullptr ? epi.ld_aux : epi.ld_blend;
  static const bool no_side = getenv("WGAN_B200_NO_SIDE_TMA") != nullptr;
  bool side = !no_side && side_ptr != nullptr && !(epi.aux != nullptr && epi.blend_x != nullptr) &&
              (reinterpret_cast<uintptr_t>(side_ptr) & 15) == 0 && (side_ld & 3) == 0;
  static const bool no_sk = getenv("WGAN_B200_NO_STREAMK") != nullptr;
  // pairs per tile up to which stream-K replaces split-K + finish.  4 = the 12288-row layer-1 products (48 / 26 tiles on
  // 74 pairs); 6 would also take the 4096-row ones (16 tiles) and was measured SLOWER on the whole iteration
  // (379.9 vs 389.0 iters/s, explicit form: the owner's serial fix-up over 4-5 partials costs more than the finish pass)
  static const int sk_max_contrib = getenv("WGAN_B200_SK_MAX_CONTRIB") ? atoi(getenv("WGAN_B200_SK_MAX_CONTRIB")) : 4;
  const int sk_pairs = h->sm_cap / 2;
  // (512-row pair tiles -- two 128-row sub-tiles per CTA sharing the B tile, to halve the L2 -> SM re-reads of the
  // L2-resident operand of the two 6605-deep layer-1 products -- and burst issue in the producer were built, parity-
  // tested and measured in round 2: TMA loads fell from 651 to 488 MB as predicted but the products got SLOWER
  // (77 -> 92 us, 78 -> 104 us: four 48 KB stages cover less latency and the stream-K fix-up tail doubles), bursts
  // changed nothing, and the extra loops in the single-thread issue paths cost 3 % on every GEMM.  Removed again;
  // DESIGN.md section 4.1 has the numbers, scripts/ubench/l1_stream.cu the load-pattern study.)
  int msub = 1;
  if (!two_cta && h->gemm_msub == 2 && M >= 2 * BM) msub = 2;   // single-CTA kernel, opt-in: measured slower on every WGAN shape
  int nacc, m_tiles, tiles, cluster, m_groups, stages;
  size_t smem;
  int splits[2], kbps[2], total_splits;
  bool streamk;
  for (;;) {
    nacc = (2 * msub * bn <= 512) ? 2 : 1;
    const int tile_rows = two_cta ? 2 * BM : BM * msub;
    m_tiles = cdiv(M, tile_rows);
    tiles = m_tiles * n_tiles;
    // (1-CTA only) CTAs of a cluster take adjacent M tiles and share the B tile through TMA multicast
    cluster = two_cta ? 1 : h->gemm_cluster;
    while (cluster > 1 && (m_tiles < cluster || (bn / 32) % cluster != 0)) cluster >>= 1;
    m_groups = cdiv(m_tiles, cluster);
    smem = two_cta ? gemm_smem_bytes(bn / 2, 1, &stages, side) : gemm_smem_bytes(bn, msub, &stages, side);

    // split-K policy: fill ~2 waves when the output has few tiles, at least 8 k-blocks per split
    total_splits = 0;
    for (int i = 0; i < nparts; ++i) {
      const int total_kb = cdiv(parts[i].K, BK);
      int sp = 1;
      if (force_splits < 0) {
        sp = 1;
      } else if (force_splits > 0) {
        sp = force_splits;
      } else {
        // smallest split count whose work units fill >= 85 % of the last wave of persistent CTAs (pairs),
        // with at least 8 K blocks per split; otherwise the best-filling one
        const int slots = h->num_sms / ctas_per_tile;
        const int units = tiles * nparts;
        // short reductions (K < 1024) are not split: the finish pass would cost more than the idle SMs
        const int max_sp = total_kb < 32 ? 1 : std::max(1, std::min(total_kb / 8, 64));
        double best_eff = 0.0;
        for (int c = 1; c <= max_sp; ++c) {
          const double eff = static_cast<double>(units) * c / (static_cast<double>(cdiv(units * c, slots)) * slots);
          if (eff > best_eff + 1e-9) { best_eff = eff; sp = c; }
          if (eff >= 0.85) { sp = c; break; }
        }
      }
      if (sp > total_kb) sp = total_kb;
      if (sp < 1) sp = 1;
      int per = cdiv(total_kb, sp);
      sp = cdiv(total_kb, per);                         // no empty split
      splits[i] = sp;
      kbps[i] = per;
      total_splits += sp;
    }
    // Stream-K (CTA-pair kernel): where split-K would be chosen for a single fused product whose tiles are at most
    // ~4 pairs apiece, every pair instead streams one contiguous range of (tile, K block)
    // units; partial accumulators meet in the workspace and the tile's owner runs the fused epilogue -- no slabs, no
    // finish pass, no idle pairs.  force_splits < 0 forces it (tests), > 0 forces split-K.
    streamk = false;
    if (!no_sk && two_cta && nparts == 1 && !raw && cdiv(parts[0].K, BK) >= 2 &&
        ((force_splits == 0 && splits[0] > 1 && tiles * sk_max_contrib >= sk_pairs) || force_splits < 0) &&
        static_cast<long long>(tiles) * cdiv(parts[0].K, BK) >= sk_pairs &&
        static_cast<size_t>(sk_pairs) * 2 * SK_SLOT_FLOATS <= h->splitws_n) {
      streamk = true;
      splits[0] = 1;
      kbps[0] = cdiv(parts[0].K, BK);
      total_splits = 1;
    }
    break;
  }
  if (total_splits > 1 && slab * (ws_slab0 + total_splits) > h->splitws_n) {
    if (nparts > 1) FAIL("gemm: split workspace too small for %d slabs of %zu floats", total_splits, slab);
    splits[0] = 1;
    kbps[0] = cdiv(parts[0].K, BK);
    total_splits = 1;
  }
  const bool to_ws = total_splits > 1 || (raw && raw_force_ws);
  if (to_ws && slab * (ws_slab0 + total_splits) > h->splitws_n) FAIL("gemm: split workspace too small");
  const bool direct = !to_ws;
  float* out = direct ? D : h->splitws + slab * ws_slab0;
  if (!direct && side) {                               // split partials skip the fused epilogue: no side ring
    side = false;
    smem = two_cta ? gemm_smem_bytes(bn / 2, 1, &stages, false) : gemm_smem_bytes(bn, msub, &stages, false);
  }

  int slab0 = 0;
  for (int i = 0; i < nparts; ++i) {
    const Part& pt = parts[i];
    GemmParams p;
    memset(&p, 0, sizeof p);
    if (!a_mn) TRY(make_map_2d(h, &p.tma_a, pt.A, pt.K, M, pt.lda, 32, two_cta ? BM : BM * msub, false));
    else       TRY(make_map_mn3d(h, &p.tma_a, pt.A, M, pt.K, pt.lda, two_cta ? BM / 32 : BM * msub / 32));
    if (!b_mn) TRY(make_map_2d(h, &p.tma_b, pt.B, pt.K, N, pt.ldb, 32, two_cta ? bn / 2 : bn / cluster, false));
    else       TRY(make_map_mn3d(h, &p.tma_b, pt.B, N, pt.K, pt.ldb, two_cta ? bn / 64 : bn / 32 / cluster));
    TRY(make_map_out(h, &p.tma_d, out + (direct ? 0 : slab * slab0), N, M, splits[i], ldd, slab));
    if (direct && epi.blend_x != nullptr)
      TRY(make_map_out(h, &p.tma_d2, blend_out, N, M, 1, ldd, slab));
    if (side) {
      cuuint64_t sdims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M)};
      cuuint64_t sstr[1] = {static_cast<cuuint64_t>(side_ld) * 4};
      cuuint32_t sbox[2] = {32, 32};
      cuuint32_t sone[2] = {1, 1};
      CUresult r = h->encode(&p.tma_side, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(side_ptr), sdims, sstr,
                             sbox, sone, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(side operand) failed with %d", (int)r);
    }
    p.side_mode = side ? 1 : 0;
    p.M = M; p.N = N; p.K = pt.K;
    p.bn = bn;
    p.splits = direct ? 1 : splits[i];
    p.kb_per_split = kbps[i];
    p.m_tiles = m_tiles; p.n_tiles = n_tiles;
    p.stages = stages;
    p.total_kb = cdiv(pt.K, BK);
    p.cluster = cluster;
    p.m_groups = m_groups;
    p.msub = msub;
    p.nacc = nacc;
    // operands that stream from HBM are prefetched into L2 a few K blocks ahead of the smem ring
    static const int l2_prefetch_mb = getenv("WGAN_B200_L2_PREFETCH") ? atoi(getenv("WGAN_B200_L2_PREFETCH")) : 0;
    const size_t pf_min = static_cast<size_t>(l2_prefetch_mb) << 20;      // operands at least this large are prefetched
    p.prefetch_a = (l2_prefetch_mb > 0 && static_cast<size_t>(M) * pt.K * 4 >= pf_min) ? 1 : 0;
    p.prefetch_b = (l2_prefetch_mb > 0 && static_cast<size_t>(N) * pt.K * 4 >= pf_min) ? 1 : 0;
    // partials that go to the workspace stay raw (p.epi is all-zero) even when this part has a
    // single split; finish_splitk_kernel applies the epilogue after summing the slabs
    if (direct && !raw) p.epi = epi;
    else if (!direct) p.splits = splits[i];
    if (streamk) {
      p.streamk = 1;
      p.sk_units = static_cast<long long>(tiles) * p.total_kb;
      p.sk_ws = h->splitws;
      p.sk_flags = h->sk_flags;
      {
        // Unit ranges.  Phase stamps of the layer-1 products (scripts/diag_stamps.py): every pair finishes its MMAs
        // within ~4 us of the others, a partial's dump + flag takes ~4 us, and a tile's owner then needs 6-10 us for
        // the fix-up; a pair whose LAST segment is a partial therefore gets `shift` units less (its dump overlaps the
        // others' last K blocks) and the rest is spread over all pairs.
        static const int shift_env = getenv("WGAN_B200_SK_SHIFT") ? atoi(getenv("WGAN_B200_SK_SHIFT")) : 9;
        const int P = sk_pairs, KB = p.total_kb;
        const long long U = p.sk_units;
        const int shift = static_cast<int>(std::min<long long>(shift_env, U / P / 8));
        long long b[80];
        for (int q = 0; q <= P; ++q) b[q] = static_cast<long long>(q) * U / P;
        for (int iter = 0; iter < 3 && shift > 0; ++iter) {
          bool late[80];
          int nlate = 0;
          for (int q = 0; q < P; ++q) {
            const long long te = (b[q + 1] - 1) / KB;                 // tile of the pair's last unit
            late[q] = b[q + 1] > b[q] && std::max(b[q], te * KB) > te * KB;   // its last segment starts inside that tile
            nlate += late[q] ? 1 : 0;
          }
          const double base = (static_cast<double>(U) + static_cast<double>(shift) * nlate) / P;
          double acc = 0.0;
          for (int q = 0; q < P; ++q) {
            acc += base - (late[q] ? shift : 0);
            b[q + 1] = std::min<long long>(U, std::max<long long>(b[q] + 1, static_cast<long long>(acc + 0.5)));
          }
          b[P] = U;
        }
        for (int q = 0; q <= P; ++q) p.sk_b[q] = static_cast<int>(b[q]);
      }
    }
    const int grid = streamk ? sk_pairs : m_groups * n_tiles * splits[i];   // cluster tiles; launch_tc caps at residency
    if (two_cta) {
      if (!a_mn && b_mn) TRY((launch_tc2<false, true>(h, p, smem, grid, s)));
      else if (!a_mn && !b_mn) TRY((launch_tc2<false, false>(h, p, smem, grid, s)));
      else if (a_mn && b_mn) TRY((launch_tc2<true, true>(h, p, smem, grid, s)));
      else FAIL("gemm: A MN-major with B K-major is not instantiated");
    } else if (!a_mn && b_mn) TRY((launch_tc<false, true>(h, p, smem, grid, s)));
    else if (!a_mn && !b_mn) TRY((launch_tc<false, false>(h, p, smem, grid, s)));
    else if (a_mn && b_mn) TRY((launch_tc<true, true>(h, p, smem, grid, s)));
    else FAIL("gemm: A MN-major with B K-major is not instantiated");
    slab0 += splits[i];
  }
  if (raw) {
    *raw_slabs = direct ? 0 : total_splits;
    return 0;
  }
  if (!direct) {
    TRY(launch_finish(h, h->splitws, slab, total_splits, D, ldd, M, N, epi, s));
    if (sumsq_tiles) *sumsq_tiles = 1;
  } else if (sumsq_tiles) {
    *sumsq_tiles = n_tiles * (NUM_EPI_WARPS / 4);
  }
  return 0;
}

inline EpiParams epi_none() {
  EpiParams e;
  memset(&e, 0, sizeof e);
  return e;
}
inline EpiParams epi_bias_lrelu(const float* bias, float alpha) {
  EpiParams e = epi_none();
  e.bias = bias; e.act = 1; e.alpha = alpha;
  return e;
}
inline EpiParams epi_mask(const float* aux, int ld_aux, float alpha) {
  EpiParams e = epi_none();
  e.aux = aux; e.ld_aux = ld_aux; e.alpha = alpha;
  return e;
}

int gemm1(wgan_handle* h, bool a_mn, bool b_mn, const float* A, int lda, const float* B, int ldb,
          float* D, int ldd, int M, int N, int K, const EpiParams& e, cudaStream_t s,
          int* sumsq_tiles = nullptr) {
  Part p{A, lda, B, ldb, K};
  return run_gemm(h, a_mn, b_mn, &p, 1, D, ldd, M, N, e, sumsq_tiles, 0, s);
}

// ---------------------------------------------------------------------------------------------
// Row-local fused chains (mid_chain_sm100.cuh).  Available on the tensor-core path when the critic's hidden
// width fits one N tile (H <= 256, H % 8 == 0); otherwise the callers use the separate GEMMs + kernels.
// ---------------------------------------------------------------------------------------------
bool mid_chain_available(const wgan_handle* h) {
  static const bool off = getenv("WGAN_B200_NO_MID_FUSE") != nullptr;
  return !off && h->cfg.gemm_mode == 0 && h->Hd <= 256 && (h->Hd % 8) == 0;
}

// `x` carries the mode's scalar / pointer fields; this fills the tensor maps and the geometry.
int run_mid_chain(wgan_handle* h, int mode, int rows, const float* A, const float* B1, int ldb1, const float* B2,
                  int ldb2, float* o1, float* o2, float* o3, MidParams x, cudaStream_t s) {
  const int H = h->Hd, ld = h->Hdl;
  const int bn = cdiv(H, 32) * 32, KB = cdiv(H, 32);
  const size_t b_bytes = static_cast<size_t>(bn) * 128, a_bytes = static_cast<size_t>(KB) * 16384;
  const size_t fixed = 1024 + 2048 + 2048 + 256;       // alignment slack, b2 | w3 vectors, row partials, barriers
  int stages = static_cast<int>((232448 - fixed - a_bytes) / b_bytes);
  if (stages > MID_MAX_STAGES) stages = MID_MAX_STAGES;
  if (stages < 2) FAIL("mid chain: hidden width %d does not fit", H);
  const size_t smem = fixed + a_bytes + stages * b_bytes;
  TRY(make_map_2d(h, &x.tma_a, A, H, rows, ld, 32, BM, false));
  TRY(make_map_mn3d(h, &x.tma_b1, B1, H, H, ldb1, bn / 32));
  if (mode == 0) TRY(make_map_2d(h, &x.tma_b2, B2, H, H, ldb2, 32, bn, false));
  else           TRY(make_map_mn3d(h, &x.tma_b2, B2, H, H, ldb2, bn / 32));
  const size_t slab = round64(static_cast<size_t>(rows) * ld);
  x.o1 = o1;
  TRY(make_map_out(h, &x.tma_o2, o2, H, rows, 1, ld, slab));
  TRY(make_map_out(h, &x.tma_o3, o3, H, rows, 1, ld, slab));
  x.rows = rows; x.H = H; x.bn = bn; x.kblocks = KB; x.stages = stages;
  x.alpha = h->cfg.alpha;
  x.ld = ld;
  if (!h->attr_mid[mode]) {
    if (mode == 0) CK(cudaFuncSetAttribute(mid_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    else           CK(cudaFuncSetAttribute(mid_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    h->attr_mid[mode] = true;
  }
  if (mode == 0) KLAUNCH(mid_chain_kernel<0>, cdiv(rows, BM), MID_THREADS, smem, s, x);
  else           KLAUNCH(mid_chain_kernel<1>, cdiv(rows, BM), MID_THREADS, smem, s, x);
  return post_launch(h, "mid_chain_kernel");
}

// Critic layer 2 forward, layer 3 + upstream deltas, layer-2 dgrad on `rows` stacked rows whose coefficient
// segments are `seg` rows long  [W:113-123,137]: fused chain when available, else GEMM + head kernel + GEMM.
int critic_middle(wgan_handle* h, int rows, int seg, float c0, float c1, float c2, cudaStream_t s) {
  const int Hd = h->Hd, Hdl = h->Hdl;
  const float a = h->cfg.alpha;
  const float* W2 = h->P(T_W2); const int ldw2 = h->td[T_W2].ld;
  if (mid_chain_available(h)) {
    MidParams x;
    memset(&x, 0, sizeof x);
    x.b2 = h->P(T_B2); x.w3 = h->P(T_W3); x.b3 = h->P(T_B3); x.dout = h->dout;
    x.seg = seg; x.c0 = c0; x.c1 = c1; x.c2 = c2;
    x.bits1 = h->bits1; x.bits2 = h->bits2;
    return run_mid_chain(h, 0, rows, h->h1, W2, ldw2, W2, ldw2, h->h2, h->d2, h->d1, x, s);
  }
  TRY(gemm1(h, false, true, h->h1, Hdl, W2, ldw2, h->h2, Hdl, rows, Hd, Hd, epi_bias_lrelu(h->P(T_B2), a), s));
  KLAUNCH(critic_head_kernel, ew_grid(static_cast<size_t>(rows) * 32, 256, h->num_sms), 256, 0, s,
      h->h2, Hdl, rows, Hd, h->P(T_W3), h->P(T_B3), h->dout, h->d2, seg, c0, c1, c2, a);
  TRY(post_launch(h, "critic_head_kernel"));
  // delta1 / g1 = (delta2 W2^T) . slope(h1)
  return gemm1(h, false, false, h->d2, Hdl, W2, ldw2, h->d1, Hdl, rows, Hd, Hd, epi_mask(h->h1, Hdl, a), s);
}

// Up to three column-sum jobs in one launch pair: out_j[n] = sum_seg scale * sum_r X_seg[r, n].  `zero_out`
// (optional) is one extra float to clear; `sums` (optional) are loss-scalar sums evaluated by extra blocks of the
// first launch.
int colsum_jobs(wgan_handle* h, const ColsumJob* jobs, int njobs, cudaStream_t s, float* zero_out = nullptr,
                const SumJob* sums = nullptr, int nsums = 0) {
  ColsumArgs a;
  memset(&a, 0, sizeof a);
  int maxrows = 1, maxN = 1;
  for (int j = 0; j < njobs; ++j) {
    a.job[j] = jobs[j];
    if (jobs[j].N > maxN) maxN = jobs[j].N;
    for (int i = 0; i < jobs[j].nseg; ++i)
      if (jobs[j].seg[i].rows > maxrows) maxrows = jobs[j].seg[i].rows;
  }
  int RC = cdiv(maxrows, 64);
  if (RC > 64) RC = 64;
  if (RC < 1) RC = 1;
  a.part_stride = h->colpart_stride;
  a.njobs = njobs;
  a.RC = RC;
  a.zero_out = zero_out;
  for (int j = 0; j < 3; ++j) a.block0[j + 1] = a.block0[j] + (j < njobs ? cdiv(jobs[j].N, 128) * RC : 0);
  SumJobs sj;
  memset(&sj, 0, sizeof sj);
  for (int j = 0; j < nsums; ++j) sj.job[j] = sums[j];
  const int grid = a.block0[njobs] + nsums;
  KLAUNCH(colsum_partial_kernel, grid, 256, 0, s, a, h->colpart, sj, nsums);
  TRY(post_launch(h, "colsum_partial_kernel"));
  dim3 g2(cdiv(maxN, 256), njobs);
  KLAUNCH(colsum_final_kernel, g2, 256, 0, s, a, h->colpart, RC);
  return post_launch(h, "colsum_final_kernel");
}

// Bring a caller matrix into a TMA-legal layout (16 B aligned, ld % 4 == 0); copies only if needed.
int stage_input(wgan_handle* h, const float* src, int ld, int rows, int cols, float* staging,
                int ld_staging, const float** out_ptr, int* out_ld, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (ld & 3) == 0) {
    *out_ptr = src; *out_ld = ld;
    return 0;
  }
  CK(cudaMemcpy2DAsync(staging, static_cast<size_t>(ld_staging) * 4, src, static_cast<size_t>(ld) * 4,
                       static_cast<size_t>(cols) * 4, rows, cudaMemcpyDeviceToDevice, s));
  *out_ptr = staging; *out_ld = ld_staging;
  return 0;
}

// Generator forward [W:61-83]: gh1, gh2 and x~ (into `xt`, ld `ldxt`).
int generator_forward(wgan_handle* h, const float* z, int ldz, int rows, float* xt, int ldxt,
                      cudaStream_t s, const float* blend_x = nullptr, int ld_blend = 0,
                      const float* blend_eps = nullptr, float* blend_out = nullptr) {
  const float a = h->cfg.alpha;
  TRY(gemm1(h, false, true, z, ldz, h->P(T_WG1), h->td[T_WG1].ld, h->gh1, h->Hgl, rows, h->Hg, h->Z,
            epi_bias_lrelu(h->P(T_BG1), a), s));
  TRY(gemm1(h, false, true, h->gh1, h->Hgl, h->P(T_WG2), h->td[T_WG2].ld, h->gh2, h->Hgl, rows, h->Hg,
            h->Hg, epi_bias_lrelu(h->P(T_BG2), a), s));
  EpiParams e3 = epi_bias_lrelu(h->P(T_BG3), a);
  e3.blend_x = blend_x; e3.ld_blend = ld_blend; e3.blend_eps = blend_eps;
  Part p3{h->gh2, h->Hgl, h->P(T_WG3), h->td[T_WG3].ld, h->Hg};
  return run_gemm(h, false, true, &p3, 1, xt, ldxt, rows, h->G, e3, nullptr, 0, s, blend_out);
}

// Every entry point that writes the shared activation workspace calls this first: outstanding prepare tokens die.
inline void invalidate_workspace(wgan_handle* h) { ++h->ws_gen; }

int check_rows(wgan_handle* h, int rows, int global_batch) {
  if (rows <= 0 || rows > h->R) FAIL("rows=%d outside (0, max_batch=%d]", rows, h->R);
  if (global_batch < rows) FAIL("global_batch=%d < rows=%d", global_batch, rows);
  return 0;
}

int adam_apply(wgan_handle* h, int net, cudaStream_t s) {
  const size_t n4 = (h->arena_n[net] - 4) / 4;     // scalar tail excluded
  KLAUNCH(adam_kernel, ew_grid(n4, 256, h->num_sms), 256, 0, s, 
      h->arena[net][WGAN_KIND_PARAM], h->arena[net][WGAN_KIND_GRAD], h->arena[net][WGAN_KIND_M],
      h->arena[net][WGAN_KIND_V], n4, h->powers[net], h->cfg.learn_rate, h->cfg.beta1, h->cfg.beta2,
      h->cfg.epsilon);
  return post_launch(h, "adam_kernel");
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int wgan_padded_ld(int cols) { return cols == 1 ? 1 : round4(cols); }
int wgan_tensor_count(void) { return 12; }

const char* wgan_last_error(const wgan_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int64_t wgan_launch_count(const wgan_handle* h) {
  return (h ? h->launches : 0) + static_cast<int64_t>(g_pipeline_launches.load());
}

void wgan_destroy(wgan_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (int r = 0; r < WGAN_DP_MAX_WORLD; ++r)
    if (h->dp_peer_ipc[r] && h->dp_peer[r] != nullptr) cudaIpcCloseMemHandle(h->dp_peer[r]);
  for (int n = 0; n < 2; ++n) {
    for (int k = 0; k < 4; ++k)
      if (k != WGAN_KIND_GRAD) cudaFree(h->arena[n][k]);      // the gradient arenas live in the DP window
    cudaFree(h->powers[n]);
  }
  if (h->dp_stream) cudaStreamDestroy(h->dp_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->aux_fork) cudaEventDestroy(h->aux_fork);
  if (h->aux_join) cudaEventDestroy(h->aux_join);
  if (h->dp_fork) cudaEventDestroy(h->dp_fork);
  if (h->dp_join) cudaEventDestroy(h->dp_join);
  cudaFree(h->dp_win);
  cudaFree(h->sk_flags);
  float* bufs[] = {h->zbuf, h->gh1, h->gh2, h->xcat, h->h1, h->h2, h->d2, h->d1, h->dout, h->v2, h->gd2,
                   h->gd1, h->dz, h->nsq, h->crow, h->pen, h->norms, h->colpart, h->splitws, h->sample_tmp,
                   h->gram_k, h->gram_s, h->gram_t};
  for (float* b : bufs) cudaFree(b);
  cudaFree(h->bits1); cudaFree(h->bits2);
  delete h;
}

int wgan_create(const wgan_config* cfg, wgan_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return -2; }
  *out = nullptr;
  wgan_handle* h = new wgan_handle();
  h->cfg = *cfg;
  h->launches = 0;
  h->last_gx_rows = 0;
  h->gemm_cluster = 1;
  h->ws_gen = 0;
  h->prep_token = h->gprep_token = 0; h->prep_rows = h->gprep_rows = 0;
  h->sk_flags = nullptr;
  memset(h->attr_tc, 0, sizeof h->attr_tc); memset(h->attr_tc2, 0, sizeof h->attr_tc2);
  memset(h->attr_mid, 0, sizeof h->attr_mid); memset(h->max_clusters, 0, sizeof h->max_clusters);
  h->dp_win = nullptr; h->dp_win_bytes = 0; h->dp_connected = false;
  h->dp_world = cfg->dp_world > 1 ? cfg->dp_world : 1;
  h->dp_rank = cfg->dp_world > 1 ? cfg->dp_rank : 0;
  // Few fat blocks: 16 x 1024 threads.  Alone the exchange is fastest with one 256-thread block per SM (critic arena, 2
  // GPUs: 32 us vs 42 us), but next to the kernels it overlaps with the small grid wins, and fatter blocks more so --
  // measured iters/s on one box each: 8 GPUs 3106 (16 x 1024) vs 3035 (8 x 1024) vs 3042 (32 x 256) vs 2745 (148 x 256);
  // 2 GPUs 785 (8 x 1024) vs 783 (16 x 512) vs 771 (32 x 256) vs 738 (148 x 128)  (profiles/r02_scaling.txt)
  h->dp_blocks = 16;
  h->dp_early_trigger = false;
  h->dp_stream = nullptr; h->dp_fork = nullptr; h->dp_join = nullptr;
  h->aux_stream = nullptr; h->aux_fork = nullptr; h->aux_join = nullptr;
  h->aux_overlap = getenv("WGAN_B200_NO_AUX_OVERLAP") == nullptr;
  if (const char* env = getenv("WGAN_B200_DP_BLOCKS")) {
    const int b = atoi(env);
    if (b >= 1 && b <= DP_MAX_BLOCKS) h->dp_blocks = b;
  }
  for (int r = 0; r < WGAN_DP_MAX_WORLD; ++r) { h->dp_peer[r] = nullptr; h->dp_peer_ipc[r] = false; }
  h->gemm_2cta = 1;
  if (const char* env = getenv("WGAN_B200_GEMM_2CTA")) h->gemm_2cta = atoi(env) != 0;
  h->prepare_reserve = 0;
  if (const char* env = getenv("WGAN_B200_PREPARE_SM_RESERVE")) h->prepare_reserve = atoi(env) > 0 ? atoi(env) : 0;
  h->gemm_msub = 0;
  if (const char* env = getenv("WGAN_B200_GEMM_MSUB")) {
    int c = atoi(env);
    if (c == 1 || c == 2) h->gemm_msub = c;
  }
  if (const char* env = getenv("WGAN_B200_GEMM_CLUSTER")) {
    int c = atoi(env);
    if (c == 1 || c == 2 || c == 4) h->gemm_cluster = c;
  }
  for (int n = 0; n < 2; ++n) { for (int k = 0; k < 4; ++k) h->arena[n][k] = nullptr; h->powers[n] = nullptr; }
  h->zbuf = h->gh1 = h->gh2 = h->xcat = h->h1 = h->h2 = h->d2 = h->d1 = h->dout = h->v2 = h->gd2 = h->gd1 =
      h->dz = h->nsq = h->crow = h->pen = h->norms = h->colpart = h->splitws = h->sample_tmp = nullptr;
  h->gram_k = h->gram_s = h->gram_t = nullptr;
  h->bits1 = h->bits2 = nullptr;
  auto fail = [&](const std::string& m) { g_create_error = m; wgan_destroy(h); return -1; };

  h->Z = cfg->noise_input_size; h->Hg = cfg->inflate_to_size; h->G = cfg->gex_size;
  h->Hd = cfg->disc_internal_size; h->R = cfg->max_batch;
  if (h->Z <= 0 || h->Hg <= 0 || h->G <= 0 || h->Hd <= 0 || h->R <= 0) return fail("non-positive dimension in config");
  if (cfg->gemm_mode != 0 && cfg->gemm_mode != 1) return fail("gemm_mode must be 0 (TF32 tcgen05) or 1 (FP32 verify)");
  if (cfg->gp_form != 0 && cfg->gp_form != 1) return fail("gp_form must be 0 (explicit chain) or 1 (Gram form)");
  if (cfg->dp_world > WGAN_DP_MAX_WORLD) return fail("dp_world exceeds WGAN_DP_MAX_WORLD");
  if (cfg->dp_world > 1 && (cfg->dp_rank < 0 || cfg->dp_rank >= cfg->dp_world)) return fail("dp_rank outside [0, dp_world)");
  h->Zl = round4(h->Z); h->Hgl = round4(h->Hg); h->Gl = round4(h->G); h->Hdl = round4(h->Hd);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail("no CUDA device: this library has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("device ordinal out of range");
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail("cudaSetDevice failed");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail("this library is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
  h->num_sms = prop.multiProcessorCount;
  h->sm_cap = h->num_sms;
  if (h->dp_blocks > h->num_sms) h->dp_blocks = h->num_sms;
  if (h->prepare_reserve > h->num_sms / 2) h->prepare_reserve = h->num_sms / 2;

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return fail("cuTensorMapEncodeTiled entry point not available");
  h->encode = reinterpret_cast<EncodeTiledFn>(fn);

  {
    float cdf[12];
    double acc = 0.0, term = exp(-1.0);
    for (int k = 0; k < 12; ++k) {
      if (k > 0) term /= k;
      acc += term;
      cdf[k] = static_cast<float>(acc);
    }
    if (cudaMemcpyToSymbol(c_poisson1_cdf, cdf, sizeof cdf) != cudaSuccess) return fail("constant upload failed");
  }

  // ---- arenas in TF variable creation order [W:128-130,140-142] ----
  static const char* names[12] = {
      "generator/gen_dense1/kernel", "generator/gen_dense1/bias", "generator/gen_dense2/kernel",
      "generator/gen_dense2/bias", "generator/gen_output/kernel", "generator/gen_output/bias",
      "discriminator/disc_dense1/kernel", "discriminator/disc_dense1/bias",
      "discriminator/disc_dense2/kernel", "discriminator/disc_dense2/bias",
      "discriminator/disc_output/kernel", "discriminator/disc_output/bias"};
  const int shp[12][2] = {{h->Z, h->Hg}, {1, h->Hg}, {h->Hg, h->Hg}, {1, h->Hg}, {h->Hg, h->G}, {1, h->G},
                          {h->G, h->Hd}, {1, h->Hd}, {h->Hd, h->Hd}, {1, h->Hd}, {h->Hd, 1}, {1, 1}};
  size_t off[2] = {0, 0};
  for (int t = 0; t < 12; ++t) {
    TensorDesc& d = h->td[t];
    d.name = names[t]; d.net = t < 6 ? WGAN_NET_GENERATOR : WGAN_NET_CRITIC;
    d.rows = shp[t][0]; d.cols = shp[t][1];
    d.ld = (d.rows == 1) ? d.cols : wgan_padded_ld(d.cols);
    d.offset = off[d.net];
    off[d.net] += round64(static_cast<size_t>(d.rows) * d.ld);
  }
  // both gradient arenas and the data-parallel flags share ONE allocation, the window peers map (wgan_dp_connect)
  for (int n = 0; n < 2; ++n) h->arena_n[n] = off[n] + 4;
  h->dp_grad_off[WGAN_NET_CRITIC] = 0;
  h->dp_grad_off[WGAN_NET_GENERATOR] = round64(h->arena_n[WGAN_NET_CRITIC]);
  h->dp_stage_off = h->dp_grad_off[WGAN_NET_GENERATOR] + round64(h->arena_n[WGAN_NET_GENERATOR]);
  {
    const size_t big = h->arena_n[0] > h->arena_n[1] ? h->arena_n[0] : h->arena_n[1];
    h->dp_stage_stride = h->dp_world > 1 ? round64(big / h->dp_world + 8) : 0;   // >= the largest slice
  }
  h->dp_flags_off = h->dp_stage_off + static_cast<size_t>(h->dp_world > 1 ? h->dp_world : 0) * h->dp_stage_stride;
  h->dp_win_bytes = (h->dp_flags_off + DP_FLAG_WORDS) * 4;
  if (cudaMalloc(&h->dp_win, h->dp_win_bytes) != cudaSuccess) return fail("cudaMalloc(gradient window) failed");
  if (cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->aux_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->aux_join, cudaEventDisableTiming) != cudaSuccess)
    return fail("creating the auxiliary stream failed");
  if (h->dp_world > 1) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&h->dp_stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->dp_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->dp_join, cudaEventDisableTiming) != cudaSuccess)
      return fail("creating the exchange stream failed");
  }
  cudaMemset(h->dp_win, 0, h->dp_win_bytes);
  h->dp_peer[h->dp_rank] = h->dp_win;
  for (int n = 0; n < 2; ++n) {
    for (int k = 0; k < 4; ++k) {
      if (k == WGAN_KIND_GRAD) { h->arena[n][k] = h->dp_win + h->dp_grad_off[n]; continue; }
      if (cudaMalloc(&h->arena[n][k], h->arena_n[n] * 4) != cudaSuccess) return fail("cudaMalloc(arena) failed");
      cudaMemset(h->arena[n][k], 0, h->arena_n[n] * 4);
    }
    if (cudaMalloc(&h->powers[n], 16) != cudaSuccess) return fail("cudaMalloc(powers) failed");
    float pw[4] = {cfg->beta1, cfg->beta2, 0.f, 0.f};        // beta powers start at beta [A:123-128]
    cudaMemcpy(h->powers[n], pw, 16, cudaMemcpyHostToDevice);
  }

  // ---- workspace ----
  const size_t R = h->R;
  h->nsq_tiles_max = cdiv(h->G, 32) * (NUM_EPI_WARPS / 4);
  h->colpart_stride = 64 * (h->G > h->Hg ? (h->G > h->Hd ? h->G : h->Hd) : (h->Hg > h->Hd ? h->Hg : h->Hd));
  struct { float** p; size_t n; } allocs[] = {
      {&h->zbuf, R * h->Zl}, {&h->gh1, R * h->Hgl}, {&h->gh2, R * h->Hgl}, {&h->xcat, 3 * R * h->Gl},
      {&h->h1, 3 * R * h->Hdl}, {&h->h2, 3 * R * h->Hdl}, {&h->d2, 3 * R * h->Hdl}, {&h->d1, 3 * R * h->Hdl},
      {&h->dout, 3 * R}, {&h->v2, R * h->Hdl}, {&h->gd2, R * h->Hgl}, {&h->gd1, R * h->Hgl}, {&h->dz, R * h->Zl},
      {&h->nsq, static_cast<size_t>(h->nsq_tiles_max) * R}, {&h->crow, R}, {&h->pen, R}, {&h->norms, R},
      {&h->colpart, static_cast<size_t>(3) * h->colpart_stride},
      {&h->sample_tmp, R * h->Gl},
      {&h->gram_k, static_cast<size_t>(h->Hd) * h->Hdl}, {&h->gram_s, static_cast<size_t>(h->Hd) * h->Hdl},
      {&h->gram_t, R * h->Hdl}};
  for (auto& a : allocs) {
    if (cudaMalloc(a.p, (a.n + 64) * 4) != cudaSuccess) return fail("cudaMalloc(workspace) failed (max_batch too large?)");
    cudaMemset(*a.p, 0, (a.n + 64) * 4);
  }
  for (uint32_t** b : {&h->bits1, &h->bits2}) {
    if (cudaMalloc(b, 3 * R * 8 * sizeof(uint32_t)) != cudaSuccess) return fail("cudaMalloc(slope bits) failed");
    cudaMemset(*b, 0, 3 * R * 8 * sizeof(uint32_t));
  }
  h->splitws_n = static_cast<size_t>(3 * h->num_sms + 8) * BM * 256;
  {
    // a two-part wgrad of the largest weight must fit with one slab per part
    size_t w1 = round64(static_cast<size_t>(h->G) * h->Hdl), w3 = round64(static_cast<size_t>(h->Hg) * h->Gl);
    size_t need = 2 * (w1 > w3 ? w1 : w3);
    if (h->splitws_n < need) h->splitws_n = need;
  }
  if (cudaMalloc(&h->splitws, h->splitws_n * 4) != cudaSuccess) return fail("cudaMalloc(split workspace) failed");
  {
    const size_t nflags = static_cast<size_t>(h->num_sms / 2 + 1) * 2 * NUM_EPI_WARPS;
    if (cudaMalloc(&h->sk_flags, nflags * 4) != cudaSuccess) return fail("cudaMalloc(stream-K flags) failed");
    cudaMemset(h->sk_flags, 0, nflags * 4);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return fail("device error during create");
  *out = h;
  return 0;
}

int wgan_get_tensor_info(const wgan_handle* h, int index, wgan_tensor_info* out) {
  if (!h || !out || index < 0 || index >= 12) return -2;
  const TensorDesc& d = h->td[index];
  memset(out, 0, sizeof *out);
  strncpy(out->name, d.name, sizeof out->name - 1);
  out->net = d.net; out->rows = d.rows; out->cols = d.cols; out->ld = d.ld;
  out->offset = static_cast<int64_t>(d.offset);
  return 0;
}

int wgan_get_tensor(wgan_handle* h, int kind, int index, float* host_dst) {
  if (!h) return -2;
  if (kind < 0 || kind > 3 || index < 0 || index >= 12 || !host_dst) FAIL("get_tensor: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  const TensorDesc& d = h->td[index];
  CK(cudaMemcpy2D(host_dst, static_cast<size_t>(d.cols) * 4, h->arena[d.net][kind] + d.offset,
                  static_cast<size_t>(d.ld) * 4, static_cast<size_t>(d.cols) * 4, d.rows, cudaMemcpyDeviceToHost));
  return 0;
}

int wgan_set_tensor(wgan_handle* h, int kind, int index, const float* host_src) {
  if (!h) return -2;
  if (kind < 0 || kind > 3 || index < 0 || index >= 12 || !host_src) FAIL("set_tensor: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  const TensorDesc& d = h->td[index];
  CK(cudaMemcpy2D(h->arena[d.net][kind] + d.offset, static_cast<size_t>(d.ld) * 4, host_src,
                  static_cast<size_t>(d.cols) * 4, static_cast<size_t>(d.cols) * 4, d.rows, cudaMemcpyHostToDevice));
  return 0;
}

int wgan_get_adam_powers(wgan_handle* h, int net, float* b1p, float* b2p) {
  if (!h) return -2;
  if (net < 0 || net > 1 || !b1p || !b2p) FAIL("get_adam_powers: bad argument");
  float pw[2];
  CK(cudaMemcpy(pw, h->powers[net], 8, cudaMemcpyDeviceToHost));
  *b1p = pw[0]; *b2p = pw[1];
  return 0;
}

int wgan_set_adam_powers(wgan_handle* h, int net, float b1p, float b2p) {
  if (!h) return -2;
  if (net < 0 || net > 1) FAIL("set_adam_powers: bad argument");
  float pw[2] = {b1p, b2p};
  CK(cudaMemcpy(h->powers[net], pw, 8, cudaMemcpyHostToDevice));
  return 0;
}

int wgan_get_arena(wgan_handle* h, int net, int kind, float** dev_ptr, size_t* n_floats) {
  if (!h) return -2;
  if (net < 0 || net > 1 || kind < 0 || kind > 3 || !dev_ptr || !n_floats) FAIL("get_arena: bad argument");
  *dev_ptr = h->arena[net][kind];
  *n_floats = h->arena_n[net];
  return 0;
}

int wgan_get_scalars(wgan_handle* h, int net, float* host4, wgan_stream_t stream) {
  if (!h) return -2;
  if (net < 0 || net > 1 || !host4) FAIL("get_scalars: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaMemcpyAsync(host4, h->scalars(net), 16, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Critic step gradients  (reference: obj_d [W:137], var_list d_params [W:154])
// ---------------------------------------------------------------------------------------------
// Phase A of the critic step: everything that does not depend on the critic's parameters -- input staging,
// generator forward x~ = G(z) and the interpolate x^.  In data-parallel runs the host issues it for update
// k+1 while the gradient all-reduce of update k is in flight.
static int critic_phase_a(wgan_handle* h, const float* x_dev, int ldx, const float* z_dev, int ldz,
                          const float* eps_dev, int rows, cudaStream_t s) {
  const bool gp = h->cfg.lambda_gp != 0.0f;
  const int Bl = rows, nh = gp ? 1 : 0;
  const int G = h->G, Gl = h->Gl;
  const bool gram = gp && h->cfg.gp_form == 1;
  const int real0 = gram ? Bl : (1 + nh) * Bl;
  const float* z; int lz;
  TRY(stage_input(h, z_dev, ldz, Bl, h->Z, h->zbuf, h->Zl, &z, &lz, s));
  // the real cells always live in their slot of the stacked buffer (layer 1 and its weight gradient are single
  // products over the stacked rows, and the weight gradient's column-chunked reads never touch caller memory)
  float* xreal_slot = h->xcat + static_cast<size_t>(real0) * Gl;
  if (x_dev != xreal_slot || ldx != Gl)
    CK(cudaMemcpy2DAsync(xreal_slot, static_cast<size_t>(Gl) * 4, x_dev, static_cast<size_t>(ldx) * 4,
                         static_cast<size_t>(G) * 4, Bl, cudaMemcpyDeviceToDevice, s));
  if (gram) return generator_forward(h, z, lz, Bl, h->xcat, Gl, s);   // rows [fake | real]; x^ is never formed
  const float* x = xreal_slot; const int lx = Gl;
  float* xt = h->xcat;
  float* xh = h->xcat + static_cast<size_t>(Bl) * Gl;    // interpolates, later overwritten by gx
  // the interpolate x^ = eps x + (1 - eps) x~ is a second output of the generator-output GEMM
  static const bool fuse_blend = getenv("WGAN_B200_NO_FUSE_BLEND") == nullptr;
  if (gp && fuse_blend) {
    TRY(generator_forward(h, z, lz, Bl, xt, Gl, s, x, lx, eps_dev, xh));
  } else {
    TRY(generator_forward(h, z, lz, Bl, xt, Gl, s));
    if (gp) {
      KLAUNCH(blend_kernel, ew_grid(static_cast<size_t>(Bl) * (Gl / 4), 256, h->num_sms), 256, 0, s, 
          x, lx, xt, Gl, eps_dev, xh, Gl, Bl, Gl / 4);
      TRY(post_launch(h, "blend_kernel"));
    }
  }
  return 0;
}

// Critic gradients with the penalty in Gram form (cfg.gp_form == 1; SURVEY 8d shortcuts i + ii).  Same
// loss and gradients as the explicit chain in real arithmetic; every product with the gene dimension on
// the interpolate rows is replaced by one with K = W1^T W1 or S = (c g1)^T g1 [Hd, Hd]:
//   a1(x^) = eps a1(x) + (1 - eps) a1(x~)            (layer 1 is linear before its activation)
//   n_b^2  = g1_b K g1_b^T,   v1 = c (g1 K) . s1,    dW1 += W1 S
// Stacked row order here: [fake | real | interpolate]; x~ and x are rows [0, 2 Bl) of xcat.
static int critic_grads_gram(wgan_handle* h, const float* eps_dev, int Bl, int global_batch, cudaStream_t s) {
  const int T = 3 * Bl, T2 = 2 * Bl;
  const int Hd = h->Hd, Hdl = h->Hdl, G = h->G, Gl = h->Gl;
  const float a = h->cfg.alpha, invB = 1.0f / static_cast<float>(global_batch);
  const float* W1 = h->P(T_W1); const int ldw1 = h->td[T_W1].ld;
  const float* W2 = h->P(T_W2); const int ldw2 = h->td[T_W2].ld;
  const size_t hat0 = static_cast<size_t>(T2) * Hdl;
  float* h1_hat = h->h1 + hat0;
  float* h2_hat = h->h2 + hat0;
  float* g1 = h->d1 + hat0;

  // layer 1: raw products of [fake; real], then bias + activation + the interpolate's blend  [W:107-111]
  {
    int slabs = 0;
    Part p1{h->xcat, Gl, W1, ldw1, G};
    TRY(run_gemm(h, false, true, &p1, 1, h->h1, Hdl, T2, Hd, epi_none(), nullptr, 0, s, nullptr, 0, &slabs, false));
    const float* part = slabs > 0 ? h->splitws : h->h1;
    const size_t n = static_cast<size_t>(Bl) * ((Hd + 3) / 4);
    KLAUNCH(l1_gram_finish_kernel, ew_grid(n, 256, h->num_sms), 256, 0, s, part,
            round64(static_cast<size_t>(T2) * Hdl), slabs > 0 ? slabs : 1, h->h1, Hdl, Bl, Hd, h->P(T_B1), eps_dev, a);
    TRY(post_launch(h, "l1_gram_finish_kernel"));
  }
  // layer 2, layer 3 + upstream deltas (fake +1/B, real -1/B, interpolate g2 = w3 . s2), delta1 / g1
  TRY(critic_middle(h, T, Bl, invB, -invB, 1.0f, s));

  // K = W1^T W1, T = g1 K, per-row norm / coefficient, v1 (over h1_hat), T <- c g1, v2 = (v1 W2) . s2
  TRY(gemm1(h, true, true, W1, ldw1, W1, ldw1, h->gram_k, Hdl, Hd, Hd, G, epi_none(), s));
  h->last_gx_rows = Bl;
  if (mid_chain_available(h)) {
    MidParams x;
    memset(&x, 0, sizeof x);
    x.bits1 = h->bits1 + static_cast<size_t>(T2) * 8; x.bits2 = h->bits2 + static_cast<size_t>(T2) * 8;   // interpolate rows
    x.two_lambda_over_B = 2.0f * h->cfg.lambda_gp * invB;
    x.crow = h->crow; x.pen = h->pen; x.norms = h->norms;
    TRY(run_mid_chain(h, 1, Bl, g1, h->gram_k, Hdl, W2, ldw2, h->gram_t, h1_hat, h->v2, x, s));
  } else {
    TRY(gemm1(h, false, true, g1, Hdl, h->gram_k, Hdl, h->gram_t, Hdl, Bl, Hd, Hd, epi_none(), s));
    KLAUNCH(gp_coef_gram_kernel, ew_grid(static_cast<size_t>(Bl) * 32, 256, h->num_sms), 256, 0, s,
        h->gram_t, g1, h1_hat, Hdl, Bl, Hd, 2.0f * h->cfg.lambda_gp * invB, h->crow, h->pen, h->norms, a);
    TRY(post_launch(h, "gp_coef_gram_kernel"));
    TRY(gemm1(h, false, true, h1_hat, Hdl, W2, ldw2, h->v2, Hdl, Bl, Hd, Hd, epi_mask(h2_hat, Hdl, a), s));
  }
  // S = (c g1)^T g1
  TRY(gemm1(h, true, true, h->gram_t, Hdl, g1, Hdl, h->gram_s, Hdl, Hd, Hd, Bl, epi_none(), s));

  // dW1 = W1 S + [x~; x]^T [delta1_fake; delta1_real]: raw slabs of both products, one finish pass
  {
    int s0 = 0, s1 = 0;
    Part ps{W1, ldw1, h->gram_s, Hdl, Hd};
    TRY(run_gemm(h, false, true, &ps, 1, h->Gr(T_W1), ldw1, G, Hd, epi_none(), nullptr, 0, s, nullptr, 0, &s0, true));
    Part pw{h->xcat, Gl, h->d1, Hdl, T2};
    TRY(run_gemm(h, true, true, &pw, 1, h->Gr(T_W1), ldw1, G, Hd, epi_none(), nullptr, 0, s, nullptr, s0, &s1, true));
    TRY(launch_finish(h, h->splitws, round64(static_cast<size_t>(G) * ldw1), s0 + s1, h->Gr(T_W1), ldw1, G, Hd,
                      epi_none(), s));
  }
  // dW2 = [h1_fake; h1_real; v1]^T [delta2_fake; delta2_real; g2]
  TRY(gemm1(h, true, true, h->h1, Hdl, h->d2, Hdl, h->Gr(T_W2), ldw2, Hd, Hd, T, epi_none(), s));
  {
    ColsumJob jobs[3];
    memset(jobs, 0, sizeof jobs);
    jobs[0].seg[0] = ColsumSeg{h->d1, T2, Hdl, 1.0f};
    jobs[0].nseg = 1; jobs[0].N = Hd; jobs[0].out = h->Gr(T_B1);
    jobs[1].seg[0] = ColsumSeg{h->d2, T2, Hdl, 1.0f};
    jobs[1].nseg = 1; jobs[1].N = Hd; jobs[1].out = h->Gr(T_B2);
    jobs[2].seg[0] = ColsumSeg{h->h2, Bl, Hdl, invB};
    jobs[2].seg[1] = ColsumSeg{h->h2 + static_cast<size_t>(Bl) * Hdl, Bl, Hdl, -invB};
    jobs[2].seg[2] = ColsumSeg{h->v2, Bl, Hdl, 1.0f};
    jobs[2].nseg = 3; jobs[2].N = Hd; jobs[2].out = h->Gr(T_W3);
    // d b3 = sum(+1/B) + sum(-1/B) over equal row counts = 0; loss scalars: mean parts of D_fake / D_real, penalty
    float* sc = h->scalars(WGAN_NET_CRITIC);
    const SumJob sums[3] = {{h->dout, sc + WGAN_S_DFAKE, Bl, invB}, {h->dout + Bl, sc + WGAN_S_DREAL, Bl, invB},
                            {h->pen, sc + WGAN_S_GP, Bl, h->cfg.lambda_gp * invB}};
    TRY(colsum_jobs(h, jobs, 3, s, h->Gr(T_B3), sums, 3));
  }
  return 0;
}

int wgan_critic_prepare(wgan_handle* h, const float* x_dev, int ldx, const float* z_dev, int ldz,
                        const float* eps_dev, int rows, wgan_stream_t stream, uint64_t* token_out) {
  if (!h) return -2;
  if (token_out) *token_out = 0;
  if (!x_dev || !z_dev) FAIL("critic_prepare: null input");
  TRY(check_rows(h, rows, rows));
  CK(cudaSetDevice(h->cfg.device));
  if (h->cfg.lambda_gp != 0.0f && !eps_dev) FAIL("critic_prepare: eps required when lambda_gp != 0");
  invalidate_workspace(h);
  h->sm_cap = h->num_sms - h->prepare_reserve;
  const int rc = critic_phase_a(h, x_dev, ldx, z_dev, ldz, eps_dev, rows, static_cast<cudaStream_t>(stream));
  h->sm_cap = h->num_sms;
  TRY(rc);
  h->prep_token = h->ws_gen; h->prep_rows = rows;
  if (token_out) *token_out = h->prep_token;
  return 0;
}

int wgan_critic_grads(wgan_handle* h, const float* x_dev, int ldx, const float* z_dev, int ldz,
                      const float* eps_dev, int rows, int global_batch, uint64_t prepared_token,
                      wgan_stream_t stream) {
  if (!h) return -2;
  if (!x_dev || !z_dev) FAIL("critic_grads: null input");
  TRY(check_rows(h, rows, global_batch));
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool gp = h->cfg.lambda_gp != 0.0f;
  if (gp && !eps_dev) FAIL("critic_grads: eps required when lambda_gp != 0");
  const int Bl = rows, nh = gp ? 1 : 0, T = (2 + nh) * Bl;
  const int Hd = h->Hd, Hdl = h->Hdl, G = h->G, Gl = h->Gl;
  const float a = h->cfg.alpha, invB = 1.0f / static_cast<float>(global_batch);
  const int real0 = (1 + nh) * Bl;                       // first real row in the stacked buffers

  // the prepared activations are still intact iff nothing has touched the workspace since that prepare call
  const bool prepared = prepared_token != 0 && prepared_token == h->prep_token && prepared_token == h->ws_gen &&
                        h->prep_rows == rows;
  invalidate_workspace(h);
  if (!prepared) TRY(critic_phase_a(h, x_dev, ldx, z_dev, ldz, eps_dev, rows, s));
  if (gp && h->cfg.gp_form == 1) return critic_grads_gram(h, eps_dev, rows, global_batch, s);
  float* xh = h->xcat + static_cast<size_t>(Bl) * Gl;

  // critic layer 1 on [fake | hat | real]  [W:107-111]
  const EpiParams e1 = epi_bias_lrelu(h->P(T_B1), a);
  const float* W1 = h->P(T_W1); const int ldw1 = h->td[T_W1].ld;
  TRY(gemm1(h, false, true, h->xcat, Gl, W1, ldw1, h->h1, Hdl, T, Hd, G, e1, s));
  // layer 2 [W:113-117], layer 3 + upstream deltas [W:119-123,137], delta1 / g1 = (delta2 W2^T) . slope(h1)
  const float* W2 = h->P(T_W2); const int ldw2 = h->td[T_W2].ld;
  TRY(critic_middle(h, T, Bl, invB, gp ? 1.0f : -invB, gp ? -invB : 0.0f, s));

  float* h1_hat = h->h1 + static_cast<size_t>(Bl) * Hdl;
  float* h2_hat = h->h2 + static_cast<size_t>(Bl) * Hdl;
  float* d1_hat = h->d1 + static_cast<size_t>(Bl) * Hdl;
  if (gp) {
    // input-gradient pass: gx = g1 W1^T, fused per-row sum of squares
    EpiParams eg = epi_none();
    eg.row_sumsq = h->nsq;
    int tiles = 1;
    TRY(gemm1(h, false, false, d1_hat, Hdl, W1, ldw1, xh, Gl, Bl, G, Hd, eg, s, &tiles));
    h->last_gx_rows = Bl;
    // norm, penalty, per-row coefficient; g1 <- c . g1
    KLAUNCH(gp_coef_kernel, ew_grid(static_cast<size_t>(Bl) * 32, 256, h->num_sms), 256, 0, s,
        h->nsq, tiles, Bl, Bl, 2.0f * h->cfg.lambda_gp * invB, h->crow, h->pen, h->norms, d1_hat, Hdl, Hd);
    TRY(post_launch(h, "gp_coef_kernel"));
    // parameter-gradient-of-norm pass: v1 = c . (gx W1) . s1 (in place over h1_hat), v2 = (v1 W2) . s2
    EpiParams ev1 = epi_mask(h1_hat, Hdl, a);
    ev1.row_scale = h->crow;
    TRY(gemm1(h, false, true, xh, Gl, W1, ldw1, h1_hat, Hdl, Bl, Hd, G, ev1, s));
    TRY(gemm1(h, false, true, h1_hat, Hdl, W2, ldw2, h->v2, Hdl, Bl, Hd, Hd, epi_mask(h2_hat, Hdl, a), s));
  }

  // Bias gradients, the layer-3 weight gradient and the loss scalars are column sums over buffers that are final by
  // now; they run on the auxiliary stream NEXT TO the two weight-gradient GEMMs below (small non-persistent grids
  // that co-reside with the persistent GEMM CTAs), joined before the function returns.
  const bool ov = h->aux_overlap;
  cudaStream_t sc_stream = s;
  if (ov) {
    CK(cudaEventRecord(h->aux_fork, s));
    CK(cudaStreamWaitEvent(h->aux_stream, h->aux_fork, 0));
    sc_stream = h->aux_stream;
  }
  {
    ColsumJob jobs[3];
    memset(jobs, 0, sizeof jobs);
    jobs[0].seg[0] = ColsumSeg{h->d1, Bl, Hdl, 1.0f};
    jobs[0].seg[1] = ColsumSeg{h->d1 + static_cast<size_t>(real0) * Hdl, Bl, Hdl, 1.0f};
    jobs[0].nseg = 2; jobs[0].N = Hd; jobs[0].out = h->Gr(T_B1);
    jobs[1].seg[0] = ColsumSeg{h->d2, Bl, Hdl, 1.0f};
    jobs[1].seg[1] = ColsumSeg{h->d2 + static_cast<size_t>(real0) * Hdl, Bl, Hdl, 1.0f};
    jobs[1].nseg = 2; jobs[1].N = Hd; jobs[1].out = h->Gr(T_B2);
    jobs[2].seg[0] = ColsumSeg{h->h2, Bl, Hdl, invB};
    jobs[2].seg[1] = ColsumSeg{h->h2 + static_cast<size_t>(real0) * Hdl, Bl, Hdl, -invB};
    jobs[2].seg[2] = ColsumSeg{h->v2, Bl, Hdl, 1.0f};
    jobs[2].nseg = gp ? 3 : 2; jobs[2].N = Hd; jobs[2].out = h->Gr(T_W3);
    // d b3 = sum(+1/B) + sum(-1/B) over equal row counts = 0; loss scalars: mean parts of D_fake / D_real and the
    // penalty (an empty sum clears it when there is none)
    float* sc = h->scalars(WGAN_NET_CRITIC);
    const SumJob sums[3] = {{h->dout, sc + WGAN_S_DFAKE, Bl, invB}, {h->dout + real0, sc + WGAN_S_DREAL, Bl, invB},
                            {h->pen, sc + WGAN_S_GP, gp ? Bl : 0, h->cfg.lambda_gp * invB}};
    TRY(colsum_jobs(h, jobs, 3, sc_stream, h->Gr(T_B3), sums, 3));
  }
  if (ov) CK(cudaEventRecord(h->aux_join, h->aux_stream));
  // weight gradients
  TRY(gemm1(h, true, true, h->xcat, Gl, h->d1, Hdl, h->Gr(T_W1), ldw1, G, Hd, T, epi_none(), s));
  TRY(gemm1(h, true, true, h->h1, Hdl, h->d2, Hdl, h->Gr(T_W2), ldw2, Hd, Hd, T, epi_none(), s));
  if (ov) CK(cudaStreamWaitEvent(s, h->aux_join, 0));
  return 0;
}

int wgan_critic_apply(wgan_handle* h, wgan_stream_t stream) {
  if (!h) return -2;
  CK(cudaSetDevice(h->cfg.device));
  return adam_apply(h, WGAN_NET_CRITIC, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// Generator step gradients  (obj_g = -mean(D(G(z))) [W:138], var_list g_params [W:155])
// ---------------------------------------------------------------------------------------------
// Phase A of the generator step: stage z and run the generator forward (does not depend on the critic).
static int generator_phase_a(wgan_handle* h, const float* z_dev, int ldz, int rows, cudaStream_t s) {
  // z is later the MN-major A operand of the first-layer weight gradient: the last 32-column chunk of an
  // MN-major tile may read up to 124 B past a row, which the engine's own buffers tolerate (slack) but a
  // caller's tensor need not -- always work on the staged copy (1.6 MB at batch 4096)
  KLAUNCH(copy2d_kernel, ew_grid(static_cast<size_t>(rows) * h->Z, 256, h->num_sms), 256, 0, s, z_dev, ldz, h->zbuf,
          h->Zl, rows, h->Z);
  TRY(post_launch(h, "copy2d_kernel"));
  return generator_forward(h, h->zbuf, h->Zl, rows, h->xcat, h->Gl, s);
}

int wgan_generator_prepare(wgan_handle* h, const float* z_dev, int ldz, int rows, wgan_stream_t stream,
                           uint64_t* token_out) {
  if (!h) return -2;
  if (token_out) *token_out = 0;
  if (!z_dev) FAIL("generator_prepare: null input");
  TRY(check_rows(h, rows, rows));
  CK(cudaSetDevice(h->cfg.device));
  invalidate_workspace(h);
  h->sm_cap = h->num_sms - h->prepare_reserve;
  const int rc = generator_phase_a(h, z_dev, ldz, rows, static_cast<cudaStream_t>(stream));
  h->sm_cap = h->num_sms;
  TRY(rc);
  h->gprep_token = h->ws_gen; h->gprep_rows = rows;
  if (token_out) *token_out = h->gprep_token;
  return 0;
}

// In-place SUM over the ranks of floats [off, off + n) of net's gradient arena (off, n multiples of 4).
static int dp_exchange_range(wgan_handle* h, int net, size_t off, size_t n, cudaStream_t s);

static int generator_grads_impl(wgan_handle* h, const float* z_dev, int ldz, int rows, int global_batch,
                                uint64_t prepared_token, wgan_stream_t stream, bool exchange) {
  if (!h) return -2;
  if (!z_dev) FAIL("generator_grads: null input");
  TRY(check_rows(h, rows, global_batch));
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Bl = rows, Hd = h->Hd, Hdl = h->Hdl, G = h->G, Gl = h->Gl, Hg = h->Hg, Hgl = h->Hgl;
  const float a = h->cfg.alpha, invB = 1.0f / static_cast<float>(global_batch);
  const bool prepared = prepared_token != 0 && prepared_token == h->gprep_token && prepared_token == h->ws_gen &&
                        h->gprep_rows == rows;
  invalidate_workspace(h);
  if (!prepared) TRY(generator_phase_a(h, z_dev, ldz, rows, s));
  const float* z = h->zbuf; const int lz = h->Zl;
  float* xt = h->xcat;
  float* d3 = h->xcat + static_cast<size_t>(Bl) * Gl;
  const float* W1 = h->P(T_W1); const int ldw1 = h->td[T_W1].ld;
  TRY(gemm1(h, false, true, xt, Gl, W1, ldw1, h->h1, Hdl, Bl, Hd, G, epi_bias_lrelu(h->P(T_B1), a), s));
  // critic layers 2 and 3, then back through layer 2 towards the generated cells
  TRY(critic_middle(h, Bl, Bl, -invB, 0.f, 0.f, s));
  TRY(gemm1(h, false, false, h->d1, Hdl, W1, ldw1, d3, Gl, Bl, G, Hd, epi_mask(xt, Gl, a), s));
  // generator backward
  const float* Wg3 = h->P(T_WG3); const int ldg3 = h->td[T_WG3].ld;
  const float* Wg2 = h->P(T_WG2); const int ldg2 = h->td[T_WG2].ld;
  const bool ov = h->aux_overlap;
  {
    // output-bias gradient (column sum of d3, 108 MB at batch 4096) on the auxiliary stream NEXT TO the output
    // layer's weight gradient, which reads the same d3
    cudaStream_t sc_stream = s;
    if (ov) {
      CK(cudaEventRecord(h->aux_fork, s));
      CK(cudaStreamWaitEvent(h->aux_stream, h->aux_fork, 0));
      sc_stream = h->aux_stream;
    }
    ColsumJob jb;
    memset(&jb, 0, sizeof jb);
    jb.seg[0] = ColsumSeg{d3, Bl, Gl, 1.0f};
    jb.nseg = 1; jb.N = G; jb.out = h->Gr(T_BG3);
    // (the loss scalar rides along: with the exchange below it must be final before the output layer's range goes out)
    float* sc = h->scalars(WGAN_NET_GENERATOR);
    const SumJob gj = {h->dout, sc + WGAN_S_DFAKE, Bl, invB};          // mean D(G(z)) for obj_g [W:138]
    TRY(colsum_jobs(h, &jb, 1, sc_stream, nullptr, &gj, 1));
    if (ov) CK(cudaEventRecord(h->aux_join, h->aux_stream));
  }
  TRY(gemm1(h, true, true, h->gh2, Hgl, d3, Gl, h->Gr(T_WG3), ldg3, Hg, G, Bl, epi_none(), s));
  if (ov) CK(cudaStreamWaitEvent(s, h->aux_join, 0));
  // Data-parallel exchange, first part: the output layer's gradients (kernel + bias, 90 % of the arena) and the
  // scalar tail are final -- sum them over the ranks on a second stream WHILE the rest of the backward pass runs
  // (measured on 2 GPUs: the 17.6 MB exchange after the step costs 62 us fully exposed; ~125 us of backward work
  // follows this point).  The small remainder is exchanged at the end.
  const size_t early_off = h->td[T_WG3].offset;
  exchange = exchange && h->dp_world > 1;
  if (exchange) {
    if (!h->dp_connected) FAIL("generator_grads_exchange: call wgan_dp_connect first");
    CK(cudaEventRecord(h->dp_fork, s));
    CK(cudaStreamWaitEvent(h->dp_stream, h->dp_fork, 0));
    TRY(dp_exchange_range(h, WGAN_NET_GENERATOR, early_off, h->arena_n[WGAN_NET_GENERATOR] - early_off, h->dp_stream));
    CK(cudaEventRecord(h->dp_join, h->dp_stream));
    h->sm_cap = h->num_sms - h->prepare_reserve;     // the persistent GEMMs next to the exchange leave SMs for it
  }
  struct CapGuard { wgan_handle* h; ~CapGuard() { h->sm_cap = h->num_sms; } } cap_guard{h};
  TRY(gemm1(h, false, false, d3, Gl, Wg3, ldg3, h->gd2, Hgl, Bl, Hg, G, epi_mask(h->gh2, Hgl, a), s));
  TRY(gemm1(h, true, true, h->gh1, Hgl, h->gd2, Hgl, h->Gr(T_WG2), ldg2, Hg, Hg, Bl, epi_none(), s));
  TRY(gemm1(h, false, false, h->gd2, Hgl, Wg2, ldg2, h->gd1, Hgl, Bl, Hg, Hg, epi_mask(h->gh1, Hgl, a), s));
  TRY(gemm1(h, true, true, z, lz, h->gd1, Hgl, h->Gr(T_WG1), h->td[T_WG1].ld, h->Z, Hg, Bl, epi_none(), s));
  {
    // the two hidden-layer bias gradients and the loss scalar in one launch
    ColsumJob jobs[2];
    memset(jobs, 0, sizeof jobs);
    jobs[0].seg[0] = ColsumSeg{h->gd2, Bl, Hgl, 1.0f};
    jobs[0].nseg = 1; jobs[0].N = Hg; jobs[0].out = h->Gr(T_BG2);
    jobs[1].seg[0] = ColsumSeg{h->gd1, Bl, Hgl, 1.0f};
    jobs[1].nseg = 1; jobs[1].N = Hg; jobs[1].out = h->Gr(T_BG1);
    TRY(colsum_jobs(h, jobs, 2, s));
  }
  if (exchange) {
    // both parts use the same staging rows and flags: the second waits for the first
    CK(cudaStreamWaitEvent(s, h->dp_join, 0));
    TRY(dp_exchange_range(h, WGAN_NET_GENERATOR, 0, early_off, s));
  }
  return 0;
}

int wgan_generator_grads(wgan_handle* h, const float* z_dev, int ldz, int rows, int global_batch,
                         uint64_t prepared_token, wgan_stream_t stream) {
  return generator_grads_impl(h, z_dev, ldz, rows, global_batch, prepared_token, stream, false);
}

int wgan_generator_grads_exchange(wgan_handle* h, const float* z_dev, int ldz, int rows, int global_batch,
                                  uint64_t prepared_token, wgan_stream_t stream) {
  return generator_grads_impl(h, z_dev, ldz, rows, global_batch, prepared_token, stream, true);
}

int wgan_generator_apply(wgan_handle* h, wgan_stream_t stream) {
  if (!h) return -2;
  CK(cudaSetDevice(h->cfg.device));
  return adam_apply(h, WGAN_NET_GENERATOR, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// Data parallelism: rendezvous + the engine's own all-reduce over NVLink peer memory
// ---------------------------------------------------------------------------------------------
namespace {
struct DpBlob {                      // WGAN_DP_BLOB_BYTES, opaque to the caller
  uint32_t magic;
  int32_t rank, world, device;
  int64_t pid;
  uint64_t ptr;                      // window base in the exporting process
  uint64_t win_bytes;
  uint64_t arena_n[2];
  int32_t blocks;
  int32_t pad;
  cudaIpcMemHandle_t ipc;            // 64 bytes
  char uuid[16];                     // physical GPU of the exporting handle
};
static_assert(sizeof(DpBlob) <= WGAN_DP_BLOB_BYTES, "DpBlob must fit the rendezvous blob");
static_assert(DP_MAX_WORLD == WGAN_DP_MAX_WORLD, "kernel and ABI disagree on the maximum group size");
constexpr uint32_t kDpMagic = 0x57474450u;   // "WGDP"
}  // namespace

int wgan_dp_export(wgan_handle* h, void* blob_out) {
  if (!h) return -2;
  if (!blob_out) FAIL("dp_export: null argument");
  CK(cudaSetDevice(h->cfg.device));
  DpBlob b;
  memset(&b, 0, sizeof b);
  b.magic = kDpMagic; b.rank = h->dp_rank; b.world = h->dp_world; b.device = h->cfg.device;
  b.pid = static_cast<int64_t>(getpid());
  b.ptr = reinterpret_cast<uint64_t>(h->dp_win);
  b.win_bytes = h->dp_win_bytes;
  b.arena_n[0] = h->arena_n[0]; b.arena_n[1] = h->arena_n[1];
  b.blocks = h->dp_blocks;
  {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->cfg.device));
    memcpy(b.uuid, prop.uuid.bytes, 16);
  }
  CK(cudaIpcGetMemHandle(&b.ipc, h->dp_win));
  memset(blob_out, 0, WGAN_DP_BLOB_BYTES);
  memcpy(blob_out, &b, sizeof b);
  return 0;
}

int wgan_dp_connect(wgan_handle* h, const void* blobs) {
  if (!h) return -2;
  if (!blobs) FAIL("dp_connect: null argument");
  if (h->dp_world <= 1) { h->dp_connected = true; return 0; }
  if (h->dp_connected) FAIL("dp_connect: already connected");
  CK(cudaSetDevice(h->cfg.device));
  const int64_t me = static_cast<int64_t>(getpid());
  bool shared_gpu = false;
  for (int r = 0; r < h->dp_world; ++r)
    for (int q = 0; q < r; ++q) {
      DpBlob a, c;
      memcpy(&a, static_cast<const char*>(blobs) + static_cast<size_t>(r) * WGAN_DP_BLOB_BYTES, sizeof a);
      memcpy(&c, static_cast<const char*>(blobs) + static_cast<size_t>(q) * WGAN_DP_BLOB_BYTES, sizeof c);
      if (memcmp(a.uuid, c.uuid, 16) == 0) shared_gpu = true;
    }
  h->dp_early_trigger = !shared_gpu && getenv("WGAN_B200_DP_NO_EARLY") == nullptr;
  for (int r = 0; r < h->dp_world; ++r) {
    DpBlob b;
    memcpy(&b, static_cast<const char*>(blobs) + static_cast<size_t>(r) * WGAN_DP_BLOB_BYTES, sizeof b);
    if (b.magic != kDpMagic || b.rank != r || b.world != h->dp_world)
      FAIL("dp_connect: blob %d is not rank %d of a %d-rank group", r, r, h->dp_world);
    if (b.win_bytes != h->dp_win_bytes || b.arena_n[0] != h->arena_n[0] || b.arena_n[1] != h->arena_n[1] ||
        b.blocks != h->dp_blocks)
      FAIL("dp_connect: rank %d was created with a different model configuration", r);
    if (r == h->dp_rank) continue;
    if (b.pid == me) {
      // same process (one host thread per GPU, or several handles on one GPU): the pointer is directly usable
      if (b.device != h->cfg.device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, h->cfg.device, b.device));
        if (!can) FAIL("dp_connect: device %d cannot access device %d", h->cfg.device, b.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        (void)cudaGetLastError();
      }
      h->dp_peer[r] = reinterpret_cast<float*>(b.ptr);
    } else {
      void* mapped = nullptr;
      CK(cudaIpcOpenMemHandle(&mapped, b.ipc, cudaIpcMemLazyEnablePeerAccess));
      h->dp_peer[r] = static_cast<float*>(mapped);
      h->dp_peer_ipc[r] = true;
    }
  }
  h->dp_connected = true;
  return 0;
}

static int dp_exchange_range(wgan_handle* h, int net, size_t off, size_t n, cudaStream_t s) {
  if ((off & 3) != 0 || (n & 3) != 0) FAIL("allreduce: range [%zu, +%zu) is not a multiple of 4 floats", off, n);
  if (n == 0) return 0;
  DpParams p;
  memset(&p, 0, sizeof p);
  for (int r = 0; r < h->dp_world; ++r) p.peer[r] = h->dp_peer[r];
  p.rank = h->dp_rank; p.world = h->dp_world;
  p.grad_off = h->dp_grad_off[net] + off;
  p.flags_off = h->dp_flags_off;
  p.stage_off = h->dp_stage_off;
  p.stage_stride = h->dp_stage_stride;
  p.n4 = n / 4;
  p.early_trigger = h->dp_early_trigger ? 1 : 0;
  static const int dp_threads = [] {
    const char* e = getenv("WGAN_B200_DP_THREADS");
    const int t = e ? atoi(e) : DP_THREADS;
    return (t >= 32 && t <= DP_THREADS && t % 32 == 0) ? t : DP_THREADS;
  }();
  KLAUNCH(dp_allreduce_kernel, h->dp_blocks, dp_threads, 0, s, p);
  return post_launch(h, "dp_allreduce_kernel");
}

int wgan_allreduce_grads(wgan_handle* h, int net, wgan_stream_t stream) {
  if (!h) return -2;
  if (net < 0 || net > 1) FAIL("allreduce_grads: bad net");
  if (h->dp_world <= 1) return 0;
  if (!h->dp_connected) FAIL("allreduce_grads: call wgan_dp_connect first");
  CK(cudaSetDevice(h->cfg.device));
  return dp_exchange_range(h, net, 0, h->arena_n[net], static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// Batched sampler: out = G(z)   (generator forward only, [W:217])
// ---------------------------------------------------------------------------------------------
int wgan_sample(wgan_handle* h, const float* z_dev, int ldz, float* out_dev, int ld_out, int rows,
                wgan_stream_t stream) {
  if (!h) return -2;
  if (!z_dev || !out_dev) FAIL("sample: null argument");
  TRY(check_rows(h, rows, rows));
  CK(cudaSetDevice(h->cfg.device));
  invalidate_workspace(h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float* z; int lz;
  TRY(stage_input(h, z_dev, ldz, rows, h->Z, h->zbuf, h->Zl, &z, &lz, s));
  const bool direct = (reinterpret_cast<uintptr_t>(out_dev) & 15) == 0 && (ld_out & 3) == 0 && ld_out >= h->G;
  if (direct) return generator_forward(h, z, lz, rows, out_dev, ld_out, s);
  TRY(generator_forward(h, z, lz, rows, h->sample_tmp, h->Gl, s));
  CK(cudaMemcpy2DAsync(out_dev, static_cast<size_t>(ld_out) * 4, h->sample_tmp, static_cast<size_t>(h->Gl) * 4,
                       static_cast<size_t>(h->G) * 4, rows, cudaMemcpyDeviceToDevice, s));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Critic forward only: d = D(x)   [W:103-124]
// ---------------------------------------------------------------------------------------------
int wgan_critic_forward(wgan_handle* h, const float* x_dev, int ldx, float* d_out_dev, int rows,
                        wgan_stream_t stream) {
  if (!h) return -2;
  if (!x_dev || !d_out_dev) FAIL("critic_forward: null argument");
  TRY(check_rows(h, rows, rows));
  CK(cudaSetDevice(h->cfg.device));
  invalidate_workspace(h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Hd = h->Hd, Hdl = h->Hdl;
  const float a = h->cfg.alpha;
  const float* x; int lx;
  TRY(stage_input(h, x_dev, ldx, rows, h->G, h->xcat, h->Gl, &x, &lx, s));
  TRY(gemm1(h, false, true, x, lx, h->P(T_W1), h->td[T_W1].ld, h->h1, Hdl, rows, Hd, h->G,
            epi_bias_lrelu(h->P(T_B1), a), s));
  TRY(gemm1(h, false, true, h->h1, Hdl, h->P(T_W2), h->td[T_W2].ld, h->h2, Hdl, rows, Hd, Hd,
            epi_bias_lrelu(h->P(T_B2), a), s));
  KLAUNCH(critic_head_kernel, ew_grid(static_cast<size_t>(rows) * 32, 256, h->num_sms), 256, 0, s, 
      h->h2, Hdl, rows, Hd, h->P(T_W3), h->P(T_B3), d_out_dev, h->d2, rows, 0.f, 0.f, 0.f, a);
  return post_launch(h, "critic_head_kernel");
}

// ---------------------------------------------------------------------------------------------
// Latent-vector fit step (build-defined workload)
// ---------------------------------------------------------------------------------------------
int wgan_latent_fit_step(wgan_handle* h, float* z_dev, float* m_dev, float* v_dev, float* powers_dev,
                         const float* x_dev, int ldx, float* loss_dev, int rows, float learn_rate,
                         wgan_stream_t stream) {
  if (!h) return -2;
  if (!z_dev || !m_dev || !v_dev || !powers_dev || !x_dev || !loss_dev) FAIL("latent_fit_step: null argument");
  if (h->Z % 4 != 0) FAIL("latent_fit_step needs noise_input_size %% 4 == 0 (got %d)", h->Z);
  if ((reinterpret_cast<uintptr_t>(z_dev) & 15) != 0) FAIL("latent_fit_step: z must be 16-byte aligned");
  TRY(check_rows(h, rows, rows));
  CK(cudaSetDevice(h->cfg.device));
  invalidate_workspace(h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Bl = rows, G = h->G, Gl = h->Gl, Hg = h->Hg, Hgl = h->Hgl, Z = h->Z;
  const float a = h->cfg.alpha;
  float* xt = h->xcat;
  float* d3 = h->xcat + static_cast<size_t>(Bl) * Gl;
  TRY(generator_forward(h, z_dev, Z, Bl, xt, Gl, s));
  KLAUNCH(latent_residual_kernel, Bl, 256, 0, s, xt, Gl, x_dev, ldx, d3, Gl, loss_dev, G, a);
  TRY(post_launch(h, "latent_residual_kernel"));
  const float* Wg3 = h->P(T_WG3); const float* Wg2 = h->P(T_WG2); const float* Wg1 = h->P(T_WG1);
  TRY(gemm1(h, false, false, d3, Gl, Wg3, h->td[T_WG3].ld, h->gd2, Hgl, Bl, Hg, G, epi_mask(h->gh2, Hgl, a), s));
  TRY(gemm1(h, false, false, h->gd2, Hgl, Wg2, h->td[T_WG2].ld, h->gd1, Hgl, Bl, Hg, Hg, epi_mask(h->gh1, Hgl, a), s));
  TRY(gemm1(h, false, false, h->gd1, Hgl, Wg1, h->td[T_WG1].ld, h->dz, Z, Bl, Z, Hg, epi_none(), s));
  const size_t n4 = static_cast<size_t>(Bl) * Z / 4;
  KLAUNCH(adam_kernel, ew_grid(n4, 256, h->num_sms), 256, 0, s, z_dev, h->dz, m_dev, v_dev, n4, powers_dev, learn_rate,
                                                          h->cfg.beta1, h->cfg.beta2, h->cfg.epsilon);
  return post_launch(h, "adam_kernel");
}

// ---------------------------------------------------------------------------------------------
// Device input pipeline
// ---------------------------------------------------------------------------------------------
static int rng_launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string(what) + ": " + cudaGetErrorString(e); return -3; }
  g_pipeline_launches.fetch_add(1);
  return 0;
}
static int upload_poisson_cdf_once() {
  static thread_local int done_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (done_dev == dev) return 0;
  float cdf[12];
  double acc = 0.0, term = exp(-1.0);
  for (int k = 0; k < 12; ++k) { if (k > 0) term /= k; acc += term; cdf[k] = static_cast<float>(acc); }
  if (cudaMemcpyToSymbol(c_poisson1_cdf, cdf, sizeof cdf) != cudaSuccess) return -1;
  done_dev = dev;
  return 0;
}

int wgan_rng_noise_prior(float* out_dev, int ld, int64_t row0, int rows, int dim, float sigma,
                         uint64_t seed, uint32_t call_id, const uint32_t* call_id_dev, wgan_stream_t stream) {
  if (!out_dev || rows <= 0 || dim <= 0 || ld < dim) { g_create_error = "rng_noise_prior: bad argument"; return -2; }
  if (upload_poisson_cdf_once() != 0) { g_create_error = "rng_noise_prior: constant upload failed"; return -1; }
  const size_t n = static_cast<size_t>(rows) * dim;
  int grid = static_cast<int>((n + 255) / 256 > kB200Sms * 16 ? kB200Sms * 16 : (n + 255) / 256);
  KLAUNCH(rng_noise_prior_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), out_dev, ld, row0, rows, dim, sigma,
                                                                               seed, call_id, 1u, call_id_dev);
  return rng_launch_check("rng_noise_prior_kernel");
}

int wgan_rng_uniform(float* out_dev, int64_t row0, int rows, uint64_t seed, uint32_t call_id,
                     const uint32_t* call_id_dev, wgan_stream_t stream) {
  if (!out_dev || rows <= 0) { g_create_error = "rng_uniform: bad argument"; return -2; }
  KLAUNCH(rng_uniform_kernel, cdiv(rows, 256), 256, 0, static_cast<cudaStream_t>(stream), out_dev, row0, rows, seed,
                                                                                     call_id, 2u, call_id_dev);
  return rng_launch_check("rng_uniform_kernel");
}

int wgan_rng_indices(int64_t* out_dev, int64_t row0, int rows, uint32_t n_cells, uint64_t seed,
                     uint32_t call_id, const uint32_t* call_id_dev, wgan_stream_t stream) {
  if (!out_dev || rows <= 0 || n_cells == 0) { g_create_error = "rng_indices: bad argument"; return -2; }
  KLAUNCH(rng_indices_kernel, cdiv(rows, 256), 256, 0, static_cast<cudaStream_t>(stream), out_dev, row0, rows, n_cells,
                                                                                     seed, call_id, 3u, call_id_dev);
  return rng_launch_check("rng_indices_kernel");
}

int wgan_counter_add(uint32_t* counter_dev, uint32_t inc, wgan_stream_t stream) {
  if (!counter_dev) { g_create_error = "counter_add: null argument"; return -2; }
  KLAUNCH(counter_add_kernel, 1, 32, 0, static_cast<cudaStream_t>(stream), counter_dev, inc);
  return rng_launch_check("counter_add_kernel");
}

int wgan_gather_rows(const float* data_dev, int ld_data, const int64_t* idx_dev, float* out_dev,
                     int ld_out, int rows, int cols, wgan_stream_t stream) {
  if (!data_dev || !idx_dev || !out_dev || rows <= 0 || cols <= 0) { g_create_error = "gather_rows: bad argument"; return -2; }
  if ((ld_data & 3) || (ld_out & 3) || (reinterpret_cast<uintptr_t>(data_dev) & 15) || (reinterpret_cast<uintptr_t>(out_dev) & 15)) {
    g_create_error = "gather_rows: data and out must be 16-byte aligned with ld % 4 == 0";
    return -2;
  }
  const int cols4 = round4(cols) / 4;
  if (round4(cols) > ld_data || round4(cols) > ld_out) { g_create_error = "gather_rows: ld smaller than padded cols"; return -2; }
  const size_t n = static_cast<size_t>(rows) * cols4;
  int grid = static_cast<int>((n + 255) / 256 > kB200Sms * 16 ? kB200Sms * 16 : (n + 255) / 256);
  KLAUNCH(gather_rows_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), data_dev, ld_data, idx_dev, out_dev,
                                                                           ld_out, rows, cols4);
  return rng_launch_check("gather_rows_kernel");
}

int wgan_draw_minibatch(const float* data_dev, int ld_data, uint32_t n_cells, float* x_out_dev, int ld_out, int cols,
                        float* z_out_dev, int ldz, int dim, float sigma, float* eps_out_dev, int64_t* idx_out_dev,
                        int64_t row0, int rows, uint64_t seed, uint32_t call_id, const uint32_t* call_id_dev,
                        wgan_stream_t stream) {
  if (!data_dev || !x_out_dev || !z_out_dev || rows <= 0 || cols <= 0 || dim <= 0 || n_cells == 0 || ldz < dim) {
    g_create_error = "draw_minibatch: bad argument";
    return -2;
  }
  if ((ld_data & 3) || (ld_out & 3) || (reinterpret_cast<uintptr_t>(data_dev) & 15) ||
      (reinterpret_cast<uintptr_t>(x_out_dev) & 15) || round4(cols) > ld_data || round4(cols) > ld_out) {
    g_create_error = "draw_minibatch: data and x_out must be 16-byte aligned with ld % 4 == 0 and ld >= padded cols";
    return -2;
  }
  if (upload_poisson_cdf_once() != 0) { g_create_error = "draw_minibatch: constant upload failed"; return -1; }
  // rows move through the bulk-copy engine (two padded rows of shared memory per block) when they fit
  const uint32_t row_bytes = static_cast<uint32_t>(round4(cols)) * 4u;
  static const bool no_bulk = getenv("WGAN_B200_NO_BULK_DRAW") != nullptr;
  if (!no_bulk && 2u * row_bytes <= 160u * 1024u) {
    static bool attr_set[16] = {};
    int dev = 0;
    (void)cudaGetDevice(&dev);
    if (!attr_set[dev & 15]) {
      if (cudaFuncSetAttribute(draw_minibatch_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) !=
          cudaSuccess) { g_create_error = "draw_minibatch: cudaFuncSetAttribute failed"; return -1; }
      attr_set[dev & 15] = true;
    }
    const int per_sm = static_cast<int>((200u * 1024u) / (2u * row_bytes + 1024u));
    const int cap = kB200Sms * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    const int grid = rows < cap ? rows : cap;
    KLAUNCH(draw_minibatch_bulk_kernel, grid, 256, 2u * row_bytes, static_cast<cudaStream_t>(stream), data_dev, ld_data,
            n_cells, x_out_dev, ld_out, row_bytes, z_out_dev, ldz, dim, sigma, eps_out_dev, idx_out_dev,
            static_cast<long>(row0), rows, seed, call_id, call_id_dev);
    return rng_launch_check("draw_minibatch_bulk_kernel");
  }
  const int grid = rows < kB200Sms * 16 ? rows : kB200Sms * 16;
  KLAUNCH(draw_minibatch_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), data_dev, ld_data, n_cells, x_out_dev,
          ld_out, round4(cols) / 4, z_out_dev, ldz, dim, sigma, eps_out_dev, idx_out_dev, static_cast<long>(row0), rows,
          seed, call_id, call_id_dev);
  return rng_launch_check("draw_minibatch_kernel");
}

// ---------------------------------------------------------------------------------------------
// Test hooks
// ---------------------------------------------------------------------------------------------
int wgan_gemm(wgan_handle* h, int a_mn, int b_mn, const float* A, int lda, const float* B, int ldb,
              float* D, int ldd, int M, int N, int K, const float* bias, int act, const float* aux,
              int ld_aux, const float* row_scale, float* row_sumsq, int* sumsq_tiles,
              int force_splits, wgan_stream_t stream) {
  if (!h) return -2;
  CK(cudaSetDevice(h->cfg.device));
  EpiParams e = epi_none();
  e.bias = bias; e.act = act; e.aux = aux; e.ld_aux = ld_aux; e.row_scale = row_scale;
  e.row_sumsq = row_sumsq; e.alpha = h->cfg.alpha;
  Part p{A, lda, B, ldb, K};
  return run_gemm(h, a_mn != 0, b_mn != 0, &p, 1, D, ldd, M, N, e, sumsq_tiles, force_splits,
                  static_cast<cudaStream_t>(stream));
}

int wgan_get_buffer(wgan_handle* h, const char* name, float** dev_ptr, int* rows, int* cols, int* ld) {
  if (!h) return -2;
  if (!name || !dev_ptr || !rows || !cols || !ld) FAIL("get_buffer: null argument");
  const int Bl = h->last_gx_rows;
  std::string n(name);
  if (n == "x_tilde") { *dev_ptr = h->xcat; *rows = h->R; *cols = h->G; *ld = h->Gl; }
  else if (n == "gx") {
    if (h->cfg.lambda_gp != 0.0f && h->cfg.gp_form == 1) FAIL("get_buffer: the Gram form never materialises gx");
    *dev_ptr = h->xcat + static_cast<size_t>(Bl) * h->Gl; *rows = Bl; *cols = h->G; *ld = h->Gl; }
  else if (n == "xcat") { *dev_ptr = h->xcat; *rows = 3 * h->R; *cols = h->G; *ld = h->Gl; }
  else if (n == "norms") { *dev_ptr = h->norms; *rows = 1; *cols = Bl; *ld = Bl; }
  else if (n == "d") { *dev_ptr = h->dout; *rows = 1; *cols = 3 * h->R; *ld = 3 * h->R; }
  else if (n == "h1") { *dev_ptr = h->h1; *rows = 3 * h->R; *cols = h->Hd; *ld = h->Hdl; }
  else if (n == "h2") { *dev_ptr = h->h2; *rows = 3 * h->R; *cols = h->Hd; *ld = h->Hdl; }
  else FAIL("get_buffer: unknown buffer '%s'", name);
  return 0;
}

}  // extern "C"

#ifdef WGAN_DIAG_STAMPS
// diagnostic build only (scripts/diag_build.sh): phase stamps [block][8] of the pair kernel's last launch
extern "C" int wgan_diag_read_stamps(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, wg::g_diag_stamps, sizeof(unsigned long long) * 160 * 8) == cudaSuccess ? 0 : -1;
}
#endif
