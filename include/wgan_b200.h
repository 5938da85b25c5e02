/* wgan_b200.h — C ABI of the B200-native WGAN(-GP) hot path.
 *
 * Drop-in boundary for the hot path of scripts/WGAN-GP_minimal.py.  The reference has no
 * FFI/plugin interface (the path is inlined in a TF1 script); each entry point below cites
 * the reference lines whose work it replaces:
 *   [W:n] = scripts/WGAN-GP_minimal.py:n      [A:n] = scripts/Adam_prediction.py:n
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative code on failure (message via wgan_last_error); no exceptions cross the ABI;
 * step calls are asynchronous on the caller's CUDA stream and never synchronise the host;
 * the caller owns every buffer it passes in; the handle owns parameters, Adam state and
 * workspace; a handle is bound to one GPU and is not thread-safe.
 * All matrices are row-major FP32, one cell (sample) per row, `ld*` = leading dimension
 * in floats.  Caller buffers are only ever read inside their logical extent [rows, cols]: a latent matrix that is
 * 16-byte aligned with ld % 4 == 0 is consumed in place by TMA, anything else -- and always the real-cell matrix
 * x, which the weight gradient also reads column-chunked -- is copied into the handle's stacked activation buffer
 * first (no copy when x already IS that slot: wgan_get_buffer "xcat", WGANTrainer.real_slot()).
 * GEMMs run as TF32 tcgen05 tensor-core kernels (gemm_mode 0) or as FP32 CUDA-core kernels (gemm_mode 1,
 * verification only).
 *
 * Data parallelism lives behind this ABI: every rank creates its handle with dp_rank / dp_world,
 * exports a small blob (wgan_dp_export), the host exchanges the blobs by any means (MPI, torch.distributed, a
 * file) and hands all of them to wgan_dp_connect; wgan_allreduce_grads then sums a gradient arena (+ the loss
 * scalars in its tail) over the ranks in place with the engine's own kernel over NVLink peer memory.
 */
#ifndef WGAN_B200_H
#define WGAN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wgan_handle wgan_handle;
typedef void* wgan_stream_t; /* cudaStream_t */

/* Module-level constants of the reference script [W:9-24] plus the build-defined
 * gradient-penalty knobs (lambda_gp = 0, n_critic = 1 is reference-literal). */
typedef struct wgan_config {
  int32_t noise_input_size;   /* [W:13] 100  */
  int32_t inflate_to_size;    /* [W:14] 600  */
  int32_t gex_size;           /* [W:15] 6605 */
  int32_t disc_internal_size; /* [W:17] 200  */
  int32_t max_batch;          /* largest per-GPU batch (rows) any call will use */
  int32_t device;             /* CUDA device ordinal */
  int32_t gemm_mode;          /* 0 = TF32 tcgen05, 1 = FP32 CUDA cores (verification) */
  int32_t gp_form;            /* gradient-penalty evaluation: 0 = explicit chain (x^ materialised),
                                 1 = Gram form (same value in real arithmetic: layer-1 pre-activation blend and
                                 K = W1^T W1, SURVEY 8d shortcuts i + ii); ignored when lambda_gp == 0 */
  float learn_rate;           /* [W:22] 1e-5 */
  float beta1;                /* [W:154-155] 0.9 */
  float beta2;                /* [W:154-155] 0.999 */
  float epsilon;              /* [A:38] 1e-8 */
  float alpha;                /* tf.nn.leaky_relu default 0.2 [W:69] */
  float lambda_gp;            /* 0 => no penalty term (reference-literal [W:137]) */
  int32_t dp_rank;            /* data-parallel rank of this handle, 0 <= dp_rank < dp_world (batch shard) */
  int32_t dp_world;           /* number of data-parallel ranks; 0 or 1 = single GPU */
} wgan_config;

/* Data-parallel rendezvous blob: opaque to the caller, produced by wgan_dp_export on every rank, gathered in rank
 * order and passed to wgan_dp_connect on every rank. */
#define WGAN_DP_MAX_WORLD 16
#define WGAN_DP_BLOB_BYTES 192

enum { WGAN_NET_GENERATOR = 0, WGAN_NET_CRITIC = 1 };
enum { WGAN_KIND_PARAM = 0, WGAN_KIND_GRAD = 1, WGAN_KIND_M = 2, WGAN_KIND_V = 3 };
/* Scalars in the tail of each net's gradient arena (globally summed by the DP all-reduce). */
enum { WGAN_S_DFAKE = 0, WGAN_S_DREAL = 1, WGAN_S_GP = 2, WGAN_S_SPARE = 3 };

typedef struct wgan_tensor_info {
  char name[64];   /* TF variable name, e.g. "discriminator/disc_dense1/kernel" [W:63-122] */
  int32_t net;     /* WGAN_NET_* */
  int32_t rows;    /* kernel: in, bias: 1 */
  int32_t cols;    /* kernel: out, bias: out */
  int32_t ld;      /* device leading dimension (floats) */
  int64_t offset;  /* float offset inside the net's arena */
} wgan_tensor_info;

/* ---- lifecycle (replaces graph construction + tf.Session [W:59-172]) ---- */
int wgan_create(const wgan_config* cfg, wgan_handle** out);
void wgan_destroy(wgan_handle* h);
const char* wgan_last_error(const wgan_handle* h); /* h may be NULL: last create error */

/* ---- state I/O in TF variable order [W:140-142]: 12 tensors, generator first ---- */
int wgan_tensor_count(void);
int wgan_get_tensor_info(const wgan_handle* h, int index, wgan_tensor_info* out);
/* Host <-> device copies of one tensor in its logical (unpadded) shape; synchronous. */
int wgan_get_tensor(wgan_handle* h, int kind, int index, float* host_dst);
int wgan_set_tensor(wgan_handle* h, int kind, int index, const float* host_src);
/* beta1_power / beta2_power of one optimizer [A:123-128]; synchronous. */
int wgan_get_adam_powers(wgan_handle* h, int net, float* beta1_power, float* beta2_power);
int wgan_set_adam_powers(wgan_handle* h, int net, float beta1_power, float beta2_power);
/* Flat device arena of one net (for the data-parallel gradient all-reduce and checkpoints).
 * n_floats includes a 4-float scalar tail (WGAN_S_*) in the GRAD arena. */
int wgan_get_arena(wgan_handle* h, int net, int kind, float** dev_ptr, size_t* n_floats);
/* Copies the 4 scalars of a net's grad arena to the host; synchronises `stream`. */
int wgan_get_scalars(wgan_handle* h, int net, float* host4, wgan_stream_t stream);

/* ---- training steps ---- */
/* Critic gradients: d(obj_d [+ GP])/d(d_params), generator constant  [W:137,141,154].
 * x [rows, gex] real cells, z [rows, noise] latent, eps [rows] interpolation weights
 * (ignored when lambda_gp == 0).  global_batch = total rows over all data-parallel ranks
 * (the 1/B of the means).  Fills the critic GRAD arena + scalars
 * {mean-part D_fake, mean-part D_real, GP}.  */
/* prepared_token: 0, or the token wgan_critic_prepare returned for exactly these inputs (see below). */
int wgan_critic_grads(wgan_handle* h, const float* x_dev, int ldx, const float* z_dev, int ldz,
                      const float* eps_dev, int rows, int global_batch, uint64_t prepared_token,
                      wgan_stream_t stream);
/* Optional: run the part of the critic step that does not depend on the critic's parameters (input
 * staging, x~ = G(z), interpolates) ahead of time, e.g. while the previous update's gradients are being
 * all-reduced.  *token_out identifies the prepared activations: pass it to the wgan_critic_grads call for the
 * same x / z / eps / rows, whose contents must not have changed in between.  Any other entry point that reuses
 * the activation workspace (sample, critic_forward, latent fit, another prepare, a generator step) invalidates
 * the token; a stale or zero token is never an error, the work is simply redone. */
int wgan_critic_prepare(wgan_handle* h, const float* x_dev, int ldx, const float* z_dev, int ldz,
                        const float* eps_dev, int rows, wgan_stream_t stream, uint64_t* token_out);
/* Adam update of the critic from its GRAD arena [A:140-151], then beta powers [A:214-225]. */
int wgan_critic_apply(wgan_handle* h, wgan_stream_t stream);
/* Generator gradients: d(obj_g)/d(g_params) through the critic [W:138,142,155];
 * scalar WGAN_S_DFAKE = mean-part of D(G(z)) (obj_g = -that). */
int wgan_generator_grads(wgan_handle* h, const float* z_dev, int ldz, int rows, int global_batch,
                         uint64_t prepared_token, wgan_stream_t stream);
/* Optional: stage z and run x~ = G(z) ahead of wgan_generator_grads(same z, rows, token) (overlap with the last
 * critic all-reduce in data-parallel runs); token rules as for wgan_critic_prepare. */
int wgan_generator_prepare(wgan_handle* h, const float* z_dev, int ldz, int rows, wgan_stream_t stream,
                           uint64_t* token_out);
int wgan_generator_apply(wgan_handle* h, wgan_stream_t stream);

/* ---- data parallelism: batch-sharded ranks, one SUM all-reduce of a net's gradient arena per optimizer step
 *      (the step it sits inside: minimize = gradients -> [all-reduce] -> apply_adam, [W:154-155],
 *      [A:140-151]).  The exchange is the engine's own kernel over NVLink peer memory (no NCCL), all cross-GPU
 *      traffic being posted writes: every rank pushes its contribution to slice r into rank r's staging rows,
 *      rank r adds the contributions in rank order and pushes the sum into every arena, so all replicas hold
 *      bit-identical gradients and stay bit-identical after Adam. ---- */
/* Writes this rank's WGAN_DP_BLOB_BYTES rendezvous blob. */
int wgan_dp_export(wgan_handle* h, void* blob_out);
/* blobs: dp_world * WGAN_DP_BLOB_BYTES bytes, blob of rank i at offset i * WGAN_DP_BLOB_BYTES.  Maps the peers'
 * gradient windows (cudaIpc across processes, direct peer access inside one process).  Collective: every rank
 * must call it before the first wgan_allreduce_grads. */
int wgan_dp_connect(wgan_handle* h, const void* blobs);
/* In-place SUM over all ranks of net's GRAD arena (incl. the WGAN_S_* scalar tail); asynchronous on `stream`,
 * collective (every rank calls it for the same net in the same order).  A no-op when dp_world <= 1. */
int wgan_allreduce_grads(wgan_handle* h, int net, wgan_stream_t stream);
/* wgan_generator_grads + the exchange of the generator's gradient arena in one call, overlapped: as soon as the
 * output layer's gradients (kernel + bias, 90 % of the arena) and the loss scalar are final they are summed over the
 * ranks on an internal high-priority stream while the rest of the backward pass [W:155] runs on `stream`; the
 * remaining tensors follow at the end.  Same results as wgan_generator_grads followed by
 * wgan_allreduce_grads(WGAN_NET_GENERATOR) bit for bit; collective; equal to wgan_generator_grads when
 * dp_world <= 1.  Capturable into a CUDA graph (the internal stream forks from and joins `stream`). */
int wgan_generator_grads_exchange(wgan_handle* h, const float* z_dev, int ldz, int rows, int global_batch,
                                  uint64_t prepared_token, wgan_stream_t stream);

/* ---- sampling: G(z) for a batch of cells (replaces the batch-1 loop [W:217-221]) ---- */
int wgan_sample(wgan_handle* h, const float* z_dev, int ldz, float* out_dev, int ld_out, int rows,
                wgan_stream_t stream);

/* ---- critic forward only: d_out[rows] = D(x) (replaces fetching D_real / D_fake [W:129-130]) ---- */
int wgan_critic_forward(wgan_handle* h, const float* x_dev, int ldx, float* d_out_dev, int rows,
                        wgan_stream_t stream);

/* ---- latent-vector fit (build-defined): one Adam step on z minimising
 *      sum_g (G(z) - x)^2 per cell with G frozen.  z_dev [rows, noise] (ld = noise) is
 *      updated in place; m_dev/v_dev same shape; powers_dev = {beta1_power, beta2_power};
 *      loss_dev [rows] receives the per-cell loss before the update. ---- */
int wgan_latent_fit_step(wgan_handle* h, float* z_dev, float* m_dev, float* v_dev, float* powers_dev,
                         const float* x_dev, int ldx, float* loss_dev, int rows, float learn_rate,
                         wgan_stream_t stream);

/* ---- device input pipeline (replaces host NumPy [W:91-94,188-193]); Philox4x32-10 ----
 * The draw is a pure function of (seed, call id, global row, column).  The effective call id is
 * call_id + *call_id_dev (call_id_dev may be NULL): a device-resident counter, advanced by
 * wgan_counter_add, lets a captured CUDA graph draw fresh numbers on every replay. */
int wgan_rng_noise_prior(float* out_dev, int ld, int64_t row0, int rows, int dim, float sigma,
                         uint64_t seed, uint32_t call_id, const uint32_t* call_id_dev, wgan_stream_t stream);
int wgan_rng_uniform(float* out_dev, int64_t row0, int rows, uint64_t seed, uint32_t call_id,
                     const uint32_t* call_id_dev, wgan_stream_t stream);
int wgan_rng_indices(int64_t* out_dev, int64_t row0, int rows, uint32_t n_cells, uint64_t seed,
                     uint32_t call_id, const uint32_t* call_id_dev, wgan_stream_t stream);
int wgan_counter_add(uint32_t* counter_dev, uint32_t inc, wgan_stream_t stream);
int wgan_gather_rows(const float* data_dev, int ld_data, const int64_t* idx_dev, float* out_dev,
                     int ld_out, int rows, int cols, wgan_stream_t stream);

/* Everything one critic update draws, in ONE launch (replaces [W:188-193] per update): for minibatch row r
 * (global row row0 + r) x_out[r,:] = data[idx_r,:] with idx_r the wgan_rng_indices draw, z_out[r,:] the
 * wgan_rng_noise_prior draw and eps_out[r] the wgan_rng_uniform draw of the same (seed, call id) -- bit-identical
 * to calling those three and wgan_gather_rows.  eps_out_dev and idx_out_dev may be NULL.  x_out_dev is normally
 * the real-cell slot of the stacked activation buffer (wgan_get_buffer "xcat"). */
int wgan_draw_minibatch(const float* data_dev, int ld_data, uint32_t n_cells, float* x_out_dev, int ld_out, int cols,
                        float* z_out_dev, int ldz, int dim, float sigma, float* eps_out_dev, int64_t* idx_out_dev,
                        int64_t row0, int rows, uint64_t seed, uint32_t call_id, const uint32_t* call_id_dev,
                        wgan_stream_t stream);

/* ---- introspection / test hooks ---- */
/* D[M,N] = epi(A * B^T-ish): a_mn / b_mn select MN-major operands (see csrc/gemm_sm100.cuh).
 * act: 0 none, 1 LeakyReLU(alpha).  Any of bias/aux/row_scale/row_sumsq may be NULL.
 * force_splits == 0 lets the engine pick (split-K or stream-K), > 0 forces that many split-K slabs, < 0 forces
 * stream-K.  *sumsq_tiles receives the number of row_sumsq tiles.
 * Unlike the step entry points (which stage or own their operands) this hook hands A / B to TMA as they are: an
 * MN-major operand is read in whole 32-float chunks, so its LAST row must be followed by 124 readable bytes
 * (one spare row is enough). */
int wgan_gemm(wgan_handle* h, int a_mn, int b_mn, const float* A, int lda, const float* B, int ldb,
              float* D, int ldd, int M, int N, int K, const float* bias, int act, const float* aux,
              int ld_aux, const float* row_scale, float* row_sumsq, int* sumsq_tiles,
              int force_splits, wgan_stream_t stream);
/* Named internal buffers ("norms", "gx", "x_tilde", ...) for parity tests; NULL if unknown. */
int wgan_get_buffer(wgan_handle* h, const char* name, float** dev_ptr, int* rows, int* cols, int* ld);
/* Number of kernels launched by this handle since creation (bench's gpu_launches). */
int64_t wgan_launch_count(const wgan_handle* h);
int wgan_padded_ld(int cols);

#ifdef __cplusplus
}
#endif
#endif /* WGAN_B200_H */
