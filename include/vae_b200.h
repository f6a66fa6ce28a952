/*
 * vae_b200.h -- C ABI of the B200-native VAE ELBO train-step library (libvae_b200.so).
 *
 * The reference (virajmehta/vae-training) has no FFI / plugin interface: it is pure
 * Python on JAX/Flax.  The drop-in boundary is therefore the set of Python call
 * seams listed in SURVEY.md 8(b); each entry point below names the reference
 * call site it replaces (paths are into the reference tree).  A jax.ffi handler
 * would bind exactly these symbols.
 *
 * Conventions
 *   - Every pointer documented "device" is a CUDA device pointer owned by the
 *     caller, 16-byte aligned.  The callee never allocates, frees or retains a
 *     pointer past return.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*; NULL is the
 *     legacy default stream).  No entry point synchronises the device.
 *   - Return value: 0 on success, negative VAE_ERR_* otherwise; the message is
 *     available from vae_last_error() (thread-local).
 *   - No global mutable state besides CUDA's per-device function attributes, which are only ever
 *     set to the same constants (maximum dynamic shared memory): re-entrant across host threads /
 *     streams.  Launch-shaping environment knobs (VAE_SMALL_*, VAE_MLP_CLUSTER, VAE_TC_CLUSTER) are
 *     read when a pack is prepared / a GEMM is launched.
 *   - There is no CPU fallback anywhere behind this header.
 *
 * Tensors are float32 row-major; PRNG keys are uint32[2] (jax.random legacy
 * threefry2x32 layout).
 */
#ifndef VAE_B200_H_
#define VAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAE_B200_ABI_VERSION 4
#define VAE_MAX_HIDDEN 8

enum {
  VAE_OK = 0,
  VAE_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  VAE_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  VAE_ERR_WORKSPACE = -3, /* workspace too small */
  VAE_ERR_UNSUPPORTED = -4
};

/* Model description.  Mirrors the keyword arguments VAE.partial(...) is given at
 * vae.py:57-59 plus the optimiser hyper-parameters of vae.py:113 and the batch
 * size of run.py:14. */
typedef struct VaeDesc {
  int32_t batch;                      /* B          run.py:14                       */
  int32_t data_dim;                   /* D          vae.py:51 (prod(dataset.shape)) */
  int32_t latent_dim;                 /* L          run.py:22                       */
  int32_t n_enc_hidden;               /* len(encoder_layer_sizes) - 1   vae.py:53   */
  int32_t enc_hidden[VAE_MAX_HIDDEN];
  int32_t n_dec_hidden;               /* len(decoder_layer_sizes) - 1   vae.py:54   */
  int32_t dec_hidden[VAE_MAX_HIDDEN];
  int32_t sig_decoder;                /* dataset_name == "sigmoid"  networks.py:75  */
  int32_t tunable_decoder_var;        /* -tdv                       networks.py:70  */
  int32_t reserved_;                  /* keeps the doubles 8-byte aligned           */
  /* Hyper-parameters are doubles because the reference holds them as Python floats:
   * legacy flax evaluates (1 - beta) in double before rounding to float32, and the
   * kernels reproduce that: b = (float)beta, 1-b = (float)(1.0 - beta),
   * bias correction 1 - powf((float)beta, t). */
  double epsilon0;                    /* --epsilon                  run.py:32       */
  double lr;                          /* -lr                        run.py:15       */
  double beta1, beta2, adam_eps;      /* flax.optim.Adam defaults .9 .999 1e-8      */
} VaeDesc;

/* Flat parameter layout (params, adam_m, adam_v and gradients all use it):
 *   for net in Encoder, Decoder, [SigDecoder if sig_decoder]:
 *     for layer i: kernel (in, out) row-major, then bias (out)
 *   epsilon_p (L)                    networks.py:69
 *   epsilon   (1) if tunable_decoder_var   networks.py:71
 */
int64_t vae_param_count(const VaeDesc* desc);

/* Synthetic generator description (run.py:46-54, datasets.py:55-98,163-228,230-279). */
enum { VAE_DATA_SPHERE = 0, VAE_DATA_LINEAR_GAUSSIAN = 1, VAE_DATA_SIGMOID = 2 };
typedef struct VaeDataDesc {
  int32_t kind;            /* VAE_DATA_*                                          */
  int32_t dim;             /* -dd  dataset_dimension                              */
  int32_t padding_dim;     /* --padding_dim                                       */
  int32_t intrinsic_dim;   /* -did (linear_gaussian only; run.py:37 default 3)    */
} VaeDataDesc;
/* Row width of a generated batch: dim+padding (+1 for sigmoid, datasets.py:235). */
int32_t vae_data_width(const VaeDataDesc* d);
/* Normals consumed per generated row: dim (sphere, sigmoid) / intrinsic_dim (linear). */
int32_t vae_data_normals_per_row(const VaeDataDesc* d);

/* ------------------------------------------------------------------ jax.random
 * Replaces random.split (vae.py:56,125; model.py:140,154,208; datasets.py:51).
 * d_key: device uint32[2]; d_out: device uint32[num*2]. */
int vae_random_split(const uint32_t* d_key, uint32_t* d_out, int32_t num, void* stream);
/* Raw threefry2x32 bits over iota(n) (jax._random_bits). */
int vae_random_bits(const uint32_t* d_key, uint32_t* d_out, int64_t n, void* stream);
/* Replaces random.normal(key, shape) float32 (model.py:227; datasets.py:76,173,185,237,242).
 * Bits are exact; floats are bit-identical to oracle/jax_random.py. */
int vae_random_normal(const uint32_t* d_key, float* d_out, int64_t n, void* stream);
/* Rows [row_begin, row_end) of the (rows, cols) array random.normal(key, (rows, cols))
 * would return, written densely to d_out ((row_end-row_begin) x cols).  Used by the
 * data-parallel path so every rank draws exactly its own slice of eps. */
int vae_random_normal_rows(const uint32_t* d_key, float* d_out, int64_t rows, int64_t cols,
                           int64_t row_begin, int64_t row_end, void* stream);

/* Data-parallel shard of random.normal(key, (rows, cols)), rows even: global rows [lo, hi)
 * (lo <= hi <= rows/2) followed by their threefry partners [rows/2 + lo, rows/2 + hi), written
 * densely as (2*(hi-lo), cols).  Row b and row b + rows/2 come out of the same threefry
 * blocks, so a shard made of such pairs computes no block it does not use. */
int vae_random_normal_pairs(const uint32_t* d_key, float* d_out, int64_t rows, int64_t cols,
                            int64_t lo, int64_t hi, void* stream);

/* -------------------------------------------------------------------- datasets
 * Replaces <Dataset>.get_batch(size) given the already-split batch key
 * (datasets.py:75-84, 183-195, 240-249).  d_A: device generator matrix
 * (dim x intrinsic_dim for linear, dim x 1 for sigmoid, ignored for sphere).
 * d_out: device (size, vae_data_width) float32. */
int vae_dataset_batch(const VaeDataDesc* d, const uint32_t* d_key, const float* d_A,
                      int32_t size, float* d_out, void* stream);
/* The same rows/partners shard of <Dataset>.get_batch(size): (2*(hi-lo), width). */
int vae_dataset_batch_pairs(const VaeDataDesc* d, const uint32_t* d_key, const float* d_A,
                            int32_t size, int32_t lo, int32_t hi, float* d_out, void* stream);
/* Replaces <Dataset>.score_batch(batch) (datasets.py:67-73, 201-206, 255-261).
 * d_scores: device float[2] = {first, second} statistic in the reference's dict order
 * (second is 0 for linear_gaussian). */
int vae_dataset_score(const VaeDataDesc* d, const float* d_A, const float* d_batch,
                      int32_t size, float* d_scores, void* stream);

/* ------------------------------------------------------------------- VAE step
 * Workspace the step / loss / sample entry points need, in bytes, for `rows` rows. */
size_t vae_workspace_bytes(const VaeDesc* desc, int32_t rows);

/* Replaces VAE.train_step(optimizer, batch, z1, z2, epsilon) (networks.py:87-101,
 * called at vae.py:129).  params/adam_m/adam_v: device float[vae_param_count],
 * updated in place; d_step: device int32 (number of updates applied so far,
 * incremented); d_x: device (B, D); d_z: device (B, L+D) = [z1 | z2] exactly as
 * model.py:227 draws it; d_loss: device float[1]. */
int vae_train_step(const VaeDesc* desc, float* params, float* adam_m, float* adam_v,
                   int32_t* d_step, const float* d_x, const float* d_z, float* d_loss,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Gradient only (no update): d_grad device float[vae_param_count].  Used by the
 * data-parallel path (allreduce between gradient and update) and by parity tests.
 * grad_scale_rows: the B the mean at networks.py:98 divides by (global batch). */
int vae_loss_and_grad(const VaeDesc* desc, const float* params, const float* d_x,
                      const float* d_z, int32_t rows, int32_t grad_scale_rows,
                      float* d_loss, float* d_grad, void* workspace, size_t workspace_bytes,
                      void* stream);
/* The same gradient with per-segment completion events, for overlapping the data-parallel allreduce with the
 * backward pass ("bucketed per layer and overlapped with backward").  vae_grad_segments lists the
 * flat-gradient ranges in the order the backward pass completes them (decoder layers last -> first, the latent
 * block epsilon_p/epsilon, encoder layers last -> first; one segment for two-decoder models); it returns the
 * count (out may be NULL).  ready_events: one cudaEvent_t per segment, recorded on `stream` as soon as that range
 * of d_grad is final.  vae_adam_update_range applies flax Adam (networks.py:100) to [begin, end) only and does
 * not advance the step counter; vae_adam_step_increment advances it once all ranges are done. */
typedef struct VaeGradSegment { int64_t begin, end; } VaeGradSegment;
int32_t vae_grad_segments(const VaeDesc* desc, VaeGradSegment* out, int32_t max_out);
int vae_loss_and_grad_events(const VaeDesc* desc, const float* params, const float* d_x, const float* d_z,
                             int32_t rows, int32_t grad_scale_rows, float* d_loss, float* d_grad,
                             void* workspace, size_t workspace_bytes, void* const* ready_events, void* stream);
int vae_adam_update_range(const VaeDesc* desc, float* params, float* adam_m, float* adam_v, const int32_t* d_step,
                          const float* d_grad, int64_t begin, int64_t end, void* stream);
int vae_adam_step_increment(int32_t* d_step, void* stream);
/* flax.optim.Adam.apply_gradient (networks.py:100) on a flat gradient. */
int vae_adam_update(const VaeDesc* desc, float* params, float* adam_m, float* adam_v,
                    int32_t* d_step, const float* d_grad, void* stream);

/* Replaces VAE.loss(model, batch, z1, z2, epsilon) (networks.py:103-113, called at
 * vae.py:135) on `rows` rows.  d_out: device float[3] = {loss, dkl.mean, mse.mean};
 * logvar_e / epsilon are parameters the caller already holds. */
int vae_loss(const VaeDesc* desc, const float* params, const float* d_x, const float* d_z,
             int32_t rows, float* d_out, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces the sampling-mode forward (networks.py:62-65 through vae.py:191-201).
 * d_current_epsilon: device float[1] (self.current_epsilon); d_xhat: device (rows, D). */
int vae_sample(const VaeDesc* desc, const float* params, const float* d_z, int32_t rows,
               const float* d_current_epsilon, float* d_xhat, void* workspace,
               size_t workspace_bytes, void* stream);

/* --------------------------------------------------------- fused training loop
 * One job = one model of the sweep.  Replaces the body of the hot loop
 * model.py:221-222 (dataset.get_batch + train_one_batch: key splits, eps draw,
 * generator, forward, loss, backward, Adam) for n_steps consecutive steps, for
 * every job, in ONE persistent launch.  Eval / plot events (model.py:214-220) are
 * scheduled by the caller between launches.  All pointers are device pointers. */
typedef struct VaeTrainJob {
  VaeDesc desc;
  VaeDataDesc data;
  float* params;            /* [P] in/out */
  float* adam_m;            /* [P] in/out */
  float* adam_v;            /* [P] in/out */
  int32_t* step;            /* [1] in/out */
  uint32_t* model_key;      /* [2] in/out  carried model-side key (row 1 rule)   */
  uint32_t* dataset_key;    /* [2] in/out  carried dataset-side key (row 0 rule) */
  const float* A;           /* generator matrix or NULL                          */
  float* loss_out;          /* [n_steps] out: per-step train loss                */
  void* workspace;          /* vae_train_fused_workspace_bytes(job) bytes, or NULL if 0 */
} VaeTrainJob;

/* Note on padding rows: generated batches have exactly-zero padding columns, so the fused kernels skip
 * the encoder-kernel rows that multiply them (gradient exactly 0).  If the caller supplies NON-zero Adam
 * moments for those rows (e.g. a checkpoint of a -dn noisy run) the kernels detect it at launch and apply
 * the g = 0 Adam update to them every step, exactly like flax Adam / vae_adam_update would. */
/* >0 if a fused persistent kernel exists for this job's shape (1 = single-CTA
 * kernel, 2 = thread-block-cluster kernel), else 0. */
int vae_train_fused_supported(const VaeTrainJob* job);
size_t vae_train_fused_workspace_bytes(const VaeTrainJob* job);
/* Resident CTAs (= models) per SM of the job's fused kernel on the current device: the
 * pack size that fills the GPU is this x the SM count.  0 if unsupported, <0 on error. */
int vae_train_fused_ctas_per_sm(const VaeTrainJob* job);
/* Jobs of this shape that are resident at once on the current device (CTAs per SM x SMs for
 * class 1, co-resident clusters for class 2): the pack size that fills the GPU. */
int vae_train_fused_resident_jobs(const VaeTrainJob* job);
/* Device scratch the launch needs for its job table. */
size_t vae_train_fused_scratch_bytes(int32_t n_jobs);
/* jobs: HOST array of n_jobs descriptors, all of the same vae_train_fused_supported()
 * class.  d_scratch: device buffer of vae_train_fused_scratch_bytes(n_jobs) bytes
 * (the job table is copied into it on `stream`; it must stay untouched until the
 * launch has completed). */
int vae_train_fused(const VaeTrainJob* jobs, int32_t n_jobs, int64_t n_steps, void* d_scratch,
                    size_t scratch_bytes, void* stream);

/* Prepared pack: the same persistent launch with every per-job cost paid ONCE.  vae_pack_prepare
 * validates the jobs, copies the device job table into d_scratch (on `stream`) and records the launch
 * plan (kernel class, models per CTA / cluster size, shared-memory size; the environment knobs are read
 * here) in the caller-owned POD `plan`; vae_pack_train(plan, n_steps) is then exactly ONE kernel launch
 * with no host-side per-job work, so short chunks (tens of steps) are not host-bound.  The plan holds
 * d_scratch and the jobs' device pointers: all of them must stay alive and unmoved while it is used.
 * vae_train_fused(...) == vae_pack_prepare(...) followed by vae_pack_train(...).
 * Replaces the loop body model.py:221-222 as called from model.py:213 (one call per chunk of the loop). */
typedef struct VaePackPlan {
  int32_t abi;            /* VAE_B200_ABI_VERSION the plan was built by            */
  int32_t cls;            /* vae_train_fused_supported() class of the pack          */
  int32_t n_jobs;
  int32_t shape_id;       /* class 1: kernel instance                               */
  int32_t homogeneous;    /* class 1: every job has the same shape and batch        */
  int32_t gm;             /* class 1: models per CTA                                */
  int32_t cluster;        /* CTAs per thread-block cluster (class 2: CTAs per model)  */
  int32_t nbuf;           /* class 1: eps buffers (2 = drawn one step ahead)        */
  uint64_t smem_bytes;    /* dynamic shared memory per CTA                          */
  void* d_jobs;           /* device job table (= d_scratch)                         */
  int32_t device;         /* CUDA device the plan was prepared on                   */
  int32_t tc;             /* class 2: 1 = tensor-core kernel (fused_mlp_tc.cu), 0 = FFMA cluster kernel */
  int32_t reserved_[4];
} VaePackPlan;
int vae_pack_prepare(const VaeTrainJob* jobs, int32_t n_jobs, void* d_scratch, size_t scratch_bytes,
                     VaePackPlan* plan, void* stream);
int vae_pack_train(const VaePackPlan* plan, int64_t n_steps, void* stream);

/* ------------------------------------------------- tensor-core GEMM building blocks
 * Used by the per-op step when a Dense layer is a true dense contraction (configs[4]:
 * width 4096, batch 65 536).  FP32-level accuracy through a 3xTF32 split; see
 * vae_training_b200/csrc/gemm_tc.cu.  They replace the XLA dot inside flax.nn.Dense
 * (networks.py:33-34) and its two transposes in the backward pass.
 *
 * hi = TF32(x), lo = TF32(x - hi), both rounded to nearest: hi + lo = x to ~2^-23. */
int vae_split_tf32(const float* d_x, float* d_hi, float* d_lo, int64_t n, void* stream);
/* x (rows x cols, leading dimension ldx) -> hiT, loT (cols x rows, leading dimension ldt). */
int vae_transpose_split_tf32(const float* d_x, int64_t ldx, float* d_hiT, float* d_loT, int64_t ldt,
                             int32_t rows, int32_t cols, void* stream);
/* C[M,N] = epi(A[M,K] * B[N,K]^T), A = Ah + Al, B = Bh + Bl (both row-major, K contiguous;
 * lda, ldb multiples of 4 floats, operands 16-byte aligned).  tcgen05.mma kind::tf32 with the
 * accumulator in tensor memory, operands staged by TMA.  flags: 1 add bias[N], 2 relu,
 * 4 multiply by (mask[M,N] > 0). */
int vae_gemm_nt_3xtf32(int32_t M, int32_t N, int32_t K, const float* d_Ah, const float* d_Al, int64_t lda,
                       const float* d_Bh, const float* d_Bl, int64_t ldb, float* d_C, int64_t ldc,
                       const float* d_bias, const float* d_mask, int64_t ldm, int32_t flags, void* stream);
/* C[M,N] = At[K,M]^T * Bt[K,N]: both operands K-row-major (M- / N-contiguous), read through MN-major tcgen05
 * descriptors (TMA SWIZZLE_128B_ATOM_32B boxes).  This is dW = X^T dY (the transposes XLA inserts for
 * flax.nn.Dense's kernel gradient, networks.py:99) taken straight from the batch-major split activations and
 * gradients: At = split(X) (K = batch rows), Bt = split(dY).  lda, ldb multiples of 4 floats. */
int vae_gemm_tn_3xtf32(int32_t M, int32_t N, int32_t K, const float* d_Ath, const float* d_Atl, int64_t lda,
                       const float* d_Bth, const float* d_Btl, int64_t ldb, float* d_C, int64_t ldc, void* stream);

/* ----------------------------------------------------------------------- misc */
const char* vae_last_error(void);
int vae_abi_version(void);
/* Number of CUDA kernels this library has launched from the calling host thread so far
 * (thread-local tally; bench.py reports its difference over the timed region as gpu_launches). */
int64_t vae_launch_count(void);
/* FP32 FFMA throughput probe (the roofline denominator for the small-model
 * configs): runs `iters` dependent-chain FFMA batches on every SM; d_sink device
 * float[>= blocks*threads]; *flops_out = total flops issued. */
int vae_ffma_probe(float* d_sink, int32_t blocks, int32_t threads, int32_t iters,
                   double* flops_out, void* stream);

/* tcgen05 operand-form probe (hardware self-check, not on the training path): copies two caller-built
 * shared-memory operand images verbatim, issues `ksteps` M128 x N x K8 kind::tf32 MMAs with the given
 * descriptor high words (everything except the start address; a_step / b_step bytes are added to the start
 * addresses per k-step; the k loop runs `repeat` times) and instruction descriptor, and returns the 128 x N
 * FP32 accumulator in d_out[0 .. 128 N) and the elapsed SM cycles (first issue -> completion) in d_out[128 N].
 * Documents what fused_mlp_tc.cu assumes about K-/MN-major SWIZZLE_128B descriptors, N = 104 and the
 * truncation of raw FP32 operands to TF32. */
int vae_umma_probe(const float* d_a_img, int32_t a_bytes, const float* d_b_img, int32_t b_bytes,
                   uint64_t desc_a_hi, uint64_t desc_b_hi, uint32_t idesc, int32_t ksteps, int32_t a_step,
                   int32_t b_step, int32_t N, int32_t repeat, float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAE_B200_H_ */
