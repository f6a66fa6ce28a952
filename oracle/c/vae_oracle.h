/* ORACLE (test infrastructure, not product code) -- plain-C restatement of the VAE
 * train-step hot path of virajmehta/vae-training.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call this.
 *
 * It mirrors oracle/*.py function for function (each cites the reference file:line it
 * follows) and exists because the NumPy oracle is too slow to serve as the timed CPU
 * baseline.  PARITY UNPINNED against the real JAX/Flax runtime (which cannot run here:
 * no jax/flax wheels, no network); pinned against the Random123 / documented jax.random
 * known answers through tests/test_oracle_c.py.
 */
#ifndef VAE_ORACLE_H_
#define VAE_ORACLE_H_
#include <stdint.h>

#define OR_MAX_HIDDEN 8

typedef struct OrDesc {            /* same field order as VaeDesc in include/vae_b200.h */
  int32_t batch, data_dim, latent_dim;
  int32_t n_enc_hidden, enc_hidden[OR_MAX_HIDDEN];
  int32_t n_dec_hidden, dec_hidden[OR_MAX_HIDDEN];
  int32_t sig_decoder, tunable_decoder_var, reserved_;
  double epsilon0, lr, beta1, beta2, adam_eps;   /* Python floats in the reference */
} OrDesc;

typedef struct OrData { int32_t kind, dim, padding_dim, intrinsic_dim; } OrData;  /* 0 sphere 1 linear 2 sigmoid */

typedef struct OrJob {
  OrDesc desc; OrData data;
  float *params, *adam_m, *adam_v;  /* [P] in/out (host) */
  int32_t* step;                    /* [1] */
  uint32_t *model_key, *dataset_key;/* [2] each, carried keys in/out */
  const float* A;
  float* loss_out;                  /* [n_steps] */
} OrJob;

/* jax.random (oracle/jax_random.py) */
void or_threefry2x32(uint32_t k0, uint32_t k1, uint32_t* x0, uint32_t* x1);
void or_random_bits(const uint32_t key[2], uint32_t* out, int64_t n);
void or_split(const uint32_t key[2], uint32_t* out, int32_t num);
void or_normal(const uint32_t key[2], float* out, int64_t n);
/* datasets (oracle/datasets.py) */
int32_t or_data_width(const OrData* d);
void or_dataset_batch(const OrData* d, const uint32_t key[2], const float* A, int32_t size, float* out);
/* step (oracle/vae_step.py) */
int64_t or_param_count(const OrDesc* d);
/* loss + flat gradient; work = NULL lets the callee allocate.  Returns the loss. */
float or_loss_and_grad(const OrDesc* d, const float* params, const float* x, const float* z, float* grad);
void or_adam_update(const OrDesc* d, float* params, float* m, float* v, int32_t* step, const float* grad);
/* the hot loop body model.py:221-222 for n_steps iterations (oracle/trainer.py) */
int or_train_loop(const OrJob* job, int64_t n_steps);
/* independent jobs spread over n_threads host threads */
int or_train_loop_mt(const OrJob* jobs, int32_t n_jobs, int64_t n_steps, int32_t n_threads);
#endif
