/* ORACLE (test infrastructure) -- see vae_oracle.h.  Float32 restatement of
 *   networks.py:26-44   FullyConnectedNetwork.apply (Dense, relu between layers)
 *   networks.py:59-84   VAE.apply
 *   networks.py:87-101  VAE.train_step (KL, Gaussian NLL, mean, value_and_grad, Adam)
 *   vae.py:123-130      train_one_batch (split, eps draw, z1/z2 slicing)
 *   model.py:221-222    the hot-loop body
 * mirroring oracle/vae_step.py and oracle/trainer.py.  The backward pass is the
 * hand-derived one of SURVEY.md Appendix B.  This TU is compiled with vectorisation
 * (it is the timed CPU baseline); tolerance-level, not bit-level, agreement with the
 * NumPy oracle is expected and tested. */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "vae_oracle.h"

#define LOG_2PI 1.8378770664093453f

typedef struct Net { int n; int sizes[OR_MAX_HIDDEN + 2]; int64_t w[OR_MAX_HIDDEN + 1], b[OR_MAX_HIDDEN + 1]; } Net;
typedef struct Lay { Net enc, dec, sig; int has_sig; int64_t off_lv, off_eps, n_params; int maxw; } Lay;

static void make_layout(const OrDesc* d, Lay* L) {
  memset(L, 0, sizeof(*L));
  L->enc.n = d->n_enc_hidden + 1; L->dec.n = d->n_dec_hidden + 1; L->sig.n = L->dec.n;
  L->enc.sizes[0] = d->data_dim;
  for (int i = 0; i < d->n_enc_hidden; ++i) L->enc.sizes[i + 1] = d->enc_hidden[i];
  L->enc.sizes[L->enc.n] = d->latent_dim;
  L->dec.sizes[0] = d->latent_dim;
  for (int i = 0; i < d->n_dec_hidden; ++i) L->dec.sizes[i + 1] = d->dec_hidden[i];
  L->dec.sizes[L->dec.n] = d->data_dim;
  memcpy(L->sig.sizes, L->dec.sizes, sizeof(L->dec.sizes));
  int64_t off = 0; int mw = 1;
  Net* nets[3] = {&L->enc, &L->dec, &L->sig};
  L->has_sig = d->sig_decoder != 0;
  for (int q = 0; q < (L->has_sig ? 3 : 2); ++q) {
    Net* n = nets[q];
    for (int i = 0; i < n->n; ++i) {
      n->w[i] = off; off += (int64_t)n->sizes[i] * n->sizes[i + 1];
      n->b[i] = off; off += n->sizes[i + 1];
      if (n->sizes[i] > mw) mw = n->sizes[i];
      if (n->sizes[i + 1] > mw) mw = n->sizes[i + 1];
    }
  }
  L->off_lv = off; off += d->latent_dim;
  if (d->tunable_decoder_var) { L->off_eps = off; off += 1; } else L->off_eps = -1;
  L->n_params = off; L->maxw = mw;
}

int64_t or_param_count(const OrDesc* d) { Lay L; make_layout(d, &L); return L.n_params; }

/* y[B,n] = x[B,m] W[m,n] + b ; relu optional */
static void dense_fwd(const float* x, int ldx, const float* W, const float* b, float* y, int B, int m, int n, int relu) {
  for (int r = 0; r < B; ++r) {
    float* yr = y + (size_t)r * n;
    const float* xr = x + (size_t)r * ldx;
    for (int j = 0; j < n; ++j) yr[j] = b[j];
    for (int k = 0; k < m; ++k) {
      const float xv = xr[k];
      const float* wr = W + (size_t)k * n;
      for (int j = 0; j < n; ++j) yr[j] += xv * wr[j];
    }
    if (relu) for (int j = 0; j < n; ++j) yr[j] = yr[j] > 0.f ? yr[j] : 0.f;
  }
}

/* dW[m,n] = h^T d ; db[n] = sum_r d ; optionally dh[B,m] = (d W^T) * (h > 0 if mask) (+= if acc) */
static void dense_bwd(const float* h, int ldh, const float* d, const float* W, float* dW, float* db, float* dh,
                      int B, int m, int n, int mask, int acc) {
  memset(dW, 0, sizeof(float) * (size_t)m * n);
  memset(db, 0, sizeof(float) * n);
  for (int r = 0; r < B; ++r) {
    const float* dr = d + (size_t)r * n;
    const float* hr = h + (size_t)r * ldh;
    for (int j = 0; j < n; ++j) db[j] += dr[j];
    for (int k = 0; k < m; ++k) {
      const float hv = hr[k];
      float* wr = dW + (size_t)k * n;
      for (int j = 0; j < n; ++j) wr[j] += hv * dr[j];
    }
    if (dh) {
      float* o = dh + (size_t)r * m;
      for (int k = 0; k < m; ++k) {
        const float* wr = W + (size_t)k * n;
        float s = 0.f;
#pragma omp simd reduction(+ : s)
        for (int j = 0; j < n; ++j) s += dr[j] * wr[j];
        if (mask && !(hr[k] > 0.f)) s = 0.f;
        o[k] = acc ? o[k] + s : s;
      }
    }
  }
}

typedef struct Work {
  float* enc_act[OR_MAX_HIDDEN + 1]; float* dec_act[OR_MAX_HIDDEN + 1]; float* sig_act[OR_MAX_HIDDEN + 1];
  float *zs, *dxh, *dpre, *ga, *gb, *dzs, *grad, *z, *x;
  float* base;
} Work;

static void work_alloc(const OrDesc* d, const Lay* L, Work* w, int gen_dim) {
  const size_t B = (size_t)d->batch;
  size_t tot = 0;
  for (int i = 0; i < L->enc.n; ++i) tot += B * L->enc.sizes[i + 1];
  for (int i = 0; i < L->dec.n; ++i) tot += 2 * B * L->dec.sizes[i + 1];
  tot += B * d->latent_dim * 2 + B * d->data_dim * 3 + 2 * B * L->maxw + (size_t)L->n_params;
  tot += B * (d->latent_dim + d->data_dim) + B * (size_t)gen_dim + 64;
  w->base = (float*)malloc(sizeof(float) * tot);
  float* p = w->base;
  for (int i = 0; i < L->enc.n; ++i) { w->enc_act[i] = p; p += B * L->enc.sizes[i + 1]; }
  for (int i = 0; i < L->dec.n; ++i) { w->dec_act[i] = p; p += B * L->dec.sizes[i + 1]; }
  for (int i = 0; i < L->dec.n; ++i) { w->sig_act[i] = p; p += B * L->dec.sizes[i + 1]; }
  w->zs = p; p += B * d->latent_dim; w->dzs = p; p += B * d->latent_dim;
  w->dxh = p; p += B * d->data_dim; w->dpre = p; p += B * d->data_dim; w->x = p; p += B * d->data_dim;
  w->ga = p; p += B * L->maxw; w->gb = p; p += B * L->maxw;
  w->grad = p; p += L->n_params;
  w->z = p;
}

static void mlp_fwd(const Net* n, const float* P, const float* in, int ldin, float** acts, int B) {
  const float* h = in; int ld = ldin;
  for (int i = 0; i < n->n; ++i) {
    dense_fwd(h, ld, P + n->w[i], P + n->b[i], acts[i], B, n->sizes[i], n->sizes[i + 1], i + 1 < n->n);
    h = acts[i]; ld = n->sizes[i + 1];
  }
}

/* consumes dout; writes d(input) to d_in when d_in != NULL */
static void mlp_bwd(const Net* n, const float* P, float* G, const float* in, int ldin, float** acts, const float* dout,
                    float* ga, float* gb, float* d_in, int acc_in, int B) {
  const float* d = dout;
  for (int i = n->n - 1; i >= 0; --i) {
    const float* h = i == 0 ? in : acts[i - 1];
    const int ldh = i == 0 ? ldin : n->sizes[i];
    float* dst = NULL; int acc = 0;
    if (i > 0) dst = (d == ga) ? gb : ga;
    else if (d_in) { dst = d_in; acc = acc_in; }
    dense_bwd(h, ldh, d, P + n->w[i], G + n->w[i], G + n->b[i], dst, B, n->sizes[i], n->sizes[i + 1], i > 0, acc);
    d = dst;
  }
}

static float loss_and_grad(const OrDesc* d, const Lay* L, Work* w, const float* P, const float* x, const float* z,
                           float* G) {
  const int B = d->batch, Ld = d->latent_dim, D = d->data_dim, ldz = Ld + D;
  const float fB = (float)B;
  const float* lv = P + L->off_lv;
  const float eps0 = (float)d->epsilon0;
  const float e = d->tunable_decoder_var ? P[L->off_eps] * eps0 : eps0;
  const float sd = expf(0.5f * e), var = expf(e);
  mlp_fwd(&L->enc, P, x, D, w->enc_act, B);
  const float* mu = w->enc_act[L->enc.n - 1];
  float s[512];
  for (int l = 0; l < Ld; ++l) s[l] = expf(0.5f * lv[l]);
  double s_mu2 = 0;
  for (int r = 0; r < B; ++r)
    for (int l = 0; l < Ld; ++l) {
      const float m = mu[(size_t)r * Ld + l];
      w->zs[(size_t)r * Ld + l] = m + s[l] * z[(size_t)r * ldz + l];
      s_mu2 += (double)(m * m);
    }
  mlp_fwd(&L->dec, P, w->zs, Ld, w->dec_act, B);
  const float* dec = w->dec_act[L->dec.n - 1];
  float* sgp = NULL;
  if (L->has_sig) { mlp_fwd(&L->sig, P, w->zs, Ld, w->sig_act, B); sgp = w->sig_act[L->dec.n - 1]; }
  double s_r2 = 0, s_rz = 0;
  for (int r = 0; r < B; ++r)
    for (int j = 0; j < D; ++j) {
      const size_t i = (size_t)r * D + j;
      const float z2 = z[(size_t)r * ldz + Ld + j];
      float xh = dec[i], sg = 0.f;
      if (sgp) { sg = 1.f / (1.f + expf(-sgp[i])); xh += sg; }
      xh += z2 * sd;
      const float res = xh - x[i];
      s_r2 += (double)(res * res); s_rz += (double)(res * z2);
      const float g = res / (var * fB);
      w->dxh[i] = g;
      if (sgp) w->dpre[i] = g * sg * (1.f - sg);
    }
  float klc = 0.f;
  for (int l = 0; l < Ld; ++l) klc += 1.f + lv[l] - expf(lv[l]);
  const float dkl = (-0.5f * klc * fB + 0.5f * (float)s_mu2) / fB;
  const float mse = (0.5f * (float)s_r2 / var + 0.5f * (LOG_2PI + e) * fB * (float)D) / fB;
  if (d->tunable_decoder_var)
    G[L->off_eps] = eps0 * (-0.5f * (float)s_r2 / var + 0.5f * fB * (float)D + 0.5f * sd * (float)s_rz / var) / fB;
  mlp_bwd(&L->dec, P, G, w->zs, Ld, w->dec_act, w->dxh, w->ga, w->gb, w->dzs, 0, B);
  if (L->has_sig) mlp_bwd(&L->sig, P, G, w->zs, Ld, w->sig_act, w->dpre, w->ga, w->gb, w->dzs, 1, B);
  for (int l = 0; l < Ld; ++l) {
    float a = 0.f;
    for (int r = 0; r < B; ++r) a += w->dzs[(size_t)r * Ld + l] * z[(size_t)r * ldz + l];
    G[L->off_lv + l] = a * 0.5f * s[l] - 0.5f * (1.f - expf(lv[l]));
  }
  for (int r = 0; r < B; ++r)
    for (int l = 0; l < Ld; ++l) w->dzs[(size_t)r * Ld + l] += mu[(size_t)r * Ld + l] / fB;
  mlp_bwd(&L->enc, P, G, x, D, w->enc_act, w->dzs, w->ga, w->gb, NULL, 0, B);
  return dkl + mse;
}

static void adam(const OrDesc* d, int64_t n, float* P, float* M, float* V, int32_t* step, const float* G) {
  /* legacy flax.optim.Adam [3P]: Python-float hyper-parameters; (1 - beta) is evaluated in
   * double and rounded, beta ** t is a float32 power of the float32 beta */
  const float t = (float)(step[0] + 1);
  const float b1 = (float)d->beta1, b2 = (float)d->beta2, lr = (float)d->lr, eps = (float)d->adam_eps;
  const float omb1 = (float)(1.0 - d->beta1), omb2 = (float)(1.0 - d->beta2);
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  for (int64_t i = 0; i < n; ++i) {
    const float g = G[i];
    const float m = b1 * M[i] + omb1 * g;
    const float v = b2 * V[i] + omb2 * g * g;
    M[i] = m; V[i] = v;
    P[i] = P[i] - lr * (m / bc1) / (sqrtf(v / bc2) + eps);
  }
  step[0] += 1;
}

float or_loss_and_grad(const OrDesc* d, const float* params, const float* x, const float* z, float* grad) {
  Lay L; make_layout(d, &L);
  Work w; work_alloc(d, &L, &w, 0);
  const float loss = loss_and_grad(d, &L, &w, params, x, z, grad);
  free(w.base);
  return loss;
}

void or_adam_update(const OrDesc* d, float* params, float* m, float* v, int32_t* step, const float* grad) {
  Lay L; make_layout(d, &L);
  adam(d, L.n_params, params, m, v, step, grad);
}

int or_train_loop(const OrJob* job, int64_t n_steps) {
  const OrDesc* d = &job->desc;
  Lay L; make_layout(d, &L);
  if (d->latent_dim > 512) return -1;
  Work w; work_alloc(d, &L, &w, 0);
  uint32_t mk[2] = {job->model_key[0], job->model_key[1]};
  uint32_t dk[2] = {job->dataset_key[0], job->dataset_key[1]};
  const int ldz = d->latent_dim + d->data_dim;
  for (int64_t s = 0; s < n_steps; ++s) {
    uint32_t sp[4];
    or_split(dk, sp, 2);                       /* datasets.py:50-52: row 0 carried, row 1 returned */
    const uint32_t bkey[2] = {sp[2], sp[3]};
    dk[0] = sp[0]; dk[1] = sp[1];
    or_dataset_batch(&job->data, bkey, job->A, d->batch, w.x);            /* model.py:221 */
    or_split(mk, sp, 2);                       /* vae.py:125: row 0 consumed, row 1 carried */
    const uint32_t lkey[2] = {sp[0], sp[1]};
    mk[0] = sp[2]; mk[1] = sp[3];
    or_normal(lkey, w.z, (int64_t)d->batch * ldz);                        /* model.py:227 */
    const float loss = loss_and_grad(d, &L, &w, job->params, w.x, w.z, w.grad);   /* networks.py:87-99 */
    adam(d, L.n_params, job->params, job->adam_m, job->adam_v, job->step, w.grad); /* networks.py:100 */
    if (job->loss_out) job->loss_out[s] = loss;
  }
  job->model_key[0] = mk[0]; job->model_key[1] = mk[1];
  job->dataset_key[0] = dk[0]; job->dataset_key[1] = dk[1];
  free(w.base);
  return 0;
}

typedef struct MtArg { const OrJob* jobs; int n_jobs; int64_t n_steps; int next; pthread_mutex_t mu; int rc; } MtArg;

static void* mt_worker(void* p) {
  MtArg* a = (MtArg*)p;
  for (;;) {
    pthread_mutex_lock(&a->mu);
    const int i = a->next++;
    pthread_mutex_unlock(&a->mu);
    if (i >= a->n_jobs) break;
    if (or_train_loop(&a->jobs[i], a->n_steps) != 0) a->rc = -1;
  }
  return NULL;
}

int or_train_loop_mt(const OrJob* jobs, int32_t n_jobs, int64_t n_steps, int32_t n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_jobs) n_threads = n_jobs;
  MtArg a; a.jobs = jobs; a.n_jobs = n_jobs; a.n_steps = n_steps; a.next = 0; a.rc = 0;
  pthread_mutex_init(&a.mu, NULL);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, mt_worker, &a);
  for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
  free(th);
  pthread_mutex_destroy(&a.mu);
  return a.rc;
}
