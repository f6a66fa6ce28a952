// api.cu -- error handling, shape/layout helpers and misc entry points of libvae_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local int64_t g_launches = 0;
void count_launch() { ++g_launches; }
int64_t launch_count() { return g_launches; }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return VAE_ERR_CUDA;
}

// Layer sizes per vae.py:53-54; flat layout per include/vae_b200.h.
int make_layout(const VaeDesc* d, Layout* L) {
  if (!d) { set_error("desc is NULL"); return VAE_ERR_INVALID; }
  if (d->batch <= 0 || d->data_dim <= 0 || d->latent_dim <= 0) {
    set_error("batch/data_dim/latent_dim must be positive (got %d/%d/%d)", d->batch, d->data_dim, d->latent_dim);
    return VAE_ERR_INVALID;
  }
  if (d->n_enc_hidden < 0 || d->n_enc_hidden > VAE_MAX_HIDDEN || d->n_dec_hidden < 0 || d->n_dec_hidden > VAE_MAX_HIDDEN) {
    set_error("hidden layer count out of range [0,%d]", VAE_MAX_HIDDEN);
    return VAE_ERR_INVALID;
  }
  memset(L, 0, sizeof(*L));
  L->n_enc = d->n_enc_hidden + 1;
  L->n_dec = d->n_dec_hidden + 1;
  L->enc_sizes[0] = d->data_dim;
  for (int i = 0; i < d->n_enc_hidden; ++i) {
    if (d->enc_hidden[i] <= 0) { set_error("enc_hidden[%d] must be positive", i); return VAE_ERR_INVALID; }
    L->enc_sizes[i + 1] = d->enc_hidden[i];
  }
  L->enc_sizes[L->n_enc] = d->latent_dim;
  L->dec_sizes[0] = d->latent_dim;
  for (int i = 0; i < d->n_dec_hidden; ++i) {
    if (d->dec_hidden[i] <= 0) { set_error("dec_hidden[%d] must be positive", i); return VAE_ERR_INVALID; }
    L->dec_sizes[i + 1] = d->dec_hidden[i];
  }
  L->dec_sizes[L->n_dec] = d->data_dim;
  int64_t off = 0;
  int mw = 1;
  for (int i = 0; i < L->n_enc; ++i) {
    L->enc_w[i] = off; off += (int64_t)L->enc_sizes[i] * L->enc_sizes[i + 1];
    L->enc_b[i] = off; off += L->enc_sizes[i + 1];
    if (L->enc_sizes[i] > mw) mw = L->enc_sizes[i];
    if (L->enc_sizes[i + 1] > mw) mw = L->enc_sizes[i + 1];
  }
  for (int i = 0; i < L->n_dec; ++i) {
    L->dec_w[i] = off; off += (int64_t)L->dec_sizes[i] * L->dec_sizes[i + 1];
    L->dec_b[i] = off; off += L->dec_sizes[i + 1];
    if (L->dec_sizes[i] > mw) mw = L->dec_sizes[i];
    if (L->dec_sizes[i + 1] > mw) mw = L->dec_sizes[i + 1];
  }
  if (d->sig_decoder) {
    for (int i = 0; i < L->n_dec; ++i) {
      L->sig_w[i] = off; off += (int64_t)L->dec_sizes[i] * L->dec_sizes[i + 1];
      L->sig_b[i] = off; off += L->dec_sizes[i + 1];
    }
  }
  L->off_eps_p = off; off += d->latent_dim;
  if (d->tunable_decoder_var) { L->off_eps = off; off += 1; } else { L->off_eps = -1; }
  L->n_params = off;
  L->max_width = mw;
  return 0;
}

}  // namespace vae

extern "C" {

int64_t vae_param_count(const VaeDesc* desc) {
  vae::Layout L;
  int rc = vae::make_layout(desc, &L);
  return rc ? (int64_t)rc : L.n_params;
}

int32_t vae_data_width(const VaeDataDesc* d) {
  if (!d) return VAE_ERR_INVALID;
  return d->dim + d->padding_dim + (d->kind == VAE_DATA_SIGMOID ? 1 : 0);
}

int32_t vae_data_normals_per_row(const VaeDataDesc* d) {
  if (!d) return VAE_ERR_INVALID;
  return d->kind == VAE_DATA_LINEAR_GAUSSIAN ? d->intrinsic_dim : d->dim;
}

const char* vae_last_error(void) { return vae::g_err; }
int vae_abi_version(void) { return VAE_B200_ABI_VERSION; }
int64_t vae_launch_count(void) { return vae::launch_count(); }

}  // extern "C"

// ------------------------------------------------------------------ FFMA probe
// Eight independent FFMA chains per thread; 2 flops per FFMA.
__global__ void ffma_probe_kernel(float* sink, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.999f, c = 1e-4f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int vae_ffma_probe(float* d_sink, int32_t blocks, int32_t threads, int32_t iters,
                              double* flops_out, void* stream) {
  VAE_REQUIRE(d_sink && blocks > 0 && threads > 0 && threads <= 1024 && iters > 0, "bad ffma probe arguments");
  ffma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_sink, iters);
  VAE_LAUNCH_OK("ffma_probe_kernel");
  if (flops_out) *flops_out = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  return 0;
}
