// common.cuh -- shared device helpers for libvae_b200 (sm_100a).
//
// RNG restates jax.random's legacy threefry2x32 path (call sites: model.py:227,
// vae.py:56,125, datasets.py:51,76,185,242 in the reference).  Every float op on the
// normal() path is an explicitly rounded __f*_rn intrinsic (fused exactly where the oracle
// fuses, never by ptxas' choice): results are bit-identical to oracle/jax_random.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vae_b200.h"

namespace vae {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch();      // thread-local tally behind vae_launch_count()

#define VAE_CUDA_OK(call)                                   \
  do {                                                      \
    int _rc = ::vae::check_cuda((call), #call);             \
    if (_rc != 0) return _rc;                               \
  } while (0)
#define VAE_LAUNCH_OK(name)                                 \
  do {                                                      \
    int _rc = ::vae::check_cuda(cudaGetLastError(), name);  \
    if (_rc != 0) return _rc;                               \
    ::vae::count_launch();                                  \
  } while (0)
#define VAE_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) {                                          \
      ::vae::set_error(__VA_ARGS__);                        \
      return VAE_ERR_INVALID;                               \
    }                                                       \
  } while (0)

// ---------------------------------------------------------------- threefry2x32
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

#define VAE_TF_ROUND(r)  \
  x0 += x1;              \
  x1 = rotl32(x1, r);    \
  x1 ^= x0;

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0; x1 += k1;
  VAE_TF_ROUND(13) VAE_TF_ROUND(15) VAE_TF_ROUND(26) VAE_TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  VAE_TF_ROUND(17) VAE_TF_ROUND(29) VAE_TF_ROUND(16) VAE_TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  VAE_TF_ROUND(13) VAE_TF_ROUND(15) VAE_TF_ROUND(26) VAE_TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  VAE_TF_ROUND(17) VAE_TF_ROUND(29) VAE_TF_ROUND(16) VAE_TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  VAE_TF_ROUND(13) VAE_TF_ROUND(15) VAE_TF_ROUND(26) VAE_TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
}

// Two independent blocks interleaved statement by statement (ILP for latency-bound callers).
#define VAE_TF_ROUND2(r)   \
  a0 += a1; b0 += b1;      \
  a1 = rotl32(a1, r); b1 = rotl32(b1, r); \
  a1 ^= a0; b1 ^= b0;
__device__ __forceinline__ void threefry2x32_x2(uint32_t ka0, uint32_t ka1, uint32_t& a0, uint32_t& a1,
                                                uint32_t kb0, uint32_t kb1, uint32_t& b0, uint32_t& b1) {
  const uint32_t ka2 = ka0 ^ ka1 ^ 0x1BD11BDAu, kb2 = kb0 ^ kb1 ^ 0x1BD11BDAu;
  a0 += ka0; a1 += ka1; b0 += kb0; b1 += kb1;
  VAE_TF_ROUND2(13) VAE_TF_ROUND2(15) VAE_TF_ROUND2(26) VAE_TF_ROUND2(6)
  a0 += ka1; a1 += ka2 + 1u; b0 += kb1; b1 += kb2 + 1u;
  VAE_TF_ROUND2(17) VAE_TF_ROUND2(29) VAE_TF_ROUND2(16) VAE_TF_ROUND2(24)
  a0 += ka2; a1 += ka0 + 2u; b0 += kb2; b1 += kb0 + 2u;
  VAE_TF_ROUND2(13) VAE_TF_ROUND2(15) VAE_TF_ROUND2(26) VAE_TF_ROUND2(6)
  a0 += ka0; a1 += ka1 + 3u; b0 += kb0; b1 += kb1 + 3u;
  VAE_TF_ROUND2(17) VAE_TF_ROUND2(29) VAE_TF_ROUND2(16) VAE_TF_ROUND2(24)
  a0 += ka1; a1 += ka2 + 4u; b0 += kb1; b1 += kb2 + 4u;
  VAE_TF_ROUND2(13) VAE_TF_ROUND2(15) VAE_TF_ROUND2(26) VAE_TF_ROUND2(6)
  a0 += ka2; a1 += ka0 + 5u; b0 += kb2; b1 += kb0 + 5u;
}

// random.split(key): rows (y0[0], y0[1]) and (y1[0], y1[1]) of blocks (0,2), (1,3).
__device__ __forceinline__ void split2(uint32_t k0, uint32_t k1, uint32_t row0[2], uint32_t row1[2]) {
  uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
  threefry2x32(k0, k1, a0, a1);
  threefry2x32(k0, k1, b0, b1);
  row0[0] = a0; row0[1] = b0;
  row1[0] = a1; row1[1] = b1;
}

// Element i of jax._random_bits(key, n): block j = i mod half consumes counters
// (j, j+half) (second counter 0 when it is the odd-length pad).
__device__ __forceinline__ uint32_t random_bits_at(uint32_t k0, uint32_t k1, uint32_t i, uint32_t n) {
  const uint32_t half = (n + 1u) >> 1;
  const bool lane1 = i >= half;
  const uint32_t j = lane1 ? i - half : i;
  uint32_t x0 = j, x1 = (j + half < n) ? j + half : 0u;
  threefry2x32(k0, k1, x0, x1);
  return lane1 ? x1 : x0;
}

// ------------------------------------------------------------ normal from bits
// w = -log1p(-t), t in [0,1): mirrors oracle/jax_random.py:neg_log1p_neg op for op.
__device__ __forceinline__ float neg_log1p_neg(float t) {
  const float a = __fsub_rn(1.0f, t);
  const float c = __fsub_rn(__fsub_rn(1.0f, a), t);
  const uint32_t bits = __float_as_uint(a);
  int k = (int)(bits >> 23) - 127;
  uint32_t mb = (bits & 0x007FFFFFu) | 0x3F800000u;
  const bool big = mb >= 0x3FB504F3u;
  mb = big ? mb - 0x00800000u : mb;
  k += big ? 1 : 0;
  const float m = __uint_as_float(mb);
  const float inv_scale = __uint_as_float((uint32_t)(127 - k) << 23);
  const float f = __fmaf_rn(c, inv_scale, __fsub_rn(m, 1.0f));
  const float fk = (float)k;
  const float z = __fmul_rn(f, f);
  float p = 7.0376836292e-2f;
  p = __fmaf_rn(p, f, -1.1514610310e-1f);
  p = __fmaf_rn(p, f, 1.1676998740e-1f);
  p = __fmaf_rn(p, f, -1.2420140846e-1f);
  p = __fmaf_rn(p, f, 1.4249322787e-1f);
  p = __fmaf_rn(p, f, -1.6668057665e-1f);
  p = __fmaf_rn(p, f, 2.0000714765e-1f);
  p = __fmaf_rn(p, f, -2.4999993993e-1f);
  p = __fmaf_rn(p, f, 3.3333331174e-1f);
  float y = __fmul_rn(__fmul_rn(p, f), z);
  y = __fmaf_rn(fk, -2.12194440e-4f, y);
  y = __fmaf_rn(-0.5f, z, y);
  float r = __fadd_rn(f, y);
  r = __fmaf_rn(fk, 0.693359375f, r);
  return -r;
}

// Giles single-precision erf_inv (XLA ErfInv32), fused Horner steps (see oracle/jax_random.py).
__device__ __forceinline__ float erf_inv_f32(float u) {
  const float w = neg_log1p_neg(__fmul_rn(u, u));
  const bool central = w < 5.0f;
  const float ww = central ? __fsub_rn(w, 2.5f) : __fsub_rn(__fsqrt_rn(w), 3.0f);
  float p = central ? 2.81022636e-08f : -0.000200214257f;
  p = __fmaf_rn(p, ww, central ? 3.43273939e-07f : 0.000100950558f);
  p = __fmaf_rn(p, ww, central ? -3.5233877e-06f : 0.00134934322f);
  p = __fmaf_rn(p, ww, central ? -4.39150654e-06f : -0.00367342844f);
  p = __fmaf_rn(p, ww, central ? 0.00021858087f : 0.00573950773f);
  p = __fmaf_rn(p, ww, central ? -0.00125372503f : -0.0076224613f);
  p = __fmaf_rn(p, ww, central ? -0.00417768164f : 0.00943887047f);
  p = __fmaf_rn(p, ww, central ? 0.246640727f : 1.00167406f);
  p = __fmaf_rn(p, ww, central ? 1.50140941f : 2.83297682f);
  return __fmul_rn(p, u);
}

// random.normal: sqrt(2) * erf_inv(max(lo, f*2 + lo)), lo = nextafter(-1, 0).
__device__ __forceinline__ float normal_from_bits(uint32_t bits) {
  const float lo = -0.99999994f;
  const float f = __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));
  return __fmul_rn(1.41421354f, erf_inv_f32(u));
}

// Same values as normal_from_bits, bit for bit; the tail polynomial (|z| >~ 2.93, 0.34% of
// draws) is taken through a warp-uniform branch so the common path carries no selects and
// no IEEE sqrt.
__device__ __forceinline__ float normal_from_bits_fast(uint32_t bits) {
  const float lo = -0.99999994f;
  const float f = __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));
  const float w = neg_log1p_neg(__fmul_rn(u, u));
  float p;
  if (__all_sync(__activemask(), w < 5.0f)) {
    const float ww = __fsub_rn(w, 2.5f);
    p = 2.81022636e-08f;
    p = __fmaf_rn(p, ww, 3.43273939e-07f);
    p = __fmaf_rn(p, ww, -3.5233877e-06f);
    p = __fmaf_rn(p, ww, -4.39150654e-06f);
    p = __fmaf_rn(p, ww, 0.00021858087f);
    p = __fmaf_rn(p, ww, -0.00125372503f);
    p = __fmaf_rn(p, ww, -0.00417768164f);
    p = __fmaf_rn(p, ww, 0.246640727f);
    p = __fmaf_rn(p, ww, 1.50140941f);
  } else {
    const bool central = w < 5.0f;
    const float ww = central ? __fsub_rn(w, 2.5f) : __fsub_rn(__fsqrt_rn(w), 3.0f);
    p = central ? 2.81022636e-08f : -0.000200214257f;
    p = __fmaf_rn(p, ww, central ? 3.43273939e-07f : 0.000100950558f);
    p = __fmaf_rn(p, ww, central ? -3.5233877e-06f : 0.00134934322f);
    p = __fmaf_rn(p, ww, central ? -4.39150654e-06f : -0.00367342844f);
    p = __fmaf_rn(p, ww, central ? 0.00021858087f : 0.00573950773f);
    p = __fmaf_rn(p, ww, central ? -0.00125372503f : -0.0076224613f);
    p = __fmaf_rn(p, ww, central ? -0.00417768164f : 0.00943887047f);
    p = __fmaf_rn(p, ww, central ? 0.246640727f : 1.00167406f);
    p = __fmaf_rn(p, ww, central ? 1.50140941f : 2.83297682f);
  }
  return __fmul_rn(1.41421354f, __fmul_rn(p, u));
}

// Deterministic exp / sigmoid for the generators (mirrors oracle/datasets.py:det_expf).
__device__ __forceinline__ float det_expf(float x) {
  const float n = floorf(__fadd_rn(__fmul_rn(1.44269504088896341f, x), 0.5f));
  float r = __fsub_rn(x, __fmul_rn(n, 0.693359375f));
  r = __fsub_rn(r, __fmul_rn(n, -2.12194440e-4f));
  const float z = __fmul_rn(r, r);
  float p = 1.9875691500e-4f;
  p = __fadd_rn(__fmul_rn(p, r), 1.3981999507e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 8.3334519073e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 4.1665795894e-2f);
  p = __fadd_rn(__fmul_rn(p, r), 1.6666665459e-1f);
  p = __fadd_rn(__fmul_rn(p, r), 5.0000001201e-1f);
  float y = __fmul_rn(p, z);
  y = __fadd_rn(y, r);
  y = __fadd_rn(y, 1.0f);
  const float scale = __uint_as_float((uint32_t)((int)n + 127) << 23);
  return __fmul_rn(y, scale);
}
__device__ __forceinline__ float det_sigmoid(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, det_expf(-x)));
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Float32 Adam constants derived from the double hyper-parameters the way legacy flax
// does (see VaeDesc in include/vae_b200.h).
struct AdamHyper { float lr, b1, b2, omb1, omb2, eps; };
inline AdamHyper make_adam_hyper(const VaeDesc* d) {
  AdamHyper h;
  h.lr = (float)d->lr; h.b1 = (float)d->beta1; h.b2 = (float)d->beta2;
  h.omb1 = (float)(1.0 - d->beta1); h.omb2 = (float)(1.0 - d->beta2); h.eps = (float)d->adam_eps;
  return h;
}

// Host-side shape helpers (api.cu).
struct Layout {
  int n_enc, n_dec;                 // number of Dense layers per net
  int enc_sizes[VAE_MAX_HIDDEN + 2];
  int dec_sizes[VAE_MAX_HIDDEN + 2];
  int64_t enc_w[VAE_MAX_HIDDEN + 1], enc_b[VAE_MAX_HIDDEN + 1];
  int64_t dec_w[VAE_MAX_HIDDEN + 1], dec_b[VAE_MAX_HIDDEN + 1];
  int64_t sig_w[VAE_MAX_HIDDEN + 1], sig_b[VAE_MAX_HIDDEN + 1];
  int64_t off_eps_p, off_eps;       // off_eps = -1 when not tunable
  int64_t n_params;
  int max_width;
};
int make_layout(const VaeDesc* d, Layout* out);

}  // namespace vae
