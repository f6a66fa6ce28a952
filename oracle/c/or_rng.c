/* ORACLE (test infrastructure) -- see vae_oracle.h.  Compiled with -ffp-contract=off:
 * every float operation here is a single IEEE binary32 op (or an explicit fmaf) in the exact order of
 * oracle/jax_random.py and oracle/datasets.py, so results are bit-identical to them.
 *
 * Restates third-party jax.random (un-vendored, un-pinned; README.md:8-13) as used at
 * model.py:29,140,154,208,227, vae.py:56,125, datasets.py:35,51,76,173,185,237,242, and
 * the generators datasets.py:75-84, 183-195, 240-249. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "vae_oracle.h"

static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

#define TF_ROUND(r) x0 += x1; x1 = rotl(x1, r); x1 ^= x0;
static inline void tf(uint32_t k0, uint32_t k1, uint32_t* px0, uint32_t* px1) {
  uint32_t x0 = *px0, x1 = *px1;
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0; x1 += k1;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
  *px0 = x0; *px1 = x1;
}

void or_threefry2x32(uint32_t k0, uint32_t k1, uint32_t* x0, uint32_t* x1) { tf(k0, k1, x0, x1); }

/* _random_bits over iota(n): block j consumes counters (j, j+half) (oracle: threefry_counts) */
void or_random_bits(const uint32_t key[2], uint32_t* out, int64_t n) {
  const int64_t half = (n + 1) / 2;
  const uint32_t k0 = key[0], k1 = key[1];
  for (int64_t j = 0; j < half; ++j) {
    uint32_t x0 = (uint32_t)j, x1 = (j + half < n) ? (uint32_t)(j + half) : 0u;
    tf(k0, k1, &x0, &x1);
    out[j] = x0;
    if (j + half < n) out[j + half] = x1;
  }
}

void or_split(const uint32_t key[2], uint32_t* out, int32_t num) { or_random_bits(key, out, 2 * (int64_t)num); }

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* oracle/jax_random.py: neg_log1p_neg */
static inline float neg_log1p_neg(float t) {
  const float a = 1.0f - t;
  const float c = (1.0f - a) - t;
  const uint32_t bits = f2u(a);
  int k = (int)(bits >> 23) - 127;
  uint32_t mb = (bits & 0x007FFFFFu) | 0x3F800000u;
  const int big = mb >= 0x3FB504F3u;
  mb = big ? mb - 0x00800000u : mb;
  k += big;
  const float m = u2f(mb);
  const float inv_scale = u2f((uint32_t)(127 - k) << 23);
  const float f = fmaf(c, inv_scale, m - 1.0f);
  const float fk = (float)k;
  const float z = f * f;
  float p = 7.0376836292e-2f;
  p = fmaf(p, f, -1.1514610310e-1f);
  p = fmaf(p, f, 1.1676998740e-1f);
  p = fmaf(p, f, -1.2420140846e-1f);
  p = fmaf(p, f, 1.4249322787e-1f);
  p = fmaf(p, f, -1.6668057665e-1f);
  p = fmaf(p, f, 2.0000714765e-1f);
  p = fmaf(p, f, -2.4999993993e-1f);
  p = fmaf(p, f, 3.3333331174e-1f);
  float y = (p * f) * z;
  y = fmaf(fk, -2.12194440e-4f, y);
  y = fmaf(-0.5f, z, y);
  float r = f + y;
  r = fmaf(fk, 0.693359375f, r);
  return -r;
}

static const float ERFINV_C[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f,
                                  0.00021858087f, -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
static const float ERFINV_T[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f,
                                  0.00573950773f, -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};

static inline float erf_inv(float u) {
  const float w = neg_log1p_neg(u * u);
  const int central = w < 5.0f;
  const float ww = central ? w - 2.5f : sqrtf(w) - 3.0f;
  const float* c = central ? ERFINV_C : ERFINV_T;
  float p = c[0];
  for (int i = 1; i < 9; ++i) p = fmaf(p, ww, c[i]);
  return p * u;
}

static inline float normal_from_bits(uint32_t bits) {
  const float lo = -0.99999994f;
  const float f = u2f((bits >> 9) | 0x3F800000u) - 1.0f;
  float u = f * 2.0f + lo;
  u = u < lo ? lo : u;
  return 1.41421354f * erf_inv(u);
}

void or_normal(const uint32_t key[2], float* out, int64_t n) {
  const int64_t half = (n + 1) / 2;
  const uint32_t k0 = key[0], k1 = key[1];
  for (int64_t j = 0; j < half; ++j) {
    uint32_t x0 = (uint32_t)j, x1 = (j + half < n) ? (uint32_t)(j + half) : 0u;
    tf(k0, k1, &x0, &x1);
    out[j] = normal_from_bits(x0);
    if (j + half < n) out[j + half] = normal_from_bits(x1);
  }
}

/* oracle/datasets.py: det_expf / det_sigmoid */
static inline float det_expf(float x) {
  const float n = floorf(1.44269504088896341f * x + 0.5f);
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  const float z = r * r;
  float p = 1.9875691500e-4f;
  p = p * r + 1.3981999507e-3f;
  p = p * r + 8.3334519073e-3f;
  p = p * r + 4.1665795894e-2f;
  p = p * r + 1.6666665459e-1f;
  p = p * r + 5.0000001201e-1f;
  float y = p * z;
  y = y + r;
  y = y + 1.0f;
  return y * u2f((uint32_t)((int)n + 127) << 23);
}
static inline float det_sigmoid(float x) { return 1.0f / (1.0f + det_expf(-x)); }

int32_t or_data_width(const OrData* d) { return d->dim + d->padding_dim + (d->kind == 2 ? 1 : 0); }

/* datasets.py:75-84 (sphere), :183-195 (linear_gaussian), :240-249 (sigmoid) */
void or_dataset_batch(const OrData* d, const uint32_t key[2], const float* A, int32_t size, float* out) {
  const int g = d->kind == 1 ? d->intrinsic_dim : d->dim;
  const int width = or_data_width(d);
  float* nrm = (float*)malloc(sizeof(float) * (size_t)size * g + 4);
  or_normal(key, nrm, (int64_t)size * g);
  memset(out, 0, sizeof(float) * (size_t)size * width);
  for (int b = 0; b < size; ++b) {
    const float* gr = nrm + (size_t)b * g;
    float* o = out + (size_t)b * width;
    if (d->kind == 0) {
      float acc = gr[0] * gr[0];
      for (int j = 1; j < d->dim; ++j) acc = acc + gr[j] * gr[j];
      const float rad = sqrtf(acc);
      for (int j = 0; j < d->dim; ++j) o[j] = gr[j] / rad;
    } else if (d->kind == 1) {
      for (int i = 0; i < d->dim; ++i) {
        float acc = A[i * g] * gr[0];
        for (int k = 1; k < g; ++k) acc = acc + A[i * g + k] * gr[k];
        o[i] = acc;
      }
    } else {
      float acc = gr[0] * A[0];
      o[0] = gr[0];
      for (int j = 1; j < d->dim; ++j) { o[j] = gr[j]; acc = acc + gr[j] * A[j]; }
      o[d->dim] = det_sigmoid(acc);
    }
  }
  free(nrm);
}
