// fused_small.cu -- persistent fused training loop for the single-Dense VAEs of the
// linear / sigmoid sweeps (seed_linpadding_expts.sh, sigmoid_vae_padding_expts.sh).
//
// One CTA (128 threads) owns one model for n_steps consecutive steps.  Everything the
// reference does per iteration of the hot loop (model.py:221-222) happens on-chip:
//   key splits (vae.py:125, datasets.py:51) -> eps = normal(key,(B,L+D)) (model.py:227)
//   -> dataset batch (datasets.py:183-195 / 240-249 / 75-84) -> VAE.apply
//   (networks.py:61-84) -> sampled ELBO (networks.py:94-98) -> hand-derived backward
//    -> legacy flax Adam (vae.py:113).
// Parameters, Adam moments and every activation live in shared memory for the whole
// launch; the only HBM traffic per step is the 4-byte loss store.
//
// Exact work skipping: generated batches have exactly-zero padding columns, so encoder
// kernel rows d >= DR receive a gradient of exactly 0 and Adam leaves them (and their
// zero moments) bit-for-bit unchanged; those rows are never loaded.
//
// Per step: phase A (all threads) threefry + normal draws into smem; phase B (thread per
// batch row) generator transform, forward, loss partials, backward to dmu; phase C
// (thread per parameter) dW = sum_b h*d, Adam.  Two CTA barriers per step.
#include "common.cuh"

namespace vae {

constexpr int kT = 128;

struct FusedJobDev {
  float* params; float* adam_m; float* adam_v; int32_t* step;
  uint32_t* model_key; uint32_t* dataset_key; const float* A; float* loss_out;
  int32_t batch; int32_t tunable; int32_t shape_id; int32_t pad_;
  float eps0, lr, b1, b2, omb1, omb2, adam_eps;
  float* dbg;
};

__host__ __device__ constexpr int up4(int n) { return (n + 3) & ~3; }
__host__ __device__ constexpr int oddi(int n) { return n | 1; }

template <int D_, int L_, int DR_, bool SIG_, int KIND_, int G_>
struct SmallCfg {
  static constexpr int D = D_, L = L_, DR = DR_, KIND = KIND_, G = G_;
  static constexpr bool SIG = SIG_;
  static constexpr int DD = KIND == VAE_DATA_SIGMOID ? DR - 1 : DR;   // generator -dd
  static constexpr int LP = up4(L), DP = up4(D);
  // padded on-chip parameter block (float offsets, all multiples of 4)
  static constexpr int oWe = 0;                          // [DR][LP]
  static constexpr int oBe = oWe + DR * LP;              // [LP]
  static constexpr int oWd = oBe + LP;                   // [L][DP]
  static constexpr int oBd = oWd + L * DP;               // [DP]
  static constexpr int oWs = oBd + DP;                   // [L][DP]   (SIG)
  static constexpr int oBs = oWs + (SIG ? L * DP : 0);   // [DP]      (SIG)
  static constexpr int oLv = oBs + (SIG ? DP : 0);       // [LP]  epsilon_p
  static constexpr int oEp = oLv + LP;                   // [4]   epsilon
  static constexpr int PP = oEp + 4;
  static constexpr int NDEC = SIG ? 2 : 1;
  static constexpr int NA = DR * L + L + (L * D + D) * NDEC + L + 1;   // active params
  static constexpr int NPT = (NA + kT - 1) / kT;
  static constexpr int AN = KIND == VAE_DATA_LINEAR_GAUSSIAN ? DD * G : (KIND == VAE_DATA_SIGMOID ? DD : 0);
  // activation row strides (odd => row-per-thread accesses are bank-conflict free)
  static constexpr int SZ = oddi(L + D), SG = oddi(G), SX = oddi(DR), SL = oddi(L), SD = oddi(D);
  // smem layout in floats given B rows
  struct Lay {
    int P, M, V, sl, el, kt, scal, one, red, A, keys, desc, z, gn, xs, zs, dx, dp, dm, q, total;
  };
  __host__ __device__ static Lay layout(int B) {
    Lay y; int o = 0;
    y.P = o; o += PP; y.M = o; o += PP; y.V = o; o += PP;
    y.sl = o; o += LP; y.el = o; o += LP;
    y.kt = o; o += 2 * LP;         // [2 slots][LP]  1 + lv - e^lv (KL term of the loss)
    y.scal = o; o += 8;            // [2 slots][e, sd, inv_var, -]
    y.one = o; o += 4;             // 1.0f (the "h" of bias-like gradients)
    y.red = o; o += 16;            // [3][4] warp partials
    y.A = o; o += up4(AN + 1);
    y.keys = o; o += 8;            // uint32 [2 slots][lat0 lat1 bat0 bat1]
    y.desc = o; o += NPT * kT * 2; // int2 per (k, thread)
    y.z = o; o += up4(B * SZ);
    y.gn = o; o += up4(B * SG);
    y.xs = o; o += up4(B * SX);
    y.zs = o; o += up4(B * SL);
    y.dx = o; o += up4(B * SD);
    y.dp = o; o += SIG ? up4(B * SD) : 0;
    y.dm = o; o += up4(B * SL);
    y.q = o; o += up4(B * SL);
    y.total = o;
    return y;
  }
  // flat (include/vae_b200.h) parameter index -> padded on-chip offset, -1 if never touched
  __device__ static int padded_of(int g, bool tunable) {
    if (g < D * L) { const int i = g / L, j = g - i * L; return i < DR ? oWe + i * LP + j : -1; }
    g -= D * L;
    if (g < L) return oBe + g;
    g -= L;
    if (g < L * D) { const int i = g / D, j = g - i * D; return oWd + i * DP + j; }
    g -= L * D;
    if (g < D) return oBd + g;
    g -= D;
    if (SIG) {
      if (g < L * D) { const int i = g / D, j = g - i * D; return oWs + i * DP + j; }
      g -= L * D;
      if (g < D) return oBs + g;
      g -= D;
    }
    if (g < L) return oLv + g;
    g -= L;
    return (tunable && g == 0) ? oEp : -1;
  }
  static constexpr int NFLAT = D * L + L + (L * D + D) * NDEC + L;   // + 1 if tunable
};

// descriptor packing: .x = h_off | d_off << 16 ; .y = poff | h_str << 16 | d_str << 23 | kind << 30
__device__ __forceinline__ int2 pack_desc(int h_off, int h_str, int d_off, int d_str, int poff, int kind) {
  return make_int2(h_off | (d_off << 16), poff | (h_str << 16) | (d_str << 23) | (kind << 30));
}

template <class C>
__device__ __forceinline__ int2 make_desc(int a, const typename C::Lay& y, bool tunable) {
  constexpr int D = C::D, L = C::L, DR = C::DR;
  const int ones = y.one;
  if (a < DR * L) { const int i = a / L, j = a - i * L; return pack_desc(y.xs + i, C::SX, y.dm + j, C::SL, C::oWe + i * C::LP + j, 0); }
  a -= DR * L;
  if (a < L) return pack_desc(ones, 0, y.dm + a, C::SL, C::oBe + a, 0);
  a -= L;
  if (a < L * D) { const int i = a / D, j = a - i * D; return pack_desc(y.zs + i, C::SL, y.dx + j, C::SD, C::oWd + i * C::DP + j, 0); }
  a -= L * D;
  if (a < D) return pack_desc(ones, 0, y.dx + a, C::SD, C::oBd + a, 0);
  a -= D;
  if constexpr (C::SIG) {
    if (a < L * D) { const int i = a / D, j = a - i * D; return pack_desc(y.zs + i, C::SL, y.dp + j, C::SD, C::oWs + i * C::DP + j, 0); }
    a -= L * D;
    if (a < D) return pack_desc(ones, 0, y.dp + a, C::SD, C::oBs + a, 0);
    a -= D;
  }
  if (a < L) return pack_desc(ones, 0, y.q + a, C::SL, C::oLv + a, 1);
  a -= L;
  if (a == 0 && tunable) return pack_desc(ones, 0, ones, 0, C::oEp, 2);
  return pack_desc(ones, 0, ones, 0, 0, 3);     // kind 3: nothing to do
}

template <class C>
__device__ void fused_small_run(const FusedJobDev& job, int64_t n_steps, float* sm) {
  constexpr int D = C::D, L = C::L, DR = C::DR, DD = C::DD, G = C::G, LP = C::LP, DP = C::DP;
  constexpr int LD = L + D;
  const int t = threadIdx.x;
  const int B = job.batch;
  const bool tunable = job.tunable != 0;
  const typename C::Lay y = C::layout(B);
  float* P = sm + y.P; float* M = sm + y.M; float* V = sm + y.V;
  uint32_t* keys = reinterpret_cast<uint32_t*>(sm + y.keys);
  int2* desc = reinterpret_cast<int2*>(sm + y.desc);

  // ---- load state
  for (int i = t; i < 3 * C::PP; i += kT) sm[y.P + i] = 0.f;
  __syncthreads();
  const int nflat = C::NFLAT + (tunable ? 1 : 0);
  for (int g = t; g < nflat; g += kT) {
    const int o = C::padded_of(g, tunable);
    if (o >= 0) { P[o] = job.params[g]; M[o] = job.adam_m[g]; V[o] = job.adam_v[g]; }
  }
  for (int i = t; i < C::AN; i += kT) sm[y.A + i] = job.A[i];
#pragma unroll
  for (int k = 0; k < C::NPT; ++k) desc[k * kT + t] = make_desc<C>(k * kT + t, y, tunable);
  if (t == 0) sm[y.one] = 1.0f;
  __syncthreads();
  // Derived per-step constants are double-buffered by step parity: phase B/C of step s
  // read slot s&1, the owning thread publishes the post-update values into the other slot.
  if (t < L) {
    const float lv = P[C::oLv + t], elv = expf(lv);
    sm[y.sl + t] = expf(0.5f * lv); sm[y.el + t] = elv;
    sm[y.kt + t] = 1.f + lv - elv; sm[y.kt + LP + t] = 1.f + lv - elv;
  }
  if (t == 0) {
    const float e = tunable ? P[C::oEp] * job.eps0 : job.eps0;
    for (int q = 0; q < 2; ++q) { sm[y.scal + 4 * q] = e; sm[y.scal + 4 * q + 1] = expf(0.5f * e); sm[y.scal + 4 * q + 2] = expf(-e); }
  }
  // Adam bias-correction powers (double running products: error ~ t * 1e-16)
  const int32_t step0 = job.step[0];
  double b1t = pow((double)job.b1, (double)step0), b2t = pow((double)job.b2, (double)step0);

  // ---- key chains: lanes 0,1 own the model-side split, lanes 2,3 the dataset-side one
  uint32_t ck0 = 0, ck1 = 0;           // carried key of this lane's chain
  if (t < 4) {
    const uint32_t* src = t < 2 ? job.model_key : job.dataset_key;
    ck0 = src[0]; ck1 = src[1];
  }
  auto advance_keys = [&](int slot) {   // call with t < 4, converged lanes 0..3
    uint32_t x0 = (uint32_t)(t & 1), x1 = (uint32_t)(t & 1) + 2u;
    threefry2x32(ck0, ck1, x0, x1);     // x0 = row0[t&1], x1 = row1[t&1]
    // model chain: row 0 consumed (latent key), row 1 carried (vae.py:125)
    // dataset chain: row 0 carried, row 1 returned (datasets.py:51)
    const uint32_t used = t < 2 ? x0 : x1;
    const uint32_t carry = t < 2 ? x1 : x0;
    const uint32_t other = __shfl_xor_sync(0xFu, carry, 1);
    ck0 = (t & 1) ? other : carry;
    ck1 = (t & 1) ? carry : other;
    keys[slot * 4 + t] = used;          // [lat0 lat1 bat0 bat1]
  };
  if (t < 4 && n_steps > 0) advance_keys(0);
  __syncthreads();

  const float invB = 1.0f / (float)B;
  const int half_e = (B * LD + 1) >> 1, n_e = B * LD;
  const int half_d = (B * G + 1) >> 1, n_d = B * G;

  for (int64_t s = 0; s < n_steps; ++s) {
    const int slot = (int)(s & 1);
    // ================= phase A: random draws
    {
      const uint32_t lk0 = keys[slot * 4 + 0], lk1 = keys[slot * 4 + 1];
      const uint32_t bk0 = keys[slot * 4 + 2], bk1 = keys[slot * 4 + 3];
      if (t < 4 && s + 1 < n_steps) advance_keys(slot ^ 1);
      float* z = sm + y.z; float* gn = sm + y.gn;
      for (int w = t; w < half_e + half_d; w += kT) {
        if (w < half_e) {
          const int j = w, j2 = w + half_e;
          uint32_t x0 = (uint32_t)j, x1 = j2 < n_e ? (uint32_t)j2 : 0u;
          threefry2x32(lk0, lk1, x0, x1);
          { const int r = j / LD; z[r * C::SZ + (j - r * LD)] = normal_from_bits(x0); }
          if (j2 < n_e) { const int r = j2 / LD; z[r * C::SZ + (j2 - r * LD)] = normal_from_bits(x1); }
        } else {
          const int j = w - half_e, j2 = j + half_d;
          uint32_t x0 = (uint32_t)j, x1 = j2 < n_d ? (uint32_t)j2 : 0u;
          threefry2x32(bk0, bk1, x0, x1);
          { const int r = j / G; gn[r * C::SG + (j - r * G)] = normal_from_bits(x0); }
          if (j2 < n_d) { const int r = j2 / G; gn[r * C::SG + (j2 - r * G)] = normal_from_bits(x1); }
        }
      }
    }
    __syncthreads();
    // ================= phase B: one thread per batch row
    float kl_acc = 0.f, r2_acc = 0.f, rz_acc = 0.f;
    {
      const float sd = sm[y.scal + 4 * slot + 1], inv_var = sm[y.scal + 4 * slot + 2];
      const float gscale = inv_var * invB;
      for (int b = t; b < B; b += kT) {
        const float* zr = sm + y.z + b * C::SZ;
        const float* gr = sm + y.gn + b * C::SG;
        // ---- generator transform (bit-exact with oracle/datasets.py)
        float x[DR];
        if (C::KIND == VAE_DATA_SPHERE) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < DD; ++j) { x[j] = gr[j]; const float sq = __fmul_rn(x[j], x[j]); acc = j == 0 ? sq : __fadd_rn(acc, sq); }
          const float rad = __fsqrt_rn(acc);
#pragma unroll
          for (int j = 0; j < DD; ++j) x[j] = __fdiv_rn(x[j], rad);
        } else if (C::KIND == VAE_DATA_LINEAR_GAUSSIAN) {
          const float* A = sm + y.A;
#pragma unroll
          for (int i = 0; i < DD; ++i) {
            float acc = __fmul_rn(A[i * G], gr[0]);
#pragma unroll
            for (int k = 1; k < G; ++k) acc = __fadd_rn(acc, __fmul_rn(A[i * G + k], gr[k]));
            x[i] = acc;
          }
        } else {
          const float* A = sm + y.A;
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < DD; ++j) { x[j] = gr[j]; const float pr = __fmul_rn(x[j], A[j]); acc = j == 0 ? pr : __fadd_rn(acc, pr); }
          x[DD] = det_sigmoid(acc);
        }
        float* xrow = sm + y.xs + b * C::SX;
#pragma unroll
        for (int d = 0; d < DR; ++d) xrow[d] = x[d];
        // ---- encoder: mu = x We + be   (only the DR non-padding columns contribute)
        float mu[LP];
#pragma unroll
        for (int q = 0; q < LP / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(P + C::oBe + 4 * q);
          mu[4 * q] = v.x; mu[4 * q + 1] = v.y; mu[4 * q + 2] = v.z; mu[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int d = 0; d < DR; ++d) {
#pragma unroll
          for (int q = 0; q < LP / 4; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(P + C::oWe + d * LP + 4 * q);
            mu[4 * q] = fmaf(x[d], w.x, mu[4 * q]); mu[4 * q + 1] = fmaf(x[d], w.y, mu[4 * q + 1]);
            mu[4 * q + 2] = fmaf(x[d], w.z, mu[4 * q + 2]); mu[4 * q + 3] = fmaf(x[d], w.w, mu[4 * q + 3]);
          }
        }
        // ---- reparameterise, decoder(s)
        float dec[DP], sp[C::SIG ? DP : 1];
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(P + C::oBd + 4 * q);
          dec[4 * q] = v.x; dec[4 * q + 1] = v.y; dec[4 * q + 2] = v.z; dec[4 * q + 3] = v.w;
          if constexpr (C::SIG) {
            const float4 u = *reinterpret_cast<const float4*>(P + C::oBs + 4 * q);
            sp[4 * q] = u.x; sp[4 * q + 1] = u.y; sp[4 * q + 2] = u.z; sp[4 * q + 3] = u.w;
          }
        }
        float* zsrow = sm + y.zs + b * C::SL;
#pragma unroll
        for (int l = 0; l < L; ++l) {
          kl_acc = fmaf(mu[l], mu[l], kl_acc);
          const float zsl = fmaf(sm[y.sl + l], zr[l], mu[l]);
          zsrow[l] = zsl;
#pragma unroll
          for (int q = 0; q < DP / 4; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(P + C::oWd + l * DP + 4 * q);
            dec[4 * q] = fmaf(zsl, w.x, dec[4 * q]); dec[4 * q + 1] = fmaf(zsl, w.y, dec[4 * q + 1]);
            dec[4 * q + 2] = fmaf(zsl, w.z, dec[4 * q + 2]); dec[4 * q + 3] = fmaf(zsl, w.w, dec[4 * q + 3]);
            if constexpr (C::SIG) {
              const float4 u = *reinterpret_cast<const float4*>(P + C::oWs + l * DP + 4 * q);
              sp[4 * q] = fmaf(zsl, u.x, sp[4 * q]); sp[4 * q + 1] = fmaf(zsl, u.y, sp[4 * q + 1]);
              sp[4 * q + 2] = fmaf(zsl, u.z, sp[4 * q + 2]); sp[4 * q + 3] = fmaf(zsl, u.w, sp[4 * q + 3]);
            }
          }
        }
        // ---- output, residual, d x_hat  (dec[] becomes d x_hat, sp[] becomes d sig_pre)
        float* dxrow = sm + y.dx + b * C::SD;
        float* dprow = sm + y.dp + b * C::SD;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          if (d < D) {
            const float z2 = zr[L + d];
            float xh = dec[d];
            float sg = 0.f;
            if constexpr (C::SIG) { sg = 1.f / (1.f + expf(-sp[d])); xh += sg; }
            xh = fmaf(z2, sd, xh);
            const float r = xh - (d < DR ? x[d < DR ? d : 0] : 0.f);
            r2_acc = fmaf(r, r, r2_acc);
            rz_acc = fmaf(r, z2, rz_acc);
            const float g = r * gscale;
            dec[d] = g; dxrow[d] = g;
            if constexpr (C::SIG) { const float gp = g * sg * (1.f - sg); sp[d] = gp; dprow[d] = gp; }
          } else {
            dec[d] = 0.f;
            if constexpr (C::SIG) sp[d] = 0.f;
          }
        }
        // ---- dzs = dxhat Wd^T (+ dpre Ws^T); dmu = dzs + mu/B; q = dzs * z1
        float* dmrow = sm + y.dm + b * C::SL;
        float* qrow = sm + y.q + b * C::SL;
#pragma unroll
        for (int l = 0; l < L; ++l) {
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int q = 0; q < DP / 4; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(P + C::oWd + l * DP + 4 * q);
            a0 = fmaf(dec[4 * q], w.x, a0); a1 = fmaf(dec[4 * q + 1], w.y, a1);
            a0 = fmaf(dec[4 * q + 2], w.z, a0); a1 = fmaf(dec[4 * q + 3], w.w, a1);
            if constexpr (C::SIG) {
              const float4 u = *reinterpret_cast<const float4*>(P + C::oWs + l * DP + 4 * q);
              a0 = fmaf(sp[4 * q], u.x, a0); a1 = fmaf(sp[4 * q + 1], u.y, a1);
              a0 = fmaf(sp[4 * q + 2], u.z, a0); a1 = fmaf(sp[4 * q + 3], u.w, a1);
            }
          }
          const float dz = a0 + a1;
          dmrow[l] = fmaf(mu[l], invB, dz);
          qrow[l] = dz * zr[l];
        }
      }
      kl_acc = warp_sum(kl_acc); r2_acc = warp_sum(r2_acc); rz_acc = warp_sum(rz_acc);
      if ((t & 31) == 0) {
        sm[y.red + 0 + (t >> 5)] = kl_acc; sm[y.red + 4 + (t >> 5)] = r2_acc; sm[y.red + 8 + (t >> 5)] = rz_acc;
      }
    }
    __syncthreads();
#ifdef VAE_FUSED_DEBUG
    if (s == 0 && job.dbg) {
      for (int i = t; i < y.total; i += kT) job.dbg[i] = sm[i];
      if (t == 0) { int* q = (int*)(job.dbg + y.total); q[0] = y.z; q[1] = y.gn; q[2] = y.xs; q[3] = y.zs; q[4] = y.dx; q[5] = y.dm; q[6] = y.red; q[7] = y.P; }
    }
    __syncthreads();
#endif
    // ================= phase C: one thread per parameter: gradient + Adam
    {
      b1t *= (double)job.b1; b2t *= (double)job.b2;
      const float ibc1 = 1.0f / (1.0f - (float)b1t), ibc2 = 1.0f / (1.0f - (float)b2t);
      const float e = sm[y.scal + 4 * slot], sd = sm[y.scal + 4 * slot + 1], inv_var = sm[y.scal + 4 * slot + 2];
      const float s_mu2 = (sm[y.red + 0] + sm[y.red + 1]) + (sm[y.red + 2] + sm[y.red + 3]);
      const float s_r2 = (sm[y.red + 4] + sm[y.red + 5]) + (sm[y.red + 6] + sm[y.red + 7]);
      const float s_rz = (sm[y.red + 8] + sm[y.red + 9]) + (sm[y.red + 10] + sm[y.red + 11]);
      const float fB = (float)B;
      if (t == 32) {   // loss = mean_b(Dkl + mse)   (networks.py:94-98)
        float klc = 0.f;
#pragma unroll
        for (int l = 0; l < L; ++l) klc += sm[y.kt + slot * LP + l];
        const float dkl = (-0.5f * klc * fB + 0.5f * s_mu2) * invB;
        const float mse = (0.5f * inv_var * s_r2 + 0.5f * (1.8378770664093453f + e) * fB * (float)D) * invB;
        job.loss_out[s] = dkl + mse;
      }
      float newp[C::NPT]; int newo[C::NPT]; int newk[C::NPT];
#pragma unroll
      for (int k = 0; k < C::NPT; ++k) {
        const int2 dsc = desc[k * kT + t];
        const int h_off = dsc.x & 0xFFFF, d_off = (dsc.x >> 16) & 0xFFFF;
        const int poff = dsc.y & 0xFFFF, h_str = (dsc.y >> 16) & 0x7F, d_str = (dsc.y >> 23) & 0x7F;
        const int kind = (dsc.y >> 30) & 3;
        newk[k] = kind; newo[k] = poff; newp[k] = 0.f;
        if (kind == 3) continue;
        const float* hp = sm + h_off; const float* dp = sm + d_off;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int b = kind == 2 ? B : 0;
        for (; b + 4 <= B; b += 4) {
          a0 = fmaf(hp[0], dp[0], a0);
          a1 = fmaf(hp[h_str], dp[d_str], a1);
          a2 = fmaf(hp[2 * h_str], dp[2 * d_str], a2);
          a3 = fmaf(hp[3 * h_str], dp[3 * d_str], a3);
          hp += 4 * h_str; dp += 4 * d_str;
        }
        for (; b < B; ++b) { a0 = fmaf(hp[0], dp[0], a0); hp += h_str; dp += d_str; }
        float g = (a0 + a1) + (a2 + a3);
        if (kind == 1) {          // d loss / d epsilon_p[l]
          const int l = poff - C::oLv;
          g = g * 0.5f * sm[y.sl + l] - 0.5f * (1.f - sm[y.el + l]);
        } else if (kind == 2) {   // d loss / d epsilon (param): eps0 * dL/de
          g = job.eps0 * invB * (-0.5f * inv_var * s_r2 + 0.5f * fB * (float)D + 0.5f * sd * inv_var * s_rz);
        }
        const float m = job.b1 * M[poff] + job.omb1 * g;
        const float v = job.b2 * V[poff] + job.omb2 * g * g;
        M[poff] = m; V[poff] = v;
        newp[k] = P[poff] - job.lr * (m * ibc1) / (sqrtf(v * ibc2) + job.adam_eps);
      }
      // Phase C reads activations, this thread's own P/M/V/sl/el entries and the
      // slot-parity constants only, so new values can be published without a barrier.
#pragma unroll
      for (int k = 0; k < C::NPT; ++k) {
        if (newk[k] == 3) continue;
        P[newo[k]] = newp[k];
        if (newk[k] == 1) {
          const int l = newo[k] - C::oLv;
          const float elv = expf(newp[k]);
          sm[y.sl + l] = expf(0.5f * newp[k]); sm[y.el + l] = elv;
          sm[y.kt + (slot ^ 1) * LP + l] = 1.f + newp[k] - elv;
        }
        if (newk[k] == 2) {
          const float e2 = newp[k] * job.eps0;
          float* sc = sm + y.scal + 4 * (slot ^ 1);
          sc[0] = e2; sc[1] = expf(0.5f * e2); sc[2] = expf(-e2);
        }
      }
    }
    // no barrier here: phase A of the next step touches only z / gn / keys, which phase C
    // does not read; the A->B barrier orders these parameter writes before the next reads.
  }
  __syncthreads();
  // ---- store state
  for (int g = t; g < nflat; g += kT) {
    const int o = C::padded_of(g, tunable);
    if (o >= 0) { job.params[g] = P[o]; job.adam_m[g] = M[o]; job.adam_v[g] = V[o]; }
  }
  if (t == 0) job.step[0] = step0 + (int32_t)n_steps;
  if (t < 4) {
    uint32_t* dst = t < 2 ? job.model_key : job.dataset_key;
    if ((t & 1) == 0) { dst[0] = ck0; dst[1] = ck1; }
  }
}

// (id, D, L, DR, SIG, KIND, G): every single-Dense shape of the two sweeps, plus one
// sphere shape without hidden layers (generator coverage).
#define VAE_SMALL_SHAPES(X)                                   \
  X(0, 12, 20, 3, false, VAE_DATA_LINEAR_GAUSSIAN, 3)         \
  X(1, 20, 20, 3, false, VAE_DATA_LINEAR_GAUSSIAN, 3)         \
  X(2, 12, 20, 6, false, VAE_DATA_LINEAR_GAUSSIAN, 3)         \
  X(3, 20, 20, 6, false, VAE_DATA_LINEAR_GAUSSIAN, 3)         \
  X(4, 12, 20, 9, false, VAE_DATA_LINEAR_GAUSSIAN, 3)         \
  X(5, 20, 10, 9, false, VAE_DATA_LINEAR_GAUSSIAN, 3)         \
  X(6, 20, 10, 12, false, VAE_DATA_LINEAR_GAUSSIAN, 3)        \
  X(7, 7, 6, 4, true, VAE_DATA_SIGMOID, 3)                    \
  X(8, 17, 8, 4, true, VAE_DATA_SIGMOID, 3)                   \
  X(9, 22, 16, 6, true, VAE_DATA_SIGMOID, 5)                  \
  X(10, 11, 10, 6, true, VAE_DATA_SIGMOID, 5)                 \
  X(11, 15, 13, 8, true, VAE_DATA_SIGMOID, 7)                 \
  X(12, 28, 24, 8, true, VAE_DATA_SIGMOID, 7)                 \
  X(13, 6, 6, 3, false, VAE_DATA_SPHERE, 3)

__global__ void __launch_bounds__(kT)
fused_small_kernel(const FusedJobDev* __restrict__ jobs, int64_t n_steps) {
  extern __shared__ __align__(16) float sm[];
  const FusedJobDev job = jobs[blockIdx.x];
  switch (job.shape_id) {
#define X(id, D, L, DR, SIG, KIND, G) \
    case id: fused_small_run<SmallCfg<D, L, DR, SIG, KIND, G>>(job, n_steps, sm); break;
    VAE_SMALL_SHAPES(X)
#undef X
    default: break;
  }
}

// Returns the shape id (>= 0) or -1; *smem_floats receives the CTA's shared memory need.
static int small_shape_of(const VaeTrainJob* j, int* smem_floats) {
  const VaeDesc& d = j->desc;
  const VaeDataDesc& g = j->data;
  if (d.n_enc_hidden != 0 || d.n_dec_hidden != 0) return -1;
  if (d.batch <= 0 || d.batch > 4096) return -1;
  const int width = g.dim + g.padding_dim + (g.kind == VAE_DATA_SIGMOID ? 1 : 0);
  if (width != d.data_dim) return -1;
  if ((d.sig_decoder != 0) != (g.kind == VAE_DATA_SIGMOID)) return -1;
  const int dr = g.dim + (g.kind == VAE_DATA_SIGMOID ? 1 : 0);
  const int gg = g.kind == VAE_DATA_LINEAR_GAUSSIAN ? g.intrinsic_dim : g.dim;
#define X(id, D, L, DR, SIG, KIND, G)                                                        \
  if (d.data_dim == D && d.latent_dim == L && dr == DR && g.kind == KIND && gg == G) {       \
    *smem_floats = SmallCfg<D, L, DR, SIG, KIND, G>::layout(d.batch).total;                  \
    return id;                                                                               \
  }
  VAE_SMALL_SHAPES(X)
#undef X
  return -1;
}

constexpr int kMaxSmem = 227 * 1024;

}  // namespace vae

extern "C" {

int vae_train_fused_supported(const VaeTrainJob* job) {
  if (!job) return 0;
  int fl = 0;
  const int id = vae::small_shape_of(job, &fl);
  if (id < 0 || (size_t)fl * 4 > (size_t)vae::kMaxSmem) return 0;
  return 1;
}

size_t vae_train_fused_workspace_bytes(const VaeTrainJob* job) { (void)job; return 0; }

size_t vae_train_fused_scratch_bytes(int32_t n_jobs) {
  return n_jobs > 0 ? (size_t)n_jobs * sizeof(vae::FusedJobDev) : 0;
}

int vae_train_fused(const VaeTrainJob* jobs, int32_t n_jobs, int64_t n_steps, void* d_scratch,
                    size_t scratch_bytes, void* stream) {
  VAE_REQUIRE(jobs && n_jobs > 0 && n_steps >= 0, "vae_train_fused: bad arguments");
  VAE_REQUIRE(d_scratch && scratch_bytes >= vae_train_fused_scratch_bytes(n_jobs),
              "vae_train_fused: scratch too small (need %zu bytes)", vae_train_fused_scratch_bytes(n_jobs));
  if (n_steps == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  vae::FusedJobDev* host = (vae::FusedJobDev*)malloc((size_t)n_jobs * sizeof(vae::FusedJobDev));
  VAE_REQUIRE(host, "vae_train_fused: out of host memory");
  int max_fl = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const VaeTrainJob& j = jobs[i];
    int fl = 0;
    const int id = vae::small_shape_of(&j, &fl);
    if (id < 0 || (size_t)fl * 4 > (size_t)vae::kMaxSmem) {
      free(host);
      vae::set_error("vae_train_fused: job %d has no fused kernel (D=%d L=%d hidden=%d/%d)", i,
                     j.desc.data_dim, j.desc.latent_dim, j.desc.n_enc_hidden, j.desc.n_dec_hidden);
      return VAE_ERR_UNSUPPORTED;
    }
    if (!j.params || !j.adam_m || !j.adam_v || !j.step || !j.model_key || !j.dataset_key || !j.loss_out ||
        (j.data.kind != VAE_DATA_SPHERE && !j.A)) {
      free(host);
      vae::set_error("vae_train_fused: job %d has a NULL device pointer", i);
      return VAE_ERR_INVALID;
    }
    if (fl > max_fl) max_fl = fl;
    vae::FusedJobDev& o = host[i];
    o.params = j.params; o.adam_m = j.adam_m; o.adam_v = j.adam_v; o.step = j.step;
    o.model_key = j.model_key; o.dataset_key = j.dataset_key; o.A = j.A; o.loss_out = j.loss_out;
    o.batch = j.desc.batch; o.tunable = j.desc.tunable_decoder_var; o.shape_id = id; o.pad_ = 0; o.dbg = (float*)j.workspace;
    const vae::AdamHyper hp = vae::make_adam_hyper(&j.desc);
    o.eps0 = (float)j.desc.epsilon0; o.lr = hp.lr; o.b1 = hp.b1; o.b2 = hp.b2; o.omb1 = hp.omb1; o.omb2 = hp.omb2;
    o.adam_eps = hp.eps;
  }
  // pageable-source async copies are staged before the call returns, so `host` may be freed
  cudaError_t e = cudaMemcpyAsync(d_scratch, host, (size_t)n_jobs * sizeof(vae::FusedJobDev),
                                  cudaMemcpyHostToDevice, s);
  free(host);
  VAE_CUDA_OK(e);
  const size_t smem = (size_t)max_fl * sizeof(float);
  VAE_CUDA_OK(cudaFuncSetAttribute(vae::fused_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vae::fused_small_kernel<<<n_jobs, vae::kT, smem, s>>>((const vae::FusedJobDev*)d_scratch, n_steps);
  VAE_LAUNCH_OK("fused_small_kernel");
  return 0;
}

}  // extern "C"
