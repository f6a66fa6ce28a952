// step_generic.cu -- shape-generic VAE step: one kernel per logical op, any layer
// sizes / batch.  This is the path behind vae_train_step / vae_loss_and_grad /
// vae_loss / vae_sample / vae_adam_update; the fused persistent kernels
// (fused_small.cu, fused_cluster.cu) specialise it for the sweep shapes.
//
// Reference: networks.py:26-44 (MLP), :59-84 (VAE.apply), :87-101 (train_step),
// :103-113 (loss); legacy flax.optim.Adam (vae.py:113).  Backward is hand-derived
// (SURVEY.md Appendix B), checked against oracle/vae_step.py.
#include "common.cuh"

namespace vae {

// ------------------------------------------------------------------ tiled SGEMM
// C[M,N] (ldc) = epilogue( sum_k A(i,k) * B(k,j) ),  A(i,k) = A[i*sAm + k*sAk],
// B(k,j) = B[k*sBk + j*sBn].  FP32 FFMA, 64x64x16 tiles, 4x4 register blocking.
enum { GEMM_BIAS = 1, GEMM_RELU = 2, GEMM_MASK = 4, GEMM_ACC = 8 };

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int64_t sAm, int64_t sAk,
             const float* __restrict__ B, int64_t sBk, int64_t sBn, float* __restrict__ C,
             int64_t ldc, const float* __restrict__ bias, const float* __restrict__ mask,
             int64_t ldm, int flags) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (sAk == 1);
  const bool b_nfast = (sBn == 1);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * 256;
      int m, k;
      if (a_kfast) { m = e >> 4; k = e & 15; } else { k = e >> 6; m = e & 63; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? A[(int64_t)gm * sAm + (int64_t)gk * sAk] : 0.f;
      int n, kb;
      if (b_nfast) { kb = e >> 6; n = e & 63; } else { n = e >> 4; kb = e & 15; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < N && gkb < K) ? B[(int64_t)gkb * sBk + (int64_t)gn * sBn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (flags & GEMM_BIAS) v += bias[gn];
      if (flags & GEMM_RELU) v = fmaxf(v, 0.f);
      if (flags & GEMM_MASK) v = mask[(int64_t)gm * ldm + gn] > 0.f ? v : 0.f;
      float* c = C + (int64_t)gm * ldc + gn;
      if (flags & GEMM_ACC) v += *c;
      *c = v;
    }
  }
}

constexpr int kColParts = 64;
// column sums with the rows split over up to kColParts blocks when there are many rows
static int colsum(cudaStream_t s, const float* Y, int64_t ldy, const float* Z, int64_t ldz, int rows, int n, float* out,
                  float* scratch /* kColParts * n floats, or nullptr */);

static int gemm(cudaStream_t s, int M, int N, int K, const float* A, int64_t sAm, int64_t sAk,
                const float* B, int64_t sBk, int64_t sBn, float* C, int64_t ldc,
                const float* bias, const float* mask, int64_t ldm, int flags) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  sgemm_kernel<<<grid, 256, 0, s>>>(M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, bias, mask, ldm, flags);
  VAE_LAUNCH_OK("sgemm_kernel");
  return 0;
}

// out[j] = sum_r Y[r*ldy + j] * (Z ? Z[r*ldz + j] : 1)      (bias grads, dlogvar sums)
// gridDim.y > 1: block y reduces its contiguous slice of the rows into out[y*n + j]
// (deterministic two-stage reduction; the second stage is this kernel again on the partials).
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ Z, int64_t ldz,
              int rows, int n, float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int rbeg = blockIdx.y * per, rend = min(rows, rbeg + per);
  out += (int64_t)blockIdx.y * n;
  float acc = 0.f;
  if (col < n) {
    for (int r = rbeg + ty; r < rend; r += 8) {
      float v = Y[(int64_t)r * ldy + col];
      if (Z) v *= Z[(int64_t)r * ldz + col];
      acc += v;
    }
  }
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sh[q][tx];
    out[col] = t;
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;   // valid in warp 0
}

static int colsum(cudaStream_t s, const float* Y, int64_t ldy, const float* Z, int64_t ldz, int rows, int n, float* out,
                  float* scratch) {
  int parts = scratch ? rows / 512 : 1;
  if (parts > kColParts) parts = kColParts;
  if (parts <= 1) {
    colsum_kernel<<<dim3((n + 31) / 32, 1), 256, 0, s>>>(Y, ldy, Z, ldz, rows, n, out);
    VAE_LAUNCH_OK("colsum_kernel");
    return 0;
  }
  colsum_kernel<<<dim3((n + 31) / 32, parts), 256, 0, s>>>(Y, ldy, Z, ldz, rows, n, scratch);
  VAE_LAUNCH_OK("colsum_kernel(stage 1)");
  colsum_kernel<<<dim3((n + 31) / 32, 1), 256, 0, s>>>(scratch, n, nullptr, 0, parts, n, out);
  VAE_LAUNCH_OK("colsum_kernel(stage 2)");
  return 0;
}

// zs = mu + exp(lv/2) * z1 ; partial sum of mu^2.   (networks.py:73-74, :94)
__global__ void __launch_bounds__(256)
reparam_kernel(const float* __restrict__ mu, const float* __restrict__ z, int ldz,
               const float* __restrict__ lv, int rows, int L, int sampling,
               float* __restrict__ zs, float* __restrict__ part_mu2) {
  __shared__ float sh[8];
  const int64_t total = (int64_t)rows * L;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / L), l = (int)(i % L);
    const float z1 = z[(int64_t)r * ldz + l];
    if (sampling) {                       // networks.py:62-65: mu = 0, logvar = 0
      zs[i] = z1;
    } else {
      const float m = mu[i];
      zs[i] = fmaf(expf(0.5f * lv[l]), z1, m);
      acc = fmaf(m, m, acc);
    }
  }
  const float t = block_sum_256(acc, sh);
  if (threadIdx.x == 0 && part_mu2) part_mu2[blockIdx.x] = t;
}

// x_hat = dec [+ sigmoid(sig_pre)] + z2 * exp(e/2);  r = x_hat - x;
// d x_hat = r / (var * Bscale);  d sig_pre = d x_hat * sg (1 - sg).   (networks.py:75-83,95-96)
__global__ void __launch_bounds__(256)
output_kernel(const float* __restrict__ dec, float* __restrict__ sig, const float* __restrict__ z,
              int ldz, int L, const float* __restrict__ x, const float* __restrict__ eps_ptr,
              float eps_mul, int rows, int D, float inv_bscale, int want_grad,
              float* __restrict__ dxhat, float* __restrict__ xhat_out,
              float* __restrict__ part_r2, float* __restrict__ part_rz) {
  __shared__ float sh[8];
  const float e = eps_ptr ? eps_ptr[0] * eps_mul : eps_mul;
  const float sd = expf(0.5f * e);
  const float inv_var = expf(-e);
  const int64_t total = (int64_t)rows * D;
  float a_r2 = 0.f, a_rz = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), d = (int)(i % D);
    const float z2 = z[(int64_t)r * ldz + L + d];
    float xh = dec[i];
    float sg = 0.f;
    if (sig) {
      sg = 1.f / (1.f + expf(-sig[i]));
      xh += sg;
    }
    xh = fmaf(z2, sd, xh);
    if (xhat_out) xhat_out[i] = xh;
    if (x) {
      const float res = xh - x[i];
      a_r2 = fmaf(res, res, a_r2);
      a_rz = fmaf(res, z2, a_rz);
      if (want_grad) {
        const float g = res * inv_var * inv_bscale;
        dxhat[i] = g;
        if (sig) sig[i] = g * sg * (1.f - sg);
      }
    }
  }
  const float t0 = block_sum_256(a_r2, sh);
  const float t1 = block_sum_256(a_rz, sh);
  if (threadIdx.x == 0 && part_r2) { part_r2[blockIdx.x] = t0; part_rz[blockIdx.x] = t1; }
}

// One block.  loss = (1/Bs) sum_rows [ -0.5 sum_l(1 + lv - e^lv - mu^2) + sum_d(0.5 r^2/var + 0.5(log 2pi + e)) ]
// out[0..2] = loss, dkl.mean, mse.mean ; g_eps = eps0 * dL/de.
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const float* __restrict__ part_mu2, int n_mu2, const float* __restrict__ part_r2,
                     const float* __restrict__ part_rz, int n_out, const float* __restrict__ lv, int L,
                     const float* __restrict__ eps_ptr, float eps_mul, int rows, int D, float inv_bscale,
                     float* __restrict__ out, int n_out_vals, float* __restrict__ g_eps) {
  __shared__ float sh[8];
  float a = 0.f, b = 0.f, c = 0.f, k = 0.f;
  for (int i = threadIdx.x; i < n_mu2; i += blockDim.x) a += part_mu2[i];
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) { b += part_r2[i]; c += part_rz[i]; }
  for (int l = threadIdx.x; l < L; l += blockDim.x) { const float v = lv[l]; k += 1.f + v - expf(v); }
  const float s_mu2 = block_sum_256(a, sh);
  const float s_r2 = block_sum_256(b, sh);
  const float s_rz = block_sum_256(c, sh);
  const float s_lv = block_sum_256(k, sh);
  if (threadIdx.x == 0) {
    const float e = eps_ptr ? eps_ptr[0] * eps_mul : eps_mul;
    const float inv_var = expf(-e);
    const float sd = expf(0.5f * e);
    const float fr = (float)rows;
    const float dkl = (-0.5f * s_lv * fr + 0.5f * s_mu2) * inv_bscale;
    const float mse = (0.5f * inv_var * s_r2 + 0.5f * (1.8378770664093453f + e) * fr * (float)D) * inv_bscale;
    out[0] = dkl + mse;
    if (n_out_vals > 1) { out[1] = dkl; out[2] = mse; }
    if (g_eps) {
      const float de = (-0.5f * inv_var * s_r2 + 0.5f * fr * (float)D + 0.5f * sd * inv_var * s_rz) * inv_bscale;
      g_eps[0] = de * eps_mul;
    }
  }
}

// dmu = dzs + mu / Bscale (in place on dzs);  g_lv[l] = colsum(dzs*z1)[l] * 0.5 * s_l
//   - 0.5 (1 - e^lv_l) * rows / Bscale.       
__global__ void __launch_bounds__(256)
latent_grad_kernel(float* __restrict__ dzs, const float* __restrict__ mu, const float* __restrict__ lv,
                   int rows, int L, float inv_bscale, float* __restrict__ g_lv) {
  const int64_t total = (int64_t)rows * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    dzs[i] = fmaf(mu[i], inv_bscale, dzs[i]);
  if (blockIdx.x == 0) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float v = lv[l];
      g_lv[l] = g_lv[l] * 0.5f * expf(0.5f * v) - 0.5f * (1.f - expf(v)) * (float)rows * inv_bscale;
    }
  }
}

// Legacy flax.optim.Adam.apply_param_gradient [3P].
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
            const float* __restrict__ g, const int32_t* __restrict__ step, int64_t n, AdamHyper hp) {
  const double t = (double)(step[0] + 1);
  const float bc1 = 1.0f - (float)pow((double)hp.b1, t);     // 1 - powf(b, t), correctly rounded
  const float bc2 = 1.0f - (float)pow((double)hp.b2, t);
  const float b1 = hp.b1, b2 = hp.b2, lr = hp.lr, eps = hp.eps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + hp.omb1 * gi;
    const float vi = b2 * v[i] + hp.omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float mh = mi / bc1, vh = vi / bc2;
    p[i] = p[i] - lr * mh / (sqrtf(vh) + eps);
  }
}
__global__ void step_inc_kernel(int32_t* step) { step[0] += 1; }

// ------------------------------------------------------------------- workspace
struct Workspace {
  float* enc_act[VAE_MAX_HIDDEN + 1];
  float* dec_act[VAE_MAX_HIDDEN + 1];
  float* sig_act[VAE_MAX_HIDDEN + 1];
  float* zs;
  float* dxhat;
  float* ga;
  float* gb;
  float* dzs;
  float* part;     // 3 * kMaxPart
  float* colpart;  // kColParts * max_width (two-stage column sums)
  float* grad;     // n_params
  // tensor-core route (use_tc): 3xTF32 operand staging
  bool tc;
  // split(layer input) of every encoder / decoder layer, rows x in_i: written by the forward pass (A operand of
  // Y = X W^T), read again by the backward pass as the MN-major A operand of dW = X^T dY (vae_gemm_tn_3xtf32) -- the
  // activations are split once per step and never transposed
  float *enc_xh[VAE_MAX_HIDDEN + 1], *enc_xl[VAE_MAX_HIDDEN + 1];
  float *dec_xh[VAE_MAX_HIDDEN + 1], *dec_xl[VAE_MAX_HIDDEN + 1];
  float *dyh, *dyl;     // split(dY)  rows x maxw: A of dX = dY W^T and MN-major B of dW (reused layer after layer)
  float *wh, *wl;       // split(W) / transpose_split(W) maxw x maxw
  size_t bytes;
};

// Dense layers are routed to the tcgen05 3xTF32 GEMM (gemm_tc.cu) when they are true dense
// contractions: many rows, wide layers, and every dimension a multiple of 4 floats (TMA
// strides).  Everything else stays on the FP32 FFMA tiles above.
static bool use_tc(const Layout& L, int rows) {
  if (rows < 1024 || rows % 4 != 0 || L.max_width < 256) return false;
  for (int i = 0; i <= L.n_enc; ++i) if (L.enc_sizes[i] % 4) return false;
  for (int i = 0; i <= L.n_dec; ++i) if (L.dec_sizes[i] % 4) return false;
  return true;
}
constexpr int kMaxPart = 1184;   // 148 * 8 blocks

static inline size_t al(size_t n_floats) { return (n_floats * sizeof(float) + 255) & ~(size_t)255; }

static void carve(const VaeDesc* d, const Layout& L, int rows, void* base, Workspace* w) {
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t n_floats) { float* r = (float*)(p + off); off += al(n_floats); return r; };
  for (int i = 0; i < L.n_enc; ++i) w->enc_act[i] = take((size_t)rows * L.enc_sizes[i + 1]);
  for (int i = 0; i < L.n_dec; ++i) w->dec_act[i] = take((size_t)rows * L.dec_sizes[i + 1]);
  for (int i = 0; i < L.n_dec; ++i) w->sig_act[i] = d->sig_decoder ? take((size_t)rows * L.dec_sizes[i + 1]) : nullptr;
  w->zs = take((size_t)rows * d->latent_dim);
  w->dxhat = take((size_t)rows * d->data_dim);
  w->ga = take((size_t)rows * L.max_width);
  w->gb = take((size_t)rows * L.max_width);
  w->dzs = take((size_t)rows * d->latent_dim);
  w->part = take(3 * kMaxPart);
  w->colpart = take((size_t)kColParts * L.max_width);
  w->grad = take((size_t)L.n_params);
  w->tc = use_tc(L, rows);
  if (w->tc) {
    const size_t act = (size_t)rows * L.max_width, ww = (size_t)L.max_width * L.max_width;
    for (int i = 0; i < L.n_enc; ++i) { w->enc_xh[i] = take((size_t)rows * L.enc_sizes[i]); w->enc_xl[i] = take((size_t)rows * L.enc_sizes[i]); }
    for (int i = 0; i < L.n_dec; ++i) { w->dec_xh[i] = take((size_t)rows * L.dec_sizes[i]); w->dec_xl[i] = take((size_t)rows * L.dec_sizes[i]); }
    w->dyh = take(act); w->dyl = take(act);
    w->wh = take(ww); w->wl = take(ww);
  }
  w->bytes = off;
}

static inline int ew_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > kMaxPart) b = kMaxPart;
  if (b < 1) b = 1;
  return (int)b;
}

// Forward MLP: acts[i] = relu?(in_i W_i + b_i)   (networks.py:28-39)
static int mlp_forward(cudaStream_t s, const float* params, const int* sizes, int n_layers,
                       const int64_t* w_off, const int64_t* b_off, const float* input, int rows,
                       float* const* acts, const Workspace* tcw = nullptr, float* const* xh = nullptr,
                       float* const* xl = nullptr) {
  // tcw + xh / xl: the tensor-core route; xh[i] / xl[i] receive the TF32 split of layer i's input (kept for backward)
  const float* h = input;
  for (int i = 0; i < n_layers; ++i) {
    const int in = sizes[i], out = sizes[i + 1];
    const int flags = GEMM_BIAS | (i + 1 < n_layers ? GEMM_RELU : 0);
    if (tcw && tcw->tc) {
      // Y = X W + b as C[rows,out] = X[rows,in] * (W^T)[out,in]^T on the tensor cores
      int rc = vae_split_tf32(h, xh[i], xl[i], (int64_t)rows * in, s);
      if (rc) return rc;
      rc = vae_transpose_split_tf32(params + w_off[i], out, tcw->wh, tcw->wl, in, in, out, s);
      if (rc) return rc;
      rc = vae_gemm_nt_3xtf32(rows, out, in, xh[i], xl[i], in, tcw->wh, tcw->wl, in, acts[i], out,
                              params + b_off[i], nullptr, 0, 1 | (i + 1 < n_layers ? 2 : 0), s);
      if (rc) return rc;
      h = acts[i];
      continue;
    }
    int rc = gemm(s, rows, out, in, h, in, 1, params + w_off[i], out, 1, acts[i], out,
                  params + b_off[i], nullptr, 0, flags);
    if (rc) return rc;
    h = acts[i];
  }
  return 0;
}

// Backward MLP.  dout (rows x sizes[n]) is consumed; returns d input in *d_in (may be
// accumulated into acc_dst when acc flag set).  ga/gb are ping-pong buffers.
static int mlp_backward(cudaStream_t s, const float* params, float* grad, const int* sizes, int n_layers,
                        const int64_t* w_off, const int64_t* b_off, const float* input,
                        float* const* acts, int rows, const float* dout, float* ga, float* gb,
                        float* d_in, bool need_dx, bool acc_dx, const Workspace* tcw = nullptr,
                        void* const* ready = nullptr, float* const* xh = nullptr, float* const* xl = nullptr) {
  // ready (may be null): one cudaEvent_t per layer, recorded on `s` once that layer's kernel + bias gradient is final
  // (ready[0] = the LAST layer, the first one the backward pass finishes)
  const float* d = dout;
  for (int i = n_layers - 1; i >= 0; --i) {
    if (ready && i < n_layers - 1) VAE_CUDA_OK(cudaEventRecord((cudaEvent_t)ready[n_layers - 2 - i], s));   // layer i + 1 is complete
    const int in = sizes[i], out = sizes[i + 1];
    const float* h_in = i == 0 ? input : acts[i - 1];
    if (tcw && tcw->tc && !acc_dx) {
      // dW[in,out] = X[rows,in]^T dY[rows,out]: both operands are read MN-major straight from the batch-major splits
      // (X was split by the forward pass, dY is split once here for dW and dX) ; dX[rows,in] = dY[rows,out] * W[in,out]^T
      int rc = vae_split_tf32(d, tcw->dyh, tcw->dyl, (int64_t)rows * out, s);
      if (rc) return rc;
      rc = vae_gemm_tn_3xtf32(in, out, rows, xh[i], xl[i], in, tcw->dyh, tcw->dyl, out, grad + w_off[i], out, s);
      if (rc) return rc;
      rc = colsum(s, d, out, nullptr, 0, rows, out, grad + b_off[i], tcw->colpart);     // bias gradient
      if (rc) return rc;
      if (i > 0 || need_dx) {
        rc = vae_split_tf32(params + w_off[i], tcw->wh, tcw->wl, (int64_t)in * out, s);
        if (rc) return rc;
        float* dst = i > 0 ? ((d == ga) ? gb : ga) : d_in;
        rc = vae_gemm_nt_3xtf32(rows, in, out, tcw->dyh, tcw->dyl, out, tcw->wh, tcw->wl, out, dst, in, nullptr,
                                i > 0 ? h_in : nullptr, in, i > 0 ? 4 : 0, s);
        if (rc) return rc;
        d = dst;
      }
      continue;
    }
    // dW = h_in^T d
    int rc = gemm(s, in, out, rows, h_in, 1, in, d, out, 1, grad + w_off[i], out, nullptr, nullptr, 0, 0);
    if (rc) return rc;
    rc = colsum(s, d, out, nullptr, 0, rows, out, grad + b_off[i], tcw ? tcw->colpart : nullptr);
    if (rc) return rc;
    if (i > 0) {
      float* dst = (d == ga) ? gb : ga;
      rc = gemm(s, rows, in, out, d, out, 1, params + w_off[i], 1, out, dst, in, nullptr, h_in, in, GEMM_MASK);
      if (rc) return rc;
      d = dst;
    } else if (need_dx) {
      rc = gemm(s, rows, in, out, d, out, 1, params + w_off[i], 1, out, d_in, in, nullptr, nullptr, 0,
                acc_dx ? GEMM_ACC : 0);
      if (rc) return rc;
    }
  }
  if (ready) VAE_CUDA_OK(cudaEventRecord((cudaEvent_t)ready[n_layers - 1], s));      // layer 0 is complete
  return 0;
}

enum Mode { MODE_GRAD, MODE_LOSS, MODE_SAMPLE };

static int run(const VaeDesc* desc, const float* params, const float* d_x, const float* d_z, int rows,
               int scale_rows, Mode mode, const float* d_cur_eps, float* d_out, int n_out_vals,
               float* d_grad, float* d_xhat, void* workspace, size_t workspace_bytes, cudaStream_t s,
               void* const* ready = nullptr) {
  Layout L;
  int rc = make_layout(desc, &L);
  if (rc) return rc;
  VAE_REQUIRE(rows > 0 && scale_rows > 0, "rows must be positive");
  VAE_REQUIRE(params && d_z, "params / z pointer is NULL");
  VAE_REQUIRE(mode == MODE_SAMPLE || d_x, "x pointer is NULL");
  Workspace w;
  carve(desc, L, rows, workspace, &w);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return VAE_ERR_WORKSPACE;
  }
  const int Ld = desc->latent_dim, D = desc->data_dim, ldz = Ld + D;
  const float inv_bs = 1.0f / (float)scale_rows;
  const float eps0 = (float)desc->epsilon0;
  const float* lv = params + L.off_eps_p;
  const float* eps_ptr;
  float eps_mul;
  if (mode == MODE_SAMPLE) { eps_ptr = d_cur_eps; eps_mul = 1.f; }
  else if (desc->tunable_decoder_var) { eps_ptr = params + L.off_eps; eps_mul = eps0; }
  else { eps_ptr = nullptr; eps_mul = eps0; }
  float* part_mu2 = w.part;
  float* part_r2 = w.part + kMaxPart;
  float* part_rz = w.part + 2 * kMaxPart;

  float* mu = w.enc_act[L.n_enc - 1];
  if (mode != MODE_SAMPLE) {
    rc = mlp_forward(s, params, L.enc_sizes, L.n_enc, L.enc_w, L.enc_b, d_x, rows, w.enc_act, &w, w.enc_xh, w.enc_xl);
    if (rc) return rc;
  }
  const int g_lat = ew_grid((int64_t)rows * Ld);
  reparam_kernel<<<g_lat, 256, 0, s>>>(mu, d_z, ldz, lv, rows, Ld, mode == MODE_SAMPLE, w.zs, part_mu2);
  VAE_LAUNCH_OK("reparam_kernel");
  rc = mlp_forward(s, params, L.dec_sizes, L.n_dec, L.dec_w, L.dec_b, w.zs, rows, w.dec_act, &w, w.dec_xh, w.dec_xl);
  if (rc) return rc;
  float* sig_out = nullptr;
  if (desc->sig_decoder) {
    rc = mlp_forward(s, params, L.dec_sizes, L.n_dec, L.sig_w, L.sig_b, w.zs, rows, w.sig_act);
    if (rc) return rc;
    sig_out = w.sig_act[L.n_dec - 1];
  }
  const int g_out = ew_grid((int64_t)rows * D);
  output_kernel<<<g_out, 256, 0, s>>>(w.dec_act[L.n_dec - 1], sig_out, d_z, ldz, Ld,
                                      mode == MODE_SAMPLE ? nullptr : d_x, eps_ptr, eps_mul, rows, D,
                                      inv_bs, mode == MODE_GRAD, w.dxhat, d_xhat, part_r2, part_rz);
  VAE_LAUNCH_OK("output_kernel");
  if (mode == MODE_SAMPLE) return 0;
  float* grad = d_grad;
  loss_finalize_kernel<<<1, 256, 0, s>>>(part_mu2, g_lat, part_r2, part_rz, g_out, lv, Ld, eps_ptr, eps_mul,
                                         rows, D, inv_bs, d_out, n_out_vals,
                                         (mode == MODE_GRAD && desc->tunable_decoder_var) ? grad + L.off_eps : nullptr);
  VAE_LAUNCH_OK("loss_finalize_kernel");
  if (mode == MODE_LOSS) return 0;

  // backward: decoder(s) -> dzs
  // gradient segments in completion order (vae_grad_segments): decoder layers last -> first, the latent block
  // (epsilon_p, epsilon), encoder layers last -> first.  Two-decoder (sigmoid) models report everything at the end.
  const bool seg_events = ready != nullptr && !desc->sig_decoder;
  rc = mlp_backward(s, params, grad, L.dec_sizes, L.n_dec, L.dec_w, L.dec_b, w.zs, w.dec_act, rows,
                    w.dxhat, w.ga, w.gb, w.dzs, true, false, &w, seg_events ? ready : nullptr, w.dec_xh, w.dec_xl);
  if (rc) return rc;
  if (desc->sig_decoder) {
    rc = mlp_backward(s, params, grad, L.dec_sizes, L.n_dec, L.sig_w, L.sig_b, w.zs, w.sig_act, rows,
                      sig_out, w.ga, w.gb, w.dzs, true, true);
    if (rc) return rc;
  }
  rc = colsum(s, w.dzs, Ld, d_z, ldz, rows, Ld, grad + L.off_eps_p, w.colpart);
  if (rc) return rc;
  latent_grad_kernel<<<g_lat, 256, 0, s>>>(w.dzs, mu, lv, rows, Ld, inv_bs, grad + L.off_eps_p);
  VAE_LAUNCH_OK("latent_grad_kernel");
  if (seg_events) VAE_CUDA_OK(cudaEventRecord((cudaEvent_t)ready[L.n_dec], s));
  rc = mlp_backward(s, params, grad, L.enc_sizes, L.n_enc, L.enc_w, L.enc_b, d_x, w.enc_act, rows,
                    w.dzs, w.ga, w.gb, nullptr, false, false, &w, seg_events ? ready + L.n_dec + 1 : nullptr, w.enc_xh, w.enc_xl);
  if (rc) return rc;
  if (ready && !seg_events) VAE_CUDA_OK(cudaEventRecord((cudaEvent_t)ready[0], s));
  return 0;
}

// Flat-gradient segments in the order the backward pass completes them.
static int grad_segments(const VaeDesc* desc, VaeGradSegment* out, int max_out) {
  Layout L;
  int rc = make_layout(desc, &L);
  if (rc) return rc;
  if (desc->sig_decoder) {
    if (out && max_out >= 1) { out[0].begin = 0; out[0].end = L.n_params; }
    return 1;
  }
  const int n = L.n_dec + 1 + L.n_enc;
  if (!out) return n;
  VAE_REQUIRE(max_out >= n, "vae_grad_segments: room for %d segments is needed", n);
  int k = 0;
  for (int i = L.n_dec - 1; i >= 0; --i, ++k) { out[k].begin = L.dec_w[i]; out[k].end = L.dec_b[i] + L.dec_sizes[i + 1]; }
  out[k].begin = L.off_eps_p; out[k].end = L.n_params; ++k;
  for (int i = L.n_enc - 1; i >= 0; --i, ++k) { out[k].begin = L.enc_w[i]; out[k].end = L.enc_b[i] + L.enc_sizes[i + 1]; }
  return n;
}

}  // namespace vae

extern "C" {

size_t vae_workspace_bytes(const VaeDesc* desc, int32_t rows) {
  vae::Layout L;
  if (vae::make_layout(desc, &L) || rows <= 0) return 0;
  vae::Workspace w;
  vae::carve(desc, L, rows, nullptr, &w);
  return w.bytes;
}

int vae_loss_and_grad(const VaeDesc* desc, const float* params, const float* d_x, const float* d_z,
                      int32_t rows, int32_t grad_scale_rows, float* d_loss, float* d_grad,
                      void* workspace, size_t workspace_bytes, void* stream) {
  VAE_REQUIRE(d_loss && d_grad, "vae_loss_and_grad: output pointer is NULL");
  return vae::run(desc, params, d_x, d_z, rows, grad_scale_rows, vae::MODE_GRAD, nullptr, d_loss, 1,
                  d_grad, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

int32_t vae_grad_segments(const VaeDesc* desc, VaeGradSegment* out, int32_t max_out) {
  return vae::grad_segments(desc, out, max_out);
}

int vae_loss_and_grad_events(const VaeDesc* desc, const float* params, const float* d_x, const float* d_z,
                             int32_t rows, int32_t grad_scale_rows, float* d_loss, float* d_grad,
                             void* workspace, size_t workspace_bytes, void* const* ready_events, void* stream) {
  VAE_REQUIRE(d_loss && d_grad && ready_events, "vae_loss_and_grad_events: NULL pointer");
  return vae::run(desc, params, d_x, d_z, rows, grad_scale_rows, vae::MODE_GRAD, nullptr, d_loss, 1,
                  d_grad, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, ready_events);
}

int vae_adam_update_range(const VaeDesc* desc, float* params, float* adam_m, float* adam_v, const int32_t* d_step,
                          const float* d_grad, int64_t begin, int64_t end, void* stream) {
  vae::Layout L;
  int rc = vae::make_layout(desc, &L);
  if (rc) return rc;
  VAE_REQUIRE(params && adam_m && adam_v && d_step && d_grad, "vae_adam_update_range: NULL pointer");
  VAE_REQUIRE(0 <= begin && begin <= end && end <= L.n_params, "vae_adam_update_range: bad range");
  if (begin == end) return 0;
  int64_t blocks = (end - begin + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  vae::adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params + begin, adam_m + begin, adam_v + begin, d_grad + begin,
                                                                  d_step, end - begin, vae::make_adam_hyper(desc));
  VAE_LAUNCH_OK("adam_kernel");
  return 0;
}

int vae_adam_step_increment(int32_t* d_step, void* stream) {
  VAE_REQUIRE(d_step, "vae_adam_step_increment: NULL pointer");
  vae::step_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_step);
  VAE_LAUNCH_OK("step_inc_kernel");
  return 0;
}

int vae_adam_update(const VaeDesc* desc, float* params, float* adam_m, float* adam_v, int32_t* d_step,
                    const float* d_grad, void* stream) {
  vae::Layout L;
  int rc = vae::make_layout(desc, &L);
  if (rc) return rc;
  VAE_REQUIRE(params && adam_m && adam_v && d_step && d_grad, "vae_adam_update: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (L.n_params + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  vae::adam_kernel<<<(int)blocks, 256, 0, s>>>(params, adam_m, adam_v, d_grad, d_step, L.n_params,
                                               vae::make_adam_hyper(desc));
  VAE_LAUNCH_OK("adam_kernel");
  vae::step_inc_kernel<<<1, 1, 0, s>>>(d_step);
  VAE_LAUNCH_OK("step_inc_kernel");
  return 0;
}

int vae_train_step(const VaeDesc* desc, float* params, float* adam_m, float* adam_v, int32_t* d_step,
                   const float* d_x, const float* d_z, float* d_loss, void* workspace,
                   size_t workspace_bytes, void* stream) {
  vae::Layout L;
  int rc = vae::make_layout(desc, &L);
  if (rc) return rc;
  VAE_REQUIRE(d_loss, "vae_train_step: d_loss is NULL");
  vae::Workspace w;
  vae::carve(desc, L, desc->batch, workspace, &w);
  if (!workspace || workspace_bytes < w.bytes) {
    vae::set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return VAE_ERR_WORKSPACE;
  }
  rc = vae::run(desc, params, d_x, d_z, desc->batch, desc->batch, vae::MODE_GRAD, nullptr, d_loss, 1,
                w.grad, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
  if (rc) return rc;
  return vae_adam_update(desc, params, adam_m, adam_v, d_step, w.grad, stream);
}

int vae_loss(const VaeDesc* desc, const float* params, const float* d_x, const float* d_z, int32_t rows,
             float* d_out, void* workspace, size_t workspace_bytes, void* stream) {
  VAE_REQUIRE(d_out, "vae_loss: d_out is NULL");
  return vae::run(desc, params, d_x, d_z, rows, rows, vae::MODE_LOSS, nullptr, d_out, 3, nullptr, nullptr,
                  workspace, workspace_bytes, (cudaStream_t)stream);
}

int vae_sample(const VaeDesc* desc, const float* params, const float* d_z, int32_t rows,
               const float* d_current_epsilon, float* d_xhat, void* workspace, size_t workspace_bytes,
               void* stream) {
  VAE_REQUIRE(d_current_epsilon && d_xhat, "vae_sample: NULL pointer");
  return vae::run(desc, params, nullptr, d_z, rows, rows, vae::MODE_SAMPLE, d_current_epsilon, nullptr, 0,
                  nullptr, d_xhat, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
