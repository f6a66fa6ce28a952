// fused_mlp.cu -- persistent fused training loop for the hidden-layer VAEs of the sphere
// sweep (sphere_vae_padding_expts.sh: enc = dec = 200|200|200), one thread-block CLUSTER per
// model.
//
// The same hot-loop body as fused_small.cu (model.py:221-222: key splits, eps draw, dataset
// batch, VAE.apply networks.py:61-84, sampled ELBO :94-98, hand-derived backward, legacy flax
// Adam), but a 166-176k-parameter model (2 MB with Adam state) does not fit one SM.  A model
// runs on a cluster of CS = 16 / 8 / 4 / 2 / 1 CTAs, chosen per launch from the pack size
// (mlp_pick_cluster: few models -> many CTAs each for latency, many models -> few CTAs each for
// throughput; results are bit-identical for every CS).  The CTAs split every Dense layer by
// OUTPUT COLUMNS (at least four slices per layer; CS < 4 walks its slices in passes):
//   forward   Y[:, slice]  = act(X W[:, slice] + b[slice])              X: all rows, all K
//   backward  dW[:, slice] = X^T dY[:, slice]  (+ Adam on the slice)     the CTA owns its columns
//             dX[:, kslice] = (dY W[kslice, :]^T) * relu'(X[:, kslice])   dY: all rows, all N
// Activations and gradients are exchanged through an L2-resident workspace in FEATURE-MAJOR
// layout ([feature][batch], batch contiguous => 128-bit loads along the batch everywhere) and
// one hardware cluster barrier per layer phase; weights / Adam moments stay in L2 (they are
// touched once per step).  Per step the only HBM-bound traffic is the 4-byte loss store.
//
// Exact work skipping: the generated batch has exactly-zero padding columns, so encoder layer
// 0 uses only its first DR input rows and the other rows' gradients (exactly 0) are never
// formed (Adam leaves those weights bit-for-bit unchanged).
#include <stdlib.h>

#include "common.cuh"

namespace vae {

constexpr int kMT = 512;                       // threads per CTA
constexpr int kML = VAE_MAX_HIDDEN + 1;        // max Dense layers per net
constexpr int kMlpMaxCluster = 8;          // portable cluster size: the throughput configuration
constexpr int kMlpBigCluster = 16;         // non-portable: one model over 16 SMs (latency configuration)
constexpr int kMlpMinSlices = 4;           // a layer is cut into at least this many column slices (staging-buffer width)
constexpr int kMlpSmallCluster = 4;        // packs larger than the resident 8-CTA clusters: more models per GPU

struct MlpJobDev {
  float* params; float* adam_m; float* adam_v; int32_t* step;
  uint32_t* model_key; uint32_t* dataset_key; const float* A; float* loss_out; float* ws;
  int32_t B, BP, D, L, DR, kind, G, dd, tunable, n_enc, n_dec, maxw;
  int32_t enc_sizes[kML + 1], dec_sizes[kML + 1];
  int32_t enc_w[kML], enc_b[kML], dec_w[kML], dec_b[kML];
  int32_t off_lv, off_eps;
  int32_t ws_zT, ws_gn, ws_enc[kML], ws_dec[kML], ws_g[2], ws_dzs;
  float eps0, lr, b1, b2, omb1, omb2, adam_eps;
};

__host__ __device__ constexpr int m_up4(int n) { return (n + 3) & ~3; }
__host__ __device__ inline int m_batch_stride(int B) { int g = (B + 3) / 4; g |= 1; return 4 * g; }

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// All threads of all CTAs of the cluster.  The arrive's release and the wait's acquire are cluster-scope: every
// thread's global writes before the barrier are visible to every thread of the cluster after it (the exchange buffers
// are read with .cg loads / cp.async.cg, i.e. from L2).  A device-scope __threadfence() in front of it is redundant
// and cost 3 us per step (measured 147.5 -> 144.4 us, results bit-identical).
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// idx / d for 0 <= idx < 2^22 and quotients < 2^12 through one float multiply: (idx + 0.5) / d is at least 0.5 / d away from
// an integer, far more than the 2^-23 relative error of the two roundings, so truncation gives the exact quotient.
__device__ __forceinline__ int fast_div(int idx, float inv_d) { return __float2int_rz(((float)idx + 0.5f) * inv_d); }

struct Slice { int c0, w; };
// Slices are dealt out from the LAST rank down: the short remainder slice (5 of 200 columns at 16 CTAs) lands on CTA 0,
// which also computes the loss and the latent-variance update for the whole cluster.
__device__ __forceinline__ Slice slice_of(int n, int rank, int cs) {
  const int chunk = (n + cs - 1) / cs;
  int c0 = (cs - 1 - rank) * chunk; if (c0 > n) c0 = n;
  int c1 = c0 + chunk; if (c1 > n) c1 = n;
  return Slice{c0, c1 - c0};
}

// Asynchronous global -> shared copies (LDGSTS): every copy of a staging loop is in flight at
// once, so a whole operand costs about one L2 round trip instead of one per loop iteration.
// Data written by other CTAs of the cluster is only read after a cluster barrier (release /
// acquire at cluster scope; the barrier also invalidates L1).
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Bulk (TMA) copies for whole activation arrays: ONE instruction of one thread moves K x BP floats from the L2-resident
// workspace into shared memory (the LDGSTS form costs K * BP / 128 warp instructions at 8 cycles each: ~1.3 k cycles of
// issue per 80 KB operand); completion is counted in bytes on an mbarrier every thread waits on.
__device__ __forceinline__ uint32_t m_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m_smem(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void m_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m_smem(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void m_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = m_smem(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// dst (smem, [nrows][BP]) <- src (global, same layout); call from ONE thread, after m_expect_tx of the total
__device__ __forceinline__ void bulk_rows(float* dst, const float* src, int nrows, int BP, uint64_t* bar) {
  const uint32_t bytes = (uint32_t)(nrows * BP) * 4u;      // one copy per array (80 KB for a 200-wide layer); a multiple of 16
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(m_smem(dst)), "l"(src), "r"(bytes), "r"(m_smem(bar)) : "memory");
}
// This thread's global stores become visible to the async proxy (the peers' bulk copies after the next cluster barrier).
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// dst (smem, [nrows][BP]) <- src (global, same layout), asynchronous: cp_async_wait_all() +
// __syncthreads() before use
__device__ __forceinline__ void load_rows(float* dst, const float* src, int nrows, int BP) {
  const int n4 = (nrows * BP) >> 2;
  for (int i = threadIdx.x; i < n4; i += kMT) cp_async16(dst + 4 * i, src + 4 * i);
}

// outT[(o0+j)][b] = epi( add[o0+j] + sum_k AT[k][b] * Ws[k][j] ),  j < w, all b.
//   Ws (smem, [K][W4]) must already be staged; AT smem [K][BP].
//   relu: max(.,0).  maskT != nullptr: multiply by (maskT[(o0+j)][b] > 0) (relu backward).
//   Rows b >= B are written as 0.
struct GemmArgs {
  const float* AT; int K; const float* Ws; int W4, w; const float* add_g; int o0; bool relu;
  const float* maskT; float* outT_g; int BP, B;
};
// One task of tile_gemm: tt = 2 * tile + (half of the k range).  All 32 lanes of a warp call it together
// (the two lanes of a pair exchange their partial sums with a shuffle).
__device__ __forceinline__ void gemm_task(const GemmArgs& g, int tt, int ntile, int NG) {
  const int K = g.K, W4 = g.W4, BP = g.BP;
  const int kh = (K + 1) >> 1;
  const bool live = tt < 2 * ntile;
  const int tile = live ? tt >> 1 : ntile - 1, half = tt & 1;
  const int cg = tile / NG, rg = tile - cg * NG;
  // Blackwell packed FP32 FMA (FFMA2): accumulators are row pairs, (a.x, a.y) / (a.z, a.w) are natural
  // register pairs of the 128-bit load; only the weight has to be duplicated into a pair.  Every lane
  // of a pair performs the same RN fma as the scalar form: results are unchanged.
  float2 acc2[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { acc2[j][0] = make_float2(0.f, 0.f); acc2[j][1] = make_float2(0.f, 0.f); }
  const float* ap = g.AT + 4 * rg;
  const float* wp = g.Ws + 4 * cg;
  const int kbeg = half ? kh : 0, kend = half ? K : kh;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};          // loaded up front: the L2 round trip overlaps the k loop
  if (g.add_g && live && !half) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (4 * cg + j < g.w) bias[j] = g.add_g[g.o0 + 4 * cg + j];
  }
#pragma unroll 4
  for (int k = kbeg; k < kend; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(ap + k * BP);
    const float4 wv = *reinterpret_cast<const float4*>(wp + k * W4);
    const float2 a01 = make_float2(a.x, a.y), a23 = make_float2(a.z, a.w);
    const float wj[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 ww = make_float2(wj[j], wj[j]);
      acc2[j][0] = __ffma2_rn(ww, a01, acc2[j][0]);
      acc2[j][1] = __ffma2_rn(ww, a23, acc2[j][1]);
    }
  }
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { acc[j][0] = acc2[j][0].x; acc[j][1] = acc2[j][0].y; acc[j][2] = acc2[j][1].x; acc[j][3] = acc2[j][1].y; }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[j][r] += __shfl_xor_sync(0xffffffffu, acc[j][r], 1);
  if (!live || half) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = 4 * cg + j;
    if (col >= g.w) continue;
    float o[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float v = acc[j][r] + bias[j];
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.maskT) v = g.maskT[(g.o0 + col) * BP + 4 * rg + r] > 0.f ? v : 0.f;
      if (4 * rg + r >= g.B) v = 0.f;
      o[r] = v;
    }
    *reinterpret_cast<float4*>(g.outT_g + (size_t)(g.o0 + col) * BP + 4 * rg) = make_float4(o[0], o[1], o[2], o[3]);
  }
}
__device__ __forceinline__ void tile_gemm(const float* __restrict__ AT, int K, const float* __restrict__ Ws, int W4,
                                          int w, const float* __restrict__ add_g, int o0, bool relu,
                                          const float* __restrict__ maskT, float* __restrict__ outT_g, int BP, int B) {
  // 4x4 output tiles; the reduction over k is split between the two lanes of a pair (halves
  // of K), combined with a shuffle: twice the warps in flight for the same instruction count.
  const GemmArgs g{AT, K, Ws, W4, w, add_g, o0, relu, maskT, outT_g, BP, B};
  const int ncg = W4 >> 2, NG = BP >> 2, ntile = ncg * NG;
  for (int tt0 = 0; tt0 < 2 * ntile; tt0 += kMT) {
    if (tt0 + ((int)threadIdx.x & ~31) >= 2 * ntile) break;              // warp-uniform
    gemm_task(g, tt0 + (int)threadIdx.x, ntile, NG);
  }
}

// Ws[k][j] = Wg[k*N + c0 + j]  (forward / dW column slice), zero padded to W4
// (t, nthr): the copying threads and their count (default: the whole CTA)
__device__ __forceinline__ void stage_w_cols(float* Ws, const float* Wg, int K, int N, int c0, int w, int W4,
                                             int t = threadIdx.x, int nthr = kMT) {
  const float inv = 1.0f / (float)W4;
  for (int idx = t; idx < K * W4; idx += nthr) {
    const int k = fast_div(idx, inv), j = idx - k * W4;
    if (j < w) cp_async4(Ws + idx, Wg + (size_t)k * N + c0 + j); else Ws[idx] = 0.f;
  }
}
// Ws[n][j] = Wg[(k0+j)*N + n]   (dX: rows k0..k0+w of W, transposed), zero padded to W4
__device__ __forceinline__ void stage_w_rows_t(float* Ws, const float* Wg, int N, int k0, int w, int W4,
                                               int t = threadIdx.x, int nthr = kMT) {
  const float inv = 1.0f / (float)N;
  for (int idx = t; idx < W4 * N; idx += nthr) {
    const int j = fast_div(idx, inv), n = idx - j * N;
    if (j < w) cp_async4(Ws + n * W4 + j, Wg + (size_t)(k0 + j) * N + n); else Ws[n * W4 + j] = 0.f;
  }
}

// dWs[k][j] = sum_b HT[k][b] * dYT[c0+j][b]  (k < K, j < w);  dbs[j] = sum_b dYT[c0+j][b]
// 4x4 tiles over (k, j); the 4 k-rows of a tile are interleaved (k, k+KT, k+2KT, k+3KT) so
// the 128-bit loads of a warp's consecutive tiles hit consecutive bank groups.
__device__ __forceinline__ void dw_task(int tile, const float* __restrict__ HT, int K, const float* __restrict__ dYT,
                                        int c0, int w, int W4, float* __restrict__ dWs, int BP) {
  const int KT = (K + 3) >> 2, NG = BP >> 2;
  const int cg = tile / KT, kg = tile - cg * KT;
  const float* hp[4]; const float* dp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) hp[i] = HT + (size_t)min(kg + i * KT, K - 1) * BP;
#pragma unroll
  for (int j = 0; j < 4; ++j) dp[j] = dYT + (size_t)(c0 + min(4 * cg + j, w - 1)) * BP;
  // FFMA2: each accumulator is a pair over the (x, y) / (z, w) halves of the 128-bit batch loads
  // (even and odd batch columns are summed separately, then added)
  float2 acc2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
  for (int g = 0; g < NG; ++g) {
    float4 hv[4], dv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) hv[i] = *reinterpret_cast<const float4*>(hp[i] + 4 * g);
#pragma unroll
    for (int j = 0; j < 4; ++j) dv[j] = *reinterpret_cast<const float4*>(dp[j] + 4 * g);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = acc2[i][j];
        a = __ffma2_rn(make_float2(hv[i].x, hv[i].y), make_float2(dv[j].x, dv[j].y), a);
        a = __ffma2_rn(make_float2(hv[i].z, hv[i].w), make_float2(dv[j].z, dv[j].w), a);
        acc2[i][j] = a;
      }
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = acc2[i][j].x + acc2[i][j].y;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = kg + i * KT;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * cg + j < w) dWs[k * W4 + 4 * cg + j] = acc[i][j];
  }
}
// bias gradient of output column c0 + j
__device__ __forceinline__ float db_col(const float* __restrict__ dYT, int c0, int j, int BP) {
  const float* dp = dYT + (size_t)(c0 + j) * BP;
  float a = 0.f;
  for (int b = 0; b < BP; ++b) a += dp[b];
  return a;
}
__device__ __forceinline__ void tile_dw(const float* __restrict__ HT, int K, const float* __restrict__ dYT, int c0, int w,
                                        int W4, float* __restrict__ dWs, float* __restrict__ dbs, int BP) {
  const int KT = (K + 3) >> 2, ncg = W4 >> 2;
  for (int tile = threadIdx.x; tile < KT * ncg; tile += kMT) dw_task(tile, HT, K, dYT, c0, w, W4, dWs, BP);
  for (int j = threadIdx.x; j < w; j += kMT) dbs[j] = db_col(dYT, c0, j, BP);
}

struct MlpHyper { float lr, b1, b2, omb1, omb2, eps, ibc1, ibc2; };
__device__ __forceinline__ float m_approx_sqrt(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void adam_one(float* P, float* M, float* V, size_t i, float g, const MlpHyper& h) {
  const float m = h.b1 * M[i] + h.omb1 * g;
  const float v = h.b2 * V[i] + h.omb2 * g * g;
  M[i] = m; V[i] = v;
  P[i] = P[i] - __fdividef(h.lr * (m * h.ibc1), m_approx_sqrt(v * h.ibc2) + h.eps);
}
// Adam on this CTA's column slice of a Dense layer (kernel rows < K, and the bias).  The
// moments live in L2: each thread first issues the loads of kAB parameters, then updates
// (memory-level parallelism instead of one L2 round trip per parameter).
constexpr int kAB = 8;
__device__ __forceinline__ void adam_slice(const MlpJobDev& J, int w_off, int b_off, int K, int N, int c0, int w, int W4,
                                           const float* dWs, const float* dbs, const MlpHyper& h) {
  float* __restrict__ P = J.params; float* __restrict__ M = J.adam_m; float* __restrict__ V = J.adam_v;
  const int total = (K + 1) * w;                 // row K of the slice is the bias: one batched pass, one L2 round trip
  const float inv_w = 1.0f / (float)(w > 0 ? w : 1);
  for (int base = 0; base < total; base += kAB * kMT) {
    size_t gi[kAB]; float p[kAB], m[kAB], v[kAB], g[kAB]; bool ok[kAB];
#pragma unroll
    for (int u = 0; u < kAB; ++u) {
      const int idx = base + u * kMT + (int)threadIdx.x;
      ok[u] = idx < total;
      const int ii = ok[u] ? idx : 0;
      const int k = fast_div(ii, inv_w), j = ii - k * w;
      gi[u] = k < K ? (size_t)w_off + (size_t)k * N + c0 + j : (size_t)b_off + c0 + j;
      g[u] = k < K ? dWs[k * W4 + j] : dbs[j];
      p[u] = __ldcg(P + gi[u]); m[u] = __ldcg(M + gi[u]); v[u] = __ldcg(V + gi[u]);
    }
#pragma unroll
    for (int u = 0; u < kAB; ++u) {
      const float mn = h.b1 * m[u] + h.omb1 * g[u];
      const float vn = h.b2 * v[u] + h.omb2 * g[u] * g[u];
      if (ok[u]) {
        M[gi[u]] = mn; V[gi[u]] = vn;
        P[gi[u]] = p[u] - __fdividef(h.lr * (mn * h.ibc1), m_approx_sqrt(vn * h.ibc2) + h.eps);
      }
    }
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {   // result valid in every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kMT / 32; ++i) t += red[i];
  return t;
}

// Dense layers in execution order (overlapped schedule of the kernel, see "dual" below).
// forward q: encoder 0 .. n_enc-1, then decoder 0 .. n_dec-1; backward r: decoder n_dec-1 .. 0, then encoder n_enc-1 .. 0.
struct LayerRef {
  int K, N, w_off, b_off, i;
  const float* in_g;      // layer input in the workspace ([K][BP]); nullptr: the batch (encoder 0) / zs (decoder 0), in smem
  float* out_g;           // forward output in the workspace ([N][BP])
  bool relu, is_dec;
};
__device__ __forceinline__ LayerRef layer_ref(const MlpJobDev& J, float* ws, bool is_dec, int i) {
  LayerRef r;
  r.is_dec = is_dec; r.i = i;
  if (!is_dec) {
    r.K = i == 0 ? J.DR : J.enc_sizes[i]; r.N = J.enc_sizes[i + 1]; r.w_off = J.enc_w[i]; r.b_off = J.enc_b[i];
    r.in_g = i > 0 ? ws + J.ws_enc[i - 1] : nullptr; r.out_g = ws + J.ws_enc[i]; r.relu = i + 1 < J.n_enc;
  } else {
    r.K = J.dec_sizes[i]; r.N = J.dec_sizes[i + 1]; r.w_off = J.dec_w[i]; r.b_off = J.dec_b[i];
    r.in_g = i > 0 ? ws + J.ws_dec[i - 1] : nullptr; r.out_g = ws + J.ws_dec[i]; r.relu = i + 1 < J.n_dec;
  }
  return r;
}
__device__ __forceinline__ LayerRef fwd_ref(const MlpJobDev& J, float* ws, int q) {
  return q < J.n_enc ? layer_ref(J, ws, false, q) : layer_ref(J, ws, true, q - J.n_enc);
}
__device__ __forceinline__ LayerRef bwd_ref(const MlpJobDev& J, float* ws, int r) {
  return r < J.n_dec ? layer_ref(J, ws, true, J.n_dec - 1 - r) : layer_ref(J, ws, false, J.n_enc - 1 - (r - J.n_dec));
}
__device__ __forceinline__ uint32_t dynamic_smem_bytes() { uint32_t r; asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(r)); return r; }

// -DVAE_MLP_TIMING: thread 0 of the first and of the last CTA of cluster 0 accumulate clock64() intervals per category
// (scratch/mlp_timing.py reads them through vae_mlp_timing): where a step's time goes, as one CTA sees it.
#ifdef VAE_MLP_TIMING
enum { kTR = 0, kTFwdLoad, kTFwdGemm, kTFwdBar, kTOut, kTBwdFirstLoad, kTBwdTiles, kTBwdBar, kTBwdIssue, kTBwdAdam, kTBwdWait, kTLatent, kTTop, kTN };
__device__ unsigned long long g_mlp_t[2][16];
#define MLP_T0() long long t_mark = clock64()
#define MLP_T(cat) do { if (tid == 0 && blockIdx.x < (unsigned)CS && (rank == 0 || rank == CS - 1)) { \
    const long long t_now = clock64(); g_mlp_t[rank == 0 ? 0 : 1][cat] += (unsigned long long)(t_now - t_mark); t_mark = t_now; } } while (0)
extern "C" int vae_mlp_timing(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  int rc = (int)cudaMemcpyFromSymbol(out32, g_mlp_t, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; rc |= (int)cudaMemcpyToSymbol(g_mlp_t, z, sizeof(z)); }
  return rc;
}
#else
#define MLP_T0() do {} while (0)
#define MLP_T(cat) do {} while (0)
#endif

__global__ void __launch_bounds__(kMT, 1)
fused_mlp_kernel(const MlpJobDev* __restrict__ jobs, int64_t n_steps) {
  extern __shared__ __align__(16) float sm[];
  __shared__ MlpJobDev J;
  __shared__ float red[kMT / 32];
  __shared__ uint32_t keys[8];
  __shared__ float sl_s[64], scal_s[4];
  __shared__ __align__(8) uint64_t mbar;          // completion of the bulk operand copies (overlapped schedule)
  const int tid = threadIdx.x;
  const int rank = (int)cluster_ctarank(), CS = (int)cluster_nctarank();
  {
    const int* src = reinterpret_cast<const int*>(jobs + blockIdx.x / CS);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < (int)(sizeof(MlpJobDev) / 4); i += kMT) dst[i] = src[i];
  }
  __syncthreads();
  const int B = J.B, BP = J.BP, D = J.D, L = J.L, DR = J.DR, G = J.G, LD = L + D;
  // Column slices per layer: one per CTA, but never fewer than kMlpMinSlices -- clusters of 1 or 2 CTAs
  // (throughput packs) walk their slices in NP passes through the same staging buffer.
  const int SL = CS < kMlpMinSlices ? kMlpMinSlices : CS, NP = SL / CS;
  const int W4max = m_up4((J.maxw + SL - 1) / SL);
  // shared memory carve-up (floats)
  float* bufA = sm;                               // [maxw][BP]  layer input (H^T)
  float* bufB = bufA + (size_t)J.maxw * BP;       // [maxw][BP]  layer output gradient (dY^T)
  float* Ws = bufB + (size_t)J.maxw * BP;         // [maxw][W4max] staged weights / dW slice
  float* dbs = Ws + (size_t)J.maxw * W4max;       // [W4max]
  float* xT = dbs + W4max;                        // [DR][BP]
  float* z1T = xT + (size_t)DR * BP;              // [L][BP]
  float* muT = z1T + (size_t)L * BP;              // [L][BP]
  float* zsT = muT + (size_t)L * BP;              // [L][BP]
  float* Ws2 = zsT + (size_t)L * BP;              // [maxw][W4max]  second staging buffer (overlapped schedule only)
  float* ws = J.ws;
  // Overlapped schedule ("dual"): with one column slice per CTA (NP == 1) and room for a second staging buffer
  // (mlp_smem_bytes(.., dual); the host grants it unless VAE_MLP_DUAL=0), every backward layer runs its dX tiles and its
  // dW tiles SIDE BY SIDE on different warps (each alone occupies 7 of the 16 warps of a 16-CTA cluster), and the
  // operands of the next layer -- its input activations, its transposed weight rows, the next dY -- are fetched while
  // Adam updates this layer's slice; forward layers stage the next layer's weight slice while they multiply.  The
  // arithmetic of every output element is that of the sequential schedule: results are bit-identical.
  const int SLd = SL < 8 ? SL : 8;
  const int W4d = m_up4((J.maxw + SLd - 1) / SLd);
  const bool dual = CS >= kMlpBigCluster &&       // measured: no gain at 8 CTAs (two rounds of tiles per layer)
      dynamic_smem_bytes() >= 4u * (uint32_t)(2 * J.maxw * BP + 2 * J.maxw * W4d + W4d + DR * BP + 3 * L * BP);
  float* Wsb[2] = {Ws, Ws2};
  uint32_t mph = 0;                               // phase parity of mbar (uniform: every thread waits on every copy)
  if (tid == 0) m_mbar_init(&mbar, 1);            // ordered before its first use by the barriers below
  for (int i = tid; i < DR * BP; i += kMT) xT[i] = 0.f;

  // key chains (every CTA tracks both redundantly; lanes 0,1 model side, lanes 2,3 dataset side)
  uint32_t ck0 = 0, ck1 = 0;
  if (tid < 4) { const uint32_t* src = tid < 2 ? J.model_key : J.dataset_key; ck0 = src[0]; ck1 = src[1]; }
  const int32_t step0 = J.step[0];
  double b1t = pow((double)J.b1, (double)step0), b2t = pow((double)J.b2, (double)step0);
  MlpHyper hp; hp.lr = J.lr; hp.b1 = J.b1; hp.b2 = J.b2; hp.omb1 = J.omb1; hp.omb2 = J.omb2; hp.eps = J.adam_eps;
  const float invB = 1.0f / (float)B, fB = (float)B;
  // Encoder layer 0 rows d >= DR multiply exactly-zero padding columns (gradient exactly 0): with zero Adam
  // moments flax Adam leaves them unchanged and they are never touched.  A caller-supplied state with
  // NON-zero moments there still decays: CTA 0 then applies the g = 0 Adam update to those rows every step.
  const int skip0 = J.enc_w[0] + DR * J.enc_sizes[1], skipN = (D - DR) * J.enc_sizes[1];
  int nz = 0;
  if (rank == 0)
    for (int g = tid; g < skipN; g += kMT) nz |= (J.adam_m[skip0 + g] != 0.f || J.adam_v[skip0 + g] != 0.f) ? 1 : 0;
  const bool decay_skipped = __syncthreads_or(nz) != 0;
  cluster_barrier();

  MLP_T0();
  for (int64_t s = 0; s < n_steps; ++s) {
    // ---- keys of this step (vae.py:125, datasets.py:51)
    if (tid < 4) {
      uint32_t x0 = (uint32_t)(tid & 1), x1 = (uint32_t)(tid & 1) + 2u;
      threefry2x32(ck0, ck1, x0, x1);
      const uint32_t used = tid < 2 ? x0 : x1, carry = tid < 2 ? x1 : x0;
      const uint32_t other = __shfl_xor_sync(0xFu, carry, 1);
      ck0 = (tid & 1) ? other : carry; ck1 = (tid & 1) ? carry : other;
      keys[tid] = used;                            // [lat0 lat1 bat0 bat1]
    }
    b1t *= (double)J.b1; b2t *= (double)J.b2;
    hp.ibc1 = 1.0f / (1.0f - (float)b1t); hp.ibc2 = 1.0f / (1.0f - (float)b2t);
    // per-step constants from the (possibly just updated) parameters
    if (tid < L) sl_s[tid] = expf(0.5f * __ldcg(J.params + J.off_lv + tid));
    if (tid == 0) {
      const float e = J.tunable ? __ldcg(J.params + J.off_eps) * J.eps0 : J.eps0;
      scal_s[0] = e; scal_s[1] = expf(0.5f * e); scal_s[2] = expf(-e);
    }
    __syncthreads();
    if (dual) {     // weight slice of the first forward layer: in flight during the draws (own columns: written by this CTA only)
      const LayerRef f0 = fwd_ref(J, ws, 0);
      const Slice s0 = slice_of(f0.N, rank, SL);
      stage_w_cols(Wsb[0], J.params + f0.w_off, f0.K, f0.N, s0.c0, s0.w, m_up4(s0.w));
    }
    MLP_T(kTTop);
    // ================= phase R: random draws, spread over the whole cluster
    {
      const uint32_t lk0 = keys[0], lk1 = keys[1], bk0 = keys[2], bk1 = keys[3];
      const int n_e = B * LD, half_e = (n_e + 1) >> 1, n_d = B * G, half_d = (n_d + 1) >> 1;
      float* zT = ws + J.ws_zT; float* gn = ws + J.ws_gn;
      for (int w = rank * kMT + tid; w < half_e + half_d; w += CS * kMT) {
        if (w < half_e) {
          const int j = w, j2 = w + half_e;
          uint32_t x0 = (uint32_t)j, x1 = j2 < n_e ? (uint32_t)j2 : 0u;
          threefry2x32(lk0, lk1, x0, x1);
          { const int r = j / LD; zT[(size_t)(j - r * LD) * BP + r] = normal_from_bits(x0); }
          if (j2 < n_e) { const int r = j2 / LD; zT[(size_t)(j2 - r * LD) * BP + r] = normal_from_bits(x1); }
        } else {
          const int j = w - half_e, j2 = j + half_d;
          uint32_t x0 = (uint32_t)j, x1 = j2 < n_d ? (uint32_t)j2 : 0u;
          threefry2x32(bk0, bk1, x0, x1);
          gn[j] = normal_from_bits(x0);
          if (j2 < n_d) gn[j2] = normal_from_bits(x1);
        }
      }
    }
    cluster_barrier();
    // ---- every CTA: eps rows for the encoder noise, generator transform (bit-exact order)
    load_rows(z1T, ws + J.ws_zT, L, BP);      // waited for before the reparameterisation
    for (int b = tid; b < B; b += kMT) {
      const float* gr = ws + J.ws_gn + (size_t)b * G;
      if (J.kind == VAE_DATA_SPHERE) {
        float acc = 0.f;
        for (int j = 0; j < J.dd; ++j) { const float v = __ldcg(gr + j); const float sq = __fmul_rn(v, v); acc = j == 0 ? sq : __fadd_rn(acc, sq); }
        const float rad = __fsqrt_rn(acc);
        for (int j = 0; j < J.dd; ++j) xT[j * BP + b] = __fdiv_rn(__ldcg(gr + j), rad);
      } else if (J.kind == VAE_DATA_LINEAR_GAUSSIAN) {
        for (int i = 0; i < J.dd; ++i) {
          float acc = __fmul_rn(J.A[i * G], __ldcg(gr));
          for (int k = 1; k < G; ++k) acc = __fadd_rn(acc, __fmul_rn(J.A[i * G + k], __ldcg(gr + k)));
          xT[i * BP + b] = acc;
        }
      } else {
        float acc = 0.f;
        for (int j = 0; j < J.dd; ++j) { const float v = __ldcg(gr + j); xT[j * BP + b] = v; const float pr = __fmul_rn(v, J.A[j]); acc = j == 0 ? pr : __fadd_rn(acc, pr); }
        xT[J.dd * BP + b] = det_sigmoid(acc);
      }
    }
    __syncthreads();
    MLP_T(kTR);
    const int NLf = J.n_enc + J.n_dec;
    if (dual) {
      // ================= forward, overlapped schedule: layer q multiplies out of Wsb[q & 1] while the weight slice of
      // layer q + 1 (or the transposed rows the first backward layer needs) lands in the other buffer
      for (int q = 0; q < NLf; ++q) {
        const LayerRef r = fwd_ref(J, ws, q);
        if (q == J.n_enc) {   // reparameterise (every CTA, all rows): zs = mu + exp(lv/2) * z1   (networks.py:73-74)
          load_rows(muT, ws + J.ws_enc[J.n_enc - 1], L, BP);
          cp_async_wait_all(); __syncthreads();
          for (int i = tid; i < L * BP; i += kMT) {
            const int l = i / BP, b = i - l * BP;
            zsT[i] = b < B ? fmaf(sl_s[l], z1T[i], muT[i]) : 0.f;
          }
        }
        const float* AT = r.in_g ? bufA : (r.is_dec ? zsT : xT);
        if (r.in_g && tid == 0) { m_expect_tx(&mbar, (uint32_t)(r.K * BP) * 4u); bulk_rows(bufA, r.in_g, r.K, BP, &mbar); }
        const Slice sc = slice_of(r.N, rank, SL);
        cp_async_wait_all();
        if (r.in_g) { m_mbar_wait(&mbar, mph); mph ^= 1u; }
        __syncthreads();
        MLP_T(kTFwdLoad);
        // the warps without GEMM tiles (9 of 16 at a 13-column slice) issue the copies; the others start multiplying
        const int gw = min(kMT, (2 * (m_up4(sc.w) >> 2) * (BP >> 2) + 31) & ~31);
        const int st0 = kMT - gw >= 64 ? gw : 0, snt = kMT - st0;
        if (tid >= st0) {
          if (q + 1 < NLf) {
            const LayerRef nx = fwd_ref(J, ws, q + 1);
            const Slice sn = slice_of(nx.N, rank, SL);
            stage_w_cols(Wsb[(q + 1) & 1], J.params + nx.w_off, nx.K, nx.N, sn.c0, sn.w, m_up4(sn.w), tid - st0, snt);
          } else {
            const LayerRef nx = bwd_ref(J, ws, 0);
            const Slice nk = slice_of(nx.K, rank, SL);
            stage_w_rows_t(Wsb[NLf & 1], J.params + nx.w_off, nx.N, nk.c0, nk.w, m_up4(nk.w), tid - st0, snt);
          }
        }
        tile_gemm(AT, r.K, Wsb[q & 1], m_up4(sc.w), sc.w, J.params + r.b_off, sc.c0, r.relu, nullptr, r.out_g, BP, B);
        MLP_T(kTFwdGemm);
        fence_async_global();
        cluster_barrier();
        MLP_T(kTFwdBar);
      }
    } else {
      // ================= encoder forward
      for (int i = 0; i < J.n_enc; ++i) {
        const int K = i == 0 ? DR : J.enc_sizes[i], N = J.enc_sizes[i + 1];
        const float* AT = xT;
        if (i > 0) { load_rows(bufA, ws + J.ws_enc[i - 1], K, BP); AT = bufA; }
        for (int p = 0; p < NP; ++p) {
          const Slice sc = slice_of(N, rank + p * CS, SL);
          const int W4 = m_up4(sc.w);
          if (p > 0) __syncthreads();                              // the previous pass has finished with Ws
          stage_w_cols(Ws, J.params + J.enc_w[i], K, N, sc.c0, sc.w, W4);
          cp_async_wait_all(); __syncthreads();
          tile_gemm(AT, K, Ws, W4, sc.w, J.params + J.enc_b[i], sc.c0, i + 1 < J.n_enc, nullptr, ws + J.ws_enc[i], BP, B);
        }
        cluster_barrier();
      }
      // ---- reparameterise (every CTA, all rows): zs = mu + exp(lv/2) * z1   (networks.py:73-74)
      load_rows(muT, ws + J.ws_enc[J.n_enc - 1], L, BP);
      cp_async_wait_all(); __syncthreads();
      for (int i = tid; i < L * BP; i += kMT) {
        const int l = i / BP, b = i - l * BP;
        zsT[i] = b < B ? fmaf(sl_s[l], z1T[i], muT[i]) : 0.f;
      }
      __syncthreads();
      // ================= decoder forward
      for (int i = 0; i < J.n_dec; ++i) {
        const int K = J.dec_sizes[i], N = J.dec_sizes[i + 1];
        const float* AT = zsT;
        if (i > 0) { load_rows(bufA, ws + J.ws_dec[i - 1], K, BP); AT = bufA; }
        for (int p = 0; p < NP; ++p) {
          const Slice sc = slice_of(N, rank + p * CS, SL);
          const int W4 = m_up4(sc.w);
          if (p > 0) __syncthreads();
          stage_w_cols(Ws, J.params + J.dec_w[i], K, N, sc.c0, sc.w, W4);
          cp_async_wait_all(); __syncthreads();
          tile_gemm(AT, K, Ws, W4, sc.w, J.params + J.dec_b[i], sc.c0, i + 1 < J.n_dec, nullptr, ws + J.ws_dec[i], BP, B);
        }
        cluster_barrier();
      }
    }
    // ================= output, loss, d x_hat (every CTA, all rows; reductions used by CTA 0)
    {
      const float e = scal_s[0], sd = scal_s[1], inv_var = scal_s[2];
      load_rows(bufB, ws + J.ws_dec[J.n_dec - 1], D, BP);      // decoder output
      load_rows(bufA, ws + J.ws_zT + (size_t)L * BP, D, BP);   // z2^T
      cp_async_wait_all(); __syncthreads();
      float r2 = 0.f, rz = 0.f, mu2 = 0.f;
      const float gscale = inv_var * invB;
      for (int i = tid; i < D * BP; i += kMT) {
        const int d = i / BP, b = i - d * BP;
        float g = 0.f;
        if (b < B) {
          const float z2 = bufA[i];
          const float xh = fmaf(z2, sd, bufB[i]);
          const float r = xh - (d < DR ? xT[i] : 0.f);
          r2 = fmaf(r, r, r2); rz = fmaf(r, z2, rz);
          g = r * gscale;
        }
        bufB[i] = g;                                           // dY^T of the last decoder layer
      }
      if (rank == 0) {
        for (int i = tid; i < L * BP; i += kMT) { const int b = i % BP; if (b < B) mu2 = fmaf(muT[i], muT[i], mu2); }
        const float s_r2 = block_sum(r2, red), s_rz = block_sum(rz, red), s_mu2 = block_sum(mu2, red);
        if (tid == 0) {
          float klc = 0.f;
          for (int l = 0; l < L; ++l) { const float lv = __ldcg(J.params + J.off_lv + l); klc += 1.f + lv - expf(lv); }
          const float dkl = (-0.5f * klc * fB + 0.5f * s_mu2) * invB;
          const float mse = (0.5f * inv_var * s_r2 + 0.5f * (1.8378770664093453f + e) * fB * (float)D) * invB;
          J.loss_out[s] = dkl + mse;
          if (J.tunable) {
            const float g = J.eps0 * invB * (-0.5f * inv_var * s_r2 + 0.5f * fB * (float)D + 0.5f * sd * inv_var * s_rz);
            adam_one(J.params, J.adam_m, J.adam_v, (size_t)J.off_eps, g, hp);
          }
        }
      }
      __syncthreads();
    }
    MLP_T(kTOut);
    if (dual) {
      // ================= backward, overlapped schedule (dY^T of the layer in bufB, all N rows)
      float* WsX = Wsb[NLf & 1];          // transposed weight rows of the dX product (staged one layer ahead)
      float* WsD = Wsb[(NLf - 1) & 1];    // dW slice of this layer, read by Adam
      const int NG = BP >> 2;
      {
        const LayerRef c0 = bwd_ref(J, ws, 0);
        if (c0.in_g && tid == 0) { m_expect_tx(&mbar, (uint32_t)(c0.K * BP) * 4u); bulk_rows(bufA, c0.in_g, c0.K, BP, &mbar); }
        cp_async_wait_all();
        if (c0.in_g) { m_mbar_wait(&mbar, mph); mph ^= 1u; }
        __syncthreads();
      }
      MLP_T(kTBwdFirstLoad);
      int pp = 0;
      for (int r = 0; r < NLf; ++r) {
        const LayerRef c = bwd_ref(J, ws, r);
        const bool has_dx = c.is_dec || c.i > 0;           // the encoder's first layer has no input gradient
        const bool to_latent = c.is_dec && c.i == 0;       // dX of the decoder's first layer is dzs
        const float* HT = c.in_g ? bufA : (c.is_dec ? zsT : xT);
        const Slice sk = slice_of(c.K, rank, SL);          // dX: this CTA's rows of the layer input
        const Slice sc = slice_of(c.N, rank, SL);          // dW + Adam: this CTA's output columns
        const int K4 = m_up4(sk.w), W4 = m_up4(sc.w);
        const GemmArgs g{bufB, c.N, WsX, K4, sk.w, nullptr, sk.c0, false, c.i > 0 ? HT : nullptr,
                         to_latent ? ws + J.ws_dzs : ws + J.ws_g[pp], BP, B};
        const int ntile = (K4 >> 2) * NG;
        const int nG = has_dx ? (2 * ntile + 31) & ~31 : 0;            // whole warps of dX tasks, then the dW tiles
        const int nD = sc.w > 0 ? ((c.K + 3) >> 2) * (W4 >> 2) : 0;
        for (int base = 0; base < nG + nD; base += kMT) {
          const int wt = base + (tid & ~31);                           // warp-uniform
          if (wt >= nG + nD) break;
          if (wt < nG) {
            gemm_task(g, base + tid, ntile, NG);
          } else {
            const int tile = base + tid - nG;
            if (tile < nD) dw_task(tile, HT, c.K, bufB, sc.c0, sc.w, W4, WsD, BP);
          }
        }
        { const int j = kMT - 1 - tid; if (j < sc.w) dbs[j] = db_col(bufB, sc.c0, j, BP); }   // the last warps idle above
        MLP_T(kTBwdTiles);
        // Adam may only start once every CTA has finished reading this layer's weights (the barrier)
        if (has_dx) { fence_async_global(); cluster_barrier(); } else __syncthreads();
        MLP_T(kTBwdBar);
        if (r + 1 < NLf) {     // operands of the next layer, in flight during Adam
          const LayerRef nx = bwd_ref(J, ws, r + 1);
          if (tid == 0) {
            const int nb = to_latent ? L : c.K;
            m_expect_tx(&mbar, (uint32_t)((nb + (nx.in_g ? nx.K : 0)) * BP) * 4u);
            bulk_rows(bufB, to_latent ? ws + J.ws_dzs : ws + J.ws_g[pp], nb, BP, &mbar);
            if (nx.in_g) bulk_rows(bufA, nx.in_g, nx.K, BP, &mbar);
          }
          if (!to_latent) pp ^= 1;
          if (nx.is_dec || nx.i > 0) {
            const Slice nk = slice_of(nx.K, rank, SL);
            stage_w_rows_t(WsX, J.params + nx.w_off, nx.N, nk.c0, nk.w, m_up4(nk.w));
          }
        }
        MLP_T(kTBwdIssue);
        adam_slice(J, c.w_off, c.b_off, c.K, c.N, sc.c0, sc.w, W4, WsD, dbs, hp);
        MLP_T(kTBwdAdam);
        cp_async_wait_all();
        if (r + 1 < NLf) { m_mbar_wait(&mbar, mph); mph ^= 1u; }
        __syncthreads();
        MLP_T(kTBwdWait);
        if (to_latent) {
          // ================= latent: d lv (CTA 0), dmu = dzs + mu / B  -> dY^T of the last encoder layer
          if (rank == 0) {
            for (int l = tid >> 5; l < L; l += kMT / 32) {
              float a = 0.f;
              for (int b = tid & 31; b < B; b += 32) a = fmaf(bufB[l * BP + b], z1T[l * BP + b], a);
              a = warp_sum(a);
              if ((tid & 31) == 0) {
                const float lv = __ldcg(J.params + J.off_lv + l);
                adam_one(J.params, J.adam_m, J.adam_v, (size_t)J.off_lv + l, a * 0.5f * sl_s[l] - 0.5f * (1.f - expf(lv)), hp);
              }
            }
          }
          __syncthreads();
          for (int i = tid; i < L * BP; i += kMT) { const int b = i % BP; bufB[i] = b < B ? fmaf(muT[i], invB, bufB[i]) : 0.f; }
          __syncthreads();
          MLP_T(kTLatent);
        }
      }
    } else {
      // ================= decoder backward (dY^T of layer i in bufB, all N rows)
      int pp = 0;
      for (int i = J.n_dec - 1; i >= 0; --i) {
        const int K = J.dec_sizes[i], N = J.dec_sizes[i + 1];
        const float* HT = zsT;
        if (i > 0) { load_rows(bufA, ws + J.ws_dec[i - 1], K, BP); HT = bufA; }
        // dX slices: rows k0..k0+kw of the layer input
        float* dst = i > 0 ? ws + J.ws_g[pp] : ws + J.ws_dzs;
        for (int p = 0; p < NP; ++p) {
          const Slice sk = slice_of(K, rank + p * CS, SL);
          const int K4 = m_up4(sk.w);
          if (p > 0) __syncthreads();
          stage_w_rows_t(Ws, J.params + J.dec_w[i], N, sk.c0, sk.w, K4);
          cp_async_wait_all(); __syncthreads();
          tile_gemm(bufB, N, Ws, K4, sk.w, nullptr, sk.c0, false, i > 0 ? HT : nullptr, dst, BP, B);
        }
        __syncthreads();
        // dW slices: columns c0..c0+w of the layer output.  Adam may only start once every CTA has
        // finished reading this layer's weights (the barrier); the first slice's dW is formed before it.
        for (int p = 0; p < NP; ++p) {
          const Slice sc = slice_of(N, rank + p * CS, SL);
          const int W4 = m_up4(sc.w);
          if (sc.w > 0) tile_dw(HT, K, bufB, sc.c0, sc.w, W4, Ws, dbs, BP);
          if (p == 0) cluster_barrier(); else __syncthreads();
          if (NP == 1 && i > 0) { load_rows(bufB, ws + J.ws_g[pp], K, BP); pp ^= 1; }     // next dY^T, in flight during Adam
          adam_slice(J, J.dec_w[i], J.dec_b[i], K, N, sc.c0, sc.w, W4, Ws, dbs, hp);
          if (NP > 1) __syncthreads();
        }
        if (NP > 1 && i > 0) { load_rows(bufB, ws + J.ws_g[pp], K, BP); pp ^= 1; }       // bufB was still this layer's dY^T
        cp_async_wait_all(); __syncthreads();
      }
      // ================= latent: d lv (CTA 0), dmu = dzs + mu / B  -> dY^T of the last encoder layer
      load_rows(bufB, ws + J.ws_dzs, L, BP);
      cp_async_wait_all(); __syncthreads();
      if (rank == 0) {
        // g_lv[l] = sum_b dzs*z1 * 0.5*s_l - 0.5*(1 - e^lv_l): one warp per l
        for (int l = tid >> 5; l < L; l += kMT / 32) {
          float a = 0.f;
          for (int b = tid & 31; b < B; b += 32) a = fmaf(bufB[l * BP + b], z1T[l * BP + b], a);
          a = warp_sum(a);
          if ((tid & 31) == 0) {
            const float lv = __ldcg(J.params + J.off_lv + l);
            adam_one(J.params, J.adam_m, J.adam_v, (size_t)J.off_lv + l, a * 0.5f * sl_s[l] - 0.5f * (1.f - expf(lv)), hp);
          }
        }
      }
      __syncthreads();
      for (int i = tid; i < L * BP; i += kMT) { const int b = i % BP; bufB[i] = b < B ? fmaf(muT[i], invB, bufB[i]) : 0.f; }
      __syncthreads();
      // ================= encoder backward
      for (int i = J.n_enc - 1; i >= 0; --i) {
        const int K = i == 0 ? DR : J.enc_sizes[i], N = J.enc_sizes[i + 1];
        const float* HT = xT;
        if (i > 0) { load_rows(bufA, ws + J.ws_enc[i - 1], K, BP); HT = bufA; }
        if (i > 0) {
          for (int p = 0; p < NP; ++p) {
            const Slice sk = slice_of(K, rank + p * CS, SL);
            const int K4 = m_up4(sk.w);
            if (p > 0) __syncthreads();
            stage_w_rows_t(Ws, J.params + J.enc_w[i], N, sk.c0, sk.w, K4);
            cp_async_wait_all(); __syncthreads();
            tile_gemm(bufB, N, Ws, K4, sk.w, nullptr, sk.c0, false, HT, ws + J.ws_g[pp], BP, B);
          }
          __syncthreads();
        }
        for (int p = 0; p < NP; ++p) {
          const Slice sc = slice_of(N, rank + p * CS, SL);
          const int W4 = m_up4(sc.w);
          if (sc.w > 0) tile_dw(HT, K, bufB, sc.c0, sc.w, W4, Ws, dbs, BP);
          if (p == 0 && i > 0) cluster_barrier(); else __syncthreads();
          if (NP == 1 && i > 0) { load_rows(bufB, ws + J.ws_g[pp], K, BP); pp ^= 1; }     // next dY^T, in flight during Adam
          adam_slice(J, J.enc_w[i], J.enc_b[i], K, N, sc.c0, sc.w, W4, Ws, dbs, hp);
          if (NP > 1) __syncthreads();
        }
        if (NP > 1 && i > 0) { load_rows(bufB, ws + J.ws_g[pp], K, BP); pp ^= 1; }
        cp_async_wait_all(); __syncthreads();
      }
    }
    if (decay_skipped)
      for (int g = tid; g < skipN; g += kMT) adam_one(J.params, J.adam_m, J.adam_v, (size_t)skip0 + g, 0.f, hp);
    // the next step's phase-R barrier orders this step's parameter writes before any read
  }
  cluster_barrier();
  if (rank == 0) {
    if (tid == 0) J.step[0] = step0 + (int32_t)n_steps;
    if (tid < 4 && (tid & 1) == 0) { uint32_t* dst = tid < 2 ? J.model_key : J.dataset_key; dst[0] = ck0; dst[1] = ck1; }
  }
}

// ------------------------------------------------------------------------ host side
// dual: with the second weight-staging buffer of the overlapped schedule (the kernel compares its dynamic
// shared-memory size with exactly this figure at cs = 8, see "dual" in fused_mlp_kernel)
static size_t mlp_smem_bytes(int maxw, int BP, int DR, int L, int cs, bool dual = false) {
  const int W4max = m_up4((maxw + cs - 1) / cs);
  return sizeof(float) * ((size_t)2 * maxw * BP + (size_t)(dual ? 2 : 1) * maxw * W4max + W4max + (size_t)DR * BP +
                          (size_t)3 * L * BP);
}
// VAE_MLP_DUAL=0: sequential schedule (A/B runs); read once
static bool mlp_dual_enabled() {
  static const bool on = [] { const char* e = getenv("VAE_MLP_DUAL"); return !(e && e[0] == '0'); }();
  return on;
}

static size_t mlp_ws_floats(const Layout& Lo, const VaeDesc& d, const VaeDataDesc& g, MlpJobDev* o) {
  const int BP = m_batch_stride(d.batch);
  const int gg = g.kind == VAE_DATA_LINEAR_GAUSSIAN ? g.intrinsic_dim : g.dim;
  size_t off = 0;
  auto take = [&](size_t n) { const size_t r = off; off += (n + 3) & ~(size_t)3; return (int32_t)r; };
  const int32_t zT = take((size_t)(d.latent_dim + d.data_dim) * BP);
  const int32_t gn = take((size_t)d.batch * gg);
  int32_t enc[kML], dec[kML];
  for (int i = 0; i < Lo.n_enc; ++i) enc[i] = take((size_t)Lo.enc_sizes[i + 1] * BP);
  for (int i = 0; i < Lo.n_dec; ++i) dec[i] = take((size_t)Lo.dec_sizes[i + 1] * BP);
  const int32_t g0 = take((size_t)Lo.max_width * BP), g1 = take((size_t)Lo.max_width * BP);
  const int32_t dz = take((size_t)d.latent_dim * BP);
  if (o) {
    o->ws_zT = zT; o->ws_gn = gn; o->ws_g[0] = g0; o->ws_g[1] = g1; o->ws_dzs = dz;
    for (int i = 0; i < Lo.n_enc; ++i) o->ws_enc[i] = enc[i];
    for (int i = 0; i < Lo.n_dec; ++i) o->ws_dec[i] = dec[i];
  }
  return off;
}

// 2 if the job can run on the cluster kernel, else 0
int mlp_supported(const VaeTrainJob* j, size_t* smem, size_t* ws_bytes, size_t* smem4) {
  const VaeDesc& d = j->desc;
  const VaeDataDesc& g = j->data;
  if (d.n_enc_hidden + d.n_dec_hidden == 0 || d.sig_decoder) return 0;
  if ((d.sig_decoder != 0) != (g.kind == VAE_DATA_SIGMOID)) return 0;
  Layout Lo;
  if (make_layout(&d, &Lo)) return 0;
  const int width = g.dim + g.padding_dim + (g.kind == VAE_DATA_SIGMOID ? 1 : 0);
  if (width != d.data_dim || d.latent_dim > 64 || d.batch > 4096) return 0;
  if (Lo.n_params >= (1ll << 31)) return 0;
  const int BP = m_batch_stride(d.batch);
  const int dr = g.dim + (g.kind == VAE_DATA_SIGMOID ? 1 : 0);
  const size_t sb = mlp_smem_bytes(Lo.max_width, BP, dr, d.latent_dim, kMlpMaxCluster);
  if (sb > (size_t)227 * 1024 - 2048) return 0;
  const size_t sbd = mlp_smem_bytes(Lo.max_width, BP, dr, d.latent_dim, kMlpMaxCluster, true);
  if (smem) *smem = (mlp_dual_enabled() && sbd <= (size_t)227 * 1024 - 2048) ? sbd : sb;
  if (smem4) {      // 4-CTA clusters stage wider weight slices; 0 if that does not fit
    const size_t s4 = mlp_smem_bytes(Lo.max_width, BP, dr, d.latent_dim, kMlpSmallCluster);
    *smem4 = s4 > (size_t)227 * 1024 - 2048 ? 0 : s4;
  }
  if (ws_bytes) *ws_bytes = mlp_ws_floats(Lo, d, g, nullptr) * sizeof(float);
  return 2;
}

int mlp_fill_job(const VaeTrainJob& j, MlpJobDev* o) {
  Layout Lo;
  int rc = make_layout(&j.desc, &Lo);
  if (rc) return rc;
  memset(o, 0, sizeof(*o));
  o->params = j.params; o->adam_m = j.adam_m; o->adam_v = j.adam_v; o->step = j.step;
  o->model_key = j.model_key; o->dataset_key = j.dataset_key; o->A = j.A; o->loss_out = j.loss_out;
  o->ws = (float*)j.workspace;
  o->B = j.desc.batch; o->BP = m_batch_stride(j.desc.batch); o->D = j.desc.data_dim; o->L = j.desc.latent_dim;
  o->kind = j.data.kind; o->dd = j.data.dim;
  o->DR = j.data.dim + (j.data.kind == VAE_DATA_SIGMOID ? 1 : 0);
  o->G = j.data.kind == VAE_DATA_LINEAR_GAUSSIAN ? j.data.intrinsic_dim : j.data.dim;
  o->tunable = j.desc.tunable_decoder_var; o->n_enc = Lo.n_enc; o->n_dec = Lo.n_dec; o->maxw = Lo.max_width;
  for (int i = 0; i <= Lo.n_enc; ++i) o->enc_sizes[i] = Lo.enc_sizes[i];
  for (int i = 0; i <= Lo.n_dec; ++i) o->dec_sizes[i] = Lo.dec_sizes[i];
  for (int i = 0; i < Lo.n_enc; ++i) { o->enc_w[i] = (int32_t)Lo.enc_w[i]; o->enc_b[i] = (int32_t)Lo.enc_b[i]; }
  for (int i = 0; i < Lo.n_dec; ++i) { o->dec_w[i] = (int32_t)Lo.dec_w[i]; o->dec_b[i] = (int32_t)Lo.dec_b[i]; }
  o->off_lv = (int32_t)Lo.off_eps_p; o->off_eps = (int32_t)Lo.off_eps;
  mlp_ws_floats(Lo, j.desc, j.data, o);
  const AdamHyper hp = make_adam_hyper(&j.desc);
  o->eps0 = (float)j.desc.epsilon0; o->lr = hp.lr; o->b1 = hp.b1; o->b2 = hp.b2; o->omb1 = hp.omb1; o->omb2 = hp.omb2;
  o->adam_eps = hp.eps;
  return 0;
}

size_t mlp_job_bytes() { return sizeof(MlpJobDev); }

// Function attributes are only ever set to the same constants (see fused_small.cu:set_kernel_attrs).
static int mlp_set_attrs() {
  cudaFuncAttributes fa;
  VAE_CUDA_OK(cudaFuncGetAttributes(&fa, fused_mlp_kernel));       // the kernel also has static shared memory
  VAE_CUDA_OK(cudaFuncSetAttribute(fused_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024 - (int)fa.sharedSizeBytes));
  VAE_CUDA_OK(cudaFuncSetAttribute(fused_mlp_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  return 0;
}

static int mlp_active_clusters(int cs, size_t smem, int* out) {
  if (int rc = mlp_set_attrs()) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * 64, 1, 1);
  cfg.blockDim = dim3(kMT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  VAE_CUDA_OK(cudaOccupancyMaxActiveClusters(out, fused_mlp_kernel, &cfg));
  return 0;
}

// clusters (= models) that can be resident at once on the current device (throughput packing)
int mlp_resident_jobs(size_t smem, size_t smem4, int* out) {
  int f8 = 0, f4 = 0;
  int rc = mlp_active_clusters(kMlpMaxCluster, smem, &f8);
  if (rc) return rc;
  if (smem4 > 0)      // 1 CTA per model: the largest pack that is resident at once
    for (int cs = 4; cs >= 1; cs >>= 1) {
      int f = 0;
      if (mlp_active_clusters(cs, smem4, &f) != 0) { cudaGetLastError(); f = 0; }
      if (f > f4) f4 = f;
    }
  *out = f4 > f8 ? f4 : f8;
  return 0;
}

// Cluster size of a launch.  A model's step is a serial chain of 17 cluster-wide phases, so a
// pack that leaves SMs idle (e.g. the 2 sphere runs per GPU of the 8-GPU sweep) runs each model
// on 16 CTAs: half the GEMM / Adam / staging work per CTA and phase.  Packs that need more than
// the co-resident 8-CTA clusters use 4 CTAs per model: twice the models per GPU, and the fixed
// cost of a phase (barrier, operand loads) is paid by half as many CTAs.  All sizes give
// bit-identical results.  VAE_MLP_CLUSTER=4|8|16 overrides.
static int mlp_pick_cluster(int n_jobs, size_t smem, size_t smem4) {
  const char* e = getenv("VAE_MLP_CLUSTER");
  if (e && e[0]) {
    const int v = atoi(e);
    if (v >= 16) return kMlpBigCluster;
    if (v >= 8 || smem4 == 0) return kMlpMaxCluster;
    return v >= 4 ? 4 : v >= 2 ? 2 : 1;
  }
  int fit = 0;
  if (mlp_active_clusters(kMlpBigCluster, smem, &fit) != 0) { cudaGetLastError(); fit = 0; }
  if (n_jobs <= fit) return kMlpBigCluster;
  if (smem4 == 0) return kMlpMaxCluster;
  if (mlp_active_clusters(kMlpMaxCluster, smem, &fit) != 0) { cudaGetLastError(); return kMlpMaxCluster; }
  if (n_jobs <= fit) return kMlpMaxCluster;
  for (int cs = 4; cs > 1; cs >>= 1) {        // fewer CTAs per model = more models resident at once
    if (mlp_active_clusters(cs, smem4, &fit) != 0) { cudaGetLastError(); return kMlpSmallCluster; }
    if (n_jobs <= fit) return cs;
  }
  return 1;
}

// Chooses the cluster size for the pack (occupancy queries, environment) -- once, at prepare time.
int mlp_prepare(int n_jobs, size_t smem, size_t smem4, VaePackPlan* plan) {
  if (int rc = mlp_set_attrs()) return rc;
  const int cs = mlp_pick_cluster(n_jobs, smem, smem4);
  if (cs <= kMlpSmallCluster) smem = smem4;
  plan->cluster = cs; plan->smem_bytes = (uint64_t)smem;
  return 0;
}

int mlp_launch(const VaePackPlan* plan, int64_t n_steps, cudaStream_t s) {
  const int cs = plan->cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(plan->n_jobs * cs), 1, 1);
  cfg.blockDim = dim3(kMT, 1, 1);
  cfg.dynamicSmemBytes = (size_t)plan->smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cs > 1 ? 1 : 0;
  VAE_CUDA_OK(cudaLaunchKernelEx(&cfg, fused_mlp_kernel, (const MlpJobDev*)plan->d_jobs, n_steps));
  count_launch();
  return 0;
}

}  // namespace vae
