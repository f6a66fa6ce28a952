// fused_small.cu -- persistent fused training loop for the single-Dense VAEs of the
// linear / sigmoid sweeps (seed_linpadding_expts.sh, sigmoid_vae_padding_expts.sh).
//
// 128 threads own one model for n_steps consecutive steps; a CTA holds one model (heterogeneous
// packs, fallback) or four (homogeneous packs: 512 threads, see "Shared eps draw" below).
// Everything the reference does per iteration of the hot loop (model.py:221-222) happens on-chip:
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
// Per step (two CTA barriers):
//   phase A (small_draw)         threefry blocks -> normals -> eps (B, L+D) and generator
//                                normals (B, g) in smem; lanes 0-3 advance the key chains;
//                                models of a CTA (or of a 2-CTA cluster) that hold equal model
//                                keys split the eps draw between them
//   phase B (thread per row)     generator transform, forward, loss partials, backward to
//                                dmu; operands of the weight gradients are stored
//                                TRANSPOSED ([feature][batch], batch contiguous)
//   phase C (thread per tile)    dW tiles (TI x TJ) = sum_b h[i][b] d[j][b] with 128-bit
//                                loads along the batch, split-K over KS adjacent lanes,
//                                shuffle reduce, Adam, publish new weights
#include <stdlib.h>

#include "common.cuh"

namespace vae {

constexpr int kT = 128;
constexpr int kGroupBatch = 100;      // batch size the GM > 1 kernels are compiled for (run.py:14 default)
#ifndef VAE_SMALL_MAXB
#define VAE_SMALL_MAXB 8
#endif
// 1: the models of a CTA that share an eps stream (GM > 1) keep ONE copy of it, in the z buffer of the group's
// lowest-numbered model, instead of storing every normal into each member's buffer (see "Shared eps draw")
// Launch set-up experiment (-DVAE_SMALL_PREFETCH=1): every global load of the set-up phase (state, Adam step, key chains,
// the CTA neighbours' model keys) issued right after the job record arrives, BEFORE the shared-memory zero fill, instead
// of in four to eight dependent round trips between shared-memory stores.  Measured on B200: a launch of ONE model gets
// 2 us shorter (116.8 -> 114.7 us per 20 steps), but the register allocation of the step loop shifts and the packed
// kernel loses 3 % in steady state (117.8 -> 114.4 M steps/s at 2000-step launches, 108.1 -> 106.1 M at 20): off.
#ifndef VAE_SMALL_PREFETCH
#define VAE_SMALL_PREFETCH 0
#endif
#ifndef VAE_SMALL_LEADZ
#define VAE_SMALL_LEADZ 1
#endif

struct FusedJobDev {
  float* params; float* adam_m; float* adam_v; int32_t* step;
  uint32_t* model_key; uint32_t* dataset_key; const float* A; float* loss_out;
  int32_t batch; int32_t tunable; int32_t shape_id; int32_t pad_;
  float eps0, lr, b1, b2, omb1, omb2, adam_eps;
};

__host__ __device__ constexpr int up4(int n) { return (n + 3) & ~3; }
__host__ __device__ constexpr int oddi(int n) { return n | 1; }
__host__ __device__ constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
// batch stride of the transposed operand arrays: a multiple of 4 floats whose quarter is
// odd, so 128-bit loads of 8 different feature rows hit 8 different bank groups
__host__ __device__ inline int batch_stride(int B) { int g = (B + 3) / 4; g |= 1; return 4 * g; }

template <int D_, int L_, int DR_, bool SIG_, int KIND_, int G_, int TI_, int TJ_>
struct SmallCfg {
  static constexpr int D = D_, L = L_, DR = DR_, KIND = KIND_, G = G_, TI = TI_, TJ = TJ_;
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
  static constexpr int NFLAT = D * L + L + (L * D + D) * NDEC + L;   // + 1 if tunable
  static constexpr int AN = KIND == VAE_DATA_LINEAR_GAUSSIAN ? DD * G : (KIND == VAE_DATA_SIGMOID ? DD : 0);
  // row-major activation strides (odd => row-per-thread accesses are bank-conflict free)
  static constexpr int SZ = oddi(L + D), SG = oddi(G);
  // gradient segments: (h rows I) x (d rows J); the last h row is the all-ones row, i.e.
  // the bias gradient.  seg 0: [x;1] x dmu, seg 1: [zs;1] x dxhat, seg 2: [zs;1] x dpre,
  // seg 3: [1] x (dzs*z1)  (epsilon_p)
  static constexpr int I0 = DR + 1, J0 = L, I1 = L + 1, J1 = D, I3 = 1, J3 = L;
  static constexpr int nJ0 = cdiv(J0, TJ), nJ1 = cdiv(J1, TJ), nJ3 = cdiv(J3, TJ);
  static constexpr int T0 = cdiv(I0, TI) * nJ0, T1 = cdiv(I1, TI) * nJ1, T2 = SIG ? T1 : 0, T3 = nJ3;
  static constexpr int NT = T0 + T1 + T2 + T3;
  static constexpr int KS = NT * 8 <= kT ? 8 : NT * 4 <= kT ? 4 : NT * 2 <= kT ? 2 : 1;
  static constexpr int NTASK = NT * KS;
  static constexpr int ROUNDS = cdiv(NTASK, kT);
  static constexpr int EPL = cdiv(TI * TJ, KS);          // tile elements updated per lane
  // CTAs per SM the per-shape kernel is compiled for (register cap), from the shared-memory
  // footprint at the reference batch size 100
  static constexpr int SMEM100 = 4 * (3 * PP + 4 * LP + 40 + up4(AN + 1) + 8 + up4(100 * SZ) + up4(100 * SG) +
                                      (DR + 1 + L + 1 + 2 * L + D * NDEC) * 100) + 1024;
  static constexpr int MINB_RAW = (227 * 1024) / SMEM100;
  static constexpr int MINB = MINB_RAW < 1 ? 1 : MINB_RAW > VAE_SMALL_MAXB ? VAE_SMALL_MAXB : MINB_RAW;
  struct Lay {
    int P, M, V, sl, el, kt, scal, red, A, keys, z, gn, hx, hz, dm, dx, dp, q, total, BP;
  };
  // nbuf = 2 when the launch uses clusters: eps (z) and the generator normals (gn) are then
  // double-buffered by step parity (they are drawn one step ahead, see fused_small_run)
  __host__ __device__ static Lay layout(int B, int nbuf) {
    Lay y; int o = 0;
    const int BP = batch_stride(B);
    y.BP = BP;
    y.P = o; o += PP; y.M = o; o += PP; y.V = o; o += PP;
    y.sl = o; o += LP; y.el = o; o += LP;
    y.kt = o; o += 2 * LP;         // [2 slots][LP]  1 + lv - e^lv (KL term of the loss)
    y.scal = o; o += 16;           // [2 slots][e, sd, inv_var, -, 1/bc1, 1/bc2, -, -]
    y.red = o; o += 16;            // [3][4] warp partials
    y.A = o; o += up4(AN + 1);
    y.keys = o; o += 8;            // uint32 [2 slots][lat0 lat1 bat0 bat1]
    y.z = o; o += nbuf * up4(B * SZ);
    y.gn = o; o += nbuf * up4(B * SG);
    y.hx = o; o += (DR + 1) * BP;  // transposed from here on: [feature][BP]
    y.hz = o; o += (L + 1) * BP;
    y.dm = o; o += L * BP;
    y.dx = o; o += D * BP;
    y.dp = o; o += SIG ? D * BP : 0;
    y.q = o; o += L * BP;
    y.total = up4(o);
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
};

// sqrt.approx (1 MUFU + fix-up, exact at 0): ample for Adam's denominator
__device__ __forceinline__ float approx_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// ---- thread-block-cluster helpers (shared eps draw, see fused_small_run)
__device__ __forceinline__ uint32_t cl_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared::cta address -> the same offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t cl_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cl_st(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// One step's random draws: this CTA's slice [e_lo, e_hi) of the eps block pairs (all of them
// when the CTA draws privately) and, when do_ds, all block pairs of the generator normals, as
// ONE index space spread over the CTA's threads; each thread-iteration owns an adjacent pair
// of threefry blocks (two independent chains in flight).  kq = [latent key | batch key].
// remote: eps normals are stored into the z buffer of every cluster CTA in gmask (DSMEM).
// mode 0: eps goes to this model's own z; mode 1: to the z of every cluster CTA in gmask
// (DSMEM); mode 2: to the z of every model of THIS CTA in gmask (models are mstride floats apart,
// this model is number mi).
template <class C>
__device__ __noinline__ void small_draw(float* zdst, float* gdst, const uint32_t* kq, int B, int e_lo, int e_hi,
                                        uint32_t gmask, int mode, int t, int mi, int mstride) {
  constexpr int LD = C::L + C::D, G = C::G;
  const uint32_t lk0 = kq[0], lk1 = kq[1], bk0 = kq[2], bk1 = kq[3];
  const int n_e = B * LD, half_e = (n_e + 1) >> 1, n_d = B * G, half_d = (n_d + 1) >> 1;
  const int ne = e_hi - e_lo, nd = (half_d + 1) >> 1;
  const uint32_t zaddr = (uint32_t)__cvta_generic_to_shared(zdst);
  // Iterations are dealt out from the LAST thread down, so the partial final round lands on warp 3
  // -- the warp with the least phase-C work (phase C -> next draw has no barrier in between).
#pragma unroll 1
  for (int it = kT - 1 - t; it < ne + nd; it += kT) {
    const bool isd = it >= ne;
    const int pi = isd ? it - ne : e_lo + it;
    const uint32_t k0 = isd ? bk0 : lk0, k1 = isd ? bk1 : lk1;
    const int half = isd ? half_d : half_e, n = isd ? n_d : n_e;
    const int ja = 2 * pi, jb = 2 * pi + 1;
    const bool vb = jb < half, v1 = ja + half < n, v3 = vb && jb + half < n;
    uint32_t a0 = (uint32_t)ja, a1 = v1 ? (uint32_t)(ja + half) : 0u;
    uint32_t b0 = vb ? (uint32_t)jb : 0u, b1 = v3 ? (uint32_t)(jb + half) : 0u;
    threefry2x32_x2(k0, k1, a0, a1, k0, k1, b0, b1);
    const float n0 = normal_from_bits_fast(a0), n1 = normal_from_bits_fast(a1);
    const float n2 = normal_from_bits_fast(b0), n3 = normal_from_bits_fast(b1);
    auto off_of = [&](int j) -> int {
      const int oe = C::SZ == LD ? j : (j / LD) * C::SZ + j % LD;
      const int od = C::SG == G ? j : (j / G) * C::SG + j % G;
      return isd ? od : oe;
    };
    const int o0 = off_of(ja), o1 = off_of(v1 ? ja + half : ja), o2 = off_of(vb ? jb : ja), o3 = off_of(v3 ? jb + half : ja);
    if (isd || mode == 0) {
      float* dst = isd ? gdst : zdst;
#if VAE_SMALL_LEADZ
      // contiguous rows (odd L + D, odd G) and an even half: the normals of blocks (ja, ja + 1) are neighbours in
      // both halves of the array -> two 64-bit stores (ja is even, the buffers are 16-byte aligned)
      const bool lin = isd ? C::SG == G : C::SZ == LD;
      if (lin && v3 && (half & 1) == 0) {
        *reinterpret_cast<float2*>(dst + ja) = make_float2(n0, n2);
        *reinterpret_cast<float2*>(dst + ja + half) = make_float2(n1, n3);
      } else
#endif
      {
        dst[o0] = n0;
        if (v1) dst[o1] = n1;
        if (vb) dst[o2] = n2;
        if (v3) dst[o3] = n3;
      }
    } else if (mode == 2) {           // the same four stores into the z of every group member of this CTA
      for (uint32_t m = gmask; m; m &= m - 1u) {
        float* dst = zdst + ((__ffs((int)m) - 1) - mi) * mstride;
        dst[o0] = n0;
        if (v1) dst[o1] = n1;
        if (vb) dst[o2] = n2;
        if (v3) dst[o3] = n3;
      }
    } else {                          // ... of every group member of the cluster
      for (uint32_t m = gmask; m; m &= m - 1u) {
        const uint32_t rb = cl_mapa(zaddr, (uint32_t)__ffs((int)m) - 1u);
        cl_st(rb + 4u * (uint32_t)o0, n0);
        if (v1) cl_st(rb + 4u * (uint32_t)o1, n1);
        if (vb) cl_st(rb + 4u * (uint32_t)o2, n2);
        if (v3) cl_st(rb + 4u * (uint32_t)o3, n3);
      }
    }
  }
}

// Shared eps draw.  The reference's model key is always PRNGKey(0) (model.py:29), so the runs
// of a sweep that have the same (B, L + D) draw the SAME eps stream (vae.py:125-126,
// model.py:227) and differ only in their dataset chain.  When the launch groups CTAs into
// thread-block clusters, every CTA compares its carried model key with its cluster peers' ON
// THE DEVICE; the CTAs that hold equal keys form a group, each member draws 1/|group| of the
// step's threefry blocks and stores the normals into every member's shared memory through
// DSMEM.  Values are bit-identical to a private draw; a CTA whose key matches nobody is a
// group of one and draws everything itself.
// Cluster launches run the draws ONE STEP AHEAD into double-buffered z / gn:
//   B_s (reads z[s&1]) ; draw_{s+1} -> z[(s+1)&1] of the group ; cluster arrive ; CTA barrier ;
//   C_s ; cluster wait ; CTA barrier
// so the only cluster-wide dependency ("eps of step s+1 is complete everywhere") has the
// whole of phase C as slack between arrive and wait.  The wait of iteration s-1 also proves
// that every peer finished B_{s-1}, the last reader of the buffer draw_{s+1} overwrites.
// GM > 1: one CTA of GM * 128 threads runs GM models side by side (threads [128 m, 128 m + 128)
// own model m and its slice of shared memory); the models of a CTA that hold equal keys form
// the eps-sharing group (plain shared-memory stores, CTA barriers only) -- used without clusters.
// BT > 0 fixes the batch size at compile time (the grouped kernels are built for the reference's
// batch of 100, run.py:14): the whole shared-memory layout becomes immediate offsets.
#ifdef VAE_SMALL_TIMING
__device__ unsigned long long g_small_t[8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SMALL_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_small_t[i] = gtimer(); } while (0)
extern "C" int vae_small_timing(unsigned long long* out8) { cudaDeviceSynchronize(); return (int)cudaMemcpyFromSymbol(out8, g_small_t, 64); }
#else
#define SMALL_STAMP(i) do {} while (0)
#endif

template <class C, int GM, int BT>
__device__ void fused_small_run(const FusedJobDev* jobs, int64_t n_steps, float* sm_cta) {
  SMALL_STAMP(0);
  constexpr int D = C::D, L = C::L, DR = C::DR, DD = C::DD, G = C::G, LP = C::LP, DP = C::DP;
  constexpr int TI = C::TI, TJ = C::TJ, KS = C::KS;
  constexpr int LD = L + D;
  const int mi = GM > 1 ? (int)threadIdx.x / kT : 0;             // model within the CTA
  const int t = GM > 1 ? (int)threadIdx.x % kT : (int)threadIdx.x;
  const FusedJobDev job = jobs[(int)blockIdx.x * GM + mi];
  const int B = BT > 0 ? BT : job.batch;
  const bool tunable = job.tunable != 0;
  const uint32_t csize = cl_nctarank(), crank = cl_ctarank();
  const bool pipe = csize > 1;           // cluster launch: draws run one step ahead (double-buffered)
  const typename C::Lay y = C::layout(B, pipe ? 2 : 1);
  float* sm = sm_cta + mi * y.total;       // y.total is a multiple of 4 floats
  const int BP = y.BP, NG = BP >> 2;
  float* P = sm + y.P; float* M = sm + y.M; float* V = sm + y.V;
  uint32_t* keys = reinterpret_cast<uint32_t*>(sm + y.keys);

  // ---- load state
  SMALL_STAMP(1);
  const int nflat = C::NFLAT + (tunable ? 1 : 0);
  constexpr int kSkip0 = DR * L, kSkipN = (D - DR) * L;
  constexpr int kNI = (C::NFLAT + 1 + kT - 1) / kT;                  // state elements per thread
  constexpr int kNA = (C::AN + kT - 1) / kT, kNS = (kSkipN + kT - 1) / kT;
  constexpr bool kPrefetch = VAE_SMALL_PREFETCH != 0 && kNI <= 4 && kNA <= 2 && kNS <= 2;
  float pf_p[kPrefetch ? kNI : 1], pf_m[kPrefetch ? kNI : 1], pf_v[kPrefetch ? kNI : 1], pf_a[kPrefetch && kNA > 0 ? kNA : 1];
  bool pf_skip = false;
  int32_t pf_step = 0;
  uint32_t pf_k0 = 0, pf_k1 = 0, pf_pk0[GM > 1 ? GM : 1], pf_pk1[GM > 1 ? GM : 1];
  int32_t pf_pb[GM > 1 ? GM : 1], pf_ps[GM > 1 ? GM : 1];
  if constexpr (kPrefetch) {
    if constexpr (GM > 1) {                      // the neighbours' records do not wait for this model's own
      const FusedJobDev* peers = jobs + (int)blockIdx.x * GM;
#pragma unroll
      for (int r = 0; r < GM; ++r) {
        const uint32_t* pk = peers[r].model_key;
        pf_pb[r] = peers[r].batch; pf_ps[r] = peers[r].shape_id;
        pf_pk0[r] = pk[0]; pf_pk1[r] = pk[1];
      }
    }
#pragma unroll
    for (int u = 0; u < kNI; ++u) {
      const int g = t + u * kT;
      const bool ok = g < nflat && C::padded_of(g, tunable) >= 0;
      pf_p[u] = ok ? job.params[g] : 0.f; pf_m[u] = ok ? job.adam_m[g] : 0.f; pf_v[u] = ok ? job.adam_v[g] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kNS; ++u) {
      const int g = t + u * kT;
      if (g < kSkipN) pf_skip |= job.adam_m[kSkip0 + g] != 0.f || job.adam_v[kSkip0 + g] != 0.f;
    }
#pragma unroll
    for (int u = 0; u < kNA; ++u) { const int i = t + u * kT; pf_a[u] = i < C::AN ? job.A[i] : 0.f; }
    pf_step = job.step[0];
    if (t < 4) { const uint32_t* src = t < 2 ? job.model_key : job.dataset_key; pf_k0 = src[0]; pf_k1 = src[1]; }
  }
  for (int i = t; i < y.total; i += kT) sm[i] = 0.f;
  __syncthreads();
  SMALL_STAMP(2);
  if constexpr (kPrefetch) {
#pragma unroll
    for (int u = 0; u < kNI; ++u) {
      const int g = t + u * kT;
      const int o = g < nflat ? C::padded_of(g, tunable) : -1;
      if (o >= 0) { P[o] = pf_p[u]; M[o] = pf_m[u]; V[o] = pf_v[u]; }
    }
  } else {
    for (int g = t; g < nflat; g += kT) {
      const int o = C::padded_of(g, tunable);
      if (o >= 0) { P[o] = job.params[g]; M[o] = job.adam_m[g]; V[o] = job.adam_v[g]; }
    }
  }
  // Encoder kernel rows d >= DR (flat [DR*L, D*L)) multiply exactly-zero padding columns: their gradient
  // is exactly 0 and, with zero Adam moments, flax Adam leaves them unchanged -- they are never loaded.
  // A caller-supplied state with NON-zero moments there (e.g. a checkpoint of a noisy-padding run) still
  // decays: those rows then take the g = 0 Adam update in global memory every step (rare, slow, exact).
  if constexpr (kPrefetch) {
    if (pf_skip) sm[y.red + 12] = 1.0f;                                                             // benign race: all store 1
#pragma unroll
    for (int u = 0; u < kNA; ++u) { const int i = t + u * kT; if (i < C::AN) sm[y.A + i] = pf_a[u]; }
  } else {
    if constexpr (kSkipN > 0) {
      for (int g = t; g < kSkipN; g += kT)
        if (job.adam_m[kSkip0 + g] != 0.f || job.adam_v[kSkip0 + g] != 0.f) sm[y.red + 12] = 1.0f;   // benign race: all store 1
    }
    for (int i = t; i < C::AN; i += kT) sm[y.A + i] = job.A[i];
  }
  for (int b = t; b < B; b += kT) { sm[y.hx + DR * BP + b] = 1.0f; sm[y.hz + L * BP + b] = 1.0f; }   // ones rows
  __syncthreads();
  // Derived per-step constants are double-buffered by step parity: phase B/C of step s
  // read slot s&1, the owning thread publishes the post-update values into the other slot.
  const bool decay_skipped = kSkipN > 0 && sm[y.red + 12] != 0.f;
  if (t < L) {
    const float lv = P[C::oLv + t], elv = expf(lv);
    sm[y.sl + t] = expf(0.5f * lv); sm[y.el + t] = elv;
    sm[y.kt + t] = 1.f + lv - elv; sm[y.kt + LP + t] = 1.f + lv - elv;
  }
  if (t == 0) {
    const float e = tunable ? P[C::oEp] * job.eps0 : job.eps0;
    for (int q = 0; q < 2; ++q) { sm[y.scal + 8 * q] = e; sm[y.scal + 8 * q + 1] = expf(0.5f * e); sm[y.scal + 8 * q + 2] = expf(-e); }
  }
  // Adam bias corrections 1/(1 - beta^t): thread kT-1 keeps double running products of the
  // float32 betas and publishes the float32 factors for the NEXT step (slot parity)
  const int32_t step0 = kPrefetch ? pf_step : job.step[0];
  double b1t = 0.0, b2t = 0.0;
  if (t == kT - 1) {
    b1t = pow((double)job.b1, (double)step0 + 1.0); b2t = pow((double)job.b2, (double)step0 + 1.0);
    sm[y.scal + 4] = 1.0f / (1.0f - (float)b1t); sm[y.scal + 5] = 1.0f / (1.0f - (float)b2t);
  }

  // ---- key chains: lanes 0,1 own the model-side split, lanes 2,3 the dataset-side one
  uint32_t ck0 = 0, ck1 = 0;           // carried key of this lane's chain
  if constexpr (kPrefetch) { ck0 = pf_k0; ck1 = pf_k1; }
  else if (t < 4) {
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
  // ---- eps-sharing group inside the cluster (a cluster of one: no cluster instructions at all)
  // members: the cluster's CTAs (GM == 1) or this CTA's models (GM > 1, never both)
  const uint32_t self = GM > 1 ? (uint32_t)mi : crank;
  const uint32_t nmemb = GM > 1 ? (uint32_t)GM : csize;
  uint32_t gmask = 1u << self;
  if constexpr (kPrefetch && GM > 1) {
    uint32_t mk0 = 0u, mk1 = 0u;
#pragma unroll
    for (int r = 0; r < GM; ++r) if (r == mi) { mk0 = pf_pk0[r]; mk1 = pf_pk1[r]; }
    gmask = 0u;
#pragma unroll
    for (int r = 0; r < GM; ++r)
      if (pf_pk0[r] == mk0 && pf_pk1[r] == mk1 && pf_pb[r] == job.batch && pf_ps[r] == job.shape_id) gmask |= 1u << r;
  } else if (nmemb > 1) {
    const uint32_t mk0 = job.model_key[0], mk1 = job.model_key[1];
    const FusedJobDev* peers = GM > 1 ? jobs + (int)blockIdx.x * GM : jobs + ((int)blockIdx.x - (int)crank);
    gmask = 0u;
    for (uint32_t r = 0; r < nmemb; ++r) {
      const FusedJobDev& o = peers[r];
      if (o.model_key[0] == mk0 && o.model_key[1] == mk1 && o.batch == job.batch && o.shape_id == job.shape_id)
        gmask |= 1u << r;
    }
  }
  const int gsize = __popc(gmask), gidx = __popc(gmask & ((1u << self) - 1u));
  constexpr bool kLeadZ = VAE_SMALL_LEADZ != 0 && GM > 1;
  const int draw_mode = gsize == 1 || kLeadZ ? 0 : GM > 1 ? 2 : 1;
  // kLeadZ: the group's eps is stored once, into its lowest-numbered member's z; every member reads it there
  const int lead = kLeadZ ? __ffs((int)gmask) - 1 : mi;
  if (kLeadZ && t == 0) reinterpret_cast<int*>(sm + y.red)[13] = lead;
  __syncthreads();
  SMALL_STAMP(3);

  const float invB = 1.0f / (float)B;
  const int npair_e = (((B * LD + 1) >> 1) + 1) >> 1;     // eps block pairs per step
  int e_lo = 0, e_hi = npair_e;                            // this CTA's slice of them
  if (gsize > 1) {
    const int chunk = (npair_e + gsize - 1) / gsize;
    e_lo = min(npair_e, gidx * chunk);
    e_hi = min(npair_e, e_lo + chunk);
  }
  const int zstride = up4(B * C::SZ), gstride = up4(B * C::SG);
  // draws of step q: keys[q&1] -> z / gn buffer (q&1 when pipelined); lanes 0-3 then advance the
  // key chains for step q+1 into the other slot
  auto draw_step = [&](int64_t q) {
    const int qs = (int)(q & 1), buf = pipe ? qs : 0;
    small_draw<C>((kLeadZ ? sm_cta + lead * y.total : sm) + y.z + buf * zstride, sm + y.gn + buf * gstride, keys + qs * 4, B, e_lo, e_hi,
                  gmask, draw_mode, t, mi, y.total);
    if (t < 4 && q + 1 < n_steps) advance_keys(qs ^ 1);
  };
  if (pipe) {
    cl_arrive(); cl_wait();       // every peer's shared memory is initialised and its keys are read
    if (n_steps > 0) draw_step(0);
    cl_arrive(); cl_wait();       // eps of step 0 complete everywhere
  }

  for (int64_t s = 0; s < n_steps; ++s) {
    const int slot = (int)(s & 1);
    const int zb = pipe ? slot : 0;
    // ================= phase A: random draws (private mode: in step; cluster mode: one step ahead)
    if (!pipe) { draw_step(s); __syncthreads(); }
    // ================= phase B: one thread per batch row
    // With 100 rows for 128 threads the fourth warp of a model has 4 busy lanes.  In the grouped kernel (4 models per
    // CTA, B = 100) model 0's fourth warp takes the 4 leftover rows of ALL four models (lane -> model lane / 4, row
    // 96 + lane % 4) and the other models' fourth warps skip phase B: one nearly empty warp pass instead of four.
    // The arithmetic per row and the order of every sum are unchanged (see the reduction below).
    constexpr bool kPackTail = GM == 4 && BT == 100;
    const bool tail_warp = kPackTail && t >= 96;
    const int mb = tail_warp ? ((t - 96) >> 2) & 3 : mi;                       // the model whose row this thread works on
    float* smB = tail_warp ? sm_cta + mb * y.total : sm;
    const float* Pb = smB + y.P;
    const float* zB = !kLeadZ ? smB + y.z : sm_cta + (tail_warp ? reinterpret_cast<const int*>(smB + y.red)[13] : lead) * y.total + y.z;
    const int b_first = !tail_warp ? t : (mi == 0 && t < 112) ? 96 + ((t - 96) & 3) : B;
    float kl_acc = 0.f, r2_acc = 0.f, rz_acc = 0.f;
    {
      const float sd = smB[y.scal + 8 * slot + 1], inv_var = smB[y.scal + 8 * slot + 2];
      const float gscale = inv_var * invB;
      for (int b = b_first; b < B; b += kT) {
        const float* zr = zB + zb * zstride + b * C::SZ;
        const float* gr = smB + y.gn + zb * gstride + b * C::SG;
        // ---- generator transform (bit-exact with oracle/datasets.py)
        float x[DR];
        if constexpr (C::KIND == VAE_DATA_SPHERE) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < DD; ++j) { x[j] = gr[j]; const float sq = __fmul_rn(x[j], x[j]); acc = j == 0 ? sq : __fadd_rn(acc, sq); }
          const float rad = __fsqrt_rn(acc);
#pragma unroll
          for (int j = 0; j < DD; ++j) x[j] = __fdiv_rn(x[j], rad);
        } else if constexpr (C::KIND == VAE_DATA_LINEAR_GAUSSIAN) {
          const float* A = smB + y.A;
#pragma unroll
          for (int i = 0; i < DD; ++i) {
            float acc = __fmul_rn(A[i * G], gr[0]);
#pragma unroll
            for (int k = 1; k < G; ++k) acc = __fadd_rn(acc, __fmul_rn(A[i * G + k], gr[k]));
            x[i] = acc;
          }
        } else {
          const float* A = smB + y.A;
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < DD; ++j) { x[j] = gr[j]; const float pr = __fmul_rn(x[j], A[j]); acc = j == 0 ? pr : __fadd_rn(acc, pr); }
          x[DD] = det_sigmoid(acc);
        }
        float* hxc = smB + y.hx + b;
#pragma unroll
        for (int d = 0; d < DR; ++d) hxc[d * BP] = x[d];
        // ---- encoder: mu = x We + be   (only the DR non-padding columns contribute)
        float mu[LP];
#pragma unroll
        for (int q = 0; q < LP / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(Pb + C::oBe + 4 * q);
          mu[4 * q] = v.x; mu[4 * q + 1] = v.y; mu[4 * q + 2] = v.z; mu[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int d = 0; d < DR; ++d) {
#pragma unroll
          for (int q = 0; q < LP / 4; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(Pb + C::oWe + d * LP + 4 * q);
            mu[4 * q] = fmaf(x[d], w.x, mu[4 * q]); mu[4 * q + 1] = fmaf(x[d], w.y, mu[4 * q + 1]);
            mu[4 * q + 2] = fmaf(x[d], w.z, mu[4 * q + 2]); mu[4 * q + 3] = fmaf(x[d], w.w, mu[4 * q + 3]);
          }
        }
        // ---- reparameterise, decoder(s)
        float dec[DP], sp[C::SIG ? DP : 1];
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(Pb + C::oBd + 4 * q);
          dec[4 * q] = v.x; dec[4 * q + 1] = v.y; dec[4 * q + 2] = v.z; dec[4 * q + 3] = v.w;
          if constexpr (C::SIG) {
            const float4 u = *reinterpret_cast<const float4*>(Pb + C::oBs + 4 * q);
            sp[4 * q] = u.x; sp[4 * q + 1] = u.y; sp[4 * q + 2] = u.z; sp[4 * q + 3] = u.w;
          }
        }
        float* hzc = smB + y.hz + b;
#pragma unroll
        for (int l = 0; l < L; ++l) {
          kl_acc = fmaf(mu[l], mu[l], kl_acc);
          const float zsl = fmaf(smB[y.sl + l], zr[l], mu[l]);
          hzc[l * BP] = zsl;
#pragma unroll
          for (int q = 0; q < DP / 4; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(Pb + C::oWd + l * DP + 4 * q);
            dec[4 * q] = fmaf(zsl, w.x, dec[4 * q]); dec[4 * q + 1] = fmaf(zsl, w.y, dec[4 * q + 1]);
            dec[4 * q + 2] = fmaf(zsl, w.z, dec[4 * q + 2]); dec[4 * q + 3] = fmaf(zsl, w.w, dec[4 * q + 3]);
            if constexpr (C::SIG) {
              const float4 u = *reinterpret_cast<const float4*>(Pb + C::oWs + l * DP + 4 * q);
              sp[4 * q] = fmaf(zsl, u.x, sp[4 * q]); sp[4 * q + 1] = fmaf(zsl, u.y, sp[4 * q + 1]);
              sp[4 * q + 2] = fmaf(zsl, u.z, sp[4 * q + 2]); sp[4 * q + 3] = fmaf(zsl, u.w, sp[4 * q + 3]);
            }
          }
        }
        // ---- output, residual, d x_hat  (dec[] becomes d x_hat, sp[] becomes d sig_pre)
        float* dxc = smB + y.dx + b;
        float* dpc = smB + y.dp + b;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          if (d < D) {
            const float z2 = zr[L + d];
            float xh = dec[d];
            float sg = 0.f;
            if constexpr (C::SIG) { sg = fast_sigmoid(sp[d]); xh += sg; }
            xh = fmaf(z2, sd, xh);
            const float r = xh - (d < DR ? x[d < DR ? d : 0] : 0.f);
            r2_acc = fmaf(r, r, r2_acc);
            rz_acc = fmaf(r, z2, rz_acc);
            const float g = r * gscale;
            dec[d] = g; dxc[d * BP] = g;
            if constexpr (C::SIG) { const float gp = g * sg * (1.f - sg); sp[d] = gp; dpc[d * BP] = gp; }
          } else {
            dec[d] = 0.f;
            if constexpr (C::SIG) sp[d] = 0.f;
          }
        }
        // ---- dzs = dxhat Wd^T (+ dpre Ws^T); dmu = dzs + mu/B; q = dzs * z1
        float* dmc = smB + y.dm + b;
        float* qc = smB + y.q + b;
#pragma unroll
        for (int l = 0; l < L; ++l) {
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int q = 0; q < DP / 4; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(Pb + C::oWd + l * DP + 4 * q);
            a0 = fmaf(dec[4 * q], w.x, a0); a1 = fmaf(dec[4 * q + 1], w.y, a1);
            a0 = fmaf(dec[4 * q + 2], w.z, a0); a1 = fmaf(dec[4 * q + 3], w.w, a1);
            if constexpr (C::SIG) {
              const float4 u = *reinterpret_cast<const float4*>(Pb + C::oWs + l * DP + 4 * q);
              a0 = fmaf(sp[4 * q], u.x, a0); a1 = fmaf(sp[4 * q + 1], u.y, a1);
              a0 = fmaf(sp[4 * q + 2], u.z, a0); a1 = fmaf(sp[4 * q + 3], u.w, a1);
            }
          }
          const float dz = a0 + a1;
          dmc[l * BP] = fmaf(mu[l], invB, dz);
          qc[l * BP] = dz * zr[l];
        }
      }
      if (tail_warp) {
        // four lanes per model: the same pairing as the last two levels of warp_sum ((l0 + l2) + (l1 + l3)); a full
        // warp_sum over this warp would only have added exact zeros before them
        if (mi == 0) {
          kl_acc += __shfl_xor_sync(0xffffffffu, kl_acc, 2); r2_acc += __shfl_xor_sync(0xffffffffu, r2_acc, 2); rz_acc += __shfl_xor_sync(0xffffffffu, rz_acc, 2);
          kl_acc += __shfl_xor_sync(0xffffffffu, kl_acc, 1); r2_acc += __shfl_xor_sync(0xffffffffu, r2_acc, 1); rz_acc += __shfl_xor_sync(0xffffffffu, rz_acc, 1);
          if (t < 112 && (t & 3) == 0) { smB[y.red + 0 + 3] = kl_acc; smB[y.red + 4 + 3] = r2_acc; smB[y.red + 8 + 3] = rz_acc; }
        }
      } else {
        kl_acc = warp_sum(kl_acc); r2_acc = warp_sum(r2_acc); rz_acc = warp_sum(rz_acc);
        if ((t & 31) == 0) {
          sm[y.red + 0 + (t >> 5)] = kl_acc; sm[y.red + 4 + (t >> 5)] = r2_acc; sm[y.red + 8 + (t >> 5)] = rz_acc;
        }
      }
    }
    if (pipe) {                     // next step's draws; then "my share of eps_{s+1} is stored everywhere"
      if (s + 1 < n_steps) draw_step(s + 1);
      cl_arrive();
    }
    __syncthreads();
    // ================= phase C: gradient tiles + Adam
    {
      const float* scs = sm + y.scal + 8 * slot;
      const float e = scs[0], sd = scs[1], inv_var = scs[2], ibc1 = scs[4], ibc2 = scs[5];
      const float s_mu2 = (sm[y.red + 0] + sm[y.red + 1]) + (sm[y.red + 2] + sm[y.red + 3]);
      const float s_r2 = (sm[y.red + 4] + sm[y.red + 5]) + (sm[y.red + 6] + sm[y.red + 7]);
      const float s_rz = (sm[y.red + 8] + sm[y.red + 9]) + (sm[y.red + 10] + sm[y.red + 11]);
      const float fB = (float)B;
      auto adam = [&](int poff, float g) -> float {
        const float m = job.b1 * M[poff] + job.omb1 * g;
        const float v = job.b2 * V[poff] + job.omb2 * g * g;
        M[poff] = m; V[poff] = v;
        const float p = P[poff] - __fdividef(job.lr * (m * ibc1), approx_sqrt(v * ibc2) + job.adam_eps);
        P[poff] = p;
        return p;
      };
      if (t == kT - 1) {   // loss = mean_b(Dkl + mse) (networks.py:94-98); epsilon parameter
        float klc = 0.f;
#pragma unroll
        for (int l = 0; l < L; ++l) klc += sm[y.kt + slot * LP + l];
        const float dkl = (-0.5f * klc * fB + 0.5f * s_mu2) * invB;
        const float mse = (0.5f * inv_var * s_r2 + 0.5f * (1.8378770664093453f + e) * fB * (float)D) * invB;
        job.loss_out[s] = dkl + mse;
        float* scn = sm + y.scal + 8 * (slot ^ 1);
        b1t *= (double)job.b1; b2t *= (double)job.b2;
        scn[4] = 1.0f / (1.0f - (float)b1t); scn[5] = 1.0f / (1.0f - (float)b2t);
        if (tunable) {     // d loss / d epsilon (param) = eps0 * dL/de
          const float g = job.eps0 * invB * (-0.5f * inv_var * s_r2 + 0.5f * fB * (float)D + 0.5f * sd * inv_var * s_rz);
          const float e2 = adam(C::oEp, g) * job.eps0;
          scn[0] = e2; scn[1] = expf(0.5f * e2); scn[2] = expf(-e2);
        }
      }
      if constexpr (kSkipN > 0) {
        if (decay_skipped) {          // g = 0 Adam on the never-loaded encoder rows (same arithmetic as adam())
          for (int g = t; g < kSkipN; g += kT) {
            const int gi = kSkip0 + g;
            const float m = job.b1 * job.adam_m[gi] + job.omb1 * 0.f;
            const float v = job.b2 * job.adam_v[gi] + job.omb2 * 0.f * 0.f;
            job.adam_m[gi] = m; job.adam_v[gi] = v;
            job.params[gi] = job.params[gi] - __fdividef(job.lr * (m * ibc1), approx_sqrt(v * ibc2) + job.adam_eps);
          }
        }
      }
#pragma unroll 1
      for (int rnd = 0; rnd < C::ROUNDS; ++rnd) {
        if (rnd * kT + (t & ~31) >= C::NTASK) break;                    // warp-uniform
        const int gt = rnd * kT + t;
        const bool valid = gt < C::NTASK;
        const int task = valid ? gt : C::NTASK - 1;
        const int tile = task / KS, ks = task % KS;
        // decode segment
        int seg, loc;
        if (tile < C::T0) { seg = 0; loc = tile; }
        else if (tile < C::T0 + C::T1) { seg = 1; loc = tile - C::T0; }
        else if (tile < C::T0 + C::T1 + C::T2) { seg = 2; loc = tile - C::T0 - C::T1; }
        else { seg = 3; loc = tile - C::T0 - C::T1 - C::T2; }
        const int nJ = seg == 0 ? C::nJ0 : seg == 3 ? C::nJ3 : C::nJ1;
        const int I = seg == 0 ? C::I0 : seg == 3 ? C::I3 : C::I1;
        const int J = seg == 0 ? C::J0 : seg == 3 ? C::J3 : C::J1;
        const int hb = seg == 0 ? y.hx : seg == 3 ? y.hz + L * BP : y.hz;
        const int db = seg == 0 ? y.dm : seg == 1 ? y.dx : seg == 2 ? y.dp : y.q;
        const int ti = loc / nJ, tj = loc - ti * nJ;
        const int i0 = ti * TI, j0 = tj * TJ;
        const float* hp[TI]; const float* dq[TJ];
#pragma unroll
        for (int ii = 0; ii < TI; ++ii) hp[ii] = sm + hb + min(i0 + ii, I - 1) * BP;
#pragma unroll
        for (int jj = 0; jj < TJ; ++jj) dq[jj] = sm + db + min(j0 + jj, J - 1) * BP;
        float acc[TI][TJ];
#pragma unroll
        for (int ii = 0; ii < TI; ++ii)
#pragma unroll
          for (int jj = 0; jj < TJ; ++jj) acc[ii][jj] = 0.f;
#pragma unroll 2
        for (int g4 = ks; g4 < NG; g4 += KS) {
          float4 hv[TI], dv[TJ];
#pragma unroll
          for (int ii = 0; ii < TI; ++ii) hv[ii] = *reinterpret_cast<const float4*>(hp[ii] + 4 * g4);
#pragma unroll
          for (int jj = 0; jj < TJ; ++jj) dv[jj] = *reinterpret_cast<const float4*>(dq[jj] + 4 * g4);
#pragma unroll
          for (int ii = 0; ii < TI; ++ii)
#pragma unroll
            for (int jj = 0; jj < TJ; ++jj) {
              float a = acc[ii][jj];
              a = fmaf(hv[ii].x, dv[jj].x, a); a = fmaf(hv[ii].y, dv[jj].y, a);
              a = fmaf(hv[ii].z, dv[jj].z, a); a = fmaf(hv[ii].w, dv[jj].w, a);
              acc[ii][jj] = a;
            }
        }
        // split-K reduction over the KS adjacent lanes of this tile
#pragma unroll
        for (int off = 1; off < KS; off <<= 1)
#pragma unroll
          for (int ii = 0; ii < TI; ++ii)
#pragma unroll
            for (int jj = 0; jj < TJ; ++jj) acc[ii][jj] += __shfl_xor_sync(0xffffffffu, acc[ii][jj], off);
        // each lane of the group updates the tile elements e = r*KS + ks
#pragma unroll
        for (int r = 0; r < C::EPL; ++r) {
          const int el = r * KS + ks;
          float g = 0.f;
#pragma unroll
          for (int q = 0; q < TI * TJ; ++q) g = (el == q) ? acc[q / TJ][q % TJ] : g;
          const int i = i0 + el / TJ, j = j0 + el % TJ;
          if (valid && el < TI * TJ && i < I && j < J) {
            if (seg == 3) {          // d loss / d epsilon_p[j]
              g = g * 0.5f * sm[y.sl + j] - 0.5f * (1.f - sm[y.el + j]);
              const float lv = adam(C::oLv + j, g);
              const float elv = expf(lv);
              sm[y.sl + j] = expf(0.5f * lv); sm[y.el + j] = elv;
              sm[y.kt + (slot ^ 1) * LP + j] = 1.f + lv - elv;
            } else {
              const int IW = seg == 0 ? DR : L;                       // row IW is the bias row
              const int ldp = seg == 0 ? LP : DP;
              const int oW = seg == 0 ? C::oWe : seg == 1 ? C::oWd : C::oWs;
              const int oB = seg == 0 ? C::oBe : seg == 1 ? C::oBd : C::oBs;
              adam(i < IW ? oW + i * ldp + j : oB + j, g);
            }
          }
        }
      }
    }
    // no barrier here: phase A of the next step touches only z / gn / keys, which phase C
    // does not read; the A->B barrier orders the parameter writes before the next reads.
    if (pipe) { cl_wait(); __syncthreads(); }     // eps_{s+1} complete everywhere; phase C done in this CTA
  }
  __syncthreads();
  SMALL_STAMP(4);
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
  SMALL_STAMP(5);
}

// (id, D, L, DR, SIG, KIND, G, TI, TJ): every single-Dense shape of the two sweeps, plus one
// sphere shape without hidden layers (generator coverage).  TI x TJ = gradient tile.
#define VAE_SMALL_SHAPES(X)                                         \
  X(0, 12, 20, 3, false, VAE_DATA_LINEAR_GAUSSIAN, 3, 2, 4)         \
  X(1, 20, 20, 3, false, VAE_DATA_LINEAR_GAUSSIAN, 3, 2, 4)         \
  X(2, 12, 20, 6, false, VAE_DATA_LINEAR_GAUSSIAN, 3, 2, 4)         \
  X(3, 20, 20, 6, false, VAE_DATA_LINEAR_GAUSSIAN, 3, 2, 4)         \
  X(4, 12, 20, 9, false, VAE_DATA_LINEAR_GAUSSIAN, 3, 2, 4)         \
  X(5, 20, 10, 9, false, VAE_DATA_LINEAR_GAUSSIAN, 3, 2, 4)         \
  X(6, 20, 10, 12, false, VAE_DATA_LINEAR_GAUSSIAN, 3, 2, 4)        \
  X(7, 7, 6, 4, true, VAE_DATA_SIGMOID, 3, 2, 4)                    \
  X(8, 17, 8, 4, true, VAE_DATA_SIGMOID, 3, 2, 4)                   \
  X(9, 22, 16, 6, true, VAE_DATA_SIGMOID, 5, 4, 4)                  \
  X(10, 11, 10, 6, true, VAE_DATA_SIGMOID, 5, 2, 4)                 \
  X(11, 15, 13, 8, true, VAE_DATA_SIGMOID, 7, 4, 4)                 \
  X(12, 28, 24, 8, true, VAE_DATA_SIGMOID, 7, 4, 4)                 \
  X(13, 6, 6, 3, false, VAE_DATA_SPHERE, 3, 2, 2)

// heterogeneous packs: one launch, dispatch on the job's shape
__global__ void __launch_bounds__(kT)
fused_small_kernel(const FusedJobDev* __restrict__ jobs, int64_t n_steps) {
  extern __shared__ __align__(16) float sm[];
  const FusedJobDev job = jobs[blockIdx.x];
  switch (job.shape_id) {
#define X(id, D, L, DR, SIG, KIND, G, TI, TJ) \
    case id: fused_small_run<SmallCfg<D, L, DR, SIG, KIND, G, TI, TJ>, 1, 0>(jobs, n_steps, sm); break;
    VAE_SMALL_SHAPES(X)
#undef X
    default: break;
  }
}

// homogeneous packs: a kernel per shape, so registers / occupancy are the shape's own
template <class C, int GM>
__global__ void __launch_bounds__(kT * GM, (C::MINB / GM) > 0 ? (C::MINB / GM) : 1)
fused_small_kernel_one(const FusedJobDev* __restrict__ jobs, int64_t n_steps) {
  extern __shared__ __align__(16) float sm[];
  fused_small_run<C, GM, (GM > 1 ? kGroupBatch : 0)>(jobs, n_steps, sm);
}

// Returns the shape id (>= 0) or -1; *smem_floats receives the CTA's shared memory need.
static int small_shape_of(const VaeTrainJob* j, int* smem_floats, int nbuf = 1) {
  const VaeDesc& d = j->desc;
  const VaeDataDesc& g = j->data;
  if (d.n_enc_hidden != 0 || d.n_dec_hidden != 0) return -1;
  if (d.batch <= 0 || d.batch > 4096) return -1;
  const int width = g.dim + g.padding_dim + (g.kind == VAE_DATA_SIGMOID ? 1 : 0);
  if (width != d.data_dim) return -1;
  if ((d.sig_decoder != 0) != (g.kind == VAE_DATA_SIGMOID)) return -1;
  const int dr = g.dim + (g.kind == VAE_DATA_SIGMOID ? 1 : 0);
  const int gg = g.kind == VAE_DATA_LINEAR_GAUSSIAN ? g.intrinsic_dim : g.dim;
#define X(id, D, L, DR, SIG, KIND, G, TI, TJ)                                                \
  if (d.data_dim == D && d.latent_dim == L && dr == DR && g.kind == KIND && gg == G) {       \
    *smem_floats = SmallCfg<D, L, DR, SIG, KIND, G, TI, TJ>::layout(d.batch, nbuf).total;          \
    return id;                                                                               \
  }
  VAE_SMALL_SHAPES(X)
#undef X
  return -1;
}

constexpr int kMaxSmem = 227 * 1024;

// VAE_SMALL_CLUSTER=n caps the cluster size, VAE_SMALL_GM=n the models per CTA (1 = off; A/B runs)
static int env_cap(const char* name, int dflt, int hi) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  const int v = atoi(e);
  return v < 1 ? 1 : v > hi ? hi : v;
}
#ifndef VAE_SMALL_GM_MODELS
#define VAE_SMALL_GM_MODELS 4
#endif
constexpr int kGM = VAE_SMALL_GM_MODELS;   // models per CTA of the grouped kernels

// How a homogeneous pack of n_jobs models is launched: GM models per CTA (eps shared inside the
// CTA) when the pack divides and GM footprints fit one SM's shared memory; otherwise one model
// per CTA in clusters of two (eps shared through DSMEM; pairs tile every GPC, larger clusters
// strand SMs); otherwise plain CTAs.
struct SmallPlan { int gm, cluster, nbuf; size_t smem; };
static SmallPlan plan_small(const VaeTrainJob* job0, int n_jobs, bool homogeneous) {
  SmallPlan p = {1, 1, 1, 0};
  int fl1 = 0, fl2 = 0;
  small_shape_of(job0, &fl1, 1);
  small_shape_of(job0, &fl2, 2);
  p.smem = (size_t)fl1 * 4;
  if (!homogeneous) return p;
  // models resident per SM under each plan, from the shared-memory footprints (+1 KB the driver
  // reserves per CTA) and the register cap of the kernels (8 x 128 threads or 2 x 512 threads):
  // sharing must not cost occupancy -- the big shapes (e.g. D = 28, L = 24: 100 KB per model) keep
  // one model per CTA and private draws.
  auto fit = [](size_t bytes, int cap) { const int n = (int)((size_t)kMaxSmem / (bytes + 1024)); return n < cap ? n : cap; };
  const int per_private = fit((size_t)fl1 * 4, 8);
  const int per_gm = kGM * fit((size_t)fl1 * 4 * kGM, 8 / kGM > 0 ? 8 / kGM : 1);
  const int per_cl = fit((size_t)fl2 * 4, 8);
  int dev = 0, sms = 148;                      // a pack that is fully resident anyway may always share
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError(); sms = 148;
  }
  if (env_cap("VAE_SMALL_GM", kGM, kGM) == kGM && n_jobs % kGM == 0 && job0->desc.batch == kGroupBatch &&
      (size_t)fl1 * 4 * kGM <= (size_t)kMaxSmem && (per_gm >= per_private || n_jobs <= per_gm * sms)) {
    p.gm = kGM; p.smem = (size_t)fl1 * 4 * kGM;
    return p;
  }
  const int cap = env_cap("VAE_SMALL_CLUSTER", 2, 8);
  if (cap >= 2 && n_jobs % 2 == 0 && (size_t)fl2 * 4 <= (size_t)kMaxSmem && (per_cl >= per_private || n_jobs <= per_cl * sms)) {
    int c = 2;
    while (c * 2 <= cap && n_jobs % (c * 2) == 0) c *= 2;
    p.cluster = c; p.nbuf = 2; p.smem = (size_t)fl2 * 4;
  }
  return p;
}

// Function attributes are per-device mutable state shared by every host thread: they are only ever set
// to the SAME constants (the 227 KB maximum, carve-out 100), at prepare time, so concurrent packs of one
// kernel with different shared-memory needs cannot race; the launch itself passes its own size.
template <class K>
static int set_kernel_attrs(K k) {
  VAE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  VAE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  return 0;
}

template <class K>
static int launch_one(K k, bool prepare_only, int n_ctas, int threads, int cluster, size_t smem, const FusedJobDev* d_jobs,
                      int64_t n_steps, cudaStream_t s) {
  if (prepare_only) return set_kernel_attrs(k);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_ctas, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cluster > 1 ? 1 : 0;
  VAE_CUDA_OK(cudaLaunchKernelEx(&cfg, k, d_jobs, n_steps));
  return 0;
}

// prepare_only: set the kernel's function attributes (vae_pack_prepare); otherwise launch (vae_pack_train)
static int launch_small(int shape_id, bool homogeneous, const SmallPlan& p, int n_jobs, const FusedJobDev* d_jobs,
                        int64_t n_steps, cudaStream_t s, bool prepare_only) {
  if (homogeneous) {
    int rc = 0;
    switch (shape_id) {
#define X(id, D, L, DR, SIG, KIND, G, TI, TJ)                                                              \
      case id:                                                                                             \
        if (p.gm == kGM)                                                                                   \
          rc = launch_one(fused_small_kernel_one<SmallCfg<D, L, DR, SIG, KIND, G, TI, TJ>, kGM>,           \
                          prepare_only, n_jobs / kGM, kT * kGM, 1, p.smem, d_jobs, n_steps, s);            \
        else                                                                                               \
          rc = launch_one(fused_small_kernel_one<SmallCfg<D, L, DR, SIG, KIND, G, TI, TJ>, 1>,             \
                          prepare_only, n_jobs, kT, p.cluster, p.smem, d_jobs, n_steps, s);                \
        break;
      VAE_SMALL_SHAPES(X)
#undef X
      default: set_error("bad shape id %d", shape_id); return VAE_ERR_INVALID;
    }
    if (rc || prepare_only) return rc;
  } else {
    if (prepare_only) return set_kernel_attrs(fused_small_kernel);
    fused_small_kernel<<<n_jobs, kT, p.smem, s>>>(d_jobs, n_steps);
  }
  VAE_LAUNCH_OK("fused_small_kernel");
  return 0;
}

// models resident per SM when a full homogeneous pack of this shape is launched (plan_small)
static int occupancy_small(int shape_id, const SmallPlan& p, int* out) {
  int blocks = 0;
  switch (shape_id) {
#define X(id, D, L, DR, SIG, KIND, G, TI, TJ)                                                              \
    case id: {                                                                                             \
      if (p.gm == kGM) {                                                                                   \
        auto k = fused_small_kernel_one<SmallCfg<D, L, DR, SIG, KIND, G, TI, TJ>, kGM>;                    \
        if (int rc_ = set_kernel_attrs(k)) return rc_;                                                     \
        VAE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k, kT * kGM, p.smem));          \
      } else {                                                                                             \
        auto k = fused_small_kernel_one<SmallCfg<D, L, DR, SIG, KIND, G, TI, TJ>, 1>;                      \
        if (int rc_ = set_kernel_attrs(k)) return rc_;                                                     \
        VAE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k, kT, p.smem));                \
      }                                                                                                    \
      break;                                                                                               \
    }
    VAE_SMALL_SHAPES(X)
#undef X
    default: set_error("bad shape id %d", shape_id); return VAE_ERR_INVALID;
  }
  *out = blocks * p.gm;
  return 0;
}

}  // namespace vae

namespace vae {

// models resident per SM as a full homogeneous pack of this shape is launched
int small_ctas_per_sm(const VaeTrainJob* job) {
  int fl = 0, out = 0;
  const int id = small_shape_of(job, &fl);
  if (id < 0 || (size_t)fl * 4 > (size_t)kMaxSmem) return 0;
  const SmallPlan p = plan_small(job, 1 << 20, true);        // a pack large enough to fill the GPU
  const int rc = occupancy_small(id, p, &out);
  return rc ? rc : out;
}

int small_supported(const VaeTrainJob* job) {
  if (!job) return 0;
  int fl = 0;
  const int id = small_shape_of(job, &fl);
  if (id < 0 || (size_t)fl * 4 > (size_t)kMaxSmem) return 0;
  return 1;
}


size_t small_job_bytes() { return sizeof(FusedJobDev); }

// Validates the jobs, uploads the device job table into d_scratch (on `s`) and records how the pack
// will be launched.  Everything per-job happens here, once; small_launch is a bare kernel launch.
int small_prepare(const VaeTrainJob* jobs, int32_t n_jobs, void* d_scratch, cudaStream_t s, VaePackPlan* plan) {
  FusedJobDev* host = (FusedJobDev*)malloc((size_t)n_jobs * sizeof(FusedJobDev));
  VAE_REQUIRE(host, "vae_pack_prepare: out of host memory");
  int max_fl = 0;
  bool homogeneous = true;
  for (int i = 0; i < n_jobs; ++i) {
    const VaeTrainJob& j = jobs[i];
    int fl = 0;
    const int id = small_shape_of(&j, &fl);
    if (id < 0 || (size_t)fl * 4 > (size_t)kMaxSmem) {
      free(host);
      set_error("vae_pack_prepare: job %d has no fused kernel (D=%d L=%d hidden=%d/%d)", i,
                     j.desc.data_dim, j.desc.latent_dim, j.desc.n_enc_hidden, j.desc.n_dec_hidden);
      return VAE_ERR_UNSUPPORTED;
    }
    if (!j.params || !j.adam_m || !j.adam_v || !j.step || !j.model_key || !j.dataset_key || !j.loss_out ||
        (j.data.kind != VAE_DATA_SPHERE && !j.A)) {
      free(host);
      set_error("vae_pack_prepare: job %d has a NULL device pointer", i);
      return VAE_ERR_INVALID;
    }
    if (fl > max_fl) max_fl = fl;
    FusedJobDev& o = host[i];
    o.params = j.params; o.adam_m = j.adam_m; o.adam_v = j.adam_v; o.step = j.step;
    o.model_key = j.model_key; o.dataset_key = j.dataset_key; o.A = j.A; o.loss_out = j.loss_out;
    o.batch = j.desc.batch; o.tunable = j.desc.tunable_decoder_var; o.shape_id = id; o.pad_ = 0;
    const AdamHyper hp = make_adam_hyper(&j.desc);
    o.eps0 = (float)j.desc.epsilon0; o.lr = hp.lr; o.b1 = hp.b1; o.b2 = hp.b2; o.omb1 = hp.omb1; o.omb2 = hp.omb2;
    o.adam_eps = hp.eps;
    if (id != host[0].shape_id || o.batch != host[0].batch) homogeneous = false;
  }
  // pageable-source async copies are staged before the call returns, so `host` may be freed
  cudaError_t e = cudaMemcpyAsync(d_scratch, host, (size_t)n_jobs * sizeof(FusedJobDev),
                                  cudaMemcpyHostToDevice, s);
  const int shape0 = host[0].shape_id;
  free(host);
  VAE_CUDA_OK(e);
  SmallPlan p = plan_small(&jobs[0], n_jobs, homogeneous);
  if (!homogeneous) p.smem = (size_t)max_fl * sizeof(float);
  plan->cls = 1; plan->n_jobs = n_jobs; plan->shape_id = shape0; plan->homogeneous = homogeneous ? 1 : 0;
  plan->gm = p.gm; plan->cluster = p.cluster; plan->nbuf = p.nbuf; plan->smem_bytes = (uint64_t)p.smem;
  plan->d_jobs = d_scratch;
  return launch_small(shape0, homogeneous, p, n_jobs, (const FusedJobDev*)d_scratch, 0, s, /*prepare_only=*/true);
}

int small_launch(const VaePackPlan* plan, int64_t n_steps, cudaStream_t s) {
  SmallPlan p = {plan->gm, plan->cluster, plan->nbuf, (size_t)plan->smem_bytes};
  return launch_small(plan->shape_id, plan->homogeneous != 0, p, plan->n_jobs, (const FusedJobDev*)plan->d_jobs, n_steps, s,
                      /*prepare_only=*/false);
}

}  // namespace vae
