// fused_mlp_tc.cu -- persistent fused training loop for the hidden-layer VAEs of the sphere sweep
// (sphere_vae_padding_expts.sh: enc = dec = 200|200|200) on the 5th-generation tensor cores: ONE CTA per
// model, every Dense layer of the step (forward, dX, dW) is a tcgen05.mma.kind::tf32 GEMM with the whole
// batch as one M = 128 tile and the accumulator in tensor memory.
//
// The hot-loop body is the one fused_mlp.cu implements with FP32 FFMA on a thread-block cluster
// (model.py:221-222: key splits, eps draw, dataset batch, VAE.apply networks.py:61-84, sampled ELBO :94-98,
// hand-derived backward, legacy flax Adam vae.py:113).  Here a model's step is a chain of 23 GEMMs
//     forward   Y[b][n]  = act(X[b][:] . WT[n][:] + bias[n])          A = X  (batch-major),   B = W^T
//     dX        dX[b][k] = (dY[b][:] . W[k][:]) * relu'(X[b][k])      A = dY (batch-major),   B = W
//     dW        dW[k][n] = XT[k][:] . dYT[n][:]   (+ Adam epilogue)    A = X^T (feature-major), B = dY^T
// -- all "NT" (both operands K-contiguous).  tcgen05 reads kind::tf32 operands K-major from shared memory
// (MN-major TF32 operands would need the SWIZZLE_128B_BASE32B layout, see umma_probe.cu), so every
// activation / gradient is kept in two global (L2-resident) copies, batch-major and feature-major [f][b],
// written by the producing GEMM's epilogue straight from tensor memory; weights are kept as W (the Adam master
// copy) and W^T [n][k] (rewritten by the Adam epilogue).
// Global layouts follow the tensor-memory drain pattern (thread = accumulator row, 8 consecutive columns per
// tcgen05.ld): a "row-per-thread" matrix X[r][c] is stored GROUP-MAJOR, [c / 8][r][8], so the 32 lanes of a
// warp write 32 x 32 contiguous bytes (one 256-bit store each) instead of 32 strided sectors; the operand
// stager reads the same arrays as 16-byte pieces.  This holds for the batch-major activations / gradients
// (r = batch row) and for W, m, v (r = k, the dW accumulator row).  Feature-major copies and W^T are plain
// row-major (their stores are lane-contiguous already).
//
// FP32-level accuracy from TF32 tensor cores: every operand is split x = hi + lo with hi = TF32(x) rounded to
// nearest and lo = x - hi (exact in FP32, |lo| <= 2^-11 |x|; the tensor core ignores the low 13 mantissa bits
// of what it is given -- measured, tests/test_gpu_umma_probe.py -- so hi is used exactly and lo to 2^-22 |x|);
// a product is accumulated as
//     A.B ~= Ah.Bh + Ah.Bl + Al.Bh        (three MMAs per K = 8 step; the dropped Al.Bl term is ~2^-22 relative
//                                          per element with either sign)
// Only algorithmic flops (2MNK) are ever reported as useful work.
//
// Pipeline per GEMM: K is walked in chunks of 16 through a shared-memory ring (3-4 stages, cut to the GEMM's
// size).  One thread feeds the ring with TMA, ONE cp.async.bulk.tensor per operand and chunk: row-major operands
// land as 64-byte rows (SWIZZLE_64B), group-major ones as two [row][32-byte] blocks (SWIZZLE_32B) -- the canonical
// K-major layouts tcgen05 reads -- with rows / K beyond the matrix zero-filled.  All 512 threads turn the landed
// chunk into hi = TF32(x) (in place) and lo = x - hi (position-preserving, so the swizzle does not matter); the
// feeding thread then issues the chunk's MMAs and a tcgen05.commit that frees the ring slot.  After the last chunk
// the accumulator (128 lanes x N columns per M tile) is drained with tcgen05.ld by all 16 warps (lane quarter =
// warp % 4, column groups dealt round-robin to the four warps of a quarter) into the fused epilogue: bias + relu,
// relu mask, reparameterisation, loss partials, or Adam.
//
// Latency configuration: a thread-block CLUSTER of C = 2, 4 or 8 CTAs per model.  Every GEMM is cut along its N
// dimension in units of 8 columns (tc_slice): forward Y[:, slice], dX[:, k-slice], dW[:, slice] + Adam on the slice;
// a CTA loads only its rows of the B operand (TMA row coordinate) and the whole A operand, multiplies M128 x Nslice
// and writes its slice of the result to the same global (L2-resident) arrays as the single-CTA kernel; one hardware
// cluster barrier (release / acquire, after a generic->async proxy fence so the peers' TMA reads see the stores)
// per GEMM on the dependent chain: 1 + 8 + 7 = 16 per step.  Layers of at most 64 columns (mu, x_hat, the latent
// gradient) belong to CTA 0 alone, which therefore also owns the loss and the epsilon / epsilon_p updates.  Every
// output element is produced by the same MMA sequence whatever C is: results are bit-identical for C = 1, 2, 4, 8.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vae {

constexpr int kTT = 512;                        // threads per CTA
constexpr int kTW = kTT / 32;                   // warps
constexpr int kTCG = kTW / 4;                   // warps per TMEM lane quarter = column-group classes
constexpr int kTKC = 16;                        // K elements per staged chunk (two K = 8 MMA steps)
constexpr int kTMaxStages = 6;
constexpr int kTMaxN = 208;                     // most B rows (accumulator columns) of one GEMM
constexpr int kTRingBytes = 176 * 1024;         // the operand ring; a GEMM cuts it into stages of its own size
constexpr int kTSmemBytes = kTRingBytes + 1024;
constexpr int kTMaxL = VAE_MAX_HIDDEN + 1;
constexpr int kTMaxCluster = 16;                // most CTAs per model (16 = non-portable cluster size, forced by VAE_MLP_TC_CLUSTER=16 only)
constexpr int kTAutoCluster = 8;                // largest cluster the library picks on its own (portable)
constexpr int kTNarrowGroups = 8;               // layers of at most this many 8-column groups are not cut (CTA 0 owns them)
// tensor maps of a model: per layer [W^T, W (group-major), input batch-major, input feature-major, dY feature-major],
// then the three batch-major gradient buffers
enum { TM_WT = 0, TM_W8 = 1, TM_INBM = 2, TM_INT = 3, TM_GT = 4, TM_PER_LAYER = 5 };
constexpr int kTGradBufs = 3;                   // dY buffers in rotation (see the backward pass)
constexpr int kTMaps = 2 * kTMaxL * TM_PER_LAYER + kTGradBufs;

// Columns [8 g0, 8 g1) of a layer with `groups` 8-column groups that CTA `rank` of a `csize`-CTA cluster owns.
__host__ __device__ inline void tc_slice(int groups, int csize, int rank, int* g0, int* g1) {
  if (csize <= 1 || groups <= kTNarrowGroups) { *g0 = 0; *g1 = rank == 0 ? groups : 0; return; }
  const int base = groups / csize, rem = groups % csize;
  *g0 = rank * base + (rank < rem ? rank : rem);
  *g1 = *g0 + base + (rank < rem ? 1 : 0);
}
// Rows of the B-operand TMA box of such a layer (the widest slice; every CTA multiplies this many columns)
__host__ __device__ inline int tc_slice_box(int groups, int csize) {
  if (csize <= 1 || groups <= kTNarrowGroups) return groups * 8;
  return ((groups + csize - 1) / csize) * 8;
}

__host__ __device__ constexpr int t_up8(int n) { return (n + 7) & ~7; }
__host__ __device__ constexpr int t_up4(int n) { return (n + 3) & ~3; }

// One Dense layer K -> N.  Offsets are in floats from the model's workspace base.
struct TcLayer {
  int32_t K, N, Kp, Np;            // logical sizes, padded to multiples of 8
  int32_t Kuse;                    // input rows that can be non-zero (encoder layer 0: DR, the non-padding columns)
  int32_t flat_w, flat_b;          // offsets in the flat parameter vector (include/vae_b200.h)
  int32_t W, Bv, WT;               // W [K][Np] (Adam master), bias [Np], WT [Np][Kp]; adam m / v at + pm_stride / 2 pm_stride
  int32_t in_bm, in_T;             // layer input: batch-major [BP][Kp], feature-major [Kp + 8][BP] (row Kuse = ones)
  int32_t out_bm, out_T;           // layer output (hidden layers): batch-major [BP][Np], feature-major [Np + 8][BP]
  int32_t in_mask, out_mask;       // relu masks (output > 0) of the input / output activation: [BP][32] bytes, one per 8 columns
  int32_t gcur;                    // which of the three gradient buffers holds this layer's dY in the backward pass
  int32_t nbox, kbox;              // B-operand box rows of the N-sliced (forward, dW) and K-sliced (dX) GEMMs
};

struct TcJobDev {
  float* params; float* adam_m; float* adam_v; int32_t* step;
  uint32_t* model_key; uint32_t* dataset_key; const float* A; float* loss_out; float* ws;
  int32_t B, BP, D, L, DR, kind, G, dd, tunable, n_enc, n_dec;
  int32_t Lp, Dp, LD;
  int32_t off_lv, off_eps;         // flat offsets of epsilon_p, epsilon
  int32_t pm_stride;               // floats between the aligned P, M and V parameter regions
  int32_t ws_z, ws_gn, ws_mu, ws_gbm[kTGradBufs], ws_gT[kTGradBufs];
  int32_t n_flat, csize;
  TcLayer enc[kTMaxL], dec[kTMaxL];
  float eps0, lr, b1, b2, omb1, omb2, adam_eps;
  alignas(64) CUtensorMap maps[kTMaps];   // read by TMA straight from the job table in global memory
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t t_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_smem(bar)), "r"(count));
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void t_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = t_smem(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void t_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t_smem(bar)) : "memory");
}
// One lane of the (converged) warp: 1 in the elected lane, 0 elsewhere.
__device__ __forceinline__ uint32_t t_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
// Marks a value the whole warp holds identically as warp-uniform for the compiler (it then lives in a uniform register
// and feeds UTCHMMA / UTMALDG directly instead of through a per-instruction ELECT + R2UR waterfall).
__device__ __forceinline__ uint32_t t_uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void t_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major shared-memory operand descriptors for the two ways TMA lands a 16-float K chunk (version 1):
//   row-major global operand  -> rows of 64 bytes, SWIZZLE_64B (layout 4): 8-row atoms 512 B apart (SBO); a K = 8
//                                step is 32 bytes further inside the row
//   group-major global operand -> two K groups of [row][32 bytes], SWIZZLE_32B (layout 6): 8-row atoms 256 B apart;
//                                a K = 8 step is the next group block (32 x rows bytes further)
__device__ __forceinline__ uint64_t t_desc(uint32_t smem_addr, int grouped) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((grouped ? 256 : 512) >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(grouped ? 6 : 4) << 61);
}
// c = F32, a = b = TF32, both K-major, N >> 3 at [17,23), M = 128 (>> 4) at [24,29)
__device__ __forceinline__ uint32_t t_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void t_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t_smem(bar)), "r"(bytes) : "memory");
}
// one chunk (16 K floats) of the operand's box rows: row-major operands are a 2-D map {K floats, rows} (box 16 x rows),
// group-major ones a 3-D map {8 floats, rows, K groups} (box 8 x rows x 2)
__device__ __forceinline__ void t_tma_chunk(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int chunk, int grouped, int row0) {
  if (grouped)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(t_smem(bar)), "r"(0), "r"(row0), "r"(2 * chunk) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(t_smem(bar)), "r"(16 * chunk), "r"(row0) : "memory");
}
__device__ __forceinline__ uint32_t t_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t t_cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void t_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
// main + correction accumulator columns (single-tile GEMMs, see tc_gemm)
__device__ __forceinline__ void t_ld8_sum(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8], c[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]) : "r"(taddr + 256u));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(c[j]);
}
__device__ __forceinline__ void t_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
// TF32(x), round to nearest (ties away): the tensor core would TRUNCATE a raw FP32 operand, which biases every
// product the same way; with hi rounded, lo = x - hi (exact in FP32) has either sign
__device__ __forceinline__ float t_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
// 256-bit global accesses (LDG/STG.E.256): one full 32-byte sector per lane
__device__ __forceinline__ void t_ldg8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void t_stg8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ float t_sqrt_approx(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct TcHyper { float lr, b1, b2, omb1, omb2, eps, ibc1, ibc2; };
__device__ __forceinline__ float t_adam(float p, float& m, float& v, float g, const TcHyper& h) {
  m = h.b1 * m + h.omb1 * g;
  v = h.b2 * v + h.omb2 * g * g;
  return p - __fdividef(h.lr * (m * h.ibc1), t_sqrt_approx(v * h.ibc2) + h.eps);
}

#ifdef VAE_TC_PROFILE
__device__ long long g_tc_prof[16];
// per-role lap timers: accumulated in registers inside a GEMM, added to the global counters once per GEMM
#define TC_LAPS() long long _tl = clock64(), _la = 0, _lb = 0
#define TC_LAP_A() do { const long long _n = clock64(); _la += _n - _tl; _tl = _n; } while (0)
#define TC_LAP_B() do { const long long _n = clock64(); _lb += _n - _tl; _tl = _n; } while (0)
#define TC_LAP_FLUSH(i, j) do { if (blockIdx.x == 0) { g_tc_prof[i] += _la; g_tc_prof[j] += _lb; } } while (0)
#define TC_T0() const long long _t0 = clock64()
#define TC_ACC(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_tc_prof[i] += clock64() - _t0; } while (0)
#else
#define TC_T0() do {} while (0)
#define TC_ACC(i) do {} while (0)
#define TC_LAPS() do {} while (0)
#define TC_LAP_A() do {} while (0)
#define TC_LAP_B() do {} while (0)
#define TC_LAP_FLUSH(i, j) do {} while (0)
#endif

// ------------------------------------------------------------------------------------------ GEMM main loop
struct TcRing {
  uint8_t* base;                  // kTRingBytes, 1024-byte aligned
  uint64_t* full;                 // [kTMaxStages] TMA bytes of the slot have landed
  uint64_t* ready;                // [kTMaxStages] the slot's hi / lo split is complete (one arrive per converter warp)
  uint64_t* empty;                // [kTMaxStages] ring slot free (tcgen05.commit of the MMAs that read it)
  uint64_t* accfull;              // accumulator complete
  float* bias_s;                  // [kTMaxN] bias of the running GEMM (forward epilogues)
  uint32_t fpar;                  // bit s: parity of full[s]'s next completion      (converter threads)
  uint32_t rpar;                  // bit s: parity of ready[s]'s next completion     (feeding thread)
  uint32_t par;                   // bit s: parity of empty[s]'s next completion     (feeding thread)
  uint32_t pend;                  // bit s: a commit on empty[s] has not been waited for yet
  uint32_t gemms;                 // GEMMs finished so far
  uint32_t tmem;                  // tensor-memory base (M tile t accumulates at column 256 t)
};

// A GEMM operand: its tensor map (box = box_rows x 16 K floats), whether the global array is group-major, box rows.
struct TcOperand { const CUtensorMap* tm; int grouped; int box_rows; int row0 = 0; };

// D[m][n] = sum_k A[m][k] B[n][k], k < kpad (multiple of 8).  A: m_tiles M tiles of 128 rows out of A.box_rows; B:
// B.box_rows rows, n_mma (multiple of 8) of them multiplied.  bias (may be null): n_bias floats copied to R.bias_s
// while the main loop runs.  The ring is cut into stages of this GEMM's own size, (A.box_rows + B.box_rows) x 16
// floats x {hi, lo}: [A hi][B hi][A lo][B lo]; up to stages - 1 chunks are in flight beyond the one being multiplied.
// Roles: lane 0 of warp 0 feeds the ring (TMA), lane 0 of warp 1 issues the MMAs (a single thread doing both spends
// ~1.4 k cycles of dependent scalar code per chunk, whatever N is -- measured), warps 2..15 split the landed chunks
// into hi / lo and hand them over through the `ready` barriers -- no block-wide barrier inside the loop, so the load of
// chunk c + 2, the split of chunk c + 1 and the MMAs of chunk c overlap.  On return the accumulator is complete and visible to tcgen05.ld in every thread.
__device__ __forceinline__ void tc_gemm(TcRing& R, const TcOperand A, int m_tiles, const TcOperand Bm, int n_mma, int kpad,
                                        const float* __restrict__ bias = nullptr, int n_bias = 0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = (kpad + kTKC - 1) / kTKC;
  const int a_bytes = A.box_rows * kTKC * 4, b_bytes = Bm.box_rows * kTKC * 4, hi_bytes = a_bytes + b_bytes;
  const int stage_bytes = 2 * hi_bytes;
  int nstages = kTRingBytes / stage_bytes;
  nstages = nstages > kTMaxStages ? kTMaxStages : nstages;

  TC_T0();
  if (tid == 0) {
    // ===================== producer thread: one TMA box per operand and chunk, as far ahead as the ring allows
    int stage = 0;
    TC_LAPS();
    for (int c = 0; c < nchunk; ++c) {
      if (R.pend & (1u << stage)) {                // the MMAs that read this slot have retired
        t_mbar_wait(&R.empty[stage], (R.par >> stage) & 1u);
        R.par ^= 1u << stage;
      }
      TC_LAP_A();
      R.pend |= 1u << stage;                       // a commit on empty[stage] will follow this chunk's MMAs
      const uint32_t sA = t_smem(R.base) + (uint32_t)(stage * stage_bytes);
      t_expect_tx(&R.full[stage], (uint32_t)hi_bytes);
      t_tma_chunk(sA, A.tm, &R.full[stage], c, A.grouped, A.row0);
      t_tma_chunk(sA + a_bytes, Bm.tm, &R.full[stage], c, Bm.grouped, Bm.row0);
      stage = stage + 1 == nstages ? 0 : stage + 1;
      TC_LAP_B();
    }
    TC_LAP_FLUSH(8, 9);
  } else if (warp == 1) {
    // ===================== MMA warp: waits for a chunk's hi / lo split, queues its MMAs (one elected lane), frees the
    // slot by commit.  The whole warp runs the loop on warp-uniform values: descriptors are built in uniform registers.
    const uint32_t idesc = t_uni(t_idesc(n_mma));
    // descriptor = constant high part | (shared address >> 4); the ring lies below 256 KB, so the address field never carries
    const uint64_t hA = t_desc(0, A.grouped), hB = t_desc(0, Bm.grouped);
    const uint32_t hiA = t_uni((uint32_t)(hA >> 32)), loA = t_uni((uint32_t)hA), hiB = t_uni((uint32_t)(hB >> 32)), loB = t_uni((uint32_t)hB);
    const uint32_t ka = t_uni((uint32_t)(A.grouped ? A.box_rows * 32 : 32) >> 4);       // one K = 8 step further, in descriptor units:
    const uint32_t kb = t_uni((uint32_t)(Bm.grouped ? Bm.box_rows * 32 : 32) >> 4);     // +32 B in a 64-byte row / the next group block
    const uint32_t ta = t_uni((uint32_t)(128 * (A.grouped ? 32 : 64)) >> 4);            // M tile 1: 128 rows further
    const uint32_t lo_ofs = t_uni((uint32_t)hi_bytes >> 4), b_ofs = t_uni((uint32_t)a_bytes >> 4), st_ofs = t_uni((uint32_t)stage_bytes >> 4);
    const uint32_t ring = t_uni(t_smem(R.base) >> 4), tmem = t_uni(R.tmem);
    const uint32_t un_chunk = t_uni((uint32_t)nchunk), un_stages = t_uni((uint32_t)nstages), u_kpad = t_uni((uint32_t)kpad);
    const uint32_t two_tiles = t_uni((uint32_t)(m_tiles != 1));
    uint32_t rpar = t_uni(R.rpar);
    auto desc = [](uint32_t hi, uint32_t lo) -> uint64_t { return ((uint64_t)hi << 32) | lo; };
    uint32_t stage = 0;
    TC_LAPS();
    for (uint32_t c = 0; c < un_chunk; ++c) {
      t_mbar_wait(&R.ready[stage], (rpar >> stage) & 1u);              // chunk c is split into hi / lo
      rpar ^= 1u << stage;
      TC_LAP_A();
      t_fence_after();
      const uint32_t sA = ring + stage * st_ofs, sB = sA + b_ofs;
      const uint32_t ksteps = min((uint32_t)kTKC, u_kpad - c * kTKC) >> 3;
      if (t_elect()) {
#pragma unroll
        for (uint32_t s = 0; s < kTKC / 8; ++s) {
          if (s < ksteps) {
            const uint64_t dBh = desc(hiB, loB | (sB + s * kb)), dBl = desc(hiB, loB | (sB + lo_ofs + s * kb));
            const uint32_t acc = (uint32_t)((c | s) != 0);
            if (!two_tiles) {
              // The tensor core accumulates FP32 with round-toward-zero.  Single-tile GEMMs (forward, dX: K up to 200)
              // keep the two correction products in their own accumulator (columns 256..) so that the small terms are
              // never truncated against the large main sum and the main chain is a third as long; the epilogue adds
              // the two.  Two-tile GEMMs (dW, K = batch) have no columns to spare and a chain of only 39 MMAs.
              const uint64_t dAh = desc(hiA, loA | (sA + s * ka)), dAl = desc(hiA, loA | (sA + lo_ofs + s * ka));
              t_mma(tmem, dAh, dBh, idesc, acc);
              t_mma(tmem + 256u, dAh, dBl, idesc, acc);
              t_mma(tmem + 256u, dAl, dBh, idesc, 1u);
            } else {
#pragma unroll
              for (uint32_t t = 0; t < 2; ++t) {
                const uint64_t dAh = desc(hiA, loA | (sA + s * ka + t * ta)), dAl = desc(hiA, loA | (sA + lo_ofs + s * ka + t * ta));
                const uint32_t d = tmem + t * 256u;
                t_mma(d, dAh, dBh, idesc, acc);
                t_mma(d, dAh, dBl, idesc, 1u);
                t_mma(d, dAl, dBh, idesc, 1u);
              }
            }
          }
        }
        t_commit(&R.empty[stage]);
        if (c == un_chunk - 1) t_commit(R.accfull);
      }
      __syncwarp();
      TC_LAP_B();
      stage = stage + 1 == un_stages ? 0 : stage + 1;
    }
    R.rpar = rpar;
    if (lane == 0) TC_LAP_FLUSH(10, 11);
  } else if (tid >= 64) {
    // ===================== converter warps: hi = TF32(x) rounded to nearest (in place), lo = x - hi
    // A chunk costs a warp ~600 cycles whatever its size (barrier wake-up, proxy fence, arrive -- measured), so with an
    // even number of ring slots the 14 warps work as two groups on alternate chunks (a slot then always belongs to the
    // same group, which keeps every thread's parity bookkeeping of the slot's `full` barrier exact).
    const int ngrp = (nstages & 1) ? 1 : 2;
    const int grp = ngrp == 2 ? (warp & 1) : 0;
    const int ct = ngrp == 2 ? ((warp - 2) >> 1) * 32 + lane : tid - 64;
    const int nconv = (kTT - 64) / ngrp;
    if (bias != nullptr)
      for (int i = tid - 64; i < n_bias; i += kTT - 64) R.bias_s[i] = bias[i];
    const int n4 = hi_bytes >> 4;
    int stage = grp;
    uint32_t use = 0;                                                  // completions of full[stage] earlier in this GEMM
    TC_LAPS();
    for (int c = grp; c < nchunk; c += ngrp) {
      t_mbar_wait(&R.full[stage], ((R.fpar >> stage) ^ use) & 1u);     // the chunk's two boxes have landed
      TC_LAP_A();
      float4* hi = reinterpret_cast<float4*>(R.base + (size_t)stage * stage_bytes);
      float4* lo = reinterpret_cast<float4*>(R.base + (size_t)stage * stage_bytes + hi_bytes);
#pragma unroll 3
      for (int p = ct; p < n4; p += nconv) {
        const float4 x = hi[p];
        // hi rounded to nearest, not the hardware's truncation of a raw operand: with truncation lo >= 0 always and the
        // dropped Al.Bl term biases every product the same way (measured: loss error 3.9e-6 instead of 2.0e-6)
        const float4 h = make_float4(t_hi(x.x), t_hi(x.y), t_hi(x.z), t_hi(x.w));
        hi[p] = h;
        lo[p] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy smem writes -> tensor-core reads
      __syncwarp();
      if (lane == 0)                                                   // ready[] counts 14 arrivals per chunk
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(t_smem(&R.ready[stage])), "r"(ngrp) : "memory");
      TC_LAP_B();
      stage += ngrp;
      if (stage >= nstages) { stage -= nstages; ++use; }
    }
    if (tid == 64) TC_LAP_FLUSH(12, 13);
  }
  // every thread's view of the `full` barriers' parities: slot s completed once per chunk c = s (mod nstages)
  for (int st = 0; st < nstages && st < nchunk; ++st)
    if (((nchunk - st + nstages - 1) / nstages) & 1) R.fpar ^= 1u << st;
  TC_ACC(0);
  if (tid < 32) __syncwarp();                      // lanes 1..31 of the producer warp park here instead of polling beside lane 0
  t_mbar_wait(R.accfull, R.gemms & 1u);
  TC_ACC(1);
  R.gemms += 1;
  if (tid == 0) { R.par ^= R.pend; R.pend = 0; }   // every outstanding commit has completed with the accumulator
  t_fence_after();
  __syncthreads();                                 // bias_s is visible; every thread is past the main loop
}

// End of an epilogue: TMEM reads and global writes of all threads are done before the next GEMM starts.
__device__ __forceinline__ void tc_epilogue_done() {
  t_fence_before();
  __syncthreads();
}
// The same across the model's cluster: this CTA's global stores are ordered before every peer's later reads -- generic
// (epilogue loads) through the barrier's release / acquire, async-proxy (TMA) through the proxy fence before it.
__device__ __forceinline__ void tc_cluster_sync(int csize) {
  t_fence_before();
  if (csize > 1) {
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

__device__ __forceinline__ float tc_block_sum(float v, float* red) {   // result valid in every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kTT / 32; ++i) t += red[i];
  return t;
}

// Element (k, n) of a layer's group-major W / m / v copy (see the header): [n / 8][k][8]
__device__ __forceinline__ size_t tc_w8(const TcLayer& Ly, int k, int n) { return ((size_t)(n >> 3) * Ly.K + k) * 8 + (n & 7); }

// Adam epilogue of dW: accumulator row = k (M tile t, lane r -> k = 128 t + r), column = n - 8 g0 for the column groups
// [g0, g1) this CTA owns.  Row Kuse is the bias (the ones row of X^T).  Updates W / m / v (group-major: one 256-bit access
// per lane and column group, lanes contiguous) and W^T.  The parameter / moment loads of the next two column groups are
// in flight while a group is updated (the state lives in L2 / HBM).
struct TcPmv { float p[8], m[8], v[8]; };
__device__ __forceinline__ void tc_pmv_load(const float* P, int pm, TcPmv& x) {
  t_ldg8(P, x.p); t_ldg8(P + pm, x.m); t_ldg8(P + 2 * pm, x.v);
}
__device__ __forceinline__ void tc_adam(const TcRing& R, float* __restrict__ ws, const TcLayer& Ly, int pm, int m_tiles,
                                        int g0, int g1, const TcHyper& hp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, cg = warp >> 2;
  for (int t = 0; t < m_tiles; ++t) {
    if (t * 128 + q * 32 > Ly.Kuse) continue;                        // no row of this warp's lane quarter is a parameter
    const int k = t * 128 + q * 32 + lane;
    const bool is_w = k < Ly.Kuse, is_b = k == Ly.Kuse, live = is_w || is_b;
    // address of column group g: kernel row k -> W + (g K + k) 8 ; bias -> Bv + 8 g
    const float* base = is_b ? ws + Ly.Bv : ws + Ly.W + (size_t)(live ? k : 0) * 8;
    const size_t gstride = is_b ? 8 : (size_t)Ly.K * 8;
    TcPmv cur, nxt;
    if (live && g0 + cg < g1) tc_pmv_load(base + (g0 + cg) * gstride, pm, cur);
    for (int g = g0 + cg; g < g1; g += kTCG) {
      if (live && g + kTCG < g1) tc_pmv_load(base + (g + kTCG) * gstride, pm, nxt);      // in flight during the update of g
      float acc[8];
      if (m_tiles == 1) t_ld8_sum(R.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - g0) * 8), acc);
      else t_ld8(R.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + (g - g0) * 8), acc);
      if (live) {
        const int n0 = g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + j < Ly.N) cur.p[j] = t_adam(cur.p[j], cur.m[j], cur.v[j], acc[j], hp);
        float* P = const_cast<float*>(base) + g * gstride;
        t_stg8(P, cur.p); t_stg8(P + pm, cur.m); t_stg8(P + 2 * pm, cur.v);
        if (is_w) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + j < Ly.N) ws[Ly.WT + (size_t)(n0 + j) * Ly.Kp + k] = cur.p[j];
        }
      }
      cur = nxt;
    }
  }
}

// dW of one layer (+ Adam) on this CTA's column groups [g0, g1): D[k][n] = sum_b X^T[k][b] dY^T[n][b], rows k incl. the
// ones row (bias), one or two M tiles.
__device__ __forceinline__ void tc_dw(TcRing& R, float* __restrict__ ws, const TcLayer& Ly, const CUtensorMap* lm, int BP, int pm,
                                      int g0, int g1, const TcHyper& hp) {
  if (g0 >= g1) return;
  const int xrows = t_up8(Ly.Kuse + 1);            // X^T rows incl. the ones row (box rows of the TM_INT map)
  const int tiles = (Ly.Kuse + 1 + 127) >> 7;
  tc_gemm(R, TcOperand{lm + TM_INT, 0, xrows}, tiles, TcOperand{lm + TM_GT, 0, Ly.nbox, g0 * 8}, Ly.nbox, BP);
  TC_T0();
  tc_adam(R, ws, Ly, pm, tiles, g0, g1, hp);
  TC_ACC(2);
  tc_epilogue_done();
}

__global__ void __launch_bounds__(kTT, 1)
fused_mlp_tc_kernel(const TcJobDev* __restrict__ jobs, int64_t n_steps) {
  extern __shared__ uint8_t tc_raw[];
  __shared__ TcJobDev J;                                            // (the tensor maps at its end are not copied)
  __shared__ __align__(8) uint64_t bars[3 * kTMaxStages + 1];
  __shared__ float bias_s[kTMaxN];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[kTW];
  __shared__ float lvred[kTW][64];
  __shared__ uint32_t keys[8];
  __shared__ float sl_s[64], scal_s[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = warp & 3, cg = warp >> 2;
  const int csize = (int)t_cluster_size(), rank = (int)t_cluster_rank();   // CTAs of this model's cluster, this CTA's place
  const int jidx = (int)blockIdx.x / csize;
  const CUtensorMap* gmaps = jobs[jidx].maps;                       // TMA reads the maps from the job table itself
  {
    const int* src = reinterpret_cast<const int*>(jobs + jidx);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < (int)(offsetof(TcJobDev, maps) / 4); i += kTT) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int i = 0; i <= 3 * kTMaxStages; ++i)
      t_mbar_init(&bars[i], (i >= kTMaxStages && i < 2 * kTMaxStages) ? kTW - 2 : 1);       // ready[]: one arrive per converter warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(t_smem(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  t_fence_before();
  __syncthreads();
  t_fence_after();
  TcRing R;
  R.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~(uintptr_t)1023);
  R.full = &bars[0]; R.ready = &bars[kTMaxStages]; R.empty = &bars[2 * kTMaxStages]; R.accfull = &bars[3 * kTMaxStages];
  R.bias_s = bias_s;
  R.fpar = 0; R.rpar = 0; R.par = 0; R.pend = 0; R.gemms = 0; R.tmem = tmem_base_s;

  const int B = J.B, BP = J.BP, D = J.D, L = J.L, DR = J.DR, G = J.G, LD = J.LD, Lp = J.Lp;
  float* __restrict__ ws = J.ws;
  const int pm = J.pm_stride;
  const int ne = J.n_enc, nd = J.n_dec;

  // ---- flat (C ABI) parameter / moment vectors -> aligned copies W [K][Np], bias [Np] (+ m, v) and W^T [Np][Kp]
  for (int net = 0; net < 2; ++net)
    for (int i = 0; i < (net ? nd : ne); ++i) {
      const TcLayer& Ly = net ? J.dec[i] : J.enc[i];
      for (int idx = rank * kTT + tid; idx < Ly.K * Ly.N; idx += csize * kTT) {
        const int k = idx / Ly.N, n = idx - k * Ly.N;
        const float p = J.params[Ly.flat_w + idx];
        const size_t o = Ly.W + tc_w8(Ly, k, n);
        ws[o] = p;
        ws[o + pm] = J.adam_m[Ly.flat_w + idx];
        ws[o + 2 * pm] = J.adam_v[Ly.flat_w + idx];
        if (k < Ly.Kuse) ws[Ly.WT + (size_t)n * Ly.Kp + k] = p;
      }
      for (int n = rank * kTT + tid; n < Ly.N; n += csize * kTT) {
        ws[Ly.Bv + n] = J.params[Ly.flat_b + n];
        ws[Ly.Bv + pm + n] = J.adam_m[Ly.flat_b + n];
        ws[Ly.Bv + 2 * pm + n] = J.adam_v[Ly.flat_b + n];
      }
      // the ones row of the layer input's feature-major copy: row Kuse (bias gradient = ones . dY)
      if (rank == 0)
        for (int b = tid; b < BP; b += kTT) ws[Ly.in_T + (size_t)Ly.Kuse * BP + b] = b < B ? 1.0f : 0.0f;
    }
  // Encoder layer 0 rows d >= DR multiply exactly-zero padding columns (gradient exactly 0): untouched unless the
  // caller supplied NON-zero Adam moments there, in which case they take the g = 0 Adam update every step.
  const TcLayer& E0 = J.enc[0];
  int nz = 0;
  for (int idx = tid; idx < (E0.K - DR) * E0.N; idx += kTT) {
    const int gi = E0.flat_w + DR * E0.N + idx;
    nz |= (J.adam_m[gi] != 0.f || J.adam_v[gi] != 0.f) ? 1 : 0;
  }
  const bool decay_skipped = __syncthreads_or(nz) != 0;

  uint32_t ck0 = 0, ck1 = 0;           // lanes 0,1 model-side chain, lanes 2,3 dataset-side chain
  if (tid < 4) { const uint32_t* src = tid < 2 ? J.model_key : J.dataset_key; ck0 = src[0]; ck1 = src[1]; }
  const int32_t step0 = J.step[0];
  double b1t = pow((double)J.b1, (double)step0), b2t = pow((double)J.b2, (double)step0);
  TcHyper hp; hp.lr = J.lr; hp.b1 = J.b1; hp.b2 = J.b2; hp.omb1 = J.omb1; hp.omb2 = J.omb2; hp.eps = J.adam_eps;
  const float invB = 1.0f / (float)B, fB = (float)B;
  const int row = q * 32 + lane;                 // accumulator row (TMEM lane) this thread drains
  const uint32_t trow = R.tmem + ((uint32_t)(q * 32) << 16);
  const int mask_ofs = rank * BP * 8;              // this CTA's private copy of the relu masks (floats from Ly.*_mask)
  __syncthreads();

  for (int64_t s = 0; s < n_steps; ++s) {
#ifdef VAE_TC_PROFILE
    const long long _ts = clock64();
#endif
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
    if (tid < L) sl_s[tid] = expf(0.5f * J.params[J.off_lv + tid]);
    if (tid == 0) {
      const float e = J.tunable ? J.params[J.off_eps] * J.eps0 : J.eps0;
      scal_s[0] = e; scal_s[1] = expf(0.5f * e); scal_s[2] = expf(-e);
    }
    __syncthreads();
    // ================= random draws: eps (B, L + D) row-major (model.py:227) and the generator normals (B, G)
    {
      const uint32_t lk0 = keys[0], lk1 = keys[1], bk0 = keys[2], bk1 = keys[3];
      const int n_e = B * LD, half_e = (n_e + 1) >> 1, n_d = B * G, half_d = (n_d + 1) >> 1;
      float* z = ws + J.ws_z; float* gn = ws + J.ws_gn;
      // eps is consumed by CTA 0 only, much later: every CTA of the cluster draws a share of it.  The generator normals
      // (few) and the batch are produced by EVERY CTA for itself -- identical values into the same global arrays -- so
      // the first GEMM of the step waits for nobody.
      for (int w = rank * kTT + tid; w < half_e; w += csize * kTT) {
        const int j = w, j2 = w + half_e;
        uint32_t x0 = (uint32_t)j, x1 = j2 < n_e ? (uint32_t)j2 : 0u;
        threefry2x32(lk0, lk1, x0, x1);
        z[j] = normal_from_bits(x0);
        if (j2 < n_e) z[j2] = normal_from_bits(x1);
      }
      for (int w = tid; w < half_d; w += kTT) {
        const int j = w, j2 = j + half_d;
        uint32_t x0 = (uint32_t)j, x1 = j2 < n_d ? (uint32_t)j2 : 0u;
        threefry2x32(bk0, bk1, x0, x1);
        gn[j] = normal_from_bits(x0);
        if (j2 < n_d) gn[j2] = normal_from_bits(x1);
      }
    }
    // Step boundary of the cluster: every CTA has finished the previous step (its dW GEMMs read activations and
    // gradients this step overwrites; its Adam wrote weights this step reads), and eps is complete.
    tc_cluster_sync(csize);
    // ================= generator transform (bit-exact op order, oracle/datasets.py): x -> batch-major + feature-major
    for (int b = tid; b < B; b += kTT) {
      const float* gr = ws + J.ws_gn + (size_t)b * G;
      auto xb = [&](int j) -> float& { return ws[E0.in_bm + ((size_t)(j >> 3) * BP + b) * 8 + (j & 7)]; };
      float* xT = ws + E0.in_T + b;
      if (J.kind == VAE_DATA_SPHERE) {
        float acc = 0.f;
        for (int j = 0; j < J.dd; ++j) { const float v = gr[j]; const float sq = __fmul_rn(v, v); acc = j == 0 ? sq : __fadd_rn(acc, sq); }
        const float rad = __fsqrt_rn(acc);
        for (int j = 0; j < J.dd; ++j) { const float v = __fdiv_rn(gr[j], rad); xb(j) = v; xT[(size_t)j * BP] = v; }
      } else if (J.kind == VAE_DATA_LINEAR_GAUSSIAN) {
        for (int i = 0; i < J.dd; ++i) {
          float acc = __fmul_rn(J.A[i * G], gr[0]);
          for (int k = 1; k < G; ++k) acc = __fadd_rn(acc, __fmul_rn(J.A[i * G + k], gr[k]));
          xb(i) = acc; xT[(size_t)i * BP] = acc;
        }
      } else {
        float acc = 0.f;
        for (int j = 0; j < J.dd; ++j) {
          const float v = gr[j]; xb(j) = v; xT[(size_t)j * BP] = v;
          const float pr = __fmul_rn(v, J.A[j]); acc = j == 0 ? pr : __fadd_rn(acc, pr);
        }
        const float sg = det_sigmoid(acc);
        xb(J.dd) = sg; xT[(size_t)J.dd * BP] = sg;
      }
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");      // the batch is read by TMA next
    __syncthreads();
#ifdef VAE_TC_PROFILE
    if (tid == 0 && blockIdx.x == 0) g_tc_prof[6] += clock64() - _ts;
#endif

    // ================= forward: encoder then decoder
    float kl_part = 0.f, r2 = 0.f, rz = 0.f;
    for (int net = 0; net < 2; ++net) {
      const int nl = net ? nd : ne;
      for (int i = 0; i < nl; ++i) {
        const TcLayer& Ly = net ? J.dec[i] : J.enc[i];
        const int kpad = i == 0 ? t_up8(Ly.Kuse) : Ly.Kp;
        const CUtensorMap* lm = gmaps + (net * kTMaxL + i) * TM_PER_LAYER;
        int g0, g1;                                      // this CTA's column groups of the layer
        tc_slice(Ly.Np >> 3, csize, rank, &g0, &g1);
        if (g0 < g1)
          tc_gemm(R, TcOperand{lm + TM_INBM, 1, 128}, 1, TcOperand{lm + TM_WT, 0, Ly.nbox, g0 * 8}, Ly.nbox, kpad, ws + Ly.Bv, Ly.N);
        TC_T0();
        const float* bias = R.bias_s;
        const bool last = i == nl - 1;
        uint64_t mbits = 0;                              // relu'(output) bytes of this thread's column groups (hidden layers)
        for (int g = g0 + cg; g < g1; g += kTCG) {
          float acc[8];
          t_ld8_sum(trow + (uint32_t)((g - g0) * 8), acc);
          if (row >= BP) continue;
          const int n0 = g * 8;
          const bool vrow = row < B;
          const size_t gofs = ((size_t)g * BP + row) * 8;           // group-major batch-major element (row, 8 g)
          if (!last) {                                   // hidden layer: relu(x W + b) (networks.py:33-39)
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = (vrow && n0 + j < Ly.N) ? fmaxf(acc[j] + bias[n0 + j], 0.f) : 0.f;
            t_stg8(ws + Ly.out_bm + gofs, y);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (n0 + j < Ly.N) ws[Ly.out_T + (size_t)(n0 + j) * BP + row] = y[j];
            uint64_t bits = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) bits |= (uint64_t)(y[j] > 0.f ? 1u : 0u) << j;
            mbits |= bits << (((g - g0) / kTCG) * 8);
          } else if (net == 0) {                         // mu; zs = mu + exp(lv/2) z1 (networks.py:67-74)
            const TcLayer& D0 = J.dec[0];
            float mu[8], zs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const bool ok = vrow && n0 + j < L;
              mu[j] = ok ? acc[j] + bias[n0 + j] : 0.f;
              kl_part = fmaf(mu[j], mu[j], kl_part);
              zs[j] = ok ? fmaf(sl_s[n0 + j], ws[J.ws_z + (size_t)row * LD + n0 + j], mu[j]) : 0.f;
            }
            t_stg8(ws + J.ws_mu + gofs, mu);
            t_stg8(ws + D0.in_bm + gofs, zs);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (n0 + j < L) ws[D0.in_T + (size_t)(n0 + j) * BP + row] = zs[j];
          } else {                                       // x_hat, residual, d x_hat (networks.py:81-83, 95-98)
            const float sd = scal_s[1], gscale = scal_s[2] * invB;
            float gq[8], xv[8];
            t_ldg8(ws + E0.in_bm + gofs, xv);            // x has the same group-major layout (columns >= DR are zero)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int d = n0 + j;
              float gv = 0.f;
              if (vrow && d < D) {
                const float z2 = ws[J.ws_z + (size_t)row * LD + L + d];
                const float xh = fmaf(z2, sd, acc[j] + bias[d]);
                const float r = xh - (d < DR ? xv[j] : 0.f);
                r2 = fmaf(r, r, r2); rz = fmaf(r, z2, rz);
                gv = r * gscale;
              }
              gq[j] = gv;
            }
            t_stg8(ws + J.ws_gbm[0] + gofs, gq);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (n0 + j < D) ws[J.ws_gT[0] + (size_t)(n0 + j) * BP + row] = gq[j];
          }
        }
        if (!last && row < BP)                           // [BP][kTCG][8 bytes]: read back by the same (row, cg) thread
          reinterpret_cast<uint2*>(ws + Ly.out_mask + mask_ofs)[row * kTCG + cg] = make_uint2((uint32_t)mbits, (uint32_t)(mbits >> 32));
        tc_cluster_sync(csize);
        TC_ACC(4);
      }
    }
    // ================= loss (networks.py:94-98) and the decoder log-variance parameter
    if (rank == 0) {                                     // the narrow mu / x_hat layers are CTA 0's: so are all loss partials
      const float s_r2 = tc_block_sum(r2, red), s_rz = tc_block_sum(rz, red), s_mu2 = tc_block_sum(kl_part, red);
      if (tid == 0) {
        const float e = scal_s[0], sd = scal_s[1], inv_var = scal_s[2];
        float klc = 0.f;
        for (int l = 0; l < L; ++l) { const float lv = J.params[J.off_lv + l]; klc += 1.f + lv - expf(lv); }
        const float dkl = (-0.5f * klc * fB + 0.5f * s_mu2) * invB;
        const float mse = (0.5f * inv_var * s_r2 + 0.5f * (1.8378770664093453f + e) * fB * (float)D) * invB;
        J.loss_out[s] = dkl + mse;
        if (J.tunable) {
          const float g = J.eps0 * invB * (-0.5f * inv_var * s_r2 + 0.5f * fB * (float)D + 0.5f * sd * inv_var * s_rz);
          float m = J.adam_m[J.off_eps], v = J.adam_v[J.off_eps];
          J.params[J.off_eps] = t_adam(J.params[J.off_eps], m, v, g, hp);
          J.adam_m[J.off_eps] = m; J.adam_v[J.off_eps] = v;
        }
      }
    }
    // ================= backward: decoder (dY of layer i in gbm[cur] / gT[cur]), latent, encoder
    int cur = 0;
    for (int net = 1; net >= 0; --net) {
      const int nl = net ? nd : ne;
      for (int i = nl - 1; i >= 0; --i) {
        const TcLayer& Ly = net ? J.dec[i] : J.enc[i];
        const bool need_dx = !(net == 0 && i == 0);
        if (need_dx) {
          // dX[b][k] = sum_n dY[b][n] W[k][n] with the OLD weights (Adam of this layer runs after)
          const bool latent = net == 1 && i == 0;       // dzs: -> d lv partials, dmu = dzs + mu / B
          uint64_t mbits = 0;                            // relu'(layer input) bytes of this thread's (row, cg) column groups
          if (!latent && row < BP) {
            const uint2 mk = reinterpret_cast<const uint2*>(ws + Ly.in_mask + mask_ofs)[row * kTCG + cg];
            mbits = (uint64_t)mk.x | ((uint64_t)mk.y << 32);
          }
          const CUtensorMap* lm = gmaps + (net * kTMaxL + i) * TM_PER_LAYER;
          int g0, g1;                                    // this CTA's groups of the layer's INPUT features (the dX columns)
          tc_slice(Ly.Kp >> 3, csize, rank, &g0, &g1);
          if (g0 < g1)
            tc_gemm(R, TcOperand{gmaps + 2 * kTMaxL * TM_PER_LAYER + cur, 1, 128}, 1, TcOperand{lm + TM_W8, 1, Ly.kbox, g0 * 8}, Ly.kbox, Ly.Np);
          TC_T0();
          const int nxt = cur + 1 == kTGradBufs ? 0 : cur + 1;
          float* obm = ws + J.ws_gbm[nxt];
          float* oT = ws + J.ws_gT[nxt];
          for (int g = g0 + cg; g < g1; g += kTCG) {
            float acc[8];
            t_ld8_sum(trow + (uint32_t)((g - g0) * 8), acc);
            const int k0 = g * 8;
            const bool vrow = row < B;
            const size_t gofs = ((size_t)g * BP + (row < BP ? row : 0)) * 8;
            float o[8];
            if (!latent) {
              const uint32_t bits = (uint32_t)(mbits >> (((g - g0) / kTCG) * 8)) & 0xFFu;
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = (vrow && ((bits >> j) & 1u)) ? acc[j] : 0.f;
            } else {
              float pr[8], muv[8];
              t_ldg8(ws + J.ws_mu + gofs, muv);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const bool ok = vrow && k0 + j < L;
                const float dz = ok ? acc[j] : 0.f;
                pr[j] = ok ? dz * ws[J.ws_z + (size_t)row * LD + k0 + j] : 0.f;
                o[j] = ok ? fmaf(muv[j], invB, dz) : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float sum = warp_sum(pr[j]);
                if (lane == 0) lvred[warp][k0 + j] = sum;
              }
            }
            if (row < BP) {
              t_stg8(obm + gofs, o);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (k0 + j < Ly.K) oT[(size_t)(k0 + j) * BP + row] = o[j];
            }
          }
          tc_cluster_sync(csize);
          TC_ACC(5);
          if (latent && rank == 0 && tid < L) {          // g_lv[l] = sum_b dzs z1 * 0.5 s_l - 0.5 (1 - e^lv_l)
            const int gsel = (tid >> 3) % kTCG;   // column group g was drained by the four warps with cg == g % kTCG
            float a = 0.f;
            for (int w = 0; w < 4; ++w) a += lvred[4 * gsel + w][tid];
            const float lv = J.params[J.off_lv + tid];
            float m = J.adam_m[J.off_lv + tid], v = J.adam_v[J.off_lv + tid];
            J.params[J.off_lv + tid] = t_adam(lv, m, v, a * 0.5f * sl_s[tid] - 0.5f * (1.f - expf(lv)), hp);
            J.adam_m[J.off_lv + tid] = m; J.adam_v[J.off_lv + tid] = v;
          }
        }
        {
          int g0, g1;
          tc_slice(Ly.Np >> 3, csize, rank, &g0, &g1);
          tc_dw(R, ws, Ly, gmaps + (net * kTMaxL + i) * TM_PER_LAYER, BP, pm, g0, g1, hp);
        }
        if (need_dx) cur = cur + 1 == kTGradBufs ? 0 : cur + 1;
      }
    }
    if (decay_skipped) {
      for (int idx = rank * kTT + tid; idx < (E0.K - DR) * E0.N; idx += csize * kTT) {
        float* P = ws + E0.W + tc_w8(E0, DR + idx / E0.N, idx % E0.N);
        float m = P[pm], v = P[2 * pm];
        P[0] = t_adam(P[0], m, v, 0.f, hp); P[pm] = m; P[2 * pm] = v;
      }
    }
    __syncthreads();
#ifdef VAE_TC_PROFILE
    if (tid == 0 && blockIdx.x == 0) g_tc_prof[7] += clock64() - _ts;
#endif
  }

  // ---- aligned copies -> flat (C ABI) vectors (every CTA of the cluster takes a share, once all Adam updates are in)
  tc_cluster_sync(csize);
  for (int net = 0; net < 2; ++net)
    for (int i = 0; i < (net ? nd : ne); ++i) {
      const TcLayer& Ly = net ? J.dec[i] : J.enc[i];
      for (int idx = rank * kTT + tid; idx < Ly.K * Ly.N; idx += csize * kTT) {
        const int k = idx / Ly.N, n = idx - k * Ly.N;
        const size_t o = Ly.W + tc_w8(Ly, k, n);
        J.params[Ly.flat_w + idx] = ws[o];
        J.adam_m[Ly.flat_w + idx] = ws[o + pm];
        J.adam_v[Ly.flat_w + idx] = ws[o + 2 * pm];
      }
      for (int n = rank * kTT + tid; n < Ly.N; n += csize * kTT) {
        J.params[Ly.flat_b + n] = ws[Ly.Bv + n];
        J.adam_m[Ly.flat_b + n] = ws[Ly.Bv + pm + n];
        J.adam_v[Ly.flat_b + n] = ws[Ly.Bv + 2 * pm + n];
      }
    }
  if (rank == 0 && tid == 0) J.step[0] = step0 + (int32_t)n_steps;
  if (rank == 0 && tid < 4 && (tid & 1) == 0) { uint32_t* dst = tid < 2 ? J.model_key : J.dataset_key; dst[0] = ck0; dst[1] = ck1; }
  t_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(R.tmem));
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TcEncodeFn tc_encode_fn() {
  static TcEncodeFn fn = nullptr;      // idempotent lookup of a driver entry point
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TcEncodeFn>(p);
  }
  return fn;
}
// Operand maps (out-of-range rows / K are zero-filled by TMA):
//   row-major [rows][ld]:        2-D {kfloats, rows}, row stride ld * 4 bytes;  box {16, box_rows}, SWIZZLE_64B
//   group-major [k / 8][ld][8]:  3-D {8, rows, kfloats / 8}, strides {32, ld * 32} bytes;  box {8, box_rows, 2}, SWIZZLE_32B
static int tc_make_map(CUtensorMap* tm, const float* ptr, int grouped, int rows, int kfloats, int ld, int box_rows) {
  TcEncodeFn enc = tc_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VAE_ERR_CUDA; }
  CUresult r;
  cuuint32_t estr[3] = {1, 1, 1};
  if (!grouped) {
    cuuint64_t gdim[2] = {(cuuint64_t)kfloats, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t gdim[3] = {8, (cuuint64_t)rows, (cuuint64_t)(kfloats / 8)};
    cuuint64_t gstr[2] = {32, (cuuint64_t)ld * 32};
    cuuint32_t box[3] = {8, (cuuint32_t)box_rows, 2};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d; grouped %d rows %d k %d ld %d box %d)", (int)r, grouped, rows, kfloats, ld, box_rows);
    return VAE_ERR_CUDA;
  }
  return 0;
}

// Workspace layout (floats).  Every offset is a multiple of 8 floats; every row stride a multiple of 8.  The caller
// zero-fills the workspace once (padding rows / columns are never written).
static size_t tc_fill(const VaeTrainJob& j, TcJobDev* o, int csize) {
  Layout Lo;
  if (make_layout(&j.desc, &Lo)) return 0;
  const VaeDesc& d = j.desc;
  const VaeDataDesc& g = j.data;
  const int B = d.batch, BP = t_up8(B), L = d.latent_dim, D = d.data_dim;
  const int DR = g.dim + (g.kind == VAE_DATA_SIGMOID ? 1 : 0);
  const int G = g.kind == VAE_DATA_LINEAR_GAUSSIAN ? g.intrinsic_dim : g.dim;
  TcJobDev J;
  memset(&J, 0, sizeof(J));
  size_t off = 0;
  auto take = [&](size_t n) { const size_t r = off; off += (n + 7) & ~(size_t)7; return (int32_t)r; };
  J.B = B; J.BP = BP; J.D = D; J.L = L; J.DR = DR; J.kind = g.kind; J.G = G; J.dd = g.dim;
  J.tunable = d.tunable_decoder_var; J.n_enc = Lo.n_enc; J.n_dec = Lo.n_dec;
  J.Lp = t_up8(L); J.Dp = t_up8(D); J.LD = L + D;
  J.off_lv = (int32_t)Lo.off_eps_p; J.off_eps = (int32_t)Lo.off_eps; J.n_flat = (int32_t)Lo.n_params; J.csize = csize;
  int maxw = 8;
  // aligned parameter region P; M and V repeat it pm_stride floats further
  for (int net = 0; net < 2; ++net)
    for (int i = 0; i < (net ? Lo.n_dec : Lo.n_enc); ++i) {
      TcLayer& Ly = net ? J.dec[i] : J.enc[i];
      const int* sz = net ? Lo.dec_sizes : Lo.enc_sizes;
      Ly.K = sz[i]; Ly.N = sz[i + 1]; Ly.Kp = t_up8(Ly.K); Ly.Np = t_up8(Ly.N);
      Ly.Kuse = (net == 0 && i == 0) ? DR : Ly.K;
      Ly.nbox = tc_slice_box(Ly.Np >> 3, csize); Ly.kbox = tc_slice_box(Ly.Kp >> 3, csize);
      Ly.flat_w = (int32_t)(net ? Lo.dec_w[i] : Lo.enc_w[i]); Ly.flat_b = (int32_t)(net ? Lo.dec_b[i] : Lo.enc_b[i]);
      Ly.W = take((size_t)Ly.K * Ly.Np); Ly.Bv = take(Ly.Np);
      if (Ly.Kp > maxw) maxw = Ly.Kp;
      if (Ly.Np > maxw) maxw = Ly.Np;
    }
  J.pm_stride = (int32_t)off;
  off *= 3;
  for (int net = 0; net < 2; ++net)
    for (int i = 0; i < (net ? Lo.n_dec : Lo.n_enc); ++i) {
      TcLayer& Ly = net ? J.dec[i] : J.enc[i];
      Ly.WT = take((size_t)Ly.Np * Ly.Kp);
    }
  // activations: layer i's input (batch-major [BP][Kp], feature-major [Kp + 8][BP]); outputs alias the next input
  for (int net = 0; net < 2; ++net) {
    const int nl = net ? Lo.n_dec : Lo.n_enc;
    for (int i = 0; i < nl; ++i) {
      TcLayer& Ly = net ? J.dec[i] : J.enc[i];
      Ly.in_bm = take((size_t)BP * Ly.Kp); Ly.in_T = take((size_t)(Ly.Kp + 8) * BP);
    }
    for (int i = 0; i + 1 < nl; ++i) {
      TcLayer& Ly = net ? J.dec[i] : J.enc[i];
      TcLayer& Nx = net ? J.dec[i + 1] : J.enc[i + 1];
      Ly.out_bm = Nx.in_bm; Ly.out_T = Nx.in_T;
      Ly.out_mask = take((size_t)BP * 8 * kTMaxCluster);   // [cluster rank][BP][32] bytes
      Nx.in_mask = Ly.out_mask;
    }
  }
  J.ws_z = take((size_t)BP * (L + D) + 8); J.ws_gn = take((size_t)BP * G + 8); J.ws_mu = take((size_t)BP * J.Lp);
  for (int b = 0; b < kTGradBufs; ++b) { J.ws_gbm[b] = take((size_t)BP * maxw); J.ws_gT[b] = take((size_t)(maxw + 8) * BP); }
  // which gradient buffer holds each layer's dY: the backward pass walks dec[nd-1] .. dec[0], enc[ne-1] .. enc[0]
  // and moves on after every layer that produces a dX.  Three buffers in rotation: in a cluster, layer i - 1's dX
  // (written two buffers on) must not land on the dY a slower peer's dW of layer i is still reading.
  {
    int cur = 0;
    for (int net = 1; net >= 0; --net)
      for (int i = (net ? Lo.n_dec : Lo.n_enc) - 1; i >= 0; --i) {
        TcLayer& Ly = net ? J.dec[i] : J.enc[i];
        Ly.gcur = cur;
        if (!(net == 0 && i == 0)) cur = (cur + 1) % kTGradBufs;
      }
  }
  if (o) {
    const float* wsb = (const float*)j.workspace;
    for (int net = 0; net < 2; ++net)
      for (int i = 0; i < (net ? Lo.n_dec : Lo.n_enc); ++i) {
        const TcLayer& Ly = net ? J.dec[i] : J.enc[i];
        CUtensorMap* lm = J.maps + (net * kTMaxL + i) * TM_PER_LAYER;
        const int xrows = t_up8(Ly.Kuse + 1);
        if (tc_make_map(lm + TM_WT, wsb + Ly.WT, 0, Ly.Np, Ly.Kp, Ly.Kp, Ly.nbox) ||               // B of the forward GEMM (row slice)
            tc_make_map(lm + TM_W8, wsb + Ly.W, 1, Ly.K, Ly.Np, Ly.K, Ly.kbox) ||                 // B of dX: rows k (slice), K = n
            tc_make_map(lm + TM_INBM, wsb + Ly.in_bm, 1, BP, Ly.Kp, BP, 128) ||                   // A of the forward GEMM
            tc_make_map(lm + TM_INT, wsb + Ly.in_T, 0, xrows, BP, BP, xrows) ||                   // X^T (+ ones row) of dW
            tc_make_map(lm + TM_GT, wsb + J.ws_gT[Ly.gcur], 0, Ly.Np, BP, BP, Ly.nbox))           // dY^T of dW (row slice)
          return 0;
      }
    for (int b = 0; b < kTGradBufs; ++b)
      if (tc_make_map(J.maps + 2 * kTMaxL * TM_PER_LAYER + b, wsb + J.ws_gbm[b], 1, BP, maxw, BP, 128)) return 0;   // A of dX
    *o = J;
    o->params = j.params; o->adam_m = j.adam_m; o->adam_v = j.adam_v; o->step = j.step;
    o->model_key = j.model_key; o->dataset_key = j.dataset_key; o->A = j.A; o->loss_out = j.loss_out;
    o->ws = (float*)j.workspace;
    const AdamHyper hp = make_adam_hyper(&j.desc);
    o->eps0 = (float)j.desc.epsilon0; o->lr = hp.lr; o->b1 = hp.b1; o->b2 = hp.b2; o->omb1 = hp.omb1; o->omb2 = hp.omb2;
    o->adam_eps = hp.eps;
  }
  return off;
}

// 1 if the tensor-core kernel can run the job (then *ws_bytes = its workspace need)
int mlp_tc_supported(const VaeTrainJob* j, size_t* ws_bytes) {
  const VaeDesc& d = j->desc;
  const VaeDataDesc& g = j->data;
  if (d.n_enc_hidden + d.n_dec_hidden == 0 || d.sig_decoder || g.kind == VAE_DATA_SIGMOID) return 0;
  Layout Lo;
  if (make_layout(&d, &Lo)) return 0;
  const int width = g.dim + g.padding_dim;
  if (width != d.data_dim || d.latent_dim > 63 || d.batch > 128 || d.batch < 1) return 0;
  if (Lo.max_width > kTMaxN - 8 || Lo.n_params >= (1ll << 30)) return 0;      // N <= 208 incl. the bias column / row
  if (ws_bytes) *ws_bytes = tc_fill(*j, nullptr, 1) * sizeof(float);      // the layout does not depend on the cluster size
  return 1;
}

size_t mlp_tc_job_bytes() { return sizeof(TcJobDev); }

int mlp_tc_fill_job(const VaeTrainJob& j, void* out, int csize) {
  return tc_fill(j, (TcJobDev*)out, csize) ? 0 : VAE_ERR_INVALID;
}

// Clusters of `cs` CTAs of the kernel that can be co-resident on the device (one CTA per SM).
static int tc_active_clusters(int cs, int* out) {
  if (cs == 1) {
    int dev = 0, sms = 0;
    VAE_CUDA_OK(cudaGetDevice(&dev));
    VAE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    *out = sms;
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cs, 1, 1);
  cfg.blockDim = dim3(kTT, 1, 1);
  cfg.dynamicSmemBytes = kTSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  VAE_CUDA_OK(cudaOccupancyMaxActiveClusters(out, fused_mlp_tc_kernel, &cfg));
  return 0;
}

// Sets the kernel's shared-memory attribute and picks the cluster size for a pack of n_jobs models: the largest of
// 8 / 4 / 2 CTAs per model that keeps the whole pack resident (latency), else one CTA per model (throughput).
// VAE_MLP_TC_CLUSTER=1|2|4|8 forces a size (A/B runs).
int mlp_tc_prepare(int n_jobs, int* cluster) {
  cudaFuncAttributes fa;
  VAE_CUDA_OK(cudaFuncGetAttributes(&fa, fused_mlp_tc_kernel));
  VAE_CUDA_OK(cudaFuncSetAttribute(fused_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024 - (int)fa.sharedSizeBytes));
  if ((size_t)kTSmemBytes + fa.sharedSizeBytes > (size_t)227 * 1024) {
    set_error("fused_mlp_tc_kernel needs %zu bytes of shared memory", (size_t)kTSmemBytes + fa.sharedSizeBytes);
    return VAE_ERR_INVALID;
  }
  int cs = 1;
  const char* e = getenv("VAE_MLP_TC_CLUSTER");
  if (e && e[0]) {
    const int v = atoi(e);
    cs = v >= 16 ? 16 : v >= 8 ? 8 : v >= 4 ? 4 : v >= 2 ? 2 : 1;
    if (cs == 16) VAE_CUDA_OK(cudaFuncSetAttribute(fused_mlp_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  } else {
    for (int c = kTAutoCluster; c > 1; c >>= 1) {
      int fit = 0;
      if (tc_active_clusters(c, &fit) != 0) { cudaGetLastError(); continue; }
      if (n_jobs <= fit) { cs = c; break; }
    }
  }
  *cluster = cs;
  return 0;
}

int mlp_tc_launch(const VaePackPlan* plan, int64_t n_steps, cudaStream_t s) {
  const int cs = plan->cluster < 1 ? 1 : plan->cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(plan->n_jobs * cs), 1, 1);
  cfg.blockDim = dim3(kTT, 1, 1);
  cfg.dynamicSmemBytes = kTSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cs > 1 ? 1 : 0;
  VAE_CUDA_OK(cudaLaunchKernelEx(&cfg, fused_mlp_tc_kernel, (const TcJobDev*)plan->d_jobs, n_steps));
  count_launch();
  return 0;
}

#ifdef VAE_TC_PROFILE
extern "C" int vae_tc_profile(long long* out16, int reset) {
  long long z[16] = {0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(z));
  if (reset) cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
  return 0;
}
#endif

}  // namespace vae
