// gemm_tc.cu -- tcgen05 / TMEM / TMA GEMM with FP32-level accuracy (3xTF32 split), for the
// one configuration of the path whose Dense layers are true dense contractions:
// BASELINE.json configs[4] (D=1024, width 4096, batch 65 536).
//
//   C[M,N] = epi( A[M,K] * B[N,K]^T )        ("NT": both operands K-contiguous, row-major)
//   C[M,N] = At[K,M]^T * Bt[K,N]              ("TN": both operands given K-row-major, i.e. M- / N-contiguous: the dW =
//                                              X^T dY form, read straight from the batch-major activations / gradients
//                                              through MN-major tcgen05 descriptors -- no transposed copies in HBM)
//
// Plain TF32 (10-bit mantissa) cannot meet the path's rel-1e-5-per-step tolerance, so every
// FP32 operand X is split as X ~= hi + lo with hi = TF32(X) and lo = TF32(X - hi) (both rounded
// to nearest; hi + lo reproduces X to ~2^-23), and the product is accumulated in FP32 in
// tensor memory as
//   A*B ~= Ah*Bh + Ah*Bl + Al*Bh         (the dropped Al*Bl term is ~2^-22 relative)
// i.e. three M128 N128 K8 products per k-step, issued as TWO tcgen05.mma.kind::tf32: Ah * [Bh ; Bl]^T
// as one N = 256 MMA (Bh and Bl are adjacent in the stage, so one descriptor spans both: Ah is read
// from shared memory once, not twice -- the kernel is shared-memory-bandwidth bound) and Al * Bh^T.
// Only algorithmic flops (2MNK) are ever reported as useful work.
//
// Structure (persistent: a CTA -- or a cluster of 2 / 4 -- walks 128x128 output tiles; 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B boxes of 128 rows x 32 floats)
//               of Ah, Al, Bh, Bl into a 3-stage smem ring, mbarrier expect_tx
//   warp 1      MMA issuer: one elected lane issues, per stage, 4 k-steps x 2 instructions of
//               tcgen05.mma.cta_group::1.kind::tf32 (M128 N256 K8 and M128 N128 K8), accumulators
//               in tensor memory; tcgen05.commit releases the smem stage / signals the epilogue
//   warps 2-5   drain + epilogue: tcgen05.ld 32x32b (each warp its own TMEM lane quarter),
//               two-level accumulation (see below), fused bias / relu / relu-mask, float4 stores
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace vae {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;             // 16 KB
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;            // Ah Al Bh Bl
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024;
constexpr int TC_THREADS = 192;
constexpr int TC_GROUP_M = 8;                                // cluster tiles along M per rasterisation group
enum { TC_BIAS = 1, TC_RELU = 2, TC_MASK = 4 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// Multicast variant: the box lands at the same smem offset (and signals the mbarrier at the same
// offset) in every CTA of the cluster named by cta_mask.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "h"(cta_mask), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tc_cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// One lane of the converged warp (1 in the elected lane).  The producer and MMA warps run their loops with all 32 lanes
// on warp-uniform values and predicate only the TMA / MMA / commit instructions on it: inside `if (lane == 0)` the
// compiler must assume divergent operands and wraps every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST waterfall
// (~70 - 100 cycles per instruction; measured on fused_mlp_tc.cu).
__device__ __forceinline__ uint32_t tc_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of
// 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand tile for kind::tf32 (TN mode): the tile is stored as 32-wide M (or N) blocks of [32 k-rows][128 B],
// one TMA box {32 floats, 32 rows} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B each (32-byte chunk index XOR (row % 4)) --
// the SWIZZLE_128B_BASE32B layout (type 1), the only one an MN-major TF32 operand is read correctly from
// (tests/test_gpu_umma_probe.py).  LBO = 4096 B between the 32-wide blocks, SBO = 512 B between 4-row groups; a K = 8
// step is 8 rows = 1024 B further.
__device__ __forceinline__ uint64_t make_mnmajor_b32_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// cute::UMMA::InstrDescriptor: c=F32, a=b=TF32, N>>3 at [17,23), M>>4 at [24,29); bits 15 / 16: A / B is MN-major
__host__ __device__ constexpr uint32_t make_tf32_idesc(int M, int N, bool mn_major = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The tensor core accumulates in FP32 with round-toward-zero: a long accumulation chain in
// TMEM picks up a systematic bias that grows linearly with K (measured: 2.5e-5 relative at
// K = 4096).  Two-level accumulation keeps FP32-level accuracy:
//   * TMEM holds two 256-column buffers (main | correction columns); a buffer accumulates a chunk of
//     TC_CHUNK k-blocks (K = 256: 32 tensor-core accumulations), then the epilogue warps drain it
//     into FP32 registers (round-to-nearest adds of main + correction) while the tensor core fills
//     the other buffer;
//   * the correction terms (2^-11 of the magnitude) sit in their own columns, so their small addends
//     are never truncated against the large main sums.
constexpr int TC_CHUNK = 8;                     // k-blocks per main-accumulator chunk
constexpr int TC_TMEM_COLS = 512;               // main0 | corr0 | main1 | corr1

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive columns of this warp's TMEM lane quarter -> 32 registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MC = 2: clusters of two CTAs that are neighbours along M share their B tiles -- each CTA
// fetches half of Bh / Bl (64 rows) and TMA-multicasts it into both CTAs' shared memory, which
// cuts the L2 -> SM operand traffic per CTA from 64 KB to 48 KB per k-block (the kernel is
// L2-bandwidth bound: ncu shows 13 TB/s of xbar reads at 62 % tensor-pipe activity for MC = 1).
// MC = 4: 2 x 2 clusters; the two CTAs with the same M tile additionally share their A tiles the
// same way (32 KB per k-block per CTA).  Cluster rank = cx + CX * cy (cx along N, cy along M).
// A stage may be refilled only when EVERY CTA that reads what this CTA multicasts has consumed it:
// tcgen05.commit arrives on the empty barrier of this CTA and of its row / column partners.
// PERSISTENT: one cluster per SM group walks the output tiles (cluster tile ct = cn + ctn * cm, n
// fastest) with the shared-memory ring, the main-accumulator ping-pong and a double-buffered
// correction accumulator all running ACROSS tiles: the TMA producer and the MMA issuer start tile
// t+1 while the epilogue warps still drain / store tile t, so neither the pipeline fill nor the
// epilogue is exposed per tile.  TMEM (512 columns): main0 | corr0 | main1 | corr1.
template <int MC, bool TN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_nt_3xtf32_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                      const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                      float* __restrict__ C, int M, int N, int K, int64_t ldc, const float* __restrict__ bias,
                      const float* __restrict__ mask, int64_t ldm, int flags, int ctn, int n_ctiles) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t mfull_bar[2];     // main accumulator chunk complete
  __shared__ __align__(8) uint64_t mempty_bar[2];    // main accumulator drained (4 epilogue warps)
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + TC_BK - 1) / TC_BK;
  const int nchunk = (nkb + TC_CHUNK - 1) / TC_CHUNK;
  constexpr int CX = MC == 4 ? 2 : 1, CY = MC >= 2 ? 2 : 1;         // cluster extent along N / along M
  const uint32_t crank = MC > 1 ? tc_cluster_ctarank() : 0u;
  const uint32_t cx = crank % CX, cy = crank / CX;
  // CTAs with the same M tile (A partners) and with the same N tile (B partners), as rank masks
  const uint16_t a_mask = (uint16_t)(((1u << CX) - 1u) << (cy * CX));
  const uint16_t b_mask = (uint16_t)((1u << cx) | (CY > 1 ? 1u << (cx + CX) : 0u));
  const uint16_t peers_mask = (uint16_t)(a_mask | b_mask);
  const int cluster_id = (int)blockIdx.x / MC, n_clusters = (int)gridDim.x / MC;
  // this CTA's tile of cluster tile ct
  // Grouped rasterisation: consecutive cluster tiles walk TC_GROUP_M tiles down M before moving one
  // tile along N, so the clusters that run at the same time share a near-square set of A / B panels
  // (less DRAM traffic per wave than sweeping all of N for two M tiles).
  const int ctm = n_ctiles / ctn;
  auto tile_of = [&](int ct, int& m0, int& n0) {
    const int gsz_full = TC_GROUP_M * ctn, grp = ct / gsz_full, rem = ct - grp * gsz_full;
    const int first_m = grp * TC_GROUP_M, gm = min(ctm - first_m, TC_GROUP_M);
    const int cm = first_m + rem % gm, cn = rem / gm;
    n0 = (cn * CX + (int)cx) * TC_BN;
    m0 = (cm * CY + (int)cy) * TC_BM;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAh)));
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAl)));
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBh)));
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBl)));
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CX + CY - 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&mfull_bar[b], 1); mbar_init(&mempty_bar[b], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // one warp allocates the tensor memory (power of two >= 32 columns) and owns the dealloc
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  if (MC > 1) tc_cluster_sync();            // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues)
    {
      uint32_t it = 0;                                       // k-blocks issued so far, across tiles
      for (int ct = cluster_id; ct < n_ctiles; ct += n_clusters) {
        int m0, n0;
        tile_of(ct, m0, n0);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = (int)(it % TC_STAGES);
          const uint32_t ph = (it / TC_STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);                 // slot free (first pass: immediately)
          uint8_t* st = smem + (size_t)s * TC_STAGE_BYTES;
          if (!tc_elect()) { __syncwarp(); continue; }
          mbar_expect_tx(&full_bar[s], TC_STAGE_BYTES);
          if (TN) {
            // TN: operand tiles are 4 blocks of [32 k-rows][32 M- or N-floats] (box x = column, y = k row); a CTA of a
            // multicast pair fetches 2 of the 4 blocks of the shared operand
            constexpr int kBlk = 32 * TC_BK * 4;                                  // 4 KB
            for (int j = 0; j < 4; ++j) {
              if (CX == 1) {
                tma_load_2d(st + 0 * TC_TILE_BYTES + j * kBlk, &tmAh, &full_bar[s], m0 + 32 * j, kb * TC_BK);
                tma_load_2d(st + 1 * TC_TILE_BYTES + j * kBlk, &tmAl, &full_bar[s], m0 + 32 * j, kb * TC_BK);
              } else if ((j >> 1) == (int)cx) {
                tma_load_2d_mc(st + 0 * TC_TILE_BYTES + j * kBlk, &tmAh, &full_bar[s], m0 + 32 * j, kb * TC_BK, a_mask);
                tma_load_2d_mc(st + 1 * TC_TILE_BYTES + j * kBlk, &tmAl, &full_bar[s], m0 + 32 * j, kb * TC_BK, a_mask);
              }
              if (CY == 1) {
                tma_load_2d(st + 2 * TC_TILE_BYTES + j * kBlk, &tmBh, &full_bar[s], n0 + 32 * j, kb * TC_BK);
                tma_load_2d(st + 3 * TC_TILE_BYTES + j * kBlk, &tmBl, &full_bar[s], n0 + 32 * j, kb * TC_BK);
              } else if ((j >> 1) == (int)cy) {
                tma_load_2d_mc(st + 2 * TC_TILE_BYTES + j * kBlk, &tmBh, &full_bar[s], n0 + 32 * j, kb * TC_BK, b_mask);
                tma_load_2d_mc(st + 3 * TC_TILE_BYTES + j * kBlk, &tmBl, &full_bar[s], n0 + 32 * j, kb * TC_BK, b_mask);
              }
            }
            __syncwarp();
            continue;
          }
          if (CX == 1) {
            tma_load_2d(st + 0 * TC_TILE_BYTES, &tmAh, &full_bar[s], kb * TC_BK, m0);
            tma_load_2d(st + 1 * TC_TILE_BYTES, &tmAl, &full_bar[s], kb * TC_BK, m0);
          } else {      // this CTA's half of the A rows, multicast to the CTAs with the same M tile
            const int half_rows = TC_BM / CX;
            const size_t off = (size_t)cx * half_rows * TC_BK * 4;
            tma_load_2d_mc(st + 0 * TC_TILE_BYTES + off, &tmAh, &full_bar[s], kb * TC_BK, m0 + (int)cx * half_rows, a_mask);
            tma_load_2d_mc(st + 1 * TC_TILE_BYTES + off, &tmAl, &full_bar[s], kb * TC_BK, m0 + (int)cx * half_rows, a_mask);
          }
          if (CY == 1) {
            tma_load_2d(st + 2 * TC_TILE_BYTES, &tmBh, &full_bar[s], kb * TC_BK, n0);
            tma_load_2d(st + 3 * TC_TILE_BYTES, &tmBl, &full_bar[s], kb * TC_BK, n0);
          } else {      // this CTA's half of the B rows, multicast to the CTAs with the same N tile
            const int half_rows = TC_BN / CY;
            const size_t off = (size_t)cy * half_rows * TC_BK * 4;
            tma_load_2d_mc(st + 2 * TC_TILE_BYTES + off, &tmBh, &full_bar[s], kb * TC_BK, n0 + (int)cy * half_rows, b_mask);
            tma_load_2d_mc(st + 3 * TC_TILE_BYTES + off, &tmBl, &full_bar[s], kb * TC_BK, n0 + (int)cy * half_rows, b_mask);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp on uniform values, one elected lane issues)
    {
      constexpr uint32_t idesc = make_tf32_idesc(TC_BM, TC_BN, TN), idesc2 = make_tf32_idesc(TC_BM, 2 * TC_BN, TN);
      uint32_t it = 0, gc = 0;                               // k-blocks, accumulator chunks so far
      for (int ct = cluster_id; ct < n_ctiles; ct += n_clusters) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = (int)(it % TC_STAGES);
          const uint32_t ph = (it / TC_STAGES) & 1u;
          const uint32_t buf = gc & 1u;
          const bool first = (kb % TC_CHUNK) == 0;
          if (first) {                                        // the drain of this buffer's previous chunk is done
            mbar_wait(&mempty_bar[buf], ((gc >> 1) & 1u) ^ 1u);
            tc_fence_after();
          }
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + (size_t)s * TC_STAGE_BYTES);
          const uint64_t dAh = TN ? make_mnmajor_b32_desc(base + 0 * TC_TILE_BYTES) : make_kmajor_sw128_desc(base + 0 * TC_TILE_BYTES);
          const uint64_t dAl = TN ? make_mnmajor_b32_desc(base + 1 * TC_TILE_BYTES) : make_kmajor_sw128_desc(base + 1 * TC_TILE_BYTES);
          const uint64_t dBh = TN ? make_mnmajor_b32_desc(base + 2 * TC_TILE_BYTES) : make_kmajor_sw128_desc(base + 2 * TC_TILE_BYTES);   // Bh, and Bl right behind it
          const uint32_t tmem_main = tmem_acc + buf * (2 * TC_BN);
          if (tc_elect()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            // per K = 8 step: +32 bytes inside the 128-byte row (K-major) / 8 rows = 1024 bytes further (MN-major)
            const uint64_t adv = (uint64_t)((TN ? k * 1024 : k * 8 * 4) >> 4);
            // Ah * [Bh ; Bl]^T in ONE N = 256 MMA: columns 0-127 main term, 128-255 first correction
            // (Ah is read from shared memory once instead of twice)
            umma_tf32(tmem_main, dAh + adv, dBh + adv, idesc2, !(first && k == 0));
            umma_tf32(tmem_main + TC_BN, dAl + adv, dBh + adv, idesc, 1u);       // second correction, N = 128
          }
          if (MC == 1) umma_commit(&empty_bar[s]);           // smem slot reusable once these MMAs retire
          else umma_commit_mc(&empty_bar[s], peers_mask);
          if ((kb % TC_CHUNK) == TC_CHUNK - 1 || kb == nkb - 1) umma_commit(&mfull_bar[buf]);
          }
          __syncwarp();
          if ((kb % TC_CHUNK) == TC_CHUNK - 1 || kb == nkb - 1) ++gc;
        }
      }
    }
  } else {
    // ===================== epilogue warps 2..5: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t gc = 0;
    for (int ct = cluster_id; ct < n_ctiles; ct += n_clusters) {
      int m0, n0;
      tile_of(ct, m0, n0);
      const int row = m0 + q * 32 + lane;
      float acc[TC_BN];
#pragma unroll
      for (int j = 0; j < TC_BN; ++j) acc[j] = 0.f;
      for (int chunk = 0; chunk < nchunk; ++chunk, ++gc) {
        const uint32_t buf = gc & 1u;
        mbar_wait(&mfull_bar[buf], (gc >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < TC_BN / 32; ++c) {               // main columns, then the chunk's correction columns
          uint32_t v[32];
          tmem_ld32(tmem_acc + lane_off + (uint32_t)(buf * 2 * TC_BN + TC_BN + c * 32), v);      // corrections first (small)
          float corr[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) corr[j] = __uint_as_float(v[j]);
          tmem_ld32(tmem_acc + lane_off + (uint32_t)(buf * 2 * TC_BN + c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]) + corr[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&mempty_bar[buf]);
      }
#pragma unroll
      for (int c = 0; c < TC_BN / 32; ++c) {
        if (row < M) {
          const int nb = n0 + c * 32;
          float* crow = C + (int64_t)row * ldc + nb;
          const float* mrow = (flags & TC_MASK) ? mask + (int64_t)row * ldm + nb : nullptr;
          const bool vec = (nb + 32 <= N) && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float x = acc[c * 32 + j + u];
              if (nb + j + u < N) {
                if (flags & TC_BIAS) x += bias[nb + j + u];
                if (flags & TC_RELU) x = fmaxf(x, 0.f);
                if (flags & TC_MASK) x = mrow[j + u] > 0.f ? x : 0.f;
              }
              o[u] = x;
            }
            if (vec) {
              *reinterpret_cast<float4*>(crow + j) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (nb + j + u < N) crow[j + u] = o[u];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC > 1) tc_cluster_sync();            // no CTA leaves while its peer may still signal its barriers
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TC_TMEM_COLS));
}

// ------------------------------------------------------------------ operand preparation
// hi = TF32(x) rounded to nearest, lo = TF32(x - hi) rounded to nearest (x - hi is exact in
// FP32 and |lo| <= 2^-11 |x|, so hi + lo reproduces x to ~2^-23); optionally transposed.
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_one(float v, float& h, float& l) { h = tf32_rn(v); l = tf32_rn(v - h); }
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      float4 h, l;
      split_one(v.x, h.x, l.x); split_one(v.y, h.y, l.y); split_one(v.z, h.z, l.z); split_one(v.w, h.w, l.w);
      *reinterpret_cast<float4*>(hi + i) = h;
      *reinterpret_cast<float4*>(lo + i) = l;
    } else {
      for (int64_t j = i; j < n; ++j) split_one(x[j], hi[j], lo[j]);
    }
  }
}
// x [R][C] (ldx) -> hiT, loT [C][R] (ldt): 32x32 smem tiles, coalesced both ways
__global__ void __launch_bounds__(256)
transpose_split_tf32_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ hiT, float* __restrict__ loT,
                            int64_t ldt, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? x[(int64_t)r * ldx + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < R) {
      float h, l;
      split_one(tile[tx][i], h, l);
      hiT[(int64_t)c * ldt + r] = h;
      loT[(int64_t)c * ldt + r] = l;
    }
  }
}

// ------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;      // idempotent lookup of a driver entry point (no mutable library state)
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rows x K row-major (leading dimension ld floats): boxes of 128 rows x 32 floats, SWIZZLE_128B
static int make_operand_map(CUtensorMap* tm, const float* ptr, int rows, int K, int64_t ld, int box_rows = TC_BM) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VAE_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return VAE_ERR_CUDA; }
  return 0;
}

// VAE_TC_CLUSTER=1|2|4 in the environment caps the cluster size (A/B comparison).  Default 2:
// measured on B200 (scratch/bench_gemm_ab.py, 65536 x 4096 x 4096 and 8192 x 4096 x 4096) the
// 2 x 2 clusters are 8-17 % SLOWER than pairs -- clusters of four strand SMs (GPCs of 16/18/20
// SMs) and the run is power-capped, so the lower L2 traffic buys nothing.
static int tc_cluster_cap() {
  const char* e = getenv("VAE_TC_CLUSTER");
  if (!e || !e[0]) return 2;
  const int v = atoi(e);
  return v >= 4 ? 4 : v >= 2 ? 2 : 1;
}

// TN operand: K rows x `cols` floats row-major (leading dimension ld): boxes of 32 k-rows x 32 floats, SWIZZLE_128B_ATOM_32B
static int make_operand_map_tn(CUtensorMap* tm, const float* ptr, int K, int cols, int64_t ld) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VAE_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)K};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)TC_BK};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (TN operand) failed (CUresult %d)", (int)r); return VAE_ERR_CUDA; }
  return 0;
}

template <int MC, bool TN = false>
static int launch_gemm(const CUtensorMap& tAh, const CUtensorMap& tAl, const CUtensorMap& tBh, const CUtensorMap& tBl,
                       float* d_C, int M, int N, int K, int64_t ldc, const float* d_bias, const float* d_mask, int64_t ldm,
                       int flags, int tiles_n, int tiles_m, cudaStream_t stream) {
  constexpr int CX = MC == 4 ? 2 : 1, CY = MC >= 2 ? 2 : 1;
  VAE_CUDA_OK(cudaFuncSetAttribute(gemm_nt_3xtf32_kernel<MC, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  // cluster tiles: ctn along N, ctm along M (an odd last M tile of a pair is all out-of-bounds: zero-filled
  // loads, masked stores); the persistent grid is one cluster per MC SMs, or fewer if there are fewer tiles
  const int ctn = tiles_n / CX, ctm = (tiles_m + CY - 1) / CY;
  const int n_ctiles = ctn * ctm;
  int dev = 0, sms = 148;
  VAE_CUDA_OK(cudaGetDevice(&dev));
  VAE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int n_clusters = sms / MC;
  if (n_clusters > n_ctiles) n_clusters = n_ctiles;
  if (n_clusters < 1) n_clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * MC, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = MC > 1 ? 1 : 0;
  VAE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_nt_3xtf32_kernel<MC, TN>, tAh, tAl, tBh, tBl, d_C, M, N, K, ldc, d_bias, d_mask, ldm,
                                 flags, ctn, n_ctiles));
  count_launch();
  return 0;
}

}  // namespace vae

extern "C" {

int vae_split_tf32(const float* d_x, float* d_hi, float* d_lo, int64_t n, void* stream) {
  VAE_REQUIRE(d_x && d_hi && d_lo && n >= 0, "vae_split_tf32: bad arguments");
  VAE_REQUIRE(((uintptr_t)d_x | (uintptr_t)d_hi | (uintptr_t)d_lo) % 16 == 0, "vae_split_tf32: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  vae::split_tf32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_x, d_hi, d_lo, n);
  VAE_LAUNCH_OK("split_tf32_kernel");
  return 0;
}

int vae_transpose_split_tf32(const float* d_x, int64_t ldx, float* d_hiT, float* d_loT, int64_t ldt, int32_t rows,
                             int32_t cols, void* stream) {
  VAE_REQUIRE(d_x && d_hiT && d_loT && rows > 0 && cols > 0 && ldx >= cols && ldt >= rows, "vae_transpose_split_tf32: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  vae::transpose_split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, ldx, d_hiT, d_loT, ldt, rows, cols);
  VAE_LAUNCH_OK("transpose_split_tf32_kernel");
  return 0;
}

int vae_gemm_nt_3xtf32(int32_t M, int32_t N, int32_t K, const float* d_Ah, const float* d_Al, int64_t lda,
                       const float* d_Bh, const float* d_Bl, int64_t ldb, float* d_C, int64_t ldc,
                       const float* d_bias, const float* d_mask, int64_t ldm, int32_t flags, void* stream) {
  VAE_REQUIRE(M > 0 && N > 0 && K > 0 && d_Ah && d_Al && d_Bh && d_Bl && d_C, "vae_gemm_nt_3xtf32: bad arguments");
  VAE_REQUIRE(lda >= K && ldb >= K && ldc >= N, "vae_gemm_nt_3xtf32: leading dimensions too small");
  VAE_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "vae_gemm_nt_3xtf32: lda / ldb must be multiples of 4 floats (TMA 16-byte strides)");
  VAE_REQUIRE((((uintptr_t)d_Ah | (uintptr_t)d_Al | (uintptr_t)d_Bh | (uintptr_t)d_Bl) & 15) == 0,
              "vae_gemm_nt_3xtf32: operands must be 16-byte aligned");
  VAE_REQUIRE(!(flags & vae::TC_BIAS) || d_bias, "vae_gemm_nt_3xtf32: bias is NULL");
  VAE_REQUIRE(!(flags & vae::TC_MASK) || (d_mask && ldm >= N), "vae_gemm_nt_3xtf32: mask is NULL / ldm too small");
  CUtensorMap tAh, tAl, tBh, tBl;
  int rc;
  const int tiles_m = (M + vae::TC_BM - 1) / vae::TC_BM, tiles_n = (N + vae::TC_BN - 1) / vae::TC_BN;
  const int cap = vae::tc_cluster_cap();
  const int mc = (cap >= 4 && tiles_m >= 2 && tiles_n % 2 == 0) ? 4 : (cap >= 2 && tiles_m >= 2) ? 2 : 1;
  const int cx = mc == 4 ? 2 : 1, cy = mc >= 2 ? 2 : 1;
  if ((rc = vae::make_operand_map(&tAh, d_Ah, M, K, lda, vae::TC_BM / cx))) return rc;
  if ((rc = vae::make_operand_map(&tAl, d_Al, M, K, lda, vae::TC_BM / cx))) return rc;
  if ((rc = vae::make_operand_map(&tBh, d_Bh, N, K, ldb, vae::TC_BN / cy))) return rc;
  if ((rc = vae::make_operand_map(&tBl, d_Bl, N, K, ldb, vae::TC_BN / cy))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mc == 4) return vae::launch_gemm<4>(tAh, tAl, tBh, tBl, d_C, M, N, K, ldc, d_bias, d_mask, ldm, flags, tiles_n, tiles_m, st);
  if (mc == 2) return vae::launch_gemm<2>(tAh, tAl, tBh, tBl, d_C, M, N, K, ldc, d_bias, d_mask, ldm, flags, tiles_n, tiles_m, st);
  return vae::launch_gemm<1>(tAh, tAl, tBh, tBl, d_C, M, N, K, ldc, d_bias, d_mask, ldm, flags, tiles_n, tiles_m, st);
}

int vae_gemm_tn_3xtf32(int32_t M, int32_t N, int32_t K, const float* d_Ath, const float* d_Atl, int64_t lda,
                       const float* d_Bth, const float* d_Btl, int64_t ldb, float* d_C, int64_t ldc, void* stream) {
  VAE_REQUIRE(M > 0 && N > 0 && K > 0 && d_Ath && d_Atl && d_Bth && d_Btl && d_C, "vae_gemm_tn_3xtf32: bad arguments");
  VAE_REQUIRE(lda >= M && ldb >= N && ldc >= N, "vae_gemm_tn_3xtf32: leading dimensions too small");
  VAE_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "vae_gemm_tn_3xtf32: lda / ldb must be multiples of 4 floats (TMA 16-byte strides)");
  VAE_REQUIRE((((uintptr_t)d_Ath | (uintptr_t)d_Atl | (uintptr_t)d_Bth | (uintptr_t)d_Btl) & 15) == 0,
              "vae_gemm_tn_3xtf32: operands must be 16-byte aligned");
  CUtensorMap tAh, tAl, tBh, tBl;
  int rc;
  const int tiles_m = (M + vae::TC_BM - 1) / vae::TC_BM, tiles_n = (N + vae::TC_BN - 1) / vae::TC_BN;
  const int cap = vae::tc_cluster_cap();
  const int mc = (cap >= 4 && tiles_m >= 2 && tiles_n % 2 == 0) ? 4 : (cap >= 2 && tiles_m >= 2) ? 2 : 1;
  if ((rc = vae::make_operand_map_tn(&tAh, d_Ath, K, M, lda))) return rc;
  if ((rc = vae::make_operand_map_tn(&tAl, d_Atl, K, M, lda))) return rc;
  if ((rc = vae::make_operand_map_tn(&tBh, d_Bth, K, N, ldb))) return rc;
  if ((rc = vae::make_operand_map_tn(&tBl, d_Btl, K, N, ldb))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mc == 4) return vae::launch_gemm<4, true>(tAh, tAl, tBh, tBl, d_C, M, N, K, ldc, nullptr, nullptr, 0, 0, tiles_n, tiles_m, st);
  if (mc == 2) return vae::launch_gemm<2, true>(tAh, tAl, tBh, tBl, d_C, M, N, K, ldc, nullptr, nullptr, 0, 0, tiles_n, tiles_m, st);
  return vae::launch_gemm<1, true>(tAh, tAl, tBh, tBl, d_C, M, N, K, ldc, nullptr, nullptr, 0, 0, tiles_n, tiles_m, st);
}

}  // extern "C"
