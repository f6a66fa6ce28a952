// umma_probe.cu -- hardware probe for the tcgen05 operand forms the hidden-layer kernel (fused_mlp_tc.cu)
// relies on: K-major and MN-major SWIZZLE_128B shared-memory descriptors for kind::tf32, odd N (104),
// and what the tensor core does with the low 13 mantissa bits of a raw FP32 operand.  One CTA copies two
// caller-built shared-memory images verbatim, issues `ksteps` M128 x N x K8 MMAs (descriptor start
// addresses advanced by a_step / b_step bytes per k-step; the whole k loop `repeat` times) and returns the
// 128 x N FP32 accumulator, followed by the elapsed SM cycles from the first issue to completion in d_out[128 N].
// tests/test_gpu_umma_probe.py builds the images in NumPy and checks the results; nothing on the
// training path calls this.
#include "common.cuh"

namespace vae {
namespace {

__device__ __forceinline__ uint32_t p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const float* __restrict__ a_img, int a_bytes, const float* __restrict__ b_img, int b_bytes,
                  uint64_t desc_a_hi, uint64_t desc_b_hi, uint32_t idesc, int ksteps, int a_step, int b_step, int N,
                  int repeat, float* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  float* sa = reinterpret_cast<float*>(smem);
  float* sb = reinterpret_cast<float*>(smem + ((a_bytes + 1023) & ~1023));
  for (int i = threadIdx.x; i < a_bytes / 4; i += 128) sa[i] = a_img[i];
  for (int i = threadIdx.x; i < b_bytes / 4; i += 128) sb[i] = b_img[i];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(p_smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(p_smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> async proxy (MMA) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const long long t_begin = clock64();
  if (threadIdx.x == 0) {
    const uint64_t da0 = desc_a_hi | (uint64_t)((p_smem_u32(sa) >> 4) & 0x3FFF);
    const uint64_t db0 = desc_b_hi | (uint64_t)((p_smem_u32(sb) >> 4) & 0x3FFF);
    for (int rep = 0; rep < repeat; ++rep)
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t da = da0 + (uint64_t)((k * a_step) >> 4), db = db0 + (uint64_t)((k * b_step) >> 4);
      const uint32_t acc = (k | rep) > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(p_smem_u32(&bar)) : "memory");
  }
  {
    const long long t0 = clock64();
    for (;;) {
      uint32_t done;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(p_smem_u32(&bar)), "r"(0u) : "memory");
      if (done) break;
      if (clock64() - t0 > 2000000000ll) __trap();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) out[128 * N] = (float)(clock64() - t_begin);      // cycles: issue of all MMAs -> completion
  const int row = threadIdx.x;                   // TMEM lane = accumulator row
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

}  // namespace
}  // namespace vae

extern "C" int vae_umma_probe(const float* d_a_img, int32_t a_bytes, const float* d_b_img, int32_t b_bytes,
                              uint64_t desc_a_hi, uint64_t desc_b_hi, uint32_t idesc, int32_t ksteps, int32_t a_step,
                              int32_t b_step, int32_t N, int32_t repeat, float* d_out, void* stream) {
  VAE_REQUIRE(d_a_img && d_b_img && d_out && a_bytes > 0 && b_bytes > 0 && a_bytes % 16 == 0 && b_bytes % 16 == 0,
              "vae_umma_probe: bad images");
  VAE_REQUIRE(N >= 8 && N <= 256 && N % 8 == 0 && ksteps > 0, "vae_umma_probe: bad shape");
  const size_t smem = (size_t)((a_bytes + 1023) & ~1023) + (size_t)b_bytes + 2048;
  VAE_REQUIRE(smem <= 200 * 1024, "vae_umma_probe: images too large");
  VAE_CUDA_OK(cudaFuncSetAttribute(vae::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  vae::umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d_a_img, a_bytes, d_b_img, b_bytes, desc_a_hi, desc_b_hi, idesc,
                                                                 ksteps, a_step, b_step, N, repeat < 1 ? 1 : repeat, d_out);
  VAE_LAUNCH_OK("umma_probe_kernel");
  return 0;
}
