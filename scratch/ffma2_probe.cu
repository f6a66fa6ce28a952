// FP32 FMA issue rates as the tile loops of fused_mlp.cu use them (register operands only), per SM sub-partition:
//   mode 0: FFMA2 (fma.rn.f32x2), three register operands, 8 accumulator pairs, fresh a / w operands every 8 instructions
//   mode 1: scalar FFMA, three register operands, 16 accumulators
//   mode 2: the tile_gemm loop itself on shared memory (LDS.128 x 2 + 8 FFMA2 per k), 100 k-steps
// one CTA, `warps` warps (warps / 4 per sub-partition); prints cycles per instruction per sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/ffma2_probe scratch/ffma2_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_ffma2(int iters, long long* out, float* sink, const float* src) {
  float2 acc[8];
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(threadIdx.x * 1e-3f, j);
  float4 a = *reinterpret_cast<const float4*>(src + 4 * (threadIdx.x & 31));
  float4 w = *reinterpret_cast<const float4*>(src + 128 + 4 * (threadIdx.x & 31));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float2 a01 = make_float2(a.x, a.y), a23 = make_float2(a.z, a.w);
    const float wj[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 ww = make_float2(wj[j], wj[j]);
      acc[j] = __ffma2_rn(ww, a01, acc[j]);
      acc[4 + j] = __ffma2_rn(ww, a23, acc[4 + j]);
    }
    a.x += 1e-9f; w.y += 1e-9f;      // keep the operands live as registers without letting the loop collapse
  }
  const long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += acc[j].x + acc[j].y;
  sink[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}
__global__ void k_ffma(int iters, long long* out, float* sink, const float* src) {
  float acc[16];
  for (int j = 0; j < 16; ++j) acc[j] = threadIdx.x * 1e-3f + j;
  float4 a = *reinterpret_cast<const float4*>(src + 4 * (threadIdx.x & 31));
  float4 w = *reinterpret_cast<const float4*>(src + 128 + 4 * (threadIdx.x & 31));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float av[4] = {a.x, a.y, a.z, a.w}, wj[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[4 * j + r] = fmaf(wj[j], av[r], acc[4 * j + r]);
    a.x += 1e-9f; w.y += 1e-9f;
  }
  const long long t1 = clock64();
  float s = 0; for (int j = 0; j < 16; ++j) s += acc[j];
  sink[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}
// the tile_gemm k loop: AT [K][BP] and Ws [K][W4] in shared memory, 4x4 tile per lane pair
__global__ void k_tile(int K, int BP, int W4, long long* out, float* sink) {
  extern __shared__ __align__(16) float sm[];
  float* AT = sm; float* Ws = sm + K * BP;
  for (int i = threadIdx.x; i < K * BP + K * W4; i += blockDim.x) sm[i] = 1e-3f * (i % 97);
  __syncthreads();
  const int NG = BP >> 2, ncg = W4 >> 2, ntile = ncg * NG;
  const int tt = threadIdx.x, tile = min(tt >> 1, ntile - 1), half = tt & 1;
  const int cg = tile / NG, rg = tile - cg * NG, kh = (K + 1) >> 1;
  float2 acc2[4][2];
  for (int j = 0; j < 4; ++j) { acc2[j][0] = make_float2(0.f, 0.f); acc2[j][1] = make_float2(0.f, 0.f); }
  const float* ap = AT + 4 * rg; const float* wp = Ws + 4 * cg;
  const int kbeg = half ? kh : 0, kend = half ? K : kh;
  const long long t0 = clock64();
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
  const long long t1 = clock64();
  float s = 0; for (int j = 0; j < 4; ++j) s += acc2[j][0].x + acc2[j][0].y + acc2[j][1].x + acc2[j][1].y;
  sink[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}
int main() {
  long long* d_out; float* d_sink; float* d_src;
  cudaMalloc(&d_out, 64 * sizeof(long long)); cudaMalloc(&d_sink, 4096 * sizeof(float)); cudaMalloc(&d_src, 4096 * sizeof(float));
  cudaMemset(d_src, 0, 4096 * sizeof(float));
  long long h[64];
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k_ffma2<<<1, warps * 32>>>(iters, d_out, d_sink, d_src); else k_ffma<<<1, warps * 32>>>(iters, d_out, d_sink, d_src);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
      const int per_iter = mode == 0 ? 8 : 16, wps = (warps + 3) / 4;
      printf("%s warps=%2d (%d per sub-partition): %.2f cycles per instruction per sub-partition, %.1f FMA/clk/SM\n",
             mode == 0 ? "FFMA2" : "FFMA ", warps, wps, (double)mx / iters / per_iter / wps,
             (double)warps * 32 * iters * 16 / mx);
    }
  cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int W4 : {16, 28})
    for (int threads : {64, 128, 224, 256, 384, 512}) {
      const int K = 200, BP = 100;
      const size_t smem = (size_t)(K * BP + K * W4) * 4;
      for (int rep = 0; rep < 2; ++rep) { k_tile<<<1, threads, smem>>>(K, BP, W4, d_out, d_sink); cudaDeviceSynchronize(); }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < threads / 32; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("tile loop K=200 W4=%d threads=%3d: %lld cycles for 100 k-steps (%.1f cycles per k-step of the slowest warp), err %s\n", W4, threads, mx,
             (double)mx / 100.0, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
