#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_store(float* p, int iters, long long* out, int mode) {
  float* base = p + (size_t)blockIdx.x * (1 << 20);
  float v[8]; for (int j = 0; j < 8; ++j) v[j] = threadIdx.x + j;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    size_t off = ((size_t)(it % 64) * blockDim.x + threadIdx.x) * 8;
    if (mode == 0) asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(base + off), "f"(v[0]),"f"(v[1]),"f"(v[2]),"f"(v[3]),"f"(v[4]),"f"(v[5]),"f"(v[6]),"f"(v[7]) : "memory");
    else if (mode == 1) { *reinterpret_cast<float4*>(base + off) = make_float4(v[0],v[1],v[2],v[3]); *reinterpret_cast<float4*>(base + off + 4) = make_float4(v[4],v[5],v[6],v[7]); }
    else { for (int j = 0; j < 8; ++j) base[(size_t)(it % 64) * blockDim.x * 8 + j * blockDim.x + threadIdx.x] = v[j]; }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
__global__ void k_load(const float* p, int iters, long long* out, float* sink) {
  const float* base = p + (size_t)blockIdx.x * (1 << 20);
  float acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    size_t off = ((size_t)(it % 64) * blockDim.x + threadIdx.x) * 8;
    float v[8];
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]),"=f"(v[1]),"=f"(v[2]),"=f"(v[3]),"=f"(v[4]),"=f"(v[5]),"=f"(v[6]),"=f"(v[7]) : "l"(base + off));
    acc += v[0] + v[7];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  float* p; long long* out; float* sink;
  cudaMalloc(&p, (size_t)148 * (1 << 20) * 4); cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  cudaMemset(p, 0, (size_t)148 * (1 << 20) * 4);
  long long h[148];
  for (int blocks : {1, 148}) for (int mode = 0; mode < 3; ++mode) {
    int iters = 2048;
    k_store<<<blocks, 512>>>(p, iters, out, mode); cudaDeviceSynchronize();
    k_store<<<blocks, 512>>>(p, iters, out, mode); cudaDeviceSynchronize();
    cudaMemcpy(h, out, blocks * 8, cudaMemcpyDeviceToHost);
    double bytes = (double)iters * 512 * 32;
    printf("store mode %d blocks %d: %.1f B/clk per SM (%lld cycles)\n", mode, blocks, bytes / h[0], h[0]);
  }
  for (int blocks : {1, 148}) {
    int iters = 2048;
    k_load<<<blocks, 512>>>(p, iters, out, sink); cudaDeviceSynchronize();
    k_load<<<blocks, 512>>>(p, iters, out, sink); cudaDeviceSynchronize();
    cudaMemcpy(h, out, blocks * 8, cudaMemcpyDeviceToHost);
    printf("load v8 blocks %d: %.1f B/clk per SM\n", blocks, (double)iters * 512 * 32 / h[0]);
  }
  return 0;
}
