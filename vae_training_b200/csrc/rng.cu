// rng.cu -- jax.random replacements: split, raw bits, normal (see common.cuh).
// Reference call sites: vae.py:56,125; model.py:140,154,208,227; datasets.py:51,76,185,242.
#include "common.cuh"

namespace vae {

__global__ void split_kernel(const uint32_t* __restrict__ key, uint32_t* __restrict__ out, int num) {
  // split(key, num) = threefry(key, iota(2*num)).reshape(num, 2); block j -> (j, j+num).
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num) return;
  uint32_t x0 = (uint32_t)j, x1 = (uint32_t)(j + num);
  threefry2x32(key[0], key[1], x0, x1);
  out[j] = x0;
  out[j + num] = x1;
}

// Thread j owns block j: counters (j, j+half) -> out[j], out[j+half].  Both stores
// are fully coalesced.
template <bool kNormal>
__global__ void bits_kernel(const uint32_t* __restrict__ key, void* __restrict__ out_, int64_t n) {
  const uint32_t k0 = key[0], k1 = key[1];
  const int64_t half = (n + 1) >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += stride) {
    uint32_t x0 = (uint32_t)j;
    const bool has1 = j + half < n;
    uint32_t x1 = has1 ? (uint32_t)(j + half) : 0u;
    threefry2x32(k0, k1, x0, x1);
    if (kNormal) {
      float* out = (float*)out_;
      out[j] = normal_from_bits(x0);
      if (has1) out[j + half] = normal_from_bits(x1);
    } else {
      uint32_t* out = (uint32_t*)out_;
      out[j] = x0;
      if (has1) out[j + half] = x1;
    }
  }
}

__global__ void normal_rows_kernel(const uint32_t* __restrict__ key, float* __restrict__ out,
                                   int64_t n, int64_t first, int64_t count) {
  const uint32_t k0 = key[0], k1 = key[1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride)
    out[t] = normal_from_bits(random_bits_at(k0, k1, (uint32_t)(first + t), (uint32_t)n));
}

// Data-parallel shard of normal(key, (rows, cols)), rows even: the shard owns global rows
// [lo, hi) and their threefry partners [lo + rows/2, hi + rows/2) (block j of the flat
// array yields elements j and j + rows*cols/2, i.e. the same column of row b and of row
// b + rows/2).  Output: (2*(hi-lo), cols), first the low rows then the partner rows.  Every
// threefry block computed is fully used.
__global__ void normal_pairs_kernel(const uint32_t* __restrict__ key, float* __restrict__ out,
                                    int64_t half, int64_t first, int64_t count) {
  const uint32_t k0 = key[0], k1 = key[1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    uint32_t x0 = (uint32_t)(first + t), x1 = (uint32_t)(first + t + half);
    threefry2x32(k0, k1, x0, x1);
    out[t] = normal_from_bits(x0);
    out[t + count] = normal_from_bits(x1);
  }
}

static inline int grid_for(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  const int64_t cap = 148 * 16;   // multiples of the SM count; grid-stride beyond that
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace vae

extern "C" {

int vae_random_split(const uint32_t* d_key, uint32_t* d_out, int32_t num, void* stream) {
  VAE_REQUIRE(d_key && d_out && num > 0, "vae_random_split: bad arguments");
  vae::split_kernel<<<(num + 63) / 64, 64, 0, (cudaStream_t)stream>>>(d_key, d_out, num);
  VAE_LAUNCH_OK("split_kernel");
  return 0;
}

int vae_random_bits(const uint32_t* d_key, uint32_t* d_out, int64_t n, void* stream) {
  VAE_REQUIRE(d_key && n >= 0 && n < (1ll << 32), "vae_random_bits: n must be in [0, 2^32)");
  if (n == 0) return 0;
  VAE_REQUIRE(d_out, "vae_random_bits: d_out is NULL");
  vae::bits_kernel<false><<<vae::grid_for((n + 1) / 2, 256), 256, 0, (cudaStream_t)stream>>>(d_key, d_out, n);
  VAE_LAUNCH_OK("bits_kernel");
  return 0;
}

int vae_random_normal(const uint32_t* d_key, float* d_out, int64_t n, void* stream) {
  VAE_REQUIRE(d_key && n >= 0 && n < (1ll << 32), "vae_random_normal: n must be in [0, 2^32)");
  if (n == 0) return 0;
  VAE_REQUIRE(d_out, "vae_random_normal: d_out is NULL");
  vae::bits_kernel<true><<<vae::grid_for((n + 1) / 2, 256), 256, 0, (cudaStream_t)stream>>>(d_key, d_out, n);
  VAE_LAUNCH_OK("normal_kernel");
  return 0;
}

int vae_random_normal_rows(const uint32_t* d_key, float* d_out, int64_t rows, int64_t cols,
                           int64_t row_begin, int64_t row_end, void* stream) {
  VAE_REQUIRE(d_key && rows >= 0 && cols >= 0 && rows * cols < (1ll << 32), "vae_random_normal_rows: bad shape");
  VAE_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= rows, "vae_random_normal_rows: bad row range");
  const int64_t count = (row_end - row_begin) * cols;
  if (count == 0) return 0;
  VAE_REQUIRE(d_out, "vae_random_normal_rows: d_out is NULL");
  vae::normal_rows_kernel<<<vae::grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(
      d_key, d_out, rows * cols, row_begin * cols, count);
  VAE_LAUNCH_OK("normal_rows_kernel");
  return 0;
}

int vae_random_normal_pairs(const uint32_t* d_key, float* d_out, int64_t rows, int64_t cols,
                            int64_t lo, int64_t hi, void* stream) {
  VAE_REQUIRE(d_key && rows >= 0 && cols >= 0 && rows * cols < (1ll << 32), "vae_random_normal_pairs: bad shape");
  VAE_REQUIRE(rows % 2 == 0, "vae_random_normal_pairs: rows must be even (got %lld)", (long long)rows);
  VAE_REQUIRE(lo >= 0 && lo <= hi && hi <= rows / 2, "vae_random_normal_pairs: bad row range");
  const int64_t count = (hi - lo) * cols;
  if (count == 0) return 0;
  VAE_REQUIRE(d_out, "vae_random_normal_pairs: d_out is NULL");
  vae::normal_pairs_kernel<<<vae::grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(
      d_key, d_out, rows * cols / 2, lo * cols, count);
  VAE_LAUNCH_OK("normal_pairs_kernel");
  return 0;
}

}  // extern "C"
