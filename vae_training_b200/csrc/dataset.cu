// dataset.cu -- the three synthetic generators and their score_batch statistics.
// Reference: datasets.py:75-84 (sphere), :183-195 (linear_gaussian), :240-249 (sigmoid);
// score_batch :67-73, :201-206, :255-261.  Operation order mirrors oracle/datasets.py
// (sequential, unfused) so generated batches are bit-identical to the oracle.
#include "common.cuh"

namespace vae {

// One thread per row.  Each needed normal is element (row*g + j) of
// random.normal(key, (size, g)).
__device__ __forceinline__ void dataset_row(const VaeDataDesc& d, uint32_t k0, uint32_t k1,
                                            const float* __restrict__ A, int row, int size,
                                            float* __restrict__ out_row) {
  const int g = d.kind == VAE_DATA_LINEAR_GAUSSIAN ? d.intrinsic_dim : d.dim;
  const uint32_t n = (uint32_t)size * (uint32_t)g;
  const uint32_t base = (uint32_t)row * (uint32_t)g;
  const int width = d.dim + d.padding_dim + (d.kind == VAE_DATA_SIGMOID ? 1 : 0);
  if (d.kind == VAE_DATA_SPHERE) {
    float acc = 0.f;
    for (int j = 0; j < d.dim; ++j) {
      const float v = normal_from_bits(random_bits_at(k0, k1, base + j, n));
      out_row[j] = v;
      const float sq = __fmul_rn(v, v);
      acc = j == 0 ? sq : __fadd_rn(acc, sq);
    }
    const float rad = __fsqrt_rn(acc);
    for (int j = 0; j < d.dim; ++j) out_row[j] = __fdiv_rn(out_row[j], rad);
    for (int j = d.dim; j < width; ++j) out_row[j] = 0.f;
  } else if (d.kind == VAE_DATA_LINEAR_GAUSSIAN) {
    // Y[i] = sum_k A[i,k] X[k]   (datasets.py:186), k sequential.
    for (int i = 0; i < d.dim; ++i) out_row[i] = 0.f;
    for (int k = 0; k < g; ++k) {
      const float xk = normal_from_bits(random_bits_at(k0, k1, base + k, n));
      for (int i = 0; i < d.dim; ++i) {
        const float pr = __fmul_rn(A[i * g + k], xk);
        out_row[i] = k == 0 ? pr : __fadd_rn(out_row[i], pr);
      }
    }
    for (int j = d.dim; j < width; ++j) out_row[j] = 0.f;
  } else {
    float acc = 0.f;
    for (int j = 0; j < d.dim; ++j) {
      const float v = normal_from_bits(random_bits_at(k0, k1, base + j, n));
      out_row[j] = v;
      const float pr = __fmul_rn(v, A[j]);
      acc = j == 0 ? pr : __fadd_rn(acc, pr);
    }
    out_row[d.dim] = det_sigmoid(acc);
    for (int j = d.dim + 1; j < width; ++j) out_row[j] = 0.f;
  }
}

__global__ void dataset_batch_kernel(VaeDataDesc d, const uint32_t* __restrict__ key,
                                     const float* __restrict__ A, int size, float* __restrict__ out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= size) return;
  const int width = d.dim + d.padding_dim + (d.kind == VAE_DATA_SIGMOID ? 1 : 0);
  dataset_row(d, key[0], key[1], A, row, size, out + (int64_t)row * width);
}

// Data-parallel shard (see normal_pairs_kernel): local row i < n_lo is global row lo + i,
// local row n_lo + i is its partner size/2 + lo + i.
__global__ void dataset_pairs_kernel(VaeDataDesc d, const uint32_t* __restrict__ key,
                                     const float* __restrict__ A, int size, int lo, int n_lo,
                                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n_lo) return;
  const int row = i < n_lo ? lo + i : size / 2 + lo + (i - n_lo);
  const int width = d.dim + d.padding_dim + (d.kind == VAE_DATA_SIGMOID ? 1 : 0);
  dataset_row(d, key[0], key[1], A, row, size, out + (int64_t)i * width);
}

// One block; double accumulation (statistics only, tolerance-level parity).
__global__ void dataset_score_kernel(VaeDataDesc d, const float* __restrict__ A,
                                     const float* __restrict__ batch, int size, float* __restrict__ scores) {
  const int width = d.dim + d.padding_dim + (d.kind == VAE_DATA_SIGMOID ? 1 : 0);
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (int r = threadIdx.x; r < size; r += blockDim.x) {
    const float* row = batch + (int64_t)r * width;
    if (d.kind == VAE_DATA_SPHERE) {
      double nr = 0, np = 0;
      for (int j = 0; j < d.dim; ++j) nr += (double)row[j] * row[j];
      for (int j = d.dim; j < width; ++j) np += (double)row[j] * row[j];
      const double e = sqrt(nr) - 1.0;
      s0 += e * e; s1 += np;
    } else if (d.kind == VAE_DATA_LINEAR_GAUSSIAN) {
      double np = 0;
      for (int j = d.dim; j < width; ++j) np += (double)row[j] * row[j];
      s0 += np;
    } else {
      double np = 0, cod = 0;
      for (int j = 0; j < d.dim; ++j) cod += (double)row[j] * A[j];
      for (int j = d.dim + 1; j < width; ++j) np += (double)row[j] * row[j];
      const double hat = row[d.dim];
      s0 += np; s1 += hat; s2 += hat * hat; s3 += cod; s4 += cod * cod;
    }
  }
  __shared__ double sh[5][256];
  sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
  sh[3][threadIdx.x] = s3; sh[4][threadIdx.x] = s4;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5] = {0, 0, 0, 0, 0};
    for (int q = 0; q < 5; ++q)
      for (int i = 0; i < blockDim.x; ++i) t[q] += sh[q][i];
    const double n = (double)size;
    if (d.kind == VAE_DATA_SPHERE) {
      scores[0] = (float)(t[0] / n); scores[1] = (float)(t[1] / n);
    } else if (d.kind == VAE_DATA_LINEAR_GAUSSIAN) {
      scores[0] = (float)(t[0] / n); scores[1] = 0.f;
    } else {
      // datasets.py:256-258 broadcasts (n,) - (n,1): mean over all (i,j) of (hat_j - cod_i)^2.
      scores[0] = (float)(t[0] / n);
      scores[1] = (float)(t[2] / n - 2.0 * (t[1] / n) * (t[3] / n) + t[4] / n);
    }
  }
}

// ---- wide rows (configs[4]: -dd 512): the per-row transform is still sequential (bit-exact
// op order), but the normals are drawn by a fully parallel pass straight into out[:, :dim]
// (row pairs b / b + size/2 share a threefry block), then transformed in place.
__global__ void wide_normals_kernel(const uint32_t* __restrict__ key, float* __restrict__ out, int width, int g,
                                    int size, int lo, int n_lo, int pairs_layout) {
  const uint32_t k0 = key[0], k1 = key[1];
  const int64_t total = (int64_t)n_lo * g;            // blocks: (local low row i, column j)
  const int64_t half = (int64_t)size * g / 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int i = (int)(t / g), j = (int)(t - (int64_t)i * g);
    const int64_t e = (int64_t)(lo + i) * g + j;      // flat element of the low row; partner = e + half
    uint32_t x0 = (uint32_t)e, x1 = (uint32_t)(e + half);
    threefry2x32(k0, k1, x0, x1);
    const int r_lo = pairs_layout ? i : lo + i;
    const int r_hi = pairs_layout ? n_lo + i : size / 2 + lo + i;
    out[(int64_t)r_lo * width + j] = normal_from_bits(x0);
    out[(int64_t)r_hi * width + j] = normal_from_bits(x1);
  }
}
// 32 rows per block (256 threads).  The row reduction keeps the oracle's sequential order (one
// thread per row walks its row left to right), but every global access is coalesced: 32x32
// tiles go through shared memory.
__global__ void __launch_bounds__(256)
wide_transform_kernel(VaeDataDesc d, const float* __restrict__ A, int nrows, float* __restrict__ out) {
  __shared__ float tile[32][33];
  __shared__ float racc[32];
  const int width = d.dim + d.padding_dim + (d.kind == VAE_DATA_SIGMOID ? 1 : 0);
  const int r0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  float acc = 0.f;
  for (int c0 = 0; c0 < d.dim; c0 += 32) {
    for (int rr = wp; rr < 32; rr += 8) {
      const int r = r0 + rr, c = c0 + lane;
      tile[rr][lane] = (r < nrows && c < d.dim) ? out[(int64_t)r * width + c] : 0.f;
    }
    __syncthreads();
    if (wp == 0) {                                   // thread `lane` owns row r0 + lane
      const int cmax = min(32, d.dim - c0);
      for (int j = 0; j < cmax; ++j) {
        const float v = tile[lane][j];
        const float t = d.kind == VAE_DATA_SPHERE ? __fmul_rn(v, v) : __fmul_rn(v, A[c0 + j]);
        acc = (c0 + j) == 0 ? t : __fadd_rn(acc, t);
      }
    }
    __syncthreads();
  }
  if (wp == 0) racc[lane] = d.kind == VAE_DATA_SPHERE ? __fsqrt_rn(acc) : det_sigmoid(acc);
  __syncthreads();
  if (d.kind == VAE_DATA_SPHERE) {
    for (int rr = wp; rr < 32; rr += 8) {
      const int r = r0 + rr;
      if (r >= nrows) continue;
      float* row = out + (int64_t)r * width;
      const float rad = racc[rr];
      for (int c = lane; c < width; c += 32) row[c] = c < d.dim ? __fdiv_rn(row[c], rad) : 0.f;
    }
  } else {      // sigmoid: columns [0, dim) already hold z; column dim = sigmoid(z . A); zeros after
    for (int rr = wp; rr < 32; rr += 8) {
      const int r = r0 + rr;
      if (r >= nrows) continue;
      float* row = out + (int64_t)r * width;
      for (int c = d.dim + lane; c < width; c += 32) row[c] = c == d.dim ? racc[rr] : 0.f;
    }
  }
}
// size even, kind sphere / sigmoid, dim >= 32: rows [lo, hi) + partners, either densely packed
// (pairs_layout) or at their global positions (lo = 0, hi = size/2)
static int wide_batch(const VaeDataDesc* d, const uint32_t* d_key, const float* d_A, int size, int lo, int hi,
                      int pairs_layout, float* d_out, cudaStream_t s) {
  const int n_lo = hi - lo, width = d->dim + d->padding_dim + (d->kind == VAE_DATA_SIGMOID ? 1 : 0);
  const int64_t total = (int64_t)n_lo * d->dim;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  wide_normals_kernel<<<(int)blocks, 256, 0, s>>>(d_key, d_out, width, d->dim, size, lo, n_lo, pairs_layout);
  VAE_LAUNCH_OK("wide_normals_kernel");
  wide_transform_kernel<<<(2 * n_lo + 31) / 32, 256, 0, s>>>(*d, d_A, 2 * n_lo, d_out);
  VAE_LAUNCH_OK("wide_transform_kernel");
  return 0;
}
static inline bool use_wide(const VaeDataDesc* d, int size) {
  return d->kind != VAE_DATA_LINEAR_GAUSSIAN && d->dim >= 32 && size % 2 == 0 && size >= 2;
}

int check_data_desc(const VaeDataDesc* d) {
  if (!d) { set_error("data desc is NULL"); return VAE_ERR_INVALID; }
  if (d->kind < 0 || d->kind > 2) { set_error("unknown dataset kind %d", d->kind); return VAE_ERR_INVALID; }
  if (d->dim <= 0 || d->padding_dim < 0) { set_error("dataset dim must be > 0 and padding >= 0"); return VAE_ERR_INVALID; }
  if (d->kind == VAE_DATA_LINEAR_GAUSSIAN && d->intrinsic_dim <= 0) { set_error("intrinsic_dim must be > 0"); return VAE_ERR_INVALID; }
  return 0;
}

}  // namespace vae

extern "C" {

int vae_dataset_batch(const VaeDataDesc* d, const uint32_t* d_key, const float* d_A,
                      int32_t size, float* d_out, void* stream) {
  int rc = vae::check_data_desc(d);
  if (rc) return rc;
  VAE_REQUIRE(d_key && size >= 0, "vae_dataset_batch: bad arguments");
  VAE_REQUIRE(d->kind == VAE_DATA_SPHERE || d_A, "vae_dataset_batch: generator matrix A is NULL");
  if (size == 0) return 0;
  VAE_REQUIRE(d_out, "vae_dataset_batch: d_out is NULL");
  VAE_REQUIRE((int64_t)size * vae_data_normals_per_row(d) < (1ll << 32), "vae_dataset_batch: too many draws");
  if (vae::use_wide(d, size)) return vae::wide_batch(d, d_key, d_A, size, 0, size / 2, 0, d_out, (cudaStream_t)stream);
  vae::dataset_batch_kernel<<<(size + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*d, d_key, d_A, size, d_out);
  VAE_LAUNCH_OK("dataset_batch_kernel");
  return 0;
}

int vae_dataset_batch_pairs(const VaeDataDesc* d, const uint32_t* d_key, const float* d_A,
                            int32_t size, int32_t lo, int32_t hi, float* d_out, void* stream) {
  int rc = vae::check_data_desc(d);
  if (rc) return rc;
  VAE_REQUIRE(d_key && size >= 0 && size % 2 == 0, "vae_dataset_batch_pairs: size must be even and >= 0");
  VAE_REQUIRE(lo >= 0 && lo <= hi && hi <= size / 2, "vae_dataset_batch_pairs: bad row range");
  VAE_REQUIRE(d->kind == VAE_DATA_SPHERE || d_A, "vae_dataset_batch_pairs: generator matrix A is NULL");
  const int n_lo = hi - lo;
  if (n_lo == 0) return 0;
  VAE_REQUIRE(d_out, "vae_dataset_batch_pairs: d_out is NULL");
  VAE_REQUIRE((int64_t)size * vae_data_normals_per_row(d) < (1ll << 32), "vae_dataset_batch_pairs: too many draws");
  if (vae::use_wide(d, size)) return vae::wide_batch(d, d_key, d_A, size, lo, hi, 1, d_out, (cudaStream_t)stream);
  vae::dataset_pairs_kernel<<<(2 * n_lo + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*d, d_key, d_A, size, lo,
                                                                                      n_lo, d_out);
  VAE_LAUNCH_OK("dataset_pairs_kernel");
  return 0;
}

int vae_dataset_score(const VaeDataDesc* d, const float* d_A, const float* d_batch,
                      int32_t size, float* d_scores, void* stream) {
  int rc = vae::check_data_desc(d);
  if (rc) return rc;
  VAE_REQUIRE(d_batch && d_scores && size > 0, "vae_dataset_score: bad arguments");
  VAE_REQUIRE(d->kind != VAE_DATA_SIGMOID || d_A, "vae_dataset_score: generator matrix A is NULL");
  vae::dataset_score_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*d, d_A, d_batch, size, d_scores);
  VAE_LAUNCH_OK("dataset_score_kernel");
  return 0;
}

}  // extern "C"
