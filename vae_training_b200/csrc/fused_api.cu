// fused_api.cu -- the fused-training entry points of the C ABI: dispatch between the
// single-CTA kernel (fused_small.cu, class 1) and the cluster kernel (fused_mlp.cu, class 2).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#ifndef VAE_MLP_TC_MIN_JOBS
#define VAE_MLP_TC_MIN_JOBS 16     // smaller packs stay on the FFMA cluster kernel (8 - 16 CTAs per model: 148 - 170 us/step vs 185; profiles/r02_mlp_kernel_crossover.txt)
#endif

namespace vae {
// fused_small.cu
int small_supported(const VaeTrainJob* job);
int small_ctas_per_sm(const VaeTrainJob* job);
size_t small_job_bytes();
int small_prepare(const VaeTrainJob* jobs, int32_t n_jobs, void* d_scratch, cudaStream_t s, VaePackPlan* plan);
int small_launch(const VaePackPlan* plan, int64_t n_steps, cudaStream_t s);
// fused_mlp.cu
struct MlpJobDev;
int mlp_supported(const VaeTrainJob* j, size_t* smem, size_t* ws_bytes, size_t* smem4 = nullptr);
int mlp_fill_job(const VaeTrainJob& j, MlpJobDev* o);
size_t mlp_job_bytes();
int mlp_prepare(int n_jobs, size_t smem, size_t smem4, VaePackPlan* plan);
int mlp_launch(const VaePackPlan* plan, int64_t n_steps, cudaStream_t s);
int mlp_resident_jobs(size_t smem, size_t smem4, int* out);
// fused_mlp_tc.cu (tensor-core kernel for the same class-2 jobs; one CTA per model)
int mlp_tc_supported(const VaeTrainJob* j, size_t* ws_bytes);
size_t mlp_tc_job_bytes();
int mlp_tc_fill_job(const VaeTrainJob& j, void* out, int csize);
int mlp_tc_prepare(int n_jobs, int* cluster);
int mlp_tc_launch(const VaePackPlan* plan, int64_t n_steps, cudaStream_t s);

// Which kernel runs a class-2 pack.  The tensor-core kernel runs one model per CTA (throughput: packs that fill the
// GPU) or per cluster of 2 / 4 / 8 CTAs (latency: small packs, e.g. the two sphere chains per GPU of the 8-GPU sweep);
// the FFMA cluster kernel of fused_mlp.cu stays as the fallback for shapes the tensor-core kernel does not take
// (batch > 128, sigmoid decoders, layers wider than 200).  VAE_MLP_TC=1 / 0 forces one or the other (A/B runs).
constexpr int kMlpTcMinJobs = VAE_MLP_TC_MIN_JOBS;
static bool use_mlp_tc(const VaeTrainJob* j, size_t* ws, int n_jobs) {
  const char* e = getenv("VAE_MLP_TC");
  if (e && e[0] == '0') return false;
  if (!(e && e[0] == '1') && n_jobs < kMlpTcMinJobs) return false;
  return mlp_tc_supported(j, ws) != 0;
}
}  // namespace vae

extern "C" {

int vae_train_fused_supported(const VaeTrainJob* job) {
  if (!job) return 0;
  if (vae::small_supported(job)) return 1;
  return vae::mlp_supported(job, nullptr, nullptr);
}

size_t vae_train_fused_workspace_bytes(const VaeTrainJob* job) {
  size_t ws = 0, wt = 0;
  if (job && !vae::small_supported(job) && vae::mlp_supported(job, nullptr, &ws)) {
    if (vae::mlp_tc_supported(job, &wt) && wt > ws) ws = wt;      // either kernel may be chosen when the pack is prepared
    return ws;
  }
  return 0;
}

int vae_train_fused_ctas_per_sm(const VaeTrainJob* job) {
  VAE_REQUIRE(job, "vae_train_fused_ctas_per_sm: job is NULL");
  if (vae::small_supported(job)) return vae::small_ctas_per_sm(job);
  return 0;
}

int vae_train_fused_resident_jobs(const VaeTrainJob* job) {
  VAE_REQUIRE(job, "vae_train_fused_resident_jobs: job is NULL");
  if (vae::small_supported(job)) {
    const int per_sm = vae::small_ctas_per_sm(job);
    if (per_sm <= 0) return per_sm;
    int dev = 0, sms = 0;
    VAE_CUDA_OK(cudaGetDevice(&dev));
    VAE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return per_sm * sms;
  }
  size_t smem = 0, smem4 = 0;
  if (!vae::mlp_supported(job, &smem, nullptr, &smem4)) return 0;
  if (vae::use_mlp_tc(job, nullptr, 1 << 20)) {                     // a GPU-filling pack runs on the tensor-core kernel: one CTA (= model) per SM
    int dev = 0, sms = 0;
    VAE_CUDA_OK(cudaGetDevice(&dev));
    VAE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return sms;
  }
  int out = 0;
  const int rc = vae::mlp_resident_jobs(smem, smem4, &out);
  return rc ? rc : out;
}

size_t vae_train_fused_scratch_bytes(int32_t n_jobs) {
  size_t per = vae::small_job_bytes() > vae::mlp_job_bytes() ? vae::small_job_bytes() : vae::mlp_job_bytes();
  if (vae::mlp_tc_job_bytes() > per) per = vae::mlp_tc_job_bytes();
  return n_jobs > 0 ? (size_t)n_jobs * per : 0;
}

int vae_pack_prepare(const VaeTrainJob* jobs, int32_t n_jobs, void* d_scratch, size_t scratch_bytes,
                     VaePackPlan* plan, void* stream) {
  VAE_REQUIRE(jobs && n_jobs > 0 && plan, "vae_pack_prepare: bad arguments");
  VAE_REQUIRE(d_scratch && scratch_bytes >= vae_train_fused_scratch_bytes(n_jobs),
              "vae_pack_prepare: scratch too small (need %zu bytes)", vae_train_fused_scratch_bytes(n_jobs));
  cudaStream_t s = (cudaStream_t)stream;
  memset(plan, 0, sizeof(*plan));
  const int cls = vae_train_fused_supported(&jobs[0]);
  for (int i = 0; i < n_jobs; ++i) {
    const int c = vae_train_fused_supported(&jobs[i]);
    if (c == 0) {
      vae::set_error("vae_pack_prepare: job %d has no fused kernel (D=%d L=%d hidden=%d/%d)", i, jobs[i].desc.data_dim,
                     jobs[i].desc.latent_dim, jobs[i].desc.n_enc_hidden, jobs[i].desc.n_dec_hidden);
      return VAE_ERR_UNSUPPORTED;
    }
    VAE_REQUIRE(c == cls, "vae_pack_prepare: jobs %d and 0 belong to different fused-kernel classes (%d vs %d)", i, c, cls);
  }
  int dev = 0;
  VAE_CUDA_OK(cudaGetDevice(&dev));
  int rc;
  if (cls == 1) {
    if ((rc = vae::small_prepare(jobs, n_jobs, d_scratch, s, plan))) return rc;
  } else {
    // class 2 on the tensor cores: one CTA per model (every job must qualify, else the FFMA cluster kernel runs the pack)
    bool tc = true;
    for (int i = 0; i < n_jobs && tc; ++i) tc = vae::use_mlp_tc(&jobs[i], nullptr, n_jobs);
    if (tc) {
      int tcs = 1;
      if ((rc = vae::mlp_tc_prepare(n_jobs, &tcs))) return rc;
      const size_t tjb = vae::mlp_tc_job_bytes();
      char* th = (char*)malloc((size_t)n_jobs * tjb);
      VAE_REQUIRE(th, "vae_pack_prepare: out of host memory");
      for (int i = 0; i < n_jobs; ++i) {
        const VaeTrainJob& j = jobs[i];
        if (!j.params || !j.adam_m || !j.adam_v || !j.step || !j.model_key || !j.dataset_key || !j.loss_out ||
            !j.workspace || (j.data.kind != VAE_DATA_SPHERE && !j.A)) {
          free(th);
          vae::set_error("vae_pack_prepare: job %d has a NULL device pointer (class 2 needs a zero-filled workspace of "
                         "vae_train_fused_workspace_bytes() bytes)", i);
          return VAE_ERR_INVALID;
        }
        rc = vae::mlp_tc_fill_job(j, th + (size_t)i * tjb, tcs);
        if (rc) { free(th); return rc; }
      }
      cudaError_t e = cudaMemcpyAsync(d_scratch, th, (size_t)n_jobs * tjb, cudaMemcpyHostToDevice, s);
      free(th);
      VAE_CUDA_OK(e);
      plan->cls = 2; plan->n_jobs = n_jobs; plan->d_jobs = d_scratch; plan->cluster = tcs; plan->tc = 1; plan->smem_bytes = 0;
      plan->abi = VAE_B200_ABI_VERSION;
      plan->device = dev;
      return 0;
    }
    // class 2: one cluster per model
    const size_t jb = vae::mlp_job_bytes();
    char* host = (char*)malloc((size_t)n_jobs * jb);
    VAE_REQUIRE(host, "vae_pack_prepare: out of host memory");
    size_t smem = 0, smem4 = 0;
    bool ok4 = true;
    for (int i = 0; i < n_jobs; ++i) {
      const VaeTrainJob& j = jobs[i];
      size_t sb = 0, wb = 0, s4 = 0;
      vae::mlp_supported(&j, &sb, &wb, &s4);
      if (s4 == 0) ok4 = false; else if (s4 > smem4) smem4 = s4;
      if (!j.params || !j.adam_m || !j.adam_v || !j.step || !j.model_key || !j.dataset_key || !j.loss_out ||
          !j.workspace || (j.data.kind != VAE_DATA_SPHERE && !j.A)) {
        free(host);
        vae::set_error("vae_pack_prepare: job %d has a NULL device pointer (class 2 needs a workspace of "
                       "vae_train_fused_workspace_bytes() = %zu bytes)", i, wb);
        return VAE_ERR_INVALID;
      }
      if (sb > smem) smem = sb;
      rc = vae::mlp_fill_job(j, (vae::MlpJobDev*)(host + (size_t)i * jb));
      if (rc) { free(host); return rc; }
    }
    cudaError_t e = cudaMemcpyAsync(d_scratch, host, (size_t)n_jobs * jb, cudaMemcpyHostToDevice, s);
    free(host);
    VAE_CUDA_OK(e);
    plan->cls = 2; plan->n_jobs = n_jobs; plan->d_jobs = d_scratch;
    if ((rc = vae::mlp_prepare(n_jobs, smem, ok4 ? smem4 : 0, plan))) return rc;
  }
  plan->abi = VAE_B200_ABI_VERSION;
  plan->device = dev;
  return 0;
}

int vae_pack_train(const VaePackPlan* plan, int64_t n_steps, void* stream) {
  VAE_REQUIRE(plan && plan->abi == VAE_B200_ABI_VERSION && plan->d_jobs && plan->n_jobs > 0 && n_steps >= 0,
              "vae_pack_train: plan was not produced by vae_pack_prepare of this library version");
  if (n_steps == 0) return 0;
  if (plan->cls == 1) return vae::small_launch(plan, n_steps, (cudaStream_t)stream);
  if (plan->cls == 2 && plan->tc) return vae::mlp_tc_launch(plan, n_steps, (cudaStream_t)stream);
  if (plan->cls == 2) return vae::mlp_launch(plan, n_steps, (cudaStream_t)stream);
  vae::set_error("vae_pack_train: bad plan class %d", plan->cls);
  return VAE_ERR_INVALID;
}

int vae_train_fused(const VaeTrainJob* jobs, int32_t n_jobs, int64_t n_steps, void* d_scratch,
                    size_t scratch_bytes, void* stream) {
  VAE_REQUIRE(jobs && n_jobs > 0 && n_steps >= 0, "vae_train_fused: bad arguments");
  if (n_steps == 0) return 0;
  VaePackPlan plan;
  const int rc = vae_pack_prepare(jobs, n_jobs, d_scratch, scratch_bytes, &plan, stream);
  return rc ? rc : vae_pack_train(&plan, n_steps, stream);
}

}  // extern "C"
