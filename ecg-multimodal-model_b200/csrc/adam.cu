// Multi-tensor Adam (reference: torch.optim.Adam defaults used by train.py:43, train_kfold.py:42,
// signal_model.py:157 -- betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).
// One launch updates every parameter of a param group: the host passes a device table of
// chunks (pointer quadruple + length); CTA i walks chunk i.  28 algorithmic bytes per parameter
// (read p, g, m, v; write p, m, v).
#include "common.h"

namespace ecgmm {

struct AdamChunk {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_over_bc1, float b1,
                                         float b2, float eps, float inv_sqrt_bc2, float gscale, float wd) {
  g *= gscale;
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = fmaf(b1, m, (1.f - b1) * g);
  v = fmaf(b2, v, (1.f - b2) * g * g);
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p -= lr_over_bc1 * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamChunk* __restrict__ chunks, float lr_over_bc1,
                                                    float b1, float b2, float eps, float inv_sqrt_bc2, float gscale,
                                                    float wd) {
  const AdamChunk c = chunks[blockIdx.x];
  const bool vec = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) |
                     reinterpret_cast<uintptr_t>(c.m) | reinterpret_cast<uintptr_t>(c.v)) & 15) == 0;
  long long i0 = 0;
  if (vec) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 p = reinterpret_cast<float4*>(c.p)[i];
      const float4 g = reinterpret_cast<const float4*>(c.g)[i];
      float4 m = reinterpret_cast<float4*>(c.m)[i];
      float4 v = reinterpret_cast<float4*>(c.v)[i];
      adam_one(p.x, g.x, m.x, v.x, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      adam_one(p.y, g.y, m.y, v.y, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      adam_one(p.z, g.z, m.z, v.z, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      adam_one(p.w, g.w, m.w, v.w, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      reinterpret_cast<float4*>(c.p)[i] = p;
      reinterpret_cast<float4*>(c.m)[i] = m;
      reinterpret_cast<float4*>(c.v)[i] = v;
    }
    i0 = n4 << 2;
  }
  for (long long i = i0 + threadIdx.x; i < c.n; i += blockDim.x) {
    float p = c.p[i], m = c.m[i], v = c.v[i];
    adam_one(p, c.g[i], m, v, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
    c.p[i] = p;
    c.m[i] = m;
    c.v[i] = v;
  }
}

// Device-resident step state for CUDA-graph replay (ecgmm.graph): the optimizer step count and a dropout seed
// offset live in device memory and are advanced by a kernel INSIDE the captured graph, so every replay sees new values
// although all kernel arguments are frozen.
__global__ void step_advance_kernel(long long* state) {
  state[0] += 1;                                            // Adam step count
  state[1] = (long long)((unsigned long long)state[1] + 0x9E3779B97F4A7C15ull);  // dropout seed offset
}

__global__ void __launch_bounds__(256) adam_dev_kernel(const AdamChunk* __restrict__ chunks,
                                                        const float* __restrict__ lr_dev, float b1, float b2, float eps,
                                                        const long long* __restrict__ step_dev, float gscale, float wd) {
  __shared__ float coef[2];
  if (threadIdx.x == 0) {
    const double step = (double)*step_dev;
    const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
    coef[0] = (float)((double)*lr_dev / bc1);
    coef[1] = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float lr_over_bc1 = coef[0], inv_sqrt_bc2 = coef[1];
  const AdamChunk c = chunks[blockIdx.x];
  const bool vec = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) |
                     reinterpret_cast<uintptr_t>(c.m) | reinterpret_cast<uintptr_t>(c.v)) & 15) == 0;
  long long i0 = 0;
  if (vec) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 p = reinterpret_cast<float4*>(c.p)[i];
      const float4 g = reinterpret_cast<const float4*>(c.g)[i];
      float4 m = reinterpret_cast<float4*>(c.m)[i];
      float4 v = reinterpret_cast<float4*>(c.v)[i];
      adam_one(p.x, g.x, m.x, v.x, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      adam_one(p.y, g.y, m.y, v.y, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      adam_one(p.z, g.z, m.z, v.z, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      adam_one(p.w, g.w, m.w, v.w, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
      reinterpret_cast<float4*>(c.p)[i] = p;
      reinterpret_cast<float4*>(c.m)[i] = m;
      reinterpret_cast<float4*>(c.v)[i] = v;
    }
    i0 = n4 << 2;
  }
  for (long long i = i0 + threadIdx.x; i < c.n; i += blockDim.x) {
    float p = c.p[i], m = c.m[i], v = c.v[i];
    adam_one(p, c.g[i], m, v, lr_over_bc1, b1, b2, eps, inv_sqrt_bc2, gscale, wd);
    c.p[i] = p;
    c.m[i] = m;
    c.v[i] = v;
  }
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_step_advance(long long* state, void* stream) {
  ECGMM_CHECK(state, ECGMM_ERR_ARG, "step_advance: null pointer");
  step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
  return check_launch("step_advance_kernel");
}

extern "C" int ecgmm_adam_step_dev(const void* chunk_table, int n_chunks, const float* lr_dev, float beta1, float beta2,
                                   float eps, float weight_decay, const long long* step_dev, float grad_scale,
                                   void* stream) {
  ECGMM_CHECK(chunk_table || n_chunks == 0, ECGMM_ERR_ARG, "adam_step_dev: null chunk table");
  ECGMM_CHECK(lr_dev && step_dev, ECGMM_ERR_ARG, "adam_step_dev: null lr / step pointer");
  if (n_chunks == 0) return ECGMM_OK;
  adam_dev_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AdamChunk*>(chunk_table), lr_dev,
                                                             beta1, beta2, eps, step_dev, grad_scale, weight_decay);
  return check_launch("adam_dev_kernel");
}

extern "C" int ecgmm_adam_chunk_bytes(void) { return (int)sizeof(AdamChunk); }

extern "C" int ecgmm_adam_step(const void* chunk_table, int n_chunks, float lr, float beta1, float beta2, float eps,
                               float weight_decay, long long step, float grad_scale, void* stream) {
  ECGMM_CHECK(chunk_table || n_chunks == 0, ECGMM_ERR_ARG, "adam_step: null chunk table");
  ECGMM_CHECK(step >= 1, ECGMM_ERR_ARG, "adam_step: step must be >= 1");
  if (n_chunks == 0) return ECGMM_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AdamChunk*>(chunk_table),
                                                         (float)(lr / bc1), beta1, beta2, eps,
                                                         (float)(1.0 / sqrt(bc2)), grad_scale, weight_decay);
  return check_launch("adam_kernel");
}
