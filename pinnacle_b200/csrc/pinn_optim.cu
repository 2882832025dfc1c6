// pinn_optim.cu -- optimizer-side kernels over the flat parameter / gradient buffers.
//
// The fused op hands back one flat gradient row per loss term (P ~ 41 k floats).  What the reference does next is a
// few dozen tiny launches and host synchronisations per step: torch.optim.Adam's foreach update over 12 tensors
// (benchmark.py:107), and -- for the learning-rate-annealing optimizer (src/optimizer/lr_adaptor.py:33-62) -- one
// backward pass per boundary term, each followed by a max|g| / sum|g| reduction and an `.item()`.  Here each of these
// is one launch over the flat buffers, with every scalar staying on the device.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/pinn_b200.h"

namespace pinn {
void set_error_string(const std::string& s);   // pinn_api.cu
}

namespace {

constexpr int kStatBlocks = 64;      // partial-reduction blocks of the statistics kernels (fixed => deterministic order)
constexpr int kStatThreads = 256;
constexpr int kMaxTerms = 16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Adam (torch.optim.Adam, amsgrad = False, maximize = False; L2 weight decay folded into the gradient) ----------
// step_dev: device counter holding the number of updates applied so far; every thread reads it, the LAST thread of the
// grid to finish increments it (ticket in step_dev[1]) so the kernel is CUDA-graph replayable with no host state.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float* step_dev, float lr, float b1, float b2, float eps, float wd) {
  const float t = step_dev[0] + 1.0f;
  const float bc1 = 1.0f - powf(b1, t);
  const float bc2 = 1.0f - powf(b2, t);
  const float step_size = lr / bc1;
  const float rs_bc2 = sqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.0f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(1.0f - b1, gi - m[i], m[i]);              // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / rs_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* ticket = reinterpret_cast<unsigned*>(step_dev + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      step_dev[0] = t;
    }
  }
}

// ---- statistics of a flat vector: out = { max|g|, sum|g|, sum g^2 } ------------------------------------------------
__global__ void stats_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ part /*[blocks][3]*/) {
  float mx = 0.f, s1 = 0.f, s2 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = fabsf(g[i]);
    mx = fmaxf(mx, a);
    s1 += a;
    s2 = fmaf(a, a, s2);
  }
  __shared__ float sm[3][kStatThreads / 32];
  mx = warp_max(mx); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = mx; sm[1][threadIdx.x >> 5] = s1; sm[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kStatThreads / 32; ++w) { mx = fmaxf(mx, sm[0][w]); s1 += sm[1][w]; s2 += sm[2][w]; }
    part[blockIdx.x * 3 + 0] = mx; part[blockIdx.x * 3 + 1] = s1; part[blockIdx.x * 3 + 2] = s2;
  }
}
__global__ void stats_final_kernel(const float* __restrict__ part, int blocks, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    float mx = 0.f, s1 = 0.f, s2 = 0.f;
    for (int b = 0; b < blocks; ++b) { mx = fmaxf(mx, part[b * 3]); s1 += part[b * 3 + 1]; s2 += part[b * 3 + 2]; }
    out[0] = mx; out[1] = s1; out[2] = s2;
  }
}

// ---- learning-rate annealing (lr_adaptor.py:33-62) on the per-term gradient rows ---------------------------------------
// G: [n_terms][ldg] gradients of the UNWEIGHTED loss terms.  Pass 1: per block, max_i |sum_{k<n_pde} G[k][i]| and, per
// boundary term t >= n_pde, sum_i |G[t][i]|.
__global__ void lra_partial_kernel(const float* __restrict__ G, int n_terms, int n_pde, int64_t P, int64_t ldg, float* __restrict__ part /*[blocks][kMaxTerms]*/) {
  float acc[kMaxTerms];
#pragma unroll
  for (int t = 0; t < kMaxTerms; ++t) acc[t] = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float r = 0.f;
    for (int k = 0; k < n_pde; ++k) r += G[k * ldg + i];
    acc[0] = fmaxf(acc[0], fabsf(r));
#pragma unroll
    for (int t = 1; t < kMaxTerms; ++t)
      if (n_pde - 1 + t < n_terms) acc[t] += fabsf(G[(int64_t)(n_pde - 1 + t) * ldg + i]);
  }
  __shared__ float sm[kMaxTerms][kStatThreads / 32];
#pragma unroll
  for (int t = 0; t < kMaxTerms; ++t) {
    const float r = (t == 0) ? warp_max(acc[t]) : warp_sum(acc[t]);
    if ((threadIdx.x & 31) == 0) sm[t][threadIdx.x >> 5] = r;
  }
  __syncthreads();
  if (threadIdx.x < kMaxTerms) {
    const int t = threadIdx.x;
    float r = sm[t][0];
    for (int w = 1; w < kStatThreads / 32; ++w) r = (t == 0) ? fmaxf(r, sm[t][w]) : r + sm[t][w];
    part[blockIdx.x * kMaxTerms + t] = r;
  }
}
// Pass 2: every block reduces the partials in the same fixed order, forms the new weights
//   w_t <- (1 - alpha) w_t + alpha * max|g_r| / (mean|g_t| * w_t),  t >= n_pde      (lr_adaptor.py:55-57)
// (block 0 stores them) and writes its slice of  grad_out = sum_t w_t G[t].
__global__ void lra_apply_kernel(const float* __restrict__ G, int n_terms, int n_pde, int64_t P, int64_t ldg, const float* __restrict__ part, int blocks,
                                 const float* __restrict__ w_in, const float* __restrict__ counts, float alpha, float* __restrict__ w_out,
                                 float* __restrict__ grad_out) {
  __shared__ float w[kMaxTerms + PINN_MAX_PDE];
  if (threadIdx.x == 0) {
    float mx = 0.f;
    for (int b = 0; b < blocks; ++b) mx = fmaxf(mx, part[b * kMaxTerms]);
    for (int t = 0; t < n_terms; ++t) {
      float wt = w_in[t];
      if (t >= n_pde) {
        float s = 0.f;
        for (int b = 0; b < blocks; ++b) s += part[b * kMaxTerms + (t - n_pde + 1)];
        const float mean = s / counts[t];
        const float lambda_hat = mx / (mean * wt);
        wt = (1.0f - alpha) * wt + alpha * lambda_hat;
      }
      w[t] = wt;
      if (blockIdx.x == 0) w_out[t] = wt;
    }
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float r = 0.f;
    for (int t = 0; t < n_terms; ++t) r = fmaf(w[t], G[(int64_t)t * ldg + i], r);
    grad_out[i] = r;
  }
}

// ---- MultiAdam (src/optimizer/multiadam.py:54-122 `sadam`, amsgrad off, no aggregated momentum) ----------------------
// One Adam moment pair PER LOSS GROUP; the update is the group-weighted sum of the groups' normalised first moments.
// G: [n_rows][ldg] gradient rows; row t belongs to group gid[t] with factor rw[t] (its loss weight), so the group gradient
// g_k = sum_{gid[t] = k} rw[t] G[t] is formed on the fly.  m, v: [n_groups][P].
constexpr int kMaxGroups = 4;
__global__ void multiadam_kernel(float* __restrict__ p, const float* __restrict__ G, int n_rows, int64_t P, int64_t ldg, const int* __restrict__ gid,
                                 const float* __restrict__ rw, int n_groups, const float* __restrict__ gw, float* __restrict__ m, float* __restrict__ v,
                                 float* step_dev, float lr, float b1, float b2, float eps, float wd) {
  const float t = step_dev[0] + 1.0f;
  const float bc1 = 1.0f - powf(b1, t);
  const float rs_bc2 = sqrtf(1.0f - powf(b2, t));
  const float step_size = lr / bc1;
  __shared__ int s_gid[kMaxTerms + PINN_MAX_PDE];
  __shared__ float s_rw[kMaxTerms + PINN_MAX_PDE], s_gw[kMaxGroups];
  if (threadIdx.x < n_rows) { s_gid[threadIdx.x] = gid[threadIdx.x]; s_rw[threadIdx.x] = rw[threadIdx.x]; }
  if (threadIdx.x < n_groups) s_gw[threadIdx.x] = gw[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float g[kMaxGroups];
#pragma unroll
    for (int k = 0; k < kMaxGroups; ++k) g[k] = 0.f;
    for (int r = 0; r < n_rows; ++r) {
      const float x = s_rw[r] * G[(int64_t)r * ldg + i];
#pragma unroll
      for (int k = 0; k < kMaxGroups; ++k)
        if (s_gid[r] == k) g[k] += x;
    }
    const float pi = p[i];
    float upd = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxGroups; ++k) {
      if (k < n_groups) {
        float gk = g[k];
        if (wd != 0.0f) gk = fmaf(wd, pi, gk);
        const float mk = b1 * m[(int64_t)k * P + i] + (1.0f - b1) * gk;
        const float vk = b2 * v[(int64_t)k * P + i] + (1.0f - b2) * gk * gk;
        m[(int64_t)k * P + i] = mk;
        v[(int64_t)k * P + i] = vk;
        upd = fmaf(s_gw[k], mk / (sqrtf(vk) / rs_bc2 + eps), upd);
      }
    }
    p[i] = pi - step_size * upd;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* ticket = reinterpret_cast<unsigned*>(step_dev + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      step_dev[0] = t;
    }
  }
}

int fail(const char* what, cudaError_t e) {
  pinn::set_error_string(std::string(what) + ": " + cudaGetErrorString(e));
  return 1;
}

}  // namespace

extern "C" {

int pinn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* step_dev, float lr, float beta1,
                   float beta2, float eps, float weight_decay, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !step_dev || n <= 0) {
    pinn::set_error_string("pinn_adam_step: null buffer or n <= 0");
    return 1;
  }
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads < 148 ? (n + threads - 1) / threads : 148);
  adam_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, step_dev, lr, beta1, beta2, eps, weight_decay);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail("pinn_adam_step", e);
}

int pinn_multiadam_step(float* params, const float* G, int n_rows, int64_t P, int64_t ldg, const int* row_group, const float* row_weight,
                        int n_groups, const float* group_weights, float* exp_avg, float* exp_avg_sq, float* step_dev, float lr, float beta1,
                        float beta2, float eps, float weight_decay, void* stream) {
  if (!params || !G || !row_group || !row_weight || !group_weights || !exp_avg || !exp_avg_sq || !step_dev || P <= 0 || ldg < P) {
    pinn::set_error_string("pinn_multiadam_step: null buffer, P <= 0 or ldg < P");
    return 1;
  }
  if (n_groups < 1 || n_groups > kMaxGroups || n_rows < 1 || n_rows > kMaxTerms + PINN_MAX_PDE) {
    pinn::set_error_string("pinn_multiadam_step: 1..4 loss groups and 1..19 gradient rows");
    return 1;
  }
  const int threads = 256;
  const int blocks = (int)((P + threads - 1) / threads < 148 ? (P + threads - 1) / threads : 148);
  multiadam_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(params, G, n_rows, P, ldg, row_group, row_weight, n_groups, group_weights, exp_avg,
                                                                exp_avg_sq, step_dev, lr, beta1, beta2, eps, weight_decay);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail("pinn_multiadam_step", e);
}

size_t pinn_optim_scratch_floats(void) { return (size_t)kStatBlocks * kMaxTerms; }

int pinn_grad_stats(const float* g, int64_t n, float* out3, float* scratch, void* stream) {
  if (!g || !out3 || !scratch || n <= 0) {
    pinn::set_error_string("pinn_grad_stats: null buffer or n <= 0");
    return 1;
  }
  stats_partial_kernel<<<kStatBlocks, kStatThreads, 0, (cudaStream_t)stream>>>(g, n, scratch);
  stats_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scratch, kStatBlocks, out3);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail("pinn_grad_stats", e);
}

int pinn_lra_combine(const float* G, int n_terms, int n_pde, int64_t P, int64_t ldg, const float* w_in, const float* counts, float alpha,
                     float* w_out, float* grad_out, float* scratch, void* stream) {
  if (!G || !w_in || !counts || !w_out || !grad_out || !scratch || P <= 0 || ldg < P) {
    pinn::set_error_string("pinn_lra_combine: null buffer, P <= 0 or ldg < P");
    return 1;
  }
  if (n_pde < 1 || n_terms < n_pde || n_terms - n_pde + 1 > kMaxTerms) {
    pinn::set_error_string("pinn_lra_combine: need 1 <= n_pde <= n_terms and at most 15 boundary terms");
    return 1;
  }
  lra_partial_kernel<<<kStatBlocks, kStatThreads, 0, (cudaStream_t)stream>>>(G, n_terms, n_pde, P, ldg, scratch);
  lra_apply_kernel<<<kStatBlocks, kStatThreads, 0, (cudaStream_t)stream>>>(G, n_terms, n_pde, P, ldg, scratch, kStatBlocks, w_in, counts, alpha, w_out,
                                                                             grad_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail("pinn_lra_combine", e);
}

}  // extern "C"
