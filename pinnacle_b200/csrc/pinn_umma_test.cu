// pinn_umma_test.cu -- single-CTA tcgen05 GEMM used to validate descriptor conventions on real hardware
// (tests/test_cuda_umma.py).  D[128][N] = A[128][K] * B[K][N], tf32 inputs, fp32 accumulate in TMEM.
//   mode 0: A in TMEM (tcgen05.st), B in shared memory, MN-major (N contiguous)   -- forward / Abar contraction
//   mode 1: A and B in shared memory, both K-major                               -- dW contraction
// `variant` bit 0 swaps the LBO/SBO descriptor fields.
#include <cuda_runtime.h>

#include <string>

#include "../../include/pinn_b200.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "umma.cuh"

namespace pinn {

__device__ __forceinline__ void mbar_init1(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma::smem_addr(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  const uint32_t addr = umma::smem_addr(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(phase)
        : "memory");
  } while (!ok);
}

__global__ void __launch_bounds__(128, 1) umma_gemm_test_kernel(int mode, int variant, int K, int N, const float* __restrict__ A,
                                                                const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid / 32;
  float* bufA = reinterpret_cast<float*>(sm);
  const int JBa = (K / 4) * 128;                 // bytes, mode 1 (r = k)
  const int JBb = (N / 4) * 128;                 // bytes, mode 0 (r = n)
  const bool a_tmem = (mode == 0 || mode == 2);
  const bool b_mn = (mode == 0 || mode == 3 || mode == 4);
  float* bufB = reinterpret_cast<float*>(sm + (!a_tmem ? 16 * JBa : 0));
  const int bytesB = b_mn ? (K / 8) * JBb : (N / 8) * JBa;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) mbar_init1(&bar);
  // zero the operand buffers (pad rows must be finite)
  for (int i = tid; i < ((!a_tmem ? 16 * JBa : 0) + bytesB) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t tD = tbase;            // columns [0, N)
  const uint32_t tA = tbase + 256;      // columns [256, 256+K)

  if (mode == 4) {
    for (int i = tid; i < bytesB / 4; i += 128) bufB[i] = (float)i;      // value = word offset inside the B buffer
  } else if (b_mn) {
    // B[k][n] -> layout X with r = n, j = k
    for (int i = tid; i < K * N; i += 128) {
      const int k = i / N, n = i % N;
      bufB[((k / 8) * JBb + (n / 4) * 128 + (k % 8) * 16 + (n % 4) * 4) / 4] = B[i];
    }
  } else {
    for (int i = tid; i < K * N; i += 128) {
      const int k = i / N, n = i % N;
      bufB[((n / 8) * JBa + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4) / 4] = B[i];
    }
  }
  if (a_tmem) {
    // A row m = tid -> TMEM lane m, columns tA + k
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = A[tid * K + k0 + q];
      umma::tmem_st8(tA + lane_base + k0, v);
    }
    umma::wait_st();
  } else {
    for (int i = tid; i < 128 * K; i += 128) {
      const int m = i / K, k = i % K;
      bufA[((m / 8) * JBa + (k / 4) * 128 + (m % 8) * 16 + (k % 4) * 4) / 4] = A[i];
    }
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();

  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_tf32(128, N, 0, b_mn ? 1 : 0);
    const int ksteps = K / 8;
    for (int ks = 0; ks < ksteps; ++ks) {
      uint64_t bd;
      if (b_mn) {
        uint32_t lbo = JBb, sbo = 128;
        if (variant & 1) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        bd = umma::make_smem_desc(umma::smem_addr(bufB) + ks * JBb, lbo, sbo);
      } else {
        bd = umma::make_smem_desc(umma::smem_addr(bufB) + ks * 256, 128, JBa);
      }
      if (a_tmem) {
        umma::mma_tf32_ts(tD, tA + ks * 8, bd, idesc, ks > 0);
      } else {
        const uint64_t ad = umma::make_smem_desc(umma::smem_addr(bufA) + ks * 256, 128, JBa);
        umma::mma_tf32_ss(tD, ad, bd, idesc, ks > 0);
      }
    }
    umma::commit(&bar);
  }
  mbar_wait_parity(&bar, 0);
  umma::fence_after_sync();
  {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int n0 = 0; n0 < N; n0 += 8) {
      float v[8];
      umma::tmem_ld8(tD + lane_base + n0, v);
      umma::wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q) D[tid * N + n0 + q] = v[q];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}


// bf16 probe.  Storage format "X16" (16-bit elements, un-swizzled core matrix = 8 rows x 16 bytes):
//     addr(r, j) = (j/8)*JB + (r/8)*128 + (j%8)*16 + (r%8)*2      bytes, JB = (R/8)*128
//   mode 10: A = X16(r=k, j=m), B = X16(r=k, j=n)  -> both K-major (K contiguous inside 16 bytes)
//   mode 11: A = X16(r=k, j=m) K-major,  B = X16(r=n, j=k) MN-major (N contiguous inside 16 bytes)
//   mode 12: A in TMEM (two bf16 per 32-bit column), B = X16(r=k, j=n) K-major
//   mode 13: A in TMEM, B MN-major
//   modes 20..23: the same four with fp16 operands (kind::f16, a_format = b_format = F16)
__device__ __forceinline__ uint16_t to16(float v, bool fp16) {
  if (fp16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16(v));
}
__global__ void __launch_bounds__(128, 1) umma_gemm_test_bf16_kernel(int mode, int variant, int K, int N, const float* __restrict__ A,
                                                                     const float* __restrict__ B, float* __restrict__ D) {
  const bool fp16 = mode >= 20;
  if (fp16) mode -= 10;
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid / 32;
  const bool a_tmem = (mode == 12 || mode == 13);
  const bool b_mn = (mode == 11 || mode == 13);
  const int JBk = (K / 8) * 128;     // r = k
  const int JBn = (N / 8) * 128;     // r = n
  uint16_t* bufA = reinterpret_cast<uint16_t*>(sm);
  const int bytesA = 16 * JBk;
  uint16_t* bufB = reinterpret_cast<uint16_t*>(sm + bytesA);
  const int bytesB = b_mn ? (K / 8) * JBn : (N / 8) * JBk;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) mbar_init1(&bar);
  for (int i = tid; i < (bytesA + bytesB) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t tD = tbase, tA = tbase + 256;

  for (int i = tid; i < K * N; i += 128) {
    const int k = i / N, n = i % N;
    const int off = b_mn ? ((k / 8) * JBn + (n / 8) * 128 + (k % 8) * 16 + (n % 8) * 2) : ((n / 8) * JBk + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2);
    bufB[off / 2] = to16(B[i], fp16);
  }
  if (a_tmem) {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int k0 = 0; k0 < K; k0 += 16) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t pr = (uint32_t)to16(A[tid * K + k0 + 2 * q], fp16) | ((uint32_t)to16(A[tid * K + k0 + 2 * q + 1], fp16) << 16);   // low half first
        v[q] = __uint_as_float(pr);
      }
      umma::tmem_st8(tA + lane_base + k0 / 2, v);
    }
    umma::wait_st();
  } else {
    for (int i = tid; i < 128 * K; i += 128) {
      const int m = i / K, k = i % K;
      bufA[((m / 8) * JBk + (k / 8) * 128 + (m % 8) * 16 + (k % 8) * 2) / 2] = to16(A[i], fp16);
    }
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();

  if (tid == 0) {
    uint32_t idesc = umma::make_idesc_bf16(128, N, 0, b_mn ? 1 : 0);
    if (fp16) idesc &= ~((1u << 7) | (1u << 10));       // a_format = b_format = 0 (F16)
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t bd;
      if (b_mn) {
        uint32_t lbo = JBn, sbo = 128;                    // K-block stride, MN-block stride
        if (variant & 1) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        bd = umma::make_smem_desc(umma::smem_addr(bufB) + ks * 2 * JBn, lbo, sbo);
      } else {
        bd = umma::make_smem_desc(umma::smem_addr(bufB) + ks * 256, 128, JBk);
      }
      if (a_tmem) {
        umma::mma_bf16_ts(tD, tA + ks * 8, bd, idesc, ks > 0);
      } else {
        const uint64_t ad = umma::make_smem_desc(umma::smem_addr(bufA) + ks * 256, 128, JBk);
        umma::mma_bf16_ss(tD, ad, bd, idesc, ks > 0);
      }
    }
    umma::commit(&bar);
  }
  mbar_wait_parity(&bar, 0);
  umma::fence_after_sync();
  {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int n0 = 0; n0 < N; n0 += 8) {
      float v[8];
      umma::tmem_ld8(tD + lane_base + n0, v);
      umma::wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q) D[tid * N + n0 + q] = v[q];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

}  // namespace pinn

extern "C" int pinn_debug_umma_gemm(int mode, int variant, int K, int N, const float* A, const float* B, float* D, void* stream) {
  if ((mode >= 10 && mode <= 13) || (mode >= 20 && mode <= 23)) {
    if (K % 16 != 0 || N % 16 != 0 || N > 256 || K > 256) return 1;
    const size_t smem16 = (size_t)16 * (K / 8) * 128 + (size_t)(K / 8) * (N / 8) * 128 + 1024;
    if (cudaFuncSetAttribute(pinn::umma_gemm_test_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16) != cudaSuccess) return 2;
    pinn::umma_gemm_test_bf16_kernel<<<1, 128, smem16, reinterpret_cast<cudaStream_t>(stream)>>>(mode, variant, K, N, A, B, D);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
  }
  if (K % 8 != 0 || N % 16 != 0 || N > 256 || K > 256 || mode < 0 || mode > 4) return 1;
  const int JBa = (K / 4) * 128, JBb = (N / 4) * 128;
  const size_t smem = (size_t)16 * JBa + (size_t)(K / 8) * JBb + (size_t)(N / 8) * JBa;
  cudaError_t e = cudaFuncSetAttribute(pinn::umma_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024);
  if (e != cudaSuccess) return 2;
  pinn::umma_gemm_test_kernel<<<1, 128, smem + 1024, reinterpret_cast<cudaStream_t>(stream)>>>(mode, variant, K, N, A, B, D);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
