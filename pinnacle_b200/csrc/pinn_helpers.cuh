// pinn_helpers.cuh -- small non-template kernels (weight packing, partial reduction, FFMA probe).
// Included by pinn_api.cu only.
#pragma once
#include <cuda_bf16.h>

#include "pinn_kernels.cuh"

namespace pinn {

// ------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------
// wpack <- [W_l^T (l=2..L) | W_l (l=2..L)] from the flat parameter vector
__global__ void pinn_pack_weights(const float* __restrict__ params, float* __restrict__ wpack, int H, int D, int L) {
  const long long offHid = (long long)H * D + H;
  const long long strideHid = (long long)H * H + H;
  const int HH = H * H;
  const int total = (L - 1) * HH;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int l = i / HH, e = i % HH;
    const int a = e / H, b = e % H;
    const float* W = params + offHid + (long long)l * strideHid;
    wpack[i] = W[b * H + a];                              // transposed: [k][n] = W[n][k]
    wpack[(size_t)total + i] = W[e];                      // native
  }
}

// Tensor-core path: every MMA A-operand matrix (W_l, W_l^T for l = 2..L, and the output layer zero-padded) is
// pre-split into its three bf16 terms and stored in the exact per-thread order the kernel's fetch_W() consumes:
//   wperm[((mat*12 + slot)*512 + tid)] (uint4),  tid = 128*g + 32*q + lane  <->  row j = 32*q + lane, K slice [32*g, 32*g+32)
//   slot = 4*split + quarter  holds packed pairs (k, k+1) for k = 32*g + 8*quarter + {0,2,4,6}
// so one warp-wide 128-bit load is 512 contiguous bytes.  mat order: W_2..W_L, W_2^T..W_L^T, W_out(padded).
__device__ __forceinline__ uint32_t pack_bf16x2_h(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__global__ void pinn_pack_weights_tc(const float* __restrict__ params, uint4* __restrict__ wperm, int H, int D, int L, int M) {
  const long long offHid = (long long)H * D + H;
  const long long strideHid = (long long)H * H + H;
  const long long offW6 = offHid + (long long)(L - 1) * strideHid;
  const int nmat = 2 * (L - 1) + 1;
  const int total = nmat * 4 * 512;                 // one work item = (mat, quarter, tid): 8 consecutive k of one row
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
    const int tid = it % 512, quarter = (it / 512) % 4, mat = it / 2048;
    const int warp = tid >> 5, lane = tid & 31, q = warp & 3, g = warp >> 2;
    const int row = 32 * q + lane, k0 = 32 * g + 8 * quarter;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      float x = 0.f;
      if (k < H) {
        if (mat < L - 1) {
          if (row < H) x = params[offHid + (long long)mat * strideHid + (long long)row * H + k];            // W_l[row][k]
        } else if (mat < 2 * (L - 1)) {
          if (row < H) x = params[offHid + (long long)(mat - (L - 1)) * strideHid + (long long)k * H + row];  // W_l^T[row][k] = W_l[k][row]
        } else {
          if (row < M) x = params[offW6 + (long long)row * H + k];
        }
      }
      v[i] = x;
    }
    // Forward operands (W_l and the output layer): the leading split w1 keeps 8 significant bits but is never finer than 2^-14
    // of the row's leading bit, so that all products w1 * a1 of a dot product sit on one coarse grid and the tensor core's
    // accumulate of them is exact (pinn_tc2_kernels.cuh header).  w1 + w2 + w3 == w still holds bit for bit for every weight
    // within 7 binades of the row maximum (below that the loss is < 2^-31 of the maximum).  Transposed operands keep the
    // plain floating split.
    float floor_grid = 0.f;
    if (mat < L - 1 || mat == 2 * (L - 1)) {
      unsigned mx = 0u;
      if (mat < L - 1 ? row < H : row < M) {
        const float* wr = (mat < L - 1) ? params + offHid + (long long)mat * strideHid + (long long)row * H : params + offW6 + (long long)row * H;
        for (int k = 0; k < H; ++k) {
          const unsigned b = __float_as_uint(wr[k]) & 0x7fffffffu;
          mx = b > mx ? b : mx;
        }
      }
      const int e = (int)(mx >> 23);                       // leading bit 2^(e - 127)
      floor_grid = (e > 14) ? __uint_as_float((unsigned)(e - 14) << 23) : 0.f;
    }
    uint32_t p1[4], p2[4], p3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = v[2 * i], x1 = v[2 * i + 1];
      if (floor_grid > 0.f) {
        // grid of the 8-bit floating split of x: 2^(e_x - 7); the coarser of that and the floor
        const float g0 = fmaxf(__uint_as_float((__float_as_uint(x0) & 0x7f800000u)) * 0.0078125f, floor_grid);
        const float g1 = fmaxf(__uint_as_float((__float_as_uint(x1) & 0x7f800000u)) * 0.0078125f, floor_grid);
        p1[i] = pack_bf16x2_h(rintf(x0 / g0) * g0, rintf(x1 / g1) * g1);       // exact: <= 8 significant bits
      } else {
        p1[i] = pack_bf16x2_h(x0, x1);
      }
      float r0 = x0 - __uint_as_float(p1[i] << 16), r1 = x1 - __uint_as_float(p1[i] & 0xFFFF0000u);
      p2[i] = pack_bf16x2_h(r0, r1);
      r0 -= __uint_as_float(p2[i] << 16);
      r1 -= __uint_as_float(p2[i] & 0xFFFF0000u);
      p3[i] = pack_bf16x2_h(r0, r1);
    }
    uint4* dst = wperm + (size_t)mat * 12 * 512 + tid;
    dst[(0 * 4 + quarter) * 512] = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    dst[(1 * 4 + quarter) * 512] = make_uint4(p2[0], p2[1], p2[2], p2[3]);
    dst[(2 * 4 + quarter) * 512] = make_uint4(p3[0], p3[1], p3[2], p3[3]);
  }
}

// out[s*P + i] = sum_cta gpart[cta][s*Pstride + i];  out[S*P + k] = inv_n * sum_cta lpart[cta][k].
// When gwt != nullptr (tensor-core path) the hidden-layer weight entries come from the transposed partials
// gwt[cta][s][l][k/4][n][k%4] instead.
__global__ void pinn_reduce_partials(const float* __restrict__ gpart, const double* __restrict__ lpart, int grid, int S, long long P,
                                     long long Pstride, int n_pde, double inv_n, float* __restrict__ grad_out,
                                     float* __restrict__ loss_out, const float* __restrict__ gwt, int H, int D, int L) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)S * P) {
    const long long sidx = i / P, e = i % P;
    const long long offHid = (long long)H * D + H, strideHid = (long long)H * H + H;
    double s = 0.0;
    bool hidden_w = false;
    long long toff = 0;
    if (gwt != nullptr && e >= offHid && e < offHid + (long long)(L - 1) * strideHid) {
      const long long r = (e - offHid) % strideHid, l = (e - offHid) / strideHid;
      if (r < (long long)H * H) {
        hidden_w = true;
        const int n = (int)(r / H), k = (int)(r % H);
        toff = (l * 25 + k / 4) * 512 + n * 4 + (k % 4);
      }
    }
    if (hidden_w) {
      const size_t per = (size_t)(L - 1) * 25 * 512;
      for (int c = 0; c < grid; ++c) s += (double)gwt[((size_t)c * S + sidx) * per + toff];
    } else {
      for (int c = 0; c < grid; ++c) s += (double)gpart[((size_t)c * S + sidx) * Pstride + e];
    }
    grad_out[i] = (float)s;
  }
  if (blockIdx.x == 0 && threadIdx.x < n_pde) {
    double s = 0.0;
    for (int c = 0; c < grid; ++c) s += lpart[(size_t)c * PINN_MAX_PDE + threadIdx.x];
    loss_out[threadIdx.x] = (float)(s * inv_n);
  }
}

// FP32 FFMA throughput probe (roofline denominator)
__global__ void pinn_ffma_probe(float* out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 0.001f * (float)(threadIdx.x + i);
  const float b = 1.0000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) out[0] = s;
}

}  // namespace pinn
