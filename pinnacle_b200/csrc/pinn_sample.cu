// pinn_sample.cu -- collocation points drawn on the device (`PDEPointResampler` on device).
//
// The reference redraws its residual points on the host: dde.callbacks.PDEPointResampler (deepxde/callbacks.py:540-575) calls
// PDE.resample_train_points (deepxde/data/pde.py:186-193), which re-runs geom.random_points(num_domain, "pseudo") --
// np.random.rand scaled into the bounding box (geometry_nd.py Hypercube.random_points) with rejection of the points that fall
// into a subtracted shape (csg.py CSGDifference.random_points) -- and the new float32 array is then uploaded by the next
// train step.  At 16 Mi rows that is a few hundred ms of numpy and a 200 MB copy per resample.  Here one launch writes the new
// rows straight into the staged device array.
//
// Every row is a pure function of (key, draw, global row index): Philox4x32-10 on the counter (row_lo, row_hi, attempt,
// block | draw << 8) gives four 32-bit words = four coordinates of block `block` (coordinates 4*block .. 4*block+3); a point
// that falls into a hole is redrawn with attempt + 1.  No generator state and no global atomics: the rows a rank draws do not
// depend on how the set is sharded, and the oracle (oracle/sample_ref.py) restates the same integer and float32 arithmetic bit
// for bit.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/pinn_b200.h"

namespace pinn {
void set_error_string(const std::string& s);   // pinn_api.cu
}

namespace {

constexpr int kRows = 256;          // rows per CTA = threads per CTA
constexpr int kMaxAttempts = 256;   // a row still rejected after this many draws keeps its last draw (5-ball in its cube: 0.836^256 = 1e-20)

struct U4 { uint32_t x, y, z, w; };

// What the host precomputes once per launch (all of it exact, so the float32 result is the one the header describes):
// the ten round keys of Philox (key + r * Weyl constant), lo, and (hi - lo) * 2^-24 -- a power-of-two scaling commutes with
// the rounding of the product, fl(fl(u * 2^-24) * span) == fl(u * (span * 2^-24)) for the 24-bit integer u.
struct SampleParams {
  uint32_t rk[10][2];
  float lo[PINN_MAX_DIM];
  float span24[PINN_MAX_DIM];
};

__device__ __forceinline__ U4 philox4x32_10(U4 c, const SampleParams& P) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = U4{hi1 ^ c.y ^ P.rk[r][0], lo1, hi0 ^ c.w ^ P.rk[r][1], lo0};
  }
  return c;
}

__device__ __forceinline__ uint32_t word(const U4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }

// One draw: coordinates p[0..D) of global row g, attempt `attempt`, and (REJECT) whether the point is rejected: outside the kept
// ball or inside a hole.
template <int D, bool REJECT>
__device__ __forceinline__ bool draw_point(const pinn_region& R, const SampleParams& P, uint64_t g, uint32_t attempt, uint32_t draw,
                                           float (&p)[D]) {
#pragma unroll
  for (int b = 0; b < (D + 3) / 4; ++b) {
    const U4 v = philox4x32_10(U4{(uint32_t)g, (uint32_t)(g >> 32), attempt, (uint32_t)b | (draw << 8)}, P);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * b + j;
      if (c < D) p[c] = __fadd_rn(__fmul_rn((float)(word(v, j) >> 8), P.span24[c]), P.lo[c]);
    }
  }
  if constexpr (!REJECT) return false;
  bool rejected = false;
  if (R.keep_dims > 0) {                                     // the domain is a ball (HeatND's Hypersphere): outside is redrawn
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (j < R.keep_dims) {
        const float dj = __fsub_rn(p[j], R.keep_center[j]);
        s = __fadd_rn(s, __fmul_rn(dj, dj));
      }
    rejected = s > __fmul_rn(R.keep_radius, R.keep_radius);
  }
  for (int h = 0; h < R.n_holes && !rejected; ++h) {
    const pinn_hole& H = R.holes[h];
    const int k = H.n_dims;                                  // 1..min(3, D), validated on the host
    if (H.shape == PINN_HOLE_BALL) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < (D < 3 ? D : 3); ++j)
        if (j < k) {
          const float dj = __fsub_rn(p[j], H.a[j]);
          s = __fadd_rn(s, __fmul_rn(dj, dj));
        }
      rejected = s <= __fmul_rn(H.b[0], H.b[0]);
    } else {
      bool all = true;
#pragma unroll
      for (int j = 0; j < (D < 3 ? D : 3); ++j)
        if (j < k) all = all && p[j] >= H.a[j] && p[j] <= H.b[j];
      rejected = all;
    }
  }
  return rejected;
}

// x[row][0..D) for rows [0, n_rows).  A CTA owns kRows consecutive rows and stages them in shared memory.  Round 0: every
// thread draws its own row.  REJECT: the rows that were rejected are collected in a shared list and redrawn by the FIRST threads
// of the CTA in the next round (attempt + 1), and so on: a warp never waits for its slowest lane's rejections, the cost of a
// redraw round is proportional to the rows still pending (13 % of the rows need a second attempt on the reference's
// 17-disk domain; done lane by lane that kept every warp busy for its unluckiest row, 3-4 full attempts).  The tile goes
// out with unit-stride stores when rows are contiguous (ld == D), otherwise only the first D of every ld columns are touched.
template <int D, bool REJECT>
__global__ void __launch_bounds__(kRows) sample_kernel(float* __restrict__ x, int64_t ld, int64_t n_rows, const __grid_constant__ pinn_region R,
                                                      const __grid_constant__ SampleParams P, uint32_t draw, uint64_t first_row) {
  __shared__ float tile[kRows * D];
  __shared__ int pending[REJECT ? 2 : 1][REJECT ? kRows : 1];
  __shared__ int n_pending[2];
  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  const int64_t left = n_rows - row0;
  const int rows_here = (int)(left < kRows ? left : kRows);
  if constexpr (REJECT) {
    if (t < 2) n_pending[t] = 0;
    __syncthreads();
  }
  float p[D];
  if (t < rows_here) {
    const bool rejected = draw_point<D, REJECT>(R, P, first_row + (uint64_t)(row0 + t), 0u, draw, p);
#pragma unroll
    for (int c = 0; c < D; ++c) tile[t * D + c] = p[c];
    if constexpr (REJECT)
      if (rejected) pending[0][atomicAdd(&n_pending[0], 1)] = t;
  }
  if constexpr (REJECT) {
    for (uint32_t attempt = 1; attempt < (uint32_t)kMaxAttempts; ++attempt) {
      const int cur = (attempt - 1) & 1, nxt = cur ^ 1;
      __syncthreads();
      const int n = n_pending[cur];
      if (n == 0) break;                                     // uniform: read after the barrier
      if (t < n) {
        const int r = pending[cur][t];
        const bool rejected = draw_point<D, true>(R, P, first_row + (uint64_t)(row0 + r), attempt, draw, p);
#pragma unroll
        for (int c = 0; c < D; ++c) tile[r * D + c] = p[c];  // the last draw stays if every attempt is rejected
        if (rejected) pending[nxt][atomicAdd(&n_pending[nxt], 1)] = r;
      }
      __syncthreads();
      if (t == 0) n_pending[cur] = 0;                        // consumed; it is the next round's output list
    }
  }
  __syncthreads();
  const int count = rows_here * D;
  if (ld == D) {
    float* base = x + row0 * D;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const int i = t + k * kRows;
      if (i < count) base[i] = tile[i];
    }
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const int i = t + k * kRows;
      if (i < count) x[(row0 + i / D) * ld + i % D] = tile[i];
    }
  }
}

template <int D>
cudaError_t launch_sample(float* x, int64_t ld, int64_t n_rows, const pinn_region& R, const SampleParams& P, uint32_t draw, uint64_t first_row,
                          unsigned blocks, cudaStream_t st) {
  if (R.n_holes > 0 || R.keep_dims > 0)
    sample_kernel<D, true><<<blocks, kRows, 0, st>>>(x, ld, n_rows, R, P, draw, first_row);
  else
    sample_kernel<D, false><<<blocks, kRows, 0, st>>>(x, ld, n_rows, R, P, draw, first_row);
  return cudaGetLastError();
}

}  // namespace

extern "C" {

int pinn_sample_points(float* x, int64_t ld, int64_t n_rows, const pinn_region* region, uint64_t key, uint32_t draw, uint64_t first_row,
                       void* stream) {
  if (!x || !region || n_rows < 0 || ld < 1) {
    pinn::set_error_string("pinn_sample_points: null buffer, negative row count or ld < 1");
    return 1;
  }
  const pinn_region& R = *region;
  if (R.d < 1 || R.d > PINN_MAX_DIM || ld < R.d || R.n_holes < 0 || R.n_holes > PINN_MAX_HOLES || draw >= (1u << 24)) {
    pinn::set_error_string("pinn_sample_points: need 1 <= d <= PINN_MAX_DIM <= ld, at most PINN_MAX_HOLES holes and draw < 2^24");
    return 1;
  }
  for (int c = 0; c < R.d; ++c)
    if (!(R.hi[c] >= R.lo[c])) {
      pinn::set_error_string("pinn_sample_points: empty bounding box (hi < lo or NaN)");
      return 1;
    }
  if (R.keep_dims < 0 || R.keep_dims > R.d || (R.keep_dims > 0 && !(R.keep_radius > 0.f))) {
    pinn::set_error_string("pinn_sample_points: keep_dims must be 0..d and the kept ball needs a positive radius");
    return 1;
  }
  for (int h = 0; h < R.n_holes; ++h) {
    const pinn_hole& H = R.holes[h];
    if ((H.shape != PINN_HOLE_BALL && H.shape != PINN_HOLE_BOX) || H.n_dims < 1 || H.n_dims > 3 || H.n_dims > R.d) {
      pinn::set_error_string("pinn_sample_points: a hole is a ball or a box over the first 1..3 coordinates");
      return 1;
    }
  }
  if (n_rows == 0) return 0;
  const int64_t blocks = (n_rows + kRows - 1) / kRows;
  if (blocks > 0x7fffffffLL) {
    pinn::set_error_string("pinn_sample_points: more than 2^31 * 256 rows in one call");
    return 1;
  }
  SampleParams P;
  for (int r = 0; r < 10; ++r) {
    P.rk[r][0] = (uint32_t)key + (uint32_t)r * 0x9E3779B9u;
    P.rk[r][1] = (uint32_t)(key >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  for (int c = 0; c < PINN_MAX_DIM; ++c) {
    const volatile float span = c < R.d ? R.hi[c] - R.lo[c] : 0.f;      // one float32 subtraction, then an exact scaling
    P.lo[c] = c < R.d ? R.lo[c] : 0.f;
    P.span24[c] = span * 5.9604644775390625e-08f;
  }
  cudaError_t e = cudaSuccess;
  cudaStream_t st = (cudaStream_t)stream;
  switch (R.d) {
    case 1: e = launch_sample<1>(x, ld, n_rows, R, P, draw, first_row, (unsigned)blocks, st); break;
    case 2: e = launch_sample<2>(x, ld, n_rows, R, P, draw, first_row, (unsigned)blocks, st); break;
    case 3: e = launch_sample<3>(x, ld, n_rows, R, P, draw, first_row, (unsigned)blocks, st); break;
    case 4: e = launch_sample<4>(x, ld, n_rows, R, P, draw, first_row, (unsigned)blocks, st); break;
    case 5: e = launch_sample<5>(x, ld, n_rows, R, P, draw, first_row, (unsigned)blocks, st); break;
    default: e = launch_sample<6>(x, ld, n_rows, R, P, draw, first_row, (unsigned)blocks, st); break;
  }
  if (e != cudaSuccess) {
    pinn::set_error_string(std::string("pinn_sample_points: ") + cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

}  // extern "C"
