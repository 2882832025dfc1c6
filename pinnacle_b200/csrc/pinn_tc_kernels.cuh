// pinn_tc_kernels.cuh -- tensor-core (tcgen05 / TMEM) formulation of the fused PDE-residual step, sm_100a.
//
// Same algorithm and interface as pinn_fused_kernel (pinn_kernels.cuh); the H x H layer contractions run on the
// 5th-generation tensor cores with split-precision operands:  x = b1 + b2 + b3 (three bf16 terms, 24 mantissa
// bits), products b1b1 + b1b2 + b2b1 + b2b2 + b1b3 + b3b1 accumulated in fp32 in TMEM ("bf16x3", 6 MMAs per
// k-step, relative product error ~2^-26 -- at least the accuracy of 3xTF32; bf16 is used instead of tf32 because
// un-swizzled MN-major operands exist only for 16-bit types, which lets ONE shared-memory copy of an activation
// matrix serve all three contractions; validated on hardware by pinn_debug_umma_gemm).
//
// Orientation: neurons on the M axis / TMEM lanes, (jet channel, point) pairs on the N axis / TMEM columns:
//     forward      Z^l [n][r]    = W_l   [n][k] * A^{l-1}[k][r]     A-operand W_l   in TMEM,  B = A^{l-1} smem, MN-major
//     reverse (ii) Abar^{l-1}[k][r] = W_l^T[k][n] * Zbar^l[n][r]     A-operand W_l^T in TMEM,  B = Zbar^l  smem, MN-major
//     reverse (i)  dW_l[n][k]    = Zbar^l[n][r] * A^{l-1}[k][r]^T    A = Zbar^l smem K-major,  B = A^{l-1} smem K-major
// so thread (lane j, column group g) of the 16 epilogue warps owns neuron j for a quad of points and all jet
// channels -- exactly what the element-wise tanh-jet stage needs.  Shared-memory storage of every activation
// matrix ("X16"):  addr(r, j) = (j/8)*JB + (r/8)*128 + (j%8)*16 + (r%8)*2 bytes, JB = (R/8)*128.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jet_math.cuh"
#include "pinn_kernels.cuh"
#include "umma.cuh"

namespace pinn {

constexpr int kTcHP = 112;   // neuron count padded to a multiple of 16 (K of the bf16 MMA, N of the dW MMA)
constexpr int kTcRows = 128; // UMMA M
constexpr int kTcWMat = 12 * 512;   // uint4 per pre-split operand matrix
constexpr int kTcGwtLayer4 = 25 * 128;   // float4 per hidden-layer weight-gradient partial, layout [k/4][n (128)][k%4]

struct TcLaunchParams {
  KindParams kp;
  int L, n_aux, n_pde, n_seeds;
  long long n_rows, P, Ppad;
  float inv_n;
  const float* params;
  // pre-split bf16x3 A-operands in per-thread fetch order (pinn_pack_weights_tc): 12*512 uint4 per matrix
  const uint4* wn;     // W_l,   l = 2..L
  const uint4* wt;     // W_l^T, l = 2..L
  const uint4* w6p;    // output layer, zero-padded to 128 x 112
  const float* x;
  const float* aux;
  const float* seeds;
  float* stash;        // [grid][L][NQ][C][512] float4
  float* gpart;        // [grid][n_seeds][Ppad]   (hidden-layer weight entries unused on this path)
  float* gwt;          // [grid][n_seeds][L-1][25][128][4]  hidden-layer weight-gradient partials
  double* lpart;
  float* resid_out;
};

template <class SIG_, int M_, int H_, int T_, bool BWD_>
struct TCfg {
  using SIG = SIG_;
  static constexpr int M = M_, H = H_, T = T_;
  static constexpr bool BWD = BWD_;
  static constexpr int C = SIG::C, D = SIG::D;
  static constexpr int NTH = 512;
  static constexpr int R = (C * T + 15) / 16 * 16;   // rows = N of the forward MMA (row c*T + p; rows >= C*T are zero padding)
  static constexpr int NQUADS = T / 4;            // quads of 4 points in a tile, dealt round-robin to the 4 column groups
  static constexpr int NQ = (NQUADS + 3) / 4;     // quads per thread (groups g >= NQUADS % 4 may own one less)
  static constexpr int JB = (R / 8) * 128;        // bytes between 8-neuron blocks
  static constexpr int ARR = 14 * JB;             // one bf16 split array (112 neuron rows)
  static constexpr int NARR = BWD ? 6 : 3;
  // dW flush: private read-modify-write partial (cheaper per element, needs the longer MMA window of a 32-point
  // tile to hide its load latency) or 128-bit REDs (fire-and-forget; better for 16-point tiles)
  static constexpr bool RMW_FLUSH = (NQ >= 2);
  // TMEM columns
  static constexpr int tcD = 0, tcDW = 160, tcW = 272, tcTotal = 512;
  // shared memory carve-up (bytes)
  static constexpr int oArr = 0;
  static constexpr int oSmall = NARR * ARR;
  static constexpr int oW1 = oSmall;                           // float [H*D]
  static constexpr int oBias = oW1 + H * D * 4;                // float [kMaxLayers*H]
  static constexpr int oW6 = oBias + kMaxLayers * H * 4;       // float [M*H]
  static constexpr int oB6 = oW6 + M * H * 4;                  // float [4]
  static constexpr int oX = oB6 + 16;                          // float [T*D]
  static constexpr int oAux = oX + T * D * 4;                  // float [T*4]
  static constexpr int oLw = oAux + T * PINN_MAX_AUX * 4;      // per-point direction weights of the Laplacian channel [T][PINN_MAX_DIM]
  static constexpr int oUo = oLw + (SIG::kLapPW ? T * PINN_MAX_DIM * 4 : 0);       // float [M*R]
  static constexpr int oUb = oUo + M * R * 4;                  // float [M*R]
  static constexpr int oRs = oUb + M * R * 4;                  // float [3*T]
  static constexpr int oRed = oRs + PINN_MAX_PDE * T * 4;      // float [4*128]
  static constexpr int oLred = (oRed + 4 * 128 * 4 + 7) / 8 * 8;  // double [3*T]
  static constexpr int oBar = oLred + PINN_MAX_PDE * T * 8;    // 2 x uint64 + tmem slot
  static constexpr int oKp = oBar + 32;
  static constexpr int nBytes = oKp + (int)((sizeof(KindParams) + 15) / 16 * 16);
  static constexpr size_t smem_bytes = (size_t)nBytes;

  static_assert(H == 100, "tensor-core path is instantiated for H = 100");
  static_assert(T % 4 == 0 && T >= 8 && T <= 48, "T");
  static_assert(R % 16 == 0 && R <= 160, "R must fit the TMEM map");
  static_assert(oSmall % 16 == 0 && oUo % 16 == 0, "alignment");
  static_assert(nBytes - oSmall >= 2 * JB, "A-operand over-read of the last Zbar array must stay inside the allocation");
  static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
};

// ------------------------------------------------------------------------------------------------
// bf16x3 split helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x (low 16 bits) = lo
  return *reinterpret_cast<const uint32_t*>(&v);
}
// (x0, x1) -> three packed bf16 pairs p1 + p2 + p3 ~= x
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
#ifdef PINN_EXP_NO_SPLIT    // what-if timing experiment only
  p1 = pack_bf16x2(x0, x1); p2 = p1; p3 = p1;
  return;
#endif
  p1 = pack_bf16x2(x0, x1);
  float r0 = x0 - __uint_as_float(p1 << 16);
  float r1 = x1 - __uint_as_float(p1 & 0xFFFF0000u);
  p2 = pack_bf16x2(r0, r1);
  r0 -= __uint_as_float(p2 << 16);
  r1 -= __uint_as_float(p2 & 0xFFFF0000u);
  p3 = pack_bf16x2(r0, r1);
}

// (x0, x1) -> two packed bf16 pairs p2 + p3 ~= x (16 significant bits: the remainder of a value whose leading 8 bits went elsewhere)
__device__ __forceinline__ void split2_pair(float x0, float x1, uint32_t& p2, uint32_t& p3) {
  p2 = pack_bf16x2(x0, x1);
  const float r0 = x0 - __uint_as_float(p2 << 16);
  const float r1 = x1 - __uint_as_float(p2 & 0xFFFF0000u);
  p3 = pack_bf16x2(r0, r1);
}

// Accurate tanh for the pipelined kernel's forward sweep: evaluated in double and rounded once, so the error is at most half an
// ulp and carries no bias.  tanhf is within 2 ulp but slightly biased, and a bias that is coherent over all points acts like
// a weight perturbation -- 1.3e-5 of gradient error on the nearly converged Laplace problem, with the FFMA kernel
// as with the exact-main-term tensor path.  The double-precision pipe issues one warp instruction per clock per SM (measured:
// 500 tanh per point are ~1 ms per 1 Mi points at 17 instructions), so the count matters: 16 + 1 MUFU instead of libdevice's ~70.
//   |x| >= 2^-6:  tanh x = 1 - 2 / (1 + E), E = e^(2x) = 2^n * exp(y), y = 2x - n ln2, |y| <= 0.36, degree-8 polynomial
//                 (remainder y^9/9! < 1e-9 * y / 2: what half an ulp of tanh tolerates through dE/E), reciprocal by rcp.approx
//                 (2^-20) + one Newton step (2^-40: < 6e-11 of tanh >= 2^-6)
//   |x| <  2^-6:  x (1 - x^2 (1/3 - x^2 (2/15 - 17 x^2 / 315))) in fp32: the bracket's rounding errors are scaled by x^2 < 2^-12
__device__ __forceinline__ float tanh_dp(float z) {
  const float az = fminf(fabsf(z), 10.0f);                          // tanh(10) rounds to 1 in fp32
  const float z2 = z * z;
  const float small = fmaf(z, -z2 * fmaf(-z2, fmaf(-z2, 0.053968254f, 0.13333334f), 0.33333334f), z);
  const float nf = rintf(az * 2.8853900817779268f);                 // any integer within ~0.52 of 2x log2(e) will do
  const float a2 = az + az;
  const double y = fma((double)nf, -0.6931471805599453, (double)a2);
#ifdef PINN_TANH_ESTRIN      // what-if: the same polynomial by Estrin's scheme (dependent chain 4 deep instead of 8, 3 more instructions)
  const double y2 = y * y, y4 = y2 * y2;
  const double e01 = 1.0 + y, e23 = fma(1.6666666666666666e-01, y, 0.5), e45 = fma(8.333333333333333e-03, y, 4.1666666666666664e-02),
               e67 = fma(1.984126984126984e-04, y, 1.388888888888889e-03);
  double p = fma(fma(e67, y2, e45), y4, fma(e23, y2, e01));
  p = fma(2.48015873015873e-05 * y4, y4, p);
#elif !defined(PINN_TANH_TAIL32)   // the whole degree-8 polynomial in double: 16 DP instructions + 1 MUFU
  double p = 2.48015873015873e-05;
  p = fma(p, y, 1.984126984126984e-04);
  p = fma(p, y, 1.388888888888889e-03);
  p = fma(p, y, 8.333333333333333e-03);
  p = fma(p, y, 4.1666666666666664e-02);
  p = fma(p, y, 1.6666666666666666e-01);
  p = fma(p, y, 0.5);
  p = fma(p, y, 1.0);
  p = fma(p, y, 1.0);
#else
  // -DPINN_TANH_TAIL32 (tried, not adopted): exp(y) = 1 + y + y^2/2 in double + the tail y^3 (1/6 + y/24 + .. + y^5/40320) in fp32
  // -- 12 DP instructions + 1 MUFU.  Host check over 11 M inputs in [2^-6, 10]: max error 0.536 ulp (0.509 with the full double
  // polynomial), 0.13 % of the results one ulp off the correctly rounded value (0.025 %).  On the GPU the kernels were no faster
  // (same box: 9.84 / 12.35 ms against 9.82 / 12.32 ms) and the gradient error at trained weights rose by 2-10 %.
  const float yf = fmaf(nf, -0.69314718f, a2);
  float tq = 2.48015873e-05f;
  tq = fmaf(tq, yf, 1.98412698e-04f);
  tq = fmaf(tq, yf, 1.38888889e-03f);
  tq = fmaf(tq, yf, 8.33333333e-03f);
  tq = fmaf(tq, yf, 4.16666667e-02f);
  tq = fmaf(tq, yf, 1.66666667e-01f);
  const float tail = (yf * yf) * (yf * tq);
  double p = fma(y, 0.5, 1.0);
  p = fma(p, y, (double)tail);
  p = p + 1.0;
#endif
  const double s = __hiloint2double(__double2hiint(p) + ((int)nf << 20), __double2loint(p)) + 1.0;     // 1 + e^(2x)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  r = fma(r, fma(-s, r, 1.0), r);
  const float big = copysignf((float)fma(-2.0, r, 1.0), z);
  return az < 0.015625f ? small : big;
}

// tanh for the tensor-core kernels' element-wise stages.
__device__ __forceinline__ float tanh_tc(float z) {
#ifdef PINN_TC_FAST_TANH
  // 1 - 2/(1 + e^{2z}) on the MUFU unit: 5 instructions, absolute error ~1e-7 (relative error grows as |z| -> 0)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
#else
  return tanhf(z);
#endif
}
// ACC: residual signatures of the pipelined kernel (see tanh_dp); boundary-term signatures keep tanhf
template <bool ACC>
__device__ __forceinline__ float tanh_tcx(float z) {
#if defined(PINN_TC_FAST_TANH) || defined(PINN_TC_TANHF)
  return tanh_tc(z);
#else
  if constexpr (ACC) return tanh_dp(z);
  else return tanh_tc(z);
#endif
}

__device__ __forceinline__ void tmem_ld4u(uint32_t taddr, float (&v)[4]) { umma::tmem_ld4(taddr, v); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ __forceinline__ void mbar_init_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(umma::smem_addr(bar)), "r"(n) : "memory");
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <class CFG>
__global__ void __launch_bounds__(CFG::NTH, 1) pinn_tc_kernel(const TcLaunchParams p) {
  using SIG = typename CFG::SIG;
  constexpr int C = CFG::C, D = CFG::D, M = CFG::M, H = CFG::H, T = CFG::T, R = CFG::R, NQ = CFG::NQ, JB = CFG::JB, ARR = CFG::ARR,
                NTH = CFG::NTH;
  constexpr bool BWD = CFG::BWD;

  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* arrA[3] = {smraw + 0 * ARR, smraw + 1 * ARR, smraw + 2 * ARR};   // A^{l} splits
  unsigned char* arrZ[3] = {smraw + 3 * ARR, smraw + 4 * ARR, smraw + 5 * ARR};   // Zbar^{l} splits (BWD)
  float* W1s = reinterpret_cast<float*>(smraw + CFG::oW1);
  float* bs = reinterpret_cast<float*>(smraw + CFG::oBias);
  float* W6s = reinterpret_cast<float*>(smraw + CFG::oW6);
  float* b6s = reinterpret_cast<float*>(smraw + CFG::oB6);
  float* xs = reinterpret_cast<float*>(smraw + CFG::oX);
  float* auxs = reinterpret_cast<float*>(smraw + CFG::oAux);
  float* lws = reinterpret_cast<float*>(smraw + CFG::oLw);
  float* uo = reinterpret_cast<float*>(smraw + CFG::oUo);
  float* ub = reinterpret_cast<float*>(smraw + CFG::oUb);
  float* rs = reinterpret_cast<float*>(smraw + CFG::oRs);
  float* redbuf = reinterpret_cast<float*>(smraw + CFG::oRed);
  double* lred = reinterpret_cast<double*>(smraw + CFG::oLred);
  uint64_t* bar0 = reinterpret_cast<uint64_t*>(smraw + CFG::oBar);        // forward / (ii) / output MMAs
  uint64_t* bar1 = bar0 + 1;                                              // dW MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar0 + 2);
  KindParams* kps = reinterpret_cast<KindParams*>(smraw + CFG::oKp);
  // direction weights of the Laplacian channel at tile point `pt`: per point (LapSig<D, true>) or the descriptor's constants
  auto lw_at = [&](int pt) -> const float* {
    if constexpr (SIG::kLapPW) return lws + pt * PINN_MAX_DIM;
    else return kps->lapw;
  };

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;         // TMEM lane quarter, column group
  const int j = 32 * q + lane;                   // neuron owned by this thread
#ifdef PINN_EXP_NO_ELEMENTWISE   // what-if: only MMAs, syncs, operand staging, flush
  const bool act = false;
#else
  const bool act = j < H;
#endif
  const uint32_t lane_off = (uint32_t)(32 * q) << 16;
  const int L = p.L;
  const int S = BWD ? p.n_seeds : 0;
  const long long P = p.P;

  const long long offB1 = (long long)H * D;
  const long long offHid = offB1 + H;
  const long long strideHid = (long long)H * H + H;
  const long long offW6 = offHid + (long long)(L - 1) * strideHid;
  const long long offB6 = offW6 + (long long)M * H;

  // ---------------- one-time setup ----------------
  for (int i = tid; i < H * D; i += NTH) W1s[i] = p.params[i];
  for (int i = tid; i < H; i += NTH) bs[i] = p.params[offB1 + i];
  for (int l = 2; l <= L; ++l)
    for (int i = tid; i < H; i += NTH) bs[(l - 1) * H + i] = p.params[offHid + (long long)(l - 2) * strideHid + (long long)H * H + i];
  for (int i = tid; i < M * H; i += NTH) W6s[i] = p.params[offW6 + i];
  if (tid < M) b6s[tid] = p.params[offB6 + tid];
  for (int i = tid; i < CFG::NARR * ARR / 16; i += NTH) reinterpret_cast<uint4*>(smraw)[i] = make_uint4(0u, 0u, 0u, 0u);   // pad rows stay zero
  if (tid == 0) {
    mbar_init_n(bar0, 1);
    mbar_init_n(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *kps = p.kp;
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, CFG::tcTotal);
  float* gp = nullptr;
  if constexpr (BWD) {
    gp = p.gpart + (size_t)blockIdx.x * S * p.Ppad;
    for (long long i = tid; i < (long long)S * p.Ppad; i += NTH) gp[i] = 0.f;
    if constexpr (CFG::RMW_FLUSH) {
      float4* gz = reinterpret_cast<float4*>(p.gwt) + (size_t)blockIdx.x * S * (L - 1) * kTcGwtLayer4;
      for (long long i = tid; i < (long long)S * (L - 1) * kTcGwtLayer4; i += NTH) gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __threadfence();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // shfl: lets ptxas treat TMEM addresses as warp-uniform
  const uint32_t tD = tbase + CFG::tcD, tDW = tbase + CFG::tcDW, tW = tbase + CFG::tcW;

  float4* stash = reinterpret_cast<float4*>(p.stash) + (size_t)blockIdx.x * L * NQ * C * NTH;
#ifdef PINN_EXP_NO_STASH   // what-if: no stash memory traffic at all (results wrong)
  float4 fake_stash_reg = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
  auto stash_at = [&](int l, int qi, int c) -> float4* { return &fake_stash_reg; };
#else
  auto stash_at = [&](int l /*1-based layer*/, int qi, int c) -> float4* { return stash + ((size_t)((l - 1) * NQ + qi) * C + c) * NTH + tid; };
#endif

  uint32_t ph0 = 0, ph1 = 0;
  bool w_ready = false;      // W_2 already staged in TMEM by the previous tile's tail
  double lacc[PINN_MAX_PDE] = {0.0, 0.0, 0.0};

  // W (128 x 112 fp32, zero padded, row j) -> three bf16 splits in TMEM columns tW + s*56 + k/2.
  // Split in two steps so the global-load latency hides behind the wait for the MMAs that still read the
  // previous W: fetch_W() right after an MMA batch is issued, put_W() once that batch has completed.
  uint32_t ws1[16], ws2[16], ws3[16];
  auto fetch_W = [&](const uint4* Wp) {
#ifdef PINN_EXP_NO_W
    return;
#endif
    const uint4* src = Wp + tid;                   // slot-major, thread-minor: every warp load is 512 contiguous bytes
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      const uint4 a = src[(0 * 4 + qd) * NTH], b = src[(1 * 4 + qd) * NTH], c = src[(2 * 4 + qd) * NTH];
      ws1[4 * qd] = a.x; ws1[4 * qd + 1] = a.y; ws1[4 * qd + 2] = a.z; ws1[4 * qd + 3] = a.w;
      ws2[4 * qd] = b.x; ws2[4 * qd + 1] = b.y; ws2[4 * qd + 2] = b.z; ws2[4 * qd + 3] = b.w;
      ws3[4 * qd] = c.x; ws3[4 * qd + 1] = c.y; ws3[4 * qd + 2] = c.z; ws3[4 * qd + 3] = c.w;
    }
  };
  auto put_W = [&]() {
#ifdef PINN_EXP_NO_W
    return;
#endif
    const uint32_t col = tW + lane_off + 16 * g;
    if (g < 3) {
      tmem_st16(col, ws1);
      tmem_st16(col + 56, ws2);
      tmem_st16(col + 112, ws3);
    } else {
      tmem_st8u(col, ws1);
      tmem_st8u(col + 56, ws2);
      tmem_st8u(col + 112, ws3);
    }
    umma::wait_st();
  };

  // four consecutive rows r0..r0+3 of neuron j -> the three split arrays (8-byte stores)
  auto store_quad = [&](unsigned char* const (&arr)[3], int r0, float v0, float v1, float v2, float v3) {
    uint32_t a1, a2, a3, b1, b2, b3;
    split3_pair(v0, v1, a1, a2, a3);
    split3_pair(v2, v3, b1, b2, b3);
    const int off = (j >> 3) * JB + (r0 >> 3) * 128 + (j & 7) * 16 + (r0 & 7) * 2;
#ifdef PINN_EXP_NO_STS      // what-if timing experiment only (results wrong)
    if (a1 == 0x12345678u) *reinterpret_cast<uint2*>(arr[0] + off) = make_uint2(a1 ^ a2 ^ a3, b1 ^ b2 ^ b3);
#else
    *reinterpret_cast<uint2*>(arr[0] + off) = make_uint2(a1, b1);
    *reinterpret_cast<uint2*>(arr[1] + off) = make_uint2(a2, b2);
    *reinterpret_cast<uint2*>(arr[2] + off) = make_uint2(a3, b3);
#endif
  };

  // D (+)= Wop[TMEM] * B[smem, MN-major]: 7 k-steps x 6 split products, small terms first.
  // Issue functions are called by all of warp 0 (see umma::mma_bf16_ts_e).
  auto issue_rows = [&](unsigned char* const (&barr)[3], uint64_t* bar) {
    const uint32_t el = umma::elect_one();
    constexpr uint32_t idesc = umma::make_idesc_bf16(kTcRows, R, 0, 1);
    const int wa[6] = {0, 2, 1, 0, 1, 0};
    const int ab[6] = {2, 0, 1, 1, 0, 0};
    // Split-product major, k-step minor: the five correction products are summed while the accumulator is still
    // ~2^-8 of its final size, so only the last 7 MMAs (b1*b1) round at full magnitude (the tensor core truncates
    // its fp32 accumulate; 42 full-magnitude truncations cost ~10x the FFMA path's error, 7 do not).
    const uint32_t lo0 = umma::smem_desc_lo(umma::smem_addr(barr[0]), JB);   // the three split arrays are ARR bytes apart
    constexpr uint32_t hi = umma::smem_desc_hi(128);
    uint32_t accum = 0;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
#pragma unroll
      for (int ks = 0; ks < kTcHP / 16; ++ks) {
#ifdef PINN_EXP_ONE_MMA     // what-if: (almost) no tensor work
        if (t != 5 || ks != 0) continue;
#endif
        umma::mma_bf16_ts_e2(tD, tW + wa[t] * 56 + ks * 8, lo0 + (uint32_t)((ab[t] * ARR + ks * 2 * JB) >> 4), hi, idesc, accum, el);
        accum = 1;
      }
    }
    umma::commit_e(bar, el);
  };
  // DW = Zbar[smem, K-major] * A[smem, K-major]^T : R/16 k-steps x 6 split products
  auto issue_dw = [&](uint64_t* bar) {
    constexpr uint32_t idesc = umma::make_idesc_bf16(kTcRows, kTcHP, 0, 0);
    const int za[6] = {0, 2, 1, 0, 1, 0};
    const int ab[6] = {2, 0, 1, 1, 0, 0};
    const uint32_t el = umma::elect_one();
    const uint32_t zlo0 = umma::smem_desc_lo(umma::smem_addr(arrZ[0]), 128);
    const uint32_t alo0 = umma::smem_desc_lo(umma::smem_addr(arrA[0]), 128);
    constexpr uint32_t hi = umma::smem_desc_hi(JB);
    uint32_t accum = 0;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
#pragma unroll
      for (int ks = 0; ks < R / 16; ++ks) {
#ifdef PINN_EXP_ONE_MMA
        if (t != 5 || ks != 0) continue;
#endif
        umma::mma_bf16_ss_e2(tDW, zlo0 + (uint32_t)((za[t] * ARR + ks * 256) >> 4), hi, alo0 + (uint32_t)((ab[t] * ARR + ks * 256) >> 4), hi, idesc,
                             accum, el);
        accum = 1;
      }
    }
    umma::commit_e(bar, el);
  };
  // all threads: publish generic-proxy smem writes + TMEM stores, sync, then one thread issues
  auto sync_for_mma = [&]() {
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  };
  auto wait_bar = [&](uint64_t* bar, uint32_t& ph) {
    mbar_wait(bar, ph);
    ph ^= 1;
    umma::fence_after_sync();
  };
  // sum a per-thread partial over the 4 column groups (fixed order) and RED it to one address
  auto reduce_groups_red = [&](float v, float* dst) {
#ifdef PINN_EXP_NO_RED
    if (v == 12345.f) atomicAdd(dst, v);
    return;
#endif
    redbuf[g * 128 + j] = v;
    __syncthreads();
    if (g == 0 && act) atomicAdd(dst, (redbuf[j] + redbuf[128 + j]) + (redbuf[256 + j] + redbuf[384 + j]));
    __syncthreads();
  };

  const long long ntiles = (p.n_rows + T - 1) / T;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * T;
    for (int i = tid; i < T * D; i += NTH) {
      const long long r = row0 + i / D;
      xs[i] = (r < p.n_rows) ? p.x[row0 * D + i] : 0.f;
    }
    if (p.n_aux > 0) {
      for (int i = tid; i < T * p.n_aux; i += NTH) {
        const int pt = i / p.n_aux, a = i % p.n_aux;
        const long long r = row0 + pt;
        auxs[pt * PINN_MAX_AUX + a] = (r < p.n_rows) ? p.aux[r * p.n_aux + a] : 0.f;
      }
    }
    if constexpr (SIG::kLapPW) {
      __syncthreads();                        // the weights read this tile's aux columns
      if (tid < T) lap_point_weights(*kps, auxs + tid * PINN_MAX_AUX, D, lws + tid * PINN_MAX_DIM);
    }
    __syncthreads();

    // ================= layer 1 (d -> H), CUDA cores =================
    if (act) {
      float w[D];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = W1s[j * D + i];
      const float b = bs[j];
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        if (4 * qi + g >= CFG::NQUADS) continue;      // warp-uniform: this group owns no quad in this slot
        const int pb = 4 * (4 * qi + g);
        float a[C][4];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          float z0 = b;
#pragma unroll
          for (int i = 0; i < D; ++i) z0 = fmaf(w[i], xs[(pb + pp) * D + i], z0);
#pragma unroll
          for (int c = 0; c < C; ++c) a[c][pp] = 0.f;
          a[0][pp] = z0;
          static_for<D>([&](auto J) {
            constexpr int jj = decltype(J)::value;
            if constexpr (SIG::order(jj) >= 1) a[SIG::base(jj)][pp] = w[jj];
          });
        }
        // the stash keeps t0 = tanh(z_0) in channel 0 instead of z_0 (nothing downstream needs z_0 itself)
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) a[0][pp] = tanh_tc(a[0][pp]);
        if constexpr (BWD) {
#pragma unroll
          for (int c = 0; c < C; ++c) *stash_at(1, qi, c) = make_float4(a[c][0], a[c][1], a[c][2], a[c][3]);
        }
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          float z[C], av[C];
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] = a[c][pp];
          tanh_jet_fwd_t0<SIG, true>(z, z[0], av, lw_at(pb + pp));
#pragma unroll
          for (int c = 0; c < C; ++c) a[c][pp] = av[c];
        }
#pragma unroll
        for (int c = 0; c < C; ++c) store_quad(arrA, c * T + pb, a[c][0], a[c][1], a[c][2], a[c][3]);
      }
    }

    // ================= layers 2..L (+ output layer as one more "layer" with W6 padded) =================
    if (!w_ready) {
      fetch_W(p.wn);
      put_W();
    }
    w_ready = false;
    for (int l = 2; l <= L + 1; ++l) {
      sync_for_mma();
      if (warp == 0) issue_rows(arrA, bar0);
      // operand of the next MMA batch: W_{l+1}, the padded output layer, or (reverse sweep) W_L^T
      const uint4* nextW = (l < L) ? p.wn + (size_t)(l - 1) * kTcWMat
                                   : (l == L) ? p.w6p
                                              : (BWD ? p.wt + (size_t)(L - 2) * kTcWMat
                                                     : (tile + gridDim.x < ntiles ? p.wn : nullptr));
      if (nextW != nullptr) fetch_W(nextW);
      wait_bar(bar0, ph0);
      if (nextW != nullptr) put_W();
      if (!BWD && l == L + 1 && nextW != nullptr) w_ready = true;
      if (l <= L) {
        const float b = act ? bs[(l - 1) * H + j] : 0.f;
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          if (4 * qi + g >= CFG::NQUADS) continue;      // warp-uniform: this group owns no quad in this slot
          const int pb = 4 * (4 * qi + g);
          float zq[C][4];
#pragma unroll
          for (int c = 0; c < C; ++c) umma::tmem_ld4(tD + lane_off + c * T + pb, zq[c]);
          umma::wait_ld();
          if (act) {
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) zq[0][pp] = tanh_tc(zq[0][pp] + b);      // channel 0 holds t0 from here on
            if constexpr (BWD) {
#pragma unroll
              for (int c = 0; c < C; ++c) *stash_at(l, qi, c) = make_float4(zq[c][0], zq[c][1], zq[c][2], zq[c][3]);
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              float z[C], av[C];
#pragma unroll
              for (int c = 0; c < C; ++c) z[c] = zq[c][pp];
              tanh_jet_fwd_t0<SIG>(z, z[0], av, lw_at(pb + pp));
#pragma unroll
              for (int c = 0; c < C; ++c) zq[c][pp] = av[c];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) store_quad(arrA, c * T + pb, zq[c][0], zq[c][1], zq[c][2], zq[c][3]);
          }
        }
      } else {
        // output jets: lanes o < M of D hold u[o][r]
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          if (4 * qi + g >= CFG::NQUADS) continue;      // warp-uniform: this group owns no quad in this slot
          const int pb = 4 * (4 * qi + g);
          float uq[C][4];
#pragma unroll
          for (int c = 0; c < C; ++c) umma::tmem_ld4(tD + lane_off + c * T + pb, uq[c]);
          umma::wait_ld();
          if (j < M) {
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
              for (int pp = 0; pp < 4; ++pp) uo[j * R + c * T + pb + pp] = uq[c][pp] + (c == 0 ? b6s[j] : 0.f);
          }
        }
      }
      umma::fence_before_sync();
      __syncthreads();
    }

    // ================= residual + loss =================
    if (tid < T) {
      float U[M][C];
#pragma unroll
      for (int o = 0; o < M; ++o)
#pragma unroll
        for (int c = 0; c < C; ++c) U[o][c] = uo[o * R + c * T + tid];
      float r[PINN_MAX_PDE];
      Residual<SIG, M>::eval(*kps, auxs + tid * PINN_MAX_AUX, U, r);
      const bool valid = row0 + tid < p.n_rows;
#pragma unroll
      for (int k = 0; k < PINN_MAX_PDE; ++k) {
        const float rk = valid ? r[k] : 0.f;
        rs[k * T + tid] = rk;
        lacc[k] += (double)rk * (double)rk;
      }
      if (p.resid_out != nullptr && valid)
        for (int k = 0; k < p.n_pde; ++k) p.resid_out[(row0 + tid) * p.n_pde + k] = r[k];
    }

    if constexpr (BWD) {
      for (int s = 0; s < S; ++s) {
        float* gs = gp + (size_t)s * p.Ppad;
        // ---- adjoint of the residual
        if (tid < T) {
          float U[M][C], Ub[M][C];
#pragma unroll
          for (int o = 0; o < M; ++o)
#pragma unroll
            for (int c = 0; c < C; ++c) U[o][c] = uo[o * R + c * T + tid];
          float rb[PINN_MAX_PDE];
#pragma unroll
          for (int k = 0; k < PINN_MAX_PDE; ++k) rb[k] = (k < p.n_pde) ? p.seeds[s * p.n_pde + k] * (kps->lin_seed ? (row0 + tid < p.n_rows ? 1.0f : 0.f) : 2.0f * rs[k * T + tid] * p.inv_n) : 0.f;
          Residual<SIG, M>::adjoint(*kps, auxs + tid * PINN_MAX_AUX, U, rb, Ub);
#pragma unroll
          for (int o = 0; o < M; ++o)
#pragma unroll
            for (int c = 0; c < C; ++c) ub[o * R + c * T + tid] = Ub[o][c];
        }
        __syncthreads();
        if (tid < M) {
          float sacc = 0.f;
          for (int pq = 0; pq < T; ++pq) sacc += ub[tid * R + pq];
          atomicAdd(gs + offB6 + tid, sacc);
        }

        // ---- layers L..1.  abar^l comes from ub/W6 (l == L) or from D (written by (ii) of layer l+1)
        for (int l = L; l >= 1; --l) {
          float dbias = 0.f;
          float dw6[M];
#pragma unroll
          for (int o = 0; o < M; ++o) dw6[o] = 0.f;
          float dw1[D];
#pragma unroll
          for (int i = 0; i < D; ++i) dw1[i] = 0.f;
#pragma unroll
          for (int qi = 0; qi < NQ; ++qi) {
            if (4 * qi + g >= CFG::NQUADS) continue;      // warp-uniform: this group owns no quad in this slot
            const int pb = 4 * (4 * qi + g);
            float ab[C][4];
            if (l < L) {
#pragma unroll
              for (int c = 0; c < C; ++c) umma::tmem_ld4(tD + lane_off + c * T + pb, ab[c]);
              umma::wait_ld();
            }
            if (act) {
              float zq[C][4];
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float4 v = *stash_at(l, qi, c);
                zq[c][0] = v.x; zq[c][1] = v.y; zq[c][2] = v.z; zq[c][3] = v.w;
              }
              if (l == L) {
                // abar^L = W6^T ubar ; dW6 = ubar a^L
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                  float z[C], av[C];
#pragma unroll
                  for (int c = 0; c < C; ++c) z[c] = zq[c][pp];
                  tanh_jet_fwd_t0<SIG>(z, z[0], av, lw_at(pb + pp));       // z[0] is the stashed t0
#pragma unroll
                  for (int c = 0; c < C; ++c) {
                    float v = 0.f;
#pragma unroll
                    for (int o = 0; o < M; ++o) {
                      const float u = ub[o * R + c * T + pb + pp];
                      v = fmaf(u, W6s[o * H + j], v);
                      dw6[o] = fmaf(u, av[c], dw6[o]);
                    }
                    ab[c][pp] = v;
                  }
                }
              }
#pragma unroll
              for (int pp = 0; pp < 4; ++pp) {
                float z[C], a_[C], zb[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  z[c] = zq[c][pp];
                  a_[c] = ab[c][pp];
                  zb[c] = 0.f;
                }
                tanh_jet_bwd_t0<SIG>(z, z[0], a_, zb, lw_at(pb + pp));  // z[0] is the stashed t0
                dbias += zb[0];
                if (l == 1) {
                  static_for<D>([&](auto J) {
                    constexpr int jj = decltype(J)::value;
                    dw1[jj] = fmaf(zb[0], xs[(pb + pp) * D + jj], dw1[jj]);
                    if constexpr (SIG::order(jj) >= 1) dw1[jj] += zb[SIG::base(jj)];
                  });
                }
#pragma unroll
                for (int c = 0; c < C; ++c) ab[c][pp] = zb[c];
              }
              if (l > 1) {
#pragma unroll
                for (int c = 0; c < C; ++c) store_quad(arrZ, c * T + pb, ab[c][0], ab[c][1], ab[c][2], ab[c][3]);
                // a^{l-1} from the stash
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  const float4 v = *stash_at(l - 1, qi, c);
                  zq[c][0] = v.x; zq[c][1] = v.y; zq[c][2] = v.z; zq[c][3] = v.w;
                }
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                  float z[C], av[C];
#pragma unroll
                  for (int c = 0; c < C; ++c) z[c] = zq[c][pp];
                  tanh_jet_fwd_t0<SIG>(z, z[0], av, lw_at(pb + pp));       // z[0] is the stashed t0
#pragma unroll
                  for (int c = 0; c < C; ++c) zq[c][pp] = av[c];
                }
#pragma unroll
                for (int c = 0; c < C; ++c) store_quad(arrA, c * T + pb, zq[c][0], zq[c][1], zq[c][2], zq[c][3]);
              }
            }
          }
          // small gradients: bias of layer l (+ output weights at l == L, first-layer weights at l == 1)
          float* gb = gs + ((l == 1) ? offB1 : offHid + (long long)(l - 2) * strideHid + (long long)H * H);
          reduce_groups_red(dbias, gb + j);
          if (l == L) {
#pragma unroll
            for (int o = 0; o < M; ++o) reduce_groups_red(dw6[o], gs + offW6 + o * H + j);
          }
          if (l == 1) {
#pragma unroll
            for (int i = 0; i < D; ++i) reduce_groups_red(dw1[i], gs + j * D + i);
            break;
          }
          // ---- tensor-core contractions of layer l (W_l^T is already in TMEM)
          sync_for_mma();
          if (warp == 0) {
            issue_dw(bar1);
            issue_rows(arrZ, bar0);
          }
          // next operand: W_{l-1}^T, or W_L^T again for the next seed, or W_2 for the next tile's forward
          const uint4* nextW = (l - 1 >= 2) ? p.wt + (size_t)(l - 3) * kTcWMat
                                            : (s + 1 < S) ? p.wt + (size_t)(L - 2) * kTcWMat
                                                          : (tile + gridDim.x < ntiles ? p.wn : nullptr);
          if (nextW != nullptr) fetch_W(nextW);
          if (l == 2 && s + 1 >= S && nextW != nullptr) w_ready = true;
          wait_bar(bar1, ph1);
          {
            const int k0 = 28 * g;
            float v[7][4];
#pragma unroll
            for (int i = 0; i < 7; ++i) umma::tmem_ld4(tDW + lane_off + k0 + 4 * i, v[i]);
            umma::wait_ld();
            if constexpr (CFG::RMW_FLUSH) {
              // Flush dW_l into this CTA's private partial by plain read-modify-write (single writer, so no atomics:
              // 10^4 scalar-equivalent REDs per layer cost ~10k cycles per tile).  Partial layout [k/4][n = j][k%4]:
              // consecutive lanes touch consecutive 16-byte words.
              float4* gt = reinterpret_cast<float4*>(p.gwt) + ((size_t)(blockIdx.x * S + s) * (L - 1) + (l - 2)) * kTcGwtLayer4;
              float4 old[7];
#pragma unroll
              for (int i = 0; i < 7; ++i)
                if (k0 + 4 * i < H) old[i] = gt[(size_t)(k0 / 4 + i) * 128 + j];
#pragma unroll
              for (int i = 0; i < 7; ++i)
                if (k0 + 4 * i < H)
                  gt[(size_t)(k0 / 4 + i) * 128 + j] = make_float4(old[i].x + v[i][0], old[i].y + v[i][1], old[i].z + v[i][2], old[i].w + v[i][3]);
            } else {
              float* gw = gs + offHid + (long long)(l - 2) * strideHid + (long long)j * H;
#pragma unroll
              for (int i = 0; i < 7; ++i)
                if (act && k0 + 4 * i < H) atomicAdd(reinterpret_cast<float4*>(gw + k0 + 4 * i), make_float4(v[i][0], v[i][1], v[i][2], v[i][3]));
            }
          }
          wait_bar(bar0, ph0);       // abar^{l-1} is now in D
          if (nextW != nullptr) put_W();
          umma::fence_before_sync();
          __syncthreads();
        }
      }
    }
    __syncthreads();
  }

  // ---------------- per-CTA loss partials ----------------
  if (tid < T) {
#pragma unroll
    for (int k = 0; k < PINN_MAX_PDE; ++k) lred[k * T + tid] = lacc[k];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid < PINN_MAX_PDE) {
    double ssum = 0.0;
    for (int pq = 0; pq < T; ++pq) ssum += lred[tid * T + pq];
    p.lpart[(size_t)blockIdx.x * PINN_MAX_PDE + tid] = ssum;
  }
  if (warp == 0) umma::tmem_dealloc(tbase, CFG::tcTotal);
}

}  // namespace pinn
