// pinn_kernels.cuh -- fused Taylor-jet MLP forward + PDE residual + MSE + hand-derived reverse (sm_100a).
//
// One persistent CTA per SM walks tiles of T collocation points.  Per tile:
//   forward : layer 1 on CUDA cores (skinny), layers 2..L as register-tiled FP32 FFMA contractions
//             Z[(c,p), n] = sum_k A[k][(c,p)] * W^T[k][n]  with every jet channel c of every point p a row;
//             the H x H weight matrix of each layer is staged into shared memory by one TMA bulk copy
//             (cp.async.bulk + mbarrier) that overlaps the element-wise tanh-jet stage of the previous layer;
//             pre-activations z^l are stashed to a per-CTA, L2-resident scratch slab.
//   residual: per point, from the output jets; r^2 accumulated in fp64 per CTA (deterministic).
//   reverse : per seed row, top-down: zbar^l from the stashed z^l, then
//             dW_l += Zbar^T A^{l-1}   (contraction over all (c,p) rows of the tile)  and
//             Abar^{l-1} = Zbar W_l    (same register tiling as the forward).
//             Weight / bias gradients go to a per-CTA partial buffer with RED.ADD (one writer per
//             address => deterministic); a second kernel sums the partials in fixed order.
//
// Shared-memory activation layout is neuron-major: buf[n][c*T + p], leading dimension LD == 4 (mod 32),
// so all three contractions read float4 along the point/row axis without bank conflicts.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "jet_math.cuh"

namespace pinn {

constexpr int kMaxLayers = 8;

struct LaunchParams {
  KindParams kp;
  int L;
  int n_aux;
  int n_pde;
  int n_seeds;
  long long n_rows;
  long long P;
  float inv_n;
  const float* params;   // flat, nn.Linear order
  const float* wpack;    // [(L-1) x H*H transposed | (L-1) x H*H native], 16B aligned
  const float* x;
  const float* aux;
  const float* seeds;    // [n_seeds, n_pde]
  float* stash;          // [grid][L][E4][NT] float4
  float* gpart;          // [grid][n_seeds][P]
  double* lpart;         // [grid][PINN_MAX_PDE]
  float* resid_out;      // nullable [n_rows, n_pde]
};

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  const uint32_t addr = smem_u32(bar);
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

// ------------------------------------------------------------------------------------------
// compile-time configuration
// ------------------------------------------------------------------------------------------
template <class SIG_, int M_, int H_, int T_, int TN_, bool BWD_>
struct KCfg {
  using SIG = SIG_;
  static constexpr int M = M_, H = H_, T = T_, TN = TN_;
  static constexpr bool BWD = BWD_;
  static constexpr int C = SIG::C, D = SIG::D, TP = 4;
  static constexpr int NPG = T / TP, NNG = H / TN, NT = NPG * NNG, NTP = (NT + 31) / 32 * 32;
  static constexpr int CT = C * T;
  static constexpr int LD = CT + ((4 - CT % 32 + 32) % 32);
  static constexpr int E4 = C * TN;  // float4 (= 4 points) per thread per layer
  // outer-product (dW) thread tiling
  static constexpr int OTK = (H % 10 == 0) ? 10 : 8, OTN = NT / OTK, ON = H / OTN, OK = H / OTK;     // H = 100: 10 x 20 threads; H = 64, 128: 8 x 16 / 8 x 32
  // shared memory carve-up, in floats
  static constexpr int oWs = 0;
  static constexpr int oA = oWs + H * H;
  static constexpr int oB = oA + H * LD;
  static constexpr int oW1 = oB + (BWD ? H * LD : 0);
  static constexpr int oBias = oW1 + H * D;
  static constexpr int oW6 = oBias + kMaxLayers * H;
  static constexpr int oB6 = oW6 + M * H;
  static constexpr int oX = oB6 + 4;
  static constexpr int oAux = oX + T * D;
  static constexpr int oUo = oAux + T * PINN_MAX_AUX;
  static constexpr int oUb = oUo + M * CT;
  static constexpr int oRs = oUb + M * CT;
  static constexpr int oLred = (oRs + PINN_MAX_PDE * T + 1) / 2 * 2;      // doubles, 8B aligned
  static constexpr int oBar = oLred + 2 * PINN_MAX_PDE * T;
  static constexpr int oKp = oBar + 2;
  static constexpr int nFloats = oKp + (int)((sizeof(KindParams) + 3) / 4);
  static constexpr size_t smem_bytes = (size_t)nFloats * 4;

  static_assert(T % TP == 0 && H % TN == 0, "tile shape");
  static_assert(TN == 2 || TN == 4, "TN");
  static_assert(NPG == 4 || NPG == 8, "point groups must fit a power-of-two lane group");
  static_assert(NT % OTK == 0 && H % OTN == 0 && H % OTK == 0, "outer-product tiling");
  static_assert((H * H * 4) % 16 == 0 && H % 4 == 0, "TMA bulk copy granularity");
  static_assert(LD % 4 == 0 && LD % 32 == 4, "LD");
  static_assert(oA % 4 == 0 && oB % 4 == 0 && oUb % 4 == 0 && oUo % 4 == 0, "float4 alignment");
  static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
};

template <int TN>
struct VecLoad;
template <>
struct VecLoad<4> {
  __device__ __forceinline__ static void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct VecLoad<2> {
  __device__ __forceinline__ static void ld(const float* p, float (&v)[2]) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
};

// acc[c][p][n] += sum_k In[k][c*T + 4*pg + p] * Mat[k][TN*ng + n]
template <class CFG>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ In, const float* __restrict__ Mat,
                                          float (&acc)[CFG::C][4][CFG::TN], int pg, int ng) {
  constexpr int C = CFG::C, T = CFG::T, TN = CFG::TN, H = CFG::H, LD = CFG::LD;
  const float* ap = In + 4 * pg;
  const float* mp = Mat + TN * ng;
#pragma unroll 2
  for (int k = 0; k < H; ++k) {
    float mv[TN];
    VecLoad<TN>::ld(mp + k * H, mv);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 av = *reinterpret_cast<const float4*>(ap + k * LD + c * T);
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        acc[c][0][n] = fmaf(av.x, mv[n], acc[c][0][n]);
        acc[c][1][n] = fmaf(av.y, mv[n], acc[c][1][n]);
        acc[c][2][n] = fmaf(av.z, mv[n], acc[c][2][n]);
        acc[c][3][n] = fmaf(av.w, mv[n], acc[c][3][n]);
      }
    }
  }
}

// g[n][k] += sum_r Zb[n][r] * Ap[k][r]   (r over all C*T rows of the tile), accumulated with RED.ADD
template <class CFG>
__device__ __forceinline__ void gemm_outer_red(const float* __restrict__ Zb, const float* __restrict__ Ap, float* __restrict__ g, int tid) {
  constexpr int ON = CFG::ON, OK = CFG::OK, OTN = CFG::OTN, OTK = CFG::OTK, LD = CFG::LD, CT = CFG::CT, H = CFG::H;
  const int tn = tid % OTN, tk = tid / OTN;
  float o[ON][OK];
#pragma unroll
  for (int i = 0; i < ON; ++i)
#pragma unroll
    for (int j = 0; j < OK; ++j) o[i][j] = 0.f;
  const float* zp = Zb + tn * LD;
  const float* pp = Ap + tk * LD;
#pragma unroll 1
  for (int r = 0; r < CT; r += 4) {
    float4 zv[ON], pv[OK];
#pragma unroll
    for (int i = 0; i < ON; ++i) zv[i] = *reinterpret_cast<const float4*>(zp + i * OTN * LD + r);
#pragma unroll
    for (int j = 0; j < OK; ++j) pv[j] = *reinterpret_cast<const float4*>(pp + j * OTK * LD + r);
#pragma unroll
    for (int i = 0; i < ON; ++i)
#pragma unroll
      for (int j = 0; j < OK; ++j) {
        o[i][j] = fmaf(zv[i].x, pv[j].x, o[i][j]);
        o[i][j] = fmaf(zv[i].y, pv[j].y, o[i][j]);
        o[i][j] = fmaf(zv[i].z, pv[j].z, o[i][j]);
        o[i][j] = fmaf(zv[i].w, pv[j].w, o[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < ON; ++i)
#pragma unroll
    for (int j = 0; j < OK; ++j) atomicAdd(g + (tn + i * OTN) * H + (tk + j * OTK), o[i][j]);
}

template <class CFG>
__device__ __forceinline__ void frag_store_smem(float* buf, const float (&acc)[CFG::C][4][CFG::TN], int pg, int ng) {
#pragma unroll
  for (int n = 0; n < CFG::TN; ++n)
#pragma unroll
    for (int c = 0; c < CFG::C; ++c)
      *reinterpret_cast<float4*>(buf + (CFG::TN * ng + n) * CFG::LD + c * CFG::T + 4 * pg) =
          make_float4(acc[c][0][n], acc[c][1][n], acc[c][2][n], acc[c][3][n]);
}

template <class CFG>
__device__ __forceinline__ void frag_store_stash(float4* st, const float (&acc)[CFG::C][4][CFG::TN], int tid) {
#pragma unroll
  for (int n = 0; n < CFG::TN; ++n)
#pragma unroll
    for (int c = 0; c < CFG::C; ++c)
      st[(size_t)(n * CFG::C + c) * CFG::NT + tid] = make_float4(acc[c][0][n], acc[c][1][n], acc[c][2][n], acc[c][3][n]);
}

// PRECISE: layer 1 (its derivative channels are the weight columns, the same for every point: see tanh_jet_fwd_t0)
template <class CFG, bool PRECISE = false>
__device__ __forceinline__ void frag_tanh_fwd(float (&acc)[CFG::C][4][CFG::TN], const float* lw) {
#pragma unroll
  for (int n = 0; n < CFG::TN; ++n)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float z[CFG::C], a[CFG::C];
#pragma unroll
      for (int c = 0; c < CFG::C; ++c) z[c] = acc[c][p][n];
      tanh_jet_fwd<typename CFG::SIG, PRECISE>(z, a, lw);
#pragma unroll
      for (int c = 0; c < CFG::C; ++c) acc[c][p][n] = a[c];
    }
}

// reload z^l from the stash and turn it into a^l in registers
template <class CFG>
__device__ __forceinline__ void frag_load_tanh(const float4* st, float (&acc)[CFG::C][4][CFG::TN], int tid, const float* lw) {
#pragma unroll
  for (int n = 0; n < CFG::TN; ++n) {
    float4 zv[CFG::C];
#pragma unroll
    for (int c = 0; c < CFG::C; ++c) zv[c] = st[(size_t)(n * CFG::C + c) * CFG::NT + tid];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float z[CFG::C], a[CFG::C];
#pragma unroll
      for (int c = 0; c < CFG::C; ++c) z[c] = (p == 0) ? zv[c].x : (p == 1) ? zv[c].y : (p == 2) ? zv[c].z : zv[c].w;
      tanh_jet_fwd<typename CFG::SIG>(z, a, lw);
#pragma unroll
      for (int c = 0; c < CFG::C; ++c) acc[c][p][n] = a[c];
    }
  }
}

// acc holds abar^l on entry and zbar^l on exit
template <class CFG>
__device__ __forceinline__ void frag_tanh_bwd(const float4* st, float (&acc)[CFG::C][4][CFG::TN], int tid, const float* lw) {
#pragma unroll
  for (int n = 0; n < CFG::TN; ++n) {
    float4 zv[CFG::C];
#pragma unroll
    for (int c = 0; c < CFG::C; ++c) zv[c] = st[(size_t)(n * CFG::C + c) * CFG::NT + tid];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float z[CFG::C], ab[CFG::C], zb[CFG::C];
#pragma unroll
      for (int c = 0; c < CFG::C; ++c) {
        z[c] = (p == 0) ? zv[c].x : (p == 1) ? zv[c].y : (p == 2) ? zv[c].z : zv[c].w;
        ab[c] = acc[c][p][n];
        zb[c] = 0.f;
      }
      tanh_jet_bwd<typename CFG::SIG>(z, ab, zb, lw);
#pragma unroll
      for (int c = 0; c < CFG::C; ++c) acc[c][p][n] = zb[c];
    }
  }
}

template <int NPG>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
  v += __shfl_xor_sync(mask, v, 1);
  v += __shfl_xor_sync(mask, v, 2);
  if (NPG == 8) v += __shfl_xor_sync(mask, v, 4);
  return v;
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <class CFG>
__global__ void __launch_bounds__(CFG::NTP, 1) pinn_fused_kernel(const LaunchParams p) {
  using SIG = typename CFG::SIG;
  constexpr int C = CFG::C, D = CFG::D, M = CFG::M, H = CFG::H, T = CFG::T, TN = CFG::TN, NPG = CFG::NPG, NT = CFG::NT,
                NTP = CFG::NTP, CT = CFG::CT, LD = CFG::LD, E4 = CFG::E4;
  constexpr bool BWD = CFG::BWD;
  constexpr uint32_t kWBytes = H * H * 4;

  extern __shared__ __align__(128) float smem[];
  float* Ws = smem + CFG::oWs;
  float* bufA = smem + CFG::oA;
  float* bufB = smem + CFG::oB;
  float* W1s = smem + CFG::oW1;
  float* bs = smem + CFG::oBias;
  float* W6s = smem + CFG::oW6;
  float* b6s = smem + CFG::oB6;
  float* xs = smem + CFG::oX;
  float* auxs = smem + CFG::oAux;
  float* uo = smem + CFG::oUo;
  float* ub = smem + CFG::oUb;
  float* rs = smem + CFG::oRs;
  double* lred = reinterpret_cast<double*>(smem + CFG::oLred);
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + CFG::oBar);
  KindParams* kps = reinterpret_cast<KindParams*>(smem + CFG::oKp);

  const int tid = threadIdx.x;
  const bool active = tid < NT;
  const int pg = tid % NPG, ng = tid / NPG;
  const unsigned amask = __ballot_sync(0xffffffffu, active);
  const int L = p.L;
  const int S = BWD ? p.n_seeds : 0;
  const long long P = p.P;

  // parameter offsets (nn.Linear order)
  const long long offB1 = (long long)H * D;
  const long long offHid = offB1 + H;                       // W_2
  const long long strideHid = (long long)H * H + H;
  const long long offW6 = offHid + (long long)(L - 1) * strideHid;
  const long long offB6 = offW6 + (long long)M * H;

  for (int i = tid; i < H * D; i += NTP) W1s[i] = p.params[i];
  for (int i = tid; i < H; i += NTP) bs[i] = p.params[offB1 + i];
  for (int l = 2; l <= L; ++l)
    for (int i = tid; i < H; i += NTP) bs[(l - 1) * H + i] = p.params[offHid + (long long)(l - 2) * strideHid + (long long)H * H + i];
  for (int i = tid; i < M * H; i += NTP) W6s[i] = p.params[offW6 + i];
  if (tid < M) b6s[tid] = p.params[offB6 + tid];
  if (tid == 0) {
    mbar_init(wbar, 1);
    *kps = p.kp;
  }

  float* gp = nullptr;
  if constexpr (BWD) {
    gp = p.gpart + (size_t)blockIdx.x * S * P;
    for (long long i = tid; i < (long long)S * P; i += NTP) gp[i] = 0.f;
    __threadfence();
  }
  __syncthreads();

  const float* wT = p.wpack;                                  // transposed W_2..W_L
  const float* wN = p.wpack + (size_t)(L - 1) * H * H;        // native W_2..W_L
  float4* stash = reinterpret_cast<float4*>(p.stash) + (size_t)blockIdx.x * L * E4 * NT;
  constexpr size_t kSlot = (size_t)E4 * NT;

  double lacc[PINN_MAX_PDE] = {0.0, 0.0, 0.0};
  uint32_t wphase = 0;
  const long long ntiles = (p.n_rows + T - 1) / T;
  if (tid == 0 && (long long)blockIdx.x < ntiles) tma_load_1d(Ws, wT, kWBytes, wbar);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * T;
    const bool more_tiles = tile + gridDim.x < ntiles;
    for (int i = tid; i < T * D; i += NTP) {
      const long long r = row0 + i / D;
      xs[i] = (r < p.n_rows) ? p.x[row0 * D + i] : 0.f;
    }
    if (p.n_aux > 0) {
      for (int i = tid; i < T * p.n_aux; i += NTP) {
        const int pt = i / p.n_aux, a = i % p.n_aux;
        const long long r = row0 + pt;
        auxs[pt * PINN_MAX_AUX + a] = (r < p.n_rows) ? p.aux[r * p.n_aux + a] : 0.f;
      }
    }
    __syncthreads();

    float acc[C][4][TN];

    // ---------------- layer 1 (d -> H): skinny, CUDA cores ----------------
    if (active) {
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        const int nn = TN * ng + n;
        float w[D];
#pragma unroll
        for (int j = 0; j < D; ++j) w[j] = W1s[nn * D + j];
        const float b = bs[nn];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const int pt = 4 * pg + pp;
          float z0 = b;
#pragma unroll
          for (int j = 0; j < D; ++j) z0 = fmaf(w[j], xs[pt * D + j], z0);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c][pp][n] = 0.f;
          acc[0][pp][n] = z0;
          static_for<D>([&](auto J) {
            constexpr int j = decltype(J)::value;
            if constexpr (SIG::order(j) >= 1) acc[SIG::base(j)][pp][n] = w[j];
          });
        }
      }
      if constexpr (BWD) frag_store_stash<CFG>(stash, acc, tid);
      frag_tanh_fwd<CFG, true>(acc, kps->lapw);
      frag_store_smem<CFG>(bufA, acc, pg, ng);
    }
    __syncthreads();

    // ---------------- layers 2..L (H -> H): FFMA contractions ----------------
    for (int l = 2; l <= L; ++l) {
      mbar_wait(wbar, wphase);
      wphase ^= 1;
      if (active) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp)
#pragma unroll
            for (int n = 0; n < TN; ++n) acc[c][pp][n] = 0.f;
        gemm_rows<CFG>(bufA, Ws, acc, pg, ng);
#pragma unroll
        for (int n = 0; n < TN; ++n) {
          const float b = bs[(l - 1) * H + TN * ng + n];
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) acc[0][pp][n] += b;
        }
      }
      __syncthreads();   // every read of bufA and Ws is done
      if (tid == 0) {
        if (l < L) tma_load_1d(Ws, wT + (size_t)(l - 1) * H * H, kWBytes, wbar);
        else if (BWD) tma_load_1d(Ws, wN + (size_t)(L - 2) * H * H, kWBytes, wbar);
        else if (more_tiles) tma_load_1d(Ws, wT, kWBytes, wbar);
      }
      if (active) {
        if constexpr (BWD) frag_store_stash<CFG>(stash + (size_t)(l - 1) * kSlot, acc, tid);
        frag_tanh_fwd<CFG>(acc, kps->lapw);
        frag_store_smem<CFG>(bufA, acc, pg, ng);
      }
      __syncthreads();
    }

    // ---------------- output layer (H -> m) on the jets ----------------
    for (int r = tid; r < CT; r += NTP) {
      float s[M];
#pragma unroll
      for (int o = 0; o < M; ++o) s[o] = (r < T) ? b6s[o] : 0.f;
      for (int n = 0; n < H; ++n) {
        const float a = bufA[n * LD + r];
#pragma unroll
        for (int o = 0; o < M; ++o) s[o] = fmaf(W6s[o * H + n], a, s[o]);
      }
#pragma unroll
      for (int o = 0; o < M; ++o) uo[o * CT + r] = s[o];
    }
    __syncthreads();

    // ---------------- residual + loss ----------------
    if (tid < T) {
      float U[M][C];
#pragma unroll
      for (int o = 0; o < M; ++o)
#pragma unroll
        for (int c = 0; c < C; ++c) U[o][c] = uo[o * CT + c * T + tid];
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
        float* gs = gp + (size_t)s * P;
        if (s > 0) {
          // bufA was consumed by the previous reverse sweep: rebuild a^L from the stash
          if (active) {
            frag_load_tanh<CFG>(stash + (size_t)(L - 1) * kSlot, acc, tid, kps->lapw);
            frag_store_smem<CFG>(bufA, acc, pg, ng);
          }
        }
        // ---- adjoint of the residual: ubar[o][c][p]
        if (tid < T) {
          float U[M][C], Ub[M][C];
#pragma unroll
          for (int o = 0; o < M; ++o)
#pragma unroll
            for (int c = 0; c < C; ++c) U[o][c] = uo[o * CT + c * T + tid];
          float rb[PINN_MAX_PDE];
#pragma unroll
          for (int k = 0; k < PINN_MAX_PDE; ++k)
            rb[k] = (k < p.n_pde) ? p.seeds[s * p.n_pde + k] * (kps->lin_seed ? (row0 + tid < p.n_rows ? 1.0f : 0.f) : 2.0f * rs[k * T + tid] * p.inv_n) : 0.f;
          Residual<SIG, M>::adjoint(*kps, auxs + tid * PINN_MAX_AUX, U, rb, Ub);
#pragma unroll
          for (int o = 0; o < M; ++o)
#pragma unroll
            for (int c = 0; c < C; ++c) ub[o * CT + c * T + tid] = Ub[o][c];
        }
        __syncthreads();

        // ---- output layer: dW_{L+1}, db_{L+1}, abar^L
        for (int idx = tid; idx < M * H; idx += NTP) {
          const int o = idx / H, n = idx % H;
          float sacc = 0.f;
          const float* ar = bufA + n * LD;
          const float* ur = ub + o * CT;
#pragma unroll 4
          for (int r = 0; r < CT; r += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(ar + r);
            const float4 u4 = *reinterpret_cast<const float4*>(ur + r);
            sacc = fmaf(a4.x, u4.x, sacc);
            sacc = fmaf(a4.y, u4.y, sacc);
            sacc = fmaf(a4.z, u4.z, sacc);
            sacc = fmaf(a4.w, u4.w, sacc);
          }
          atomicAdd(gs + offW6 + idx, sacc);
        }
        if (tid < M) {
          float sacc = 0.f;
          for (int q = 0; q < T; ++q) sacc += ub[tid * CT + q];
          atomicAdd(gs + offB6 + tid, sacc);
        }
        if (active) {
#pragma unroll
          for (int c = 0; c < C; ++c)
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
              for (int n = 0; n < TN; ++n) {
                float v = 0.f;
#pragma unroll
                for (int o = 0; o < M; ++o) v = fmaf(ub[o * CT + c * T + 4 * pg + pp], W6s[o * H + TN * ng + n], v);
                acc[c][pp][n] = v;
              }
        }
        __syncthreads();   // bufA (a^L) and ub are free again

        // ---- hidden layers, top-down
        for (int l = L; l >= 1; --l) {
          if (active) {
            frag_tanh_bwd<CFG>(stash + (size_t)(l - 1) * kSlot, acc, tid, kps->lapw);     // acc: abar^l -> zbar^l
            // bias gradient: sum over points of the value-channel adjoint
            float* gb = gs + ((l == 1) ? offB1 : offHid + (long long)(l - 2) * strideHid + (long long)H * H);
#pragma unroll
            for (int n = 0; n < TN; ++n) {
              float v = (acc[0][0][n] + acc[0][1][n]) + (acc[0][2][n] + acc[0][3][n]);
              v = group_sum<NPG>(v, amask);
              if (pg == 0) atomicAdd(gb + TN * ng + n, v);
            }
            if (l == 1) {
              // dW_1[n][j] = sum_p zbar_0 x_j + zbar_{1,(j)}
#pragma unroll
              for (int n = 0; n < TN; ++n) {
                static_for<D>([&](auto J) {
                  constexpr int j = decltype(J)::value;
                  float v = 0.f;
#pragma unroll
                  for (int pp = 0; pp < 4; ++pp) {
                    v = fmaf(acc[0][pp][n], xs[(4 * pg + pp) * D + j], v);
                    if constexpr (SIG::order(j) >= 1) v += acc[SIG::base(j)][pp][n];
                  }
                  v = group_sum<NPG>(v, amask);
                  if (pg == 0) atomicAdd(gs + (TN * ng + n) * D + j, v);
                });
              }
            } else {
              frag_store_smem<CFG>(bufB, acc, pg, ng);
              frag_load_tanh<CFG>(stash + (size_t)(l - 2) * kSlot, acc, tid, kps->lapw);   // a^{l-1}
              frag_store_smem<CFG>(bufA, acc, pg, ng);
            }
          }
          if (l == 1) break;
          __syncthreads();
          mbar_wait(wbar, wphase);     // native W_l
          wphase ^= 1;
          if (active) {
            gemm_outer_red<CFG>(bufB, bufA, gs + offHid + (long long)(l - 2) * strideHid, tid);
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
              for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int n = 0; n < TN; ++n) acc[c][pp][n] = 0.f;
            gemm_rows<CFG>(bufB, Ws, acc, pg, ng);
          }
          __syncthreads();
          if (tid == 0) {
            if (l - 1 >= 2) tma_load_1d(Ws, wN + (size_t)(l - 3) * H * H, kWBytes, wbar);
            else if (s + 1 < S) tma_load_1d(Ws, wN + (size_t)(L - 2) * H * H, kWBytes, wbar);
            else if (more_tiles) tma_load_1d(Ws, wT, kWBytes, wbar);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---------------- per-CTA loss partials (fixed order => deterministic) ----------------
  if (tid < T) {
#pragma unroll
    for (int k = 0; k < PINN_MAX_PDE; ++k) lred[k * T + tid] = lacc[k];
  }
  __syncthreads();
  if (tid < PINN_MAX_PDE) {
    double s = 0.0;
    for (int q = 0; q < T; ++q) s += lred[tid * T + q];
    p.lpart[(size_t)blockIdx.x * PINN_MAX_PDE + tid] = s;
  }
}

}  // namespace pinn
