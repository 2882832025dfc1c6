// jet_math.cuh -- per-(point, neuron) Taylor-jet arithmetic and per-point PDE residuals.
//
// Everything here is a pure function of registers, usable from device code and (for the host
// harness under tests/host_harness, which checks the hand-derived adjoints against autograd
// without a GPU) from plain C++.
//
// Jet layout of one neuron: channel 0 is the value; then, for every input direction j in
// order, the Taylor coefficients c_k = (1/k!) d^k/ds^k f(x + s e_j), k = 1..K_j.
// (replaces the repeated reverse passes of deepxde/gradients.py:47-52,222-228.)
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/pinn_b200.h"

#if defined(__CUDACC__)
#define PINN_HD __host__ __device__ __forceinline__
#else
#define PINN_HD inline
#endif

namespace pinn {

// Multiply / add / subtract that the compiler may not contract into an fma: the error-free transformations below rely on the
// rounding of each step (nvcc fuses a * b + c by default).  The host build has no fma contraction to begin with.
#if defined(__CUDA_ARCH__)
#define PINN_MUL(a, b) __fmul_rn((a), (b))
#define PINN_ADD(a, b) __fadd_rn((a), (b))
#define PINN_SUB(a, b) __fsub_rn((a), (b))
#else
#define PINN_MUL(a, b) ((a) * (b))
#define PINN_ADD(a, b) ((a) + (b))
#define PINN_SUB(a, b) ((a) - (b))
#endif

// ---------------------------------------------------------------------------------------------
// compile-time jet signature
// ---------------------------------------------------------------------------------------------
template <int N>
struct IC {
  static constexpr int value = N;
  PINN_HD constexpr operator int() const { return N; }
};

template <int I, int N, class F>
PINN_HD void static_for_impl(F&& f) {
  if constexpr (I < N) {
    f(IC<I>{});
    static_for_impl<I + 1, N>(f);
  }
}
template <int N, class F>
PINN_HD void static_for(F&& f) {
  static_for_impl<0, N>(f);
}

template <int... KS>
struct JetSig {
  static constexpr bool kLap = false;
  static constexpr bool kLapPW = false;
  static constexpr int D = sizeof...(KS);
  static constexpr int C = 1 + (0 + ... + KS);
  // residual signatures (some direction beyond first order) take the double-evaluated tanh in the FFMA kernel; the
  // boundary-term signatures (VALUE: all 0, FLUX: all 1), which always run on the FFMA kernel, keep tanhf (see pinn_tanh)
  static constexpr bool kAccTanh = ((KS >= 2) || ...);
  PINN_HD static constexpr int order(int j) {
    constexpr int o[sizeof...(KS)] = {KS...};
    return o[j];
  }
  // channel of the first-order coefficient of direction j
  PINN_HD static constexpr int base(int j) {
    int b = 1;
    for (int i = 0; i < j; ++i) b += order(i);
    return b;
  }
  PINN_HD static constexpr int chan(int j, int k) { return base(j) + k - 1; }
};

// Forward-Laplacian signature (SURVEY.md §8d "forward-Laplacian collapse C -> d+2"): for residuals that use the second
// derivatives only through one weighted sum  s = sum_j w_j d2u/dx_j^2  (every LINEAR-kind class: Poisson, heat, wave,
// Helmholtz), carry per neuron  [value, du/dx_1 .. du/dx_D, s]  instead of separate second-order coefficients per
// direction.  Linear layers act per channel as before; tanh couples them:
//   a_0 = t,  a_j = t' z_j,  a_s = t' z_s + t'' sum_j w_j z_j^2      (t = tanh z_0, t' = 1 - t^2, t'' = -2 t t').
// PW_: the direction weights carry a per-point factor (lap_point_weights) -- a separate instantiation, so that the
// constant-weight kernels keep their loop-invariant weight loads.
template <int D_, bool PW_ = false>
struct LapSig {
  static constexpr bool kLap = true;
  static constexpr bool kLapPW = PW_;
  static constexpr bool kAccTanh = true;
  static constexpr int D = D_;
  static constexpr int C = D_ + 2;
  static constexpr int lap = D_ + 1;                        // channel of s
  PINN_HD static constexpr int order(int) { return 1; }     // every direction carries its first derivative
  PINN_HD static constexpr int base(int j) { return 1 + j; }
  PINN_HD static constexpr int chan(int j, int) { return 1 + j; }
};

// ---------------------------------------------------------------------------------------------
// tanh on jets.  With t = tanh(z), w = 1 - t^2 as power series in s:
//   t_k = (1/k) sum_{i=1..k} i z_i w_{k-i},   w_k = - sum_{i=0..k} t_i t_{k-i}
// One transcendental per neuron (t_0).
// ---------------------------------------------------------------------------------------------
template <int K>
PINN_HD void tanh_series(const float (&zz)[K + 1], float t0, float w0, float (&t)[K + 1], float (&w)[K + 1]) {
  t[0] = t0;
  w[0] = w0;
#pragma unroll
  for (int k = 1; k <= K; ++k) {
    float s = 0.f;
#pragma unroll
    for (int i = 1; i <= k; ++i) s = fmaf((float)i * zz[i], w[k - i], s);
    t[k] = s * (1.0f / (float)k);
    float ww = 0.f;
#pragma unroll
    for (int i = 0; i <= k; ++i) ww = fmaf(t[i], t[k - i], ww);
    w[k] = -ww;
  }
}

// t0 = tanh(z[0]) supplied by the caller (a[0] of the forward pass, stashed by the tensor-core kernel)
// `lw`: the D direction weights of a forward-Laplacian signature (LapSig); ignored by JetSig signatures.
// PRECISE (forward-Laplacian signatures, layer 1 of the forward sweep): q = sum_j w_j z_j^2 enters a_s as an unevaluated fp32 pair.
template <class SIG, bool PRECISE = false>
PINN_HD void tanh_jet_fwd_t0(const float (&z)[SIG::C], float t0, float (&a)[SIG::C], const float* lw = nullptr);
template <class SIG>
PINN_HD void tanh_jet_bwd_t0(const float (&z)[SIG::C], float t0, const float (&abar)[SIG::C], float (&zbar)[SIG::C], const float* lw = nullptr);

// tanh of a pre-activation for the FP32 FFMA kernel (and the host build of this header): evaluated in double and rounded
// once.  Near a minimum the parameter gradient is a sum of per-point terms orders of magnitude larger than the result, and
// one or two ulp of (slightly biased) error in tanh then decide the last digits: on the nearly converged Laplace problem the
// gradient's distance from the fp64 reference drops from 1.3e-5 (tanhf) to 5e-7 with this, below the reference's own fp32
// evaluation (1.4e-6); profiles/r1_lap_summary.md.  Costs the FFMA kernel 19 %, so the boundary-term signatures (which
// always run on it, inside the default path) stay on tanhf: SIG::kAccTanh.  The tensor-core kernels keep tanhf
// (tanh_tc): their error there is set by the tensor core's truncating accumulate, not by tanh.
// -DPINN_TANH_MODE=0 restores tanhf, 2 = expm1f-based t/(t+2).
#ifndef PINN_TANH_MODE
#define PINN_TANH_MODE 1
#endif
PINN_HD float pinn_tanh(float z) {
#if PINN_TANH_MODE == 1
  return (float)tanh((double)z);
#elif PINN_TANH_MODE == 2
  const float t = expm1f(2.0f * z);
  return (z > 9.0f) ? 1.0f : t / (t + 2.0f);
#else
  return tanhf(z);
#endif
}

template <class SIG, bool PRECISE = false>
PINN_HD void tanh_jet_fwd(const float (&z)[SIG::C], float (&a)[SIG::C], const float* lw = nullptr) {
#ifdef PINN_EXP_NO_JET
  tanh_jet_fwd_t0<SIG, PRECISE>(z, z[0], a, lw);
#else
  tanh_jet_fwd_t0<SIG, PRECISE>(z, SIG::kAccTanh ? pinn_tanh(z[0]) : tanhf(z[0]), a, lw);
#endif
}

template <class SIG, bool PRECISE>
PINN_HD void tanh_jet_fwd_t0(const float (&z)[SIG::C], float t0, float (&a)[SIG::C], const float* lw) {
#ifdef PINN_EXP_NO_JET      // what-if timing experiment only
  for (int c = 0; c < SIG::C; ++c) a[c] = z[c] * 0.5f;
  return;
#endif
  const float w0 = fmaf(-t0, t0, 1.0f);
  a[0] = t0;
#ifdef PINN_LAP_CHEAP_FWD       // A/B builds: the plain rule everywhere
  constexpr bool kPrecise = false;
#else
  constexpr bool kPrecise = PRECISE;
#endif
  if constexpr (SIG::kLap && kPrecise) {
    // LAYER 1's version of the rule below (forward sweep).  In layer 1 the derivative channels are the weight columns,
    // z_j = W1[n][j], and z_s = 0: q = sum_j w_j W1[n][j]^2 is a per-neuron CONSTANT, so its half-ulp rounding is the same
    // relative perturbation of a_s = -2 t t' q at every point -- coherent over the points like a weight perturbation.  Near a
    // minimum the gradient error is (2/N) sum_i delta r_i dr_i/dtheta to leading order (the error of the residual VALUES, which
    // seed the reverse sweep; derivative errors enter multiplied by the small r_i), and a coherent delta r is not averaged out:
    // on the nearly converged Laplace problem the plain rule put the FP32 kernel at 1.1e-5 of the fp64 gradient where
    // per-direction second-order jets stay at 1.4e-6 (host build of this header, tests/host_harness; seeding the reverse sweep
    // with the fp64 residuals gives 1.2e-7 for both).  Carrying q as an unevaluated fp32 pair (error-free squares, two-sum)
    // into the combination brings the Laplacian form to 2.1e-6; applying it to the hidden layers as well adds nothing (1.7e-6)
    // and cost 3-5 % of the kernel, so they and the reverse sweep's recomputation of a^{l-1} keep the plain rule.
    float qh = 0.f, ql = 0.f;
#pragma unroll
    for (int j = 0; j < SIG::D; ++j) {
      const float zj = z[1 + j];
      a[1 + j] = w0 * zj;
      if (lw[j] == 0.f) continue;                       // a direction the Laplacian channel does not use (time, in the heat kinds)
      const float tj = PINN_MUL(lw[j], zj);
      const float p = PINN_MUL(tj, zj);
      const float ep = fmaf(fmaf(lw[j], zj, -tj), zj, fmaf(tj, zj, -p));    // exact error of p (both roundings), to first order
      if (j == 0) {
        qh = p;
        ql = ep;
      } else {
        const float s = PINN_ADD(qh, p);                // two-sum: s + e == qh + p exactly
        const float bb = PINN_SUB(s, qh);
        ql += PINN_ADD(PINN_ADD(PINN_SUB(qh, PINN_SUB(s, bb)), PINN_SUB(p, bb)), ep);
        qh = s;
      }
    }
    // z_s - 2 t qh in ONE rounding (an fma's product is exact), relative to the cancelled result; then the tail of q
    const float m2 = -2.0f * t0;
    a[SIG::lap] = w0 * fmaf(m2, ql, fmaf(m2, qh, z[SIG::lap]));
    return;
  } else if constexpr (SIG::kLap) {
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < SIG::D; ++j) {
      q = fmaf(lw[j] * z[1 + j], z[1 + j], q);
      a[1 + j] = w0 * z[1 + j];
    }
    a[SIG::lap] = w0 * fmaf(-2.0f * t0, q, z[SIG::lap]);
    return;
  } else {
    static_for<SIG::D>([&](auto J) {
    constexpr int j = decltype(J)::value;
    constexpr int K = SIG::order(j);
    constexpr int b = SIG::base(j);
    if constexpr (K >= 1) {
      float zz[K + 1], t[K + 1], w[K + 1];
      zz[0] = 0.f;
#pragma unroll
      for (int k = 1; k <= K; ++k) zz[k] = z[b + k - 1];
      tanh_series<K>(zz, t0, w0, t, w);
#pragma unroll
      for (int k = 1; k <= K; ++k) a[b + k - 1] = t[k];
    }
  });
  }
}

// Adjoint of tanh_jet_fwd: given z and abar = dLoss/da, returns zbar = dLoss/dz.
template <class SIG>
PINN_HD void tanh_jet_bwd(const float (&z)[SIG::C], const float (&abar)[SIG::C], float (&zbar)[SIG::C], const float* lw = nullptr) {
#ifdef PINN_EXP_NO_JET
  for (int c = 0; c < SIG::C; ++c) zbar[c] = abar[c] * 0.5f + z[c] * 1e-9f;
#else
  tanh_jet_bwd_t0<SIG>(z, SIG::kAccTanh ? pinn_tanh(z[0]) : tanhf(z[0]), abar, zbar, lw);
#endif
}

// adjoint of the forward-Laplacian tanh rule
template <class SIG>
PINN_HD void tanh_lap_bwd(const float (&z)[SIG::C], float t0, const float (&abar)[SIG::C], float (&zbar)[SIG::C], const float* lw) {
  const float w0 = fmaf(-t0, t0, 1.0f);
  const float as = abar[SIG::lap];
  float q = 0.f, w0bar = 0.f;
  const float k = -4.0f * t0 * w0 * as;          // d a_s / d z_j = t'' * 2 w_j z_j = -2 t t' * 2 w_j z_j
#pragma unroll
  for (int j = 0; j < SIG::D; ++j) {
    const float zj = z[1 + j];
    q = fmaf(lw[j] * zj, zj, q);
    w0bar = fmaf(abar[1 + j], zj, w0bar);
    zbar[1 + j] = fmaf(w0, abar[1 + j], k * lw[j] * zj);
  }
  zbar[SIG::lap] = w0 * as;
  w0bar = fmaf(as, fmaf(-2.0f * t0, q, z[SIG::lap]), w0bar);      // a_s = w0 (z_s - 2 t q)
  float t0bar = abar[0] - 2.0f * w0 * q * as;                       // d a_s / d t (w0 fixed)
  t0bar = fmaf(-2.0f * t0, w0bar, t0bar);                           // w0 = 1 - t^2
  zbar[0] = w0 * t0bar;
}

template <class SIG>
PINN_HD void tanh_jet_bwd_t0(const float (&z)[SIG::C], float t0, const float (&abar)[SIG::C], float (&zbar)[SIG::C], const float* lw) {
  if constexpr (SIG::kLap) {
    tanh_lap_bwd<SIG>(z, t0, abar, zbar, lw);
    return;
  } else {
  const float w0 = fmaf(-t0, t0, 1.0f);
  float t0bar = abar[0];
  float w0bar = 0.f;
  static_for<SIG::D>([&](auto J) {
    constexpr int j = decltype(J)::value;
    constexpr int K = SIG::order(j);
    constexpr int b = SIG::base(j);
    if constexpr (K >= 1) {
      float zz[K + 1], t[K + 1], w[K + 1];
      zz[0] = 0.f;
#pragma unroll
      for (int k = 1; k <= K; ++k) zz[k] = z[b + k - 1];
      tanh_series<K>(zz, t0, w0, t, w);
      float tb[K + 1], wb[K + 1], zb[K + 1];
#pragma unroll
      for (int k = 0; k <= K; ++k) {
        tb[k] = (k >= 1) ? abar[b + k - 1] : 0.f;
        wb[k] = 0.f;
        zb[k] = 0.f;
      }
#pragma unroll
      for (int k = K; k >= 1; --k) {
        // t_k = (1/k) sum_i i z_i w_{k-i}
        const float g = tb[k] * (1.0f / (float)k);
#pragma unroll
        for (int i = 1; i <= k; ++i) {
          zb[i] = fmaf((float)i * w[k - i], g, zb[i]);
          wb[k - i] = fmaf((float)i * zz[i], g, wb[k - i]);
        }
        // w_{k-1} = - sum_i t_i t_{k-1-i}   (k-1 >= 1; w_0 is handled once, below)
        if (k - 1 >= 1) {
          const float h = -2.0f * wb[k - 1];
#pragma unroll
          for (int i = 0; i <= k - 1; ++i) tb[i] = fmaf(t[k - 1 - i], h, tb[i]);
        }
      }
      t0bar += tb[0];
      w0bar += wb[0];
#pragma unroll
      for (int k = 1; k <= K; ++k) zbar[b + k - 1] = zb[k];
    }
  });
  t0bar = fmaf(-2.0f * t0, w0bar, t0bar);   // w_0 = 1 - t_0^2
  zbar[0] = w0 * t0bar;                     // t_0 = tanh(z_0)
  }
}

// ---------------------------------------------------------------------------------------------
// per-point residuals and their adjoints
// U[o][c]: Taylor-jet of network output o.  Derivatives: u_j = U[.][chan(j,1)],
// u_jj = 2 U[.][chan(j,2)], u_jjjj = 24 U[.][chan(j,4)].
// ---------------------------------------------------------------------------------------------
struct KindParams {
  int kind;
  int aux_sel[PINN_MAX_AUXSEL];
  float c[PINN_MAX_CONSTS];
  float lapw[PINN_MAX_DIM];   // direction weights of the forward-Laplacian channel (set_kind_params; unused by JetSig signatures)
  int lap_pw;                 // 1: the weights carry a per-point factor (aux column aux_sel[LIN_SEL_C2 + j]), see lap_point_weights
  int lin_seed;               // 1: reverse seeds multiply sum_rows r_k (pinn_fwd_bwd_linear) instead of the mean-square loss
};

enum : int { LIN_CU = 0, LIN_C1 = 1, LIN_C2 = 1 + PINN_MAX_DIM, LIN_SEL_SRC = 0, LIN_SEL_CU = 1, LIN_SEL_C2 = 2 };

PINN_HD float aux_or_one(const float* aux, int sel) { return sel >= 0 ? aux[sel] : 1.0f; }

// Direction weights of the forward-Laplacian channel AT ONE POINT: the descriptor's constants, times the per-point
// multiplier of each second-order term when the terms do not share one (Wave2D_Heterogeneous: u_xx + u_yy - u_tt / c(x),
// wave.py:85-89).  A per-point weight is constant through the layers, so the tanh rule is unchanged.
PINN_HD void lap_point_weights(const KindParams& kp, const float* aux, int D, float* out) {
  for (int j = 0; j < D; ++j) out[j] = kp.lap_pw ? kp.lapw[j] * aux_or_one(aux, kp.aux_sel[LIN_SEL_C2 + j]) : kp.lapw[j];
}

template <class SIG, int M>
struct Residual {
  static constexpr int C = SIG::C;
  static constexpr int D = SIG::D;

  // true when (SIG, M) can host the kind
  PINN_HD static constexpr bool sig22() { return D >= 2 && (SIG::kLap || (SIG::order(0) >= 2 && SIG::order(1) >= 2)); }
  // u_xx + u_yy of output row Uo, and its adjoint: two second-order Taylor coefficients, or the forward-Laplacian channel
  // (whose direction weights the host sets to (1, 1, 0..) for the kinds below)
  PINN_HD static float lap_xy(const float (&Uo)[C]) {
    if constexpr (SIG::kLap) return Uo[SIG::lap];
    else return 2.0f * Uo[SIG::chan(0, 2)] + 2.0f * Uo[SIG::chan(1, 2)];
  }
  PINN_HD static void lap_xy_adj(float (&Ubo)[C], float g) {
    if constexpr (SIG::kLap) Ubo[SIG::lap] = g;
    else Ubo[SIG::chan(0, 2)] = Ubo[SIG::chan(1, 2)] = 2.0f * g;
  }

  PINN_HD static bool valid(int kind) {
    if constexpr (SIG::kLap) {   // the forward-Laplacian channel only serves residuals that use one weighted sum of second derivatives
      switch (kind) {
        case PINN_KIND_LINEAR: return M == 1;
        case PINN_KIND_NS2D: return M == 3 && D == 2;
        case PINN_KIND_NS2D_UNSTEADY: return M == 3 && D == 3;
        case PINN_KIND_BURGERS2D: return M == 2 && D == 3;
        case PINN_KIND_GRAYSCOTT: return M == 2 && D == 3;
        case PINN_KIND_HEAT_NLSRC: return M == 1 && D == 3;
        case PINN_KIND_DIV_AGRAD: return M == 2 && (D == 2 || D == 3);
        default: return false;
      }
    }
    switch (kind) {
      case PINN_KIND_LINEAR: return M == 1;
      case PINN_KIND_BURGERS1D: return M == 1 && D == 2 && SIG::order(0) >= 2 && SIG::order(1) >= 1;
      case PINN_KIND_BURGERS2D: return M == 2 && D == 3 && sig22() && SIG::order(D - 1) >= 1;
      case PINN_KIND_NS2D: return M == 3 && D == 2 && sig22();
      case PINN_KIND_NS2D_UNSTEADY: return M == 3 && D == 3 && sig22() && SIG::order(D - 1) >= 1;
      case PINN_KIND_HEAT_NLSRC: return M == 1 && D == 3 && sig22() && SIG::order(D - 1) >= 1;
      case PINN_KIND_GRAYSCOTT: return M == 2 && D == 3 && sig22() && SIG::order(D - 1) >= 1;
      case PINN_KIND_KS: return M == 1 && D == 2 && SIG::order(0) >= 4 && SIG::order(1) >= 1;
      case PINN_KIND_VALUE: return C == 1;
      case PINN_KIND_FLUX: return C == 1 + D && SIG::order(0) == 1;
      case PINN_KIND_DIV_AGRAD: return M == 2 && sig22() && (D == 2 || (D == 3 && SIG::order(D - 1) >= 1));
      case PINN_KIND_GP_LINEAR2:
      case PINN_KIND_GP_BURGERS1D: return gp4();
      default: return false;
    }
  }

  // gPINN kinds: order-3 jets along e0, e1, e0+e1, e0-e1 of a one-output net (see include/pinn_b200.h)
  PINN_HD static constexpr bool gp4() {
    if constexpr (SIG::kLap) return false;
    else return M == 1 && D == 4 && SIG::order(0) >= 3 && SIG::order(1) >= 3 && SIG::order(2) >= 3 && SIG::order(3) >= 3;
  }
  // the partial derivatives a gPINN residual in two inputs (x, y) can use, from the four directional jets:
  //   2 c(2,2) = u_xx + 2 u_xy + u_yy;  6 c(2,3) = u_xxx + 3 u_xxy + 3 u_xyy + u_yyy;  6 c(3,3) = u_xxx - 3 u_xxy + 3 u_xyy - u_yyy
  struct Gp {
    float u, ux, uy, uxx, uyy, uxy, uxxx, uyyy, uxxy, uxyy;
  };
  PINN_HD static Gp gp_derivs(const float (&Uo)[C]) {
    Gp g{};
    if constexpr (gp4()) {
      const float c01 = Uo[SIG::chan(0, 1)], c02 = Uo[SIG::chan(0, 2)], c03 = Uo[SIG::chan(0, 3)];
      const float c11 = Uo[SIG::chan(1, 1)], c12 = Uo[SIG::chan(1, 2)], c13 = Uo[SIG::chan(1, 3)];
      const float c22 = Uo[SIG::chan(2, 2)], c23 = Uo[SIG::chan(2, 3)], c33 = Uo[SIG::chan(3, 3)];
      g.u = Uo[0]; g.ux = c01; g.uy = c11;
      g.uxx = 2.0f * c02; g.uyy = 2.0f * c12; g.uxy = c22 - c02 - c12;
      g.uxxx = 6.0f * c03; g.uyyy = 6.0f * c13;
      g.uxxy = (c23 - c33) - 2.0f * c13;
      g.uxyy = (c23 + c33) - 2.0f * c03;
    }
    return g;
  }
  // transpose of gp_derivs: adjoints of the partial derivatives -> adjoints of the jet channels
  PINN_HD static void gp_derivs_adj(const Gp& b, float (&Ubo)[C]) {
    if constexpr (gp4()) {
      Ubo[0] = b.u;
      Ubo[SIG::chan(0, 1)] = b.ux;
      Ubo[SIG::chan(1, 1)] = b.uy;
      Ubo[SIG::chan(0, 2)] = 2.0f * b.uxx - b.uxy;
      Ubo[SIG::chan(1, 2)] = 2.0f * b.uyy - b.uxy;
      Ubo[SIG::chan(2, 2)] = b.uxy;
      Ubo[SIG::chan(0, 3)] = 6.0f * b.uxxx - 2.0f * b.uxyy;
      Ubo[SIG::chan(1, 3)] = 6.0f * b.uyyy - 2.0f * b.uxxy;
      Ubo[SIG::chan(2, 3)] = b.uxxy + b.uxyy;
      Ubo[SIG::chan(3, 3)] = b.uxyy - b.uxxy;
    }
  }

  // r[k], k < n_pde
  PINN_HD static void eval(const KindParams& kp, const float* aux, const float (&U)[M][C], float (&r)[PINN_MAX_PDE]) {
    r[0] = r[1] = r[2] = 0.f;
    const float* c = kp.c;
    switch (kp.kind) {
      case PINN_KIND_LINEAR: {
        if constexpr (M == 1 && SIG::kLap) {
          // U[0][lap] = sum_j c2_j d2u/dx_j^2 already; the c2 terms share one per-point multiplier (host-checked), or each
          // carries its own inside the weights (lap_pw)
          float mul = 1.0f;
#pragma unroll
          for (int j = 0; j < D; ++j)
            if (!kp.lap_pw && c[LIN_C2 + j] != 0.f && kp.aux_sel[LIN_SEL_C2 + j] >= 0) mul = aux[kp.aux_sel[LIN_SEL_C2 + j]];
          float acc = mul * U[0][SIG::lap];
#pragma unroll
          for (int j = 0; j < D; ++j) acc = fmaf(c[LIN_C1 + j], U[0][1 + j], acc);
          acc = fmaf(c[LIN_CU] * aux_or_one(aux, kp.aux_sel[LIN_SEL_CU]), U[0][0], acc);
          if (kp.aux_sel[LIN_SEL_SRC] >= 0) acc += aux[kp.aux_sel[LIN_SEL_SRC]];
          r[0] = acc;
        } else if constexpr (M == 1) {
          float acc = 0.f;
          static_for<D>([&](auto J) {
            constexpr int j = decltype(J)::value;
            if constexpr (SIG::order(j) >= 2) {
              const float c2 = c[LIN_C2 + j];
              if (c2 != 0.f) acc = fmaf(2.0f * c2 * aux_or_one(aux, kp.aux_sel[LIN_SEL_C2 + j]), U[0][SIG::chan(j, 2)], acc);
            }
          });
          static_for<D>([&](auto J) {
            constexpr int j = decltype(J)::value;
            if constexpr (SIG::order(j) >= 1) acc = fmaf(c[LIN_C1 + j], U[0][SIG::chan(j, 1)], acc);
          });
          acc = fmaf(c[LIN_CU] * aux_or_one(aux, kp.aux_sel[LIN_SEL_CU]), U[0][0], acc);
          if (kp.aux_sel[LIN_SEL_SRC] >= 0) acc += aux[kp.aux_sel[LIN_SEL_SRC]];
          r[0] = acc;
        }
      } break;
      case PINN_KIND_BURGERS1D: {
        if constexpr (M == 1 && D == 2) {
          if constexpr (SIG::order(0) >= 2 && SIG::order(1) >= 1) {
            const float u = U[0][0], ux = U[0][SIG::chan(0, 1)], uxx = 2.0f * U[0][SIG::chan(0, 2)], ut = U[0][SIG::chan(1, 1)];
            r[0] = ut + u * ux - c[0] * uxx;
          }
        }
      } break;
      case PINN_KIND_BURGERS2D: {
        if constexpr (M == 2 && D == 3) {
          if constexpr (sig22() && SIG::order(2) >= 1) {
            constexpr int X1 = SIG::chan(0, 1), Y1 = SIG::chan(1, 1), T1 = SIG::chan(2, 1);
            const float u1 = U[0][0], u2 = U[1][0];
            r[0] = U[0][T1] + u1 * U[0][X1] + u2 * U[0][Y1] - c[0] * lap_xy(U[0]);
            r[1] = U[1][T1] + u1 * U[1][X1] + u2 * U[1][Y1] - c[0] * lap_xy(U[1]);
          }
        }
      } break;
      case PINN_KIND_NS2D:
      case PINN_KIND_NS2D_UNSTEADY: {
        if constexpr (M == 3 && (D == 2 || D == 3)) {
          if constexpr (sig22()) {
            constexpr int X1 = SIG::chan(0, 1), Y1 = SIG::chan(1, 1);
            const float u = U[0][0], v = U[1][0];
            const float cv = 1.0f - c[1];        // c[1] = 1: Stokes form without the convective terms (ns.py:51-67, linear=True)
            float mx = cv * (u * U[0][X1] + v * U[0][Y1]) + U[2][X1] - c[0] * lap_xy(U[0]);
            float my = cv * (u * U[1][X1] + v * U[1][Y1]) + U[2][Y1] - c[0] * lap_xy(U[1]);
            if constexpr (D == 3) {
              if constexpr (SIG::order(2) >= 1) {
                if (kp.kind == PINN_KIND_NS2D_UNSTEADY) {
                  constexpr int T1 = SIG::chan(2, 1);
                  mx += U[0][T1];
                  my += U[1][T1] + aux[kp.aux_sel[0]];
                }
              }
            }
            r[0] = mx;
            r[1] = my;
            r[2] = U[0][X1] + U[1][Y1];
          }
        }
      } break;
      case PINN_KIND_HEAT_NLSRC: {
        if constexpr (M == 1 && D == 3) {
          if constexpr (sig22() && SIG::order(2) >= 1) {
            const float u = U[0][0];
            const float lap = lap_xy(U[0]);
            r[0] = U[0][SIG::chan(2, 1)] - c[0] * lap - c[1] * sinf(c[2] * u * u) * aux[kp.aux_sel[0]];
          }
        }
      } break;
      case PINN_KIND_GRAYSCOTT: {
        if constexpr (M == 2 && D == 3) {
          if constexpr (sig22() && SIG::order(2) >= 1) {
            constexpr int T1 = SIG::chan(2, 1);
            const float u = U[0][0], v = U[1][0];
            const float lapu = lap_xy(U[0]), lapv = lap_xy(U[1]);
            r[0] = U[0][T1] - (c[0] * lapu + c[2] * (1.0f - u) - u * v * v);
            r[1] = U[1][T1] - (c[1] * lapv - c[3] * v + u * v * v);
          }
        }
      } break;
      case PINN_KIND_DIV_AGRAD: {
        if constexpr (M == 2 && (D == 2 || D == 3)) {
          if constexpr (sig22()) {
            constexpr int X1 = SIG::chan(0, 1), Y1 = SIG::chan(1, 1);
            const float a = U[1][0];
            float acc = c[1] * (a * lap_xy(U[0]) + U[1][X1] * U[0][X1] + U[1][Y1] * U[0][Y1]) + c[2] * aux[kp.aux_sel[0]];
            if constexpr (D == 3) acc = fmaf(c[0], U[0][SIG::chan(2, 1)], acc);
            r[0] = acc;
          }
        }
      } break;
      case PINN_KIND_KS: {
        if constexpr (M == 1 && D == 2) {
          if constexpr (SIG::order(0) >= 4 && SIG::order(1) >= 1) {
            const float u = U[0][0], ux = U[0][SIG::chan(0, 1)];
            r[0] = U[0][SIG::chan(1, 1)] + c[0] * u * ux + c[1] * 2.0f * U[0][SIG::chan(0, 2)] + c[2] * 24.0f * U[0][SIG::chan(0, 4)];
          }
        }
      } break;
      case PINN_KIND_GP_LINEAR2: {
        if constexpr (gp4()) {
          const Gp g = gp_derivs(U[0]);
          // per-point multipliers the reference treats as constants under autograd (table look-ups: Poisson2D_ManyArea's a(x))
          const float m2 = aux_or_one(aux, kp.aux_sel[3]), mu = aux_or_one(aux, kp.aux_sel[4]);
          const float cu = c[LIN_CU] * mu, c1x = c[LIN_C1], c1y = c[LIN_C1 + 1], c2x = c[LIN_C2] * m2, c2y = c[LIN_C2 + 1] * m2;
          r[0] = cu * g.u + c1x * g.ux + c1y * g.uy + c2x * g.uxx + c2y * g.uyy + (kp.aux_sel[0] >= 0 ? aux[kp.aux_sel[0]] : 0.f);
          r[1] = cu * g.ux + c1x * g.uxx + c1y * g.uxy + c2x * g.uxxx + c2y * g.uxyy + (kp.aux_sel[1] >= 0 ? aux[kp.aux_sel[1]] : 0.f);
          r[2] = cu * g.uy + c1x * g.uxy + c1y * g.uyy + c2x * g.uxxy + c2y * g.uyyy + (kp.aux_sel[2] >= 0 ? aux[kp.aux_sel[2]] : 0.f);
        }
      } break;
      case PINN_KIND_GP_BURGERS1D: {
        if constexpr (gp4()) {
          // inputs (x, t): uy = u_t, uxy = u_xt, uyy = u_tt, uxxy = u_xxt
          const Gp g = gp_derivs(U[0]);
          r[0] = g.uy + g.u * g.ux - c[0] * g.uxx;
          r[1] = g.uxy + g.ux * g.ux + g.u * g.uxx - c[0] * g.uxxx;
          r[2] = g.uyy + g.uy * g.ux + g.u * g.uxy - c[0] * g.uxxy;
        }
      } break;
      case PINN_KIND_VALUE: {
        const int comp = (int)c[0];
        float u = U[0][0];
#pragma unroll
        for (int o = 1; o < M; ++o) u = (comp == o) ? U[o][0] : u;
        r[0] = u - aux[kp.aux_sel[0]];
      } break;
      case PINN_KIND_FLUX: {
        if constexpr (C == 1 + D) {
          const int comp = (int)c[0];
          float acc = -aux[kp.aux_sel[0]];
#pragma unroll
          for (int o = 0; o < M; ++o) {
            if (comp == o) {
              if (kp.aux_sel[1] >= 0) acc = fmaf(aux[kp.aux_sel[1]], U[o][0], acc);
#pragma unroll
              for (int jj = 0; jj < D; ++jj) acc = fmaf(aux[kp.aux_sel[2 + jj]], U[o][1 + jj], acc);
            }
          }
          r[0] = acc;
        }
      } break;
      default: break;
    }
  }

  // Ubar = sum_k rbar[k] * d r_k / dU
  PINN_HD static void adjoint(const KindParams& kp, const float* aux, const float (&U)[M][C], const float (&rb)[PINN_MAX_PDE],
                              float (&Ub)[M][C]) {
#pragma unroll
    for (int o = 0; o < M; ++o)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) Ub[o][ch] = 0.f;
    const float* c = kp.c;
    switch (kp.kind) {
      case PINN_KIND_LINEAR: {
        if constexpr (M == 1 && SIG::kLap) {
          const float g = rb[0];
          float mul = 1.0f;
#pragma unroll
          for (int j = 0; j < D; ++j)
            if (!kp.lap_pw && c[LIN_C2 + j] != 0.f && kp.aux_sel[LIN_SEL_C2 + j] >= 0) mul = aux[kp.aux_sel[LIN_SEL_C2 + j]];
          Ub[0][SIG::lap] = mul * g;
#pragma unroll
          for (int j = 0; j < D; ++j) Ub[0][1 + j] = c[LIN_C1 + j] * g;
          Ub[0][0] = c[LIN_CU] * aux_or_one(aux, kp.aux_sel[LIN_SEL_CU]) * g;
        } else if constexpr (M == 1) {
          const float g = rb[0];
          static_for<D>([&](auto J) {
            constexpr int j = decltype(J)::value;
            if constexpr (SIG::order(j) >= 2) {
              const float c2 = c[LIN_C2 + j];
              if (c2 != 0.f) Ub[0][SIG::chan(j, 2)] = 2.0f * c2 * aux_or_one(aux, kp.aux_sel[LIN_SEL_C2 + j]) * g;
            }
            if constexpr (SIG::order(j) >= 1) Ub[0][SIG::chan(j, 1)] = c[LIN_C1 + j] * g;
          });
          Ub[0][0] = c[LIN_CU] * aux_or_one(aux, kp.aux_sel[LIN_SEL_CU]) * g;
        }
      } break;
      case PINN_KIND_BURGERS1D: {
        if constexpr (M == 1 && D == 2) {
          if constexpr (SIG::order(0) >= 2 && SIG::order(1) >= 1) {
            const float g = rb[0];
            Ub[0][0] = g * U[0][SIG::chan(0, 1)];
            Ub[0][SIG::chan(0, 1)] = g * U[0][0];
            Ub[0][SIG::chan(0, 2)] = -2.0f * c[0] * g;
            Ub[0][SIG::chan(1, 1)] = g;
          }
        }
      } break;
      case PINN_KIND_BURGERS2D: {
        if constexpr (M == 2 && D == 3) {
          if constexpr (sig22() && SIG::order(2) >= 1) {
            constexpr int X1 = SIG::chan(0, 1), Y1 = SIG::chan(1, 1), T1 = SIG::chan(2, 1);
            const float u1 = U[0][0], u2 = U[1][0];
            Ub[0][0] = rb[0] * U[0][X1] + rb[1] * U[1][X1];
            Ub[1][0] = rb[0] * U[0][Y1] + rb[1] * U[1][Y1];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              Ub[o][X1] = rb[o] * u1;
              Ub[o][Y1] = rb[o] * u2;
              lap_xy_adj(Ub[o], -c[0] * rb[o]);
              Ub[o][T1] = rb[o];
            }
          }
        }
      } break;
      case PINN_KIND_NS2D:
      case PINN_KIND_NS2D_UNSTEADY: {
        if constexpr (M == 3 && (D == 2 || D == 3)) {
          if constexpr (sig22()) {
            constexpr int X1 = SIG::chan(0, 1), Y1 = SIG::chan(1, 1);
            const float u = U[0][0], v = U[1][0];
            const float cv = 1.0f - c[1];
            const float gx = rb[0], gy = rb[1], gc = rb[2];
            const float cx = cv * gx, cy = cv * gy;
            Ub[0][0] = cx * U[0][X1] + cy * U[1][X1];
            Ub[1][0] = cx * U[0][Y1] + cy * U[1][Y1];
            Ub[0][X1] = cx * u + gc;
            Ub[0][Y1] = cx * v;
            Ub[1][X1] = cy * u;
            Ub[1][Y1] = cy * v + gc;
            Ub[2][X1] = gx;
            Ub[2][Y1] = gy;
            lap_xy_adj(Ub[0], -c[0] * gx);
            lap_xy_adj(Ub[1], -c[0] * gy);
            if constexpr (D == 3) {
              if constexpr (SIG::order(2) >= 1) {
                if (kp.kind == PINN_KIND_NS2D_UNSTEADY) {
                  constexpr int T1 = SIG::chan(2, 1);
                  Ub[0][T1] = gx;
                  Ub[1][T1] = gy;
                }
              }
            }
          }
        }
      } break;
      case PINN_KIND_HEAT_NLSRC: {
        if constexpr (M == 1 && D == 3) {
          if constexpr (sig22() && SIG::order(2) >= 1) {
            const float u = U[0][0], g = rb[0];
            Ub[0][0] = -c[1] * cosf(c[2] * u * u) * 2.0f * c[2] * u * aux[kp.aux_sel[0]] * g;
            lap_xy_adj(Ub[0], -c[0] * g);
            Ub[0][SIG::chan(2, 1)] = g;
          }
        }
      } break;
      case PINN_KIND_GRAYSCOTT: {
        if constexpr (M == 2 && D == 3) {
          if constexpr (sig22() && SIG::order(2) >= 1) {
            constexpr int T1 = SIG::chan(2, 1);
            const float u = U[0][0], v = U[1][0];
            Ub[0][0] = rb[0] * (c[2] + v * v) - rb[1] * v * v;
            Ub[1][0] = rb[0] * 2.0f * u * v + rb[1] * (c[3] - 2.0f * u * v);
            Ub[0][T1] = rb[0];
            Ub[1][T1] = rb[1];
            lap_xy_adj(Ub[0], -c[0] * rb[0]);
            lap_xy_adj(Ub[1], -c[1] * rb[1]);
          }
        }
      } break;
      case PINN_KIND_DIV_AGRAD: {
        if constexpr (M == 2 && (D == 2 || D == 3)) {
          if constexpr (sig22()) {
            constexpr int X1 = SIG::chan(0, 1), Y1 = SIG::chan(1, 1);
            const float g = c[1] * rb[0];
            lap_xy_adj(Ub[0], g * U[1][0]);
            Ub[0][X1] = g * U[1][X1];
            Ub[0][Y1] = g * U[1][Y1];
            Ub[1][0] = g * lap_xy(U[0]);
            Ub[1][X1] = g * U[0][X1];
            Ub[1][Y1] = g * U[0][Y1];
            if constexpr (D == 3) Ub[0][SIG::chan(2, 1)] = c[0] * rb[0];
          }
        }
      } break;
      case PINN_KIND_KS: {
        if constexpr (M == 1 && D == 2) {
          if constexpr (SIG::order(0) >= 4 && SIG::order(1) >= 1) {
            const float g = rb[0];
            Ub[0][0] = c[0] * U[0][SIG::chan(0, 1)] * g;
            Ub[0][SIG::chan(0, 1)] = c[0] * U[0][0] * g;
            Ub[0][SIG::chan(0, 2)] = 2.0f * c[1] * g;
            Ub[0][SIG::chan(0, 4)] = 24.0f * c[2] * g;
            Ub[0][SIG::chan(1, 1)] = g;
          }
        }
      } break;
      case PINN_KIND_GP_LINEAR2: {
        if constexpr (gp4()) {
          const float m2 = aux_or_one(aux, kp.aux_sel[3]), mu = aux_or_one(aux, kp.aux_sel[4]);
          const float cu = c[LIN_CU] * mu, c1x = c[LIN_C1], c1y = c[LIN_C1 + 1], c2x = c[LIN_C2] * m2, c2y = c[LIN_C2 + 1] * m2;
          const float g0 = rb[0], g1 = rb[1], g2 = rb[2];
          Gp b{};
          b.u = cu * g0;
          b.ux = c1x * g0 + cu * g1;
          b.uy = c1y * g0 + cu * g2;
          b.uxx = c2x * g0 + c1x * g1;
          b.uyy = c2y * g0 + c1y * g2;
          b.uxy = c1y * g1 + c1x * g2;
          b.uxxx = c2x * g1;
          b.uyyy = c2y * g2;
          b.uxxy = c2x * g2;
          b.uxyy = c2y * g1;
          gp_derivs_adj(b, Ub[0]);
        }
      } break;
      case PINN_KIND_GP_BURGERS1D: {
        if constexpr (gp4()) {
          const Gp g = gp_derivs(U[0]);
          const float g0 = rb[0], g1 = rb[1], g2 = rb[2], nu = c[0];
          Gp b{};
          b.u = g0 * g.ux + g1 * g.uxx + g2 * g.uxy;
          b.ux = g0 * g.u + 2.0f * g1 * g.ux + g2 * g.uy;
          b.uy = g0 + g2 * g.ux;
          b.uxx = -nu * g0 + g1 * g.u;
          b.uyy = g2;
          b.uxy = g1 + g2 * g.u;
          b.uxxx = -nu * g1;
          b.uxxy = -nu * g2;
          gp_derivs_adj(b, Ub[0]);
        }
      } break;
      case PINN_KIND_VALUE: {
        const int comp = (int)c[0];
#pragma unroll
        for (int o = 0; o < M; ++o) Ub[o][0] = (comp == o) ? rb[0] : 0.f;
      } break;
      case PINN_KIND_FLUX: {
        if constexpr (C == 1 + D) {
          const int comp = (int)c[0];
#pragma unroll
          for (int o = 0; o < M; ++o) {
            if (comp == o) {
              Ub[o][0] = (kp.aux_sel[1] >= 0) ? aux[kp.aux_sel[1]] * rb[0] : 0.f;
#pragma unroll
              for (int jj = 0; jj < D; ++jj) Ub[o][1 + jj] = aux[kp.aux_sel[2 + jj]] * rb[0];
            }
          }
        }
      } break;
      default: break;
    }
  }
};

// Host-side: can this descriptor run on the forward-Laplacian signature LapSig<d>?
//  * LINEAR kind, one output: no direction beyond second order, at least two second-order directions (else nothing is
//    saved); weights = c2_j, times each term's own per-point multiplier when the terms do not share one;
//  * the kinds whose only second derivatives are u_xx + u_yy (NS2D, NS2D_UNSTEADY, BURGERS2D, GRAYSCOTT, HEAT_NLSRC)
//    and DIV_AGRAD (outputs u, a) with their canonical orders (2,2) / (2,2,1); weights = (1, 1, 0).
// returns 0 (no), 1 (yes, constant direction weights) or 2 (yes, per-point direction weights: lap_point_weights)
inline int lap_eligible(int kind, int d, int m, const int* order, const int* aux_sel, const float* consts) {
  if (kind == PINN_KIND_NS2D) return d == 2 && m == 3 && order[0] == 2 && order[1] == 2;
  if (kind == PINN_KIND_NS2D_UNSTEADY || kind == PINN_KIND_BURGERS2D || kind == PINN_KIND_GRAYSCOTT || kind == PINN_KIND_HEAT_NLSRC)
    return d == 3 && order[0] == 2 && order[1] == 2 && order[2] == 1;
  if (kind == PINN_KIND_DIV_AGRAD)
    return m == 2 && order[0] == 2 && order[1] == 2 && (d == 2 || (d == 3 && order[2] == 1));
  if (kind != PINN_KIND_LINEAR || m != 1 || d < 2) return 0;
  int n2 = 0, sel = -2;
  bool common = true;
  for (int j = 0; j < d; ++j) {
    if (order[j] < 1 || order[j] > 2) return 0;
    if (order[j] == 2) ++n2;
    if (order[j] < 2 && consts[LIN_C2 + j] != 0.f) return 0;
    if (consts[LIN_C2 + j] != 0.f) {
      const int sj = aux_sel[LIN_SEL_C2 + j];
      if (sel == -2) sel = sj;
      else if (sel != sj) common = false;
    }
  }
  if (n2 < 2) return 0;
  return common ? 1 : 2;
}

// descriptor -> KindParams (the kernels' copy of kind, selectors, constants and Laplacian weights)
inline void set_kind_params(KindParams& kp, int kind, int d, const int* aux_sel, const float* consts, int lap_mode = 0) {
  kp.kind = kind;
  kp.lap_pw = (lap_mode == 2) ? 1 : 0;
  kp.lin_seed = 0;
  for (int i = 0; i < PINN_MAX_AUXSEL; ++i) kp.aux_sel[i] = aux_sel[i];
  for (int i = 0; i < PINN_MAX_CONSTS; ++i) kp.c[i] = consts[i];
  for (int j = 0; j < PINN_MAX_DIM; ++j) kp.lapw[j] = (kind == PINN_KIND_LINEAR) ? (j < d ? consts[LIN_C2 + j] : 0.f) : (j < 2 ? 1.0f : 0.f);
}

}  // namespace pinn
