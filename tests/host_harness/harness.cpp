// TEST INFRASTRUCTURE -- host build of the device math (pinnacle_b200/csrc/jet_math.cuh).
//
// Runs the same tanh-jet recurrences, hand-derived adjoints and residual functors the CUDA kernels
// use, but with plain per-point loops for the layer contractions, so the MATH can be checked against
// the autograd oracle in the build container (no GPU).  Never linked into the product.
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "../../pinnacle_b200/csrc/jet_math.cuh"
#include "../../pinnacle_b200/csrc/pinn_configs.h"

using namespace pinn;


// ---- emulation of the tensor-core contraction (bf16x3, 6 products, k-steps of 16, fp32 accumulate) for accuracy studies ----
#include <cstring>
#include <cstdlib>
static inline float bf16_rn(float x) { uint32_t b; std::memcpy(&b, &x, 4); uint32_t r = b + 0x7FFFu + ((b >> 16) & 1u); r &= 0xFFFF0000u; float y; std::memcpy(&y, &r, 4); return y; }
static inline void split3(float x, float (&s)[3]) { s[0] = bf16_rn(x); float r = x - s[0]; s[1] = bf16_rn(r); r -= s[1]; s[2] = bf16_rn(r); }
static inline float trunc_f32(double v) { float f = (float)v; if (std::fabs((double)f) > std::fabs(v)) { uint32_t b; std::memcpy(&b, &f, 4); b -= 1; std::memcpy(&f, &b, 4); } return f; }
// Contraction models (harness_set_tc_mode / harness_set_tc_modes), ranked on the CPU before GPU time is spent:
//   0  plain fp32 FMA chain
//   1  the kernels' scheme: floating bf16x3 splits, six products in issue order (small terms first, main term last), one
//      accumulator, fp32 accumulate that truncates once per k-step (the round-1 model)
//   2  same with a round-to-nearest accumulate (what an ideal fp32 accumulate would give)
//   3  as 1, main term over two accumulators (even / odd k-steps)
//   4  as 1 with the accumulate the hardware probe found (scripts/umma_round_probe.py): the 16 products and the accumulator
//      are aligned to the largest exponent among them, each addend's magnitude truncated to a 26-bit window, summed exactly,
//      the sum truncated (toward zero) to 24 bits
//   10 fixed-grid digits (three signed 8-bit digits relative to the row / column maximum: every product of a group sits on
//      one grid, so a group's sum is exact in fp32), one accumulator, groups in the order 2^-16, 2^-8, main; hardware accumulate
//   11 fixed-grid digits, main group alone in its own accumulator (exact), small groups in a second one, added in
//      round-to-nearest fp32 by the epilogue
//   12 floating bf16x3 (as 4) with the main term in its own accumulator
//   13 as 10 with the main group first
static int g_tc_fwd = -1, g_tc_bwd = -1;
static inline void tc_init() { if (g_tc_fwd < 0) { const char* e = getenv("PINN_HARNESS_TC"); g_tc_fwd = g_tc_bwd = e ? atoi(e) : 0; } }
extern "C" void harness_set_tc_mode(int m) { g_tc_fwd = g_tc_bwd = m; }
extern "C" void harness_set_tc_modes(int fwd, int bwd) { g_tc_fwd = fwd; g_tc_bwd = bwd; }
static inline int tc_mode() { tc_init(); return g_tc_fwd; }

// one tcgen05 kind::f16 accumulate step as probed on B200: D <- RZ24( trunc26(D) + sum_i trunc26(p_i) )
static float hw_acc(float acc, const double* p, int n) {
  double mx = std::fabs((double)acc);
  for (int i = 0; i < n; ++i) mx = std::fmax(mx, std::fabs(p[i]));
  if (mx == 0.0) return 0.f;
  int e; std::frexp(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)  ->  leading bit 2^(e-1)
  const double grid = std::ldexp(1.0, e - 1 - 25);
  double sum = std::trunc((double)acc / grid);
  for (int i = 0; i < n; ++i) sum += std::trunc(p[i] / grid);
  return trunc_f32(sum * grid);
}
// three signed 8-bit digits of x / scale:  x ~ scale * (d[0] + d[1] 2^-8 + d[2] 2^-16), |d[i]| <= 128 (round to nearest)
static inline void digits3(float x, float inv_scale, float (&d)[3]) {
  float v = x * inv_scale;
  d[0] = std::nearbyint(v); v = (v - d[0]) * 256.f;
  d[1] = std::nearbyint(v); v = (v - d[1]) * 256.f;
  d[2] = std::nearbyint(v);
}
static inline float pow2_scale(float mx) {           // power of two s with mx / s <= 128
  if (mx == 0.f) return 1.f;
  int e; std::frexp(mx, &e);                         // mx < 2^e
  return std::ldexp(1.f, e - 7);
}
// dot(w[0..n), a[0..n)) as the rows contraction computes it under `mode`
static float tc_dot(const float* w, const float* a, int n, int astride, int mode) {
  static const int wa[6] = {1, 1, 2, 0, 0, 0}, ab[6] = {1, 0, 0, 2, 1, 0};
  if (mode >= 30 && mode <= 39) {
    // "exact main term": w1 = 8 significant bits of w, but never finer than 2^-14 of the row's leading bit (so w1 + bf16(w - w1)
    // + bf16(rest) = w exactly for every weight within 7 binades of the largest); a1 = a rounded to a grid of A1 bits below the
    // leading bit of the column maximum over each block of 32 consecutive k (one warp's lanes); a2, a3 = floating bf16 splits of the
    // remainder.  w1*a1 accumulates alone (every product on one coarse grid: the sums are exact), the other five products in a
    // second accumulator, the two are added in round-to-nearest fp32.
    const int A1 = (mode == 31) ? 8 : (mode == 32) ? 6 : 7;
    float mw = 0.f;
    for (int k = 0; k < n; ++k) mw = std::fmax(mw, std::fabs(w[k]));
    int ew; std::frexp(mw, &ew);                                // mw < 2^ew
    const float gw_floor = std::ldexp(1.f, ew - 1 - 14);
    float wp[128][3], ap[128][3];
    for (int k0 = 0; k0 < n; k0 += 32) {
      float ma = 0.f;
      for (int k = k0; k < k0 + 32 && k < n; ++k) ma = std::fmax(ma, std::fabs(a[(size_t)k * astride]));
      int ea; std::frexp(ma, &ea);
      const float ga = (ma == 0.f) ? 1.f : std::ldexp(1.f, ea - A1);
      for (int k = k0; k < k0 + 32 && k < n; ++k) {
        const float av = a[(size_t)k * astride];
        ap[k][0] = std::nearbyint(av / ga) * ga;
        float r = av - ap[k][0];
        ap[k][1] = bf16_rn(r); r -= ap[k][1];
        ap[k][2] = bf16_rn(r);
        const float wv = w[k];
        int e1; std::frexp(wv, &e1);
        const float g1 = std::fmax(std::ldexp(1.f, e1 - 8), gw_floor);
        wp[k][0] = (wv == 0.f) ? 0.f : std::nearbyint(wv / g1) * g1;
        r = wv - wp[k][0];
        wp[k][1] = bf16_rn(r); r -= wp[k][1];
        wp[k][2] = bf16_rn(r);
      }
    }
    float dm = 0.f, dsm = 0.f;
    static const int swa[5] = {1, 1, 2, 0, 0}, sab[5] = {1, 0, 0, 2, 1};
    for (int k0 = 0; k0 < n; k0 += 16) {
      double p[16]; int m = 0;
      for (int k = k0; k < k0 + 16 && k < n; ++k) p[m++] = (double)wp[k][0] * (double)ap[k][0];
      dm = hw_acc(dm, p, m);
    }
    for (int t = 0; t < 5; ++t)
      for (int k0 = 0; k0 < n; k0 += 16) {
        double p[16]; int m = 0;
        for (int k = k0; k < k0 + 16 && k < n; ++k) p[m++] = (double)wp[k][swa[t]] * (double)ap[k][sab[t]];
        dsm = hw_acc(dsm, p, m);
      }
    if (mode == 33) return hw_acc(dsm, nullptr, 0) + dm;      // (unused)
    return dm + dsm;
  }
  if (mode == 20 || mode == 21 || mode == 22) {        // diagnostics: fp32 FMA chain on quantised operands (20: W, 21: activations); 22: W * (1 - 2^-24)
    float mw = 0.f, ma = 0.f;
    for (int k = 0; k < n; ++k) { mw = std::fmax(mw, std::fabs(w[k])); ma = std::fmax(ma, std::fabs(a[(size_t)k * astride])); }
    const float sw = pow2_scale(mw), sa = pow2_scale(ma);
    float acc = 0.f;
    for (int k = 0; k < n; ++k) {
      float wk = w[k], ak = a[(size_t)k * astride];
      if (mode == 20) { float d[3]; digits3(wk, 1.f / sw, d); wk = (float)(((double)d[0] + d[1] / 256.0 + d[2] / 65536.0) * sw); }
      if (mode == 21) { float d[3]; digits3(ak, 1.f / sa, d); ak = (float)(((double)d[0] + d[1] / 256.0 + d[2] / 65536.0) * sa); }
      acc = fmaf(wk, ak, acc);
    }
    if (mode == 22) acc = (float)((double)acc * (1.0 - 1.0 / 16777216.0));
    return acc;
  }
  if (mode >= 10 && mode != 12) {
    float mw = 0.f, ma = 0.f;
    for (int k = 0; k < n; ++k) { mw = std::fmax(mw, std::fabs(w[k])); ma = std::fmax(ma, std::fabs(a[(size_t)k * astride])); }
    const float sw = pow2_scale(mw), sa = pow2_scale(ma);
    float dw[128][3], da[128][3];
    for (int k = 0; k < n; ++k) { digits3(w[k], 1.f / sw, dw[k]); digits3(a[(size_t)k * astride], 1.f / sa, da[k]); }
    const double gs[3] = {1.0, 1.0 / 256.0, 1.0 / 65536.0};
    // group g = i + j (digit indices): 0 main, 1 -> 2^-8, 2 -> 2^-16
    float acc = 0.f, acc_main = 0.f;
    const int order10[3] = {2, 1, 0}, order13[3] = {0, 1, 2};
    const int* order = (mode == 13) ? order13 : order10;
    for (int gi = 0; gi < 3; ++gi) {
      const int g = order[gi];
      for (int i = 0; i <= g; ++i) {                  // w digit i, a digit g - i: one MMA batch (7 k-steps) per product
        const int j = g - i;
        for (int k0 = 0; k0 < n; k0 += 16) {
          double p[16]; int m = 0;
          for (int k = k0; k < k0 + 16 && k < n; ++k) p[m++] = (double)dw[k][i] * (double)da[k][j] * gs[g];
          if (mode == 11 && g == 0) acc_main = hw_acc(acc_main, p, m);
          else acc = hw_acc(acc, p, m);
        }
      }
    }
    return (acc_main + acc) * (sw * sa);              // fp32 round-to-nearest add, exact power-of-two scaling
  }
  float acc = 0.f, acc2 = 0.f;     // acc2: second accumulator (mode 3: odd k-steps of the main term; mode 12: the whole main term)
  for (int t = 0; t < 6; ++t)
    for (int k0 = 0; k0 < n; k0 += 16) {
      double s = 0.0, p[16]; int m = 0;
      for (int k = k0; k < k0 + 16 && k < n; ++k) { float ws[3], as[3]; split3(w[k], ws); split3(a[(size_t)k * astride], as); p[m] = (double)ws[wa[t]] * (double)as[ab[t]]; s += p[m++]; }
      if (mode == 4 || mode == 12) {
        if (mode == 12 && t == 5) acc2 = hw_acc(acc2, p, m); else acc = hw_acc(acc, p, m);
        continue;
      }
      if (mode == 3 && t == 5 && ((k0 / 16) & 1)) { acc2 = trunc_f32((double)acc2 + s); continue; }
      const double v = (double)acc + s;
      acc = (mode == 1 || mode == 3) ? trunc_f32(v) : (float)v;
    }
  return acc + acc2;
}

// diagnostic (scripts/lap_error_study.py): seed the reverse sweep with externally supplied residual values (e.g. the fp64
// reference's) instead of the ones this forward pass computed -- separates the two error terms of the gradient near a minimum
static const double* g_resid_override = nullptr;
extern "C" void harness_set_resid_override(const double* r) { g_resid_override = r; }

template <class SIG, int M>
static void run_cfg(const pinn_kind_desc* d, const float* params, const float* x, const float* aux, long long n, double inv_n,
                    const float* seeds, int S, double* grad, double* loss, float* resid) {
  constexpr int C = SIG::C, D = SIG::D;
  const int H = d->H, L = d->L;
  const long long P = (long long)D * H + H + (long long)(L - 1) * ((long long)H * H + H) + (long long)H * M + M;
  std::vector<long long> offW(L + 2), offB(L + 2);
  offW[1] = 0; offB[1] = (long long)H * D;
  for (int l = 2; l <= L; ++l) { offW[l] = (long long)H * D + H + (long long)(l - 2) * ((long long)H * H + H); offB[l] = offW[l] + (long long)H * H; }
  offW[L + 1] = (long long)H * D + H + (long long)(L - 1) * ((long long)H * H + H); offB[L + 1] = offW[L + 1] + (long long)M * H;
  KindParams kp;
  set_kind_params(kp, d->kind, d->d, d->aux_sel, d->consts, SIG::kLap ? (SIG::kLapPW ? 2 : 1) : 0);
  for (int k = 0; k < PINN_MAX_PDE; ++k) loss[k] = 0.0;
  for (long long i = 0; i < (long long)S * P; ++i) grad[i] = 0.0;

  std::vector<float> z((size_t)(L + 1) * H * C), a((size_t)(L + 1) * H * C), ab((size_t)H * C), zb((size_t)H * C), abp((size_t)H * C);
  for (long long pt = 0; pt < n; ++pt) {
    const float* xp = x + pt * D;
    float auxv[PINN_MAX_AUX] = {};
    for (int q = 0; q < d->n_aux; ++q) auxv[q] = aux[pt * d->n_aux + q];
    float lwp[PINN_MAX_DIM] = {};
    if constexpr (SIG::kLapPW) lap_point_weights(kp, auxv, D, lwp);
    else if constexpr (SIG::kLap) for (int j = 0; j < D; ++j) lwp[j] = kp.lapw[j];
    // ---- forward
    for (int nn = 0; nn < H; ++nn) {
      float zz[C];
      for (int c = 0; c < C; ++c) zz[c] = 0.f;
      float z0 = params[offB[1] + nn];
      for (int j = 0; j < D; ++j) z0 = fmaf(params[offW[1] + nn * D + j], xp[j], z0);
      zz[0] = z0;
      static_for<D>([&](auto J) {
        constexpr int j = decltype(J)::value;
        if constexpr (SIG::order(j) >= 1) zz[SIG::base(j)] = params[offW[1] + nn * D + j];
      });
      float aa[C];
      tanh_jet_fwd<SIG, true>(zz, aa, lwp);      // layer 1: the pair rule (jet_math.cuh)
      for (int c = 0; c < C; ++c) { z[((size_t)1 * H + nn) * C + c] = zz[c]; a[((size_t)1 * H + nn) * C + c] = aa[c]; }
    }
    for (int l = 2; l <= L; ++l) {
      for (int nn = 0; nn < H; ++nn) {
        float zz[C];
        for (int c = 0; c < C; ++c) zz[c] = 0.f;
        if (tc_mode()) {
          for (int c = 0; c < C; ++c) zz[c] = tc_dot(&params[offW[l] + (long long)nn * H], &a[((size_t)(l - 1) * H) * C + c], H, C, g_tc_fwd);
        } else
        for (int k = 0; k < H; ++k) {
          const float w = params[offW[l] + (long long)nn * H + k];
          for (int c = 0; c < C; ++c) zz[c] = fmaf(a[((size_t)(l - 1) * H + k) * C + c], w, zz[c]);
        }
        zz[0] += params[offB[l] + nn];
        float aa[C];
        tanh_jet_fwd<SIG>(zz, aa, lwp);
        for (int c = 0; c < C; ++c) { z[((size_t)l * H + nn) * C + c] = zz[c]; a[((size_t)l * H + nn) * C + c] = aa[c]; }
      }
    }
    float U[M][C];
    for (int o = 0; o < M; ++o)
      for (int c = 0; c < C; ++c) {
        float s = (c == 0) ? params[offB[L + 1] + o] : 0.f;
        for (int k = 0; k < H; ++k) s = fmaf(params[offW[L + 1] + (long long)o * H + k], a[((size_t)L * H + k) * C + c], s);
        U[o][c] = s;
      }
    float r[PINN_MAX_PDE];
    Residual<SIG, M>::eval(kp, auxv, U, r);
    for (int k = 0; k < d->n_pde; ++k) {
      loss[k] += (double)r[k] * (double)r[k];
      if (resid) resid[pt * d->n_pde + k] = r[k];
    }
    // ---- reverse, per seed
    for (int s = 0; s < S; ++s) {
      double* g = grad + (size_t)s * P;
      float rb[PINN_MAX_PDE] = {0, 0, 0};
      for (int k = 0; k < d->n_pde; ++k)
        rb[k] = seeds[s * d->n_pde + k] * 2.0f * (g_resid_override ? (float)g_resid_override[pt * d->n_pde + k] : r[k]) * (float)inv_n;
      float Ub[M][C];
      Residual<SIG, M>::adjoint(kp, auxv, U, rb, Ub);
      for (int o = 0; o < M; ++o) {
        g[offB[L + 1] + o] += Ub[o][0];
        for (int k = 0; k < H; ++k)
          for (int c = 0; c < C; ++c) g[offW[L + 1] + (long long)o * H + k] += (double)Ub[o][c] * a[((size_t)L * H + k) * C + c];
      }
      for (int k = 0; k < H; ++k)
        for (int c = 0; c < C; ++c) {
          float v = 0.f;
          for (int o = 0; o < M; ++o) v = fmaf(Ub[o][c], params[offW[L + 1] + (long long)o * H + k], v);
          ab[(size_t)k * C + c] = v;
        }
      for (int l = L; l >= 1; --l) {
        for (int nn = 0; nn < H; ++nn) {
          float zz[C], aa[C], zz_b[C];
          for (int c = 0; c < C; ++c) { zz[c] = z[((size_t)l * H + nn) * C + c]; aa[c] = ab[(size_t)nn * C + c]; zz_b[c] = 0.f; }
          tanh_jet_bwd<SIG>(zz, aa, zz_b, lwp);
          for (int c = 0; c < C; ++c) zb[(size_t)nn * C + c] = zz_b[c];
          g[offB[l] + nn] += zz_b[0];
        }
        if (l == 1) {
          for (int nn = 0; nn < H; ++nn) {
            static_for<D>([&](auto J) {
              constexpr int j = decltype(J)::value;
              double v = (double)zb[(size_t)nn * C + 0] * xp[j];
              if constexpr (SIG::order(j) >= 1) v += zb[(size_t)nn * C + SIG::base(j)];
              g[offW[1] + nn * D + j] += v;
            });
          }
        } else {
          for (int nn = 0; nn < H; ++nn)
            for (int k = 0; k < H; ++k) {
              double v = 0.0;
              for (int c = 0; c < C; ++c) v += (double)zb[(size_t)nn * C + c] * a[((size_t)(l - 1) * H + k) * C + c];
              g[offW[l] + (long long)nn * H + k] += v;
            }
          for (int k = 0; k < H; ++k)
            for (int c = 0; c < C; ++c) {
              float v = 0.f;
              if (g_tc_bwd > 0) {
                float wcol[128];
                for (int nn = 0; nn < H; ++nn) wcol[nn] = params[offW[l] + (long long)nn * H + k];
                v = tc_dot(wcol, &zb[c], H, C, g_tc_bwd);
              } else
              for (int nn = 0; nn < H; ++nn) v = fmaf(zb[(size_t)nn * C + c], params[offW[l] + (long long)nn * H + k], v);
              abp[(size_t)k * C + c] = v;
            }
          ab.swap(abp);
        }
      }
    }
  }
  for (int k = 0; k < PINN_MAX_PDE; ++k) loss[k] *= inv_n;
}

template <int... KS>
struct SigOfH {
  using type = JetSig<KS...>;
};
template <int K0>
struct SigOfH<K0> {
  using type = std::conditional_t<(K0 < -10), LapSig<(K0 < -10 ? -K0 - 10 : 1), true>,
                                  std::conditional_t<(K0 < 0), LapSig<(K0 < 0 ? -K0 : 1)>, JetSig<(K0 < 0 ? 0 : K0)>>>;
};

extern "C" int harness_run(const pinn_kind_desc* d, const float* params, const float* x, const float* aux, long long n, double inv_n,
                           const float* seeds, int S, double* grad, double* loss, float* resid) {
  // same selection rule as the library (pinn_api.cu find_config): forward-Laplacian signature where eligible
  const char* nl = getenv("PINN_NO_LAPLACIAN");
  const int want_lap = (nl != nullptr && nl[0] == '1') ? 0 : lap_eligible(d->kind, d->d, d->m, d->order, d->aux_sel, d->consts);
#define TRY(ID, H_, M_, T_, TN_, TCT_, ...)                                                                \
  {                                                                                                    \
    using SIG = typename SigOfH<__VA_ARGS__>::type;                                                    \
    const int o[] = {__VA_ARGS__};                                                                     \
    bool same = (SIG::D == d->d) && (M_ == d->m) && ((SIG::kLap ? (SIG::kLapPW ? 2 : 1) : 0) == want_lap);                           \
    for (int j = 0; same && !SIG::kLap && j < SIG::D; ++j) same = (o[j] == d->order[j]);               \
    if (same && Residual<SIG, M_>::valid(d->kind)) {                                                   \
      run_cfg<SIG, M_>(d, params, x, aux, n, inv_n, seeds, S, grad, loss, resid);                      \
      return 0;                                                                                        \
    }                                                                                                  \
  }
  PINN_FOR_EACH_CONFIG(TRY)
#undef TRY
  return 1;
}
