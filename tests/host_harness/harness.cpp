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
static int g_tc_mode = -1;   // 0 off, 1 truncating accumulate, 2 round-to-nearest accumulate
static inline int tc_mode() { if (g_tc_mode < 0) { const char* e = getenv("PINN_HARNESS_TC"); g_tc_mode = e ? atoi(e) : 0; } return g_tc_mode; }
extern "C" void harness_set_tc_mode(int m) { g_tc_mode = m; }   // 0 plain fp32 FMA, 1 tensor-core model (truncating accumulate), 2 same with RN accumulate, 3 truncating, main term over two accumulators
// dot(w[0..n), a[0..n)) as the kernels' rows contraction computes it
static float tc_dot(const float* w, const float* a, int n, int astride) {
  static const int wa[6] = {1, 1, 2, 0, 0, 0}, ab[6] = {1, 0, 0, 2, 1, 0};
  float acc = 0.f, acc2 = 0.f;     // acc2: mode 3 only -- odd k-steps of the main term go to a second (truncating) accumulator
  for (int t = 0; t < 6; ++t)
    for (int k0 = 0; k0 < n; k0 += 16) {
      double s = 0.0;
      for (int k = k0; k < k0 + 16 && k < n; ++k) { float ws[3], as[3]; split3(w[k], ws); split3(a[(size_t)k * astride], as); s += (double)ws[wa[t]] * (double)as[ab[t]]; }
      if (tc_mode() == 3 && t == 5 && ((k0 / 16) & 1)) { acc2 = trunc_f32((double)acc2 + s); continue; }
      const double v = (double)acc + s;
      acc = (tc_mode() == 1 || tc_mode() == 3) ? trunc_f32(v) : (float)v;
    }
  return acc + acc2;
}

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
      tanh_jet_fwd<SIG>(zz, aa, lwp);
      for (int c = 0; c < C; ++c) { z[((size_t)1 * H + nn) * C + c] = zz[c]; a[((size_t)1 * H + nn) * C + c] = aa[c]; }
    }
    for (int l = 2; l <= L; ++l) {
      for (int nn = 0; nn < H; ++nn) {
        float zz[C];
        for (int c = 0; c < C; ++c) zz[c] = 0.f;
        if (tc_mode()) {
          for (int c = 0; c < C; ++c) zz[c] = tc_dot(&params[offW[l] + (long long)nn * H], &a[((size_t)(l - 1) * H) * C + c], H, C);
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
      for (int k = 0; k < d->n_pde; ++k) rb[k] = seeds[s * d->n_pde + k] * 2.0f * r[k] * (float)inv_n;
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
              if (tc_mode()) {
                float wcol[128];
                for (int nn = 0; nn < H; ++nn) wcol[nn] = params[offW[l] + (long long)nn * H + k];
                v = tc_dot(wcol, &zb[c], H, C);
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
