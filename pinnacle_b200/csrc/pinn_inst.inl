// pinn_inst.inl -- instantiates the forward and forward+reverse kernels of ONE configuration.
// Included by the generated inst_<ID>.cu files (see build.py), which specialise get_config<ID>().
#include "pinn_configs.h"
#include "pinn_dispatch.h"

#include <type_traits>

namespace pinn {
namespace {

// orders pack -> signature: (k_1, .., k_d) = per-direction Taylor orders; a single negative entry -d = the
// forward-Laplacian signature LapSig<d> (value, d first derivatives, one weighted sum of second derivatives)
template <int... KS>
struct SigOf {
  using type = JetSig<KS...>;
  static constexpr int lap = 0;
};
// -d: LapSig<d> with constant direction weights; -(10 + d): LapSig<d, true>, per-point direction weights
template <int K0>
struct SigOf<K0> {
  using type = std::conditional_t<(K0 < -10), LapSig<(K0 < -10 ? -K0 - 10 : 1), true>,
                                  std::conditional_t<(K0 < 0), LapSig<(K0 < 0 ? -K0 : 1)>, JetSig<(K0 < 0 ? 0 : K0)>>>;
  static constexpr int lap = K0 < -10 ? 2 : (K0 < 0 ? 1 : 0);
};

template <int ID, int H, int M, int T, int TN, int TCT, int... KS>
struct Inst {
  using SIG = typename SigOf<KS...>::type;
  using CF = KCfg<SIG, M, H, T, TN, false>;
  using CB = KCfg<SIG, M, H, T, TN, true>;

  static bool valid_kind(int kind) { return Residual<SIG, M>::valid(kind); }

  static cudaError_t prepare() {
    cudaError_t e = cudaFuncSetAttribute(pinn_fused_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pinn_fused_kernel<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB::smem_bytes);
  }

  static cudaError_t launch(const LaunchParams& p, bool bwd, int grid, cudaStream_t st) {
    if (bwd)
      pinn_fused_kernel<CB><<<grid, CB::NTP, CB::smem_bytes, st>>>(p);
    else
      pinn_fused_kernel<CF><<<grid, CF::NTP, CF::smem_bytes, st>>>(p);
    return cudaGetLastError();
  }

  template <int TT>
  struct Tc {
    using TF = TCfg<SIG, M, H, TT, false>;
    using TB = TCfg<SIG, M, H, TT, true>;
    // the pipelined two-sub-tile kernel (pinn_tc2_kernels.cuh) exists where two sub-tiles fit in shared memory
    static constexpr bool kPipe = (TT == 32 && SIG::C <= 5) || (TT == 16 && SIG::C == 6);
    static cudaError_t prepare() {
      cudaError_t e = cudaFuncSetAttribute(pinn_tc_kernel<TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF::smem_bytes);
      if (e != cudaSuccess) return e;
      if constexpr (kPipe) {
        using PF = TC2fg<SIG, M, H, false>;
        using PB = TC2fg<SIG, M, H, true>;
        e = cudaFuncSetAttribute(pinn_tc2_kernel<PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF::smem_bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(pinn_tc2_kernel<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB::smem_bytes);
        if (e != cudaSuccess) return e;
      }
      return cudaFuncSetAttribute(pinn_tc_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB::smem_bytes);
    }
    static cudaError_t launch(const TcLaunchParams& p, bool bwd, bool pipelined, int grid, cudaStream_t st) {
      if constexpr (kPipe) {
        if (pipelined) {
          using PF = TC2fg<SIG, M, H, false>;
          using PB = TC2fg<SIG, M, H, true>;
          if (bwd)
            pinn_tc2_kernel<PB><<<grid, PB::NTHALL, PB::smem_bytes, st>>>(p);
          else
            pinn_tc2_kernel<PF><<<grid, PF::NTHALL, PF::smem_bytes, st>>>(p);
          return cudaGetLastError();
        }
      }
      if (bwd)
        pinn_tc_kernel<TB><<<grid, TB::NTH, TB::smem_bytes, st>>>(p);
      else
        pinn_tc_kernel<TF><<<grid, TF::NTH, TF::smem_bytes, st>>>(p);
      return cudaGetLastError();
    }
    static void fill(ConfigInfo& c) {
      c.tc_T = TT; c.tc_NQ = TB::NQ; c.tc_rmw = TB::RMW_FLUSH ? 1 : 0; c.tc_pipe = kPipe ? 1 : 0; c.tc_smem_fwd = TF::smem_bytes; c.tc_smem_bwd = TB::smem_bytes;
      c.tc_prepare = &prepare; c.tc_launch = &launch;
    }
  };

  static const ConfigInfo* info() {
    static const ConfigInfo ci = [] {
      ConfigInfo c{};
      c.id = ID; c.H = H; c.m = M; c.d = SIG::D; c.C = SIG::C;
      const int o[] = {KS...};
      c.lap = SigOf<KS...>::lap;
      for (int j = 0; j < PINN_MAX_DIM; ++j) c.order[j] = j < SIG::D ? (c.lap ? 2 : o[j]) : 0;
      c.T = T; c.NT = CB::NT; c.NTP = CB::NTP; c.E4 = CB::E4;
      c.smem_fwd = CF::smem_bytes; c.smem_bwd = CB::smem_bytes;
      c.valid_kind = &valid_kind; c.prepare = &prepare; c.launch = &launch;
      c.tc_T = 0; c.tc_NQ = 0; c.tc_rmw = 0; c.tc_pipe = 0; c.tc_smem_fwd = c.tc_smem_bwd = 0; c.tc_prepare = nullptr; c.tc_launch = nullptr;
      if constexpr (TCT > 0) Tc<TCT>::fill(c);
      return c;
    }();
    return &ci;
  }
};

}  // namespace

}  // namespace pinn
