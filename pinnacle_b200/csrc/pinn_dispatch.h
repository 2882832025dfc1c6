// pinn_dispatch.h -- type-erased handle to one instantiated kernel configuration.
#pragma once
#include <cuda_runtime.h>

#include "pinn_kernels.cuh"
#include "pinn_tc_kernels.cuh"
#include "pinn_tc2_kernels.cuh"

namespace pinn {

struct ConfigInfo {
  int id, H, m, d, C;
  int lap;   // 1: forward-Laplacian signature LapSig<d> (serves eligible LINEAR-kind descriptors whatever their order[]); 2: with per-point direction weights
  int order[PINN_MAX_DIM];
  int T, NT, NTP, E4;
  size_t smem_fwd, smem_bwd;
  bool (*valid_kind)(int);
  cudaError_t (*prepare)();
  cudaError_t (*launch)(const LaunchParams&, bool bwd, int grid, cudaStream_t);
  // tensor-core (tcgen05) variant; tc_T == 0 when this configuration has none
  int tc_T, tc_NQ, tc_rmw;
  int tc_pipe;   // 1: the pipelined kernel (pinn_tc2_kernels.cuh: 32-point tiles, 2 stash quads per thread, RMW dW partials) exists too
  size_t tc_smem_fwd, tc_smem_bwd;
  cudaError_t (*tc_prepare)();
  cudaError_t (*tc_launch)(const TcLaunchParams&, bool bwd, bool pipelined, int grid, cudaStream_t);
};

template <int ID>
const ConfigInfo* get_config();

#define PINN_DECLARE_CONFIG(ID, ...) template <> const ConfigInfo* get_config<ID>();

}  // namespace pinn
