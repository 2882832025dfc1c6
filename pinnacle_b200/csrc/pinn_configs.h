// pinn_configs.h -- the (jet signature, output dim, tile shape) combinations this build instantiates.
//   X(ID, H, M, T, TN, TCT, orders...)   T/TN: FFMA kernel tile; TCT: points per tile of the tcgen05 kernel (0 = none;
//                                        8: two of the four column groups of element-wise threads own a quad)
// ID  problem classes
//  0  Poisson1D                                    (2)
//  1  Burgers1D                                    (2,1)
//  2  Poisson2D_Classic, PoissonBoltzmann2D, Poisson2D_ManyArea, Wave1D, Helmholtz2D   (2,2)
//  3  NS2D_LidDriven / BackStep / Classic          (2,2)  m=3
//  4  Heat2D_{Multiscale,VaryingCoef,ComplexGeometry,LongTime}   (2,2,1)
//  5  Burgers2D, GrayScott                         (2,2,1) m=2
//  6  NS2D_LongTime                                (2,2,1) m=3
//  7  Poisson3D_ComplexGeometry, Wave2D_*          (2,2,2)
//  8  KuramotoSivashinsky                          (4,1)
//  9  PoissonND                                    (2,2,2,2,2)
// 10  HeatND                                       (2,2,2,2,2,1)
// 11-18 value-type BC/IC/point-set terms (PINN_KIND_VALUE): jet-free signatures (0,..,0), d x m = 2x1 2x3 3x1 3x2 3x3 1x1 5x1 6x1
// 21-24 forward-Laplacian signatures LapSig<d> (orders entry -d), d = 2, 3, 5, 6: C = d + 2 channels for LINEAR-kind
//       residuals whose second derivatives enter through one weighted sum (Poisson*, Heat2D_{Multiscale,VaryingCoef,
//       ComplexGeometry}, Wave*, Helmholtz2D, PoissonND, HeatND); chosen automatically (jet_math.cuh: lap_eligible)
// 25-27 the same for the kinds whose only second derivatives are u_xx + u_yy: NS2D (d=2, m=3), Burgers2D / Gray-Scott
//       (d=3, m=2), NS2D_LongTime (d=3, m=3); Heat2D_LongTime uses 22
// 19-20 flux-type boundary terms (PINN_KIND_FLUX: Neumann / affine Robin): first-order signatures (1,1), (1,1,1), m = 1
// 29-30 PoissonInv (inverse.py:19-24; outputs u, a): (2,2) m=2 and its forward-Laplacian form LapSig<2> m=2 (HeatInv uses 5 / 26)
// 31    LapSig<3> with per-point direction weights (orders entry -(10+d)): Wave2D_Heterogeneous, u_xx + u_yy - u_tt / c(x,y)
// 33-42 / 43-52  hidden widths 64 and 128 (benchmark.py:65 --hidden-layers, src/utils/args.py:4-12), FP32 FFMA kernel only: Burgers1D (2,1);
//       the forward-Laplacian signatures of the Poisson2D / Heat2D / NS2D families (-2, -3, -2 with m = 3); Kuramoto-Sivashinsky (4,1);
//       value-type boundary terms 2x1, 2x3, 3x1; flux-type boundary terms (1,1), (1,1,1)
// 53-65 / 66-78  the remaining classes at hidden widths 64 and 128 (FFMA kernel): Burgers2D / Gray-Scott / HeatInv (-3, m = 2), Poisson1D (2),
//       PoissonND (-5), HeatND (-6), NS2D_LongTime (-3, m = 3), Wave2D_Heterogeneous ((2,2,2): its per-point direction weights exist in the tensor-core kernels only), PoissonInv (-2, m = 2); value-type boundary
//       terms 3x2, 3x3, 1x1, 5x1, 6x1; HeatND's 6-dimensional Neumann term
// 32    gPINN kinds (GP_LINEAR2, GP_BURGERS1D: residual + its input gradient, two inputs embedded along e0, e1, e0+e1, e0-e1): (3,3,3,3)
// 28    the same for the 6-dimensional Neumann term of HeatND (heat.py:296-301): (1,1,1,1,1,1)
#pragma once

#define PINN_FOR_EACH_CONFIG(X)        \
  X(0, 100, 1, 32, 4, 32, 2)               \
  X(1, 100, 1, 32, 4, 32, 2, 1)            \
  X(2, 100, 1, 32, 4, 32, 2, 2)            \
  X(3, 100, 3, 32, 4, 32, 2, 2)            \
  X(4, 100, 1, 32, 4, 16, 2, 2, 1)         \
  X(5, 100, 2, 32, 4, 16, 2, 2, 1)         \
  X(6, 100, 3, 32, 4, 16, 2, 2, 1)         \
  X(7, 100, 1, 16, 2, 16, 2, 2, 2)         \
  X(8, 100, 1, 32, 4, 16, 4, 1)            \
  X(9, 100, 1, 16, 2, 8, 2, 2, 2, 2, 2)   \
  X(10, 100, 1, 16, 2, 8, 2, 2, 2, 2, 2, 1) \
  X(11, 100, 1, 32, 4, 0, 0, 0)            \
  X(12, 100, 3, 32, 4, 0, 0, 0)            \
  X(13, 100, 1, 32, 4, 0, 0, 0, 0)         \
  X(14, 100, 2, 32, 4, 0, 0, 0, 0)         \
  X(15, 100, 3, 32, 4, 0, 0, 0, 0)         \
  X(16, 100, 1, 32, 4, 0, 0)               \
  X(17, 100, 1, 32, 4, 0, 0, 0, 0, 0, 0)   \
  X(18, 100, 1, 32, 4, 0, 0, 0, 0, 0, 0, 0) \
  X(19, 100, 1, 32, 4, 0, 1, 1)            \
  X(20, 100, 1, 32, 4, 0, 1, 1, 1)         \
  X(21, 100, 1, 32, 4, 32, -2)             \
  X(22, 100, 1, 32, 4, 32, -3)             \
  X(23, 100, 1, 16, 2, 16, -5)             \
  X(24, 100, 1, 16, 2, 16, -6)             \
  X(25, 100, 3, 32, 4, 32, -2)             \
  X(26, 100, 2, 32, 4, 32, -3)             \
  X(27, 100, 3, 32, 4, 32, -3)             \
  X(28, 100, 1, 16, 2, 0, 1, 1, 1, 1, 1, 1) \
  X(29, 100, 2, 32, 4, 32, 2, 2)           \
  X(30, 100, 2, 32, 4, 32, -2)             \
  X(31, 100, 1, 32, 4, 32, -13)           \
  X(32, 100, 1, 16, 2, 0, 3, 3, 3, 3)     \
  X(33, 64, 1, 32, 4, 0, 2, 1)            \
  X(34, 64, 1, 32, 4, 0, -2)              \
  X(35, 64, 1, 32, 4, 0, -3)              \
  X(36, 64, 3, 32, 4, 0, -2)              \
  X(37, 64, 1, 32, 4, 0, 4, 1)            \
  X(38, 64, 1, 32, 4, 0, 0, 0)            \
  X(39, 64, 3, 32, 4, 0, 0, 0)            \
  X(40, 64, 1, 32, 4, 0, 0, 0, 0)         \
  X(41, 64, 1, 32, 4, 0, 1, 1)            \
  X(42, 64, 1, 32, 4, 0, 1, 1, 1)         \
  X(43, 128, 1, 16, 2, 0, 2, 1)           \
  X(44, 128, 1, 16, 2, 0, -2)             \
  X(45, 128, 1, 16, 2, 0, -3)             \
  X(46, 128, 3, 16, 2, 0, -2)             \
  X(47, 128, 1, 16, 2, 0, 4, 1)           \
  X(48, 128, 1, 16, 2, 0, 0, 0)           \
  X(49, 128, 3, 16, 2, 0, 0, 0)           \
  X(50, 128, 1, 16, 2, 0, 0, 0, 0)        \
  X(51, 128, 1, 16, 2, 0, 1, 1)           \
  X(52, 128, 1, 16, 2, 0, 1, 1, 1)        \
  X(53, 64, 2, 32, 4, 0, -3) \
  X(54, 64, 1, 32, 4, 0, 2) \
  X(55, 64, 1, 16, 2, 0, -5) \
  X(56, 64, 1, 16, 2, 0, -6) \
  X(57, 64, 3, 32, 4, 0, -3) \
  X(58, 64, 1, 16, 2, 0, 2, 2, 2) \
  X(59, 64, 2, 32, 4, 0, -2) \
  X(60, 64, 2, 32, 4, 0, 0, 0, 0) \
  X(61, 64, 3, 32, 4, 0, 0, 0, 0) \
  X(62, 64, 1, 32, 4, 0, 0) \
  X(63, 64, 1, 32, 4, 0, 0, 0, 0, 0, 0) \
  X(64, 64, 1, 32, 4, 0, 0, 0, 0, 0, 0, 0) \
  X(65, 64, 1, 16, 2, 0, 1, 1, 1, 1, 1, 1) \
  X(66, 128, 2, 16, 2, 0, -3) \
  X(67, 128, 1, 16, 2, 0, 2) \
  X(68, 128, 1, 16, 2, 0, -5) \
  X(69, 128, 1, 16, 2, 0, -6) \
  X(70, 128, 3, 16, 2, 0, -3) \
  X(71, 128, 1, 16, 2, 0, 2, 2, 2) \
  X(72, 128, 2, 16, 2, 0, -2) \
  X(73, 128, 2, 16, 2, 0, 0, 0, 0) \
  X(74, 128, 3, 16, 2, 0, 0, 0, 0) \
  X(75, 128, 1, 16, 2, 0, 0) \
  X(76, 128, 1, 16, 2, 0, 0, 0, 0, 0, 0) \
  X(77, 128, 1, 16, 2, 0, 0, 0, 0, 0, 0, 0) \
  X(78, 128, 1, 16, 2, 0, 1, 1, 1, 1, 1, 1)

#define PINN_NUM_CONFIGS 79
