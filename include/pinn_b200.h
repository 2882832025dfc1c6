/*
 * pinn_b200.h -- C ABI of the fused PDE-residual + parameter-gradient path (libpinn_b200.so).
 *
 * One call evaluates, for a tanh MLP u(x) and a block of collocation points, the PDE residual
 * r_k(x) of one PINNacle problem, the mean-square loss per residual term and (optionally) the
 * gradient of seeded combinations of those losses w.r.t. all network parameters.
 *
 * Reference interfaces this boundary replaces (paths relative to the PINNacle tree):
 *   - deepxde/model.py:258-279      outputs_losses(): net(x), data.losses(), stack
 *   - deepxde/nn/pytorch/fnn.py:35-48        FNN.forward
 *   - deepxde/gradients.py:160-181,258-283   dde.grad.jacobian / dde.grad.hessian
 *   - src/pde/<problem>.py          the pde(x, u) residual closures (one "kind" each, see enum)
 *   - deepxde/data/pde.py:147-152 + deepxde/losses.py:16-23   mean of squared residual rows
 *   - deepxde/model.py:306-315      closure(): sum(losses).backward() -> p.grad
 *
 * Conventions
 *   - All pointers are DEVICE pointers (fp32 unless stated), caller-owned; the library allocates
 *     nothing and keeps no state between calls.  `stream` is a cudaStream_t passed as void*.
 *   - `params` is the flat parameter vector in torch.nn.Linear / net.parameters() order:
 *     W0[H,d], b0[H], W1[H,H], b1[H], ..., W_L[m,H], b_L[m]   (row-major, y = x W^T + b).
 *   - `x` is [n_rows, d] row-major, `aux` is [n_rows, n_aux] row-major or NULL when n_aux == 0.
 *   - Losses are returned as  loss_k = inv_n_global * sum_rows r_k^2 , i.e. this rank's share of
 *     the global mean; summing the outputs of all ranks (one all-reduce) gives the global value.
 *   - Every function returns 0 on success, non-zero on error; pinn_last_error() describes the
 *     last error of the calling thread.  Unsupported descriptors are errors -- there is no
 *     fallback path.
 */
#ifndef PINN_B200_H_
#define PINN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_MAX_DIM 6
#define PINN_MAX_CONSTS 24
#define PINN_MAX_AUXSEL 8
#define PINN_MAX_PDE 3
#define PINN_MAX_AUX 8
#define PINN_MAX_SEEDS 4

/* Residual kinds.  Taylor coefficients: u_j = du/dx_j, u_jj = d2u/dx_j^2. */
enum pinn_kind {
  /* r = cu*[aux]*u + sum_j c1_j u_j + sum_j c2_j*[aux]*u_jj + [aux]   (poisson.py:54-58,111-120,179-193,
   * 253-258,288-296; heat.py:34-41,88-93,147-152,265-275; wave.py:22-26,85-89,144-149; helmholtz.py:19-27)
   * consts[0]=cu, consts[1+j]=c1_j, consts[7+j]=c2_j; aux_sel[0]=source column, aux_sel[1]=multiplier of
   * the cu term, aux_sel[2+j]=multiplier of the c2_j term (-1 = none). */
  PINN_KIND_LINEAR = 0,
  PINN_KIND_BURGERS1D = 1,     /* burgers.py:21-25   u_t + u u_x - c0 u_xx                      */
  PINN_KIND_BURGERS2D = 2,     /* burgers.py:66-80   2 outputs, 2 residuals, c0 = nu            */
  PINN_KIND_NS2D = 3,          /* ns.py:141-160      (u,v,p): momentum x/y + continuity, c0 = nu; c1 = 1 drops the convective terms (ns.py:51-67) */
  PINN_KIND_NS2D_UNSTEADY = 4, /* ns.py:330-352      + u_t, v_t, aux[aux_sel[0]] added to mom-y  */
  PINN_KIND_HEAT_NLSRC = 5,    /* heat.py:214-222    u_t - c0 lap u - c1 sin(c2 u^2) aux          */
  PINN_KIND_GRAYSCOTT = 6,     /* chaotic.py:21-31   c0,c1 = eps; c2 = b; c3 = d                  */
  PINN_KIND_KS = 7,            /* chaotic.py:77-83   u_t + c0 u u_x + c1 u_xx + c2 u_xxxx         */
  /* value-type boundary / initial / point-set term: r = u_c - g(x), c = (int)consts[0], g = aux[aux_sel[0]]
   * (deepxde/icbc/boundary_conditions.py:73-80,214-241; initial_conditions.py:29-36).  Jet-free signature (all orders 0). */
  PINN_KIND_VALUE = 8,
  /* flux-type boundary term (Neumann / affine Robin): r = sum_j n_j(x) du_c/dx_j + beta(x) u_c - g(x), c = (int)consts[0];
   * aux_sel[0] = g column, aux_sel[1] = beta column (-1: none), aux_sel[2+j] = column of the outward normal component
   * n_j (deepxde/icbc/boundary_conditions.py:44-57,83-105: NeumannBC.error, RobinBC.error with func affine in u).
   * First-order jets in every direction: signature (1,..,1). */
  PINN_KIND_FLUX = 9,
  /* inverse problems with an unknown diffusion field (src/pde/inverse.py:19-24 PoissonInv, 129-136 HeatInv), outputs (u, a):
   *   r = c0 u_t + c1 div(a grad u) + c2 aux[aux_sel[0]],   div(a grad u) = a (u_xx + u_yy) + a_x u_x + a_y u_y
   * (the reference differentiates the products a u_x, a u_y).  d = 2 (c0 = 0) or 3 (x, y, t); m = 2; orders (2,2[,1]). */
  PINN_KIND_DIV_AGRAD = 10,
  /* gPINN (src/pde/baseclass.py:151-180 use_gepinn, benchmark.py:91-92 --method gepinn): the residual of a one-output problem
   * in two inputs AND its gradient with respect to the inputs as loss terms -- r_0 = r, r_1 = dr/dx_0, r_2 = dr/dx_1 (the order
   * pde_wrapper appends them in).  dr/dx_j needs the mixed partials u_01, u_001, u_011: they come from order-3 Taylor jets along
   * the FOUR directions e0, e1, e0+e1, e0-e1 (polarisation: D^3_{e0+-e1} u = u_000 +- 3 u_001 + 3 u_011 +- u_111).  The caller
   * presents the 2-input network as a 4-"input" one: rows (x0, x1, 0, 0) and first-layer weights W1 [e0 e1 e0+e1 e0-e1], so
   * d = 4, orders (3,3,3,3); the gradient with respect to the embedded first layer maps back by the transposed direction matrix.
   *   GP_LINEAR2:   r = cu u + sum_{j<2} (c1_j u_j + c2_j u_jj) + src   (consts as LINEAR, constant coefficients;
   *                 aux_sel[0..2] = columns of src, dsrc/dx0, dsrc/dx1, -1 = none; aux_sel[3], aux_sel[4] = per-point multipliers of
   *                 the c2 terms / the cu term that the reference treats as constants under autograd (table look-ups), -1 = none)
   *   GP_BURGERS1D: r = u_t + u u_x - c0 u_xx,  inputs (x, t)                                   (burgers.py:21-25) */
  PINN_KIND_GP_LINEAR2 = 11,
  PINN_KIND_GP_BURGERS1D = 12
};

typedef struct pinn_kind_desc {
  int32_t kind;                    /* enum pinn_kind */
  int32_t d;                       /* input dimension  (1..PINN_MAX_DIM) */
  int32_t m;                       /* output dimension (1..3) */
  int32_t H;                       /* hidden width */
  int32_t L;                       /* number of hidden tanh layers (>= 2) */
  int32_t n_pde;                   /* residual terms (1..PINN_MAX_PDE) */
  int32_t n_aux;                   /* per-point aux columns (0..PINN_MAX_AUX) */
  int32_t order[PINN_MAX_DIM];     /* highest pure derivative carried per input direction */
  int32_t aux_sel[PINN_MAX_AUXSEL];
  float consts[PINN_MAX_CONSTS];
} pinn_kind_desc;

/* Number of fp32 parameters P of the network the descriptor names. */
int64_t pinn_param_count(const pinn_kind_desc* desc);

/* 0 if this build has kernels for the descriptor (jet signature, m, H, kind), else non-zero. */
int pinn_supported(const pinn_kind_desc* desc);

/* Bytes of device scratch the two calls below need for up to n_seeds reverse passes. */
size_t pinn_workspace_bytes(const pinn_kind_desc* desc, int n_seeds);

/* Forward only: loss_out[k] = inv_n_global * sum_rows r_k^2, k < n_pde.
 * resid_out (nullable): [n_rows, n_pde] residual matrix.  Replaces model.py:258-279 for
 * `_test()` / `predict(operator=pde)` style evaluations (model.py:476-499, 857-871). */
int pinn_fwd(const pinn_kind_desc* desc, const float* params, const float* x, const float* aux,
             int64_t n_rows, double inv_n_global, float* loss_out, float* resid_out,
             void* workspace, size_t workspace_bytes, void* stream);

/* Forward + hand-derived reverse.  seeds: device [n_seeds, n_pde] (grad_output over the loss terms
 * for each reverse pass, 1 <= n_seeds <= PINN_MAX_SEEDS).  out: device [n_seeds*P + n_pde]:
 *   out[s*P + i]      = d( sum_k seeds[s][k] * loss_k ) / d params[i]      (written, not accumulated)
 *   out[n_seeds*P+k]  = loss_k
 * so that a single all-reduce(SUM) of `out` yields global gradients and losses
 * (replaces model.py:306-315 and the per-term backward passes of src/optimizer/lr_adaptor.py:33-62,
 * multiadam.py:247-261). */
int pinn_fwd_bwd(const pinn_kind_desc* desc, const float* params, const float* x, const float* aux,
                 int64_t n_rows, double inv_n_global, const float* seeds, int n_seeds, float* out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Same pass with LINEAR seeds:  out[s*P + i] = d( sum_k seeds[s][k] * sum_rows r_k ) / d params[i]  -- the gradient of the
 * residual SUMS, which src/optimizer/ntk.py:24-58 builds with loss_fn = (y - x).sum() for its NTK-trace statistics.
 * out[n_seeds*P + k] still holds sum_rows r_k^2 (inv_n_global = 1). */
int pinn_fwd_bwd_linear(const pinn_kind_desc* desc, const float* params, const float* x, const float* aux, int64_t n_rows,
                        const float* seeds, int n_seeds, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Kernel selection: 0 = automatic (tcgen05 tensor-core kernel where one is instantiated -- its pipelined variant
 * where that exists -- else the FP32 FFMA kernel), 1 = FP32 FFMA kernel, 2 = tensor-core kernel, one tile at a time
 * (calls fail where none exists), 3 = pipelined tensor-core kernel (two sub-tiles in flight, warp-specialised;
 * falls back to 2 for configurations it is not instantiated for).  The environment variable
 * PINN_IMPL=ffma|tc|tc2 has the same effect when this is left at 0.  All kernels implement the same interface to
 * the same tolerance; the switch exists so they can be measured against each other. */
int pinn_set_impl(int impl);
/* 1 if the descriptor's configuration has a tensor-core kernel in this build. */
int pinn_has_tensor_core_kernel(const pinn_kind_desc* desc);
/* The kernel pinn_fwd / pinn_fwd_bwd would run for this descriptor under the current selection: 1 = FP32 FFMA,
 * 2 = tensor cores (one tile at a time), 3 = tensor cores (pipelined); 0 = unsupported descriptor / selection. */
int pinn_selected_impl(const pinn_kind_desc* desc);
/* Jet channels per neuron the kernels carry for this descriptor: 1 + sum(order) normally; d + 2 when the residual
 * uses its second derivatives only through one weighted sum (LINEAR kind) and the forward-Laplacian signature
 * [u, du/dx_1..d, sum_j c2_j d2u/dx_j^2] is propagated instead of separate second-order coefficients.  Same results
 * to rounding; pinn_set_laplacian_collapse(0) (or PINN_NO_LAPLACIAN=1) switches the collapse off for comparison. */
int pinn_effective_channels(const pinn_kind_desc* desc);
int pinn_set_laplacian_collapse(int enable);

/* Number of kernel launches the last pinn_fwd / pinn_fwd_bwd call of this thread enqueued. */
int pinn_last_launch_count(void);

/* ---- diagnostics (not part of the hot path; unlike the calls above these two manage their own resources) ----
 * FP32 FFMA micro-benchmark used for the compute roofline denominator: runs `iters` dependent-free
 * FFMA per thread on every SM and returns achieved TFLOP/s (negative on error).  Allocates and frees 4 bytes of device
 * memory, launches on the default stream and synchronises the device: call it outside timed regions and CUDA graphs. */
double pinn_ffma_peak_tflops(int iters);

/* Hardware self-test of the tcgen05 descriptor conventions the fused kernels rely on (caller-owned buffers, no allocation): one CTA computes
 * D[128][N] = A[128][K] * B[K][N] (tf32 x tf32 -> fp32, row-major device arrays) with the operands staged exactly
 * as the fused kernels stage them.  mode 0: A in TMEM, B in shared memory MN-major; mode 1: A and B in shared
 * memory, K-major.  variant bit 0 swaps the descriptor LBO/SBO fields (diagnostic).  K % 8 == 0, N % 16 == 0. */
int pinn_debug_umma_gemm(int mode, int variant, int K, int N, const float* A, const float* B, float* D, void* stream);

/* ---- optimizer-side kernels over the flat parameter / gradient buffers (one launch each, no host synchronisation) ----
 *
 * pinn_adam_step: torch.optim.Adam's update (amsgrad off; weight_decay = L2 term added to the gradient), the optimizer
 * benchmark.py:107 constructs, over the flat buffers instead of 12 tensors.  step_dev: device float[2] = {updates applied
 * so far, 0}; the kernel reads it for the bias corrections and increments it, so a captured CUDA graph can replay it. */
int pinn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* step_dev,
                   float lr, float beta1, float beta2, float eps, float weight_decay, void* stream);
/* One MultiAdam update (src/optimizer/multiadam.py:54-122 `sadam`; amsgrad off, no aggregated momentum -- what
 * benchmark.py:108 constructs) from gradient rows G[n_rows][ldg]: row t belongs to loss group row_group[t] (device int32)
 * with factor row_weight[t] (its loss weight); per group k one Adam moment pair exp_avg / exp_avg_sq [n_groups][P]; the
 * parameter update is  -lr/(1-beta1^t) * sum_k group_weights[k] * m_k / (sqrt(v_k)/sqrt(1-beta2^t) + eps).
 * The reference runs one backward pass per group and a stack/copy of every moment tensor per parameter per step. */
int pinn_multiadam_step(float* params, const float* G, int n_rows, int64_t P, int64_t ldg, const int32_t* row_group,
                        const float* row_weight, int n_groups, const float* group_weights, float* exp_avg, float* exp_avg_sq,
                        float* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay, void* stream);
/* Floats of device scratch pinn_grad_stats / pinn_lra_combine need. */
size_t pinn_optim_scratch_floats(void);
/* out3 = { max|g|, sum|g|, sum g^2 } of a flat vector: the statistics src/optimizer/lr_adaptor.py:37-41,50-53 and
 * ntk.py:38-58 compute with one torch reduction + .item() per parameter tensor. */
int pinn_grad_stats(const float* g, int64_t n, float* out3, float* scratch, void* stream);
/* One learning-rate-annealing update (src/optimizer/lr_adaptor.py:33-62, mode "max") from the per-term gradient rows
 * G[n_terms][ldg] of the UNWEIGHTED loss terms (PDE terms first):
 *   m_r = max_i |sum_{k<n_pde} G[k][i]|;   for t >= n_pde:  w_t <- (1-alpha) w_t + alpha m_r / (sum_i|G[t][i]| / counts[t] * w_t)
 *   grad_out[i] = sum_t w_t G[t][i]        (the gradient lr_adaptor.py:59-62 back-propagates, with the new weights)
 * w_in / w_out: device [n_terms] (may not alias); counts: device [n_terms], the element count the reference's mean
 * runs over (parameters whose gradient is not None).  At most 15 boundary terms. */
int pinn_lra_combine(const float* G, int n_terms, int n_pde, int64_t P, int64_t ldg, const float* w_in, const float* counts,
                     float alpha, float* w_out, float* grad_out, float* scratch, void* stream);

/* ---- collocation points drawn on the device (replaces the host resampling of dde.callbacks.PDEPointResampler,
 * deepxde/callbacks.py:563-568 -> PDE.resample_train_points, deepxde/data/pde.py:186-193 -> geom.random_points(n, "pseudo")) ----
 *
 * Region = axis-aligned box [lo, hi) in d coordinates, optionally intersected with one closed ball over the first keep_dims
 * coordinates (keep_dims = 0: none), minus up to PINN_MAX_HOLES closed holes, each a ball (a = centre, b[0] = radius) or a box
 * (a = low corner, b = high corner) over the FIRST n_dims coordinates: the shapes src/pde builds its domains from (Rectangle /
 * Cuboid / Hypercube minus Disk / Sphere / Rectangle; HeatND's Hypersphere; times a TimeDomain). */
#define PINN_MAX_HOLES 32
enum pinn_hole_shape { PINN_HOLE_BALL = 0, PINN_HOLE_BOX = 1 };
typedef struct pinn_hole {
  int32_t shape;  /* pinn_hole_shape */
  int32_t n_dims; /* 1..3 */
  float a[3];
  float b[3];
} pinn_hole;
typedef struct pinn_region {
  int32_t d; /* 1..PINN_MAX_DIM */
  int32_t n_holes;
  float lo[PINN_MAX_DIM];
  float hi[PINN_MAX_DIM];
  int32_t keep_dims;                /* 0, or 1..d: points outside the ball below are redrawn */
  float keep_radius;
  float keep_center[PINN_MAX_DIM];
  pinn_hole holes[PINN_MAX_HOLES];
} pinn_region;
/* x[r][0..d) <- uniform point number first_row + r of the stream (key, draw), r in [0, n_rows); x is a device array with row
 * stride ld >= d floats (columns d..ld are left alone); `region` is a HOST struct, copied into the launch.  Row g of a stream
 * is a pure function of (key, draw, g) -- Philox4x32-10 on the counter (g_lo, g_hi, attempt, block | draw << 8), coordinate
 * 4*block + j = lo + (word_j >> 8) * 2^-24 * (hi - lo) in float32, a point outside the kept ball or inside a hole redrawn with
 * attempt + 1 (at most 256 attempts) -- so ranks that draw disjoint row ranges of the same stream produce the rows one rank would.  draw < 2^24. */
int pinn_sample_points(float* x, int64_t ld, int64_t n_rows, const pinn_region* region, uint64_t key, uint32_t draw,
                       uint64_t first_row, void* stream);

const char* pinn_last_error(void);
const char* pinn_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* PINN_B200_H_ */
