// pinn_api.cu -- C ABI (include/pinn_b200.h): descriptor validation, workspace carve-up, launches.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>

#include "pinn_configs.h"
#include "pinn_dispatch.h"
#include "pinn_helpers.cuh"

namespace pinn {

#define PINN_DECL(ID, ...) template <> const ConfigInfo* get_config<ID>();
PINN_FOR_EACH_CONFIG(PINN_DECL)
#undef PINN_DECL

static const ConfigInfo* const* all_configs() {
  static const ConfigInfo* table[PINN_NUM_CONFIGS] = {
#define PINN_ENTRY(ID, ...) get_config<ID>(),
      PINN_FOR_EACH_CONFIG(PINN_ENTRY)
#undef PINN_ENTRY
  };
  return table;
}

static thread_local std::string g_err;
static thread_local int g_launches = 0;
// Process-wide configuration knobs (the programmatic form of the PINN_IMPL / PINN_NO_LAPLACIAN environment variables):
// atomics that a call only READS.  They select among kernels that compute the same result, so a call racing with a
// setter runs one kernel or the other, never a mixture; no call mutates library state.
static std::atomic<int> g_no_lap{0};   // pinn_set_laplacian_collapse(0)
static std::atomic<int> g_impl{0};     // 0 = auto (tensor cores, pipelined where instantiated), 1 = FP32 FFMA kernel, 2 = tensor cores, one tile at a
                                       // time (error if absent), 3 = pipelined tensor-core kernel (falls back to 2 where not instantiated)

static int effective_impl() {
  int impl = g_impl.load(std::memory_order_relaxed);
  if (impl == 0) {
    const char* e = getenv("PINN_IMPL");
    if (e != nullptr) {
      if (strcmp(e, "ffma") == 0) impl = 1;
      if (strcmp(e, "tc") == 0) impl = 2;
      if (strcmp(e, "tc2") == 0) impl = 3;
    }
  }
  return impl;
}
// auto prefers the pipelined kernel where it is instantiated
static bool tc_pipelined() { const int impl = effective_impl(); return impl == 3 || impl == 0; }

static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

void set_error_string(const std::string& s) { g_err = s; }   // for the other translation units of the library (pinn_optim.cu)

static const ConfigInfo* find_config(const pinn_kind_desc* d, std::string* why) {
  if (d == nullptr) { *why = "null descriptor"; return nullptr; }
  if (d->d < 1 || d->d > PINN_MAX_DIM) { *why = "input dimension out of range"; return nullptr; }
  if (d->L < 2 || d->L > kMaxLayers) { *why = "L (hidden tanh layers) must be in 2.." + std::to_string(kMaxLayers); return nullptr; }
  if (d->n_pde < 1 || d->n_pde > PINN_MAX_PDE) { *why = "n_pde out of range"; return nullptr; }
  if (d->n_aux < 0 || d->n_aux > PINN_MAX_AUX) { *why = "n_aux out of range"; return nullptr; }
  for (int i = 0; i < PINN_MAX_AUXSEL; ++i)
    if (d->aux_sel[i] >= d->n_aux) { *why = "aux selector beyond n_aux"; return nullptr; }
  const ConfigInfo* const* tab = all_configs();
  // forward-Laplacian signature first where the descriptor allows it (fewer channels for the same result);
  // PINN_NO_LAPLACIAN=1 keeps the per-direction second-order jets (for A/B measurements)
  static const bool no_lap = [] { const char* e = getenv("PINN_NO_LAPLACIAN"); return e != nullptr && e[0] == '1'; }();
  // (per-point direction weights -- lap_eligible() == 2 -- exist in the tensor-core kernels only)
  const int lap_mode = lap_eligible(d->kind, d->d, d->m, d->order, d->aux_sel, d->consts);
  if (!no_lap && !g_no_lap.load(std::memory_order_relaxed) && (lap_mode == 1 || (lap_mode == 2 && effective_impl() != 1))) {
    for (int i = 0; i < PINN_NUM_CONFIGS; ++i) {
      const ConfigInfo* c = tab[i];
      if (c != nullptr && c->lap == lap_mode && c->H == d->H && c->m == d->m && c->d == d->d && c->valid_kind(d->kind)) return c;
    }
  }
  for (int i = 0; i < PINN_NUM_CONFIGS; ++i) {
    const ConfigInfo* c = tab[i];
    if (c == nullptr || c->lap || c->H != d->H || c->m != d->m || c->d != d->d) continue;
    bool same = true;
    for (int j = 0; j < d->d; ++j) same = same && (c->order[j] == d->order[j]);
    if (!same) continue;
    if (!c->valid_kind(d->kind)) continue;
    return c;
  }
  char buf[256];
  snprintf(buf, sizeof buf, "no kernel instantiated for kind=%d d=%d m=%d H=%d order=(%d,%d,%d,%d,%d,%d)", d->kind, d->d, d->m, d->H,
           d->order[0], d->order[1], d->order[2], d->order[3], d->order[4], d->order[5]);
  *why = buf;
  return nullptr;
}

static long long param_count(const pinn_kind_desc* d) {
  return (long long)d->d * d->H + d->H + (long long)(d->L - 1) * ((long long)d->H * d->H + d->H) + (long long)d->H * d->m + d->m;
}

struct DeviceInfo {
  int dev = -1;
  int sms = 0;
  bool ok = false;
  std::string err;
};

static const DeviceInfo& device_info() {
  static thread_local DeviceInfo di;
  static thread_local int cached_dev = -1;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    di.ok = false;
    di.err = std::string("cudaGetDevice: ") + cudaGetErrorString(e);
    return di;
  }
  if (dev != cached_dev || !di.ok) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
      di.ok = false;
      di.err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
      return di;
    }
    if (prop.major != 10) {
      di.ok = false;
      di.err = "libpinn_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
      return di;
    }
    di.sms = prop.multiProcessorCount;
    di.dev = dev;
    di.ok = true;
    cached_dev = dev;
  }
  return di;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  size_t off_wpack, off_stash, off_gpart, off_gwt, off_lpart, total;
  long long pstride;
};

static Workspace carve(const pinn_kind_desc* d, const ConfigInfo* c, int grid, int n_seeds, bool tc) {
  Workspace w;
  size_t o = 0;
  w.off_wpack = o;
  if (tc)
    o = align_up(o + (size_t)(2 * (d->L - 1) + 1) * kTcWMat * sizeof(uint4), 256);
  else
    o = align_up(o + (size_t)2 * (d->L - 1) * d->H * d->H * sizeof(float), 256);
  w.off_stash = o;
  if (tc)
    o = align_up(o + (size_t)grid * d->L * (c->tc_pipe && c->tc_NQ < 2 ? 2 : c->tc_NQ) * c->C * 512 * 16, 256);
  else
    o = align_up(o + (size_t)grid * d->L * c->E4 * c->NT * 16, 256);
  w.off_gpart = o;
  w.pstride = (param_count(d) + 3) / 4 * 4;
  o = align_up(o + (size_t)grid * (n_seeds > 0 ? n_seeds : 0) * w.pstride * sizeof(float), 256);
  w.off_gwt = o;
  if (tc) o = align_up(o + (size_t)grid * (n_seeds > 0 ? n_seeds : 0) * (d->L - 1) * kTcGwtLayer4 * 16, 256);
  w.off_lpart = o;
  o = align_up(o + (size_t)grid * PINN_MAX_PDE * sizeof(double), 256);
  w.total = o;
  return w;
}

static bool want_tc(const ConfigInfo* c, std::string* why) {
  const int impl = effective_impl();
  if (impl == 1) return false;
  if (c->tc_T > 0) return true;
  if (impl >= 2) *why = "PINN_IMPL=tc but this configuration has no tensor-core kernel";
  return false;
}

static int run(const pinn_kind_desc* desc, const float* params, const float* x, const float* aux, int64_t n_rows, double inv_n,
               const float* seeds, int n_seeds, float* grad_out, float* loss_out, float* resid_out, void* workspace,
               size_t workspace_bytes, void* stream_v, int lin_seed = 0) {
  g_launches = 0;
  std::string why;
  const ConfigInfo* c = find_config(desc, &why);
  if (c == nullptr) return fail(why);
  const bool bwd = n_seeds > 0;
  if (bwd && n_seeds > PINN_MAX_SEEDS) return fail("n_seeds exceeds PINN_MAX_SEEDS");
  if (params == nullptr || loss_out == nullptr || workspace == nullptr) return fail("null pointer argument");
  if (n_rows < 0) return fail("negative n_rows");
  if (n_rows > 0 && x == nullptr) return fail("x is null");
  if (desc->n_aux > 0 && n_rows > 0 && aux == nullptr) return fail("descriptor needs aux columns but aux is null");
  if (bwd && (seeds == nullptr || grad_out == nullptr)) return fail("null seeds/out");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail("workspace must be 256-byte aligned");
  const DeviceInfo& di = device_info();
  if (!di.ok) return fail(di.err);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  const int max_grid = di.sms;
  std::string why_tc;
  const bool tc = want_tc(c, &why_tc);
  if (!why_tc.empty()) return fail(why_tc);
  const Workspace ws = carve(desc, c, max_grid, n_seeds, tc);
  if (workspace_bytes < ws.total) return fail("workspace too small: need " + std::to_string(ws.total) + " bytes");

  // cudaFuncSetAttribute (dynamic shared memory above 48 KB) is per device: one process may drive several
  constexpr int kMaxDevices = 64;
  static thread_local bool prepared_tab[kMaxDevices][PINN_NUM_CONFIGS] = {};
  if (di.dev < 0 || di.dev >= kMaxDevices) return fail("device ordinal out of range");
  bool* prepared = prepared_tab[di.dev];
  if (!prepared[c->id]) {
    cudaError_t e = c->prepare();
    if (e == cudaSuccess && c->tc_T > 0) e = c->tc_prepare();
    if (e != cudaSuccess) return fail(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    prepared[c->id] = true;
  }

  const long long P = param_count(desc);
  const bool pipe = tc && c->tc_pipe && tc_pipelined();
  const int tileT = tc ? (pipe ? 32 : c->tc_T) : c->T;
  const long long ntiles = (n_rows + tileT - 1) / tileT;
  int grid = (int)(ntiles < max_grid ? ntiles : max_grid);
  char* wsb = static_cast<char*>(workspace);
  float* wpack = reinterpret_cast<float*>(wsb + ws.off_wpack);
  float* gpart = reinterpret_cast<float*>(wsb + ws.off_gpart);
  double* lpart = reinterpret_cast<double*>(wsb + ws.off_lpart);

  if (grid > 0 && tc) {
    uint4* wperm = reinterpret_cast<uint4*>(wpack);
    const uint4* wn = wperm;
    const uint4* wt = wperm + (size_t)(desc->L - 1) * kTcWMat;
    const uint4* w6p = wperm + (size_t)2 * (desc->L - 1) * kTcWMat;
    const int total = (2 * (desc->L - 1) + 1) * 4 * 512;
    pinn_pack_weights_tc<<<(total + 255) / 256, 256, 0, st>>>(params, wperm, desc->H, desc->d, desc->L, desc->m);
    ++g_launches;
    TcLaunchParams lp{};
    set_kind_params(lp.kp, desc->kind, desc->d, desc->aux_sel, desc->consts, c->lap);
    lp.kp.lin_seed = lin_seed;
    lp.L = desc->L; lp.n_aux = desc->n_aux; lp.n_pde = desc->n_pde; lp.n_seeds = n_seeds;
    lp.n_rows = n_rows; lp.P = P; lp.Ppad = ws.pstride; lp.inv_n = (float)inv_n;
    lp.params = params; lp.wn = wn; lp.wt = wt; lp.w6p = w6p; lp.x = x; lp.aux = aux; lp.seeds = seeds;
    lp.stash = reinterpret_cast<float*>(wsb + ws.off_stash);
    lp.gpart = gpart; lp.lpart = lpart; lp.resid_out = resid_out;
    lp.gwt = reinterpret_cast<float*>(wsb + ws.off_gwt);
    cudaError_t e = c->tc_launch(lp, bwd, pipe, grid, st);
    if (e != cudaSuccess) return fail(std::string("tensor-core kernel launch: ") + cudaGetErrorString(e));
    ++g_launches;
  } else if (grid > 0) {
    const int total = (desc->L - 1) * desc->H * desc->H;
    pinn_pack_weights<<<(total + 255) / 256, 256, 0, st>>>(params, wpack, desc->H, desc->d, desc->L);
    ++g_launches;

    LaunchParams lp{};
    set_kind_params(lp.kp, desc->kind, desc->d, desc->aux_sel, desc->consts, c->lap);
    lp.kp.lin_seed = lin_seed;
    lp.L = desc->L;
    lp.n_aux = desc->n_aux;
    lp.n_pde = desc->n_pde;
    lp.n_seeds = n_seeds;
    lp.n_rows = n_rows;
    lp.P = ws.pstride;
    lp.inv_n = (float)inv_n;
    lp.params = params;
    lp.wpack = wpack;
    lp.x = x;
    lp.aux = aux;
    lp.seeds = seeds;
    lp.stash = reinterpret_cast<float*>(wsb + ws.off_stash);
    lp.gpart = gpart;
    lp.lpart = lpart;
    lp.resid_out = resid_out;
    cudaError_t e = c->launch(lp, bwd, grid, st);
    if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
    ++g_launches;
  }
  const long long SP = (long long)n_seeds * P;
  const long long work = SP > 0 ? SP : 1;
  pinn_reduce_partials<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(gpart, lpart, grid, n_seeds, P, ws.pstride, desc->n_pde, inv_n, grad_out,
                                                                       loss_out, (tc && bwd && (pipe || c->tc_rmw)) ? reinterpret_cast<const float*>(wsb + ws.off_gwt) : nullptr,
                                                                       desc->H, desc->d, desc->L);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("reduce launch: ") + cudaGetErrorString(e));
  return 0;
}

}  // namespace pinn

using namespace pinn;

extern "C" {

int64_t pinn_param_count(const pinn_kind_desc* desc) { return desc ? param_count(desc) : -1; }

int pinn_supported(const pinn_kind_desc* desc) {
  std::string why;
  if (find_config(desc, &why) == nullptr) return fail(why);
  return 0;
}

size_t pinn_workspace_bytes(const pinn_kind_desc* desc, int n_seeds) {
  std::string why;
  const ConfigInfo* c = find_config(desc, &why);
  if (c == nullptr) { fail(why); return 0; }
  if (n_seeds < 0 || n_seeds > PINN_MAX_SEEDS) { fail("n_seeds out of range"); return 0; }
  const DeviceInfo& di = device_info();
  if (!di.ok) { fail(di.err); return 0; }
  const size_t a = carve(desc, c, di.sms, n_seeds, false).total;
  const size_t b = c->tc_T > 0 ? carve(desc, c, di.sms, n_seeds, true).total : 0;
  return a > b ? a : b;
}

int pinn_fwd(const pinn_kind_desc* desc, const float* params, const float* x, const float* aux, int64_t n_rows, double inv_n_global,
             float* loss_out, float* resid_out, void* workspace, size_t workspace_bytes, void* stream) {
  return run(desc, params, x, aux, n_rows, inv_n_global, nullptr, 0, nullptr, loss_out, resid_out, workspace, workspace_bytes, stream);
}

int pinn_fwd_bwd(const pinn_kind_desc* desc, const float* params, const float* x, const float* aux, int64_t n_rows, double inv_n_global,
                 const float* seeds, int n_seeds, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_seeds < 1) return fail("n_seeds must be >= 1");
  if (desc == nullptr || out == nullptr) return fail("null pointer argument");
  const long long P = param_count(desc);
  return run(desc, params, x, aux, n_rows, inv_n_global, seeds, n_seeds, out, out + (size_t)n_seeds * P, nullptr, workspace,
             workspace_bytes, stream);
}

int pinn_fwd_bwd_linear(const pinn_kind_desc* desc, const float* params, const float* x, const float* aux, int64_t n_rows,
                        const float* seeds, int n_seeds, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_seeds < 1) return fail("n_seeds must be >= 1");
  if (desc == nullptr || out == nullptr) return fail("null pointer argument");
  const long long P = param_count(desc);
  return run(desc, params, x, aux, n_rows, 1.0, seeds, n_seeds, out, out + (size_t)n_seeds * P, nullptr, workspace, workspace_bytes, stream, 1);
}

int pinn_last_launch_count(void) { return g_launches; }

int pinn_set_impl(int impl) {
  if (impl < 0 || impl > 3) return fail("impl must be 0 (auto), 1 (ffma), 2 (tensor cores) or 3 (pipelined tensor cores)");
  g_impl.store(impl, std::memory_order_relaxed);
  return 0;
}

int pinn_has_tensor_core_kernel(const pinn_kind_desc* desc) {
  std::string why;
  const ConfigInfo* c = find_config(desc, &why);
  return (c != nullptr && c->tc_T > 0) ? 1 : 0;
}

int pinn_set_laplacian_collapse(int enable) {
  g_no_lap.store(enable ? 0 : 1, std::memory_order_relaxed);
  return 0;
}

int pinn_effective_channels(const pinn_kind_desc* desc) {
  std::string why;
  const ConfigInfo* c = find_config(desc, &why);
  return c == nullptr ? 0 : c->C;
}

int pinn_selected_impl(const pinn_kind_desc* desc) {
  std::string why;
  const ConfigInfo* c = find_config(desc, &why);
  if (c == nullptr) return 0;
  std::string why_tc;
  if (!want_tc(c, &why_tc)) return why_tc.empty() ? 1 : 0;
  return (c->tc_pipe && tc_pipelined()) ? 3 : 2;
}

double pinn_ffma_peak_tflops(int iters) {
  const DeviceInfo& di = device_info();
  if (!di.ok) { fail(di.err); return -1.0; }
  float* out = nullptr;
  if (cudaMalloc(&out, 4) != cudaSuccess) { fail("cudaMalloc"); return -1.0; }
  const int blocks = di.sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  pinn_ffma_probe<<<blocks, threads>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  pinn_ffma_probe<<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (e != cudaSuccess || ms <= 0.f) { fail(std::string("ffma probe: ") + cudaGetErrorString(e)); return -1.0; }
  const double flops = 2.0 * 8 * 16 * (double)iters * threads * blocks;
  return flops / (ms * 1e-3) / 1e12;
}

const char* pinn_last_error(void) { return g_err.c_str(); }

#define PINN_STR2(x) #x
#define PINN_STR(x) PINN_STR2(x)
const char* pinn_build_info(void) {
  return "libpinn_b200 sm_100a fp32-ffma + tcgen05 bf16x3 jets; configs=" PINN_STR(PINN_NUM_CONFIGS) "; built " __DATE__ " " __TIME__;
}

}  // extern "C"
