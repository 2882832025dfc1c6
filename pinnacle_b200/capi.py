"""ctypes binding of ``libpinn_b200.so`` (C ABI declared in ``include/pinn_b200.h``).

The library is the product's only compute path.  If it has not been built
(``python pinnacle_b200/csrc/build.py`` or ``__graft_entry__.build()``) loading raises
``FusedLibraryMissing`` -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from .kinds import CKindDesc, CRegion, KindDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
# PINN_B200_LIB: an alternative build of the same library (what-if builds under variants/, see scripts/whatif_times.sh)
LIB_PATH = os.environ.get("PINN_B200_LIB") or os.path.join(_HERE, "csrc", "libpinn_b200.so")

EXPORTS = (
    "pinn_param_count", "pinn_supported", "pinn_workspace_bytes", "pinn_fwd", "pinn_fwd_bwd", "pinn_fwd_bwd_linear",
    "pinn_last_launch_count", "pinn_ffma_peak_tflops", "pinn_last_error", "pinn_build_info", "pinn_debug_umma_gemm", "pinn_set_impl", "pinn_has_tensor_core_kernel",
    "pinn_selected_impl", "pinn_effective_channels", "pinn_set_laplacian_collapse",
    "pinn_adam_step", "pinn_optim_scratch_floats", "pinn_grad_stats", "pinn_lra_combine", "pinn_multiadam_step",
    "pinn_sample_points",
)


class FusedLibraryMissing(RuntimeError):
    pass


class FusedOpError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FusedLibraryMissing(
            f"{LIB_PATH} not found: build it with `python pinnacle_b200/csrc/build.py` "
            "(nvcc, sm_100a). The fused path has no fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER(CKindDesc)
    vp = ctypes.c_void_p
    lib.pinn_param_count.argtypes = [P]
    lib.pinn_param_count.restype = ctypes.c_int64
    lib.pinn_supported.argtypes = [P]
    lib.pinn_supported.restype = ctypes.c_int
    lib.pinn_workspace_bytes.argtypes = [P, ctypes.c_int]
    lib.pinn_workspace_bytes.restype = ctypes.c_size_t
    lib.pinn_fwd.argtypes = [P, vp, vp, vp, ctypes.c_int64, ctypes.c_double, vp, vp, vp, ctypes.c_size_t, vp]
    lib.pinn_fwd.restype = ctypes.c_int
    lib.pinn_fwd_bwd.argtypes = [P, vp, vp, vp, ctypes.c_int64, ctypes.c_double, vp, ctypes.c_int, vp, vp, ctypes.c_size_t, vp]
    lib.pinn_fwd_bwd.restype = ctypes.c_int
    lib.pinn_last_launch_count.argtypes = []
    lib.pinn_last_launch_count.restype = ctypes.c_int
    lib.pinn_ffma_peak_tflops.argtypes = [ctypes.c_int]
    lib.pinn_ffma_peak_tflops.restype = ctypes.c_double
    lib.pinn_debug_umma_gemm.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    lib.pinn_debug_umma_gemm.restype = ctypes.c_int
    lib.pinn_set_impl.argtypes = [ctypes.c_int]
    lib.pinn_set_impl.restype = ctypes.c_int
    lib.pinn_has_tensor_core_kernel.argtypes = [P]
    lib.pinn_has_tensor_core_kernel.restype = ctypes.c_int
    lib.pinn_effective_channels.argtypes = [P]
    lib.pinn_effective_channels.restype = ctypes.c_int
    lib.pinn_set_laplacian_collapse.argtypes = [ctypes.c_int]
    lib.pinn_set_laplacian_collapse.restype = ctypes.c_int
    V, F, I64 = ctypes.c_void_p, ctypes.c_float, ctypes.c_int64
    lib.pinn_adam_step.argtypes = [V, V, V, V, I64, V, F, F, F, F, F, V]
    lib.pinn_adam_step.restype = ctypes.c_int
    lib.pinn_multiadam_step.argtypes = [V, V, ctypes.c_int, I64, I64, V, V, ctypes.c_int, V, V, V, V, F, F, F, F, F, V]
    lib.pinn_multiadam_step.restype = ctypes.c_int
    lib.pinn_fwd_bwd_linear.argtypes = [P, V, V, V, I64, V, ctypes.c_int, V, V, ctypes.c_size_t, V]
    lib.pinn_fwd_bwd_linear.restype = ctypes.c_int
    lib.pinn_optim_scratch_floats.argtypes = []
    lib.pinn_optim_scratch_floats.restype = ctypes.c_size_t
    lib.pinn_grad_stats.argtypes = [V, I64, V, V, V]
    lib.pinn_grad_stats.restype = ctypes.c_int
    lib.pinn_lra_combine.argtypes = [V, ctypes.c_int, ctypes.c_int, I64, I64, V, V, F, V, V, V, V]
    lib.pinn_lra_combine.restype = ctypes.c_int
    lib.pinn_selected_impl.argtypes = [P]
    lib.pinn_selected_impl.restype = ctypes.c_int
    lib.pinn_sample_points.argtypes = [V, I64, I64, ctypes.POINTER(CRegion), ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, V]
    lib.pinn_sample_points.restype = ctypes.c_int
    lib.pinn_last_error.argtypes = []
    lib.pinn_last_error.restype = ctypes.c_char_p
    lib.pinn_build_info.argtypes = []
    lib.pinn_build_info.restype = ctypes.c_char_p
    _lib = lib
    return lib


def last_error() -> str:
    return load().pinn_last_error().decode()


def _check(rc: int, what: str):
    if rc != 0:
        raise FusedOpError(f"{what}: {last_error()}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise FusedOpError(f"{name} must be a CUDA tensor (got {t.device}); the fused path has no CPU implementation")
    if t.dtype != dtype:
        raise FusedOpError(f"{name} must be {dtype} (got {t.dtype})")
    if not t.is_contiguous():
        raise FusedOpError(f"{name} must be contiguous")


def supported(desc: KindDesc) -> bool:
    c = desc.to_c()
    return load().pinn_supported(ctypes.byref(c)) == 0


def check_supported(desc: KindDesc):
    c = desc.to_c()
    _check(load().pinn_supported(ctypes.byref(c)), f"descriptor {desc.name or desc.kind} unsupported")


def workspace_bytes(desc: KindDesc, n_seeds: int) -> int:
    c = desc.to_c()
    n = load().pinn_workspace_bytes(ctypes.byref(c), int(n_seeds))
    if n == 0:
        raise FusedOpError(f"pinn_workspace_bytes: {last_error()}")
    return int(n)


class Workspace:
    """Caller-owned device scratch (torch caching allocator), grown on demand.

    One buffer per (device, stream): calls enqueued on one stream serialise, so they may share scratch; calls on different
    streams (two ops running concurrently, a graph being captured on a side stream) get different buffers and cannot race
    on the stash / partial accumulators.  A buffer whose pointer was handed out while its stream was capturing is baked
    into a CUDA graph: it is kept alive for the life of the process and never replaced in its slot by regrowth, so a later
    replay can never write into memory the caching allocator has given to someone else."""

    def __init__(self):
        self._buf = {}
        self._captured = []            # buffers a CUDA graph may hold raw pointers to

    def get(self, device: torch.device, nbytes: int) -> torch.Tensor:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(device)
        capturing = torch.cuda.is_current_stream_capturing()
        key = (device.type, idx, int(stream.cuda_stream))
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes + 256:
            b = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            self._buf[key] = b
        if capturing and not any(c is b for c in self._captured):
            self._captured.append(b)
        off = (-b.data_ptr()) % 256
        return b[off:off + nbytes]

    def n_buffers(self) -> int:
        return len(self._buf)


_WS = Workspace()


# PINN_NVTX=1: every C-ABI call below is bracketed by an NVTX range named after it and its residual kind, so that a timeline
# (ncu --nvtx, nsys where available) shows which loss term a kernel belongs to (tracing)
_NVTX = os.environ.get("PINN_NVTX", "0") == "1"


class _nvtx:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _NVTX:
            torch.cuda.nvtx.range_pop()


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def fwd(desc: KindDesc, params: torch.Tensor, x: torch.Tensor, aux: Optional[torch.Tensor], inv_n_global: float,
        want_residual: bool = False):
    """pinn_fwd: returns (loss[n_pde] device tensor, residual[n_rows,n_pde] or None)."""
    lib = load()
    _req(params, "params"); _req(x, "x")
    if aux is not None:
        _req(aux, "aux")
    n = x.shape[0]
    if params.numel() != desc.n_params:
        raise FusedOpError(f"params has {params.numel()} elements, descriptor needs {desc.n_params}")
    if x.dim() != 2 or x.shape[1] != desc.d:
        raise FusedOpError(f"x must be [N,{desc.d}]")
    if desc.n_aux and (aux is None or tuple(aux.shape) != (n, desc.n_aux)):
        raise FusedOpError(f"aux must be [N,{desc.n_aux}]")
    c = desc.to_c()
    with torch.cuda.device(x.device), _nvtx(f"pinn_fwd:{desc.name or desc.kind}"):
        ws = _WS.get(x.device, workspace_bytes(desc, 0))
        loss = torch.empty(desc.n_pde, dtype=torch.float32, device=x.device)
        resid = torch.empty((n, desc.n_pde), dtype=torch.float32, device=x.device) if want_residual else None
        rc = lib.pinn_fwd(ctypes.byref(c), _ptr(params), _ptr(x), _ptr(aux), n, float(inv_n_global), _ptr(loss), _ptr(resid),
                          _ptr(ws), ws.numel(), _stream_ptr(x.device))
    _check(rc, "pinn_fwd")
    return loss, resid


def fwd_bwd(desc: KindDesc, params: torch.Tensor, x: torch.Tensor, aux: Optional[torch.Tensor], inv_n_global: float,
            seeds: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pinn_fwd_bwd: returns the flat device buffer [S*P grads | n_pde losses]."""
    lib = load()
    _req(params, "params"); _req(x, "x"); _req(seeds, "seeds")
    if aux is not None:
        _req(aux, "aux")
    n = x.shape[0]
    if params.numel() != desc.n_params:
        raise FusedOpError(f"params has {params.numel()} elements, descriptor needs {desc.n_params}")
    if x.dim() != 2 or x.shape[1] != desc.d:
        raise FusedOpError(f"x must be [N,{desc.d}]")
    if desc.n_aux and (aux is None or tuple(aux.shape) != (n, desc.n_aux)):
        raise FusedOpError(f"aux must be [N,{desc.n_aux}]")
    if seeds.dim() != 2 or seeds.shape[1] != desc.n_pde:
        raise FusedOpError(f"seeds must be [S,{desc.n_pde}]")
    S = seeds.shape[0]
    c = desc.to_c()
    with torch.cuda.device(x.device), _nvtx(f"pinn_fwd_bwd:{desc.name or desc.kind}"):
        ws = _WS.get(x.device, workspace_bytes(desc, S))
        if out is None:
            out = torch.empty(S * desc.n_params + desc.n_pde, dtype=torch.float32, device=x.device)
        else:
            _req(out, "out")
            if out.numel() != S * desc.n_params + desc.n_pde:
                raise FusedOpError("out has the wrong size")
        rc = lib.pinn_fwd_bwd(ctypes.byref(c), _ptr(params), _ptr(x), _ptr(aux), n, float(inv_n_global), _ptr(seeds), S,
                              _ptr(out), _ptr(ws), ws.numel(), _stream_ptr(x.device))
    _check(rc, "pinn_fwd_bwd")
    return out


IMPL_AUTO, IMPL_FFMA, IMPL_TC, IMPL_TC_PIPELINED = 0, 1, 2, 3


def set_impl(impl: int):
    """0 = auto, 1 = FP32 FFMA kernel, 2 = tcgen05 tensor-core kernel, 3 = pipelined tcgen05 kernel (falls back to 2 where not instantiated)."""
    _check(load().pinn_set_impl(int(impl)), "pinn_set_impl")


def has_tensor_core_kernel(desc: KindDesc) -> bool:
    c = desc.to_c()
    return bool(load().pinn_has_tensor_core_kernel(ctypes.byref(c)))


def effective_channels(desc: KindDesc) -> int:
    """Jet channels per neuron the kernels carry for ``desc`` (d + 2 under the forward-Laplacian collapse, else 1 + sum(order))."""
    c = desc.to_c()
    return int(load().pinn_effective_channels(ctypes.byref(c)))


def set_laplacian_collapse(enable: bool):
    _check(load().pinn_set_laplacian_collapse(1 if enable else 0), "pinn_set_laplacian_collapse")


def selected_impl(desc: KindDesc) -> int:
    """Kernel the next call would run for ``desc``: IMPL_FFMA, IMPL_TC or IMPL_TC_PIPELINED (0 = unsupported)."""
    c = desc.to_c()
    return int(load().pinn_selected_impl(ctypes.byref(c)))


IMPL_NAMES = {1: "pinn_fused_kernel (FP32 FFMA)", 2: "pinn_tc_kernel (tcgen05 bf16x3)",
              3: "pinn_tc2_kernel (tcgen05 bf16x3, warp-specialised, two sub-tiles in flight)"}


def last_launch_count() -> int:
    return int(load().pinn_last_launch_count())


def ffma_peak_tflops(iters: int = 4096) -> float:
    v = float(load().pinn_ffma_peak_tflops(int(iters)))
    if v < 0:
        raise FusedOpError(f"pinn_ffma_peak_tflops: {last_error()}")
    return v


def build_info() -> str:
    return load().pinn_build_info().decode()


# ---- optimizer-side kernels over the flat buffers -------------------------------------------------------
_optim_scratch = {}


def _scratch(device) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    b = _optim_scratch.get(key)
    if b is None:
        b = torch.empty(int(load().pinn_optim_scratch_floats()), dtype=torch.float32, device=device)
        _optim_scratch[key] = b
    return b


def adam_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step_dev: torch.Tensor,
              lr: float, beta1: float, beta2: float, eps: float, weight_decay: float = 0.0):
    """pinn_adam_step: one launch of torch.optim.Adam's update over flat fp32 buffers; `step_dev` = device float[2]."""
    for t, nm in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq"), (step_dev, "step_dev")):
        _req(t, nm)
    n = params.numel()
    if grads.numel() != n or exp_avg.numel() != n or exp_avg_sq.numel() != n or step_dev.numel() < 2:
        raise FusedOpError("adam_step: buffer sizes differ")
    with torch.cuda.device(params.device):
        rc = load().pinn_adam_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), n, _ptr(step_dev), float(lr), float(beta1),
                                   float(beta2), float(eps), float(weight_decay), _stream_ptr(params.device))
    _check(rc, "pinn_adam_step")


def grad_stats(g: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pinn_grad_stats: device tensor [max|g|, sum|g|, sum g^2] of a flat vector, no host synchronisation."""
    _req(g, "g")
    if out is None:
        out = torch.empty(3, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        rc = load().pinn_grad_stats(_ptr(g), g.numel(), _ptr(out), _ptr(_scratch(g.device)), _stream_ptr(g.device))
    _check(rc, "pinn_grad_stats")
    return out


def lra_combine(G: torch.Tensor, n_pde: int, w_in: torch.Tensor, counts: torch.Tensor, alpha: float, w_out: torch.Tensor,
                grad_out: torch.Tensor):
    """pinn_lra_combine: G [n_terms, P] rows of the unweighted terms -> new weights (w_out) and grad_out = sum_t w_t G[t]."""
    for t, nm in ((G, "G"), (w_in, "w_in"), (counts, "counts"), (w_out, "w_out"), (grad_out, "grad_out")):
        _req(t, nm)
    if G.dim() != 2:
        raise FusedOpError("lra_combine: G must be [n_terms, P]")
    T, P = G.shape
    if w_in.numel() != T or w_out.numel() != T or counts.numel() != T or grad_out.numel() != P or w_in.data_ptr() == w_out.data_ptr():
        raise FusedOpError("lra_combine: buffer sizes differ or w_in aliases w_out")
    with torch.cuda.device(G.device):
        rc = load().pinn_lra_combine(_ptr(G), int(T), int(n_pde), int(P), int(G.stride(0)), _ptr(w_in), _ptr(counts), float(alpha), _ptr(w_out),
                                     _ptr(grad_out), _ptr(_scratch(G.device)), _stream_ptr(G.device))
    _check(rc, "pinn_lra_combine")


def multiadam_step(params: torch.Tensor, G: torch.Tensor, row_group: torch.Tensor, row_weight: torch.Tensor, group_weights: torch.Tensor,
                   exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step_dev: torch.Tensor, lr: float, beta1: float, beta2: float, eps: float,
                   weight_decay: float = 0.0):
    """pinn_multiadam_step: one MultiAdam update from the gradient rows G [n_rows, P] (row t: group row_group[t], factor row_weight[t])."""
    for t, nm in ((params, "params"), (G, "G"), (row_weight, "row_weight"), (group_weights, "group_weights"), (exp_avg, "exp_avg"),
                  (exp_avg_sq, "exp_avg_sq"), (step_dev, "step_dev")):
        _req(t, nm)
    _req(row_group, "row_group", torch.int32)
    T, P = G.shape
    ng = group_weights.numel()
    if params.numel() != P or row_group.numel() != T or row_weight.numel() != T or exp_avg.numel() != ng * P or exp_avg_sq.numel() != ng * P:
        raise FusedOpError("multiadam_step: buffer sizes differ")
    with torch.cuda.device(params.device):
        rc = load().pinn_multiadam_step(_ptr(params), _ptr(G), int(T), int(P), int(G.stride(0)), _ptr(row_group), _ptr(row_weight), int(ng),
                                        _ptr(group_weights), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(step_dev), float(lr), float(beta1), float(beta2),
                                        float(eps), float(weight_decay), _stream_ptr(params.device))
    _check(rc, "pinn_multiadam_step")


def fwd_bwd_linear(desc: KindDesc, params: torch.Tensor, x: torch.Tensor, aux: Optional[torch.Tensor], seeds: torch.Tensor,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pinn_fwd_bwd_linear: rows d(sum_k seeds[s,k] * sum_rows r_k)/dtheta (gradient of the residual SUMS, ntk.py:24-58);
    returns the flat device buffer [S*P | sum_rows r_k^2]."""
    lib = load()
    _req(params, "params"); _req(x, "x"); _req(seeds, "seeds")
    if aux is not None:
        _req(aux, "aux")
    n = x.shape[0]
    if params.numel() != desc.n_params or x.dim() != 2 or x.shape[1] != desc.d:
        raise FusedOpError("fwd_bwd_linear: params / x do not match the descriptor")
    if desc.n_aux and (aux is None or tuple(aux.shape) != (n, desc.n_aux)):
        raise FusedOpError(f"aux must be [N,{desc.n_aux}]")
    if seeds.dim() != 2 or seeds.shape[1] != desc.n_pde:
        raise FusedOpError(f"seeds must be [S,{desc.n_pde}]")
    S = seeds.shape[0]
    c = desc.to_c()
    with torch.cuda.device(x.device), _nvtx(f"pinn_fwd_bwd_linear:{desc.name or desc.kind}"):
        ws = _WS.get(x.device, workspace_bytes(desc, S))
        if out is None:
            out = torch.empty(S * desc.n_params + desc.n_pde, dtype=torch.float32, device=x.device)
        else:
            _req(out, "out")
            if out.numel() != S * desc.n_params + desc.n_pde:
                raise FusedOpError("out has the wrong size")
        rc = lib.pinn_fwd_bwd_linear(ctypes.byref(c), _ptr(params), _ptr(x), _ptr(aux), n, _ptr(seeds), S, _ptr(out), _ptr(ws), ws.numel(),
                                     _stream_ptr(x.device))
    _check(rc, "pinn_fwd_bwd_linear")
    return out


def sample_points(x: torch.Tensor, region: CRegion, key: int, draw: int, first_row: int = 0):
    """pinn_sample_points: fill the rows of the device array ``x`` [n, ld] (its first ``region.d`` columns) with points
    first_row .. first_row + n of the uniform stream (key, draw) over ``region`` -- one launch, nothing crosses PCIe."""
    _req(x, "x")
    if x.dim() != 2 or x.shape[1] < region.d:
        raise FusedOpError(f"sample_points: x must be [n, >= {region.d}], got {tuple(x.shape)}")
    with torch.cuda.device(x.device):
        rc = load().pinn_sample_points(_ptr(x), x.shape[1], x.shape[0], ctypes.byref(region), int(key) & 0xFFFFFFFFFFFFFFFF, int(draw),
                                       int(first_row), _stream_ptr(x.device))
    _check(rc, "pinn_sample_points")
    return x
