"""The fused PDE-loss op as a differentiable PyTorch node.

``FusedPDEResidual`` owns one problem's descriptor and its device-resident collocation points
(+ per-point aux columns) and exposes ``losses(params)``: a vector of per-term mean-square PDE
losses whose ``backward`` can be called repeatedly with different ``grad_output`` vectors, exactly
like the list of 0-d loss tensors ``dde.data.PDE.losses`` builds (``deepxde/data/pde.py:147-152``)
and the reference optimizers back-propagate through (``deepxde/model.py:306-315``,
``src/optimizer/lr_adaptor.py:33-62``, ``src/optimizer/multiadam.py:247-261``).

Two gradient policies:

* ``eager``  -- forward runs ``pinn_fwd_bwd`` with the identity as seed matrix and keeps
  G[n_pde, P]; every later backward is ``g @ G`` (one kernel pass serves all backward calls);
* ``lazy``   -- forward runs ``pinn_fwd``; each backward runs ``pinn_fwd_bwd`` with the incoming
  ``grad_output`` as the single seed row (cached per distinct seed within one forward).

``auto`` picks eager for n_pde == 1, lazy otherwise.

Multi-GPU: every rank holds a contiguous shard of the rows and the same parameters; the flat
``[grads | losses]`` buffer the kernels write is all-reduced once (``torch.distributed``, NCCL over
NVLink) inside the op, so optimizers that compute non-linear statistics of per-term gradients
(``lr_adaptor.py:37-56``, ``ntk.py:38-63``) see global values.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import capi
from .kinds import KIND_FLUX, KindDesc
from .problems import ProblemSpec

# The compute backend: always the C ABI.  Kept as module attributes so the gloo/CPU tests of the
# sharding + all-reduce plumbing can substitute a checker explicitly; nothing in the product does.
_backend_fwd = capi.fwd
_backend_fwd_bwd = capi.fwd_bwd
_backend_fwd_bwd_linear = capi.fwd_bwd_linear      # gradient of the residual SUMS (NTK statistics); no CPU stand-in


def shard_bounds(n_rows: int, rank: int, world: int):
    """Contiguous row block of rank ``rank`` (sizes differ by at most one row)."""
    base, rem = divmod(n_rows, world)
    beg = rank * base + min(rank, rem)
    return beg, beg + base + (1 if rank < rem else 0)


def _all_reduce(buf: torch.Tensor, group):
    import torch.distributed as dist
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)


class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op: "FusedPDEResidual", *params: torch.Tensor):
        flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        desc = op.desc
        need_grad = any(ctx.needs_input_grad[1:])
        ctx.op = op
        ctx.shapes = [p.shape for p in params]
        ctx.mode = "none"
        if need_grad and op.policy_resolved == "eager":
            out = _backend_fwd_bwd(desc, flat, op.x, op.aux, op.inv_n_global, op._eye())
            op.stats["fwd_bwd_calls"] += 1
            if op.world > 1:
                _all_reduce(out, op.group)
                op.stats["allreduce_calls"] += 1
            S, P = desc.n_pde, desc.n_params
            ctx.G = out[:S * P].view(S, P)
            ctx.mode = "eager"
            return out[S * P:].clone()
        loss, _ = _backend_fwd(desc, flat, op.x, op.aux, op.inv_n_global)
        op.stats["fwd_calls"] += 1
        if op.world > 1:
            _all_reduce(loss, op.group)
            op.stats["allreduce_calls"] += 1
        if need_grad:
            ctx.mode = "lazy"
            ctx.flat = flat
            ctx.cache = []
        return loss

    @staticmethod
    def backward(ctx, g):
        op = ctx.op
        desc = op.desc
        if ctx.mode == "none":
            return (None,) + tuple(None for _ in ctx.shapes)
        g = g.detach().to(torch.float32).contiguous()
        if ctx.mode == "eager":
            flat_grad = g @ ctx.G
        else:
            flat_grad = None
            for seed, grad in ctx.cache:
                if torch.equal(seed, g):
                    flat_grad = grad
                    break
            if flat_grad is None:
                out = _backend_fwd_bwd(desc, ctx.flat, op.x, op.aux, op.inv_n_global, g.view(1, -1))
                op.stats["fwd_bwd_calls"] += 1
                if op.world > 1:
                    _all_reduce(out, op.group)
                    op.stats["allreduce_calls"] += 1
                flat_grad = out[:desc.n_params]
                ctx.cache.append((g.clone(), flat_grad))
        grads, off = [], 0
        for i, shp in enumerate(ctx.shapes):
            n = int(np.prod(shp))
            grads.append(flat_grad[off:off + n].view(shp) if ctx.needs_input_grad[1 + i] else None)
            off += n
        if desc.kind == KIND_FLUX and desc.aux_sel[1] < 0:
            # n.grad u does not depend on the output-layer bias: autograd reports None for that tensor in the reference, and
            # its LRA wrapper counts only the entries of non-None gradients (src/optimizer/lr_adaptor.py:50-54)
            grads[-1] = None
        return (None,) + tuple(grads)


class FusedPDEResidual:
    """Device-resident state of the fused op for one problem and one net shape."""

    def __init__(self, spec: ProblemSpec, H: int = 100, L: int = 5, policy: str = "auto", device=None, group=None,
                 check_library: bool = True):
        if policy not in ("auto", "eager", "lazy"):
            raise ValueError(f"unknown policy {policy!r}")
        self.spec = spec
        self.desc: KindDesc = spec.desc.with_net(H, L)
        self.policy = policy
        self.policy_resolved = ("eager" if self.desc.n_pde == 1 else "lazy") if policy == "auto" else policy
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.group = group
        self.world, self.rank = 1, 0
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        if check_library:
            capi.check_supported(self.desc)      # unsupported descriptor => error now, never a fallback
        self.x: Optional[torch.Tensor] = None
        self.aux: Optional[torch.Tensor] = None
        self.n_global = 0
        self.inv_n_global = 0.0
        self._eye_t = None
        self.stats = dict(fwd_calls=0, fwd_bwd_calls=0, allreduce_calls=0, h2d_bytes=0, uploads=0)

    def _eye(self):
        if self._eye_t is None or self._eye_t.device != self.device:
            self._eye_t = torch.eye(self.desc.n_pde, dtype=torch.float32, device=self.device)
        return self._eye_t

    def set_points(self, x, shard: bool = True, n_global: Optional[int] = None):
        """Upload the PDE rows and evaluate the per-point aux columns once.

        ``x``: numpy array or CPU/GPU tensor [N, d].  Default (``n_global=None``): every rank passes the
        same GLOBAL array and keeps its contiguous shard.  With ``n_global`` given the rows are already
        this rank's own share of ``n_global`` rows in total (no slicing)."""
        xt = torch.as_tensor(x)
        embed = getattr(self.spec, "embed", None)
        if embed is not None and xt.dim() == 2 and xt.shape[1] == embed.shape[0]:
            # embedded problem (gPINN): the kernels' extra "inputs" are jet directions, their coordinates are zero
            xt = torch.cat([xt, xt.new_zeros(xt.shape[0], self.desc.d - xt.shape[1])], dim=1)
        if xt.dim() != 2 or xt.shape[1] != self.desc.d:
            raise ValueError(f"points must be [N,{self.desc.d}], got {tuple(xt.shape)}")
        if n_global is not None:
            self.n_global = int(n_global)
        else:
            self.n_global = int(xt.shape[0])
            if shard and self.world > 1:
                beg, end = shard_bounds(self.n_global, self.rank, self.world)
                xt = xt[beg:end]
        if self.n_global == 0:
            raise ValueError("empty point set")
        self.inv_n_global = 1.0 / self.n_global
        xt = xt.to(torch.float32)
        if xt.device != self.device:
            self.stats["h2d_bytes"] += xt.numel() * 4
            self.stats["uploads"] += 1
            if self.device.type == "cuda":
                # the host -> device copy runs on a stream of its own: it does not queue behind whatever the compute stream is
                # still doing (another problem's kernels, the previous step's update), only the consumer waits for it
                up = getattr(self, "_upload_stream", None)
                if up is None:
                    up = self._upload_stream = torch.cuda.Stream(device=self.device)
                cur = torch.cuda.current_stream(self.device)
                with torch.cuda.stream(up):
                    xt = xt.contiguous().to(self.device, non_blocking=True)
                cur.wait_stream(up)
                xt.record_stream(cur)
            else:
                xt = xt.contiguous().to(self.device, non_blocking=True)
        self.x = xt.contiguous()
        self.aux = None
        if self.desc.n_aux:
            self.aux = self.spec.aux(self.x).to(torch.float32).contiguous()
        return self

    def set_rows(self, x_dev: torch.Tensor, aux_dev: Optional[torch.Tensor], n_global: Optional[int] = None):
        """Install already device-resident rows and aux columns (used for fused boundary terms)."""
        self.x = x_dev.to(torch.float32).contiguous()
        self.aux = None if aux_dev is None else aux_dev.to(torch.float32).contiguous()
        self.n_global = int(n_global if n_global is not None else self.x.shape[0])
        self.inv_n_global = 1.0 / max(self.n_global, 1)
        return self

    def losses(self, params: Sequence[torch.Tensor]) -> torch.Tensor:
        """Per-term mean-square PDE losses [n_pde] (global mean over all ranks' rows)."""
        if self.x is None:
            raise RuntimeError("set_points() has not been called")
        return _FusedLossFn.apply(self, *params)

    @torch.no_grad()
    def residual(self, params: Sequence[torch.Tensor]) -> torch.Tensor:
        """Residual matrix [n_local_rows, n_pde] (forward only; model.predict(operator=pde) analogue)."""
        flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        _, res = _backend_fwd(self.desc, flat, self.x, self.aux, self.inv_n_global, True)
        self.stats["fwd_calls"] += 1
        return res


def _is_adaptive_activation_net(net) -> bool:
    """DNN_LAAF / DNN_GAAF of src/model/laaf.py (benchmark.py --method laaf / gaaf): Sequential of LAAFlayer(fc, n, a) + Linear."""
    layers = getattr(net, "layers", None)
    if layers is None or not isinstance(getattr(net, "a", None), torch.Tensor) or len(layers) < 3:
        return False
    return all(hasattr(l, "fc") and hasattr(l, "a") and hasattr(l, "n") for l in list(layers)[:-1]) and isinstance(layers[-1], torch.nn.Linear)


def adaptive_activation_params(net) -> List[torch.Tensor]:
    """Effective FNN parameters of a layer-wise / global adaptive-activation net (src/model/laaf.py:5-80).

    LAAFlayer computes tanh(n * a * (W x + b)) with a trainable slope a (per neuron: DNN_LAAF, per layer: DNN_GAAF) and
    the constant n = 10.  That IS a plain tanh layer with W' = diag(n a) W, b' = n a * b, so the fused kernels run
    unchanged on (W', b') and autograd carries their gradient rows back to W, b and a through these few element-wise
    products (dL/da = n * (sum_i dL/dW'_ji W_ji + dL/db'_j b_j), summed over neurons for the global variant)."""
    out = []
    layers = list(net.layers)
    for lay in layers[:-1]:
        if not isinstance(lay.activation, torch.nn.Tanh):
            raise capi.FusedOpError(f"fused path supports the tanh activation only (got {lay.activation!r})")
        s = (lay.n * lay.a).reshape(-1)                    # [H] (LAAF) or [1] (GAAF)
        out += [s.reshape(-1, 1) * lay.fc.weight, s * lay.fc.bias]
    out += [layers[-1].weight, layers[-1].bias]
    return out


def _is_pfnn(net) -> bool:
    """dde.nn.PFNN (deepxde/nn/pytorch/fnn.py:50-150) / src.model.fnn.PFNN (with split_mask): every layer a ModuleList of one
    Linear per output -- the `recommend_net` of the inverse problems (src/pde/inverse.py:30-32, 143-148)."""
    layers = getattr(net, "layers", None)
    return (isinstance(layers, torch.nn.ModuleList) and len(layers) >= 3
            and all(isinstance(l, torch.nn.ModuleList) and all(isinstance(f, torch.nn.Linear) for f in l) for l in layers)
            and len({len(l) for l in layers}) == 1)


def pfnn_params(net) -> List[torch.Tensor]:
    """Effective FNN parameters of a fully parallel PFNN: independent sub-networks are ONE dense net whose hidden weight
    matrices are block diagonal (W' = blockdiag(W_0, .., W_{m-1}), b' = cat(b_j)), whose first layer stacks the branches
    (times each branch's input mask, ``split_mask``, when the net has one) and whose output layer puts branch j's
    Linear(h_j -> 1) in row j.  The fused kernels run unchanged on the dense form; autograd carries the gradient of the
    blocks back to the branch parameters (the off-diagonal zeros are constants)."""
    act = getattr(net, "activation", None)
    if getattr(act, "__name__", "") != "tanh" and not isinstance(act, torch.nn.Tanh):
        raise capi.FusedOpError(f"fused path supports the tanh activation only (got {act!r})")
    if getattr(net, "_input_transform", None) is not None or getattr(net, "_output_transform", None) is not None:
        raise capi.FusedOpError("fused path does not support input/output transforms (deepxde/nn/pytorch/nn.py:13-23)")
    layers = list(net.layers)
    m = len(layers[0])
    mask = getattr(net, "split_mask", None)
    out = []
    first = []
    for j, f in enumerate(layers[0]):
        w = f.weight
        if mask is not None:
            w = w * torch.as_tensor(mask[j], dtype=w.dtype, device=w.device).reshape(1, -1)
        first.append(w)
    out += [torch.cat(first, dim=0), torch.cat([f.bias for f in layers[0]])]
    for lay in layers[1:-1]:
        out += [torch.block_diag(*[f.weight for f in lay]), torch.cat([f.bias for f in lay])]
    last = layers[-1]
    if any(f.out_features != 1 for f in last):
        raise capi.FusedOpError("PFNN output sub-layers must map to one output each")
    rows = []
    widths = [f.in_features for f in last]
    for j, f in enumerate(last):
        left, right = sum(widths[:j]), sum(widths[j + 1:])
        rows.append(torch.nn.functional.pad(f.weight, (left, right)))
    out += [torch.cat(rows, dim=0), torch.cat([f.bias for f in last])]
    return out


def embedded_param_provider(base_fn, embed: np.ndarray):
    """Parameters of the embedded network of a gPINN problem: the first-layer weight times the direction matrix V (so that the
    kernels' input directions are the columns of V); autograd carries the gradient back to the real W1 through the product."""
    state = {}

    def fn():
        params = list(base_fn())
        w1 = params[0]
        V = state.get("V")
        if V is None or V.device != w1.device:
            V = state["V"] = torch.as_tensor(embed, dtype=w1.dtype).to(w1.device)
        return [w1 @ V] + params[1:]

    return fn


def net_param_provider(net):
    """(fn, dynamic): ``fn()`` returns the 2(L+1) effective ``nn.Linear``-order tensors the fused op differentiates with
    respect to.  Plain FNN: the parameters themselves (static list).  Adaptive-activation nets: recomputed from the live
    W, b, a on every call (dynamic)."""
    if _is_adaptive_activation_net(net):
        return (lambda: adaptive_activation_params(net)), True
    if _is_pfnn(net):
        return (lambda: pfnn_params(net)), True
    params = net_linear_params(net)
    return (lambda: params), False


def net_linear_params(net) -> List[torch.Tensor]:
    """The 2(L+1) nn.Linear tensors of a dde.nn.FNN-style net, in net.parameters() order
    (deepxde/nn/pytorch/fnn.py:25-33), after checking the net is one the fused path supports."""
    linears = getattr(net, "linears", None)
    if linears is None:
        raise capi.FusedOpError("fused path needs an FNN with a `.linears` ModuleList (deepxde/nn/pytorch/fnn.py)")
    if getattr(net, "_input_transform", None) is not None or getattr(net, "_output_transform", None) is not None:
        raise capi.FusedOpError("fused path does not support input/output transforms (deepxde/nn/pytorch/nn.py:13-23)")
    act = getattr(net, "activation", None)
    if isinstance(act, list) or act is None or getattr(act, "__name__", "") != "tanh":
        raise capi.FusedOpError(f"fused path supports the tanh activation only (got {act!r})")
    out = []
    for lin in linears:
        if lin.bias is None:
            raise capi.FusedOpError("fused path needs biased Linear layers")
        out += [lin.weight, lin.bias]
    return out


def net_shape(net):
    """(d, H, L, m) of an FNN (or of the FNN an adaptive-activation net is equivalent to); raises when hidden widths differ."""
    if _is_pfnn(net):
        layers = list(net.layers)
        d = layers[0][0].in_features
        widths = [sum(f.out_features for f in lay) for lay in layers[:-1]]
        if len(set(widths)) != 1:
            raise capi.FusedOpError("fused path needs equal hidden widths")
        for a, b in zip(layers[:-1], layers[1:]):
            if [f.out_features for f in a] != [f.in_features for f in b]:
                raise capi.FusedOpError("PFNN branches must keep their widths from layer to layer")
        return d, widths[0], len(layers) - 1, len(layers[0])
    lins = [l.fc for l in list(net.layers)[:-1]] + [net.layers[-1]] if _is_adaptive_activation_net(net) else list(net.linears)
    d = lins[0].in_features
    H = lins[0].out_features
    m = lins[-1].out_features
    for lin in lins[1:-1]:
        if lin.in_features != H or lin.out_features != H:
            raise capi.FusedOpError("fused path needs equal hidden widths")
    if lins[-1].in_features != H:
        raise capi.FusedOpError("fused path needs equal hidden widths")
    return d, H, len(lins) - 1, m
