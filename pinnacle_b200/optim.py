"""Optimizers over ONE flat parameter buffer.

The fused op already produces the gradient as a flat vector; the reference then spreads it over 12 ``nn.Linear`` tensors
and updates them with the foreach ``torch.optim.Adam`` (benchmark.py:107), and its learning-rate-annealing wrapper
(src/optimizer/lr_adaptor.py) runs one extra backward pass plus ``.item()`` synchronisations per boundary term.  Here:

* ``FlatAdam``   -- ``torch.optim.Adam`` semantics (same hyper-parameters, same update) as one kernel launch
  (``pinn_adam_step``) over a flat buffer the network's parameters are views of.  Works as a plain
  ``torch.optim.Optimizer`` (closure protocol, ``opt.losses``) with any loss; with a ``FusedPDESolver`` whose terms are all
  fused it also skips autograd: the solver writes the weighted gradient straight into ``flat_grad``.
* ``FusedLRA``   -- the ``LR_Adaptor`` algorithm (mode "max") on the per-term gradient ROWS of the fused op: one fused
  pass per term, then ``pinn_lra_combine`` computes max|g_r|, mean|g_t|, the new weights and the re-weighted total
  gradient in two launches with every scalar on the device -- no per-term backward passes, no host synchronisation.

Both need the CUDA library; there is no CPU implementation.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import capi


class FlatAdam(torch.optim.Optimizer):
    """Adam (amsgrad off) over a flat fp32 buffer; ``p.data`` / ``p.grad`` of every parameter become views of
    ``self.flat`` / ``self.flat_grad``.  ``zero_grad()`` zeroes the flat gradient in one memset and keeps the views
    attached (``set_to_none`` is ignored: a parameter no loss term reaches sees a zero gradient instead of being
    skipped, which is the same update as long as it has never received a gradient)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(params)
        if not params:
            raise ValueError("FlatAdam got an empty parameter list")
        # the group carries torch.optim.Adam's remaining keys at their defaults so that a FlatAdam checkpoint loads into the
        # reference's stock Adam (its step() reads them from the restored group)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                                      capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False))
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam supports one parameter group")
        dev = params[0].device
        for p in params:
            if p.device != dev or p.dtype != torch.float32 or not p.is_cuda:
                raise capi.FusedOpError("FlatAdam needs fp32 CUDA parameters on one device (no CPU implementation)")
        self._params = params
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(2, dtype=torch.float32, device=dev)     # {updates applied, ticket}: see pinn_adam_step
        self._slices = []
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                self._slices.append((off, k))
                off += k
        self._attach_grads()

    def _attach_grads(self):
        for p, (off, k) in zip(self._params, self._slices):
            want = self.flat_grad[off:off + k]
            if p.grad is None or p.grad.data_ptr() != want.data_ptr():
                p.grad = want.view(p.shape)

    def zero_grad(self, set_to_none: bool = True):
        self.flat_grad.zero_()
        self._attach_grads()

    def _gather_grads(self):
        """Bring gradients that autograd re-pointed (p.grad replaced instead of accumulated into) back into flat_grad."""
        for p, (off, k) in zip(self._params, self._slices):
            want = self.flat_grad[off:off + k]
            if p.grad is None:
                want.zero_()
            elif p.grad.data_ptr() != want.data_ptr():
                want.copy_(p.grad.reshape(-1))
        self._attach_grads()

    def apply_update(self):
        """One Adam update of ``flat`` from ``flat_grad`` (one kernel launch)."""
        g = self.param_groups[0]
        capi.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_dev, g["lr"], g["betas"][0], g["betas"][1],
                       g["eps"], g["weight_decay"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._gather_grads()
        self.apply_update()
        return loss

    # ---- checkpointing (deepxde/model.py:945-948 saves opt.state_dict(); restore() loads it back) -------------------
    def state_dict(self):
        """torch.optim.Adam's own layout -- ``state[i] = {step, exp_avg, exp_avg_sq}`` per parameter, as copies shaped like
        the parameter -- so a checkpoint written here loads into the reference's ``torch.optim.Adam`` (benchmark.py:107) and
        back; ``flat_state`` carries the same moments as flat vectors (the fast path of ``load_state_dict``)."""
        sd = super().state_dict()
        step = self.step_dev[:1].detach().clone().cpu().reshape(())
        sd["state"] = {i: {"step": step.clone(), "exp_avg": self.exp_avg[off:off + k].detach().clone().view(p.shape),
                           "exp_avg_sq": self.exp_avg_sq[off:off + k].detach().clone().view(p.shape)}
                       for i, (p, (off, k)) in enumerate(zip(self._params, self._slices))}
        sd["flat_state"] = {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.step_dev[:1].clone()}
        return sd

    def load_state_dict(self, state_dict):
        """Accepts a FlatAdam checkpoint (``flat_state``) or one written by ``torch.optim.Adam`` for the same parameters
        (per-parameter ``exp_avg`` / ``exp_avg_sq`` / ``step`` are scattered into the flat buffers).  A checkpoint with
        optimizer state in neither form raises: resuming with silently reset moments is not an option."""
        state_dict = dict(state_dict)
        flat_state = state_dict.pop("flat_state", None)
        per_param = state_dict.get("state", {}) or {}
        state_dict["state"] = {}                        # the flat buffers are the state; nothing lives in self.state
        super().load_state_dict(state_dict)
        if flat_state is not None:
            self.exp_avg.copy_(flat_state["exp_avg"])
            self.exp_avg_sq.copy_(flat_state["exp_avg_sq"])
            self.step_dev.zero_()
            self.step_dev[:1].copy_(flat_state["step"])
            return
        if not per_param:                               # an optimizer that has never stepped
            self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.step_dev.zero_()
            return
        ids = sorted(per_param)
        if len(ids) != len(self._params):
            raise capi.FusedOpError(f"FlatAdam.load_state_dict: checkpoint has state for {len(ids)} parameters, optimizer has {len(self._params)}")
        steps = set()
        for i, (p, (off, k)) in zip(ids, zip(self._params, self._slices)):
            st = per_param[i]
            if not all(key in st for key in ("exp_avg", "exp_avg_sq", "step")):
                raise capi.FusedOpError("FlatAdam.load_state_dict: per-parameter state is not torch.optim.Adam's (exp_avg, exp_avg_sq, step)")
            if st["exp_avg"].numel() != k:
                raise capi.FusedOpError("FlatAdam.load_state_dict: parameter shapes differ from the checkpoint's")
            self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(float(st["step"]))
        if len(steps) != 1:
            raise capi.FusedOpError(f"FlatAdam.load_state_dict: parameters carry different step counts {sorted(steps)}")
        self.step_dev.zero_()
        self.step_dev[:1].fill_(steps.pop())

    def fused_step(self, solver, w_dev: Optional[torch.Tensor] = None):
        """One training step of a ``FusedPDESolver`` whose loss terms are all fused, without autograd: every term's fused
        pass writes its gradient row, ``flat_grad = sum_t w_t G_t`` (one matrix-vector product), one Adam launch.
        Sets ``self.losses`` to the weighted per-term losses like the reference closure (deepxde/model.py:311)."""
        n_pde = solver.spec.desc.n_pde
        T = solver.num_loss
        dev = self.flat.device
        if w_dev is not None:            # live weights already on the device (FusedNTK): no host round trip
            if getattr(self, "_comb", None) is None or self._comb.numel() != 1 + T - n_pde:
                self._comb = torch.ones(1 + T - n_pde, dtype=torch.float32, device=dev)
            self._comb[1:] = w_dev[n_pde:]
            self._w_dev, self._w_host = None, None
        else:
            w = solver.loss_weights
            w_dev = getattr(self, "_w_dev", None)
            w_host = None if w is None else np.asarray(w, dtype=np.float32).copy()
            if w_dev is None or w_dev.numel() != T or (w_host is not None and not np.array_equal(w_host, self._w_host)) or (w_host is None) != (self._w_host is None):
                self._w_host = w_host
                self._w_dev = w_dev = torch.ones(T, dtype=torch.float32, device=dev) if w_host is None else torch.as_tensor(w_host).to(dev)
                self._comb = torch.ones(1 + T - n_pde, dtype=torch.float32, device=dev)
                self._comb[1:] = w_dev[n_pde:]
        G, losses = solver.term_gradient_rows(self.flat, w_dev[:n_pde].view(1, n_pde))
        torch.mv(G.t(), self._comb, out=self.flat_grad)
        self.apply_update()
        self.losses = losses * w_dev
        return self.losses.sum()


class FusedLRA(torch.optim.Optimizer):
    """Learning-rate annealing (src/optimizer/lr_adaptor.py:4-64, mode "max") driven by the fused op's per-term gradient
    rows.  Same constructor protocol as the reference class -- ``FusedLRA(inner, loss_weight, num_pde, alpha)`` -- with
    ``inner`` a ``FlatAdam``; ``attach(solver)`` names the ``FusedPDESolver`` whose terms it balances.  The live weights
    stay on the device (``self.weights``); ``loss_weight`` (the host array the reference mutates) is refreshed lazily by
    ``sync_loss_weight()`` so a step never waits for the GPU."""

    def __init__(self, optimizer: FlatAdam, loss_weight, num_pde: int, alpha: float = 0.1, mode: str = "max"):
        if mode != "max":
            raise ValueError("FusedLRA implements mode='max' (the only one the reference accepts, lr_adaptor.py:16)")
        if not isinstance(optimizer, FlatAdam):
            raise capi.FusedOpError("FusedLRA wraps a FlatAdam")
        super().__init__(optimizer.param_groups, dict(alpha=alpha, mode=mode))
        self.optimizer = optimizer
        self.loss_weight = loss_weight
        self.num_pde = int(num_pde)
        self.alpha = float(alpha)
        self.iter = 0
        self.solver = None
        dev = optimizer.flat.device
        w = np.asarray(loss_weight, dtype=np.float32)
        self.weights = torch.as_tensor(w.copy()).to(dev)
        self._w_next = torch.empty_like(self.weights)
        self.counts = torch.full((w.shape[0],), float(optimizer.flat.numel()), dtype=torch.float32, device=dev)
        self._rows_w = None

    def attach(self, solver):
        if solver.spec.desc.n_pde != self.num_pde or solver.num_loss != self.weights.numel():
            raise capi.FusedOpError("FusedLRA: loss_weight / num_pde do not match the solver's loss terms")
        self.solver = solver
        return self

    def zero_grad(self, set_to_none: bool = True):
        self.optimizer.zero_grad(set_to_none)

    def sync_loss_weight(self):
        """Copy the device weights into the host ``loss_weight`` array (one D2H; call when logging)."""
        self.loss_weight[:] = self.weights.detach().cpu().numpy()
        return self.loss_weight

    def state_dict(self):
        return {"inner": self.optimizer.state_dict(), "weights": self.weights.clone(), "iter": self.iter}

    def load_state_dict(self, state_dict):
        self.optimizer.load_state_dict(state_dict["inner"])
        self.weights.copy_(state_dict["weights"])
        self.iter = int(state_dict["iter"])
        self.sync_loss_weight()

    def step(self, closure=None):
        if self.solver is None:
            raise capi.FusedOpError("FusedLRA.step(): attach(solver) first")
        self.iter += 1
        opt, n_pde = self.optimizer, self.num_pde
        w = self.weights
        T = w.numel()
        pde_w = w[:n_pde]
        # the PDE terms keep their (constant) weights; with equal weights c one row  sum_k G_k  serves both the statistic
        # max|grad sum_k L_k| and the combination c * row; unequal weights need the individual rows
        if self._rows_w is None:
            hw = np.asarray(self.loss_weight, dtype=np.float32)[:n_pde]
            self._pde_equal = bool(np.all(hw == hw[0]))
            if self._pde_equal:
                self._seeds = torch.ones(1, n_pde, dtype=torch.float32, device=w.device)
                self._rows_w = torch.empty(1 + T - n_pde, dtype=torch.float32, device=w.device)
                self._rows_w_next = torch.empty_like(self._rows_w)
                self._rows_counts = torch.cat([self.counts[:1], self.counts[n_pde:]]).contiguous()
            else:
                self._seeds = torch.eye(n_pde, dtype=torch.float32, device=w.device)
                self._rows_w = torch.empty(T, dtype=torch.float32, device=w.device)
                self._rows_w_next = torch.empty_like(self._rows_w)
                self._rows_counts = self.counts
        G, losses = self.solver.term_gradient_rows(opt.flat, self._seeds)
        tc = getattr(getattr(self.solver, "core", None), "term_counts", None)
        if tc is not None and tuple(tc) != getattr(self, "_counts_key", None):
            # per boundary term: the number of parameters its gradient reaches (P - m for Neumann terms, see gradient_rows)
            self._counts_key = tuple(tc)
            self._rows_counts = self._rows_counts.clone()
            self._rows_counts[self._rows_counts.numel() - len(tc):] = torch.as_tensor(tc, dtype=torch.float32, device=w.device)
        if self._pde_equal:
            self._rows_w[:1] = pde_w[:1]
            self._rows_w[1:] = w[n_pde:]
            n_rows_pde = 1
        else:
            self._rows_w.copy_(w)
            n_rows_pde = n_pde
        self.losses = losses * w                       # what the closure reports: weighted with the weights of this step
        capi.lra_combine(G, n_rows_pde, self._rows_w, self._rows_counts, self.alpha, self._rows_w_next, opt.flat_grad)
        w[n_pde:] = self._rows_w_next[n_rows_pde:]
        opt.apply_update()
        return torch.sum(losses * w)                   # lr_adaptor.py:60: total loss under the NEW weights


class FlatMultiAdam(torch.optim.Optimizer):
    """MultiAdam (src/optimizer/multiadam.py:125-330 with its defaults: amsgrad off, no aggregated momentum, no parameter
    scheduler) on the fused op's gradient rows: one Adam moment pair per loss group, update = group-weighted sum of the
    groups' normalised first moments -- ONE kernel launch (``pinn_multiadam_step``) after one fused pass per term, where the
    reference runs one backward pass per group and stacks / copies every moment tensor of every parameter each step.

    Same constructor protocol: ``FlatMultiAdam(params, lr, betas, eps, weight_decay, loss_group_idx=[num_pde], group_weights)``
    (benchmark.py:108); ``attach(solver)`` names the ``FusedPDESolver``.  The loss weights of the terms are read from
    ``solver.loss_weights`` (constant, as in the reference where MultiAdam differentiates the weighted losses)."""

    def __init__(self, params, lr=1e-3, betas=(0.99, 0.99), eps=1e-8, weight_decay=0.0, loss_group_idx=None, group_weights=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = FlatAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)     # owns the flat buffers / views
        self.flat = self._flat.flat
        self.loss_group_idx = list(loss_group_idx or [])
        self.n_groups = len(self.loss_group_idx) + 1
        dev = self.flat.device
        gw = torch.full((self.n_groups,), 1.0 / self.n_groups) if group_weights is None else torch.as_tensor(group_weights, dtype=torch.float32)
        self.group_weights = gw.to(dev).contiguous()
        n = self.flat.numel()
        self.exp_avg = torch.zeros(self.n_groups * n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_groups * n, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        self.solver = None
        self._rows = None

    def attach(self, solver):
        self.solver = solver
        self._rows = None
        return self

    def zero_grad(self, set_to_none: bool = True):
        self._flat.zero_grad(set_to_none)

    def state_dict(self):
        sd = super().state_dict()
        sd["flat_state"] = {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.step_dev[:1].clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat_state = state_dict.pop("flat_state", None)
        super().load_state_dict(state_dict)
        if flat_state is not None:
            self.exp_avg.copy_(flat_state["exp_avg"])
            self.exp_avg_sq.copy_(flat_state["exp_avg_sq"])
            self.step_dev.zero_()
            self.step_dev[:1].copy_(flat_state["step"])

    def step(self, closure=None):
        if self.solver is None:
            raise capi.FusedOpError("FlatMultiAdam.step(): attach(solver) first")
        solver = self.solver
        n_pde, T = solver.spec.desc.n_pde, solver.num_loss
        dev = self.flat.device
        if self._rows is None:
            w = np.ones(T, dtype=np.float32) if solver.loss_weights is None else np.asarray(solver.loss_weights, dtype=np.float32)
            bounds = [0] + self.loss_group_idx + [T]
            group_of = np.zeros(T, dtype=np.int32)
            for k in range(len(bounds) - 1):
                group_of[bounds[k]:bounds[k + 1]] = k
            # PDE terms of one group share a fused pass (seed row = their loss weights); a group boundary inside the PDE terms
            # needs one seed row per piece
            pieces = sorted(set([0, n_pde] + [b for b in self.loss_group_idx if 0 < b < n_pde]))
            seeds = np.zeros((len(pieces) - 1, n_pde), dtype=np.float32)
            row_group, row_weight = [], []
            for r in range(len(pieces) - 1):
                seeds[r, pieces[r]:pieces[r + 1]] = w[pieces[r]:pieces[r + 1]]
                row_group.append(group_of[pieces[r]])
                row_weight.append(1.0)
            for t in range(n_pde, T):
                row_group.append(group_of[t])
                row_weight.append(float(w[t]))
            self._seeds = torch.as_tensor(seeds).to(dev)
            self._row_group = torch.as_tensor(np.asarray(row_group, dtype=np.int32)).to(dev)
            self._row_weight = torch.as_tensor(np.asarray(row_weight, dtype=np.float32)).to(dev)
            self._w_dev = torch.as_tensor(w).to(dev)
            self._rows = True
        G, losses = solver.term_gradient_rows(self.flat, self._seeds)
        g = self.param_groups[0]
        capi.multiadam_step(self.flat, G, self._row_group, self._row_weight, self.group_weights, self.exp_avg, self.exp_avg_sq, self.step_dev,
                            g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"])
        self.losses = losses * self._w_dev
        return self.losses.sum()


class FusedNTK(torch.optim.Optimizer):
    """NTK-based loss balancing (src/optimizer/ntk.py:6-66) on the fused op's LINEAR gradient rows.

    The reference evaluates, every step, the gradient of the SUM of all PDE residuals and of the sum of all boundary errors
    with stock autograd (``data.losses`` with ``loss_fn = (y - x).sum()``), takes m_r = |g_r|^2, m_b = |g_b|^2, sets the PDE
    weights to (m_r + m_b) / m_r and the boundary weights to (m_r + m_b) / m_b, then lets the inner optimizer take a normal
    step under the new weights.  Here the two gradients come from one ``pinn_fwd_bwd_linear`` pass per term, the norms from
    ``pinn_grad_stats``, the weights never leave the device, and the inner ``FlatAdam`` step is the autograd-free one.
    Same constructor protocol as ``FusedLRA``; ``sync_loss_weight()`` refreshes the host array."""

    def __init__(self, optimizer: FlatAdam, loss_weight, num_pde: int):
        if not isinstance(optimizer, FlatAdam):
            raise capi.FusedOpError("FusedNTK wraps a FlatAdam")
        super().__init__(optimizer.param_groups, {})
        self.optimizer, self.loss_weight, self.num_pde, self.iter, self.solver = optimizer, loss_weight, int(num_pde), 0, None
        dev = optimizer.flat.device
        self.weights = torch.as_tensor(np.asarray(loss_weight, dtype=np.float32).copy()).to(dev)
        self._seeds = torch.ones(1, self.num_pde, dtype=torch.float32, device=dev)
        self._sr = torch.empty(3, dtype=torch.float32, device=dev)
        self._sb = torch.empty(3, dtype=torch.float32, device=dev)

    def attach(self, solver):
        if solver.spec.desc.n_pde != self.num_pde or solver.num_loss != self.weights.numel():
            raise capi.FusedOpError("FusedNTK: loss_weight / num_pde do not match the solver's loss terms")
        self.solver = solver
        return self

    def zero_grad(self, set_to_none: bool = True):
        self.optimizer.zero_grad(set_to_none)

    def sync_loss_weight(self):
        self.loss_weight[:] = self.weights.detach().cpu().numpy()
        return self.loss_weight

    def state_dict(self):
        return {"inner": self.optimizer.state_dict(), "weights": self.weights.clone(), "iter": self.iter}

    def load_state_dict(self, state_dict):
        self.optimizer.load_state_dict(state_dict["inner"])
        self.weights.copy_(state_dict["weights"])
        self.iter = int(state_dict["iter"])
        self.sync_loss_weight()

    def step(self, closure=None):
        if self.solver is None:
            raise capi.FusedOpError("FusedNTK.step(): attach(solver) first")
        self.iter += 1
        opt = self.optimizer
        G, _ = self.solver.term_gradient_rows(opt.flat, self._seeds, linear=True)
        capi.grad_stats(G[0], out=self._sr)
        g_b = G[1:].sum(dim=0) if G.shape[0] > 2 else G[1]
        capi.grad_stats(g_b.contiguous(), out=self._sb)
        m_r, m_b = self._sr[2], self._sb[2]
        tot = m_r + m_b
        self.weights[:self.num_pde] = tot / m_r            # ntk.py:60-63
        self.weights[self.num_pde:] = tot / m_b
        loss = opt.fused_step(self.solver, w_dev=self.weights)
        self.losses = opt.losses
        return loss
