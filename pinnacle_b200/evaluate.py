"""Device-resident evaluation around the fused path.

The reference evaluates through host arrays: ``Model._test()`` pulls the network outputs of every training and test row to
numpy each ``display_every`` iterations (deepxde/model.py:765-799), ``rar_wrapper`` pulls the residual of every training row
to argsort it (src/utils/rar.py:20-32) and ``TesterCallback`` pulls the predictions on the solution table
(src/utils/callbacks.py:161-197).  At 16 Mi collocation points those copies are hundreds of MB per evaluation.  Here the
same numbers are produced on the device and only scalars / the k selected indices cross PCIe:

* ``LazyOutputs`` -- what the patched ``outputs_losses`` returns as "outputs" when the net is frozen (the ``_test()`` path):
  it snapshots the flat parameter vector (160 KB, device to device) instead of running the full forward, answers the
  ``.detach().cpu().numpy()`` chain of ``deepxde.utils.to_numpy`` with an array-like that computes the outputs -- at the
  weights of the evaluation, not the later live ones -- only if somebody actually reads them (metrics, plots);
* ``topk_residual`` / ``fused_rar_wrapper`` -- the residual-adaptive refinement loop of src/utils/rar.py with the residual
  query served by the fused forward-only kernel, |r| summed over the PDE terms, its mean and its top-k taken on the device;
* ``DeviceTester`` -- TesterCallback's error metrics (same formulas, same errors.txt columns) reduced on the device.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

D2H_BYTES = {"outputs": 0, "metrics": 0, "topk": 0}      # bytes this module moved device -> host (tests read it)


def functional_forward(flat: torch.Tensor, shape, x: torch.Tensor, chunk: int = 1 << 20) -> torch.Tensor:
    """tanh-MLP outputs [N, m] from the flat ``nn.Linear``-order parameter vector (deepxde/nn/pytorch/fnn.py:35-48), in
    row chunks so that a 16 Mi-row evaluation keeps its temporaries at chunk x H floats."""
    d, H, L, m = shape
    ws, off = [], 0
    dims = [d] + [H] * L + [m]
    for a, b in zip(dims[:-1], dims[1:]):
        ws.append((flat[off:off + a * b].view(b, a), flat[off + a * b:off + a * b + b]))
        off += a * b + b
    out = torch.empty((x.shape[0], m), dtype=torch.float32, device=x.device)
    for beg in range(0, x.shape[0], chunk):
        h = x[beg:beg + chunk]
        for i, (w, b) in enumerate(ws):
            h = torch.addmm(b, h, w.t())
            if i < len(ws) - 1:
                h = torch.tanh(h)
        out[beg:beg + chunk] = h
    return out


class _LazyArray:
    """numpy side of ``LazyOutputs``: materialises on first use (``np.asarray``, arithmetic with an ndarray, indexing)."""

    def __init__(self, owner: "LazyOutputs"):
        self._owner = owner
        self._val: Optional[np.ndarray] = None
        self.shape = owner.shape
        self.ndim = 2
        self.dtype = np.dtype(np.float32)

    def _get(self) -> np.ndarray:
        if self._val is None:
            y = self._owner.materialize()
            D2H_BYTES["outputs"] += y.numel() * 4
            self._val = y.cpu().numpy()
        return self._val

    def __array__(self, dtype=None, copy=None):
        v = self._get()
        return v if dtype is None else v.astype(dtype, copy=False)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        return self._get()[idx]

    @property
    def materialized(self) -> bool:
        return self._val is not None

    def __getattr__(self, name):
        # anything else an ndarray offers (.T, .astype, .reshape, .mean, ...): compute the outputs and delegate
        if name.startswith("__") or name in ("_owner", "_val"):
            raise AttributeError(name)
        return getattr(self._get(), name)

    def __iter__(self):
        return iter(self._get())


class LazyOutputs:
    """Network outputs for every row of ``inputs`` at the CURRENT weights, computed only on demand.

    Quacks like the tensor ``deepxde.utils.to_numpy`` expects (``.detach().cpu().numpy()``)."""

    def __init__(self, flat_snapshot: torch.Tensor, shape, inputs, device):
        self._flat = flat_snapshot
        self._shape = shape
        self._inputs = inputs
        self._device = device
        n = inputs.shape[0]
        self.shape = (int(n), int(shape[3]))
        self._arr: Optional[_LazyArray] = None

    def materialize(self) -> torch.Tensor:
        x = torch.as_tensor(self._inputs, dtype=torch.float32).to(self._device)
        with torch.no_grad():
            return functional_forward(self._flat, self._shape, x)

    def detach(self):
        return self

    def cpu(self):
        return self

    def numpy(self) -> _LazyArray:
        if self._arr is None:
            self._arr = _LazyArray(self)
        return self._arr


@torch.no_grad()
def topk_residual(op, params: Sequence[torch.Tensor], k: int):
    """(indices [k] ascending by error as ``np.argsort(err)[-k:]`` returns them, mean error) of err = sum_terms |r| over the
    rows staged in the fused op -- src/utils/rar.py:21-30 with every reduction on the device."""
    res = op.residual(params)                                  # [N, n_pde], device
    err = res.abs().sum(dim=1)
    k = min(int(k), err.shape[0])
    vals, idx = torch.topk(err, k, largest=True, sorted=True)
    idx = torch.flip(idx, dims=[0])                            # ascending, the largest last
    out = torch.cat([idx.to(torch.float64), err.mean(dtype=torch.float64).reshape(1)]).cpu().numpy()
    D2H_BYTES["topk"] += out.nbytes
    return out[:-1].astype(np.int64), float(out[-1])


def fused_rar_wrapper(pde, model, conf):
    """``src.utils.rar.rar_wrapper`` (rar.py:4-34) for a model patched by ``enable_fused``: same schedule, same anchors, the
    residual of the whole training set reduced to its ``count`` worst rows on the device."""
    data = model.data
    train = model.train

    def wrapper(*args, **kwargs):
        total_iter = kwargs["iterations"]
        interval, count = conf["interval"], conf["count"]
        assert total_iter % interval == 0
        kwargs["iterations"] = interval
        for i in range(total_iter // interval):
            if i == 0:
                train(*args, **kwargs)
                continue
            X = model.train_state.X_train
            idx, mean_err = model.fused_topk_residual(X, count)
            print(f"mean residual: {mean_err}")
            data.add_anchors(X[idx])
            train(*args, **kwargs, disregard_previous_best=True, save_model=False)

    return wrapper


class DeviceTester:
    """TesterCallback's error metrics (src/utils/callbacks.py:161-176: mse, mae, mxe, l1re, l2re, crmse against the solution
    table or ``ref_sol`` samples) computed on the device: the test set is uploaded once, each evaluation returns six scalars.

    Usable as a deepxde callback (duck-typed: the hooks ``CallbackList`` calls) or directly through ``evaluate()``."""

    def __init__(self, test_x=None, test_y=None, log_every: Optional[int] = 100, verbose: bool = False):
        self.test_x, self.test_y = test_x, test_y
        self.log_every, self.verbose = log_every, verbose
        self.model = None
        self.indexes, self.rows = [], []
        self.valid_epoch = 0
        self._since = 0
        self._dev = None

    # ---- deepxde.callbacks.Callback protocol -------------------------------------------------------------------
    def set_model(self, model):
        self.model = model

    def init(self):
        pass

    def on_epoch_begin(self): pass
    def on_batch_begin(self): pass
    def on_batch_end(self): pass
    def on_predict_begin(self): pass
    def on_predict_end(self): pass

    def on_train_begin(self):
        if self.test_x is None:
            pde = self.model.pde
            if pde.ref_sol is not None:                       # callbacks.py:134-146
                n = 2500 if pde.input_dim == 2 else 20000
                fn = getattr(self.model.data.geom, "uniform_points", None) or self.model.data.geom.random_points
                self.test_x = fn(n, boundary=False)
                self.test_y = pde.ref_sol(self.test_x)
            elif pde.ref_data is not None:
                mask = np.isnan(pde.ref_data).any(axis=1)
                self.test_x = pde.ref_data[~mask, :pde.input_dim]
                self.test_y = pde.ref_data[~mask, pde.input_dim:]
        self._stage(next(self.model.net.parameters()).device)

    def on_epoch_end(self):
        self._since += 1
        self.valid_epoch += 1
        if self.test_x is None or self.log_every is None or self._since < self.log_every:
            return
        self._since = 0
        self.indexes.append(self.valid_epoch)
        self.rows.append(self.evaluate(self.model.net))

    def on_train_end(self):
        path = getattr(self.model, "model_save_path", None)
        if path and self.rows:
            r = np.asarray(self.rows)
            nan = np.full(len(r), np.nan)
            # errors.txt columns as callbacks.py:200-206 writes them
            np.savetxt(path + "/errors.txt", np.array([self.indexes, r[:, 1], r[:, 0], r[:, 2], r[:, 3], r[:, 4], r[:, 5], nan, nan, nan]).T,
                       header="epochs, maes, mses, mxes, l1res, l2res, crmses, frmses(low, mid, high)")

    # ---- the metrics -------------------------------------------------------------------------------------------
    def _stage(self, device):
        self._dev = device
        self._x = torch.as_tensor(np.asarray(self.test_x), dtype=torch.float32).to(device)
        y = torch.as_tensor(np.asarray(self.test_y)).to(device).to(torch.float64)
        self._y = y
        self._sol_l1 = y.abs().mean()
        self._sol_l2 = (y ** 2).mean().sqrt()

    @torch.no_grad()
    def evaluate(self, net) -> np.ndarray:
        """[mse, mae, mxe, l1re, l2re, crmse] at the net's current weights; 48 bytes device -> host."""
        if self._dev is None:
            self._stage(next(net.parameters()).device)
        was = net.training
        net.eval()
        y = net(self._x).to(torch.float64)
        net.train(was)
        diff = y - self._y
        mse = (diff ** 2).mean()
        mae = diff.abs().mean()
        out = torch.stack([mse, mae, diff.abs().max(), mae / self._sol_l1, mse.sqrt() / self._sol_l2, diff.mean().abs()]).cpu().numpy()
        D2H_BYTES["metrics"] += out.nbytes
        if self.verbose:
            print("Validation: epoch {} MSE {:.5f} MAE {:.5f} MXE {:.5f} L1RE {:.5f} L2RE {:.5f} CRMSE {:.5f}".format(self.valid_epoch, *out))
        return out

    @property
    def l2res(self):
        return [float(r[4]) for r in self.rows]
