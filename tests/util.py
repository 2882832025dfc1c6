"""Shared test helpers (golden fixtures, error metrics, the oracle-backed test double)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from golden import common
from pinnacle_b200 import kinds as K

GOLDEN_NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(common.GOLDEN_DIR, "*.npz"))
                      if not os.path.basename(f).startswith(("l2re_", "gepinn_", "regions")))   # l2re_*: training runs (test_cuda_l2re.py); gepinn_*: test_gpinn*.py; regions: test_resample.py

# Parity tolerances (BASELINE.json north_star: residuals, losses and gradients within 1e-5 relative in fp32)
RTOL = 1e-5


class Golden:
    def __init__(self, name):
        g = np.load(common.fixture_path(name))
        self.name = name
        self.g = g
        self.desc = K.KindDesc(int(g["kind"]), int(g["d"]), int(g["m"]), tuple(int(k) for k in g["order"]), int(g["n_pde"]),
                               int(g["n_aux"]), [float(c) for c in g["consts"]], [int(a) for a in g["aux_sel"]], name=name)
        self.seed = int(g["seed"])
        self.flat64 = common.golden_params(self.desc.layer_shapes(), self.seed)
        self.idx = g["idx"]
        self.probes = common.probe_vectors(self.desc.n_params, self.seed)

    def inputs(self, dtype):
        nd = np.float32 if dtype == torch.float32 else np.float64
        tag = "32" if dtype == torch.float32 else "64"
        flat = torch.as_tensor(self.flat64.astype(nd))
        x = torch.as_tensor(self.g["x"].astype(nd))
        aux = torch.as_tensor(self.g[f"aux{tag}"].astype(nd)) if self.desc.n_aux else None
        return flat, x, aux


class GpinnGolden:
    """tests/golden/gepinn_<Class>.npz (make_gpinn_golden.py): the reference with use_gepinn() on the REAL two-input network.
    ``embedded(dtype)`` is what the kernels are handed -- first layer times the direction matrix, rows padded with zeros --
    and ``back(g)`` maps a gradient with respect to the embedded parameters to the real ones (chain rule through W1 @ V)."""

    def __init__(self, name):
        self.name = name
        self.g = g = np.load(common.fixture_path("gepinn_" + name))
        self.embed = np.asarray(g["embed"], dtype=np.float64)
        # the descriptor as the generator recorded it (problems.gpinn of the live reference object); the fixture carries the aux
        # columns of its points, so no coefficient tables are needed here
        self.desc = K.KindDesc(int(g["kind"]), self.embed.shape[1], 1, (3,) * self.embed.shape[1], int(g["loss64"].shape[0]), int(g["n_aux"]),
                               [float(c) for c in g["consts"]], [int(a) for a in g["aux_sel"]], name=name + "_gepinn")
        self.base_d = int(g["base_d"])
        self.base = K.KindDesc(K.KIND_LINEAR, self.base_d, 1, (2,) * self.base_d, name=name)   # only the layer shapes of the real net matter
        self.seed = int(g["seed"])
        self.flat64 = common.golden_params(self.base.layer_shapes(), self.seed)
        self.idx = g["idx"]
        self.probes = common.probe_vectors(self.base.n_params, self.seed)

    def embedded(self, dtype):
        nd = np.float32 if dtype == torch.float32 else np.float64
        H = self.base.H
        flat = self.flat64.astype(nd)
        bd = self.base_d
        w1 = flat[:bd * H].reshape(H, bd)
        flat_e = np.concatenate([(w1 @ self.embed.astype(nd)).reshape(-1), flat[bd * H:]]).astype(nd)
        x = self.g["x"].astype(nd)
        x_e = np.concatenate([x, np.zeros((x.shape[0], self.desc.d - bd), nd)], axis=1)
        return torch.as_tensor(flat_e), torch.as_tensor(np.ascontiguousarray(x_e))

    def aux(self, x_e):
        """Aux columns of the fixture's (embedded) rows -- source and its input gradient, or table coefficients -- or None."""
        if not self.desc.n_aux:
            return None
        return torch.as_tensor(np.ascontiguousarray(self.g["aux64"])).to(x_e.dtype)

    def back(self, grad_e):
        grad_e = np.asarray(grad_e, dtype=np.float64)
        H, de = self.base.H, self.desc.d
        w1 = grad_e[..., :de * H].reshape(grad_e.shape[:-1] + (H, de)) @ self.embed.T
        return np.concatenate([w1.reshape(grad_e.shape[:-1] + (self.base_d * H,)), grad_e[..., de * H:]], axis=-1)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def assert_parity(got, ref, what, rtol=RTOL):
    """||got-ref||_2 <= rtol ||ref||_2  and  max|got-ref| <= rtol max|ref|  (SURVEY.md §8c recipe)."""
    e2, em = rel_l2(got, ref), rel_max(got, ref)
    assert e2 <= rtol and em <= rtol, f"{what}: rel-L2 {e2:.3e}, rel-max {em:.3e} exceed {rtol:g}"


# ---------------------------------------------------------------------------------------------
# oracle-backed stand-in for the C ABI, used ONLY by CPU tests of host-side plumbing
# (sharding, all-reduce, autograd semantics, the dde.Model patch).  Never used by the product.
# ---------------------------------------------------------------------------------------------
def oracle_fwd(desc, params, x, aux, inv_n_global, want_residual=False):
    from oracle import autograd_ref as O
    with torch.enable_grad():          # called from inside autograd.Function.forward (grad mode off there)
        losses, res, _ = O.pde_losses(desc, params.detach(), x, aux, inv_n=inv_n_global)
    return losses.detach().to(torch.float32), (res.detach() if want_residual else None)


def oracle_fwd_bwd(desc, params, x, aux, inv_n_global, seeds, out=None):
    from oracle import autograd_ref as O
    with torch.enable_grad():
        losses, grads, _ = O.losses_and_grads(desc, params.detach(), x, aux, seeds=seeds, inv_n=inv_n_global)
    res = torch.cat([grads.reshape(-1), losses]).to(torch.float32)
    if out is not None:
        out.copy_(res)
        return out
    return res


def oracle_fwd_bwd_linear(desc, params, x, aux, seeds, out=None):
    from oracle import autograd_ref as O
    with torch.enable_grad():
        rows, sq = O.linear_rows(desc, params.detach(), x, aux, seeds)
    res = torch.cat([rows.reshape(-1), sq]).to(torch.float32)
    if out is not None:
        out.copy_(res)
        return out
    return res


def region_holes(region_c):
    """(shape, a, b) tuples of a ctypes pinn_region, as oracle/sample_ref.py takes them."""
    out = []
    for i in range(region_c.n_holes):
        h = region_c.holes[i]
        out.append((int(h.shape), [h.a[j] for j in range(h.n_dims)], [h.b[j] for j in range(h.n_dims if h.shape == 1 else 1)]))
    return out


def oracle_sample_points(x, region_c, key, draw, first_row=0):
    """Stand-in for capi.sample_points on CPU tensors: the rows the kernel would write, from oracle/sample_ref.py."""
    from oracle import sample_ref as S
    d = region_c.d
    keep = ([region_c.keep_center[j] for j in range(region_c.keep_dims)], region_c.keep_radius) if region_c.keep_dims > 0 else None
    rows = S.sample_points(x.shape[0], [region_c.lo[j] for j in range(d)], [region_c.hi[j] for j in range(d)], region_holes(region_c),
                           key=int(key), draw=int(draw), first_row=int(first_row), keep=keep)
    x[:, :d] = torch.from_numpy(rows)
    return x


class oracle_backend:
    """Context manager swapping pinnacle_b200.fused's backend (and the device point sampler) for the oracle (CPU tests only)."""

    def __enter__(self):
        from pinnacle_b200 import capi, fused
        self._saved = (fused._backend_fwd, fused._backend_fwd_bwd, fused._backend_fwd_bwd_linear, capi.sample_points)
        fused._backend_fwd, fused._backend_fwd_bwd, fused._backend_fwd_bwd_linear = oracle_fwd, oracle_fwd_bwd, oracle_fwd_bwd_linear
        capi.sample_points = oracle_sample_points
        return self

    def __exit__(self, *a):
        from pinnacle_b200 import capi, fused
        fused._backend_fwd, fused._backend_fwd_bwd, fused._backend_fwd_bwd_linear, capi.sample_points = self._saved
