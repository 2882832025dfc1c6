"""Import the UNMODIFIED reference in-process: for oracle pinning (CPU, build container) and as the comparator of the
drop-in tests on the GPU box.

Root = /root/reference where it is mounted (the build container), else the byte-for-byte copy that
``scripts/install_reference.py`` placed under ``baseline/_ref`` (git-ignored; it travels to the GPU box).  Two optional
dependencies of the reference are absent from this image and cannot be installed, so they
are stubbed before ``import deepxde``:

* ``matplotlib`` -> MagicMock modules (plotting only);
* ``skopt``      -> a stand-in quasi-random sampler built on scipy.stats.qmc (affects which
  training points ``dde.data.PDE`` draws, not the residual/gradient arithmetic).

Nothing in ``pinnacle_b200/`` imports this module; ``smoke()`` does not either.  GPU tests and ``bench.py --impl
reference`` use it through ``baseline/_ref`` only (never /root/reference, which does not exist on the GPU box).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from unittest import mock

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ["/root/reference", os.path.join(_REPO, "baseline", "_ref")]       # (tests import this module as `refshim`)
if os.environ.get("PINN_REFERENCE_ROOT"):
    _CANDIDATES.insert(0, os.environ["PINN_REFERENCE_ROOT"])
REFERENCE_ROOT = next((c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "deepxde"))), _CANDIDATES[0])


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deepxde"))


def _install_stubs():
    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.tri", "matplotlib.colors", "matplotlib.cm",
              "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1"]:
        sys.modules.setdefault(n, mock.MagicMock(name=n))
    if "skopt" not in sys.modules:
        from scipy.stats import qmc

        class _QMC:
            def __init__(self, *a, **k):
                pass

            def generate(self, space, n):
                d = len(space)
                return qmc.Halton(d, scramble=False).random(n + 1)[1:].tolist()

        skopt = types.ModuleType("skopt")
        sampler = types.ModuleType("skopt.sampler")
        for nm in ("Lhs", "Halton", "Hammersly", "Sobol", "Grid"):
            setattr(sampler, nm, _QMC)
        skopt.sampler = sampler
        space = types.ModuleType("skopt.space")

        class Space(list):
            pass

        space.Space = Space
        skopt.space = space
        skopt.Space = Space
        sys.modules["skopt"] = skopt
        sys.modules["skopt.sampler"] = sampler
        sys.modules["skopt.space"] = space


_loaded = None


def load():
    """Returns (dde, src_pde_modules dict).  Idempotent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not mounted")
    _install_stubs()
    os.environ["DDEBACKEND"] = "pytorch"
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    import torch
    import deepxde as dde
    torch.set_default_dtype(torch.float32)
    dde.config.set_default_float("float32")
    from src.pde import baseclass, burgers, chaotic, heat, helmholtz, inverse, ns, poisson, wave
    _loaded = (dde, dict(baseclass=baseclass, burgers=burgers, chaotic=chaotic, heat=heat, helmholtz=helmholtz, inverse=inverse, ns=ns,
                         poisson=poisson, wave=wave))
    return _loaded


@contextlib.contextmanager
def _in_reference_dir():
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


def _synthetic_table(path: str) -> np.ndarray:
    """Deterministic stand-ins for coefficient / initial-condition tables whose data files were
    stripped from the reference tree (.MISSING_LARGE_BLOBS)."""
    rng = np.random.default_rng(abs(hash(os.path.basename(path))) % (2 ** 32))
    g = np.linspace(0, 1, 24)
    xx, yy = np.meshgrid(g, g, indexing="ij")
    if "burgers2d_init" in path:
        xx, yy = xx * 4, yy * 4
        val = np.sin(xx) * np.cos(yy)
    else:
        val = 1.0 + 0.5 * np.sin(3 * xx) * np.cos(2 * yy) + 0.05 * rng.standard_normal(xx.shape)
    return np.stack([xx.ravel(), yy.ravel(), val.ravel()], axis=1)


CLASS_MODULE = {
    "Burgers1D": "burgers", "Burgers2D": "burgers",
    "Poisson1D": "poisson", "Poisson2D_Classic": "poisson", "PoissonBoltzmann2D": "poisson",
    "Poisson3D_ComplexGeometry": "poisson", "Poisson2D_ManyArea": "poisson", "PoissonND": "poisson",
    "Heat2D_VaryingCoef": "heat", "Heat2D_Multiscale": "heat", "Heat2D_ComplexGeometry": "heat",
    "Heat2D_LongTime": "heat", "HeatND": "heat",
    "NS2D_Classic": "ns", "NS2D_LidDriven": "ns", "NS2D_BackStep": "ns", "NS2D_LongTime": "ns",
    "Wave1D": "wave", "Wave2D_Heterogeneous": "wave", "Wave2D_LongTime": "wave",
    "GrayScottEquation": "chaotic", "KuramotoSivashinskyEquation": "chaotic",
    "Helmholtz2D": "helmholtz",
    "PoissonInv": "inverse", "HeatInv": "inverse",
}


VARIANTS = {"NS2D_Classic_linear": ("NS2D_Classic", {"linear": True})}


def construct(name: str, **kwargs):
    """Instantiate reference problem class ``name``; solution-data files that were stripped
    from the tree are bypassed (they are not on the residual/gradient path)."""
    dde, mods = load()
    if name in VARIANTS:                      # a named constructor variant of a class (golden fixtures are keyed by name)
        name, extra = VARIANTS[name]
        kwargs = {**extra, **kwargs}
    cls = getattr(mods[CLASS_MODULE[name]], name)
    base = mods["baseclass"]
    real_loadtxt = np.loadtxt
    real_load_ref = base.BasePDE.load_ref_data

    def load_ref(self, datapath, *a, **k):
        if os.path.exists(datapath):
            return real_load_ref(self, datapath, *a, **k)
        self.ref_data = None

    def loadtxt(path, *a, **k):
        if isinstance(path, str) and not os.path.exists(path):
            return _synthetic_table(path)
        return real_loadtxt(path, *a, **k)

    with _in_reference_dir(), mock.patch.object(base.BasePDE, "load_ref_data", load_ref), \
            mock.patch.object(np, "loadtxt", loadtxt):
        return cls(**kwargs)


def reference_losses_and_grads(pde_obj, net, x_np, n_bc_rows: int = 0, seeds=None):
    """Run the reference pipeline on explicit points: net forward, ``pde_obj.pde`` (dde.grad),
    drop the first ``n_bc_rows`` rows, mean-square per term (data/pde.py:147-152), backward per seed.

    Returns (losses[n_pde], grads[S, P] flat in net.parameters() order, residual matrix)."""
    import torch
    dde, _ = load()
    x = torch.as_tensor(x_np).clone().requires_grad_(True)
    dde.grad.clear()
    u = net(x)
    f = pde_obj.pde(x, u)
    if not isinstance(f, (list, tuple)):
        f = [f]
    f = [fi[n_bc_rows:] for fi in f]
    losses = torch.stack([dde.losses.mean_squared_error(torch.zeros_like(e), e) for e in f])
    params = list(net.parameters())
    n_pde = len(f)
    if seeds is None:
        seeds = np.eye(n_pde)
    grads = []
    for s in np.asarray(seeds):
        net.zero_grad()
        tot = (losses * torch.as_tensor(s, dtype=losses.dtype)).sum()
        tot.backward(retain_graph=True)
        grads.append(torch.cat([(torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1) for p in params]).clone())
    dde.grad.clear()
    res = torch.cat([fi.reshape(-1, 1) for fi in f], dim=1).detach()
    return losses.detach(), torch.stack(grads), res
