"""Host-side mirror of the PINNacle problem classes (``src/pde/*.py``) for the fused path.

For every residual the reference defines as a ``pde(x, u)`` closure this module builds the
``KindDesc`` the CUDA op needs plus, where the residual contains a point-dependent source
term or coefficient, an ``aux_fn`` that evaluates those terms ONCE per point set (the
collocation points are static between resamples; the reference re-evaluates them every
step, or memoises them through ``src/utils/func_cache.py:5-18``).

Two entry points:

* ``make(name, **kwargs)``  -- build a spec from the class name and the same constructor
  keywords the reference class takes (defaults copied from its signature);
* ``describe_reference(pde_obj)`` -- build a spec from a live reference ``BasePDE`` object
  by matching the class name and reading the constants its closure captured.  Unknown
  classes raise ``UnsupportedProblem`` -- there is no silent fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from . import kinds as K
from .kinds import KindDesc, linear_desc


class UnsupportedProblem(ValueError):
    pass


@dataclass
class ProblemSpec:
    name: str
    desc: KindDesc
    bbox: Sequence[float]                      # [lo0, hi0, lo1, hi1, ...] in input order
    aux_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None   # x[N,d] -> aux[N,n_aux] (same dtype/device)
    # gPINN: direction matrix V [d_net, desc.d].  The kernels see a desc.d-"input" network whose rows are (x, 0, .., 0) and
    # whose first layer is W1 @ V, so that their per-direction jets are directional derivatives along the columns of V.
    embed: Optional[np.ndarray] = None
    # True: the aux columns are table look-ups the reference computes outside autograd (constants when a residual is differentiated)
    aux_constant: bool = False

    @property
    def net_d(self) -> int:
        """Input dimension of the network the user trains (desc.d unless the problem is embedded)."""
        return int(self.embed.shape[0]) if self.embed is not None else self.desc.d

    def aux(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        if self.aux_fn is None:
            return None
        a = self.aux_fn(x)
        assert a.shape == (x.shape[0], self.desc.n_aux), (a.shape, self.desc.n_aux)
        return a.contiguous()

    def sample(self, n: int, seed: int = 0, dtype=np.float32) -> np.ndarray:
        """Uniform synthetic collocation points in the bounding box."""
        rng = np.random.default_rng(seed)
        lo = np.asarray(self.bbox[0::2], dtype=np.float64)
        hi = np.asarray(self.bbox[1::2], dtype=np.float64)
        return rng.uniform(lo, hi, size=(n, self.net_d)).astype(dtype)


def _col(x, j):
    return x[:, j:j + 1]


# ----------------------------------------------------------------------------------------
# builders, one per reference class.  Each cites the residual it mirrors.
# ----------------------------------------------------------------------------------------

def burgers1d(geom=(-1, 1), time=(0, 1), nu=0.01 / np.pi) -> ProblemSpec:
    # src/pde/burgers.py:21-25   u_t + u u_x - nu u_xx
    d = KindDesc(K.KIND_BURGERS1D, 2, 1, (2, 1), consts=[nu], name="Burgers1D")
    return ProblemSpec("Burgers1D", d, list(geom) + list(time))


def burgers2d(nu=0.001, L=4, T=1) -> ProblemSpec:
    # src/pde/burgers.py:66-80
    d = KindDesc(K.KIND_BURGERS2D, 3, 2, (2, 2, 1), n_pde=2, consts=[nu], name="Burgers2D")
    return ProblemSpec("Burgers2D", d, [0, L, 0, L, 0, T])


def poisson1d(a=1) -> ProblemSpec:
    # src/pde/poisson.py:19-24   u_xx + a^2 sin(a x)
    d = linear_desc(1, (2,), c2=(1.0,), n_aux=1, src_col=0, name="Poisson1D")
    return ProblemSpec("Poisson1D", d, [0, 2 * np.pi / a], lambda x: (a ** 2) * torch.sin(a * x))


def poisson2d_classic(scale=1) -> ProblemSpec:
    # src/pde/poisson.py:54-58   u_xx + u_yy
    d = linear_desc(2, (2, 2), c2=(1.0, 1.0), name="Poisson2D_Classic")
    return ProblemSpec("Poisson2D_Classic", d, [-scale / 2, scale / 2, -scale / 2, scale / 2])


def poisson_boltzmann2d(k=8, mu=(1, 4), A=10, bbox=(-1, 1, -1, 1)) -> ProblemSpec:
    # src/pde/poisson.py:111-120   -(u_xx+u_yy) + k^2 u - f(x,y)
    def aux(xy):
        x, y = _col(xy, 0), _col(xy, 1)
        f = A * (mu[0] ** 2 + x ** 2 + mu[1] ** 2 + y ** 2) * torch.sin(mu[0] * torch.pi * x) * torch.sin(mu[1] * torch.pi * y)
        return -f

    d = linear_desc(2, (2, 2), cu=k ** 2, c2=(-1.0, -1.0), n_aux=1, src_col=0, name="PoissonBoltzmann2D")
    return ProblemSpec("PoissonBoltzmann2D", d, list(bbox), aux)


def poisson3d_complexgeometry(bbox=(0, 1, 0, 1, 0, 1), interface_z=0.5, A=(20, 100), m=(1, 10, 5), k=(8, 10), mu=(1, 1)) -> ProblemSpec:
    # src/pde/poisson.py:179-193   -mu(z) lap(u) + k(z)^2 u - f(x,y,z)
    def aux(xyz):
        x, y, z = _col(xyz, 0), _col(xyz, 1), _col(xyz, 2)
        xlen2 = x ** 2 + y ** 2 + z ** 2
        part1 = torch.exp(torch.sin(m[0] * np.pi * x) + torch.sin(m[1] * np.pi * y) + torch.sin(m[2] * np.pi * z)) * (xlen2 - 1) / (xlen2 + 1)
        part2 = torch.sin(m[0] * np.pi * x) * torch.sin(m[1] * np.pi * y) * torch.sin(m[2] * np.pi * z)
        f = A[0] * part1 + A[1] * part2
        below = z < interface_z
        mus = torch.where(below, float(mu[0]), float(mu[1])).to(xyz.dtype)
        ks = torch.where(below, float(k[0] ** 2), float(k[1] ** 2)).to(xyz.dtype)
        return torch.cat([mus, ks, -f], dim=1)

    d = linear_desc(3, (2, 2, 2), cu=1.0, c2=(-1.0, -1.0, -1.0), n_aux=3, src_col=2, cu_col=1, c2_cols=(0, 0, 0),
                    name="Poisson3D_ComplexGeometry")
    return ProblemSpec("Poisson3D_ComplexGeometry", d, list(bbox), aux)


def poisson2d_manyarea(a_cof: np.ndarray, f_cof: np.ndarray, bbox=(-10, 10, -10, 10)) -> ProblemSpec:
    # src/pde/poisson.py:222-258   a(x) lap(u) + f(x); a piecewise constant per block, f a sine series per block
    a_cof = np.asarray(a_cof, dtype=np.float64)
    f_cof = np.asarray(f_cof, dtype=np.float64)
    split = f_cof.shape[:2]
    block = np.array([(bbox[1] - bbox[0] + 2e-5) / split[0], (bbox[3] - bbox[2] + 2e-5) / split[1]])
    lo = np.array(bbox[::2], dtype=np.float64)

    def aux(xy):
        xn = xy.detach().cpu().numpy().astype(np.float64)
        red = xn - lo + 1e-5
        dom = np.floor(red / block).astype(np.int64)
        res = (red - dom * block) / block
        a = a_cof[dom[:, 0], dom[:, 1]]
        coef = f_cof[dom[:, 0], dom[:, 1]]                      # [N, fi, fj]
        f = coef[:, 0, 0].copy()
        for i in range(coef.shape[1]):
            for j in range(coef.shape[2]):
                f += coef[:, i, j] * np.sin(np.pi * i * res[:, 0]) * np.sin(np.pi * j * res[:, 1])
        out = np.stack([a, f], axis=1).astype(np.float32)     # reference stores float32 tensors (poisson.py:251)
        return torch.as_tensor(out).to(dtype=xy.dtype, device=xy.device)

    d = linear_desc(2, (2, 2), c2=(1.0, 1.0), n_aux=2, src_col=1, c2_cols=(0, 0), name="Poisson2D_ManyArea")
    return ProblemSpec("Poisson2D_ManyArea", d, list(bbox), aux, aux_constant=True)


def poisson_nd(dim=5, len=1) -> ProblemSpec:
    # src/pde/poisson.py:288-296   lap(u) + pi^2/4 sum_i sin(pi x_i / 2)
    def aux(x):
        return (np.pi ** 2) / 4 * torch.sin(np.pi / 2 * x).sum(dim=1).reshape(-1, 1)

    d = linear_desc(dim, (2,) * dim, c2=(1.0,) * dim, n_aux=1, src_col=0, name="PoissonND")
    return ProblemSpec("PoissonND", d, [0, len] * dim, aux)


def _nearest_grid_lookup(table: np.ndarray):
    """Nearest-neighbour lookup on scattered (x, y, value) rows, as
    scipy.interpolate.griddata(method='nearest') does (heat.py:28-32)."""
    from scipy.spatial import cKDTree
    tree = cKDTree(table[:, 0:2])

    def look(q: np.ndarray) -> np.ndarray:
        _, idx = tree.query(q)
        return table[idx, 2]

    return look


def heat2d_varyingcoef(heat_2d_coef: np.ndarray, bbox=(0, 1, 0, 1, 0, 5), A=200, m=(1, 5, 1)) -> ProblemSpec:
    # src/pde/heat.py:34-41   u_t - c(x,y) (u_xx+u_yy) - f(x,y,t)
    look = _nearest_grid_lookup(np.asarray(heat_2d_coef, dtype=np.float64))

    def aux(x):
        c = torch.as_tensor(look(x.detach().cpu().numpy()[:, 0:2]).astype(np.float32)).to(dtype=x.dtype, device=x.device).unsqueeze(1)
        f = A * torch.sin(m[0] * torch.pi * _col(x, 0)) * torch.sin(m[1] * torch.pi * _col(x, 1)) * torch.sin(m[2] * torch.pi * _col(x, 2))
        return torch.cat([c, -f], dim=1)

    d = linear_desc(3, (2, 2, 1), c1=(0, 0, 1.0), c2=(-1.0, -1.0, 0.0), n_aux=2, src_col=1, c2_cols=(0, 0, -1),
                    name="Heat2D_VaryingCoef")
    return ProblemSpec("Heat2D_VaryingCoef", d, list(bbox), aux)


def heat2d_multiscale(bbox=(0, 1, 0, 1, 0, 5), pde_coef=(1 / np.square(500 * np.pi), 1 / np.square(np.pi))) -> ProblemSpec:
    # src/pde/heat.py:88-93   u_t - c0 u_xx - c1 u_yy
    d = linear_desc(3, (2, 2, 1), c1=(0, 0, 1.0), c2=(-pde_coef[0], -pde_coef[1], 0.0), name="Heat2D_Multiscale")
    return ProblemSpec("Heat2D_Multiscale", d, list(bbox))


def heat2d_complexgeometry(bbox=(-8, 8, -12, 12, 0, 3)) -> ProblemSpec:
    # src/pde/heat.py:147-152   u_t - u_xx - u_yy
    d = linear_desc(3, (2, 2, 1), c1=(0, 0, 1.0), c2=(-1.0, -1.0, 0.0), name="Heat2D_ComplexGeometry")
    return ProblemSpec("Heat2D_ComplexGeometry", d, list(bbox))


def heat2d_longtime(bbox=(0, 1, 0, 1, 0, 100), k=1, m1=4, m2=2) -> ProblemSpec:
    # src/pde/heat.py:214-222   u_t - 0.001 lap(u) - 5 sin(k u^2) (1 + 2 sin(pi t/4)) sin(m1 pi x) sin(m2 pi y)
    def aux(x):
        return (1 + 2 * torch.sin(_col(x, 2) * np.pi / 4)) * torch.sin(m1 * np.pi * _col(x, 0)) * torch.sin(m2 * np.pi * _col(x, 1))

    d = KindDesc(K.KIND_HEAT_NLSRC, 3, 1, (2, 2, 1), n_aux=1, consts=[0.001, 5.0, k], aux_sel=[0], name="Heat2D_LongTime")
    return ProblemSpec("Heat2D_LongTime", d, list(bbox), aux)


def heat_nd(dim=5, T=1) -> ProblemSpec:
    # src/pde/heat.py:265-275   lap_x(u)/dim + f(x,t) - u_t ,  f = -(1/dim) |x|^2 exp(|x|^2/2 + t)
    def aux(xt):
        x2 = (xt[:, :-1] ** 2).sum(dim=1).reshape(-1, 1)
        t = xt[:, -1:]
        return -1 / dim * x2 * torch.exp(x2 / 2 + t)

    d = linear_desc(dim + 1, (2,) * dim + (1,), c1=(0.0,) * dim + (-1.0,), c2=(1.0 / dim,) * dim + (0.0,), n_aux=1,
                    src_col=0, name="HeatND")
    return ProblemSpec("HeatND", d, [-1, 1] * dim + [0, T], aux)


def _ns2d(name, nu, bbox, linear=False) -> ProblemSpec:
    # src/pde/ns.py:30-49 / 141-160 / 240-259   momentum_x, momentum_y, continuity;  linear: ns.py:51-67 (no convective terms)
    d = KindDesc(K.KIND_NS2D, 2, 3, (2, 2), n_pde=3, consts=[nu, 1.0 if linear else 0.0], name=name)
    return ProblemSpec(name, d, list(bbox))


def ns2d_classic(nu=1, bbox=(0, 8, 0, 8), linear=False) -> ProblemSpec:
    return _ns2d("NS2D_Classic", nu, bbox, linear)


def ns2d_liddriven(nu=1 / 100, bbox=(0, 1, 0, 1)) -> ProblemSpec:
    return _ns2d("NS2D_LidDriven", nu, bbox)


def ns2d_backstep(nu=1 / 100, bbox=(0, 4, 0, 2)) -> ProblemSpec:
    return _ns2d("NS2D_BackStep", nu, bbox)


def ns2d_longtime(nu=1 / 100, bbox=(0, 2, 0, 1, 0, 5)) -> ProblemSpec:
    # src/pde/ns.py:327-352   + u_t, v_t, forcing -initial_fy(x) on momentum_y
    def aux(x):
        return torch.sin(np.pi * _col(x, 0)) * torch.sin(np.pi * _col(x, 1)) * torch.sin(np.pi * _col(x, 2))

    d = KindDesc(K.KIND_NS2D_UNSTEADY, 3, 3, (2, 2, 1), n_pde=3, n_aux=1, consts=[nu], aux_sel=[0], name="NS2D_LongTime")
    return ProblemSpec("NS2D_LongTime", d, list(bbox), aux)


def wave1d(C=2, scale=1) -> ProblemSpec:
    # src/pde/wave.py:22-26   u_tt - C^2 u_xx
    d = linear_desc(2, (2, 2), c2=(-float(C ** 2), 1.0), name="Wave1D")
    return ProblemSpec("Wave1D", d, [0, scale, 0, scale])


def wave2d_heterogeneous(darcy_2d_coef: np.ndarray, bbox=(-1, 1, -1, 1, 0, 5)) -> ProblemSpec:
    # src/pde/wave.py:79-89   u_xx + u_yy - u_tt / c(x,y) ; c by linear griddata interpolation on ((x,y)+1)/2
    from scipy import interpolate
    table = np.asarray(darcy_2d_coef, dtype=np.float64)

    def aux(x):
        q = (x.detach().cpu().numpy()[:, 0:2] + 1) / 2
        c = interpolate.griddata(table[:, 0:2], table[:, 2], q)
        c = torch.as_tensor(c.astype(np.float32)).to(dtype=x.dtype, device=x.device).unsqueeze(1)
        return 1.0 / c

    d = linear_desc(3, (2, 2, 2), c2=(1.0, 1.0, -1.0), n_aux=1, c2_cols=(-1, -1, 0), name="Wave2D_Heterogeneous")
    return ProblemSpec("Wave2D_Heterogeneous", d, list(bbox), aux)


def wave2d_longtime(bbox=(0, 1, 0, 1, 0, 100), a=np.sqrt(2)) -> ProblemSpec:
    # src/pde/wave.py:144-149   u_tt - (u_xx + a^2 u_yy)
    d = linear_desc(3, (2, 2, 2), c2=(-1.0, -float(a * a), 1.0), name="Wave2D_LongTime")
    return ProblemSpec("Wave2D_LongTime", d, list(bbox))


def grayscott(bbox=(-1, 1, -1, 1, 0, 200), b=0.04, d=0.1, epsilon=(1e-5, 5e-6)) -> ProblemSpec:
    # src/pde/chaotic.py:21-31
    dd = KindDesc(K.KIND_GRAYSCOTT, 3, 2, (2, 2, 1), n_pde=2, consts=[epsilon[0], epsilon[1], b, d], name="GrayScottEquation")
    return ProblemSpec("GrayScottEquation", dd, list(bbox))


def kuramoto_sivashinsky(bbox=(0, 2 * np.pi, 0, 1), alpha=100 / 16, beta=100 / (16 * 16), gamma=100 / (16 ** 4)) -> ProblemSpec:
    # src/pde/chaotic.py:77-83   u_t + alpha u u_x + beta u_xx + gamma u_xxxx
    d = KindDesc(K.KIND_KS, 2, 1, (4, 1), consts=[alpha, beta, gamma], name="KuramotoSivashinskyEquation")
    return ProblemSpec("KuramotoSivashinskyEquation", d, list(bbox))


def helmholtz2d(scale=1, A=(4, 4), k=1) -> ProblemSpec:
    # src/pde/helmholtz.py:19-27   u_xx + u_yy + k^2 u - f(x)
    def aux(x):
        f = torch.sin(A[0] * np.pi * _col(x, 0) / scale) * torch.sin(A[1] * np.pi * _col(x, 1) / scale) * \
            (k ** 2 - np.pi ** 2 * (A[0] ** 2 + A[1] ** 2) / scale ** 2)
        return -f

    d = linear_desc(2, (2, 2), cu=k ** 2, c2=(1.0, 1.0), n_aux=1, src_col=0, name="Helmholtz2D")
    return ProblemSpec("Helmholtz2D", d, [0, scale, 0, scale], aux)


def poisson_inv(bbox=(0, 1, 0, 1)) -> ProblemSpec:
    # src/pde/inverse.py:19-24   (a u_x)_x + (a u_y)_y + f_src(x),  outputs (u, a);  f_src: inverse.py:69-79
    def a_ref(x, y):
        return 1 / (1 + x ** 2 + y ** 2 + (x - 1) ** 2 + (y - 1) ** 2)

    def aux(xy):
        x, y = _col(xy, 0), _col(xy, 1)
        return (2 * np.pi ** 2 * torch.sin(np.pi * x) * torch.sin(np.pi * y) * a_ref(x, y)
                + 2 * np.pi * ((2 * x + 1) * torch.cos(np.pi * x) * torch.sin(np.pi * y)
                               + (2 * y + 1) * torch.sin(np.pi * x) * torch.cos(np.pi * y)) * a_ref(x, y) ** 2)

    d = KindDesc(K.KIND_DIV_AGRAD, 2, 2, (2, 2), n_pde=1, n_aux=1, consts=[0.0, 1.0, 1.0], aux_sel=[0], name="PoissonInv")
    return ProblemSpec("PoissonInv", d, list(bbox), aux)


def heat_inv(bbox=(-1, 1, -1, 1, 0, 1)) -> ProblemSpec:
    # src/pde/inverse.py:129-136   u_t - [(a u_x)_x + (a u_y)_y] - f(x),  outputs (u, a);  f: inverse.py:101-113
    def aux(xyt):
        x, y, t = _col(xyt, 0), _col(xyt, 1), _col(xyt, 2)
        s, c, p = torch.sin, torch.cos, np.pi
        return torch.exp(-t) * ((4 * p ** 2 - 1) * s(p * x) * s(p * y)
                                + p ** 2 * (2 * s(p * x) ** 2 * s(p * y) ** 2 - c(p * x) ** 2 * s(p * y) ** 2 - s(p * x) ** 2 * c(p * y) ** 2))

    d = KindDesc(K.KIND_DIV_AGRAD, 3, 2, (2, 2, 1), n_pde=1, n_aux=1, consts=[1.0, -1.0, -1.0], aux_sel=[0], name="HeatInv")
    return ProblemSpec("HeatInv", d, list(bbox), aux)


BUILDERS: Dict[str, Callable[..., ProblemSpec]] = {
    "Burgers1D": burgers1d,
    "Burgers2D": burgers2d,
    "Poisson1D": poisson1d,
    "Poisson2D_Classic": poisson2d_classic,
    "PoissonBoltzmann2D": poisson_boltzmann2d,
    "Poisson3D_ComplexGeometry": poisson3d_complexgeometry,
    "Poisson2D_ManyArea": poisson2d_manyarea,
    "PoissonND": poisson_nd,
    "Heat2D_VaryingCoef": heat2d_varyingcoef,
    "Heat2D_Multiscale": heat2d_multiscale,
    "Heat2D_ComplexGeometry": heat2d_complexgeometry,
    "Heat2D_LongTime": heat2d_longtime,
    "HeatND": heat_nd,
    "NS2D_Classic": ns2d_classic,
    "NS2D_LidDriven": ns2d_liddriven,
    "NS2D_BackStep": ns2d_backstep,
    "NS2D_LongTime": ns2d_longtime,
    "Wave1D": wave1d,
    "Wave2D_Heterogeneous": wave2d_heterogeneous,
    "Wave2D_LongTime": wave2d_longtime,
    "GrayScottEquation": grayscott,
    "KuramotoSivashinskyEquation": kuramoto_sivashinsky,
    "Helmholtz2D": helmholtz2d,
    "PoissonInv": poisson_inv,
    "HeatInv": heat_inv,
}


def make(name: str, **kwargs) -> ProblemSpec:
    try:
        builder = BUILDERS[name]
    except KeyError:
        raise UnsupportedProblem(f"no fused residual kind for problem {name!r}; supported: {sorted(BUILDERS)}") from None
    return builder(**kwargs)


# ----------------------------------------------------------------------------------------
# live reference objects
# ----------------------------------------------------------------------------------------

GP_DIRECTIONS = np.array([[1.0, 0.0, 1.0, 1.0], [0.0, 1.0, 1.0, -1.0]])     # e0, e1, e0+e1, e0-e1 (include/pinn_b200.h, gPINN kinds)


def _gpinn_1d(spec: ProblemSpec) -> ProblemSpec:
    """One input (Poisson1D): loss terms [r, dr/dx].  The same kind and kernel as the two-input case with the single input
    direction repeated -- V = [e0, 0, e0, e0] makes every "y" partial vanish identically (u_xy = c22 - c02 - c12 = 0, ...) -- and
    n_pde = 2, so the third residual the kernel forms is never seeded nor reported."""
    d = spec.desc
    sel = list(d.aux_sel)
    if sel[K.LIN_SEL_CU] >= 0 or sel[K.LIN_SEL_C2] >= 0:
        raise UnsupportedProblem(f"gPINN: {spec.name} has per-point coefficients (not wired up)")
    consts = [0.0] * K.PINN_MAX_CONSTS
    consts[K.LIN_CU], consts[K.LIN_C1], consts[K.LIN_C2] = d.consts[K.LIN_CU], d.consts[K.LIN_C1], d.consts[K.LIN_C2]
    embed = np.array([[1.0, 0.0, 1.0, 1.0]])
    if sel[K.LIN_SEL_SRC] < 0:
        g = KindDesc(K.KIND_GP_LINEAR2, 4, 1, (3, 3, 3, 3), n_pde=2, consts=consts, name=spec.name + "_gepinn")
        return ProblemSpec(spec.name + "_gepinn", g, list(spec.bbox), None, embed)
    src_col, base_aux = sel[K.LIN_SEL_SRC], spec.aux_fn

    def aux_fn(x):
        with torch.enable_grad():
            xr = x[:, :1].detach().to(torch.float64).clone().requires_grad_(True)
            f = base_aux(xr)[:, src_col:src_col + 1]
            (gf,) = torch.autograd.grad(f.sum(), xr)
        return torch.cat([f.detach(), gf], dim=1).to(dtype=x.dtype, device=x.device)

    g = KindDesc(K.KIND_GP_LINEAR2, 4, 1, (3, 3, 3, 3), n_pde=2, n_aux=2, consts=consts, aux_sel=[0, 1, -1], name=spec.name + "_gepinn")
    return ProblemSpec(spec.name + "_gepinn", g, list(spec.bbox), aux_fn, embed)


def gpinn(spec: ProblemSpec) -> ProblemSpec:
    """The gPINN form of a problem (src/pde/baseclass.py:151-180 ``use_gepinn``; benchmark.py:91-92 ``--method gepinn``): loss
    terms [r, dr/dx_0, dr/dx_1].  Covered: one-output problems in two inputs whose residual is Burgers1D's or a LINEAR one with
    constant coefficients, with or without an additive source (Poisson2D_Classic, Wave1D, Helmholtz2D, PoissonBoltzmann2D), or with
    table-look-up coefficients that are constants to the reference's autograd (Poisson2D_ManyArea): the mixed
    third-order partials come from order-3 jets along four directions.  Anything else raises."""
    d = spec.desc
    if d.d == 1 and d.m == 1 and d.n_pde == 1 and d.kind == K.KIND_LINEAR:
        return _gpinn_1d(spec)
    if d.d != 2 or d.m != 1 or d.n_pde != 1:
        raise UnsupportedProblem(f"gPINN: {spec.name} has d={d.d}, m={d.m}, n_pde={d.n_pde}; the fused gPINN kinds cover one-output, "
                                 "one-residual problems in one or two inputs (three inputs need 9 jet directions of order 3)")
    if d.kind == K.KIND_BURGERS1D:
        g = KindDesc(K.KIND_GP_BURGERS1D, 4, 1, (3, 3, 3, 3), n_pde=3, consts=[d.consts[0]], name=spec.name + "_gepinn")
    elif d.kind == K.KIND_LINEAR:
        sel = list(d.aux_sel)
        c2_sels = {sel[K.LIN_SEL_C2 + j] for j in range(2) if d.consts[K.LIN_C2 + j] != 0.0}
        table_aux = getattr(spec, "aux_constant", False)
        if (sel[K.LIN_SEL_CU] >= 0 or c2_sels != {-1}) and not table_aux:
            raise UnsupportedProblem(f"gPINN: {spec.name} has per-point coefficients that autograd differentiates; the input gradient "
                                     "of its residual needs their derivatives too (not wired up)")
        if table_aux:
            # coefficients AND source come from table look-ups on the host (Poisson2D_ManyArea, poisson.py:228-252): constants to
            # the reference's autograd, so dr/dx_j carries the multipliers and no source gradient
            if len(c2_sels) != 1:
                raise UnsupportedProblem(f"gPINN: {spec.name}: the second-order terms do not share one per-point multiplier")
            g = KindDesc(K.KIND_GP_LINEAR2, 4, 1, (3, 3, 3, 3), n_pde=3, n_aux=d.n_aux, consts=list(d.consts),
                         aux_sel=[sel[K.LIN_SEL_SRC], -1, -1, c2_sels.pop(), sel[K.LIN_SEL_CU]], name=spec.name + "_gepinn")
            base_aux = spec.aux_fn
            return ProblemSpec(spec.name + "_gepinn", g, list(spec.bbox), (lambda x: base_aux(x[:, :2])), GP_DIRECTIONS.copy())
        if sel[K.LIN_SEL_SRC] >= 0:
            # additive per-point source f(x) (Helmholtz2D, PoissonBoltzmann2D): dr/dx_j gains df/dx_j.  The source closures are
            # torch formulas, so their gradient comes from autograd, evaluated in float64 once per point set like the source
            src_col, base_aux = sel[K.LIN_SEL_SRC], spec.aux_fn

            def aux_fn(x):
                with torch.enable_grad():
                    xr = x[:, :2].detach().to(torch.float64).clone().requires_grad_(True)
                    f = base_aux(xr)[:, src_col:src_col + 1]
                    (gf,) = torch.autograd.grad(f.sum(), xr)
                return torch.cat([f.detach(), gf], dim=1).to(dtype=x.dtype, device=x.device)

            g = KindDesc(K.KIND_GP_LINEAR2, 4, 1, (3, 3, 3, 3), n_pde=3, n_aux=3, consts=list(d.consts), aux_sel=[0, 1, 2], name=spec.name + "_gepinn")
            return ProblemSpec(spec.name + "_gepinn", g, list(spec.bbox), aux_fn, GP_DIRECTIONS.copy())
        g = KindDesc(K.KIND_GP_LINEAR2, 4, 1, (3, 3, 3, 3), n_pde=3, consts=list(d.consts), name=spec.name + "_gepinn")
    else:
        raise UnsupportedProblem(f"gPINN: no fused residual-gradient kind for {spec.name} (kind {K.KIND_NAMES.get(d.kind, d.kind)})")
    return ProblemSpec(spec.name + "_gepinn", g, list(spec.bbox), None, GP_DIRECTIONS.copy())


def _closure_vars(fn) -> dict:
    """Names -> values captured by a Python closure."""
    out = {}
    if fn.__closure__:
        for n, c in zip(fn.__code__.co_freevars, fn.__closure__):
            try:
                out[n] = c.cell_contents
            except ValueError:      # empty cell
                pass
    return out


def describe_reference(pde_obj) -> ProblemSpec:
    """Build the fused-path spec for a live reference problem object (``src/pde/baseclass.py:10``).

    Constants are read from the variables the reference ``pde`` closure captured, so a
    problem constructed with non-default arguments (``benchmark.py:36-58``) maps correctly.
    """
    name = type(pde_obj).__name__
    if name not in BUILDERS:
        raise UnsupportedProblem(f"no fused residual kind for reference class {name!r}; supported: {sorted(BUILDERS)}")
    cv = _closure_vars(pde_obj.pde)
    gepinn = bool(getattr(pde_obj, "num_gepinn", 0))
    if gepinn:
        # use_gepinn() replaced pde by pde_wrapper, which closes over the original closure (baseclass.py:152-153)
        if "pde_original" not in cv:
            raise UnsupportedProblem("gPINN: cannot find the wrapped residual closure (baseclass.py:152)")
        cv = _closure_vars(cv["pde_original"])
    bbox = list(pde_obj.bbox)
    if name == "Burgers1D":
        spec = burgers1d(bbox[0:2], bbox[2:4], cv["nu"])
    elif name == "Burgers2D":
        spec = burgers2d(cv["nu"], bbox[1], bbox[5])
    elif name == "Poisson1D":
        spec = poisson1d(_closure_vars(cv["f"])["a"])
    elif name == "Poisson2D_Classic":
        spec = poisson2d_classic(bbox[1] * 2)
    elif name == "PoissonBoltzmann2D":
        spec = poisson_boltzmann2d(cv["k"], cv["mu"], cv["A"], bbox)
    elif name == "Poisson3D_ComplexGeometry":
        spec = poisson3d_complexgeometry(bbox, cv["interface_z"], cv["A"], cv["m"], cv["k"], cv["mu"])
    elif name == "Poisson2D_ManyArea":
        spec = poisson2d_manyarea(pde_obj.a_cof, pde_obj.f_cof, bbox)
    elif name == "PoissonND":
        spec = poisson_nd(cv["dim"], bbox[1])
    elif name == "Heat2D_VaryingCoef":
        spec = heat2d_varyingcoef(pde_obj.heat_2d_coef, bbox, cv["A"], cv["m"])
    elif name == "Heat2D_Multiscale":
        spec = heat2d_multiscale(bbox, cv["pde_coef"])
    elif name == "Heat2D_ComplexGeometry":
        spec = heat2d_complexgeometry(bbox)
    elif name == "Heat2D_LongTime":
        spec = heat2d_longtime(bbox, cv["k"], cv["m1"], cv["m2"])
    elif name == "HeatND":
        spec = heat_nd(cv["dim"], bbox[-1])
    elif name == "NS2D_Classic":
        spec = ns2d_classic(pde_obj.nu, bbox, linear=(pde_obj.pde.__name__ == "ns_pde_linear"))
    elif name == "NS2D_LidDriven":
        spec = ns2d_liddriven(cv["nu"], bbox)
    elif name == "NS2D_BackStep":
        spec = ns2d_backstep(cv["nu"], bbox)
    elif name == "NS2D_LongTime":
        spec = ns2d_longtime(cv["nu"], bbox)
    elif name == "Wave1D":
        spec = wave1d(cv["C"], bbox[1])
    elif name == "Wave2D_Heterogeneous":
        spec = wave2d_heterogeneous(pde_obj.darcy_2d_coef, bbox)
    elif name == "Wave2D_LongTime":
        spec = wave2d_longtime(bbox, cv["a"])
    elif name == "GrayScottEquation":
        spec = grayscott(bbox, cv["b"], cv["d"], cv["epsilon"])
    elif name == "KuramotoSivashinskyEquation":
        spec = kuramoto_sivashinsky(bbox, cv["alpha"], cv["beta"], cv["gamma"])
    elif name == "Helmholtz2D":
        spec = helmholtz2d(cv["scale"], cv["A"], cv["k"])
    elif name == "PoissonInv":
        spec = poisson_inv(bbox)
    elif name == "HeatInv":
        spec = heat_inv(bbox)
    else:  # pragma: no cover
        raise UnsupportedProblem(name)
    if gepinn:
        if pde_obj.num_gepinn != pde_obj.input_dim * pde_obj.num_pde:
            raise UnsupportedProblem("gPINN: unexpected number of residual-gradient terms")
        g = gpinn(spec)
        if g.net_d != pde_obj.input_dim or g.desc.n_pde != pde_obj.num_pde + pde_obj.num_gepinn:
            raise UnsupportedProblem(f"{name}: gPINN descriptor does not match the reference object")
        return g
    if spec.desc.d != pde_obj.input_dim or spec.desc.m != pde_obj.output_dim or spec.desc.n_pde != pde_obj.num_pde:
        raise UnsupportedProblem(
            f"{name}: descriptor (d={spec.desc.d}, m={spec.desc.m}, n_pde={spec.desc.n_pde}) does not match the reference "
            f"object (d={pde_obj.input_dim}, m={pde_obj.output_dim}, n_pde={pde_obj.num_pde})")
    return spec
