"""Residual-kind descriptors for the fused PDE-residual op.

A *kind descriptor* tells the CUDA path everything it needs about one PINNacle
problem: input/output dimension, tanh-MLP shape, which Taylor-jet orders to carry per
input direction, which residual formula to evaluate and its constants.  It is the
Python mirror of ``struct pinn_kind_desc`` in ``include/pinn_b200.h`` (same field order,
same sizes) and is passed by pointer through the C-ABI.

Reference behaviour this replaces: the per-PDE ``pde(x, u)`` closures of
``src/pde/*.py`` (SURVEY.md §8a "Residual kinds") together with the
``dde.grad.jacobian/hessian`` calls inside them (``deepxde/gradients.py:160-292``).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Sequence

PINN_MAX_DIM = 6
PINN_MAX_CONSTS = 24
PINN_MAX_AUXSEL = 8
PINN_MAX_PDE = 3

# residual kinds (must match enum pinn_kind in include/pinn_b200.h)
KIND_LINEAR = 0          # r = cu*u + sum_j (c1_j u_j + c2_j u_jj) + src   (coefficients optionally per point)
KIND_BURGERS1D = 1       # src/pde/burgers.py:21-25
KIND_BURGERS2D = 2       # src/pde/burgers.py:66-80
KIND_NS2D = 3            # src/pde/ns.py:30-49,141-160,240-259
KIND_NS2D_UNSTEADY = 4   # src/pde/ns.py:330-352
KIND_HEAT_NLSRC = 5      # src/pde/heat.py:214-222
KIND_GRAYSCOTT = 6       # src/pde/chaotic.py:21-31
KIND_KS = 7              # src/pde/chaotic.py:77-83
KIND_VALUE = 8           # value-type BC/IC/point-set term: u_c - g(x)  (icbc/boundary_conditions.py:73-80,214-241)
KIND_FLUX = 9            # flux-type BC term: n(x).grad u_c + beta(x) u_c - g(x)  (icbc/boundary_conditions.py:44-57,83-105)
KIND_DIV_AGRAD = 10      # inverse problems, outputs (u, a): c0 u_t + c1 div(a grad u) + c2 src   (inverse.py:19-24, 129-136)

KIND_GP_LINEAR2 = 11     # gPINN: LINEAR residual in two inputs + its input gradient (baseclass.py:151-180); 4 jet directions, order 3
KIND_GP_BURGERS1D = 12   # gPINN: Burgers1D residual + its input gradient

KIND_NAMES = {
    KIND_LINEAR: "linear",
    KIND_BURGERS1D: "burgers1d",
    KIND_BURGERS2D: "burgers2d",
    KIND_NS2D: "ns2d",
    KIND_NS2D_UNSTEADY: "ns2d_unsteady",
    KIND_HEAT_NLSRC: "heat_nlsrc",
    KIND_GRAYSCOTT: "grayscott",
    KIND_KS: "ks",
    KIND_VALUE: "value",
    KIND_FLUX: "flux",
    KIND_DIV_AGRAD: "div_agrad",
    KIND_GP_LINEAR2: "gpinn_linear2",
    KIND_GP_BURGERS1D: "gpinn_burgers1d",
}

# LINEAR kind constant layout
LIN_CU = 0                    # consts[0]          coefficient of u
LIN_C1 = 1                    # consts[1 + j]      coefficient of du/dx_j
LIN_C2 = 1 + PINN_MAX_DIM     # consts[7 + j]      coefficient of d2u/dx_j^2
# LINEAR kind aux selectors (column of the per-point aux array, -1 = unused)
LIN_SEL_SRC = 0               # aux_sel[0]   additive per-point source term
LIN_SEL_CU = 1                # aux_sel[1]   per-point multiplier of the cu*u term
LIN_SEL_C2 = 2                # aux_sel[2+j] per-point multiplier of the c2_j*u_jj term


class CKindDesc(ctypes.Structure):
    """ctypes mirror of ``struct pinn_kind_desc`` (include/pinn_b200.h)."""

    _fields_ = [
        ("kind", ctypes.c_int32),
        ("d", ctypes.c_int32),
        ("m", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("L", ctypes.c_int32),
        ("n_pde", ctypes.c_int32),
        ("n_aux", ctypes.c_int32),
        ("order", ctypes.c_int32 * PINN_MAX_DIM),
        ("aux_sel", ctypes.c_int32 * PINN_MAX_AUXSEL),
        ("consts", ctypes.c_float * PINN_MAX_CONSTS),
    ]


PINN_MAX_HOLES = 32
HOLE_BALL, HOLE_BOX = 0, 1


class CHole(ctypes.Structure):
    """ctypes mirror of ``struct pinn_hole``."""

    _fields_ = [("shape", ctypes.c_int32), ("n_dims", ctypes.c_int32), ("a", ctypes.c_float * 3), ("b", ctypes.c_float * 3)]


class CRegion(ctypes.Structure):
    """ctypes mirror of ``struct pinn_region``: the sampling region of ``pinn_sample_points``."""

    _fields_ = [("d", ctypes.c_int32), ("n_holes", ctypes.c_int32), ("lo", ctypes.c_float * PINN_MAX_DIM), ("hi", ctypes.c_float * PINN_MAX_DIM),
                ("keep_dims", ctypes.c_int32), ("keep_radius", ctypes.c_float), ("keep_center", ctypes.c_float * PINN_MAX_DIM),
                ("holes", CHole * PINN_MAX_HOLES)]


@dataclass
class KindDesc:
    kind: int
    d: int
    m: int
    order: Sequence[int]
    n_pde: int = 1
    n_aux: int = 0
    consts: List[float] = field(default_factory=list)
    aux_sel: List[int] = field(default_factory=list)
    H: int = 100
    L: int = 5
    name: str = ""

    def __post_init__(self):
        self.order = tuple(int(k) for k in self.order)
        if len(self.order) != self.d:
            raise ValueError(f"order has {len(self.order)} entries for d={self.d}")
        if self.d > PINN_MAX_DIM:
            raise ValueError(f"d={self.d} exceeds PINN_MAX_DIM={PINN_MAX_DIM}")
        if any(k not in (0, 1, 2, 3, 4) for k in self.order):
            raise ValueError(f"unsupported jet order in {self.order}")
        if not 1 <= self.n_pde <= PINN_MAX_PDE:
            raise ValueError(f"n_pde={self.n_pde} outside 1..{PINN_MAX_PDE}")
        self.consts = [float(c) for c in self.consts]
        self.consts += [0.0] * (PINN_MAX_CONSTS - len(self.consts))
        if len(self.consts) > PINN_MAX_CONSTS:
            raise ValueError("too many constants")
        self.aux_sel = [int(a) for a in self.aux_sel]
        self.aux_sel += [-1] * (PINN_MAX_AUXSEL - len(self.aux_sel))
        if len(self.aux_sel) > PINN_MAX_AUXSEL:
            raise ValueError("too many aux selectors")
        for a in self.aux_sel:
            if a >= self.n_aux:
                raise ValueError(f"aux selector {a} but n_aux={self.n_aux}")

    # ---- jet channel bookkeeping -------------------------------------------------
    @property
    def n_chan(self) -> int:
        """C = 1 + sum_j K_j jet channels per neuron."""
        return 1 + sum(self.order)

    def chan(self, j: int, k: int) -> int:
        """Channel index of the k-th Taylor coefficient along direction j (k>=1)."""
        if not 1 <= k <= self.order[j]:
            raise ValueError(f"direction {j} carries order {self.order[j]}, asked {k}")
        return 1 + sum(self.order[:j]) + (k - 1)

    @property
    def n_params(self) -> int:
        """P = dH+H + (L-1)(H^2+H) + Hm+m, nn.Linear order (fnn.py:25-33)."""
        return self.d * self.H + self.H + (self.L - 1) * (self.H * self.H + self.H) + self.H * self.m + self.m

    def layer_shapes(self):
        sizes = [self.d] + [self.H] * self.L + [self.m]
        return [(sizes[i + 1], sizes[i]) for i in range(len(sizes) - 1)]

    # ---- canonical work model ---------------------
    def flops_fwd_per_point(self) -> float:
        C, H, L, d, m = self.n_chan, self.H, self.L, self.d, self.m
        return 2.0 * d * H + 2.0 * C * ((L - 1) * H * H + H * m)

    def flops_step_per_point(self) -> float:
        C, H, L, d, m = self.n_chan, self.H, self.L, self.d, self.m
        return 4.0 * d * H + 6.0 * C * ((L - 1) * H * H + H * m)

    def tc_executed_flops_per_point(self, n_seeds: int = 1, n_chan: int | None = None) -> float:
        """Tensor-core flops the tcgen05 kernel EXECUTES per point (bf16x3: 6 split products, neurons padded to
        128 x 112): L forward batches (layers 2..L and the zero-padded output layer) + per seed (L-1) x (dW + Abar).
        `n_chan`: channels the library actually propagates (capi.effective_channels: d+2 under the forward-Laplacian
        collapse) -- defaults to the canonical C = 1 + sum(order)."""
        per_batch = 6.0 * (112 / 16) * 2.0 * 128 * 16 * (self.n_chan if n_chan is None else n_chan)      # = C * 172032
        return per_batch * (self.L + 2.0 * (self.L - 1) * n_seeds)

    def bytes_per_point(self) -> int:
        return 4 * (self.d + self.n_aux)

    def with_net(self, H: int, L: int) -> "KindDesc":
        return KindDesc(self.kind, self.d, self.m, self.order, self.n_pde, self.n_aux,
                        list(self.consts), list(self.aux_sel), H, L, self.name)

    def to_c(self) -> CKindDesc:
        c = CKindDesc()
        c.kind, c.d, c.m, c.H, c.L = self.kind, self.d, self.m, self.H, self.L
        c.n_pde, c.n_aux = self.n_pde, self.n_aux
        for j in range(PINN_MAX_DIM):
            c.order[j] = self.order[j] if j < self.d else 0
        for i in range(PINN_MAX_AUXSEL):
            c.aux_sel[i] = self.aux_sel[i]
        for i in range(PINN_MAX_CONSTS):
            c.consts[i] = self.consts[i]
        return c


def value_desc(d: int, m: int, component: int, name: str = "") -> KindDesc:
    """Descriptor of a value-type boundary / initial / point-set term r = u_component - target (target = aux column 0)."""
    if not 0 <= component < m:
        raise ValueError(f"component {component} outside 0..{m - 1}")
    return KindDesc(KIND_VALUE, d, m, (0,) * d, 1, 1, [float(component)], [0], name=name)


def flux_desc(d: int, m: int, component: int, with_beta: bool, name: str = "") -> KindDesc:
    """Descriptor of a Neumann / affine-Robin boundary term r = n.grad u_c + beta u_c - g.
    aux columns: [g, beta (if with_beta), n_0 .. n_{d-1}]."""
    if not 0 <= component < m:
        raise ValueError(f"component {component} outside 0..{m - 1}")
    nb = 2 if with_beta else 1
    sel = [0, 1 if with_beta else -1] + [nb + j for j in range(d)]
    return KindDesc(KIND_FLUX, d, m, (1,) * d, 1, nb + d, [float(component)], sel, name=name)


def linear_desc(d: int, order: Sequence[int], cu: float = 0.0, c1: Sequence[float] = (), c2: Sequence[float] = (),
                n_aux: int = 0, src_col: int = -1, cu_col: int = -1, c2_cols: Sequence[int] = (), m: int = 1,
                name: str = "") -> KindDesc:
    """Build a LINEAR-kind descriptor: r = cu*[aux]*u + sum_j c1_j u_j + sum_j c2_j*[aux]*u_jj + [aux src]."""
    consts = [0.0] * PINN_MAX_CONSTS
    consts[LIN_CU] = cu
    for j, c in enumerate(c1):
        consts[LIN_C1 + j] = c
    for j, c in enumerate(c2):
        consts[LIN_C2 + j] = c
    for j in range(d):
        if consts[LIN_C2 + j] != 0.0 and order[j] < 2:
            raise ValueError(f"direction {j} needs order 2")
        if consts[LIN_C1 + j] != 0.0 and order[j] < 1:
            raise ValueError(f"direction {j} needs order 1")
    sel = [-1] * PINN_MAX_AUXSEL
    sel[LIN_SEL_SRC] = src_col
    sel[LIN_SEL_CU] = cu_col
    for j, c in enumerate(c2_cols):
        sel[LIN_SEL_C2 + j] = c
    return KindDesc(KIND_LINEAR, d, m, order, 1, n_aux, consts, sel, name=name)
