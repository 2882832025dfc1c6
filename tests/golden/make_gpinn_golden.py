"""Generate tests/golden/gepinn_<Class>.npz from the UNMODIFIED reference with ``use_gepinn()`` switched on
(src/pde/baseclass.py:151-180; benchmark.py:91-92 ``--method gepinn``):

    python tests/golden/make_gpinn_golden.py [Poisson2D_Classic Burgers1D]

The reference pipeline (``dde.nn.FNN`` forward, the class's ``pde`` closure wrapped by ``pde_wrapper`` -- residual plus
``dde.grad.jacobian`` of the residual per input --, mean-square per term, backward per term) is evaluated in float64 and
float32 on deterministic weights of the REAL two-input network and deterministic points.  Stored: points, per-term losses
[r, dr/dx0, dr/dx1], the residual matrix and the flat parameter gradient per term at a fixed index set (+ norms and four
random projections).  The tests present the same network to the kernels in its embedded form (first layer times the
direction matrix, rows padded with zeros) and map the gradient back.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refshim  # noqa: E402
from golden import common  # noqa: E402
from pinnacle_b200 import problems  # noqa: E402

CASES = ["Poisson2D_Classic", "Burgers1D", "Wave1D", "Helmholtz2D", "PoissonBoltzmann2D", "Poisson2D_ManyArea", "Poisson1D"]


def run(name):
    dde, _ = refshim.load()
    pde = refshim.construct(name)
    base = problems.describe_reference(pde).desc          # the plain problem: shapes of the real network
    pde.use_gepinn()
    spec = problems.describe_reference(pde)               # kind GP_*, d = 4 directions, embed = direction matrix
    desc = spec.desc
    assert spec.embed is not None and desc.n_pde == 1 + base.d
    seed = common.class_seed("gepinn_" + name)
    shapes = base.layer_shapes()
    flat64 = common.golden_params(shapes, seed)
    idx = common.golden_index_set(shapes, seed)
    probes = common.probe_vectors(base.n_params, seed)
    x64 = spec.sample(common.N_POINTS, seed=seed, dtype=np.float64).astype(np.float32).astype(np.float64)
    x_e = np.concatenate([x64, np.zeros((x64.shape[0], desc.d - base.d))], axis=1)
    aux64 = spec.aux(torch.as_tensor(x_e)).numpy() if desc.n_aux else np.zeros((x64.shape[0], 0))
    out = dict(x=x64, seed=np.int64(seed), idx=idx, kind=np.int64(desc.kind), consts=np.asarray(desc.consts, dtype=np.float64),
               embed=np.asarray(spec.embed, dtype=np.float64), n_aux=np.int64(desc.n_aux), aux_sel=np.asarray(desc.aux_sel),
               aux64=aux64, base_d=np.int64(base.d), H=np.int64(base.H), L=np.int64(base.L))
    for tag, tdt, ndt in (("64", torch.float64, np.float64), ("32", torch.float32, np.float32)):
        torch.set_default_dtype(tdt)
        dde.config.set_default_float("float64" if tag == "64" else "float32")
        net = dde.nn.FNN([base.d] + [base.H] * base.L + [1], "tanh", "Glorot normal")
        flat = torch.as_tensor(flat64.astype(ndt))
        off = 0
        with torch.no_grad():
            for p in net.parameters():
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        losses, grads, res = refshim.reference_losses_and_grads(pde, net, x64.astype(ndt))
        assert res.shape[1] == desc.n_pde
        out[f"loss{tag}"] = losses.numpy()
        out[f"resid{tag}"] = res.numpy()
        out[f"grad{tag}"] = grads.numpy()[:, idx]
        out[f"gnorm{tag}"] = grads.norm(dim=1).numpy()
        out[f"gproj{tag}"] = grads.numpy().astype(np.float64) @ probes.T
    torch.set_default_dtype(torch.float32)
    dde.config.set_default_float("float32")
    np.savez_compressed(os.path.join(HERE, f"gepinn_{name}.npz"), **out)
    print(f"{name}: losses {out['loss64']}  |grad| {out['gnorm64']}")


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        run(n)
