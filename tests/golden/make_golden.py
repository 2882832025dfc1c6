"""Generate tests/golden/<Class>.npz from the UNMODIFIED reference (run in the build container):

    python tests/golden/make_golden.py            # all classes
    python tests/golden/make_golden.py Burgers1D  # one class

For each problem class the reference pipeline (``dde.nn.FNN`` forward, the class's own ``pde``
closure with ``dde.grad.jacobian/hessian``, mean-square per term as ``dde.data.PDE.losses``
does, ``backward`` per loss term) is evaluated in float64 and float32 on deterministic weights
and points.  Stored: points, aux columns, per-term losses, the residual matrix, and the flat
parameter gradient per loss term at a fixed index set (+ norms and 4 random projections of the
full vector).  Weights are regenerated from the seed by ``common.golden_params``.
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


def run(name):
    dde, _ = refshim.load()
    pde = refshim.construct(name)
    spec = problems.describe_reference(pde)
    desc = spec.desc
    seed = common.class_seed(name)
    shapes = desc.layer_shapes()
    flat64 = common.golden_params(shapes, seed)
    idx = common.golden_index_set(shapes, seed)
    probes = common.probe_vectors(desc.n_params, seed)
    x64 = spec.sample(common.N_POINTS, seed=seed, dtype=np.float64)
    x64 = x64.astype(np.float32).astype(np.float64)       # points exactly representable in fp32
    out = dict(x=x64, seed=np.int64(seed), idx=idx, kind=np.int64(desc.kind), order=np.asarray(desc.order),
               consts=np.asarray(desc.consts, dtype=np.float64), aux_sel=np.asarray(desc.aux_sel), n_aux=np.int64(desc.n_aux),
               d=np.int64(desc.d), m=np.int64(desc.m), n_pde=np.int64(desc.n_pde))
    for tag, tdt, ndt in (("64", torch.float64, np.float64), ("32", torch.float32, np.float32)):
        torch.set_default_dtype(tdt)
        dde.config.set_default_float("float64" if tag == "64" else "float32")
        net = dde.nn.FNN([desc.d] + [desc.H] * desc.L + [desc.m], "tanh", "Glorot normal")
        flat = torch.as_tensor(flat64.astype(ndt))
        off = 0
        with torch.no_grad():
            for p in net.parameters():
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        x = x64.astype(ndt)
        losses, grads, res = refshim.reference_losses_and_grads(pde, net, x)
        aux = spec.aux(torch.as_tensor(x))
        out[f"loss{tag}"] = losses.numpy()
        out[f"resid{tag}"] = res.numpy()
        out[f"grad{tag}"] = grads.numpy()[:, idx]
        out[f"gnorm{tag}"] = grads.norm(dim=1).numpy()
        out[f"gproj{tag}"] = grads.numpy().astype(np.float64) @ probes.T
        if aux is not None:
            out[f"aux{tag}"] = aux.numpy()
    torch.set_default_dtype(torch.float32)
    dde.config.set_default_float("float32")
    np.savez_compressed(common.fixture_path(name), **out)
    print(f"{name}: N={x64.shape[0]} P={desc.n_params} n_idx={idx.size} loss64={out['loss64']}")


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(refshim.CLASS_MODULE)):
        run(n)
