"""Generate the L2RE parity fixtures (tests/golden/l2re_<case>.npz) by training the UNMODIFIED reference on CPU.

    python tests/golden/make_l2re_golden.py [--steps 1000] [--log-every 100] [--twins 2] [case ...]

For every case the reference is driven exactly as benchmark.py:84-122 does it -- ``pde = Class()``,
``dde.nn.FNN([d] + [100]*5 + [m], "tanh", "Glorot normal")``, ``pde.create_model(net)``,
``model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=ones)``, ``model.train(iterations, display_every,
callbacks=[TesterCallback(log_every)])`` -- with a fixed seed, and the L2RE trajectory is read back from the
``errors.txt`` its own TesterCallback writes (src/utils/callbacks.py:171-176, 200-208: l2re = sqrt(mean((y - y_ref)^2)) /
sqrt(mean(y_ref^2)) on the solution table ref/*.dat or on ``ref_sol`` samples).

Also stored, so that the replay needs no reference tree: the initial flat parameter vector, ``data.train_x`` (BC rows
first), ``num_bcs``, every boundary term as (beg, end, component, values) -- all terms of the chosen cases are value-type --
the test set (test_x, test_y) and the trajectories of ``--twins`` runs whose initial weights were multiplied by
(1 + 1e-6 * N(0,1)): the reference's own sensitivity band (probe p8).
"""
import argparse
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (reference class, constructor kwargs, bias std of the initial weights, training_points kwargs or None)
    "Burgers1D": ("Burgers1D", {}, 0.0, None),                       # BASELINE configs[0]: ref/burgers1d.dat
    "Poisson2D_Classic": ("Poisson2D_Classic", {}, 0.0, None),       # ref/poisson1_cg_data.dat
    "NS2D_LidDriven": ("NS2D_LidDriven", {}, 0.1, None),             # ref/lid_driven_a4.dat, non-zero biases
    "Heat2D_Multiscale": ("Heat2D_Multiscale", {}, 0.0, dict(domain=2048, boundary=512)),   # analytic ref_sol, fewer points (CPU time)
}
SEED = 20261020


def build(name, flat_init=None):
    """(dde, pde_obj, net, model) exactly as benchmark.py:84-122 builds a case."""
    import refshim
    import torch
    dde, _ = refshim.load()
    cls, kwargs, bias_std, tp = CASES[name]
    dde.config.set_random_seed(SEED)
    pde = refshim.construct(cls, **kwargs)
    if tp is not None:
        pde.training_points(**tp)
    net = dde.nn.FNN([pde.input_dim] + [100] * 5 + [pde.output_dim], "tanh", "Glorot normal").float()
    if flat_init is not None:
        set_flat(net, flat_init)
    elif bias_std > 0:
        g = torch.Generator(device="cpu").manual_seed(SEED + 1)
        with torch.no_grad():
            for lin in net.linears:
                lin.bias.copy_((torch.randn(lin.bias.shape, generator=g, device="cpu") * bias_std).to(lin.bias.device))
    opt = torch.optim.Adam(net.parameters(), 1e-3)
    with refshim._in_reference_dir():
        model = pde.create_model(net)
    model.compile(opt, loss_weights=np.ones(pde.num_loss))
    return dde, pde, net, model


def get_flat(net):
    import torch
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu().numpy().astype(np.float32)


def set_flat(net, flat):
    import torch
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            n = p.numel()
            p.copy_(torch.as_tensor(flat[off:off + n]).reshape(p.shape))
            off += n


def train_and_read(model, steps, log_every):
    """model.train with the reference's TesterCallback; returns (epochs, l2re) from the errors.txt it writes."""
    from src.utils.callbacks import TesterCallback
    with tempfile.TemporaryDirectory() as tmp:
        model.train(iterations=steps, display_every=log_every, callbacks=[TesterCallback(log_every=log_every, verbose=False, fRMSE_param={"enable": False})],
                    model_save_path=tmp, save_model=False)
        err = np.loadtxt(os.path.join(tmp, "errors.txt")).reshape(-1, 10)
    return err[:, 0].astype(np.int64), err[:, 5]


def test_set(pde, model):
    """The points / values TesterCallback.on_train_begin evaluates on (src/utils/callbacks.py:134-150)."""
    if pde.ref_sol is not None:
        n = 2500 if pde.input_dim == 2 else 20000
        fn = getattr(model.data.geom, "uniform_points", None) or model.data.geom.random_points
        tx = fn(n, boundary=False)
        return tx, pde.ref_sol(tx)
    mask = np.isnan(pde.ref_data).any(axis=1)
    return pde.ref_data[~mask, :pde.input_dim], pde.ref_data[~mask, pde.input_dim:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(CASES))
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--log-every", type=int, default=100)
    ap.add_argument("--twins", type=int, default=2)
    a = ap.parse_args()
    import torch
    from pinnacle_b200.solver import _value_bc_info
    torch.set_num_threads(os.cpu_count())
    for name in a.cases:
        dde, pde, net, model = build(name)
        flat0 = get_flat(net)
        data = model.data
        train_x = np.asarray(data.train_x, dtype=np.float32).copy()
        num_bcs = np.asarray(data.num_bcs, dtype=np.int64)
        starts = np.cumsum([0] + list(num_bcs))
        bc_comp, bc_vals = [], []
        for i, bc in enumerate(data.bcs):
            info = _value_bc_info(bc)
            assert info is not None, f"{name}: boundary term {i} ({type(bc).__name__}) is not value-type"
            v = info[1](data.train_x, int(starts[i]), int(starts[i + 1]))
            v = v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)
            bc_comp.append(info[0])
            bc_vals.append(np.broadcast_to(np.asarray(v, dtype=np.float32).reshape(-1, 1) if np.ndim(v) else np.full((1, 1), v, np.float32),
                                           (int(num_bcs[i]), 1)).copy())
        tx, ty = test_set(pde, model)
        ep, l2 = train_and_read(model, a.steps, a.log_every)
        print(f"{name}: reference L2RE {l2[0]:.5f} -> {l2[-1]:.5f} after {a.steps} Adam steps", flush=True)
        twins = []
        for t in range(a.twins):
            rs = np.random.RandomState(SEED + 100 + t)
            flat_t = (flat0.astype(np.float64) * (1.0 + 1e-6 * rs.standard_normal(flat0.shape))).astype(np.float32)
            _, _, _, model_t = build(name, flat_t)
            assert np.array_equal(np.asarray(model_t.data.train_x, dtype=np.float32), train_x), "sampler is not deterministic"
            _, l2t = train_and_read(model_t, a.steps, a.log_every)
            twins.append(l2t)
            print(f"   twin {t}: {l2t[-1]:.5f}", flush=True)
        np.savez_compressed(os.path.join(HERE, f"l2re_{name}.npz"), flat_init=flat0, train_x=train_x, num_bcs=num_bcs,
                            bc_component=np.asarray(bc_comp, dtype=np.int64), bc_values=np.concatenate(bc_vals, axis=0).astype(np.float32),
                            test_x=np.asarray(tx, dtype=np.float32), test_y=np.asarray(ty, dtype=np.float32), epochs=ep, l2re=l2,
                            l2re_twins=np.asarray(twins).reshape(a.twins, -1), steps=a.steps, log_every=a.log_every, lr=1e-3, seed=SEED)


if __name__ == "__main__":
    main()
