"""Final-L2RE parity of fixed-seed training runs (BASELINE metric "L2RE parity vs reference", configs[0]).

The fixtures tests/golden/l2re_<case>.npz hold runs of the UNMODIFIED reference on CPU (tests/golden/make_l2re_golden.py:
benchmark.py's construction, Adam lr 1e-3, 1000 iterations, L2RE by the reference's own TesterCallback,
src/utils/callbacks.py:171-176) together with the runs of twins whose initial weights were perturbed by 1e-6 -- the reference's
own sensitivity band.  Two replays on the GPU:

* ``test_l2re_replay_*``: reference-free -- the fixture's points, boundary data and initial weights through
  ``FusedPDESolver`` (fused PDE + boundary terms, default kernel), torch.optim.Adam / FlatAdam, L2RE by the same formula;
* ``test_l2re_real_reference_model_*``: the real ``dde.Model`` of baseline/_ref built as benchmark.py:84-122 does,
  ``enable_fused(model)``, ``model.train(..., callbacks=[TesterCallback])`` -- trajectory read from the errors.txt the reference's
  callback writes.

Tolerance.  The bar of the north star -- final L2RE within 5 % of the reference's -- is applied to the L2RE at the END of the run
averaged over the last three logged epochs, or twice the spread of the reference's own three runs in that statistic where that is
wider.  A single logged epoch is not a stable statistic on every case: late in the NS2D_LidDriven run the L2RE shows +10 % spikes
(Adam at lr 1e-3 on a stiff loss) that land on DIFFERENT logged epochs in the reference's own three runs (0.544 at iteration 900
in one, 0.564 at 600 in another), and in the fused runs the spike moves between 800, 900 and 1000 from one kernel build to the
next while the three-epoch mean stays within 0.3 % of the reference's.  Per logged epoch (trajectory sanity): the first epoch,
before trajectories can diverge, at the plain 5 %; the others at 10 % / twice the spread of the reference's runs at that epoch,
and from the middle of the run on at twice the widest late-stage spread.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["Burgers1D", "Poisson2D_Classic", "NS2D_LidDriven", "Heat2D_Multiscale"]


def _fixture(name):
    path = os.path.join(HERE, "golden", f"l2re_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


def _tolerance(g):
    runs = np.vstack([g["l2re"][None, :], g["l2re_twins"]])
    spread = runs.max(axis=0) - runs.min(axis=0)
    frac = np.full(runs.shape[1], 0.10)
    frac[0] = frac[-1] = 0.05
    tol = np.maximum(frac * g["l2re"], 2.0 * spread)
    # second half of the run: transient spikes of the loss (+10 % of L2RE on NS2D_LidDriven) land on different logged epochs in
    # different runs of the reference itself -- use the widest late-stage spread for all of them
    half = runs.shape[1] // 2
    tol[half:] = np.maximum(tol[half:], 2.0 * spread[half:].max())
    return tol


def _final(l2):
    """L2RE at the end of the run: mean over the last three logged epochs."""
    return np.asarray(l2)[..., -3:].mean(axis=-1)


def _check(name, l2, g, what):
    ref, tol = g["l2re"], _tolerance(g)
    dev = np.abs(l2 - ref)
    print(f"{name} [{what}] L2RE fused {l2[-1]:.5f}  reference {ref[-1]:.5f}  twins {g['l2re_twins'][:, -1]}  |diff|/ref {dev[-1] / ref[-1]:.3%}  tolerance {tol[-1] / ref[-1]:.3%}")
    assert np.all(np.isfinite(l2))
    assert np.all(dev <= tol), (name, what, l2.tolist(), ref.tolist(), tol.tolist())
    assert dev[0] <= 0.05 * ref[0]            # before the trajectories have had time to diverge: the plain 5 % bar
    # the north star's bar on the final L2RE
    runs = np.vstack([ref[None, :], g["l2re_twins"]])
    fin_ref, fin_runs = float(_final(ref)), _final(runs)
    fin_tol = max(0.05 * fin_ref, 2.0 * float(fin_runs.max() - fin_runs.min()))
    print(f"{name} [{what}] final L2RE (last three logged epochs) fused {float(_final(l2)):.5f}  reference {fin_ref:.5f}  reference runs {fin_runs}  tolerance {fin_tol / fin_ref:.3%}")
    assert abs(float(_final(l2)) - fin_ref) <= fin_tol, (name, what, float(_final(l2)), fin_ref, fin_tol)


def _l2re(y, y_ref):
    # src/utils/callbacks.py:171-176
    return float(np.sqrt(((y - y_ref) ** 2).mean()) / np.sqrt((y_ref ** 2).mean()))


@pytest.mark.parametrize("flat_opt", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_l2re_replay_through_fused_solver(name, flat_opt):
    from pinnacle_b200 import capi, problems
    from pinnacle_b200.nn import FNN
    from pinnacle_b200.optim import FlatAdam
    from pinnacle_b200.solver import FusedPDESolver, ValueBC
    g = _fixture(name)
    dev = torch.device("cuda:0")
    spec = problems.make(name)
    d = spec.desc
    net = FNN([d.d] + [100] * 5 + [d.m]).to(dev)
    off = 0
    flat0 = torch.as_tensor(g["flat_init"])
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(flat0[off:off + p.numel()].reshape(p.shape))
            off += p.numel()
    num_bcs = g["num_bcs"]
    n_bc = int(num_bcs.sum())
    starts = np.cumsum([0] + list(num_bcs))
    x = g["train_x"]
    terms = [ValueBC(int(starts[i]), int(starts[i + 1]), int(g["bc_component"][i]), g["bc_values"][int(starts[i]):int(starts[i + 1])])
             for i in range(len(num_bcs))]
    opt = FlatAdam(net.parameters(), lr=float(g["lr"])) if flat_opt else torch.optim.Adam(net.parameters(), lr=float(g["lr"]))
    solver = FusedPDESolver(spec, net, x[n_bc:], x[:n_bc], terms).compile(opt, loss_weights=np.ones(d.n_pde + len(terms)))
    assert capi.selected_impl(d.with_net(100, 5)) in (capi.IMPL_TC, capi.IMPL_TC_PIPELINED)       # the DEFAULT kernel
    tx = torch.as_tensor(g["test_x"]).to(dev)
    ty = g["test_y"]
    l2 = []
    log_every, steps = int(g["log_every"]), int(g["steps"])
    for it in range(1, steps + 1):
        solver.train_step()
        if it % log_every == 0:
            with torch.no_grad():
                l2.append(_l2re(net(tx).cpu().numpy(), ty))
    _check(name, np.asarray(l2), g, "FusedPDESolver + " + ("FlatAdam" if flat_opt else "torch Adam"))


@pytest.mark.parametrize("name", ["Burgers1D", "NS2D_LidDriven"])
def test_l2re_real_reference_model_on_the_fused_path(name):
    """benchmark.py's own construction on CUDA (trainer.py:39-41 puts every tensor on the GPU), patched by enable_fused; runs in
    a subprocess (tests/ref_gpu_driver.py) because importing the reference on a GPU box changes the default tensor type."""
    import json
    import subprocess
    import refshim
    if not refshim.available():
        pytest.skip("reference tree not installed (scripts/install_reference.py)")
    g = _fixture(name)
    r = subprocess.run([sys.executable, os.path.join(HERE, "ref_gpu_driver.py"), "l2re", name], capture_output=True, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(lines[-1][7:])
    assert res["same_points"], "the reference drew other training points than the fixture's"
    assert res["fused_calls"] >= int(g["steps"])
    _check(name, np.asarray(res["l2re"]), g, "dde.Model + enable_fused")
