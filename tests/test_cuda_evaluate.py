"""Device-resident evaluation on the GPU: a 16 Mi-row ``_test()``-style evaluation of a patched model moves
only the loss scalars to the host; the RAR selection on the device picks the rows the reference's host loop picks; the
TesterCallback metrics on the device equal their host formulas."""
import numpy as np
import pytest
import torch

from test_cuda_solver import DirichletBC, _StubData, _StubModel, _net

pytestmark = pytest.mark.gpu


def _patched_model(n_rows, nb=256, seed=3):
    from pinnacle_b200 import problems
    from pinnacle_b200.solver import enable_fused
    dev = torch.device("cuda:0")
    spec = problems.make("Poisson2D_Classic")
    x = spec.sample(n_rows, seed=seed)
    xb = spec.sample(nb, seed=seed + 1)
    vals = np.random.RandomState(1).standard_normal((nb, 1)).astype(np.float32)
    train_x = np.vstack([xb, x]).astype(np.float32)
    net = _net(spec.desc, dev, 23)
    data = _StubData(lambda x_, u_: None, [DirichletBC(0, vals)], [nb], train_x)
    model = _StubModel(data, net, torch.optim.Adam(net.parameters(), 1e-3), np.ones(2))
    enable_fused(model, spec=spec)
    return model, net, train_x, nb


def test_sixteen_million_row_test_pass_moves_under_a_kilobyte_to_the_host():
    from pinnacle_b200 import evaluate
    n = 1 << 24
    model, net, train_x, nb = _patched_model(n)
    test_x = np.vstack([train_x[:nb], train_x[nb:nb + (1 << 20)][::-1]]).astype(np.float32)
    before = dict(evaluate.D2H_BYTES)
    d2h = 0
    net.requires_grad_(False)                                   # what Model._outputs_losses does around the call (model.py:489-493)
    for _ in range(2):                                          # two display intervals: train rows, test rows, train rows, ...
        out_tr, l_tr = model.outputs_losses_train(train_x, None)
        out_te, l_te = model.outputs_losses_test(test_x, None)
        # deepxde.utils.to_numpy on both members (model.py:499)
        y_tr, y_te = out_tr.detach().cpu().numpy(), out_te.detach().cpu().numpy()
        lt, le = l_tr.detach().cpu().numpy(), l_te.detach().cpu().numpy()
        d2h += lt.nbytes + le.nbytes
    net.requires_grad_(True)
    assert evaluate.D2H_BYTES == before and d2h < 1024, (evaluate.D2H_BYTES, d2h)
    assert y_tr.shape == (n + nb, 1) and y_te.shape == (nb + (1 << 20), 1) and not y_tr.materialized
    assert np.all(np.isfinite(lt)) and np.all(np.isfinite(le))
    st = model.fused_core.op.stats
    assert st["uploads"] == 2, st                               # each point set uploaded once, then switched
    # read on demand: equal to the plain forward at the weights of that evaluation
    with torch.no_grad():
        want = net(torch.as_tensor(test_x[:4096]).cuda()).cpu().numpy()
        for p in net.parameters():
            p.mul_(1.01)
    np.testing.assert_allclose(np.asarray(y_te)[:4096], want, rtol=1e-4, atol=1e-6)
    assert evaluate.D2H_BYTES["outputs"] - before["outputs"] == y_te.shape[0] * 4


def test_rar_selection_on_the_device_matches_the_host_loop():
    from pinnacle_b200 import evaluate
    model, net, train_x, nb = _patched_model(1 << 20)
    X = train_x[nb:]
    f = model.predict(X, operator=model.data.pde)              # host path: the full residual matrix comes back (rar.py:21)
    err = np.abs(f).squeeze()
    before = evaluate.D2H_BYTES["topk"]
    idx, mean_err = model.fused_topk_residual(X, 16)
    assert evaluate.D2H_BYTES["topk"] - before == 17 * 8
    np.testing.assert_array_equal(np.sort(idx), np.sort(np.argsort(err)[-16:]))
    np.testing.assert_allclose(err[idx], np.sort(err)[-16:], rtol=0, atol=0)
    assert abs(mean_err - err.mean(dtype=np.float64)) <= 1e-6 * err.mean(dtype=np.float64)
    dev_res = model.fused_residual(X)
    assert dev_res.is_cuda and tuple(dev_res.shape) == (X.shape[0], 1)
    np.testing.assert_array_equal(dev_res.cpu().numpy(), np.asarray(f).reshape(-1, 1))


def test_device_tester_l2re_equals_the_reference_formula():
    from pinnacle_b200 import evaluate
    g = np.load(__file__.replace("test_cuda_evaluate.py", "golden/l2re_Burgers1D.npz"))
    dev = torch.device("cuda:0")
    from pinnacle_b200.nn import FNN
    net = FNN([2] + [100] * 5 + [1]).to(dev)
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.as_tensor(g["flat_init"][off:off + p.numel()]).reshape(p.shape))
            off += p.numel()
    before = evaluate.D2H_BYTES["metrics"]
    t = evaluate.DeviceTester(g["test_x"], g["test_y"], log_every=1)
    got = t.evaluate(net)
    assert evaluate.D2H_BYTES["metrics"] - before == 48
    with torch.no_grad():
        y = net(torch.as_tensor(g["test_x"]).to(dev)).cpu().numpy().astype(np.float64)
    ty = g["test_y"].astype(np.float64)
    l2re = np.sqrt(((y - ty) ** 2).mean()) / np.sqrt((ty ** 2).mean())            # src/utils/callbacks.py:171-176
    assert abs(got[4] - l2re) <= 1e-6 * l2re
    # epoch 0 of the reference's own trajectory is not logged; its first logged value differs (100 Adam steps later) -- the
    # formula is pinned by tests/test_cuda_l2re.py, which replays that trajectory
