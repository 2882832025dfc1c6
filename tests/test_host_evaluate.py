"""Device-resident evaluation (pinnacle_b200/evaluate.py), the part that runs without a GPU: lazy
``_test()`` outputs, the RAR selection and TesterCallback's metrics against their host formulas, the staged-point-set LRU."""
import numpy as np
import pytest
import torch

import refshim
from util import oracle_backend

from pinnacle_b200 import evaluate
from pinnacle_b200.nn import FNN

needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not available")


def _flat(net):
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()])


def test_lazy_outputs_keep_the_weights_of_the_evaluation_and_cost_nothing_unread():
    torch.manual_seed(0)
    net = FNN([3] + [100] * 4 + [2])
    x = np.random.RandomState(0).rand(257, 3).astype(np.float32)
    want = net(torch.as_tensor(x)).detach().numpy()
    before = evaluate.D2H_BYTES["outputs"]
    lazy = evaluate.LazyOutputs(_flat(net).clone(), (3, 100, 4, 2), x, torch.device("cpu"))
    arr = lazy.detach().cpu().numpy()                 # the chain deepxde.utils.to_numpy applies
    assert arr.shape == (257, 2) and len(arr) == 257 and not arr.materialized
    assert evaluate.D2H_BYTES["outputs"] == before     # nothing computed, nothing copied
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.5)                               # training goes on ...
    np.testing.assert_allclose(np.asarray(arr), want, rtol=1e-5, atol=1e-6)     # ... the outputs are those of the evaluation
    assert arr.materialized and evaluate.D2H_BYTES["outputs"] == before + 257 * 2 * 4
    np.testing.assert_allclose(want - arr, np.zeros_like(want), atol=1e-6)      # ndarray arithmetic (deepxde metrics)
    np.testing.assert_allclose(arr[3:5], want[3:5], rtol=1e-5, atol=1e-6)
    assert arr.T.shape == (2, 257) and arr.astype(np.float64).dtype == np.float64 and abs(float(arr.mean()) - float(want.mean())) < 1e-6


def test_topk_residual_selects_like_the_reference_rar_loop():
    rs = np.random.RandomState(1)
    res = rs.standard_normal((1000, 3)).astype(np.float32)

    class Op:
        def residual(self, params):
            return torch.as_tensor(res)

    idx, mean_err = evaluate.topk_residual(Op(), [], 7)
    # src/utils/rar.py:21-30 on the list-of-[N,1] structure model.predict(operator=pde) returns
    f = [res[:, k:k + 1] for k in range(3)]
    err = np.abs(f).squeeze()
    err = np.sum(err, axis=0)
    np.testing.assert_array_equal(idx, np.argsort(err)[-7:])
    assert abs(mean_err - np.mean(err)) < 1e-6 * np.mean(err)


def test_device_tester_metrics_equal_the_reference_formulas():
    torch.manual_seed(2)
    net = FNN([2] + [100] * 3 + [1])
    rs = np.random.RandomState(2)
    tx = rs.rand(500, 2).astype(np.float32)
    ty = rs.standard_normal((500, 1))
    t = evaluate.DeviceTester(tx, ty, log_every=1)
    got = t.evaluate(net)
    y = net(torch.as_tensor(tx)).detach().numpy().astype(np.float64)
    # src/utils/callbacks.py:161-176
    mse = ((y - ty) ** 2).mean()
    mae = np.abs(y - ty).mean()
    want = [mse, mae, np.max(np.abs(y - ty)), mae / np.abs(ty).mean(), np.sqrt(mse) / np.sqrt((ty ** 2).mean()), np.abs((y - ty).mean())]
    np.testing.assert_allclose(got, want, rtol=1e-6)


@needs_ref
def test_patched_test_path_is_lazy_and_switches_between_resident_point_sets():
    """model._test() of a patched model: the same loss_train / loss_test as the reference, y_pred_* equal to the reference's
    when read, nothing computed when not; train rows and test rows each uploaded once however often _test() alternates."""
    dde, _ = refshim.load()
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused
    torch.manual_seed(0)
    dde.config.set_random_seed(0)
    pde = refshim.construct("Burgers1D")
    pde.training_points(domain=96, boundary=32, initial=32)
    pde.num_test_points = 64
    net = dde.nn.FNN([2] + [100] * 5 + [1], "tanh", "Glorot normal").float()
    model = pde.create_model(net)
    model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
    model.train_state.set_data_train(*model.data.train_next_batch())
    model.train_state.set_data_test(*model.data.test())
    model._test()
    ref = (np.array(model.train_state.y_pred_train), np.array(model.train_state.loss_train), np.array(model.train_state.y_pred_test),
           np.array(model.train_state.loss_test))
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model)
            before = evaluate.D2H_BYTES["outputs"]
            for _ in range(3):
                model._test()
                model._train_step(model.train_state.X_train, None, None)
            model._test()
            st = model.fused_core.op.stats
            assert st["uploads"] == 0 or st["uploads"] <= 2          # CPU "device": set_points counts only real copies
            assert len(model.fused_core._staged) == 2                # train rows + test rows, switched, not re-staged
            assert evaluate.D2H_BYTES["outputs"] == before
            assert not model.train_state.y_pred_test.materialized
    finally:
        fused.FusedPDEResidual.__init__ = orig
    # a fresh, untrained twin for the value comparison (the loop above trained the net)
    torch.manual_seed(0)
    dde.config.set_random_seed(0)
    pde2 = refshim.construct("Burgers1D")
    pde2.training_points(domain=96, boundary=32, initial=32)
    pde2.num_test_points = 64
    net2 = dde.nn.FNN([2] + [100] * 5 + [1], "tanh", "Glorot normal").float()
    model2 = pde2.create_model(net2)
    model2.compile(torch.optim.Adam(net2.parameters(), 1e-3), loss_weights=np.ones(pde2.num_loss))
    model2.train_state.set_data_train(*model2.data.train_next_batch())
    model2.train_state.set_data_test(*model2.data.test())
    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model2)
            model2._test()
    finally:
        fused.FusedPDEResidual.__init__ = orig
    np.testing.assert_allclose(model2.train_state.loss_train, ref[1], rtol=2e-5)
    np.testing.assert_allclose(model2.train_state.loss_test, ref[3], rtol=2e-5)
    np.testing.assert_allclose(np.asarray(model2.train_state.y_pred_train), ref[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.asarray(model2.train_state.y_pred_test), ref[2], rtol=1e-5, atol=1e-6)


@needs_ref
def test_fused_rar_wrapper_adds_the_same_anchor_as_the_reference_loop():
    dde, _ = refshim.load()
    from src.utils.rar import rar_wrapper
    from pinnacle_b200 import fused
    from pinnacle_b200.solver import enable_fused

    def build():
        torch.manual_seed(1)
        dde.config.set_random_seed(1)
        pde = refshim.construct("Poisson2D_Classic")
        pde.training_points(domain=128, boundary=48)
        net = dde.nn.FNN([2] + [100] * 5 + [1], "tanh", "Glorot normal").float()
        model = pde.create_model(net)
        model.compile(torch.optim.Adam(net.parameters(), 1e-3), loss_weights=np.ones(pde.num_loss))
        return pde, model

    pde_a, ref_model = build()
    ref_model.train = rar_wrapper(pde_a, ref_model, {"interval": 2, "count": 3})
    ref_model.train(iterations=4, display_every=2)
    pde_b, model = build()
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            enable_fused(model)
            before = evaluate.D2H_BYTES["topk"]
            model.train = evaluate.fused_rar_wrapper(pde_b, model, {"interval": 2, "count": 3})
            model.train(iterations=4, display_every=2)
            assert evaluate.D2H_BYTES["topk"] - before == (3 + 1) * 8          # three indices and the mean, per query
    finally:
        fused.FusedPDEResidual.__init__ = orig
    a, b = ref_model.data.anchors, model.data.anchors
    assert a.shape == b.shape == (3, 2)
    np.testing.assert_array_equal(np.sort(a, axis=0), np.sort(b, axis=0))


def test_staged_point_sets_are_an_lru_of_three_and_drop_their_term_ops():
    """_StepCore.stage keeps the last three host arrays resident; a fourth evicts the oldest together with the boundary-term
    ops built on it; re-staging a resident array is a switch (no new generation)."""
    from pinnacle_b200 import fused, problems
    from pinnacle_b200.solver import FusedPDESolver, ValueBC
    spec = problems.make("Burgers1D")
    torch.manual_seed(0)
    net = FNN([2] + [100] * 5 + [1])
    xb = spec.sample(16, seed=1)
    terms = [ValueBC(0, 16, 0, np.zeros((16, 1), np.float32))]
    orig = fused.FusedPDEResidual.__init__

    def init_cpu(self, spec, H=100, L=5, policy="auto", device=None, group=None, check_library=True):
        orig(self, spec, H, L, policy, "cpu", group, False)

    fused.FusedPDEResidual.__init__ = init_cpu
    try:
        with oracle_backend():
            s = FusedPDESolver(spec, net, spec.sample(40, seed=2), xb, terms, device="cpu").compile(torch.optim.Adam(net.parameters(), 1e-3))
            arrays = [np.vstack([xb, spec.sample(40, seed=10 + i)]).astype(np.float32) for i in range(4)]
            gens = []
            for a in arrays[:3]:
                s.outputs_losses_train(a)
                gens.append(s.core.generation)
            assert len(set(gens)) == 3 and len(s.core._staged) == 3
            assert {k[0] for k in s.core._value_ops} == set(gens)
            s.outputs_losses_train(arrays[1])                       # resident: switched, not re-staged
            assert s.core.generation == gens[1] and len(s.core._staged) == 3
            s.outputs_losses_train(arrays[3])                       # evicts the least recently used (arrays[0])
            assert len(s.core._staged) == 3 and gens[0] not in {k[0] for k in s.core._value_ops}
            l_a = s.outputs_losses_train(arrays[0])[1]              # comes back as a new generation with the same losses
            s2 = FusedPDESolver(spec, net, arrays[0][16:], xb, terms, device="cpu").compile(torch.optim.Adam(net.parameters(), 1e-3))
            np.testing.assert_allclose(l_a.detach().numpy(), s2.outputs_losses_train()[1].detach().numpy(), rtol=1e-6)
    finally:
        fused.FusedPDEResidual.__init__ = orig
