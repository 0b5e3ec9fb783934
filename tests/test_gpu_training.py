"""Config 3 (scripts/train.py semantics) end to end on the GPU vs the reference's CPU training curve.

Golden: tests/golden/train_curve_mazes032.npz (tests/golden/make_golden_train.py ran the reference's
NeuralAstar + autograd + RMSprop for 4 steps on CPU).  The search is discrete, so tiny differences
between cuDNN and MKL-DNN convolutions can flip individual selections; the curve must agree to 2 %.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_curve_mazes032.npz")


def _batch():
    z = np.load(GOLD)
    B, H, W = (int(v) for v in z["shape"])
    N = H * W

    def bits(k):
        return np.unpackbits(z[k], axis=1)[:, :N].reshape(B, 1, H, W).astype(np.float32)

    def onehot(k):
        x = np.zeros((B, N), np.float32)
        x[np.arange(B), z[k]] = 1
        return x.reshape(B, 1, H, W)

    return z, bits("obst_bits"), onehot("start_idx"), onehot("goal_idx"), bits("opt_bits")


def test_training_curve_matches_reference():
    from types import SimpleNamespace

    from neural_astar.planner import NeuralAstar
    from neural_astar.utils.training import PlannerModule

    z, maps, starts, goals, opts = _batch()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # fp32 convolutions, like the CPU reference
    try:
        torch.manual_seed(1234)
        planner = NeuralAstar(encoder_input="m+", encoder_arch="CNN", encoder_depth=4, Tmax=0.25)
        init_abs = float(sum(p.detach().abs().sum() for p in planner.parameters()))
        assert init_abs == pytest.approx(float(z["init_abs"]), rel=1e-6), "same seed must give the reference's init"
        module = PlannerModule(planner, SimpleNamespace(params=SimpleNamespace(lr=1e-3))).cuda()
        opt = module.configure_optimizers()
        assert isinstance(opt, torch.optim.RMSprop)
        batch = [torch.from_numpy(x).cuda() for x in (maps, starts, goals, opts)]
        module.train()
        losses, hist_sums = [], []
        for step in range(len(z["losses"])):
            opt.zero_grad()
            loss = module.training_step(batch, step)
            loss.backward()
            opt.step()
            losses.append(float(loss))
            with torch.no_grad():
                hist_sums.append(float(module(batch[0], batch[1], batch[2]).histories.sum()))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ref = z["losses"]
    assert losses[0] == pytest.approx(float(ref[0]), rel=2e-3), (losses, ref)   # same init -> same first loss
    np.testing.assert_allclose(losses, ref, rtol=2e-2)
    final_abs = float(sum(p.detach().abs().sum() for p in planner.parameters()))
    assert final_abs == pytest.approx(float(z["final_abs"]), rel=1e-3)


def test_graphed_train_step_matches_eager():
    """utils/training.GraphedTrainStep replays the whole training step (forward, L1 loss, search + encoder backward,
    RMSprop) as one CUDA graph: same losses and parameters as the eager loop from the same initialisation, and
    constructing it does not change the model."""
    from types import SimpleNamespace

    from neural_astar.planner import NeuralAstar
    from neural_astar.utils.training import GraphedTrainStep, PlannerModule

    z, maps, starts, goals, opts = _batch()
    batch = [torch.from_numpy(x).cuda() for x in (maps, starts, goals, opts)]

    def make():
        torch.manual_seed(1234)
        planner = NeuralAstar(encoder_input="m+", encoder_arch="CNN", encoder_depth=4, Tmax=0.25)
        return PlannerModule(planner, SimpleNamespace(params=SimpleNamespace(lr=1e-3))).cuda().train()

    eager = make()
    opt = eager.configure_optimizers()
    eager_losses = []
    for step in range(3):
        opt.zero_grad()
        loss = eager.training_step(batch, step)
        loss.backward()
        opt.step()
        eager_losses.append(float(loss))

    graphed = make()
    before = float(sum(p.detach().abs().sum() for p in graphed.planner.parameters()))
    step_fn = GraphedTrainStep(graphed, batch)
    after = float(sum(p.detach().abs().sum() for p in graphed.planner.parameters()))
    assert after == before, "constructing the graphed step must leave the parameters untouched"
    graph_losses = [float(step_fn(batch)) for _ in range(3)]
    assert step_fn.replays == 3
    np.testing.assert_allclose(graph_losses, eager_losses, rtol=2e-3)
    pe = float(sum(p.detach().abs().sum() for p in eager.planner.parameters()))
    pg = float(sum(p.detach().abs().sum() for p in graphed.planner.parameters()))
    assert pg == pytest.approx(pe, rel=1e-3)
    # and the curve is the reference's (tests/golden/train_curve_mazes032.npz, fp32 convs there: 2 % like above)
    np.testing.assert_allclose(graph_losses, z["losses"][:3], rtol=3e-2)


def test_validation_metrics_on_device():
    """validation_step logs p_opt / p_exp / h_mean (utils/training.py:63-87) without leaving the GPU."""
    from types import SimpleNamespace

    from neural_astar.planner import VanillaAstar
    from neural_astar.utils.training import PlannerModule, planner_metrics

    z, maps, starts, goals, opts = _batch()
    batch = [torch.from_numpy(x[:32]).cuda() for x in (maps, starts, goals, opts)]
    module = PlannerModule(VanillaAstar(), SimpleNamespace(params=SimpleNamespace(lr=1e-3))).cuda().eval()
    with torch.no_grad():
        module.validation_step(batch, 0)
    if hasattr(module, "logged"):
        lg = module.logged
        assert lg["metrics/p_opt"] == 1.0 and lg["metrics/p_exp"] == 0.0   # planner == vanilla baseline
        assert lg["metrics/h_mean"] == pytest.approx(0.0, abs=1e-6)
    va = VanillaAstar().cuda()(batch[0], batch[1], batch[2])
    # vanilla A* finds optimal-length paths on every map (path length == optimal)
    assert torch.equal(va.paths.sum((1, 2, 3)), (batch[3].sum((1, 2, 3)) + 1).long())
    assert planner_metrics(va, va) == (1.0, 0.0, pytest.approx(0.0, abs=1e-6))
