"""The reference's own tests (/root/reference/tests/astar_test.py:5-53) re-run against this package on
CUDA tensors, plus API-level parity (intermediate frames, encoder hand-off, larger engines)."""
import numpy as np
import pytest
import torch

from golden_util import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture
def setup():
    # same fixture as the reference (astar_test.py:5-14), on the GPU
    map_designs = torch.ones((8, 1, 64, 64))
    map_designs[:, :, 24:48, 24:48] = 0
    start_maps = torch.zeros((8, 1, 64, 64))
    start_maps[:, :, 0, 0] = 1
    goal_maps = torch.zeros((8, 1, 64, 64))
    goal_maps[:, :, -1, -1] = 1
    return map_designs.cuda(), start_maps.cuda(), goal_maps.cuda()


def test_neural_astar(setup):
    from neural_astar.planner import NeuralAstar

    map_designs, start_maps, goal_maps = setup
    planner = NeuralAstar().cuda()
    output = planner(map_designs, start_maps, goal_maps)
    assert output.histories.shape == (8, 1, 64, 64) and output.paths.dtype == torch.int64
    assert output.intermediate_results == []


def test_vanilla_astar(setup):
    from neural_astar.planner import VanillaAstar

    map_designs, start_maps, goal_maps = setup
    planner = VanillaAstar().cuda()
    output = planner(map_designs, start_maps, goal_maps)
    g = Golden("fixture64_vanilla")  # the reference's output on this fixture: 1169 expansions, path length 88
    np.testing.assert_array_equal(output.histories[:2].cpu().numpy() != 0, g.bits("hist_bits") != 0)
    np.testing.assert_array_equal(output.paths[:2].cpu().numpy() != 0, g.bits("path_bits") != 0)
    assert int(output.histories[0].sum()) == 1169 and int(output.paths[0].sum()) == 88


def test_pq_astar(setup):
    from neural_astar.planner import VanillaAstar

    map_designs, start_maps, goal_maps = setup
    planner = VanillaAstar(use_differentiable_astar=True).cuda()
    output = planner(map_designs, start_maps, goal_maps)
    planner_pq = VanillaAstar(use_differentiable_astar=False)
    output_pq = planner_pq(map_designs.cpu(), start_maps.cpu(), goal_maps.cpu())   # CPU in, CPU out like the reference
    assert not output_pq.histories.is_cuda
    assert torch.allclose(output.histories.cpu(), output_pq.histories)
    assert torch.allclose(output.paths.cpu(), output_pq.paths)


def test_astar_on_rectangle(setup):
    from neural_astar.planner import NeuralAstar, VanillaAstar

    map_designs, start_maps, goal_maps = setup
    map_designs = torch.concat((map_designs, map_designs), -1)
    start_maps = torch.concat((start_maps, torch.zeros_like(start_maps)), -1)
    goal_maps = torch.concat((torch.zeros_like(goal_maps), goal_maps), -1)
    planner = NeuralAstar().cuda()
    output = planner(map_designs, start_maps, goal_maps)
    assert output.histories.shape == (8, 1, 64, 128)
    g = Golden("rect64x128_vanilla")
    out = VanillaAstar().cuda()(map_designs, start_maps, goal_maps)
    np.testing.assert_array_equal(out.histories[:2].cpu().numpy() != 0, g.bits("hist_bits") != 0)
    np.testing.assert_array_equal(out.paths[:2].cpu().numpy() != 0, g.bits("path_bits") != 0)


def test_store_intermediate_results_frames():
    """T_batch+1 frames: frame t = (closed set before step t, node selected at step t); the last frame is
    (histories, paths) (differentiable_astar.py:210-216,257-263)."""
    from neural_astar.planner import VanillaAstar

    g = Golden("mazes032_vanilla_test")
    n = 12
    out = VanillaAstar().cuda()(torch.from_numpy(g.obst[:n]).cuda(), torch.from_numpy(g.start[:n]).cuda(),
                                torch.from_numpy(g.goal[:n]).cuda(), store_intermediate_results=True)
    frames = out.intermediate_results
    hs = g.z["hist_sum"][:n]
    T_batch = int(hs.max())
    assert len(frames) == T_batch + 1
    ref_trace = g.z["trace"][:n]          # reference selections (B=100 run; per-map prefix is batch independent)
    for t in (0, 1, 5, T_batch - 1):
        sel = frames[t]["paths"].reshape(n, -1).argmax(1).cpu().numpy()
        assert frames[t]["paths"].shape == (n, 1, 32, 32) and float(frames[t]["paths"].sum()) == n
        want = np.where(t < hs, ref_trace[:, t], g.z["goal_idx"][:n])   # post-solve: goal re-selected
        np.testing.assert_array_equal(sel, want)
        closed = frames[t]["histories"].reshape(n, -1).sum(1).cpu().numpy()
        np.testing.assert_array_equal(closed, np.minimum(t, hs))
    assert torch.equal(frames[-1]["histories"], out.histories) and torch.equal(frames[-1]["paths"], out.paths)


def test_neural_astar_with_reference_checkpoint_quality():
    """System-level anchor: shipped ckpt on the test split gives p_opt 0.80 / p_exp 0.445 /
    h_mean 0.572 with the CPU reference; the cuDNN encoder (TF32 convs) perturbs costs slightly."""
    import os

    from neural_astar.planner import NeuralAstar, VanillaAstar

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = Golden("mazes032_vanilla_test")
    state = np.load(os.path.join(root, "tests", "golden", "mazes032_ckpt_planner_state.npz"))
    na = NeuralAstar(encoder_arch="CNN")
    na.load_state_dict({k: torch.from_numpy(state[k]) for k in state.files})
    na = na.cuda().eval()
    maps, start, goal = (torch.from_numpy(x).cuda() for x in (g.obst, g.start, g.goal))
    with torch.no_grad():
        out = na(maps, start, goal)
        va = VanillaAstar().cuda()(maps, start, goal)
    pl_a, pl_m = va.paths.sum((1, 2, 3)).cpu().numpy(), out.paths.sum((1, 2, 3)).cpu().numpy()
    ex_a, ex_m = va.histories.sum((1, 2, 3)).cpu().numpy(), out.histories.sum((1, 2, 3)).cpu().numpy()
    p_opt = (pl_a == pl_m).mean()
    p_exp = np.maximum((ex_a - ex_m) / ex_a, 0.0).mean()
    h_mean = 2.0 / (1.0 / (p_opt + 1e-10) + 1.0 / (p_exp + 1e-10))   # utils/training.py:71-85
    assert abs(p_opt - 0.80) <= 0.03 and abs(p_exp - 0.445) <= 0.02 and abs(h_mean - 0.572) <= 0.02
    # with the reference's CPU cost maps fed to the search, outputs are bit-identical to the reference
    gn = Golden("mazes032_neural_test")
    o2 = na.perform_astar(torch.from_numpy(gn.cost).cuda(), start, goal, maps)
    np.testing.assert_array_equal(o2.histories.cpu().numpy() != 0, gn.bits("hist_bits") != 0)
    np.testing.assert_array_equal(o2.paths.cpu().numpy() != 0, gn.bits("path_bits") != 0)


@pytest.mark.parametrize("H,W,B", [(128, 128, 6), (96, 160, 4), (256, 256, 4), (200, 300, 3), (65, 96, 3), (67, 160, 2)])
def test_large_maps_vs_oracle(oracle, H, W, B):
    """Engine 2 (shared-memory state, up to 128x128) and engine 3 (HBM workspace, e.g. 256x256 = Config 5)."""
    from neural_astar import _native

    rng = np.random.RandomState(H + W)
    obst = (rng.rand(B, 1, H, W) > 0.2).astype(np.float32)
    start = np.zeros((B, 1, H, W), np.float32)
    goal = np.zeros((B, 1, H, W), np.float32)
    for b in range(B):
        ys, xs = rng.randint(H // 4), rng.randint(W // 4)
        yg, xg = H - 1 - rng.randint(H // 4), W - 1 - rng.randint(W // 4)
        obst[b, 0, ys, xs] = obst[b, 0, yg, xg] = 1
        start[b, 0, ys, xs] = 1
        goal[b, 0, yg, xg] = 1
    learned = (obst * (0.5 + rng.rand(B, 1, H, W))).astype(np.float32)
    for cost, alias in ((obst, True), (learned, False)):
        ref = oracle.forward(cost, start, goal, obst, mode="spec")
        c = torch.from_numpy(cost).cuda()
        o = c if alias else torch.from_numpy(obst).cuda()
        hist, paths, ts, ns, _ = _native.forward(c, torch.from_numpy(start).cuda(), torch.from_numpy(goal).cuda(), o,
                                                 0.5, W * W)
        np.testing.assert_array_equal(ts.cpu().numpy(), ref.t_solve)
        np.testing.assert_array_equal(ns.cpu().numpy(), ref.n_steps)
        np.testing.assert_array_equal(hist.cpu().numpy(), ref.histories)
        np.testing.assert_array_equal(paths.cpu().numpy(), ref.paths)


@pytest.mark.parametrize("arch,inp,depth,const,shape", [("CNN", "m+", 4, None, (16, 1, 32, 32)),
                                                         ("CNNDownSize", "rgb+", 3, 10.0, (8, 3, 96, 96))])
def test_encoder_eval_fast_path_matches_module(arch, inp, depth, const, shape):
    """Eval-mode cuDNN fast path (folded BN, channels-last, fused bias+ReLU) == nn.Sequential path up to TF32."""
    from neural_astar.planner import NeuralAstar

    torch.manual_seed(3)
    na = NeuralAstar(encoder_input=inp, encoder_arch=arch, encoder_depth=depth, const=const,
                     learn_obstacles=(arch != "CNN")).cuda()
    # make BatchNorm statistics non-trivial
    na.train()
    x = torch.rand(shape, device="cuda")
    hw = 32 if arch == "CNN" else 12
    s = torch.zeros((shape[0], 1, hw, hw), device="cuda"); s[:, :, 0, 0] = 1
    g = torch.zeros_like(s); g[:, :, -1, -1] = 1
    for _ in range(3):
        na.encode(x, s, g)
    na.eval()
    with torch.no_grad():
        fast = na.encode(x, s, g)
    assert na.encoder._plan is not None
    slow = na.encode(x, s, g)           # grad enabled -> reference module path
    assert slow.requires_grad and not fast.requires_grad
    scale = 1.0 if const is None else const
    assert float((fast - slow).abs().max()) < 3e-3 * scale
    # plan is invalidated when weights change
    with torch.no_grad():
        na.encoder.model[0].weight.mul_(1.5)
        fast2 = na.encode(x, s, g)
    assert float((fast2 - fast).abs().max()) > 1e-4


def test_graphed_planner_matches_eager():
    """GraphedPlanner (CUDA-graph replay) returns the same, caller-owned outputs as the eager call."""
    import os

    from neural_astar import _native
    from neural_astar.planner import NeuralAstar, VanillaAstar
    from neural_astar.utils.inference import GraphedPlanner

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = Golden("mazes032_vanilla_test")
    state = np.load(os.path.join(root, "tests", "golden", "mazes032_ckpt_planner_state.npz"))
    na = NeuralAstar(encoder_arch="CNN")
    na.load_state_dict({k: torch.from_numpy(state[k]) for k in state.files})
    na = na.cuda().eval()
    maps, start, goal = (torch.from_numpy(x).cuda() for x in (g.obst, g.start, g.goal))
    with torch.no_grad():
        want = na(maps, start, goal)
    fast = GraphedPlanner(na, maps, start, goal)
    assert fast.native_launches_per_replay == 3     # first layer + head products + the search kernel (TAPS prologue); the rest is cuDNN
    # a different batch of the same shape through the captured graph
    perm = torch.randperm(maps.shape[0], device="cuda")
    with torch.no_grad():
        want2 = na(maps[perm], start[perm], goal[perm])
    got2 = fast(maps[perm], start[perm], goal[perm])
    got = fast(maps, start, goal)
    assert torch.equal(got.histories, want.histories) and torch.equal(got.paths, want.paths)
    assert torch.equal(got2.histories, want2.histories) and torch.equal(got2.paths, want2.paths)   # got2 not overwritten
    assert got.intermediate_results == [] and fast.replays == 2
    # vanilla planner, host inputs straight into the static buffers
    va = GraphedPlanner(VanillaAstar().cuda().eval(), maps, start, goal)
    for dst, src in zip(va.static_inputs, (g.obst, g.start, g.goal)):
        dst.copy_(torch.from_numpy(src).pin_memory(), non_blocking=True)
    out = va.replay()
    np.testing.assert_array_equal(out.histories.cpu().numpy() != 0, g.bits("hist_bits") != 0)
    with pytest.raises(ValueError):
        fast(maps[:5], start[:5], goal[:5])
    with pytest.raises(ValueError):
        GraphedPlanner(NeuralAstar().cuda().train(), maps, start, goal)


def test_warcraft_config_matches_reference_end_to_end():
    """Config 4 shape end to end on the GPU: reference weights, 96x96 RGB -> 12x12 costs -> search with
    learn_obstacles; the reference's CPU histories/paths are reproduced when fed its cost maps, and the cuDNN encoder's
    cost maps agree with the reference's to TF32 accuracy."""
    import os

    from neural_astar.planner import NeuralAstar

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "warcraft_encoder_ckpt.npz"))
    na = NeuralAstar(encoder_input="rgb+", encoder_arch="CNNDownSize", encoder_depth=3, learn_obstacles=True, const=10.0)
    na.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")})
    na = na.cuda().eval()
    x = torch.from_numpy(z["x"].astype(np.float32)).cuda()
    s, g = torch.from_numpy(z["start"]).cuda(), torch.from_numpy(z["goal"]).cuda()
    with torch.no_grad():
        cost = na.encode(x, s, g)
    assert float((cost.cpu() - torch.from_numpy(z["cost"])).abs().max()) < 3e-2          # const=10, TF32 convs
    out = na.perform_astar(torch.from_numpy(z["cost"]).cuda(), s, g, torch.ones_like(s))
    np.testing.assert_array_equal(out.histories.cpu().numpy() != 0, z["hist"] != 0)
    np.testing.assert_array_equal(out.paths.cpu().numpy() != 0, z["paths"] != 0)
