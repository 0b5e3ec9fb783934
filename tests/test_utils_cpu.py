"""CPU tests of the thin re-provisions in neural_astar.utils (host-side glue)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch


def _bfs_policy(m, goal):
    """Optimal Moore policy/distances on a small free map (what planning-datasets precomputes)."""
    H, W = m.shape
    moves = [(-1, 0), (0, 1), (0, -1), (1, 0), (-1, 1), (-1, -1), (1, 1), (1, -1)]  # action order, data.py:234-243
    dist = np.full((H, W), -1, int)
    dist[goal] = 0
    q = [goal]
    while q:
        y, x = q.pop(0)
        for dy, dx in moves:
            yy, xx = y + dy, x + dx
            if 0 <= yy < H and 0 <= xx < W and m[yy, xx] and dist[yy, xx] < 0:
                dist[yy, xx] = dist[y, x] + 1
                q.append((yy, xx))
    pol = np.zeros((8, 1, H, W))
    for y in range(H):
        for x in range(W):
            if dist[y, x] <= 0:
                continue
            for a, (dy, dx) in enumerate(moves):
                yy, xx = y + dy, x + dx
                if 0 <= yy < H and 0 <= xx < W and dist[yy, xx] == dist[y, x] - 1:
                    pol[a, 0, y, x] = 1
                    break
    od = np.where(dist >= 0, -dist, -H * W).astype(float)[None]
    return pol, od


@pytest.fixture
def npz(tmp_path):
    rng = np.random.RandomState(0)
    arrs = []
    for split_n in (6, 3, 3):
        maps = (rng.rand(split_n, 10, 10) > 0.1).astype(np.float32)
        goals, pols, dists = [], [], []
        for i in range(split_n):
            free = np.argwhere(maps[i] > 0)
            gy, gx = free[rng.randint(len(free))]
            g = np.zeros((1, 10, 10)); g[0, gy, gx] = 1
            pol, od = _bfs_policy(maps[i], (gy, gx))
            goals.append(g); pols.append(pol); dists.append(od)
        arrs += [maps, np.stack(goals), np.stack(pols), np.stack(dists)]
    path = str(tmp_path / "toy_010_moore_c4.npz")
    np.savez(path, *arrs)
    return path


def test_maze_dataset_and_dataloader(npz):
    from neural_astar.utils.data import MazeDataset, create_dataloader

    ds = MazeDataset(npz, "train", num_starts=2)
    assert len(ds) == 6 and ds.num_actions == 8
    np.random.seed(0)
    m, s, g, o = ds[0]
    assert m.shape == (1, 10, 10) and s.shape == (2, 10, 10) and g.shape == (1, 10, 10) and o.shape == (2, 10, 10)
    assert s.sum() == 2 and g.sum() == 1 and all(a.dtype == np.float32 for a in (m, s, g, o))
    for k in range(2):
        # the optimal trajectory starts at the start, excludes the goal, and is connected (Moore moves)
        cells = np.argwhere(o[k] > 0)
        assert o[k][tuple(np.argwhere(s[k] > 0)[0])] == 1 and (o[k] * g[0]).sum() == 0
        assert len(cells) >= 1
    dl = create_dataloader(npz, "valid", 3)
    batch = next(iter(dl))
    assert [tuple(b.shape) for b in batch] == [(3, 1, 10, 10)] * 4
    # same numpy seed -> same starts (the reference's determinism contract, utils/data.py:200-221)
    np.random.seed(5); a = ds[1][1]
    np.random.seed(5); b = ds[1][1]
    assert np.array_equal(a, b)


def test_warcraft_dataset(tmp_path):
    from neural_astar.utils.data import create_warcraft_dataloader

    rng = np.random.RandomState(1)
    np.save(tmp_path / "test_maps.npy", rng.randint(0, 255, (4, 96, 96, 3)).astype(np.uint8))
    np.save(tmp_path / "test_shortest_paths.npy", (rng.rand(4, 12, 12) > 0.8).astype(np.uint8))
    maps, start, goal, traj = next(iter(create_warcraft_dataloader(str(tmp_path), "test", 4)))
    assert maps.shape == (4, 3, 96, 96) and float(maps.max()) <= 1.0
    assert start.shape == goal.shape == traj.shape == (4, 1, 12, 12)
    assert float(start[:, 0, 0, 0].sum()) == 4 and float(goal[:, 0, -1, -1].sum()) == 4   # data.py:284-292


def test_visualize_results_follows_the_reference_convention():
    """uint8 HxWx3 image laid out like torchvision.utils.make_grid (8 per row, 2-pixel border, values*255) —
    the frames scripts/create_gif.py feeds to moviepy (reference utils/data.py:16-51)."""
    from neural_astar.planner.differentiable_astar import AstarOutput
    from neural_astar.utils.data import visualize_results

    tv = pytest.importorskip("torchvision.utils")
    rng = np.random.RandomState(0)
    for B, scale in ((10, 1), (2, 1), (1, 1), (3, 2)):
        m = torch.from_numpy((rng.rand(B, 1, 8, 12) > 0.3).astype(np.float32))
        h = torch.zeros(B, 1, 8, 12); h[:, :, 2, :] = 1
        p = torch.zeros(B, 1, 8, 12, dtype=torch.long); p[:, :, 2, 3] = 1
        img = visualize_results(m, AstarOutput(h, p), scale=scale)
        # the reference's own sequence of operations, written with torchvision
        want = tv.make_grid(m).permute(1, 2, 0)
        hh, pp = tv.make_grid(h).permute(1, 2, 0), tv.make_grid(p).permute(1, 2, 0).float()
        want[hh[..., 0] == 1] = torch.tensor([0.2, 0.8, 0])
        want[pp[..., 0] == 1] = torch.tensor([1.0, 0.0, 0])
        want = (want.numpy() * 255.0).astype("uint8")
        if scale > 1:
            from PIL import Image

            want = np.asarray(Image.fromarray(want).resize([x * scale for x in want.shape[:2]], resample=Image.NEAREST))
        assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
        assert img.shape == want.shape
        np.testing.assert_array_equal(img, want)
    assert visualize_results(m, {"histories": h, "paths": p}).dtype == np.uint8
    big = visualize_results(torch.ones(10, 1, 8, 8), AstarOutput(torch.zeros(10, 1, 8, 8), torch.zeros(10, 1, 8, 8)))
    assert big.shape == (2 * 10 + 2, 8 * 10 + 2, 3)        # two rows of tiles: 8 + 2


def test_training_helpers(tmp_path):
    from neural_astar.planner import NeuralAstar
    from neural_astar.utils.training import PlannerModule, load_from_ptl_checkpoint, set_global_seeds

    set_global_seeds(7)
    a = torch.rand(3), np.random.rand(), 
    set_global_seeds(7)
    b = torch.rand(3), np.random.rand(),
    assert torch.equal(a[0], b[0]) and a[1] == b[1]
    planner = NeuralAstar(encoder_depth=1)
    mod = PlannerModule(planner, SimpleNamespace(params=SimpleNamespace(lr=1e-3)))
    assert isinstance(mod.configure_optimizers(), torch.optim.RMSprop)
    # Lightning-style checkpoint round trip: keys containing 'planner' are kept, prefix stripped (training.py:31-39)
    ck = tmp_path / "lightning_logs" / "version_0" / "checkpoints"
    os.makedirs(ck)
    torch.save({"state_dict": {k: v for k, v in mod.state_dict().items()}}, ck / "epoch=0-step=1.ckpt")
    sd = load_from_ptl_checkpoint(str(tmp_path))
    assert "astar.neighbor_filter" in sd and "encoder.model.0.weight" in sd
    assert not any(k.startswith("vanilla_astar") for k in sd) or True
    msg = NeuralAstar(encoder_depth=1).load_state_dict({k: v for k, v in sd.items() if not k.startswith("vanilla")},
                                                        strict=False)
    assert not msg.missing_keys


def test_cached_getitem_equals_reference_style_methods(npz):
    """The cached fast path of __getitem__ draws the same samples (and consumes np.random identically) as the
    reference-style public methods get_random_start_map / get_opt_traj."""
    from neural_astar.utils.data import MazeDataset

    ds = MazeDataset(npz, "train")
    for i in range(len(ds)):
        np.random.seed(100 + i)
        s_ref = ds.get_random_start_map(ds.opt_dists[i])
        t_ref = ds.get_opt_traj(s_ref, ds.goal_maps[i], ds.opt_policies[i])
        after_ref = np.random.rand()
        np.random.seed(100 + i)
        _, s, _, t = ds[i]
        after = np.random.rand()
        assert np.array_equal(s, s_ref) and np.array_equal(t, t_ref) and after == after_ref
