#!/usr/bin/env python
"""Generate the golden input/output vectors under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference's Python package is imported from /root/reference/src unmodified (nothing is
copied); `segmentation_models_pytorch` and `pqdict` — imported by the reference's package
__init__ chain but never touched by the hot path — are stubbed.  Every file
records torch's version because the oracle inherits ATen CPU semantics
(differentiable_astar.py:207 exp, :68-69 div/max).

Reference entry points exercised:
  neural_astar.planner.VanillaAstar.forward      (planner/astar.py:73-102)
  neural_astar.planner.NeuralAstar.encode        (planner/astar.py:154-180)
  neural_astar.planner.differentiable_astar.DifferentiableAstar.forward (:150-267) + autograd
  neural_astar.utils.data.MazeDataset            (utils/data.py:82-245)
"""
from __future__ import annotations

import json
import os
import re
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
for name, attrs in (("segmentation_models_pytorch", {"Unet": None}), ("pqdict", {"pqdict": dict})):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
REF = "/root/reference"
sys.path.insert(0, f"{REF}/src")

import torch  # noqa: E402
from neural_astar.planner import NeuralAstar, VanillaAstar  # noqa: E402
from neural_astar.planner.differentiable_astar import DifferentiableAstar  # noqa: E402
from neural_astar.utils.data import MazeDataset  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MPD = f"{REF}/planning-datasets/data/mpd"
CKPT = f"{REF}/model/mazes_032_moore_c8/lightning_logs/version_0/checkpoints/epoch=33-step=272.ckpt"
torch.set_num_threads(8)


def pack(x) -> np.ndarray:
    x = np.asarray(x)
    return np.packbits(x.reshape(x.shape[0], -1) != 0, axis=1)


def onehot_idx(x) -> np.ndarray:
    x = np.asarray(x)
    return x.reshape(x.shape[0], -1).argmax(1).astype(np.int32)


def save(name, meta, **arrays):
    meta = dict(meta)
    meta["torch"] = torch.__version__
    meta["reference"] = "omron-sinicx/neural-astar @473edbbd (src/neural_astar/planner/differentiable_astar.py)"
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  {meta.get('desc', '')}")


def run_search(cost, start, goal, obst, g_ratio=0.5, Tmax=1.0, training=False, grad_hist=None, opt_traj=None):
    """Reference DifferentiableAstar forward (+ autograd when a loss is given)."""
    astar = DifferentiableAstar(g_ratio=g_ratio, Tmax=Tmax)
    astar.train(training)
    need_grad = grad_hist is not None or opt_traj is not None
    cost = cost.clone().requires_grad_(need_grad)
    out = astar(cost, start, goal, obst, store_intermediate_results=True)
    T_batch = len(out.intermediate_results) - 1
    trace = torch.stack([fr["paths"].reshape(fr["paths"].shape[0], -1).argmax(1) for fr in out.intermediate_results[:-1]], 1)
    res = dict(hist=out.histories.detach(), paths=out.paths.detach(), T_batch=T_batch, trace=trace.to(torch.int32))
    if need_grad:
        if opt_traj is not None:
            loss = torch.nn.L1Loss()(out.histories, opt_traj)  # utils/training.py:58
        else:
            loss = (out.histories * grad_hist).sum()
        loss.backward()
        res["grad_cost"] = cost.grad.detach()
        res["loss"] = float(loss)
    return res


def common_arrays(maps, start, goal, res, cost=None):
    H, W = maps.shape[-2:]
    a = dict(
        shape=np.array([maps.shape[0], H, W], np.int32),
        obst_bits=pack(maps[:, 0]),
        start_idx=onehot_idx(start),
        goal_idx=onehot_idx(goal),
        hist_bits=pack(res["hist"][:, 0]),
        path_bits=pack(res["paths"][:, 0]),
        hist_sum=res["hist"].sum((1, 2, 3)).numpy().astype(np.int32),
        path_sum=res["paths"].sum((1, 2, 3)).numpy().astype(np.int32),
        T_batch=np.int32(res["T_batch"]),
        trace=res["trace"].numpy(),
    )
    if cost is not None:
        a["cost"] = cost[:, 0].numpy().astype(np.float32)
    return a


def load_batch(npz, split, n, seed=1234):
    np.random.seed(seed)
    torch.manual_seed(seed)
    ds = MazeDataset(f"{MPD}/{npz}", split)
    items = [ds[i] for i in range(n)]
    maps, starts, goals, opts = (torch.from_numpy(np.stack([it[k] for it in items])) for k in range(4))
    return maps, starts, goals, opts


def load_ckpt_planner():
    sd = torch.load(CKPT, weights_only=False, map_location="cpu")["state_dict"]
    ext = {re.split("planner.", k)[-1]: v for k, v in sd.items() if "planner" in k}  # utils/training.py:31-39
    planner = NeuralAstar(encoder_arch="CNN")
    print("ckpt:", planner.load_state_dict(ext))
    planner.eval()
    return planner, ext


def main():
    # --- 1. the reference's own test fixture (tests/astar_test.py:5-14) -------------------
    maps = torch.ones((2, 1, 64, 64))
    maps[:, :, 24:48, 24:48] = 0
    start = torch.zeros((2, 1, 64, 64)); start[:, :, 0, 0] = 1
    goal = torch.zeros((2, 1, 64, 64)); goal[:, :, -1, -1] = 1
    with torch.no_grad():
        out = VanillaAstar()(maps, start, goal)
    res = run_search(maps, start, goal, maps)
    assert torch.equal(out.histories, res["hist"]) and torch.equal(out.paths, res["paths"])
    save("fixture64_vanilla", dict(desc="tests/astar_test.py fixture, VanillaAstar", g_ratio=0.5, vanilla=True),
         **common_arrays(maps, start, goal, res))

    # --- 2. non-square (tests/astar_test.py:45-53), vanilla costs -------------------------
    maps2 = torch.cat((maps, maps), -1)
    start2 = torch.cat((start, torch.zeros_like(start)), -1)
    goal2 = torch.cat((torch.zeros_like(goal), goal), -1)
    res = run_search(maps2, start2, goal2, maps2)
    save("rect64x128_vanilla", dict(desc="64x128 rectangle fixture, VanillaAstar", g_ratio=0.5, vanilla=True),
         **common_arrays(maps2, start2, goal2, res))

    # --- 3. mazes_032 test split, vanilla (Config 1/2 inputs) ----------------
    maps, starts, goals, opts = load_batch("mazes_032_moore_c8.npz", "test", 100)
    res = run_search(maps, starts, goals, maps)
    a = common_arrays(maps, starts, goals, res)
    a["opt_bits"] = pack(opts[:, 0])
    save("mazes032_vanilla_test", dict(desc="mazes_032_moore_c8 test split B=100 seed 1234, VanillaAstar",
                                       g_ratio=0.5, vanilla=True), **a)

    # --- 4. same inputs, learned costs from the shipped checkpoint ------------------------
    planner, ext = load_ckpt_planner()
    with torch.no_grad():
        cost = planner.encode(maps, starts, goals)
        out = planner(maps, starts, goals)
    res = run_search(cost, starts, goals, maps)
    assert torch.equal(out.histories, res["hist"]) and torch.equal(out.paths, res["paths"])
    a = common_arrays(maps, starts, goals, res, cost=cost)
    a["opt_bits"] = pack(opts[:, 0])
    # training-mode gradient: Tmax=0.25 (scripts/config/train.yaml), L1 loss (utils/training.py:58)
    rt = run_search(cost, starts, goals, maps, Tmax=0.25, training=True, opt_traj=opts)
    a.update(train_hist_bits=pack(rt["hist"][:, 0]), train_path_bits=pack(rt["paths"][:, 0]),
             train_T_batch=np.int32(rt["T_batch"]), train_grad_cost=rt["grad_cost"][:, 0].numpy(),
             train_loss=np.float32(rt["loss"]))
    # eval-mode gradient with a dense random upstream gradient
    gen = torch.Generator().manual_seed(7)
    G = torch.randn(cost.shape, generator=gen)
    rg = run_search(cost, starts, goals, maps, grad_hist=G)
    a.update(rand_G=G[:, 0].numpy(), rand_grad_cost=rg["grad_cost"][:, 0].numpy())
    save("mazes032_neural_test", dict(desc="mazes_032 test split B=100, costs from shipped ckpt encoder (CPU)",
                                      g_ratio=0.5, vanilla=False, train_Tmax=0.25), **a)
    np.savez_compressed(os.path.join(OUT, "mazes032_ckpt_planner_state.npz"),
                        **{k: v.numpy() for k, v in ext.items()})
    print("ckpt state:", os.path.getsize(os.path.join(OUT, "mazes032_ckpt_planner_state.npz")) / 1024, "KiB")
    # encoder parity anchor: first 4 cost maps are already in `cost`

    # --- 5. other 32x32 families, 16 test maps each, vanilla ------------------------------
    fam = ["alternating_gaps", "bugtrap_forest", "forest", "gaps_and_forest", "multiple_bugtraps",
           "shifting_gaps", "single_bugtrap"]
    ms, ss, gs = [], [], []
    for i, f in enumerate(fam):
        m_, s_, g_, _ = load_batch(f"{f}_032_moore_c8.npz", "test", 16, seed=100 + i)
        ms.append(m_); ss.append(s_); gs.append(g_)
    maps_f, starts_f, goals_f = torch.cat(ms), torch.cat(ss), torch.cat(gs)
    res = run_search(maps_f, starts_f, goals_f, maps_f)
    save("mpd032_families_vanilla", dict(desc="7 other 32x32 MPD families x16 test maps, VanillaAstar",
                                         g_ratio=0.5, vanilla=True, families=fam),
         **common_arrays(maps_f, starts_f, goals_f, res))

    # --- 6. 64x64 dataset maps, vanilla ----------------------------------------------------
    m64, s64, g64, _ = load_batch("all_064_moore_c16.npz", "test", 12, seed=5)
    res = run_search(m64, s64, g64, m64)
    save("all064_vanilla", dict(desc="all_064_moore_c16 test first 12 maps, VanillaAstar", g_ratio=0.5, vanilla=True),
         **common_arrays(m64, s64, g64, res))

    # --- 7. WarCraft-shaped 12x12 (Config 4): learned-cost range (0,10), no obstacles ------
    gen = torch.Generator().manual_seed(1234)
    B = 32
    cost12 = torch.sigmoid(torch.randn((B, 1, 12, 12), generator=gen) * 2) * 10.0
    ones = torch.ones((B, 1, 12, 12))
    s12 = torch.zeros((B, 1, 12, 12)); s12[:, :, 0, 0] = 1
    g12 = torch.zeros((B, 1, 12, 12)); g12[:, :, -1, -1] = 1
    res = run_search(cost12, s12, g12, ones)
    a = common_arrays(ones, s12, g12, res, cost=cost12)
    G = torch.randn(cost12.shape, generator=gen)
    rt = run_search(cost12, s12, g12, ones, Tmax=0.25, training=True, grad_hist=G)
    a.update(train_hist_bits=pack(rt["hist"][:, 0]), train_path_bits=pack(rt["paths"][:, 0]),
             train_T_batch=np.int32(rt["T_batch"]), rand_G=G[:, 0].numpy(),
             train_grad_cost=rt["grad_cost"][:, 0].numpy())
    rg = run_search(cost12, s12, g12, ones, grad_hist=G)
    a.update(rand_grad_cost=rg["grad_cost"][:, 0].numpy())
    save("warcraft12_synth", dict(desc="12x12 synthetic learned costs in (0,10), learn_obstacles=True, B=32",
                                  g_ratio=0.5, vanilla=False, train_Tmax=0.25), **a)

    # --- 8. edge cases ---------------------------------------------------------------------
    # 8a start == goal
    m8 = torch.ones((1, 1, 8, 8)); s8 = torch.zeros((1, 1, 8, 8)); s8[0, 0, 3, 4] = 1
    res = run_search(m8, s8, s8.clone(), m8)
    save("edge_start_is_goal", dict(desc="8x8, start==goal, B=1", g_ratio=0.5, vanilla=True),
         **common_arrays(m8, s8, s8, res))
    # 8b random-obstacle non-square maps, both orientations, B=1 each so no batch coupling
    rng = np.random.RandomState(3)
    for (H, W) in ((12, 20), (20, 12), (5, 40), (33, 31)):
        cases = []
        while len(cases) < 6:
            m_ = (rng.rand(H, W) > 0.25).astype(np.float32)
            free = np.argwhere(m_ > 0)
            a_, b_ = free[rng.randint(len(free))], free[rng.randint(len(free))]
            mt = torch.from_numpy(m_)[None, None]
            st = torch.zeros_like(mt); st[0, 0, a_[0], a_[1]] = 1
            gt = torch.zeros_like(mt); gt[0, 0, b_[0], b_[1]] = 1
            try:
                r_ = run_search(mt, st, gt, mt)
            except Exception:
                continue  # unreachable goal: the reference raises (NaN -> IndexError)
            if not bool(torch.isfinite(r_["hist"]).all()):
                continue
            cases.append((mt, st, gt, r_))
        # stored as independent B=1 problems (T_batch differs per problem)
        mm = torch.cat([c[0] for c in cases]); sm = torch.cat([c[1] for c in cases]); gm = torch.cat([c[2] for c in cases])
        hist = torch.cat([c[3]["hist"] for c in cases]); paths = torch.cat([c[3]["paths"] for c in cases])
        tb = np.array([c[3]["T_batch"] for c in cases], np.int32)
        save(f"edge_rand_{H}x{W}", dict(desc=f"{H}x{W} random obstacles, 6 independent B=1 problems (T=W*W may cap)",
                                        g_ratio=0.5, vanilla=True, independent=True),
             shape=np.array([len(cases), H, W], np.int32), obst_bits=pack(mm[:, 0]), start_idx=onehot_idx(sm),
             goal_idx=onehot_idx(gm), hist_bits=pack(hist[:, 0]), path_bits=pack(paths[:, 0]),
             hist_sum=hist.sum((1, 2, 3)).numpy().astype(np.int32), path_sum=paths.sum((1, 2, 3)).numpy().astype(np.int32),
             T_batch_each=tb)
    # 8c other g_ratio values (>= 0.5 keeps post-solve invariance, App. A.4), first 16 mazes
    for gr in (0.7, 1.0):
        res = run_search(maps[:16], starts[:16], goals[:16], maps[:16], g_ratio=gr)
        save(f"mazes032_vanilla_gr{int(gr * 10):02d}", dict(desc=f"mazes_032 test[:16], g_ratio={gr}", g_ratio=gr, vanilla=True),
             **common_arrays(maps[:16], starts[:16], goals[:16], res))
    # 8d g_ratio < 0.5 with B=1 problems (no post-solve steps possible)
    hs, ps, tb = [], [], []
    for i in range(8):
        r_ = run_search(cost[i:i + 1], starts[i:i + 1], goals[i:i + 1], maps[i:i + 1], g_ratio=0.2)
        hs.append(r_["hist"]); ps.append(r_["paths"]); tb.append(r_["T_batch"])
    hist, paths = torch.cat(hs), torch.cat(ps)
    save("mazes032_neural_gr02_b1", dict(desc="mazes_032 test[:8] learned costs, g_ratio=0.2, independent B=1",
                                         g_ratio=0.2, vanilla=False, independent=True),
         shape=np.array([8, 32, 32], np.int32), obst_bits=pack(maps[:8, 0]), start_idx=onehot_idx(starts[:8]),
         goal_idx=onehot_idx(goals[:8]), cost=cost[:8, 0].numpy(), hist_bits=pack(hist[:, 0]), path_bits=pack(paths[:, 0]),
         hist_sum=hist.sum((1, 2, 3)).numpy().astype(np.int32), path_sum=paths.sum((1, 2, 3)).numpy().astype(np.int32),
         T_batch_each=np.array(tb, np.int32))


if __name__ == "__main__":
    main()
