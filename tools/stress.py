"""Randomised differential test: CUDA engines vs the SPEC oracle over many shapes/seeds (developer tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from neural_astar import _native
from oracle import oracle
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t0 = time.time(); n = 0; cells = 0
shapes = [(32, 32), (12, 12), (31, 17), (1, 9), (9, 1), (33, 33), (64, 64), (50, 64), (64, 50), (40, 33), (33, 64), (64, 33),
          (65, 65), (70, 40), (96, 96), (128, 128), (130, 20), (20, 130)]
while time.time() - t0 < budget:
    H, W = shapes[rng.randint(len(shapes))]
    B = int(rng.randint(1, 9))
    p_obst = rng.choice([0.0, 0.1, 0.25, 0.4])
    obst = (rng.rand(B, 1, H, W) > p_obst).astype(np.float32)
    start = np.zeros_like(obst); goal = np.zeros_like(obst)
    for b in range(B):
        s, g = rng.randint(H * W, size=2)
        start[b, 0].flat[s] = 1; goal[b, 0].flat[g] = 1; obst[b, 0].flat[s] = 1; obst[b, 0].flat[g] = 1
    mode = rng.randint(3)
    if mode == 0: cost, alias = obst, True
    elif mode == 1: cost, alias = (obst * rng.choice([1.0, 0.5, 2.0])).astype(np.float32), False          # many exact ties
    else: cost, alias = (1 / (1 + np.exp(-2 * rng.randn(B, 1, H, W)))).astype(np.float32) * np.float32(rng.choice([1, 10])), False
    g_ratio = float(rng.choice([0.5, 0.5, 0.7, 1.0, 0.2, 0.0]))
    noexit = bool(rng.randint(4) == 0)
    T = int(rng.choice([W * W, max(1, W * W // 4), max(1, W)])) if not noexit else int(min(W * W, rng.randint(1, 200)))
    ref = oracle.forward(cost, start, goal, obst, g_ratio=g_ratio, mode="spec", want_trace=True, T=T, no_early_exit=noexit)
    c = torch.from_numpy(cost).cuda(); o = c if alias else torch.from_numpy(obst).cuda()
    hist, paths, ts, ns, tr = _native.forward(c, torch.from_numpy(start).cuda(), torch.from_numpy(goal).cuda(), o, g_ratio, T, True, noexit)
    ok = (np.array_equal(ts.cpu().numpy(), ref.t_solve) and np.array_equal(ns.cpu().numpy(), ref.n_steps)
          and np.array_equal(tr.cpu().numpy(), ref.trace) and np.array_equal(hist.cpu().numpy(), ref.histories)
          and np.array_equal(paths.cpu().numpy(), ref.paths))
    if not ok:
        print("MISMATCH", H, W, B, p_obst, mode, g_ratio, noexit, T); np.savez("gpurun_out/stress_fail.npz", cost=cost, start=start, goal=goal, obst=obst); sys.exit(1)
    # backward spot check on solvable 32x32-or-smaller / generic
    if rng.randint(6) == 0 and (ref.t_solve >= 0).all() and not noexit:
        G = rng.randn(B, 1, H, W).astype(np.float32)
        Tb = _native.batch_steps(ts, ns, T)
        ts_bwd = ts
        if g_ratio < 0.5:   # goal clamp is not implied by t_solve: derive it from a no-exit trace (header contract)
            Tbi = int(Tb.item())
            full = oracle.forward(cost, start, goal, obst, g_ratio=g_ratio, mode="spec", want_trace=True, T=Tbi, no_early_exit=True)
            gidx = goal.reshape(B, -1).argmax(1)
            clamped = (full.trace == gidx[:, None]).sum(1) >= 2
            ts_bwd = torch.from_numpy(np.where(clamped, 0, Tbi).astype(np.int32)).cuda()
        gc = _native.backward(c, torch.from_numpy(start).cuda(), torch.from_numpy(goal).cuda(), o, torch.from_numpy(G).cuda(), Tb, ts_bwd, g_ratio)
        want = oracle.backward(cost, start, goal, obst, G, int(Tb.item()), g_ratio=g_ratio)
        err = float(np.abs(gc.cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-3 * np.abs(G).max() / np.sqrt(W)))
        absdiff = float(np.abs(gc.cpu().numpy() - want).max())
        if not (err < 1e-5 or absdiff < 1e-6 * float(np.abs(G).max())):   # near-zero gradients: compare absolutely
            print("BWD MISMATCH", H, W, B, g_ratio, T, err); sys.exit(1)
    n += 1; cells += B
print(f"stress ok: {n} problems, {cells} maps in {time.time() - t0:.0f} s")
