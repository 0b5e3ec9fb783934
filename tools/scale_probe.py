"""Developer probe: generic-engine timings on Config-4/5-shaped synthetic inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from scipy import ndimage
from neural_astar import _native

def c5_maps(B, H, W, seed, p_obst=0.2, min_cheb=None):
    rng = np.random.RandomState(seed)
    min_cheb = min_cheb if min_cheb is not None else min(H, W) // 2
    obst = np.zeros((B, 1, H, W), np.float32); start = np.zeros_like(obst); goal = np.zeros_like(obst)
    for b in range(B):
        while True:
            m = (rng.rand(H, W) > p_obst)
            lab, n = ndimage.label(m, structure=np.ones((3, 3)))
            if n == 0: continue
            big = np.argmax(np.bincount(lab.ravel())[1:]) + 1
            cells = np.argwhere(lab == big)
            for _ in range(50):
                s, g = cells[rng.randint(len(cells))], cells[rng.randint(len(cells))]
                if max(abs(s[0]-g[0]), abs(s[1]-g[1])) >= min_cheb: break
            else: continue
            break
        obst[b, 0] = m; start[b, 0, s[0], s[1]] = 1; goal[b, 0, g[0], g[1]] = 1
    return obst, start, goal

def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2], out

for (H, W, B) in ((12, 12, 512), (64, 64, 100), (64, 64, 4096), (128, 128, 1024), (256, 256, 256), (256, 256, 1024)):
    obst, start, goal = c5_maps(min(B, 64), H, W, 1234)
    rep = (B + obst.shape[0] - 1) // obst.shape[0]
    o = torch.from_numpy(np.tile(obst, (rep, 1, 1, 1))[:B]).cuda()
    s = torch.from_numpy(np.tile(start, (rep, 1, 1, 1))[:B]).cuda()
    g = torch.from_numpy(np.tile(goal, (rep, 1, 1, 1))[:B]).cuda()
    ms, out = timeit(lambda: _native.forward(o, s, g, o, 0.5, W * W))
    hist, paths, ts, ns, _ = out
    exp = float(hist.sum()); nsmax = int(ns.max()); solved = int((ts >= 0).sum())
    print(f"{H}x{W} B={B:5d} engine={_native.lib().nastar_b200_engine_for(H, W)}: {ms:9.3f} ms  maps/s {B/ms*1e3:12.0f}  "
          f"exp/s {exp/ms*1e3:14.0f}  mean steps {exp/B:8.1f} max {nsmax}  solved {solved}/{B}  "
          f"GB/s {24*H*W*B/ms/1e6:8.1f}  ns/step(max map) {ms*1e6/nsmax:.0f}")
