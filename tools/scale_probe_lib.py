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

