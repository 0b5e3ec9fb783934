"""Config-5 inputs: synthetic 256x256 Moore grids, i.i.d. Bernoulli obstacles (p = 0.2),
start / goal drawn uniformly from the largest 8-connected free component with Chebyshev distance >= 128.
Pure NumPy/SciPy, deterministic in `seed`; used by bench.py (configs.c5) and tools/."""
import numpy as np
from scipy import ndimage


def c5_maps(B, H=256, W=256, seed=1234, p_obst=0.2, min_cheb=None):
    rng = np.random.RandomState(seed)
    min_cheb = min_cheb if min_cheb is not None else min(H, W) // 2
    obst = np.zeros((B, 1, H, W), np.float32)
    start = np.zeros_like(obst)
    goal = np.zeros_like(obst)
    structure = np.ones((3, 3))
    for b in range(B):
        while True:
            m = rng.rand(H, W) > p_obst
            lab, n = ndimage.label(m, structure=structure)
            if n == 0:
                continue
            big = np.argmax(np.bincount(lab.ravel())[1:]) + 1
            cells = np.argwhere(lab == big)
            for _ in range(50):
                s, g = cells[rng.randint(len(cells))], cells[rng.randint(len(cells))]
                if max(abs(s[0] - g[0]), abs(s[1] - g[1])) >= min_cheb:
                    break
            else:
                continue
            break
        obst[b, 0] = m
        start[b, 0, s[0], s[1]] = 1
        goal[b, 0, g[0], g[1]] = 1
    return obst, start, goal
