#!/usr/bin/env python
"""Adversarial near-tie vectors ("nothing pins the near-tie regime"), run through the reference.

    python tests/golden/make_golden_adversarial.py     # build container only (needs /root/reference)

The engine selects arg-min(f, flat index); the reference selects argmax of fl(exp(-f/sqrt(W))) * open / sum with
first-index ties (differentiable_astar.py:55-74,206-209).  The two differ only when exp/division round two DISTINCT
f values to the same softmax weight.  These 5x8 maps are built so that exactly that happens at step 1: the start S
opens A = (1,2) [lower index] and B = (3,2) with f_A = nextafter(f_B) — one ulp LARGER — and identical rounded
exp.  The reference then expands A (lower index) before B; arg-min expands B first, reaches the goal through
(3,3) and never closes A.  Expected, documented deviation: `histories` of the engine lack that ONE cell; paths are
identical.  (Found by tools-free search over random cB; the constants below are those hits.)  Natural data: 0 such
mask differences in 31 200 dataset problems / 11.6 M selections (profiles/r02_selection_sweep.txt).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import torch  # noqa: E402

CASES = [(np.float32(1.3854055), np.float32(1.3862334)), (np.float32(1.4244161), np.float32(1.4252442)),
         (np.float32(1.2728081), np.float32(1.2736362))]
H, W = 5, 8


def build(cA, cB):
    obst = np.zeros((1, 1, H, W), np.float32)
    cost = np.ones((1, 1, H, W), np.float32)
    start = np.zeros_like(obst)
    goal = np.zeros_like(obst)
    for (y, x) in [(2, 1), (1, 2), (3, 2), (3, 3), (3, 4), (1, 3)]:
        obst[0, 0, y, x] = 1
    start[0, 0, 2, 1] = 1
    goal[0, 0, 3, 4] = 1
    cost[0, 0, 2, 1] = 0.5
    cost[0, 0, 1, 2] = cA
    cost[0, 0, 3, 2] = cB
    cost[0, 0, 3, 3] = 0.01
    cost[0, 0, 3, 4] = 0.01
    cost[0, 0, 1, 3] = 5.0
    return cost, start, goal, obst


def main():
    parts = [build(a, b) for a, b in CASES]
    cost, start, goal, obst = (torch.from_numpy(np.concatenate([p[k] for p in parts])) for k in range(4))
    res = mg.run_search(cost, start, goal, obst)
    arrays = mg.common_arrays(obst, start, goal, res, cost=cost)
    mg.save("adversarial_neartie", dict(desc="5x8 maps with two open cells whose f differ by one ulp but whose rounded "
                                             "exp(-f/sqrt(W)) agree: the reference expands the lower index first",
                                        g_ratio=0.5, neartie=True, extra_cell=[1, 2]), **arrays)
    print("reference trace", res["trace"].tolist(), "closed", res["hist"].sum((1, 2, 3)).tolist())


if __name__ == "__main__":
    main()
