#!/usr/bin/env python
"""Golden vectors for g_ratio < 0.5 with batch > 1 (batch-coupled post-solve steps matter).
Runs the reference's DifferentiableAstar on CPU:  python tests/golden/make_golden_lowg.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import make_golden as mg   # sets up the reference import (stubs + sys.path)
import torch

def main():
    z = np.load(os.path.join(mg.OUT, "mazes032_neural_test.npz"))
    B, H, W = 16, 32, 32
    N = H * W
    maps = torch.from_numpy(np.unpackbits(z["obst_bits"], axis=1)[:B, :N].reshape(B, 1, H, W).astype(np.float32))
    def onehot(idx):
        x = np.zeros((B, N), np.float32); x[np.arange(B), idx[:B]] = 1; return torch.from_numpy(x.reshape(B, 1, H, W))
    starts, goals = onehot(z["start_idx"]), onehot(z["goal_idx"])
    cost = torch.from_numpy(z["cost"][:B].reshape(B, 1, H, W))
    for gr, cst, tag in ((0.0, cost * 10.0, "gr00_cost10"), (0.2, cost, "gr02"), (0.4, cost * 10.0, "gr04_cost10")):
        res = mg.run_search(cst, starts, goals, maps, g_ratio=gr)
        # how different is it from per-map early exit?  (B=1 runs)
        diff = 0
        for b in range(B):
            r1 = mg.run_search(cst[b:b+1], starts[b:b+1], goals[b:b+1], maps[b:b+1], g_ratio=gr)
            diff += int((r1["hist"] != res["hist"][b:b+1]).sum())
        print(tag, "T_batch", res["T_batch"], "hist cells differing from per-map exit:", diff)
        a = mg.common_arrays(maps, starts, goals, res, cost=cst)
        gen = torch.Generator().manual_seed(11)
        G = torch.randn(cst.shape, generator=gen)
        rg = mg.run_search(cst, starts, goals, maps, g_ratio=gr, grad_hist=G)   # reference autograd through the coupled loop
        a.update(rand_G=G[:, 0].numpy(), rand_grad_cost=rg["grad_cost"][:, 0].numpy())
        mg.save(f"mazes032_lowg_{tag}", dict(desc=f"mazes_032 test[:16], learned costs, g_ratio={gr}, batch-coupled (B=16)",
                                             g_ratio=gr, vanilla=False, lowg=True), **a)

if __name__ == "__main__":
    main()
