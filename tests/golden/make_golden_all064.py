#!/usr/bin/env python
"""The whole test split of all_064_moore_c16 (400 problems, 64x64) through the reference's VanillaAstar on CPU:
bit-packed inputs and histories / paths.  Exercises the warp-resident 64-wide engine on every test map of the
reference's 64x64 dataset instead of the 12 of all064_vanilla.npz.

    python tests/golden/make_golden_all064.py      # build container only (needs /root/reference)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import torch  # noqa: E402


def main():
    maps, starts, goals, _ = mg.load_batch("all_064_moore_c16.npz", "test", 400)
    starts = starts[:, :1]
    t0 = time.time()
    hist, paths = [], []
    va = mg.VanillaAstar()
    with torch.no_grad():
        for i in range(0, 400, 50):
            out = va(maps[i:i + 50], starts[i:i + 50], goals[i:i + 50])
            hist.append(out.histories.numpy()); paths.append(out.paths.numpy())
    hist, paths = np.concatenate(hist), np.concatenate(paths)
    mg.save("inputs_all064_test400", dict(desc=f"all_064_moore_c16 test split (400 maps, seed-1234 starts) + the reference's "
                                               f"VanillaAstar histories/paths (CPU, {time.time() - t0:.0f} s)", g_ratio=0.5, vanilla=True),
            shape=np.array([400, 64, 64], np.int32), obst_bits=mg.pack(maps[:, 0].numpy()),
            start_idx=mg.onehot_idx(starts.numpy()), goal_idx=mg.onehot_idx(goals.numpy()),
            hist_bits=mg.pack(hist[:, 0]), path_bits=mg.pack(paths[:, 0]), hist_sum=hist.sum((1, 2, 3)).astype(np.int32))
    print("expansions mean/max", hist.sum((1, 2, 3)).mean(), hist.sum((1, 2, 3)).max())


if __name__ == "__main__":
    main()
