#!/usr/bin/env python
"""All 1000 problems of mazes_032_moore_c8 (800 train + 100 valid + 100 test) as ONE small fixture:
  * inputs (bit-packed maps, start / goal indices; start positions from the reference's MazeDataset,
    /root/reference/src/neural_astar/utils/data.py:152-221, under np.random.seed(1234)) — used by bench.py's
    saturated-throughput measurement ("tile the 1000 available maps ... to fill the GPU");
  * the REFERENCE's end-to-end outputs for them: NeuralAstar with the shipped checkpoint on CPU (encoder + PyTorch
    DifferentiableAstar loop), histories / paths bit-packed — the system-level parity anchor of
    tests/test_gpu_round2.py::test_end_to_end_masks_on_all_1000_maps.

    python tests/golden/make_golden_all1000.py     # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import torch  # noqa: E402


def main():
    np.random.seed(1234)
    torch.manual_seed(1234)
    maps, starts, goals = [], [], []
    for split in ("train", "valid", "test"):
        ds = mg.MazeDataset(f"{mg.MPD}/mazes_032_moore_c8.npz", split)
        for i in range(len(ds)):
            m, s, g, _ = ds[i]
            maps.append(m); starts.append(s[:1]); goals.append(g)
    maps, starts, goals = (np.stack(x) for x in (maps, starts, goals))
    planner, _ = mg.load_ckpt_planner()
    hist, paths = [], []
    with torch.no_grad():
        for i in range(0, len(maps), 100):
            out = planner(*(torch.from_numpy(x[i:i + 100]) for x in (maps, starts, goals)))
            hist.append(out.histories.numpy()); paths.append(out.paths.numpy())
    hist, paths = np.concatenate(hist), np.concatenate(paths)
    mg.save("inputs_mazes032_all1000", dict(desc="mazes_032_moore_c8 train+valid+test (1000 maps), seed-1234 starts; inputs + "
                                                 "the reference's NeuralAstar (shipped ckpt, CPU) histories/paths"),
            shape=np.array([len(maps), 32, 32], np.int32), obst_bits=mg.pack(maps[:, 0]),
            start_idx=mg.onehot_idx(starts), goal_idx=mg.onehot_idx(goals),
            neural_hist_bits=mg.pack(hist[:, 0]), neural_path_bits=mg.pack(paths[:, 0]),
            neural_hist_sum=hist.sum((1, 2, 3)).astype(np.int32))


if __name__ == "__main__":
    main()
