#!/usr/bin/env python
"""Golden loss curve for Config 3 (scripts/train.py semantics) by RUNNING THE REFERENCE on CPU.

    python tests/golden/make_golden_train.py        # build container only (needs /root/reference)

Reference pieces exercised: NeuralAstar(Tmax=0.25) (planner/astar.py:105-152), the autograd through
DifferentiableAstar.forward (differentiable_astar.py:150-267), L1 loss + RMSprop(lr=1e-3)
(utils/training.py:52-61, scripts/config/train.yaml).  Deterministic variant of the data order: the first
100 maps of the mazes_032 train split (shuffle off), starts drawn after np.random.seed(1234).
"""
import json, os, sys, types, warnings
import numpy as np
warnings.filterwarnings("ignore")
for name, attrs in (("segmentation_models_pytorch", {"Unet": None}), ("pqdict", {"pqdict": dict})):
    m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m
sys.path.insert(0, "/root/reference/src")
import torch
from neural_astar.planner import NeuralAstar
from neural_astar.utils.data import MazeDataset

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
STEPS = 4

np.random.seed(1234); torch.manual_seed(1234)
ds = MazeDataset("/root/reference/planning-datasets/data/mpd/mazes_032_moore_c8.npz", "train")
items = [ds[i] for i in range(100)]
maps, starts, goals, opts = (torch.from_numpy(np.stack([it[k] for it in items])) for k in range(4))
torch.manual_seed(1234)
planner = NeuralAstar(encoder_input="m+", encoder_arch="CNN", encoder_depth=4, Tmax=0.25)
init_abs = float(sum(p.detach().abs().sum() for p in planner.parameters()))
opt = torch.optim.RMSprop(planner.parameters(), 1e-3)
planner.train()
losses, hist_sums = [], []
for step in range(STEPS):
    opt.zero_grad()
    out = planner(maps, starts, goals)
    loss = torch.nn.L1Loss()(out.histories, opts)
    loss.backward()
    opt.step()
    losses.append(float(loss)); hist_sums.append(float(out.histories.sum()))
    print(step, losses[-1], hist_sums[-1], flush=True)
final_abs = float(sum(p.detach().abs().sum() for p in planner.parameters()))
np.savez_compressed(os.path.join(OUT, "train_curve_mazes032.npz"),
                    meta=np.array(json.dumps(dict(desc="reference CPU training, 4 RMSprop steps, B=100 first train maps, Tmax=0.25",
                                                  torch=torch.__version__, seed=1234, lr=1e-3))),
                    obst_bits=np.packbits(maps.numpy().reshape(100, -1) != 0, axis=1),
                    start_idx=starts.numpy().reshape(100, -1).argmax(1).astype(np.int32),
                    goal_idx=goals.numpy().reshape(100, -1).argmax(1).astype(np.int32),
                    opt_bits=np.packbits(opts.numpy().reshape(100, -1) != 0, axis=1),
                    shape=np.array([100, 32, 32], np.int32),
                    losses=np.array(losses), hist_sums=np.array(hist_sums),
                    init_abs=np.float64(init_abs), final_abs=np.float64(final_abs))
print("saved", losses)
