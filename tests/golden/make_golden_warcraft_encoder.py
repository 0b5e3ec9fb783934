#!/usr/bin/env python
"""Golden for the Config-4 encoder hand-off: the reference's NeuralAstar.encode with CNNDownSize on a 96x96 RGB batch
(start/goal marks nearest-upsampled from 12x12, planner/astar.py:172-177) — weights, inputs and cost maps.
    python tests/golden/make_golden_warcraft_encoder.py     (build container only)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import make_golden as mg
import torch
from neural_astar.planner import NeuralAstar

torch.manual_seed(4321)
ref = NeuralAstar(encoder_input="rgb+", encoder_arch="CNNDownSize", encoder_depth=3, learn_obstacles=True, const=10.0)
# non-trivial BatchNorm statistics
ref.train()
gen = torch.Generator().manual_seed(1)
for _ in range(2):
    x = torch.rand(4, 3, 96, 96, generator=gen)
    s = torch.zeros(4, 1, 12, 12); s[:, :, 0, 0] = 1
    g = torch.zeros(4, 1, 12, 12); g[:, :, -1, -1] = 1
    ref.encode(x, s, g)
ref.eval()
x = torch.rand(2, 3, 96, 96, generator=gen)
s = torch.zeros(2, 1, 12, 12); s[0, 0, 0, 0] = 1; s[1, 0, 3, 7] = 1
g = torch.zeros(2, 1, 12, 12); g[0, 0, -1, -1] = 1; g[1, 0, 9, 2] = 1
with torch.no_grad():
    cost = ref.encode(x, s, g)
    out = ref(x, s, g)
state = {k: v.numpy() for k, v in ref.state_dict().items()}
np.savez_compressed(os.path.join(mg.OUT, "warcraft_encoder_ckpt.npz"), **{"sd::" + k: v for k, v in state.items()},
                    x=x.numpy().astype(np.float16).astype(np.float32), start=s.numpy(), goal=g.numpy())
# inputs are stored after an fp16 round trip to halve the file: recompute the outputs on exactly those inputs
x2 = torch.from_numpy(x.numpy().astype(np.float16).astype(np.float32))
with torch.no_grad():
    cost = ref.encode(x2, s, g); out = ref(x2, s, g)
z = dict(np.load(os.path.join(mg.OUT, "warcraft_encoder_ckpt.npz")))
z["x"] = x2.numpy().astype(np.float16)
z["cost"] = cost.numpy(); z["hist"] = out.histories.numpy().astype(np.uint8); z["paths"] = out.paths.numpy().astype(np.uint8)
np.savez_compressed(os.path.join(mg.OUT, "warcraft_encoder_ckpt.npz"), **z)
print(os.path.getsize(os.path.join(mg.OUT, "warcraft_encoder_ckpt.npz")) / 1024, "KiB", cost.shape, float(cost.min()), float(cost.max()))
