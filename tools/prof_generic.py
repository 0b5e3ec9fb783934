import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np, torch
from neural_astar import _native
from scale_probe_lib import c5_maps
H = W = int(sys.argv[1]); B = int(sys.argv[2])
obst, start, goal = c5_maps(min(B, 64), H, W, 1234)
rep = (B + obst.shape[0] - 1) // obst.shape[0]
o = torch.from_numpy(np.tile(obst, (rep, 1, 1, 1))[:B]).cuda()
s = torch.from_numpy(np.tile(start, (rep, 1, 1, 1))[:B]).cuda()
g = torch.from_numpy(np.tile(goal, (rep, 1, 1, 1))[:B]).cuda()
for _ in range(3):
    _native.forward(o, s, g, o, 0.5, W * W)
torch.cuda.synchronize()
