"""ncu target: a handful of B=100 vanilla/neural forward launches on the golden 32x32 inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from golden_util import Golden
from neural_astar import _native
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for name in ("mazes032_vanilla_test", "mazes032_neural_test"):
    g = Golden(name)
    cost = torch.from_numpy(np.tile(g.cost, (rep,1,1,1))).cuda()
    start = torch.from_numpy(np.tile(g.start, (rep,1,1,1))).cuda()
    goal = torch.from_numpy(np.tile(g.goal, (rep,1,1,1))).cuda()
    obst = cost if g.meta.get("vanilla") else torch.from_numpy(np.tile(g.obst, (rep,1,1,1))).cuda()
    for _ in range(3):
        _native.forward(cost, start, goal, obst, 0.5, 1024)
    torch.cuda.synchronize()
