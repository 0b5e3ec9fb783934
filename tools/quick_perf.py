"""Developer probe: kernel-only timings of the forward engine on the golden 32x32 inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from golden_util import Golden
from neural_astar import _native

def time_fwd(args, g_ratio, T, iters=20):
    for _ in range(3): _native.forward(*args, g_ratio, T)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); _native.forward(*args, g_ratio, T); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[0], ts[len(ts)//2]

for name in ("mazes032_vanilla_test", "mazes032_neural_test"):
    g = Golden(name)
    for rep in (1, 10, 100, 1000):
        cost = torch.from_numpy(np.tile(g.cost, (rep,1,1,1))).cuda()
        start = torch.from_numpy(np.tile(g.start, (rep,1,1,1))).cuda()
        goal = torch.from_numpy(np.tile(g.goal, (rep,1,1,1))).cuda()
        obst = cost if g.meta.get("vanilla") else torch.from_numpy(np.tile(g.obst, (rep,1,1,1))).cuda()
        B = cost.shape[0]
        best, med = time_fwd((cost, start, goal, obst), 0.5, 1024)
        steps = int(g.z["hist_sum"].sum()) * rep
        nbytes = (24 if g.meta.get("vanilla") else 28) * 1024 * B
        print(f"{name:24s} B={B:7d} best {best*1e3:9.1f} us  med {med*1e3:9.1f} us  maps/s {B/med*1e3:12.0f}  "
              f"exp/s {steps/med*1e3:14.0f}  GB/s {nbytes/med/1e6:8.1f}  ns/step/map(best) {best*1e6/ (g.z['hist_sum'].max()):.1f}" )
