"""Developer probe: engine 5 (bin16) vs the generic engine on the same maps; prints where they differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np, torch
from neural_astar import _native
from golden_util import Golden

names = sys.argv[1:] or ["large128_vanilla", "large256_vanilla"]
for name in names:
    g = Golden(name)
    o, s, gl = (torch.from_numpy(x).cuda() for x in (g.obst, g.start, g.goal))
    a = _native.forward(o, s, gl, o, 0.5, g.W * g.W)               # aliased -> engine 5
    b = _native.forward(o.clone(), s, gl, o, 0.5, g.W * g.W)       # generic
    torch.cuda.synchronize()
    ref = g.bits("hist_bits") != 0
    for m in range(g.B):
        ha, hb = a[0][m, 0].cpu().numpy() != 0, b[0][m, 0].cpu().numpy() != 0
        print(name, "map", m, "steps bin16/generic/ref", int(a[3][m]), int(b[3][m]), int(g.z["hist_sum"][m]),
              "ts", int(a[2][m]), int(b[2][m]), "hist diff vs ref: bin16", int((ha != ref[m, 0]).sum()), "generic", int((hb != ref[m, 0]).sum()))
        d = np.argwhere(ha != ref[m, 0])
        for y, x in d[:6]:
            print("    cell", (int(y), int(x)), "bin16", bool(ha[y, x]), "ref", bool(ref[m, 0][y, x]), "goal", divmod(int(g.z["goal_idx"][m]), g.W))
