"""Developer probe: engine 5 (bin16) vs the generic engine, many Config-5 maps, repeated; any difference is printed.
Usage: python tools/stress_bin16.py [n_maps] [repeats] [H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tools")]
import numpy as np, torch
from neural_astar import _native
from c5_data import c5_maps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
obst, start, goal = c5_maps(n, H, H, 4321)
o, s, g = (torch.from_numpy(x).cuda() for x in (obst, start, goal))
ref = _native.forward(o.clone(), s, g, o, 0.5, H * H)       # generic engine (cost pointer differs)
torch.cuda.synchronize()
bad = 0
for r in range(reps):
    # garbage in shared memory / different allocator state between repeats
    junk = torch.randn(1 << 20, device="cuda").sort()[0]
    a = _native.forward(o, s, g, o, 0.5, H * H)
    torch.cuda.synchronize()
    for k, nm in enumerate(("hist", "paths", "t_solve", "n_steps")):
        if not torch.equal(a[k], ref[k]):
            d = (a[k] != ref[k]).reshape(n, -1).any(1).nonzero().flatten().tolist()
            bad += 1
            print(f"rep {r}: {nm} differs on maps {d[:10]} (steps {ref[3][d[:10]].tolist()})")
print(f"stress_bin16: {n} maps {H}x{H} x {reps} repeats, mismatching outputs: {bad}; steps max {int(ref[3].max())}")
