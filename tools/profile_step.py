#!/usr/bin/env python
"""Profiling driver: a few steady-state steps of the headline forward inside a cudaProfilerStart/Stop window.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/profile_step.py [pipe|graph|eager|c5|search]

Everything before the window (cuDNN autotuning, plan building, graph capture, warm-up) stays out of the capture.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np, torch
from neural_astar import _native
from golden_util import Golden

mode = sys.argv[1] if len(sys.argv) > 1 else "pipe"
if mode == "c5":
    from c5_data import c5_maps
    n = int(os.environ.get("C5_MAPS", "296"))
    o, s, g = (torch.from_numpy(x).cuda() for x in c5_maps(n, 256, 256, 1234))
    for _ in range(2):
        _native.forward(o, s, g, o, 0.5, 65536)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    _native.forward(o, s, g, o, 0.5, 65536)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

from neural_astar.planner import NeuralAstar
from neural_astar.utils.inference import GraphedPlanner, PipelinedPlanner
g_ = Golden("mazes032_vanilla_test")
state = np.load(os.path.join(ROOT, "tests", "golden", "mazes032_ckpt_planner_state.npz"))
na = NeuralAstar(encoder_arch="CNN"); na.load_state_dict({k: torch.from_numpy(state[k]) for k in state.files})
na = na.cuda().eval()
maps, start, goal = (torch.from_numpy(x).cuda() for x in (g_.obst, g_.start, g_.goal))
with torch.no_grad():
    for _ in range(3):
        na(maps, start, goal)
    cost = na.encode(maps, start, goal)
if mode == "search":
    for _ in range(3):
        _native.forward(cost, start, goal, maps, 0.5, 1024)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    _native.forward(cost, start, goal, maps, 0.5, 1024)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
elif mode == "eager":
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    with torch.no_grad():
        for _ in range(2):
            na(maps, start, goal)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
elif mode == "graph":
    fast = GraphedPlanner(na, maps, start, goal)
    for _ in range(3):
        fast.replay()
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    for _ in range(2):
        fast.replay()
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
else:
    pipe = PipelinedPlanner(na, maps, start, goal); pipe.prepare()
    for _ in range(4):
        pipe.submit(maps, start, goal)
    pipe.drain()
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    for _ in range(3):
        pipe.submit(maps, start, goal)
    pipe.drain()
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
