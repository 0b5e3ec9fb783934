import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from oracle import oracle
maps, start, goal, _ = bench.load_problem()
planner = bench.load_planner(torch.device("cpu"))
print("cpu_count", os.cpu_count(), "torch threads", torch.get_num_threads(), "omp max", oracle.max_threads())
for n in (4, 8, 16, 32, 64, 128):
    if n > (os.cpu_count() or 1): break
    torch.set_num_threads(n); oracle.set_threads(n)
    with torch.no_grad():
        planner.encode(torch.from_numpy(maps), torch.from_numpy(start), torch.from_numpy(goal))
        t0 = time.perf_counter()
        for _ in range(3): cost = planner.encode(torch.from_numpy(maps), torch.from_numpy(start), torch.from_numpy(goal)).numpy()
        te = (time.perf_counter() - t0) / 3
    oracle.forward(cost, start, goal, maps, mode="literal")
    t0 = time.perf_counter()
    for _ in range(3): oracle.forward(cost, start, goal, maps, mode="literal")
    to = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for _ in range(3): oracle.forward(cost, start, goal, maps, mode="spec")
    ts = (time.perf_counter() - t0) / 3
    print(f"threads {n:4d}: encoder {te*1e3:8.1f} ms  literal {to*1e3:8.1f} ms  spec {ts*1e3:8.1f} ms")
