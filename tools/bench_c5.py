#!/usr/bin/env python
"""Config 5 (BASELINE.json configs[4]): synthetic 256x256 Moore grids, 1024 maps per GPU, VanillaAstar semantics.

    python tools/bench_c5.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c5.py

Maps: i.i.d. Bernoulli obstacles (p=0.2), start/goal drawn from the largest 8-connected free
component with Chebyshev distance >= 128, generator seed 1234 + rank.  64 distinct maps per rank are tiled to
1024 (the CPU generation of 1024 distinct 256x256 components dominates otherwise; say so when quoting).
Independent shards, no data-path collective; NCCL only reduces (maps, expansions, seconds).
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tools")]
import numpy as np, torch
from neural_astar import _native
from neural_astar.utils.distributed import aggregate_throughput
from scale_probe_lib import c5_maps

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H = W = 256
PER_GPU, DISTINCT = 1024, 64
obst, start, goal = c5_maps(DISTINCT, H, W, 1234 + rank, p_obst=0.2, min_cheb=128)
rep = PER_GPU // DISTINCT
o, s, g = (torch.from_numpy(np.tile(x, (rep, 1, 1, 1))).cuda() for x in (obst, start, goal))
fn = lambda: _native.forward(o, s, g, o, 0.5, W * W)
for _ in range(2): out = fn()
torch.cuda.synchronize()
if world > 1: dist.barrier()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
sec = float(np.median(ts)) * 1e-3
maps, exp, sec = aggregate_throughput(PER_GPU, float(out[0].sum()), sec, device=torch.device("cuda", local))
if rank == 0:
    print(json.dumps({"workload": "C5: 256x256 Moore grids, 1024 maps/GPU (64 distinct x16), p_obst=0.2, cheb>=128",
                      "n_gpus": world, "maps_per_s": maps / sec, "expansions_per_s": exp / sec, "ms": sec * 1e3,
                      "mean_expansions_per_map": exp / maps, "algorithmic_GBps": 24 * H * W * maps / sec / 1e9,
                      "engine": int(_native.lib().nastar_b200_engine_for(H, W))}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
