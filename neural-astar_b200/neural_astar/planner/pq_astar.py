"""`pq_astar` entry point kept for API compatibility with
/root/reference/src/neural_astar/planner/pq_astar.py:75-104.

The reference runs a per-sample heap A* on CPU NumPy arrays here.  This build has no CPU search:
the call is served by the same GPU engine (identical histories/paths on the reference's own
cross-check, tests/astar_test.py:33-42).  CPU tensors are moved to the current CUDA device and
the result is returned on the inputs' device, which preserves the reference's "CPU in, CPU out"
behaviour of this function.
"""
from __future__ import annotations

import torch

from .. import _native
from .differentiable_astar import AstarOutput


def pq_astar(pred_costs, start_maps, goal_maps, map_designs, store_intermediate_results: bool = False,
             g_ratio: float = 0.5) -> AstarOutput:
    assert (
        store_intermediate_results == False  # noqa: E712
    ), "store_intermediate_results = True is currently supported only for differentiable A*"
    src = pred_costs.device
    dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    W = pred_costs.shape[-1]
    args = [t.detach().to(dev, torch.float32) for t in (pred_costs, start_maps, goal_maps, map_designs)]
    if map_designs is pred_costs:
        args[3] = args[0]
    H = pred_costs.shape[-2]
    hist, paths, _, _, _ = _native.forward(*args, g_ratio, H * W, False)  # heap A* has no step cap
    return AstarOutput(hist.to(src), paths.to(src))
