"""`pq_astar` entry point kept for API compatibility with
/root/reference/src/neural_astar/planner/pq_astar.py:75-104.

The reference runs a per-sample heap A* on CPU NumPy arrays here.  This build has no CPU search: the call is
served by the same GPU engine, and CPU tensors are moved to the current CUDA device and the result returned on
the inputs' device ("CPU in, CPU out" like the reference function).

What is and is not preserved (be aware before relying on `use_differentiable_astar=False`):

* **Uniform costs (VanillaAstar: cost == map design, every passable cell costs 1)** — identical `histories` and
  `paths` to the reference's heap A*; this is the only case the reference itself cross-checks
  (/root/reference/tests/astar_test.py:33-42, mirrored in tests/test_gpu_module_api.py).
* **Non-uniform (learned) costs** — NOT the reference's heap-A* convention.  The reference's `solve_single`
  (/root/reference/src/neural_astar/planner/pq_astar.py:107-162) charges the cost of the NEIGHBOUR being entered
  (`f_new = f_sel - (1-g)h(sel) + g*cost[nei] + (1-g)h(nei)`), uses the heuristic WITHOUT the cost term, treats a
  cell as passable iff `map == 1`, breaks ties in `pqdict` heap order, and returns all-zero maps when the goal is
  unreachable.  The engine implements DifferentiableAstar's convention instead (cost of the SELECTED node,
  `h + cost`, ties by flat index; differentiable_astar.py:191-192,234).  With learned costs the two expand
  different nodes.  A one-time warning is issued when a non-binary cost map arrives through this entry point.
  Rebuilding `pqdict`'s heap-order tie-breaking on a GPU is out of scope (pq_astar is not
  the north-star path).
"""
from __future__ import annotations

import warnings

import torch

from .. import _native
from .differentiable_astar import AstarOutput

_warned = False


def pq_astar(pred_costs, start_maps, goal_maps, map_designs, store_intermediate_results: bool = False,
             g_ratio: float = 0.5) -> AstarOutput:
    global _warned
    assert (
        store_intermediate_results == False  # noqa: E712
    ), "store_intermediate_results = True is currently supported only for differentiable A*"
    src = pred_costs.device
    dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    W = pred_costs.shape[-1]
    if not _warned and map_designs is not pred_costs and src.type == "cpu":
        # cheap only for host tensors; device inputs are not inspected (no sync on the hot path)
        c = pred_costs.detach()
        if bool(((c != 0) & (c != 1)).any()):
            _warned = True
            warnings.warn("pq_astar on the B200 engine follows DifferentiableAstar's cost convention (cost of the "
                          "selected node, h + cost); with non-uniform costs it differs from the reference's heap A* "
                          "— see neural_astar/planner/pq_astar.py", stacklevel=2)
    args = [t.detach().to(dev, torch.float32) for t in (pred_costs, start_maps, goal_maps, map_designs)]
    if map_designs is pred_costs:
        args[3] = args[0]
    H = pred_costs.shape[-2]
    hist, paths, _, _, _ = _native.forward(*args, g_ratio, H * W, False)  # heap A* has no step cap
    return AstarOutput(hist.to(src), paths.to(src))
