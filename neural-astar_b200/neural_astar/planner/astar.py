"""VanillaAstar / NeuralAstar — the reference's nn.Module API on top of the B200 search engine.

API contract mirrored from /root/reference/src/neural_astar/planner/astar.py (:17-102 VanillaAstar,
:105-213 NeuralAstar): constructor keywords, the `.astar` / `.encoder` attributes, `encode()`,
`perform_astar()` and `forward()` returning `AstarOutput`.  `use_differentiable_astar=False` used to
select a CPU heap A* (`pq_astar`); both settings are now served by the same GPU search — the
reference's own test pins the two to identical outputs (tests/astar_test.py:33-42).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoder as _encoders
from .differentiable_astar import AstarOutput, DifferentiableAstar
from .pq_astar import pq_astar  # noqa: F401  re-exported, like the reference module does


class VanillaAstar(nn.Module):
    """A* on the given map: cost == obstacles == map design (reference :93-94).

    >>> out = VanillaAstar()(map_designs, start_maps, goal_maps)   # out.histories, out.paths
    """

    def __init__(self, g_ratio: float = 0.5, use_differentiable_astar: bool = True):
        super().__init__()
        self._configure_search(g_ratio, 1.0, use_differentiable_astar)

    def _configure_search(self, g_ratio: float, Tmax: float, use_differentiable_astar: bool) -> None:
        self.astar = DifferentiableAstar(g_ratio=g_ratio, Tmax=Tmax)
        self.g_ratio = g_ratio
        self.use_differentiable_astar = use_differentiable_astar

    def perform_astar(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor,
                      obstacles_maps: torch.Tensor, store_intermediate_results: bool = False) -> AstarOutput:
        """Positional contract of the seam (reference :48-71): (cost, start, goal, obstacles, store)."""
        if self.use_differentiable_astar:
            return self.astar(map_designs, start_maps, goal_maps, obstacles_maps, store_intermediate_results)
        return pq_astar(map_designs, start_maps, goal_maps, obstacles_maps, store_intermediate_results,
                        g_ratio=self.g_ratio)

    def forward(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor,
                store_intermediate_results: bool = False) -> AstarOutput:
        # the same tensor is passed as cost and as obstacles; the engine notices the alias and reads it once
        return self.perform_astar(map_designs, start_maps, goal_maps, map_designs, store_intermediate_results)


class NeuralAstar(VanillaAstar):
    """Neural A*: an encoder predicts the cost map, the differentiable search runs on it (reference :105-213).

    Args (same names and defaults as the reference):
        g_ratio: weight of g(v) against h(v); 0 = best-first.
        Tmax: fraction of W*W steps explored in training mode (use 0.25 when training).
        encoder_input: "m" = map only, "m+" = map + (start+goal) channel, "rgb+" = RGB + that channel.
        encoder_arch: class name in `planner.encoder` ("CNN", "CNNDownSize", "Unet").
        encoder_depth: number of conv blocks.
        learn_obstacles: hide the obstacle map from the search (all cells passable).
        const: learnable multiplier on the predicted cost (None = 1).
        use_differentiable_astar: kept for compatibility, see module docstring.
    """

    def __init__(self, g_ratio: float = 0.5, Tmax: float = 1.0, encoder_input: str = "m+", encoder_arch: str = "CNN",
                 encoder_depth: int = 4, learn_obstacles: bool = False, const: Optional[float] = None,
                 use_differentiable_astar: bool = True):
        super().__init__()
        self._configure_search(g_ratio, Tmax, use_differentiable_astar)
        self.encoder_input = encoder_input
        self.encoder = getattr(_encoders, encoder_arch)(len(encoder_input), encoder_depth, const)
        self.learn_obstacles = learn_obstacles
        if learn_obstacles:
            print("WARNING: learn_obstacles has been set to True")

    def encode(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor) -> torch.Tensor:
        """Cost maps from the encoder; with a "+" input the start/goal marks ride along as an extra channel,
        nearest-upsampled when the image is larger than the planning grid (reference :154-180)."""
        x = map_designs
        if "+" in self.encoder_input:
            marks = start_maps + goal_maps
            if marks.shape[-1] != x.shape[-1]:
                marks = F.interpolate(marks, size=x.shape[-2:], mode="nearest")
            x = torch.cat((x, marks), dim=1)
        return self.encoder(x)

    def forward(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor,
                store_intermediate_results: bool = False) -> AstarOutput:
        cost_maps = self.encode(map_designs, start_maps, goal_maps)
        passable = torch.ones_like(start_maps) if self.learn_obstacles else map_designs
        return self.perform_astar(cost_maps, start_maps, goal_maps, passable, store_intermediate_results)
