"""VanillaAstar / NeuralAstar — the nn.Module API of the reference, on the B200 engine.

Mirrors /root/reference/src/neural_astar/planner/astar.py:17-213 (constructor arguments,
attributes `.astar`, `.encoder`, `.encode()`, `.perform_astar()`, `.forward()`), so
scripts/train.py and the notebooks run unchanged.  `use_differentiable_astar=False` selected a
CPU heap A* (`pq_astar`) in the reference; here both settings run the same GPU search (the
reference's own test pins the two to identical outputs, tests/astar_test.py:33-42).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import encoder
from .differentiable_astar import AstarOutput, DifferentiableAstar
from .pq_astar import pq_astar  # noqa: F401  (re-exported like the reference)


class VanillaAstar(nn.Module):
    def __init__(self, g_ratio: float = 0.5, use_differentiable_astar: bool = True):
        """
        Vanilla A* search (reference astar.py:17-46).

        Args:
            g_ratio: ratio between g(v) + h(v). Set 0 to perform as best-first search.
            use_differentiable_astar: kept for API compatibility; both values use the GPU engine.
        """
        super().__init__()
        self.astar = DifferentiableAstar(g_ratio=g_ratio, Tmax=1.0)
        self.g_ratio = g_ratio
        self.use_differentiable_astar = use_differentiable_astar

    def perform_astar(
        self,
        map_designs: torch.Tensor,
        start_maps: torch.Tensor,
        goal_maps: torch.Tensor,
        obstacles_maps: torch.Tensor,
        store_intermediate_results: bool = False,
    ) -> AstarOutput:
        if not self.use_differentiable_astar:
            return pq_astar(map_designs, start_maps, goal_maps, obstacles_maps, store_intermediate_results,
                            g_ratio=self.g_ratio)
        return self.astar(map_designs, start_maps, goal_maps, obstacles_maps, store_intermediate_results)

    def forward(
        self,
        map_designs: torch.Tensor,
        start_maps: torch.Tensor,
        goal_maps: torch.Tensor,
        store_intermediate_results: bool = False,
    ) -> AstarOutput:
        # cost == obstacles == map design (reference astar.py:93-94); the engine detects the alias
        # and stages the plane once.
        return self.perform_astar(map_designs, start_maps, goal_maps, map_designs, store_intermediate_results)


class NeuralAstar(VanillaAstar):
    def __init__(
        self,
        g_ratio: float = 0.5,
        Tmax: float = 1.0,
        encoder_input: str = "m+",
        encoder_arch: str = "CNN",
        encoder_depth: int = 4,
        learn_obstacles: bool = False,
        const: float = None,
        use_differentiable_astar: bool = True,
    ):
        """
        Neural A* search (reference astar.py:105-152).

        Args:
            g_ratio: ratio between g(v) + h(v).
            Tmax: how much of the map the model explores during training (0.25 when training).
            encoder_input: "m+" = map design concatenated with (start + goal); "m" = map only.
            encoder_arch: encoder class name in planner.encoder ("CNN", "CNNDownSize", "Unet").
            encoder_depth: depth of the encoder.
            learn_obstacles: if the obstacles are invisible to the planner.
            const: learnable weight multiplied onto the predicted cost.
            use_differentiable_astar: kept for API compatibility.
        """
        super().__init__()
        self.astar = DifferentiableAstar(g_ratio=g_ratio, Tmax=Tmax)
        self.encoder_input = encoder_input
        encoder_cls = getattr(encoder, encoder_arch)
        self.encoder = encoder_cls(len(self.encoder_input), encoder_depth, const)
        self.learn_obstacles = learn_obstacles
        if self.learn_obstacles:
            print("WARNING: learn_obstacles has been set to True")
        self.g_ratio = g_ratio
        self.use_differentiable_astar = use_differentiable_astar

    def encode(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor) -> torch.Tensor:
        """Predict cost maps (reference astar.py:154-180)."""
        inputs = map_designs
        if "+" in self.encoder_input:
            marks = start_maps + goal_maps
            if map_designs.shape[-1] != start_maps.shape[-1]:
                marks = nn.functional.interpolate(marks, size=map_designs.shape[-2:], mode="nearest")
            inputs = torch.cat((inputs, marks), dim=1)
        return self.encoder(inputs)

    def forward(
        self,
        map_designs: torch.Tensor,
        start_maps: torch.Tensor,
        goal_maps: torch.Tensor,
        store_intermediate_results: bool = False,
    ) -> AstarOutput:
        cost_maps = self.encode(map_designs, start_maps, goal_maps)
        obstacles_maps = map_designs if not self.learn_obstacles else torch.ones_like(start_maps)
        return self.perform_astar(cost_maps, start_maps, goal_maps, obstacles_maps, store_intermediate_results)
