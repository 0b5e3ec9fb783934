"""VanillaAstar / NeuralAstar — the reference's nn.Module API on top of the B200 search engine.

API contract mirrored from /root/reference/src/neural_astar/planner/astar.py (:17-102 VanillaAstar,
:105-213 NeuralAstar): constructor keywords, the `.astar` / `.encoder` attributes, `encode()`,
`perform_astar()` and `forward()` returning `AstarOutput`.  `use_differentiable_astar=False` used to
select a CPU heap A* (`pq_astar`); both settings are now served by the same GPU search — identical for
uniform costs, which is what the reference's own test pins (tests/astar_test.py:33-42); see
planner/pq_astar.py for the cost-convention difference with learned costs.

Inference hand-off: in eval / no-grad mode on CUDA, `NeuralAstar.forward` runs
the input assembly (start+goal add, nearest upsample, concat, NHWC) as one kernel — folded into the first
convolution for the "m+" CNN (`conv1_marks`), a separate `pack_inputs` otherwise -> the encoder's cuDNN convs ->
the head's per-pixel products (`head_taps`) -> the search kernel, whose prologue finishes the encoder (9-tap
gather + bias, sigmoid, * const: encoder.py:32-34) — no ATen glue launches in between.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native
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
        if ("+" in self.encoder_input and self.encoder.fast_path_ok(map_designs) and start_maps.is_cuda
                and start_maps.dtype == torch.float32):
            head = self.encoder.head_taps_marks(map_designs, start_maps, goal_maps)   # same kernels as forward()
            if head is not None:
                return _native.cost_from_taps(*head)
        return self.encoder(self._encoder_input(map_designs, start_maps, goal_maps))

    def _encoder_input(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor) -> torch.Tensor:
        x = map_designs
        if "+" in self.encoder_input:
            if self.encoder.fast_path_ok(x) and start_maps.is_cuda and start_maps.dtype == torch.float32:
                return _native.pack_inputs(x, start_maps, goal_maps)    # one kernel, channels-last result
            marks = start_maps + goal_maps
            if marks.shape[-1] != x.shape[-1]:
                marks = F.interpolate(marks, size=x.shape[-2:], mode="nearest")
            x = torch.cat((x, marks), dim=1)
        return x

    def _head_taps(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor, out=None,
                   on_last_conv=None):
        """Encoder up to the 9-tap products of its head (EncoderBase.head_taps), from the un-assembled inputs: the
        "m+" CNN reads the three planes in its first layer's kernel, everything else packs them first."""
        if "+" in self.encoder_input and start_maps.is_cuda and start_maps.dtype == torch.float32:
            head = self.encoder.head_taps_marks(map_designs, start_maps, goal_maps, out=out, on_last_conv=on_last_conv)
            if head is not None:
                return head
        return self.encoder.head_taps(self._encoder_input(map_designs, start_maps, goal_maps), out=out,
                                      on_last_conv=on_last_conv)

    def forward(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor,
                store_intermediate_results: bool = False) -> AstarOutput:
        passable = torch.ones_like(start_maps) if self.learn_obstacles else map_designs
        fused = self._fused_forward(map_designs, start_maps, goal_maps, passable, store_intermediate_results)
        if fused is not None:
            return fused
        cost_maps = self.encode(map_designs, start_maps, goal_maps)
        return self.perform_astar(cost_maps, start_maps, goal_maps, passable, store_intermediate_results)

    def forward_pair(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor,
                     vanilla_g_ratio: float = 0.5):
        """Validation pair (reference utils/training.py:63-87) in one search launch: returns
        (outputs of this planner, outputs of VanillaAstar on the same batch, (n_closed [2B], path_len [2B]))
        or None when the single-launch form does not apply (autograd on, CPU tensors, learn_obstacles, different
        g_ratio for the baseline, g_ratio < 0.5 batches, pq_astar).  Inference only."""
        B = start_maps.shape[0]
        if (torch.is_grad_enabled() or not start_maps.is_cuda or self.learn_obstacles or map_designs.shape[1] != 1
                or not self.use_differentiable_astar or float(vanilla_g_ratio) != float(self.g_ratio)
                or (float(self.g_ratio) < 0.5 and B > 1) or map_designs.shape[-2:] != start_maps.shape[-2:]
                or map_designs.dtype != torch.float32 or start_maps.dtype != torch.float32):
            return None
        H, W = start_maps.shape[-2], start_maps.shape[-1]
        T = self.astar.num_steps(W)
        kw = {}
        cost = None
        if H <= 64 and W <= 64 and self.encoder.fast_path_ok(map_designs):
            head = self._head_taps(map_designs, start_maps, goal_maps)
            if head is not None:
                cost, bias, scale = head
                kw = dict(cost_kind=_native.COST_TAPS, cost_bias=bias, cost_scale=scale)
        if cost is None:
            cost = self.encode(map_designs, start_maps, goal_maps)
        hist, paths, _, _, _, counts = _native.forward(cost, start_maps, goal_maps, map_designs, float(self.g_ratio), T,
                                                       pair=True, want_counts=True, **kw)
        return (AstarOutput(hist[:B], paths[:B], []), AstarOutput(hist[B:], paths[B:], []), tuple(counts))

    def _fused_forward(self, map_designs, start_maps, goal_maps, passable, store_intermediate_results):
        """Inference fast path: the search kernel consumes the encoder's 9-tap partial products directly
        (NASTAR_COST_TAPS), so no cost plane, sigmoid, bias add or layout conversion is launched.  Returns None
        when it does not apply (training / autograd, CPU tensors, planning grids above 64x64, encoders without a
        single-output-channel conv head, g_ratio < 0.5 batches, pq_astar)."""
        H, W = start_maps.shape[-2], start_maps.shape[-1]
        if (not self.use_differentiable_astar or not self.encoder.fast_path_ok(map_designs) or H > 64 or W > 64
                or not start_maps.is_cuda or (float(self.g_ratio) < 0.5 and start_maps.shape[0] > 1)
                or start_maps.dtype != torch.float32 or passable.dtype != torch.float32):
            return None
        head = self._head_taps(map_designs, start_maps, goal_maps)
        if head is None or head[0].shape[1:3] != (H, W):
            return None
        taps, bias, scale = head
        return self.astar.search_from_taps(taps, bias, scale, start_maps, goal_maps, passable,
                                           store_intermediate_results)
