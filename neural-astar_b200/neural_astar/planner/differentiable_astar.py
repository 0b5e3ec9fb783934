"""Differentiable A* — B200-native engine behind the reference's module surface.

Mirrors /root/reference/src/neural_astar/planner/differentiable_astar.py:
  AstarOutput (:16-23), get_heuristic (:26-52), DifferentiableAstar (:128-267).
The T-step Python loop (:203-252), `expand` (:77-93), `_st_softmax_noexp` (:55-74) and
`backtrack` (:96-125) do not exist here as Python: they run inside one persistent sm_100a kernel
per map batch (csrc/nastar_fwd_*.cuh) reached through the C ABI in include/nastar_b200.h.
Autograd (only cost -> h -> f -> softmax -> histories is live in the reference, :237-243) is a
second kernel evaluating the closed form of SURVEY.md App. B.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional

import torch
import torch.nn as nn

from .. import _native


class AstarOutput(NamedTuple):
    """Output structure of A* search planners (reference :16-23)."""

    histories: torch.Tensor
    paths: torch.Tensor
    intermediate_results: Optional[List[dict]] = None


def get_heuristic(goal_maps: torch.Tensor, tb_factor: float = 0.001) -> torch.Tensor:
    """Chebyshev distance + tb_factor * Euclidean distance to the goal (reference :26-52).

    Kept for API compatibility (`DifferentiableAstar.get_heuristic`); the engine evaluates the same
    expression on-chip (csrc/nastar_common.cuh: heuristic()).
    """
    B, H, W = goal_maps.shape[0], goal_maps.shape[-2], goal_maps.shape[-1]
    flat = goal_maps.reshape(B, -1)
    ys = torch.arange(H, device=goal_maps.device, dtype=goal_maps.dtype).repeat_interleave(W)
    xs = torch.arange(W, device=goal_maps.device, dtype=goal_maps.dtype).repeat(H)
    gy = (flat * ys).sum(-1, keepdim=True)
    gx = (flat * xs).sum(-1, keepdim=True)
    dy, dx = (ys - gy).abs(), (xs - gx).abs()
    cheb = (dy + dx) - torch.minimum(dy, dx)
    euc = torch.sqrt(dy * dy + dx * dx)
    return (cheb + tb_factor * euc).reshape_as(goal_maps)


class _AstarSearch(torch.autograd.Function):
    """forward: libnastar_b200 search; backward: closed-form dL/dcost kernel."""

    @staticmethod
    def forward(ctx, cost_maps, start_maps, goal_maps, obstacles_maps, g_ratio, T, want_trace, goal_clamped=None):
        # goal_clamped is given only on the batch-coupled path (g_ratio < 0.5, see _coupled_steps): the search then
        # runs exactly T steps per map and the backward is told which goals were selected more than once
        no_early_exit = goal_clamped is not None
        hist, paths, t_solve, n_steps, trace = _native.forward(
            cost_maps, start_maps, goal_maps, obstacles_maps, g_ratio, T, want_trace, no_early_exit)
        ctx.g_ratio = g_ratio
        ctx.T = T
        ctx.no_early_exit = no_early_exit
        if no_early_exit:
            # encoding understood by nastar_b200_backward: t_solve < T_batch-1  <=>  gradient blocked at the goal
            t_for_bwd = torch.where(goal_clamped, torch.zeros_like(t_solve), torch.full_like(t_solve, T))
        else:
            t_for_bwd = t_solve
        ctx.save_for_backward(cost_maps, start_maps, goal_maps, obstacles_maps, t_for_bwd, n_steps)
        ctx.mark_non_differentiable(paths, t_solve, n_steps)
        if trace is not None:
            ctx.mark_non_differentiable(trace)
        return hist, paths, t_solve, n_steps, trace

    @staticmethod
    def backward(ctx, grad_hist, *_unused):
        cost, start, goal, obst, t_solve, n_steps = ctx.saved_tensors
        grad_cost = None
        if ctx.needs_input_grad[0]:
            if ctx.no_early_exit:   # the forward already ran exactly the reference's T_batch steps
                T_batch = torch.full((1,), ctx.T, dtype=torch.int32, device=cost.device)
            else:
                T_batch = _native.batch_steps(t_solve, n_steps, ctx.T)
            grad_cost = _native.backward(cost, start, goal, obst, grad_hist.contiguous(), T_batch, t_solve, ctx.g_ratio)
            if grad_cost.shape != cost.shape:  # cost had extra channels: only channel 0 is searched (:177)
                full = torch.zeros_like(cost)
                full[:, :1] = grad_cost
                grad_cost = full
        return grad_cost, None, None, None, None, None, None, None


class DifferentiableAstar(nn.Module):
    def __init__(self, g_ratio: float = 0.5, Tmax: float = 1.0):
        """
        Differentiable A* module (reference :129-148).

        Args:
            g_ratio: ratio between g(v) + h(v). Set 0 to perform as best-first search.
            Tmax: how much of the map the planner explores during training.
        """
        super().__init__()
        # kept so that reference checkpoints load with "All keys matched" (state-dict key
        # `astar.neighbor_filter`); the 3x3 Moore stencil itself is hard-wired in the kernel.
        neighbor_filter = torch.ones(1, 1, 3, 3)
        neighbor_filter[0, 0, 1, 1] = 0
        self.neighbor_filter = nn.Parameter(neighbor_filter, requires_grad=False)
        self.get_heuristic = get_heuristic
        self.g_ratio = g_ratio
        assert (Tmax > 0) & (Tmax <= 1), "Tmax must be within (0, 1]"
        self.Tmax = Tmax

    def num_steps(self, width: int) -> int:
        """Loop bound of the reference (:200-202): int(Tmax_eff * W * W), W = last dim."""
        Tmax = self.Tmax if self.training else 1.0
        return int(Tmax * width * width)

    def forward(
        self,
        cost_maps: torch.Tensor,
        start_maps: torch.Tensor,
        goal_maps: torch.Tensor,
        obstacles_maps: torch.Tensor,
        store_intermediate_results: bool = False,
    ) -> AstarOutput:
        assert cost_maps.ndim == 4
        assert start_maps.ndim == 4
        assert goal_maps.ndim == 4
        assert obstacles_maps.ndim == 4

        T = self.num_steps(cost_maps.shape[-1])
        if T < 1:
            raise ValueError("Tmax * W * W < 1: the reference loop would not execute (:203)")
        g_ratio = float(self.g_ratio)
        coupled = g_ratio < 0.5 and cost_maps.shape[0] > 1
        goal_clamped = None
        if coupled:
            # For g_ratio < 0.5 a solved map does not necessarily keep re-selecting its goal,
            # so the reference's batch-synchronous loop (:251-252) changes the outputs of already-solved maps.
            # Reproduce it: step every map without early exit, find the first step at which ALL maps select
            # their goal, and take the state after exactly that many steps.  (One host sync; the reference
            # synchronises every step.)
            T, goal_clamped = _coupled_steps(cost_maps, start_maps, goal_maps, obstacles_maps, g_ratio, T)
        hist, paths, t_solve, n_steps, trace = _AstarSearch.apply(
            cost_maps, start_maps, goal_maps, obstacles_maps, g_ratio, T, bool(store_intermediate_results), goal_clamped)

        intermediate_results: List[dict] = []
        if store_intermediate_results:
            intermediate_results = _materialise_frames(hist, paths, goal_maps, t_solve, n_steps, trace, T,
                                                       T_batch=T if coupled else None)
        return AstarOutput(hist, paths, intermediate_results)


    def search_from_taps(self, taps: torch.Tensor, bias: float, scale: float, start_maps: torch.Tensor,
                         goal_maps: torch.Tensor, obstacles_maps: torch.Tensor,
                         store_intermediate_results: bool = False) -> AstarOutput:
        """Inference-only entry of the fused encoder hand-off: `taps` [B,H,W,9] are the
        partial products of the encoder's last 3x3 conv; the kernel prologue computes
        cost = sigmoid(bias + gather(taps)) * scale (encoder.py:32-34) and searches on it.  No autograd."""
        W = start_maps.shape[-1]
        T = self.num_steps(W)
        hist, paths, t_solve, n_steps, trace = _native.forward(
            taps, start_maps, goal_maps, obstacles_maps, float(self.g_ratio), T, bool(store_intermediate_results),
            cost_kind=_native.COST_TAPS, cost_scale=scale, cost_bias=bias)
        frames: List[dict] = []
        if store_intermediate_results:
            frames = _materialise_frames(hist, paths, goal_maps, t_solve, n_steps, trace, T)
        return AstarOutput(hist, paths, frames)


def _coupled_steps(cost_maps, start_maps, goal_maps, obstacles_maps, g_ratio: float, T: int):
    """(T_batch, goal_clamped[B]) for g_ratio < 0.5: the number of iterations the reference's loop executes when
    solved maps keep evolving, and per map whether its goal is selected more than once within them — then
    clamp(hist + sel) saw 2 at the goal and blocks its gradient (differentiable_astar.py:222-223)."""
    with torch.no_grad():
        _, _, _, _, trace = _native.forward(cost_maps, start_maps, goal_maps, obstacles_maps, g_ratio, T,
                                            want_trace=True, no_early_exit=True)
        B = goal_maps.shape[0]
        goal_idx = goal_maps[:, 0].reshape(B, -1).argmax(-1).to(trace.dtype)
        at_goal = trace == goal_idx[:, None]
        all_at_goal = at_goal.all(0)
        first = torch.where(all_at_goal.any(), all_at_goal.float().argmax() + 1, torch.tensor(T, device=trace.device))
        T_batch = int(first.item())
        goal_clamped = at_goal[:, :T_batch].sum(1) >= 2
    return T_batch, goal_clamped


def _materialise_frames(hist, paths, goal_maps, t_solve, n_steps, trace, T, T_batch=None) -> List[dict]:
    """Rebuild the reference's per-step frames (:210-216, :257-263) from the selection trace.

    Frame t holds the closed set BEFORE step t and the node selected AT step t; the reference keeps
    stepping every map until the slowest one is solved, re-selecting the goal (App. A.4), so
    post-solve frames repeat the goal.  One host sync (T_batch) is inherent to returning a list.
    """
    B, _, H, W = hist.shape
    N = H * W
    if T_batch is None:
        T_batch = int(_native.batch_steps(t_solve, n_steps, T).item())
    goal_idx = goal_maps[:, 0].reshape(B, -1).argmax(-1)
    tr = trace[:, :T_batch].to(torch.int64)
    tr = torch.where(tr < 0, goal_idx[:, None].expand_as(tr), tr)
    sel = torch.zeros((T_batch, B, N), dtype=hist.dtype, device=hist.device)
    sel.scatter_(2, tr.t().unsqueeze(-1), 1.0)
    closed_before = (sel.cumsum(0) - sel).clamp_(0, 1)
    frames = [
        {"histories": closed_before[t].reshape(B, 1, H, W), "paths": sel[t].reshape(B, 1, H, W)}
        for t in range(T_batch)
    ]
    frames.append({"histories": hist.detach(), "paths": paths.detach()})
    return frames
