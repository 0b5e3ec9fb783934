"""CUDA-graph replay and encoder/search pipelining of a planner's inference forward for fixed batch shapes.

An eager `planner(map_designs, start_maps, goal_maps)` issues about ten launches from Python (pack kernel, encoder
convs, head GEMM, search kernel); at 32x32 / batch 100 the GPU work is ~0.3 ms, so the step is close to
CPU-launch bound and slows down further when several ranks share a host.

    fast = GraphedPlanner(planner, map_designs, start_maps, goal_maps)   # example batch fixes shapes/dtypes
    out = fast(map_designs, start_maps, goal_maps)                        # AstarOutput, caller-owned tensors
    out = fast.replay_host()      # inputs read from fast.host_inputs (pinned), outputs land in fast.host_outputs:
                                  # H2D copies, forward and D2H copies are ONE graph launch

    pipe = PipelinedPlanner(planner, map_designs, start_maps, goal_maps)
    for batch in batches:                      # search(k) overlaps encoder(k+1) inside one graph launch per step
        prev = pipe.submit(*batch)             # AstarOutput of the PREVIOUS batch (None for the first)
    last = pipe.drain()

Inference only (eval mode, no autograd, `store_intermediate_results=False`, `g_ratio >= 0.5` or batch 1).
"""
from __future__ import annotations

from typing import Optional

import torch

from ..planner.differentiable_astar import AstarOutput


def _check_planner(planner, map_designs):
    if planner.training:
        raise ValueError("graph capture covers the inference forward: call planner.eval() first")
    if getattr(planner, "g_ratio", 0.5) < 0.5 and map_designs.shape[0] > 1:
        raise ValueError("g_ratio < 0.5 needs a host decision per batch (batch-coupled stop) and cannot be captured")
    dev = next((p.device for p in planner.parameters()), torch.device("cuda", torch.cuda.current_device()))
    if dev.type != "cuda":
        raise ValueError("the planner must live on a CUDA device")
    return dev


def _warm(planner, inputs, dev, warmup):
    """cuDNN autotuning, plan building (one host sync) and lazy one-time initialisation, outside capture."""
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(max(1, warmup)):
            planner(*inputs)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)


class GraphedPlanner:
    def __init__(self, planner: torch.nn.Module, map_designs: torch.Tensor, start_maps: torch.Tensor,
                 goal_maps: torch.Tensor, device=None, warmup: int = 3):
        dev = _check_planner(planner, map_designs) if device is None else torch.device(device)
        self.planner = planner
        self.device = dev
        examples = (map_designs, start_maps, goal_maps)
        self._in = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in examples)
        for dst, src in zip(self._in, examples):
            dst.copy_(src)
        _warm(planner, self._in, dev, warmup)
        from .. import _native

        before = _native.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            out = planner(*self._in)
        self._out = out
        # kernels of libnastar_b200 recorded in the graph: each replay launches them again without going through
        # the library's host entry points (so nastar_b200_launch_count() does not see replays)
        self.native_launches_per_replay = _native.launch_count() - before
        self.replays = 0
        # end-to-end variant: pinned host buffers -> device -> forward -> pinned host buffers, one graph
        self._host_in = tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in examples)
        for dst, src in zip(self._host_in, examples):
            dst.copy_(src)
        self._host_out = (torch.empty(out.histories.shape, dtype=out.histories.dtype).pin_memory(),
                          torch.empty(out.paths.shape, dtype=out.paths.dtype).pin_memory())
        self._graph_host: Optional[torch.cuda.CUDAGraph] = None
        self._out_host: Optional[AstarOutput] = None

    @property
    def static_inputs(self):
        """(map_designs, start_maps, goal_maps) buffers the graph reads; fill them and call replay() to skip copies."""
        return self._in

    @property
    def static_outputs(self) -> AstarOutput:
        """Outputs of the last replay, owned by the graph (overwritten by the next replay)."""
        return self._out

    @property
    def host_inputs(self):
        """Pinned host staging buffers read by replay_host() (write the next batch here)."""
        return self._host_in

    @property
    def host_outputs(self):
        """(histories, paths) pinned host buffers written by replay_host(); valid after a stream/event sync."""
        return self._host_out

    def replay(self) -> AstarOutput:
        self.graph.replay()
        self.replays += 1
        return self._out

    def replay_host(self) -> AstarOutput:
        """One graph launch: H2D of host_inputs, the forward, D2H into host_outputs.  Returns the device-side
        outputs (graph-owned); the host copies complete in stream order."""
        if self._graph_host is None:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                ins = tuple(torch.empty_like(d) for d in self._in)
                for d, h in zip(ins, self._host_in):
                    d.copy_(h, non_blocking=True)
                out = self.planner(*ins)
                self._host_out[0].copy_(out.histories, non_blocking=True)
                self._host_out[1].copy_(out.paths, non_blocking=True)
            self._graph_host, self._out_host = g, out
        self._graph_host.replay()
        self.replays += 1
        return self._out_host

    def __call__(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor,
                 store_intermediate_results: bool = False) -> AstarOutput:
        if store_intermediate_results:
            raise ValueError("store_intermediate_results needs a host sync per call; use the eager planner")
        for dst, src in zip(self._in, (map_designs, start_maps, goal_maps)):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError(f"GraphedPlanner was captured for {tuple(dst.shape)} {dst.dtype}, got {tuple(src.shape)} {src.dtype}")
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return AstarOutput(self._out.histories.clone(), self._out.paths.clone(), [])


class PipelinedPlanner:
    """Software pipeline over consecutive batches of a NeuralAstar: while the search kernel of batch k runs on a
    hundred warps, the encoder convolutions of batch k+1 occupy the tensor cores.  Both are branches of one CUDA
    graph (captured with a forked stream), double-buffered on inputs, encoder output and results; a step is ONE
    graph launch.  Outputs are identical to the eager call's (same kernels, same order per batch).

    `host=True` builds the end-to-end variant: every step's graph also contains the H2D copy of that batch from
    `host_inputs[k % 2]` (pinned) and the D2H copy of the finished batch into `host_outputs[k % 2]`.
    """

    def __init__(self, planner: torch.nn.Module, map_designs: torch.Tensor, start_maps: torch.Tensor,
                 goal_maps: torch.Tensor, host: bool = False, warmup: int = 3):
        from ..planner.astar import NeuralAstar

        if not isinstance(planner, NeuralAstar):
            raise TypeError("PipelinedPlanner overlaps an encoder with the search: it needs a NeuralAstar")
        dev = _check_planner(planner, map_designs)
        self.planner, self.device, self.host = planner, dev, host
        examples = (map_designs, start_maps, goal_maps)
        # when the three inputs share shape and dtype (the "m+" planners) they live in ONE stacked buffer per
        # parity, so a whole batch can arrive with a single copy (submit_stacked)
        self._stacked = None
        if all(t.shape == map_designs.shape and t.dtype == map_designs.dtype for t in examples):
            self._stacked = [torch.empty((3,) + tuple(map_designs.shape), dtype=map_designs.dtype, device=dev) for _ in range(2)]
            self._in = [tuple(st[i] for i in range(3)) for st in self._stacked]
        else:
            self._in = [tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in examples) for _ in range(2)]
        for buf in self._in:
            for dst, src in zip(buf, examples):
                dst.copy_(src)
        _warm(planner, self._in[0], dev, warmup)
        with torch.no_grad():
            head = self._encode(self._in[0])
        if head is None:
            raise ValueError("this planner/shape has no fused encoder hand-off (needs a conv head and a grid <= 32x32)")
        taps, self._bias, self._scale = head
        self._taps = [torch.empty_like(taps) for _ in range(2)]
        self._host_in = self._host_out = None
        if host:
            self._host_in = [tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in examples) for _ in range(2)]
            for buf in self._host_in:
                for dst, src in zip(buf, examples):
                    dst.copy_(src)
            B, _, H, W = start_maps.shape
            self._host_out = [(torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory(),
                               torch.empty((B, 1, H, W), dtype=torch.int64).pin_memory()) for _ in range(2)]
        self._side = torch.cuda.Stream(device=dev)
        self._graphs = {}
        self._outs = {}
        self._k = 0            # batches submitted so far
        self.replays = 0
        from .. import _native

        self._native = _native
        self.native_launches = 0   # library kernels launched by graph replays (not seen by launch_count())
        self._per_graph = {}

    # -- building blocks ---------------------------------------------------------------------------------
    def _encode(self, ins, out=None):
        p = self.planner
        return p._head_taps(*ins, out=out)

    def _search(self, par):
        p = self.planner
        ins = self._in[par]
        passable = torch.ones_like(ins[1]) if p.learn_obstacles else ins[0]
        return p.astar.search_from_taps(self._taps[par], self._bias, self._scale, ins[1], ins[2], passable)

    def _graph(self, kind: str, par: int):
        """kind: 'enc' (first batch), 'full' (search of batch k-1 on parity 1-par  ||  encoder of batch k on par),
        'search' (last batch, parity par)."""
        key = (kind, par)
        if key in self._graphs:
            return self._graphs[key]
        torch.cuda.synchronize(self.device)
        before = self._native.launch_count()
        g = torch.cuda.CUDAGraph()
        out = None
        with torch.cuda.graph(g), torch.no_grad():
            main = torch.cuda.current_stream(self.device)
            if kind in ("full", "search"):
                spar = (1 - par) if kind == "full" else par
                self._side.wait_stream(main)                       # fork
                with torch.cuda.stream(self._side):
                    out = self._search(spar)
                    if self.host:
                        self._host_out[spar][0].copy_(out.histories, non_blocking=True)
                        self._host_out[spar][1].copy_(out.paths, non_blocking=True)
            if kind in ("full", "enc"):
                if self.host:
                    for d, h in zip(self._in[par], self._host_in[par]):
                        d.copy_(h, non_blocking=True)
                self._encode(self._in[par], out=self._taps[par])
            if kind in ("full", "search"):
                main.wait_stream(self._side)                       # join
        self._graphs[key] = g
        self._outs[key] = out
        self._per_graph[key] = self._native.launch_count() - before
        return g

    # -- public API --------------------------------------------------------------------------------------
    @property
    def host_inputs(self):
        """[(map_designs, start_maps, goal_maps)] x 2 pinned staging buffers; batch k is read from index k % 2."""
        return self._host_in

    @property
    def host_outputs(self):
        """[(histories, paths)] x 2 pinned result buffers; batch k lands in index k % 2."""
        return self._host_out

    def submit(self, map_designs: Optional[torch.Tensor] = None, start_maps: Optional[torch.Tensor] = None,
               goal_maps: Optional[torch.Tensor] = None) -> Optional[AstarOutput]:
        """Enqueue batch k (device tensors; omit them with host=True, the batch is then read from
        host_inputs[k % 2]).  Returns the graph-owned outputs of batch k-1 (overwritten two submits later)."""
        par = self._k & 1
        if not self.host:
            for dst, src in zip(self._in[par], (map_designs, start_maps, goal_maps)):
                dst.copy_(src, non_blocking=True)
        kind = "enc" if self._k == 0 else "full"
        g = self._graph(kind, par)
        g.replay()
        self.replays += 1
        self.native_launches += self._per_graph[(kind, par)]
        self._k += 1
        return self._outs[(kind, par)]

    def submit_stacked(self, batch: torch.Tensor) -> Optional[AstarOutput]:
        """submit() for a batch delivered as one [3, B, 1, H, W] tensor (map_designs, start_maps, goal_maps stacked):
        a single device copy instead of three."""
        if self._stacked is None or self.host:
            raise ValueError("submit_stacked needs equally shaped inputs and a device-input pipeline")
        par = self._k & 1
        self._stacked[par].copy_(batch, non_blocking=True)
        kind = "enc" if self._k == 0 else "full"
        self._graph(kind, par).replay()
        self.replays += 1
        self.native_launches += self._per_graph[(kind, par)]
        self._k += 1
        return self._outs[(kind, par)]

    def drain(self) -> Optional[AstarOutput]:
        """Finish the last submitted batch; the pipeline is empty afterwards."""
        if self._k == 0:
            return None
        par = (self._k - 1) & 1
        g = self._graph("search", par)
        g.replay()
        self.replays += 1
        self.native_launches += self._per_graph[("search", par)]
        self._k = 0
        return self._outs[("search", par)]

    def prepare(self) -> None:
        """Capture every graph variant now (otherwise captured lazily on first use, which synchronises)."""
        self._graph("enc", 0)
        for par in (0, 1):
            self._graph("full", par)
            self._graph("search", par)
