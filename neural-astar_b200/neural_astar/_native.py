"""ctypes binding of libnastar_b200.so (C ABI declared in include/nastar_b200.h).

PyTorch is used here only for device memory and the current CUDA stream; the library itself
has no torch dependency.  Loading fails loudly when the library is missing: the product has no
CPU or eager-PyTorch fallback for the search.
"""
from __future__ import annotations

import contextlib
import ctypes
import math
import os
from typing import Optional, Tuple

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # neural-astar_b200/
LIB_PATH = os.environ.get("NASTAR_B200_LIB", os.path.join(_PKG_ROOT, "lib", "libnastar_b200.so"))

ABI_VERSION = 2
NASTAR_OK = 0
TS_CAPPED = -1
TS_EXHAUSTED = -2

_f32p = ctypes.POINTER(ctypes.c_float)


class FwdParams(ctypes.Structure):
    """struct nastar_fwd_params (include/nastar_b200.h)"""

    _fields_ = [
        ("cost", ctypes.c_void_p), ("cost_stride", ctypes.c_int64),
        ("start", ctypes.c_void_p), ("start_stride", ctypes.c_int64),
        ("goal", ctypes.c_void_p), ("goal_stride", ctypes.c_int64),
        ("obst", ctypes.c_void_p), ("obst_stride", ctypes.c_int64),
        ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("g_ratio", ctypes.c_float), ("one_minus_g_ratio", ctypes.c_float),
        ("T", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("cost_kind", ctypes.c_int32), ("cost_scale", ctypes.c_float), ("cost_bias", ctypes.c_float),
        ("histories", ctypes.c_void_p), ("paths", ctypes.c_void_p),
        ("t_solve", ctypes.c_void_p), ("n_steps", ctypes.c_void_p), ("trace", ctypes.c_void_p),
        ("n_closed", ctypes.c_void_p), ("path_len", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


class BwdParams(ctypes.Structure):
    """struct nastar_bwd_params (include/nastar_b200.h)"""

    _fields_ = [
        ("cost", ctypes.c_void_p), ("cost_stride", ctypes.c_int64),
        ("start", ctypes.c_void_p), ("start_stride", ctypes.c_int64),
        ("goal", ctypes.c_void_p), ("goal_stride", ctypes.c_int64),
        ("obst", ctypes.c_void_p), ("obst_stride", ctypes.c_int64),
        ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("g_ratio", ctypes.c_float), ("one_minus_g_ratio", ctypes.c_float), ("sqrt_w", ctypes.c_float),
        ("T_batch", ctypes.c_void_p), ("t_solve", ctypes.c_void_p),
        ("grad_histories", ctypes.c_void_p), ("grad_stride", ctypes.c_int64),
        ("grad_cost", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


EXPORTS = (
    "nastar_b200_abi_version",
    "nastar_b200_forward_workspace_bytes",
    "nastar_b200_backward_workspace_bytes",
    "nastar_b200_forward",
    "nastar_b200_backward",
    "nastar_b200_batch_steps",
    "nastar_b200_engine_for",
    "nastar_b200_bin16_supported",
    "nastar_b200_pack_inputs",
    "nastar_b200_cost_from_taps",
    "nastar_b200_head_taps",
    "nastar_b200_conv1_marks",
    "nastar_b200_selftest_sqrt",
    "nastar_b200_launch_count",
    "nastar_b200_status_string",
    "nastar_b200_last_cuda_error",
)

_lib = None
_ws_cache = {}


class NativeLibraryMissing(ImportError):
    pass


def lib():
    """Load the engine. Raises NativeLibraryMissing (never falls back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build the sm_100a engine first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C neural-astar_b200). "
            "neural_astar (B200) has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.nastar_b200_abi_version.restype = ctypes.c_int
    L.nastar_b200_forward_workspace_bytes.argtypes = [ctypes.c_int32] * 3
    L.nastar_b200_forward_workspace_bytes.restype = ctypes.c_size_t
    L.nastar_b200_backward_workspace_bytes.argtypes = [ctypes.c_int32] * 3
    L.nastar_b200_backward_workspace_bytes.restype = ctypes.c_size_t
    L.nastar_b200_forward.argtypes = [ctypes.POINTER(FwdParams), ctypes.c_void_p]
    L.nastar_b200_forward.restype = ctypes.c_int
    L.nastar_b200_backward.argtypes = [ctypes.POINTER(BwdParams), ctypes.c_void_p]
    L.nastar_b200_backward.restype = ctypes.c_int
    L.nastar_b200_batch_steps.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.c_void_p]
    L.nastar_b200_batch_steps.restype = ctypes.c_int
    L.nastar_b200_engine_for.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.nastar_b200_engine_for.restype = ctypes.c_int
    L.nastar_b200_bin16_supported.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.nastar_b200_bin16_supported.restype = ctypes.c_int
    L.nastar_b200_pack_inputs.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                          ctypes.c_void_p]
    L.nastar_b200_pack_inputs.restype = ctypes.c_int
    L.nastar_b200_cost_from_taps.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    L.nastar_b200_cost_from_taps.restype = ctypes.c_int
    L.nastar_b200_head_taps.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]
    L.nastar_b200_head_taps.restype = ctypes.c_int
    L.nastar_b200_conv1_marks.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]
    L.nastar_b200_conv1_marks.restype = ctypes.c_int
    L.nastar_b200_selftest_sqrt.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    L.nastar_b200_selftest_sqrt.restype = ctypes.c_int
    L.nastar_b200_launch_count.restype = ctypes.c_uint64
    L.nastar_b200_status_string.argtypes = [ctypes.c_int]
    L.nastar_b200_status_string.restype = ctypes.c_char_p
    L.nastar_b200_last_cuda_error.restype = ctypes.c_char_p
    if L.nastar_b200_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI {L.nastar_b200_abi_version()} != expected {ABI_VERSION}")
    _lib = L
    return L


def _check(status: int, what: str) -> None:
    if status != NASTAR_OK:
        L = lib()
        msg = L.nastar_b200_status_string(status).decode()
        if status == 3:
            msg += ": " + L.nastar_b200_last_cuda_error().decode()
        raise RuntimeError(f"{what} failed: {msg}")


def host_scalars(g_ratio: float, W: int) -> Tuple[float, float, float]:
    """The Python-double scalar arithmetic of differentiable_astar.py:206-207; ctypes then rounds
    each to fp32 exactly as ATen rounds a Python scalar operand."""
    return float(g_ratio), float(1 - g_ratio), math.sqrt(W)


def _plane(x: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """Return (tensor kept alive, device pointer of channel 0, map stride in elements)."""
    if x.dtype != torch.float32:
        raise TypeError(f"planes must be float32 (got {x.dtype}); the reference path is fp32-only")
    H, W = x.shape[-2], x.shape[-1]
    if x.stride(-1) != 1 or x.stride(-2) != W:
        x = x.contiguous()
    return x, x.data_ptr(), x.stride(0)


def launch_count() -> int:
    return int(lib().nastar_b200_launch_count())


FWD_NO_EARLY_EXIT = 1
FWD_PAIR = 2
COST_PLANE, COST_LOGIT, COST_TAPS = 0, 1, 2


class SearchCounts(tuple):
    """(n_closed [B'] i32, path_len [B'] i32): histories.sum() / paths.sum() per map, written by the kernel."""


def forward(cost: torch.Tensor, start: torch.Tensor, goal: torch.Tensor, obst: torch.Tensor,
            g_ratio: float, T: int, want_trace: bool = False, no_early_exit: bool = False, *,
            pair: bool = False, cost_kind: int = COST_PLANE, cost_scale: float = 1.0, cost_bias: float = 0.0,
            want_counts: bool = False):
    """Run the search for a batch of [B,C,H,W] fp32 CUDA planes (channel 0 is used).

    Returns (histories [B',1,H,W] f32, paths [B',1,H,W] i64, t_solve [B'] i32, n_steps [B'] i32,
    trace [B',T] i32 or None), all on the inputs' device, asynchronously on the current stream; with
    `want_counts` a sixth element (n_closed [B'], path_len [B']) is appended.  B' = 2B with `pair`
    (NASTAR_FWD_PAIR: second half = the same problems with cost = obstacles), else B.
    `cost_kind` = COST_LOGIT / COST_TAPS: `cost` holds the encoder's raw output [B,1,H,W] / the 9-tap partial
    products [B,H,W,9] and the kernel prologue finishes encoder.py:32-34 itself (H, W <= 64 only).
    """
    L = lib()
    if not cost.is_cuda:
        raise RuntimeError("neural_astar (B200): the search runs on CUDA tensors only (no CPU fallback); "
                           "move the planner inputs to the GPU")
    dev = cost.device
    for t_ in (start, goal, obst):
        if t_.device != dev:
            raise RuntimeError("all planes must live on the same CUDA device")
    if start.ndim != 4:
        raise ValueError("planes are [B,C,H,W]")
    B, _, H, W = start.shape
    planes = [("start", start.detach()), ("goal", goal.detach()), ("obst", obst.detach())]
    keep = []
    p = FwdParams()
    if cost_kind == COST_TAPS:
        if cost.dtype != torch.float32 or tuple(cost.shape) != (B, H, W, 9) or not cost.is_contiguous():
            raise ValueError("COST_TAPS expects a contiguous fp32 [B,H,W,9] tensor")
        keep.append(cost)
        p.cost, p.cost_stride = cost.data_ptr(), 9 * H * W
    else:
        planes.insert(0, ("cost", cost.detach()))
    for _, t_ in planes:
        if t_.shape[0] != B or t_.shape[-2:] != (H, W):
            raise ValueError("plane shapes differ")
    for name, t_ in planes:
        k, ptr, stride = _plane(t_)
        keep.append(k)
        setattr(p, name, ptr)
        setattr(p, name + "_stride", stride)
    gr, omg, _ = host_scalars(g_ratio, W)
    p.B, p.H, p.W = B, H, W
    p.g_ratio, p.one_minus_g_ratio = gr, omg
    p.T = int(T)
    p.flags = (FWD_NO_EARLY_EXIT if no_early_exit else 0) | (FWD_PAIR if pair else 0)
    p.cost_kind, p.cost_scale, p.cost_bias = int(cost_kind), float(cost_scale), float(cost_bias)
    Bo = 2 * B if pair else B
    ws_bytes = _ws_cache.get((B, H, W))
    if ws_bytes is None:
        ws_bytes = _ws_cache[(B, H, W)] = int(L.nastar_b200_forward_workspace_bytes(B, H, W))
    guard = contextlib.nullcontext() if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)
    with guard:
        hist = torch.empty((Bo, 1, H, W), dtype=torch.float32, device=dev)
        paths = torch.empty((Bo, 1, H, W), dtype=torch.int64, device=dev)
        counters = torch.empty((4 if want_counts else 2, Bo), dtype=torch.int32, device=dev)   # [t_solve | n_steps | ...]
        t_solve, n_steps = counters[0], counters[1]
        trace = torch.empty((Bo, int(T)), dtype=torch.int32, device=dev) if want_trace else None
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
        p.histories, p.paths = hist.data_ptr(), paths.data_ptr()
        p.t_solve, p.n_steps = t_solve.data_ptr(), n_steps.data_ptr()
        if want_counts:
            p.n_closed, p.path_len = counters[2].data_ptr(), counters[3].data_ptr()
        p.trace = trace.data_ptr() if want_trace else None
        p.workspace = ws.data_ptr() if ws is not None else None
        p.workspace_bytes = ws_bytes
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(L.nastar_b200_forward(ctypes.byref(p), ctypes.c_void_p(stream)), "nastar_b200_forward")
    del keep
    if want_counts:
        return hist, paths, t_solve, n_steps, trace, SearchCounts((counters[2], counters[3]))
    return hist, paths, t_solve, n_steps, trace


def pack_inputs(map_designs: torch.Tensor, start: torch.Tensor, goal: torch.Tensor) -> torch.Tensor:
    """Encoder input of NeuralAstar.encode (reference astar.py:172-177) in ONE kernel: channels-last
    [B, C+1, Hm, Wm] = cat(map_designs, nearest-upsampled(start + goal))."""
    L = lib()
    if not map_designs.is_cuda or map_designs.dtype != torch.float32:
        raise RuntimeError("pack_inputs: fp32 CUDA tensors only")
    B, C, Hm, Wm = map_designs.shape
    H, W = start.shape[-2], start.shape[-1]
    md = map_designs.detach().contiguous()
    ks, ps, ss = _plane(start.detach())
    kg, pg, sg = _plane(goal.detach())
    dev = map_designs.device
    guard = contextlib.nullcontext() if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)
    with guard:
        out = torch.empty((B, Hm, Wm, C + 1), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(L.nastar_b200_pack_inputs(md.data_ptr(), C, Hm, Wm, ps, ss, pg, sg, B, H, W, out.data_ptr(),
                                         ctypes.c_void_p(stream)), "nastar_b200_pack_inputs")
    del ks, kg
    return out.permute(0, 3, 1, 2)   # logical NCHW view with channels-last strides


HEAD_CHANNELS = (32, 64, 128, 256)


def conv1_marks(map_designs: torch.Tensor, start: torch.Tensor, goal: torch.Tensor, w_host, bias_host) -> torch.Tensor:
    """relu(bias + conv3x3(cat(map_designs, start + goal))) as a channels-last [B,32,H,W] tensor in one kernel (the
    "m+" encoder's first layer fused with NeuralAstar.encode's input assembly).  w_host: C-contiguous float32 NumPy
    array [9,2,32] = [tap, cin, cout] (BatchNorm folded), bias_host: float32 NumPy [32]; both on the HOST (they travel
    as kernel parameters)."""
    L = lib()
    B, C, H, W = map_designs.shape
    if C != 1 or start.shape[-2:] != (H, W) or map_designs.dtype != torch.float32 or not map_designs.is_cuda:
        raise ValueError("conv1_marks: one-channel fp32 CUDA maps with start/goal maps of the same size")
    if (w_host.dtype.name != "float32" or w_host.shape != (9, 2, 32) or not w_host.flags["C_CONTIGUOUS"]
            or bias_host.dtype.name != "float32" or bias_host.shape != (32,) or not bias_host.flags["C_CONTIGUOUS"]):
        raise ValueError("conv1_marks: w_host must be a C-contiguous float32 [9,2,32] array, bias_host float32 [32]")
    md = map_designs.detach().contiguous()
    ks, ps, ss = _plane(start.detach())
    kg, pg, sg = _plane(goal.detach())
    dev = map_designs.device
    guard = contextlib.nullcontext() if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)
    with guard:
        out = torch.empty((B, H, W, 32), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(L.nastar_b200_conv1_marks(md.data_ptr(), ps, ss, pg, sg, B, H, W, w_host.ctypes.data,
                                         bias_host.ctypes.data, out.data_ptr(), ctypes.c_void_p(stream)),
               "nastar_b200_conv1_marks")
    del ks, kg
    return out.permute(0, 3, 1, 2)


def head_taps(x: torch.Tensor, w_host, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """taps [B,H,W,9] = per-pixel products of a channels-last activation x [B,C,H,W] with the head conv's folded
    weights `w_host` (a contiguous float32 NumPy array [C,9] on the HOST: it travels as a kernel parameter)."""
    L = lib()
    B, C, H, W = x.shape
    if C not in HEAD_CHANNELS or x.dtype != torch.float32 or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("head_taps expects a channels-last fp32 activation with 32/64/128/256 channels")
    if w_host.dtype.name != "float32" or w_host.shape != (C, 9) or not w_host.flags["C_CONTIGUOUS"]:
        raise ValueError("w_host must be a C-contiguous float32 [C,9] array")
    dev = x.device
    guard = contextlib.nullcontext() if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)
    with guard:
        if out is None:
            out = torch.empty((B, H, W, 9), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(L.nastar_b200_head_taps(x.data_ptr(), B * H * W, C, w_host.ctypes.data, out.data_ptr(),
                                       ctypes.c_void_p(stream)), "nastar_b200_head_taps")
    return out


def cost_from_taps(taps: torch.Tensor, bias: float, scale: float) -> torch.Tensor:
    """cost maps [B,1,H,W] = sigmoid(bias + 9-tap gather(taps)) * scale; taps is fp32 [B,H,W,9] (the same
    arithmetic as the search kernel's COST_TAPS prologue)."""
    L = lib()
    B, H, W, nine = taps.shape
    if nine != 9 or taps.dtype != torch.float32 or not taps.is_contiguous() or not taps.is_cuda:
        raise ValueError("cost_from_taps expects a contiguous fp32 CUDA [B,H,W,9] tensor")
    dev = taps.device
    guard = contextlib.nullcontext() if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)
    with guard:
        cost = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(L.nastar_b200_cost_from_taps(taps.data_ptr(), B, H, W, float(bias), float(scale), cost.data_ptr(),
                                            ctypes.c_void_p(stream)), "nastar_b200_cost_from_taps")
    return cost


def batch_steps(t_solve: torch.Tensor, n_steps: torch.Tensor, T: int) -> torch.Tensor:
    """Device-side T_batch (int32[1]); replaces the per-step host sync of differentiable_astar.py:251."""
    L = lib()
    dev = t_solve.device
    with torch.cuda.device(dev):
        out = torch.empty((1,), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(L.nastar_b200_batch_steps(t_solve.data_ptr(), n_steps.data_ptr(), t_solve.numel(), int(T),
                                         out.data_ptr(), ctypes.c_void_p(stream)), "nastar_b200_batch_steps")
    return out


def backward(cost, start, goal, obst, grad_hist: torch.Tensor, T_batch: torch.Tensor, t_solve: torch.Tensor,
             g_ratio: float) -> torch.Tensor:
    """dL/dcost [B,1,H,W] from dL/dhistories (closed form of the reference's autograd)."""
    L = lib()
    dev = cost.device
    B, _, H, W = cost.shape
    keep = []
    p = BwdParams()
    for name, t_ in (("cost", cost.detach()), ("start", start.detach()), ("goal", goal.detach()),
                     ("obst", obst.detach())):
        k, ptr, stride = _plane(t_)
        keep.append(k)
        setattr(p, name, ptr)
        setattr(p, name + "_stride", stride)
    gh, ptr, stride = _plane(grad_hist.detach().to(torch.float32))
    keep.append(gh)
    p.grad_histories, p.grad_stride = ptr, stride
    gr, omg, sq = host_scalars(g_ratio, W)
    p.B, p.H, p.W = B, H, W
    p.g_ratio, p.one_minus_g_ratio, p.sqrt_w = gr, omg, sq
    p.T_batch = T_batch.data_ptr()
    p.t_solve = t_solve.data_ptr()
    with torch.cuda.device(dev):
        grad_cost = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        ws_bytes = L.nastar_b200_backward_workspace_bytes(B, H, W)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
        p.grad_cost = grad_cost.data_ptr()
        p.workspace = ws.data_ptr() if ws is not None else None
        p.workspace_bytes = ws_bytes
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(L.nastar_b200_backward(ctypes.byref(p), ctypes.c_void_p(stream)), "nastar_b200_backward")
    del keep
    return grad_cost
