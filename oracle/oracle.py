"""ctypes front-end of the CPU oracle (oracle/astar_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Restates /root/reference/src/neural_astar/planner/differentiable_astar.py:150-267
(see the C file's header for the per-function map).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import NamedTuple, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libastar_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "astar_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libastar_oracle.so"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lp = ctypes.POINTER(ctypes.c_int64)
        lib.nastar_oracle_forward_literal.argtypes = [fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                      fp, lp, ip, ip, ip]
        lib.nastar_oracle_forward_literal.restype = ctypes.c_int
        lib.nastar_oracle_forward_spec.argtypes = [fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                   fp, lp, ip, ip, ip]
        lib.nastar_oracle_forward_spec.restype = ctypes.c_int
        lib.nastar_oracle_backward.argtypes = [fp, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, fp]
        lib.nastar_oracle_backward.restype = ctypes.c_int
        lib.nastar_oracle_num_threads.restype = ctypes.c_int
        lib.nastar_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


class OracleOutput(NamedTuple):
    histories: np.ndarray  # [B,1,H,W] float32
    paths: np.ndarray  # [B,1,H,W] int64
    t_solve: np.ndarray  # [B] int32
    n_steps: np.ndarray  # [B] int32 (spec) / T_batch broadcast (literal)
    T_batch: int
    trace: Optional[np.ndarray]  # [B,T] int32 or None


def _planes(x) -> np.ndarray:
    a = np.asarray(x, dtype=np.float32)
    assert a.ndim == 4, "planes are [B,1,H,W] (differentiable_astar.py:172-175)"
    return np.ascontiguousarray(a[:, 0])


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def scalars(g_ratio: float, W: int):
    """Host-side scalar preparation exactly as the reference does it in Python doubles
    (differentiable_astar.py:206-207) before ATen casts them to fp32."""
    return np.float32(g_ratio), np.float32(1 - g_ratio), np.float32(math.sqrt(W))


def num_steps(Tmax: float, training: bool, W: int) -> int:
    """differentiable_astar.py:200-202"""
    return int((Tmax if training else 1.0) * W * W)


def forward(cost, start, goal, obst, g_ratio=0.5, Tmax=1.0, training=False, mode="spec",
            want_trace=False, T=None, no_early_exit=False) -> OracleOutput:
    lib = _load()
    c, s, g, o = _planes(cost), _planes(start), _planes(goal), _planes(obst)
    B, H, W = c.shape
    T = int(T) if T is not None else num_steps(Tmax, training, W)
    gr, omg, sq = scalars(g_ratio, W)
    hist = np.zeros((B, H, W), np.float32)
    paths = np.zeros((B, H, W), np.int64)
    ts = np.zeros(B, np.int32)
    ns = np.zeros(B, np.int32)
    trace = np.full((B, T), -1, np.int32) if want_trace else None
    tp = _p(trace, ctypes.c_int32) if want_trace else None
    if mode == "literal":
        tb = ctypes.c_int32(0)
        rc = lib.nastar_oracle_forward_literal(_p(c, ctypes.c_float), _p(s, ctypes.c_float), _p(g, ctypes.c_float),
                                               _p(o, ctypes.c_float), B, H, W, gr, omg, sq, T,
                                               _p(hist, ctypes.c_float), _p(paths, ctypes.c_int64),
                                               _p(ts, ctypes.c_int32), tp, ctypes.byref(tb))
        T_batch = int(tb.value)
        ns[:] = T_batch
    elif mode == "spec":
        rc = lib.nastar_oracle_forward_spec(_p(c, ctypes.c_float), _p(s, ctypes.c_float), _p(g, ctypes.c_float),
                                            _p(o, ctypes.c_float), B, H, W, gr, omg, T, int(bool(no_early_exit)),
                                            _p(hist, ctypes.c_float), _p(paths, ctypes.c_int64),
                                            _p(ts, ctypes.c_int32), _p(ns, ctypes.c_int32), tp)
        T_batch = int(ns.max())
    else:
        raise ValueError(mode)
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return OracleOutput(hist[:, None], paths[:, None], ts, ns, T_batch, trace)


def backward(cost, start, goal, obst, grad_hist, T_batch: int, g_ratio=0.5) -> np.ndarray:
    lib = _load()
    c, s, g, o, gh = _planes(cost), _planes(start), _planes(goal), _planes(obst), _planes(grad_hist)
    B, H, W = c.shape
    gr, omg, sq = scalars(g_ratio, W)
    out = np.zeros((B, H, W), np.float32)
    rc = lib.nastar_oracle_backward(_p(c, ctypes.c_float), _p(s, ctypes.c_float), _p(g, ctypes.c_float),
                                    _p(o, ctypes.c_float), _p(gh, ctypes.c_float), B, H, W, gr, omg, sq,
                                    int(T_batch), _p(out, ctypes.c_float))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return out[:, None]


def set_threads(n: int) -> None:
    _load().nastar_oracle_set_threads(int(n))


def max_threads() -> int:
    return int(_load().nastar_oracle_num_threads())
