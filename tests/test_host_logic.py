"""CPU-only tests: C-ABI surface, host-side module logic, sharding (incl. world_size-2 gloo)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nastar_b200.h")


@pytest.fixture(scope="module")
def built_lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "neural-astar_b200"), "-s", "all"])
    from neural_astar import _native

    return _native


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nastar_b200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    """The shared library loads without a GPU and exports exactly what include/nastar_b200.h declares."""
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(built_lib.EXPORTS) == declared
    lib.nastar_b200_abi_version.restype = ctypes.c_int
    assert lib.nastar_b200_abi_version() == built_lib.ABI_VERSION


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C11 and carries no torch/CUDA types."""
    c = tmp_path / "t.c"
    c.write_text('#include "nastar_b200.h"\nint main(void){ nastar_fwd_params p; (void)p; '
                 'return sizeof(nastar_bwd_params) > 0 ? 0 : 1; }\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(c), "-o", str(tmp_path / "t.o")])
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    assert "torch" not in code and "cudaStream_t" not in code and "#include <cuda" not in code


def test_struct_layout_matches_ctypes(built_lib, tmp_path):
    """ctypes mirrors of the parameter structs have the C compiler's size and field offsets."""
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nastar_b200.h"\nint main(void){'
                 'printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(nastar_fwd_params), offsetof(nastar_fwd_params, T),'
                 'offsetof(nastar_fwd_params, workspace_bytes), sizeof(nastar_bwd_params),'
                 'offsetof(nastar_bwd_params, t_solve), offsetof(nastar_bwd_params, grad_cost)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    F, Bw = built_lib.FwdParams, built_lib.BwdParams
    want = [ctypes.sizeof(F), F.T.offset, F.workspace_bytes.offset, ctypes.sizeof(Bw), Bw.t_solve.offset,
            Bw.grad_cost.offset]
    assert got == want


def test_engine_dispatch_and_workspace(built_lib):
    L = built_lib.lib()
    assert L.nastar_b200_engine_for(32, 32) == 1 and L.nastar_b200_engine_for(12, 12) == 1
    assert L.nastar_b200_engine_for(64, 64) == 4 and L.nastar_b200_engine_for(33, 20) == 4
    assert L.nastar_b200_engine_for(64, 128) == 2
    assert L.nastar_b200_engine_for(128, 128) == 2
    assert L.nastar_b200_engine_for(256, 256) == 3
    assert L.nastar_b200_engine_for(0, 5) == 0 and L.nastar_b200_engine_for(100000, 100000) == 0
    assert L.nastar_b200_forward_workspace_bytes(8, 32, 32) == 0
    assert L.nastar_b200_forward_workspace_bytes(4, 256, 256) >= 4 * 256 * 256 * 9
    assert L.nastar_b200_forward_workspace_bytes(4, 64, 64) == 0
    assert L.nastar_b200_backward_workspace_bytes(4, 64, 64) == 0       # backward of 64x64: warp64 engine, all in shared memory
    assert L.nastar_b200_backward_workspace_bytes(4, 96, 64) >= 4 * 96 * 64 * 28     # generic engine: 28 B per cell
    assert L.nastar_b200_bin16_supported(256, 256) == 1 and L.nastar_b200_bin16_supported(64, 64) == 0
    assert L.nastar_b200_backward_workspace_bytes(4, 32, 32) == 0
    assert L.nastar_b200_status_string(2).decode().startswith("unsupported")


def test_invalid_arguments_are_rejected_without_a_gpu(built_lib):
    L = built_lib.lib()
    p = built_lib.FwdParams()
    assert L.nastar_b200_forward(ctypes.byref(p), None) == 1  # NASTAR_EINVAL: null planes
    assert L.nastar_b200_forward(None, None) == 1
    assert L.nastar_b200_backward(None, None) == 1


def test_glue_entry_points_reject_bad_arguments_without_a_gpu(built_lib):
    """The encoder-side entry points validate before launching: null pointers / sizes at the C ABI, CPU tensors and wrong
    weight layouts in the binding (no compute call is made on this box)."""
    import numpy as np

    L = built_lib.lib()
    assert L.nastar_b200_conv1_marks(None, None, 0, None, 0, 1, 8, 8, None, None, None, None) == 1
    assert L.nastar_b200_head_taps(None, 64, 256, None, None, None) == 1
    assert L.nastar_b200_pack_inputs is not None and L.nastar_b200_cost_from_taps is not None
    maps = torch.ones(2, 1, 8, 8)
    marks = torch.zeros(2, 1, 8, 8)
    w = np.zeros((9, 2, 32), np.float32)
    b = np.zeros((32,), np.float32)
    with pytest.raises(ValueError):
        built_lib.conv1_marks(maps, marks, marks, w, b)                      # CPU tensors
    with pytest.raises((ValueError, AttributeError)):
        built_lib.conv1_marks(maps, marks, marks, w.reshape(18, 32), b)      # wrong weight layout


def test_stream_helpers_validate_arguments():
    from neural_astar.planner import VanillaAstar
    from neural_astar.utils.inference import OverlappedPlanner, PipelinedPlanner

    with pytest.raises(ValueError):
        OverlappedPlanner(VanillaAstar(), n_streams=0)
    with pytest.raises(ValueError):
        OverlappedPlanner(VanillaAstar(), device="cpu")
    with pytest.raises(TypeError):
        PipelinedPlanner(VanillaAstar(), torch.ones(1, 1, 8, 8), torch.ones(1, 1, 8, 8), torch.ones(1, 1, 8, 8))


def test_no_cpu_fallback(built_lib):
    """CPU tensors are refused loudly; the product never routes through the oracle or eager PyTorch."""
    from neural_astar.planner import VanillaAstar

    x = torch.ones(2, 1, 8, 8)
    s = torch.zeros(2, 1, 8, 8)
    s[:, :, 0, 0] = 1
    g = torch.zeros(2, 1, 8, 8)
    g[:, :, -1, -1] = 1
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        VanillaAstar()(x, s, g)
    pkg = os.path.join(ROOT, "neural-astar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower(), f"{f} mentions the oracle"


def test_missing_library_fails_loudly(built_lib, monkeypatch, tmp_path):
    from neural_astar import _native

    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_native.NativeLibraryMissing):
        _native.lib()


def test_module_surface_matches_reference():
    """Constructor signatures, attributes and state-dict keys of the reference."""
    import inspect

    from neural_astar.planner import NeuralAstar, VanillaAstar
    from neural_astar.planner.differentiable_astar import AstarOutput, DifferentiableAstar, get_heuristic

    assert list(inspect.signature(DifferentiableAstar.__init__).parameters)[1:] == ["g_ratio", "Tmax"]
    assert list(inspect.signature(DifferentiableAstar.forward).parameters)[1:] == [
        "cost_maps", "start_maps", "goal_maps", "obstacles_maps", "store_intermediate_results"]
    assert list(inspect.signature(VanillaAstar.__init__).parameters)[1:] == ["g_ratio", "use_differentiable_astar"]
    assert list(inspect.signature(NeuralAstar.__init__).parameters)[1:] == [
        "g_ratio", "Tmax", "encoder_input", "encoder_arch", "encoder_depth", "learn_obstacles", "const",
        "use_differentiable_astar"]
    assert AstarOutput._fields == ("histories", "paths", "intermediate_results")
    na = NeuralAstar()
    keys = set(na.state_dict().keys())
    assert "astar.neighbor_filter" in keys and "encoder.model.0.weight" in keys
    assert "encoder.model.13.running_var" in keys
    nf = na.astar.neighbor_filter
    assert nf.shape == (1, 1, 3, 3) and not nf.requires_grad and float(nf.sum()) == 8 and float(nf[0, 0, 1, 1]) == 0
    assert na.astar.get_heuristic is get_heuristic
    with pytest.raises(AssertionError):
        DifferentiableAstar(Tmax=0.0)
    d = DifferentiableAstar(Tmax=0.25)
    assert d.train().num_steps(32) == 256 and d.eval().num_steps(32) == 1024  # differentiable_astar.py:200-202
    with pytest.raises(AssertionError):
        d(torch.ones(8, 8), torch.ones(8, 8), torch.ones(8, 8), torch.ones(8, 8))  # ndim asserts (:172-175)
    wc = NeuralAstar(encoder_input="rgb+", encoder_arch="CNNDownSize", encoder_depth=3, const=10.0)
    assert wc.encoder.const.item() == 10.0


def test_reference_checkpoint_loads_with_all_keys_matched():
    from neural_astar.planner import NeuralAstar

    state = np.load(os.path.join(ROOT, "tests", "golden", "mazes032_ckpt_planner_state.npz"))
    msg = NeuralAstar(encoder_arch="CNN").load_state_dict({k: torch.from_numpy(state[k]) for k in state.files})
    assert str(msg) == "<All keys matched successfully>"


def test_encoder_matches_reference_cost_maps():
    """The re-provided CNN encoder reproduces the reference encoder's cost maps on CPU (golden fixture)."""
    from golden_util import Golden
    from neural_astar.planner import NeuralAstar

    g = Golden("mazes032_neural_test")
    state = np.load(os.path.join(ROOT, "tests", "golden", "mazes032_ckpt_planner_state.npz"))
    planner = NeuralAstar(encoder_arch="CNN")
    planner.load_state_dict({k: torch.from_numpy(state[k]) for k in state.files})
    planner.eval()
    with torch.no_grad():
        cost = planner.encode(torch.from_numpy(g.obst[:8]), torch.from_numpy(g.start[:8]),
                              torch.from_numpy(g.goal[:8]))
    np.testing.assert_allclose(cost.numpy(), g.cost[:8], rtol=1e-5, atol=1e-6)
    # CNNDownSize geometry of the WarCraft config (96x96 RGB -> 12x12 costs, encoder.py:81-97)
    wc = NeuralAstar(encoder_input="rgb+", encoder_arch="CNNDownSize", encoder_depth=3, learn_obstacles=True,
                     const=10.0).eval()
    s = torch.zeros(2, 1, 12, 12)
    s[:, :, 0, 0] = 1
    with torch.no_grad():
        c = wc.encode(torch.rand(2, 3, 96, 96), s, s.flip(-1, -2))
    assert c.shape == (2, 1, 12, 12) and float(c.min()) >= 0 and float(c.max()) <= 10


def test_get_heuristic_matches_formula():
    from neural_astar.planner.differentiable_astar import get_heuristic

    goal = torch.zeros(2, 5, 7)
    goal[0, 4, 6] = 1
    goal[1, 0, 3] = 1
    h = get_heuristic(goal).numpy()
    for b, (gy, gx) in enumerate(((4, 6), (0, 3))):
        for y in range(5):
            for x in range(7):
                dy, dx = abs(y - gy), abs(x - gx)
                want = np.float32(np.float32(dy + dx - min(dy, dx)) +
                                  np.float32(0.001) * np.sqrt(np.float32(dy * dy + dx * dx)))
                assert h[b, y, x] == pytest.approx(float(want), abs=1e-6)


@pytest.mark.parametrize("n,world", [(100, 1), (100, 2), (100, 8), (13, 4), (3, 8), (0, 2), (8192, 8)])
def test_shard_range_partitions_the_batch(n, world):
    from neural_astar.utils.distributed import shard_range

    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neural_astar.utils.distributed import aggregate_throughput, shard_batch

    full = torch.arange(101 * 4).reshape(101, 1, 2, 2).float()
    (mine,) = shard_batch((full,), rank, world)
    maps, exp, sec = aggregate_throughput(mine.shape[0], float(mine.sum()), 1.0 + rank)
    q.put((rank, mine.shape[0], maps, exp, sec, float(mine[0].sum())))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    """world_size-2 run on CPU: shards are disjoint, cover the batch, and throughput aggregates as
    (SUM maps, SUM expansions, MAX seconds)."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full_sum = float(torch.arange(101 * 4).sum())
    assert [r[1] for r in res] == [51, 50]
    for r in res:
        assert r[2] == 101 and r[3] == full_sum and r[4] == 2.0
    assert res[0][5] == float(torch.arange(4).sum())
    assert res[1][5] == float(torch.arange(51 * 4, 51 * 4 + 4).sum())


# ---- host-side module logic exercised on CPU with the engine call replaced by the oracle ------------------
class _FakeNative:
    """Stands in for neural_astar._native in CPU tests: same call signatures, results from the SPEC oracle.
    (Only the Python host logic around the engine is under test here; GPU tests cover the real engine.)"""

    def __init__(self, oracle):
        self.oracle = oracle
        self.calls = []

    def forward(self, cost, start, goal, obst, g_ratio, T, want_trace=False, no_early_exit=False):
        self.calls.append((int(T), bool(want_trace), bool(no_early_exit)))
        o = self.oracle.forward(cost.detach().numpy(), start.numpy(), goal.numpy(), obst.detach().numpy(),
                                g_ratio=g_ratio, mode="spec", want_trace=True, T=T, no_early_exit=no_early_exit)
        tr = torch.from_numpy(o.trace) if want_trace else None
        return (torch.from_numpy(o.histories), torch.from_numpy(o.paths), torch.from_numpy(o.t_solve),
                torch.from_numpy(o.n_steps), tr)

    def batch_steps(self, t_solve, n_steps, T):
        v = torch.where(t_solve >= 0, t_solve + 1, torch.full_like(t_solve, T))
        return torch.clamp(v.max(), max=T).reshape(1).to(torch.int32)

    def backward(self, cost, start, goal, obst, grad_hist, T_batch, t_solve, g_ratio):
        self.calls.append(("bwd", int(T_batch.item()), t_solve.clone()))
        gc = self.oracle.backward(cost.detach().numpy(), start.numpy(), goal.numpy(), obst.detach().numpy(),
                                  grad_hist.numpy(), int(T_batch.item()), g_ratio=g_ratio)
        return torch.from_numpy(gc)


@pytest.fixture
def fake_engine(oracle, monkeypatch):
    from neural_astar.planner import differentiable_astar as da

    fake = _FakeNative(oracle)
    monkeypatch.setattr(da, "_native", fake)
    return fake


def test_intermediate_frames_host_logic(fake_engine):
    """_materialise_frames: T_batch+1 frames, frame t = (closed set before step t, node selected at t), solved maps
    repeat their goal, last frame = (histories, paths) — against the reference's own trace (golden)."""
    from golden_util import Golden
    from neural_astar.planner.differentiable_astar import DifferentiableAstar

    g = Golden("mazes032_vanilla_gr07")
    astar = DifferentiableAstar(g_ratio=g.g_ratio).eval()
    t = [torch.from_numpy(x) for x in (g.cost, g.start, g.goal, g.obst)]
    out = astar(*t, store_intermediate_results=True)
    ref = g.z["trace"]
    assert len(out.intermediate_results) == int(g.z["T_batch"]) + 1 == ref.shape[1] + 1
    sel = torch.stack([f["paths"].reshape(g.B, -1).argmax(1) for f in out.intermediate_results[:-1]], 1).numpy()
    np.testing.assert_array_equal(sel, ref)
    closed = np.stack([f["histories"].reshape(g.B, -1).sum(1).numpy() for f in out.intermediate_results[:-1]], 1)
    np.testing.assert_array_equal(closed, np.minimum(np.arange(ref.shape[1])[None, :], g.z["hist_sum"][:, None]))
    assert torch.equal(out.intermediate_results[-1]["paths"], out.paths)
    assert fake_engine.calls == [(32 * 32, True, False)]
    assert astar(*t).intermediate_results == []


@pytest.mark.parametrize("name", ["mazes032_lowg_gr00_cost10", "mazes032_lowg_gr04_cost10"])
def test_batch_coupled_host_procedure(fake_engine, name):
    """g_ratio < 0.5, B > 1: trace without early exit -> first step at which all maps select their goal -> rerun for
    exactly that many steps.  Masks, T_batch and every frame equal the reference's (golden)."""
    from golden_util import Golden
    from neural_astar.planner.differentiable_astar import DifferentiableAstar

    g = Golden(name)
    astar = DifferentiableAstar(g_ratio=g.g_ratio).eval()
    t = [torch.from_numpy(x) for x in (g.cost, g.start, g.goal, g.obst)]
    out = astar(*t, store_intermediate_results=True)
    Tb = int(g.z["T_batch"])
    assert fake_engine.calls == [(1024, True, True), (Tb, True, True)]
    np.testing.assert_array_equal(out.histories.numpy() != 0, g.bits("hist_bits") != 0)
    np.testing.assert_array_equal(out.paths.numpy() != 0, g.bits("path_bits") != 0)
    assert len(out.intermediate_results) == Tb + 1
    # B == 1 or g_ratio >= 0.5 never take the coupled path
    fake_engine.calls.clear()
    astar(*[x[:1] for x in t])
    assert fake_engine.calls == [(1024, False, False)]


def test_autograd_wiring_on_cpu(fake_engine):
    """_AstarSearch: histories carry grad, paths do not; backward gets the device-side T_batch and returns
    dL/dcost with the cost tensor's shape (extra channels receive zeros) — L1 training loss vs the reference's
    autograd gradient (golden)."""
    from golden_util import Golden
    from neural_astar.planner.differentiable_astar import DifferentiableAstar

    g = Golden("mazes032_neural_test")
    astar = DifferentiableAstar(g_ratio=0.5, Tmax=g.meta["train_Tmax"]).train()
    cost2 = torch.from_numpy(np.concatenate([g.cost, np.full_like(g.cost, 3.0)], 1)).requires_grad_(True)  # [B,2,H,W]
    start, goal, obst = (torch.from_numpy(x) for x in (g.start, g.goal, g.obst))
    out = astar(cost2, start, goal, obst)
    assert out.histories.requires_grad and not out.paths.requires_grad
    opt = torch.from_numpy(g.bits("opt_bits").astype(np.float32))
    torch.nn.L1Loss()(out.histories, opt).backward()
    kind, Tb, _ = fake_engine.calls[-1]
    assert kind == "bwd" and Tb == int(g.z["train_T_batch"]) and fake_engine.calls[0] == (256, False, False)
    grad = cost2.grad.numpy()
    assert grad.shape == cost2.shape and not grad[:, 1].any()
    ref = g.plane("train_grad_cost")
    assert np.abs(grad[:, :1] - ref).max() / np.abs(ref).max() < 1e-5


def test_coupled_backward_passes_goal_clamp_encoding(fake_engine):
    """g_ratio < 0.5: the backward receives T_batch = the coupled step count and the 0 / T encoding of 'goal selected
    at least twice' (include/nastar_b200.h, nastar_bwd_params.t_solve)."""
    from golden_util import Golden
    from neural_astar.planner.differentiable_astar import DifferentiableAstar

    g = Golden("mazes032_lowg_gr00_cost10")
    astar = DifferentiableAstar(g_ratio=g.g_ratio).eval()
    cost = torch.from_numpy(g.cost).requires_grad_(True)
    out = astar(cost, torch.from_numpy(g.start), torch.from_numpy(g.goal), torch.from_numpy(g.obst))
    (out.histories * torch.from_numpy(g.plane("rand_G").astype(np.float32))).sum().backward()
    kind, Tb, enc = fake_engine.calls[-1]
    Tbatch = int(g.z["T_batch"])
    assert kind == "bwd" and Tb == Tbatch
    twice = (g.z["trace"] == g.z["goal_idx"][:, None]).sum(1) >= 2
    np.testing.assert_array_equal(enc.numpy(), np.where(twice, 0, Tbatch))
    ref = g.plane("rand_grad_cost")
    assert np.abs(cost.grad.numpy() - ref).max() / np.abs(ref).max() < 1e-5


def test_warcraft_encoder_handoff_matches_reference():
    """Config-4 hand-off on CPU: CNNDownSize(depth 3, const 10) with the reference's weights reproduces the reference's
    cost maps on a 96x96 RGB batch whose start/goal marks are nearest-upsampled from 12x12 (planner/astar.py:172-177)."""
    from neural_astar.planner import NeuralAstar

    z = np.load(os.path.join(ROOT, "tests", "golden", "warcraft_encoder_ckpt.npz"))
    na = NeuralAstar(encoder_input="rgb+", encoder_arch="CNNDownSize", encoder_depth=3, learn_obstacles=True, const=10.0)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    assert str(na.load_state_dict(sd)) == "<All keys matched successfully>"
    na.eval()
    x = torch.from_numpy(z["x"].astype(np.float32))
    with torch.no_grad():
        cost = na.encode(x, torch.from_numpy(z["start"]), torch.from_numpy(z["goal"]))
    np.testing.assert_allclose(cost.numpy(), z["cost"], rtol=1e-5, atol=1e-5)


def test_reference_arm_thread_sweep_is_bounded(monkeypatch):
    """bench.py's reference arm picks its thread count with an ascending sweep that stops at the first slower count
    and cuts a collapsing trial off (the PyTorch loop with 128 unpinned OpenMP threads ran for minutes per batch)."""
    import importlib
    import sys
    import time

    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(os, "sched_getaffinity", lambda _pid: set(range(128)))
    state = {"threads": 0, "calls": []}
    cost = {8: 0.30, 16: 0.10, 32: 0.30, 64: 2.0, 128: 5.0}     # well separated: robust on a loaded box

    def fn():
        state["calls"].append(state["threads"])
        end = time.perf_counter() + cost[state["threads"]]
        while time.perf_counter() < end:        # Python-level loop, like the reference's: the alarm can interrupt it
            time.sleep(0.001)

    best, tried, t = bench._pick_threads(fn, lambda c: state.__setitem__("threads", c))
    assert best == 16 and tried == [8, 16, 32] and 0.05 < t < 0.28 and state["threads"] == 16
    # a collapsing count right after the first candidate is cut off at 4x the best time + 1 s
    cost.update({8: 0.10, 16: 60.0})
    t0 = time.perf_counter()
    best, tried, _ = bench._pick_threads(fn, lambda c: state.__setitem__("threads", c))
    assert best == 8 and tried == [8, 16] and time.perf_counter() - t0 < 10.0
