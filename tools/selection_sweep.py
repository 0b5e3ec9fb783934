#!/usr/bin/env python
"""LITERAL-vs-SPEC selection sweep (VERDICT r1 weak #2 / next #1c), CPU only, build container only.

The engine selects arg-min(f, flat index) over the open set (SPEC); the reference selects
argmax(exp(-f/sqrt(W)) * open / sum) with first-index ties (differentiable_astar.py:55-74,206-209; LITERAL in the
oracle, pinned step for step to the reference by tests/test_oracle_golden.py).  The two differ only when expf /
the division round two DISTINCT f values to the same softmax weight.  This sweep counts, over every map of the
shipped datasets, how often that changes anything:

    python tools/selection_sweep.py > profiles/r02_selection_sweep.txt

Inputs: all eight 32x32 MPD families and the two 64x64 sets (all_064, street mixed_064), train/valid/test, start
positions from MazeDataset under seed 1234; cost maps = (a) the map itself (VanillaAstar), (b) the shipped
checkpoint's encoder output (const = 1) and (c) the same x10 (const = 10, tighter f spacing relative to fp32 ulp).
Reported per (dataset, cost kind): maps, selections, maps whose histories / paths masks differ, differing trace
positions, and maps whose traces differ at all.
"""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import oracle  # noqa: E402

DATA = "/root/reference/planning-datasets/data"
FILES = [f"mpd/{n}_032_moore_c8.npz" for n in ("mazes", "alternating_gaps", "bugtrap_forest", "forest", "gaps_and_forest",
                                                "multiple_bugtraps", "shifting_gaps", "single_bugtrap")]
FILES += ["mpd/all_064_moore_c16.npz", "street/mixed_064_moore_c16.npz"]
MAX_MAPS_64 = 1200     # per 64x64 file (LITERAL is a dense exp per cell per step)


def encoder():
    from neural_astar.planner import NeuralAstar

    na = NeuralAstar(encoder_input="m+", encoder_arch="CNN", encoder_depth=4)
    st = np.load(os.path.join(ROOT, "tests", "golden", "mazes032_ckpt_planner_state.npz"))
    na.load_state_dict({k: torch.from_numpy(st[k]) for k in st.files})
    return na.eval()


def batches(path):
    from neural_astar.utils.data import MazeDataset

    np.random.seed(1234)
    torch.manual_seed(1234)
    out = []
    for split in ("train", "valid", "test"):
        with contextlib.redirect_stdout(io.StringIO()):    # the dataset announces its size like the reference does
            ds = MazeDataset(path, split)
        n = len(ds)
        items = [ds[i] for i in range(n)]
        out.append(tuple(np.stack([it[k] for it in items]).astype(np.float32) for k in range(3)))
    maps, start, goal = (np.concatenate([o[k] for o in out]) for k in range(3))
    return maps, start[:, :1], goal


def compare(cost, start, goal, obst):
    W = cost.shape[-1]
    lit = oracle.forward(cost, start, goal, obst, mode="literal", want_trace=True)
    spec = oracle.forward(cost, start, goal, obst, mode="spec", want_trace=True)
    dh = (lit.histories != spec.histories).reshape(len(cost), -1).any(1)
    dp = (lit.paths != spec.paths).reshape(len(cost), -1).any(1)
    # literal traces run to T_batch (post-solve steps repeat the goal); compare each map up to its own solve step
    ns = spec.n_steps
    tpos = 0
    tmaps = 0
    for b in range(len(cost)):
        d = int((lit.trace[b, :ns[b]] != spec.trace[b, :ns[b]]).sum())
        tpos += d
        tmaps += d > 0
    return dict(maps=len(cost), selections=int(ns.sum()), hist_maps=int(dh.sum()), path_maps=int(dp.sum()),
                trace_positions=tpos, trace_maps=tmaps)


def main():
    oracle.build()
    torch.set_num_threads(os.cpu_count() or 8)
    na = encoder()
    print(f"# LITERAL (reference selection) vs SPEC (engine selection); torch {torch.__version__}; "
          f"oracle threads {oracle.max_threads()}")
    print("# dataset | cost | maps | selections | maps hist differ | maps path differ | trace positions differ | maps trace differ")
    tot = {}
    for f in FILES:
        t0 = time.time()
        maps, start, goal = batches(os.path.join(DATA, f))
        if maps.shape[-1] > 32 and len(maps) > MAX_MAPS_64:
            # valid+test in full, then as many train maps as fit
            n = len(maps)
            keep = np.r_[np.arange(n - 800, n), np.arange(0, MAX_MAPS_64 - 800)]
            maps, start, goal = maps[keep], start[keep], goal[keep]
        with torch.no_grad():
            enc = torch.cat([na.encode(*(torch.from_numpy(x[i:i + 200]) for x in (maps, start, goal)))
                             for i in range(0, len(maps), 200)]).numpy()
        for kind, cost in (("vanilla", maps), ("learned x1", enc), ("learned x10", enc * np.float32(10.0))):
            r = compare(np.ascontiguousarray(cost), start, goal, maps)
            print(f"{f} | {kind} | {r['maps']} | {r['selections']} | {r['hist_maps']} | {r['path_maps']} | "
                  f"{r['trace_positions']} | {r['trace_maps']}", flush=True)
            for k, v in r.items():
                tot.setdefault(kind, {}).setdefault(k, 0)
                tot[kind][k] += v
        print(f"#   ({time.time() - t0:.0f} s)", flush=True)
    for kind, r in tot.items():
        print(f"TOTAL | {kind} | {r['maps']} | {r['selections']} | {r['hist_maps']} | {r['path_maps']} | "
              f"{r['trace_positions']} | {r['trace_maps']}")


if __name__ == "__main__":
    main()
