#!/usr/bin/env python
"""Stage the reference's own hot-path modules into oracle/_ref/ so that its PyTorch CPU DifferentiableAstar can be
TIMED on the GPU box's host cores (north star: "reported next to the reference's own PyTorch CPU
DifferentiableAstar timed on the same box's host cores").

TEST / BASELINE INFRASTRUCTURE ONLY.  oracle/_ref/ is build output: git-ignored (never committed — reference
sources do not enter the history), NOT gpurun-ignored (it travels to the GPU box like a built .so).  Only
`bench.py --impl reference` (and the cpu_baseline leg, through a subprocess) imports it; the product package
never does.  /root/reference does not exist on the GPU box, hence the staging at build() time here.

What is staged: /root/reference/src/neural_astar/planner/{__init__,astar,differentiable_astar,encoder,pq_astar}.py,
byte for byte.  Their third-party imports that the hot path never touches (`segmentation_models_pytorch`,
`pqdict`) are satisfied by empty stub modules created next to them.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/neural_astar"
DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "astar.py", "differentiable_astar.py", "encoder.py", "pq_astar.py")


def stage() -> bool:
    src = os.path.join(REF_SRC, "planner")
    if not os.path.isdir(src):
        return os.path.isdir(os.path.join(DST, "neural_astar", "planner"))    # GPU box: use what travelled
    dst = os.path.join(DST, "neural_astar", "planner")
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        a, b = os.path.join(src, f), os.path.join(dst, f)
        if not os.path.exists(b) or not filecmp.cmp(a, b, shallow=False):
            shutil.copyfile(a, b)
    open(os.path.join(DST, "neural_astar", "__init__.py"), "a").close()
    for stub, body in (("segmentation_models_pytorch", "Unet = None\n"), ("pqdict", "pqdict = dict\n")):
        os.makedirs(os.path.join(DST, stub), exist_ok=True)
        with open(os.path.join(DST, stub, "__init__.py"), "w") as fh:
            fh.write("# stub: imported by the reference's package chain, never used by DifferentiableAstar\n" + body)
    return True


if __name__ == "__main__":
    ok = stage()
    print("oracle/_ref staged" if ok else "no /root/reference and nothing staged")
    sys.exit(0)
