"""Per-layer timing of the folded encoder: fused cuDNN conv+bias+ReLU vs conv2d(+bias)+relu_, NHWC, benchmark on/off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, torch.nn.functional as F
import bench
planner = bench.load_planner(torch.device("cuda"))
plan = planner.encoder._inference_plan(torch.device("cuda"))
x = torch.rand(100, 2, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return float(np.median(ts))
inp = x
for li,(w,b,stride,pad,dil,relu,pool) in enumerate(plan):
    res = {}
    for bm in (False, True):
        with torch.backends.cudnn.flags(enabled=True, benchmark=bm, deterministic=False, allow_tf32=True):
            if relu:
                res[f"fused bm={bm}"] = t(lambda: torch.cudnn_convolution_relu(inp, w, b, stride, pad, dil, 1))
            res[f"conv2d+bias{'+relu_' if relu else ''} bm={bm}"] = t(lambda: (F.relu_(F.conv2d(inp, w, b, stride, pad, dil, 1)) if relu else F.conv2d(inp, w, b, stride, pad, dil, 1)))
            res[f"conv2d nobias bm={bm}"] = t(lambda: F.conv2d(inp, w, None, stride, pad, dil, 1))
    wn = w.contiguous()  # NCHW weights + NCHW input
    xin = inp.contiguous()
    res["NCHW conv2d+bias"] = t(lambda: F.conv2d(xin, wn, b, stride, pad, dil, 1))
    print(f"layer {li}: {tuple(w.shape)} in {tuple(inp.shape)}")
    for k,v in res.items(): print(f"    {k:34s} {v:8.1f} us")
    with torch.no_grad():
        inp = torch.cudnn_convolution_relu(inp, w, b, stride, pad, dil, 1) if relu else F.conv2d(inp, w, b, stride, pad, dil, 1)
