"""Head kernel vs a plain streaming read of the same 105 MB activation (rotating buffers > L2)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200")]
import numpy as np, torch
from neural_astar import _native
xs = [torch.randn(100, 32, 32, 256, device="cuda").permute(0, 3, 1, 2) for _ in range(4)]
w = np.ascontiguousarray(np.random.RandomState(0).randn(256, 9).astype(np.float32))
outs = [torch.empty(100, 32, 32, 9, device="cuda") for _ in range(4)]
def timed(fn, n=100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(5): fn(k)
    torch.cuda.synchronize(); a.record()
    for k in range(n): fn(k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
res = {"head_us": timed(lambda k: _native.head_taps(xs[k & 3], w, out=outs[k & 3])),
       "sum_us": timed(lambda k: torch.sum(xs[k & 3])),
       "copy_us": timed(lambda k: outs[0].copy_(outs[1]) if False else xs[(k + 1) & 3].copy_(xs[k & 3]))}
res["head_GBs"] = 104.8576e6 / res["head_us"] / 1e3
res["sum_GBs"] = 104.8576e6 / res["sum_us"] / 1e3
res["copy_GBs_rw"] = 2 * 104.8576e6 / res["copy_us"] / 1e3
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
if len(sys.argv) > 1:
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    _native.head_taps(xs[0], w, out=outs[0])
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
