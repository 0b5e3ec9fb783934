"""Developer probe: Config 3 (training step) and Config 4 (WarCraft-shaped) timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from neural_astar import _native
from neural_astar.planner import NeuralAstar
from golden_util import Golden

def ev_time(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

# ---- Config 3: training step, 32x32, B=100, Tmax=0.25, RMSprop, L1 ----
g = Golden("mazes032_vanilla_test")
maps, start, goal = (torch.from_numpy(x).cuda() for x in (g.obst, g.start, g.goal))
opt_traj = torch.from_numpy(g.bits("opt_bits").astype(np.float32)).cuda()
torch.manual_seed(1234)
planner = NeuralAstar(Tmax=0.25).cuda().train()
opt = torch.optim.RMSprop(planner.parameters(), 1e-3)
def train_step():
    opt.zero_grad(set_to_none=True)
    out = planner(maps, start, goal)
    loss = torch.nn.L1Loss()(out.histories, opt_traj)
    loss.backward(); opt.step()
ms = ev_time(train_step)
print(f"C3 train step (enc fwd+bwd, search fwd+bwd, RMSprop) B=100: {ms:.3f} ms -> {1e3/ms:.1f} steps/s, {100/ms*1e3:.0f} maps/s")
cost = planner.encode(maps, start, goal).detach().requires_grad_(True)
def search_fb():
    out = planner.perform_astar(cost, start, goal, maps)
    out.histories.sum().backward()
ms2 = ev_time(search_fb)
print(f"C3 search fwd+bwd only: {ms2*1e3:.1f} us")
ms3 = ev_time(lambda: _native.forward(cost.detach(), start, goal, maps, 0.5, 256))
print(f"C3 search fwd only (T=256): {ms3*1e3:.1f} us")

# ---- Config 4: WarCraft-shaped, B=512, 96x96 RGB -> 12x12 ----
gen = torch.Generator().manual_seed(1234)
md = torch.rand(512, 3, 96, 96, generator=gen).cuda()
s12 = torch.zeros(512, 1, 12, 12, device="cuda"); s12[:, :, 0, 0] = 1
g12 = torch.zeros(512, 1, 12, 12, device="cuda"); g12[:, :, -1, -1] = 1
torch.manual_seed(1234)
wc = NeuralAstar(encoder_input="rgb+", encoder_arch="CNNDownSize", encoder_depth=3, learn_obstacles=True, const=10.0).cuda().eval()
with torch.no_grad():
    ms4 = ev_time(lambda: wc(md, s12, g12))
    c12 = wc.encode(md, s12, g12)
    ms5 = ev_time(lambda: _native.forward(c12, s12, g12, torch.ones_like(s12), 0.5, 144))
    out = wc(md, s12, g12)
print(f"C4 NeuralAstar fwd B=512: {ms4:.3f} ms -> {512/ms4*1e3:.0f} maps/s; search only {ms5*1e3:.1f} us (incl. host launch overhead); "
      f"mean expansions {float(out.histories.sum())/512:.1f}")
