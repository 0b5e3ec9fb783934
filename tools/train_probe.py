import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from neural_astar.planner import NeuralAstar
from golden_util import Golden
g = Golden("mazes032_vanilla_test")
maps, start, goal = (torch.from_numpy(x).cuda() for x in (g.obst, g.start, g.goal))
opt_traj = torch.from_numpy(g.bits("opt_bits").astype(np.float32)).cuda()
def run(tag, cl, bench):
    torch.manual_seed(1234)
    planner = NeuralAstar(Tmax=0.25).cuda().train()
    if cl: planner.encoder.model.to(memory_format=torch.channels_last)
    opt = torch.optim.RMSprop(planner.parameters(), 1e-3)
    torch.backends.cudnn.benchmark = bench
    def step():
        opt.zero_grad(set_to_none=True)
        m = maps.contiguous(memory_format=torch.channels_last) if cl else maps
        out = planner(m, start, goal)
        loss = torch.nn.L1Loss()(out.histories, opt_traj); loss.backward(); opt.step()
    for _ in range(8): step()
    torch.cuda.synchronize(); ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{tag}: {np.median(ts):.3f} ms/step")
run("baseline NCHW", False, False)
run("NCHW + cudnn.benchmark", False, True)
run("channels_last weights", True, False)
run("channels_last + benchmark", True, True)
