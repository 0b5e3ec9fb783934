"""Training helpers with the reference's interface (/root/reference/src/neural_astar/utils/training.py).

`PlannerModule` derives from `pytorch_lightning.LightningModule` when Lightning is importable (then
scripts/train.py runs unchanged) and from `nn.Module` otherwise — this image has no Lightning, so a small
`fit()` loop is included.  Optimiser / loss / metrics follow the reference (:42-87): RMSprop(lr),
L1(histories, opt_trajs), and p_opt / p_exp / h_mean measured against VanillaAstar.
"""
from __future__ import annotations

import glob
import random
import re

import numpy as np
import torch
import torch.nn as nn

from ..planner.astar import VanillaAstar

try:  # pragma: no cover - optional dependency
    import pytorch_lightning as pl
except ImportError:
    pl = None
_ModuleBase = nn.Module if pl is None else pl.LightningModule


def set_global_seeds(seed: int) -> None:
    """Seed torch / numpy / random and make cuDNN deterministic (reference :90-106)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.deterministic = True


def load_from_ptl_checkpoint(checkpoint_path: str) -> dict:
    """State dict of the planner inside the newest Lightning *.ckpt below `checkpoint_path`: entries whose key
    mentions 'planner', with everything up to 'planner.' removed (reference :20-39)."""
    newest = sorted(glob.glob(f"{checkpoint_path}/**/*.ckpt", recursive=True))[-1]
    print(f"load {newest}")
    full = torch.load(newest, map_location="cpu", weights_only=False)["state_dict"]
    return {re.split("planner.", name)[-1]: tensor for name, tensor in full.items() if "planner" in name}


def planner_metrics(outputs, va_outputs):
    """(p_opt, p_exp, h_mean) of a planner against vanilla A*, reduced on the device.

    p_opt: share of maps whose path is as long as A*'s; p_exp: mean relative saving in explored nodes
    (clipped at 0); h_mean: their harmonic mean (reference :71-85 does this in NumPy after host copies).
    """
    len_ref, len_out = va_outputs.paths.sum((1, 2, 3)), outputs.paths.sum((1, 2, 3))
    exp_ref = va_outputs.histories.sum((1, 2, 3)).double()
    exp_out = outputs.histories.sum((1, 2, 3)).double()
    p_opt = (len_ref == len_out).double().mean()
    p_exp = ((exp_ref - exp_out) / exp_ref).clamp_min(0.0).mean()
    h_mean = 2.0 / (1.0 / (p_opt + 1e-10) + 1.0 / (p_exp + 1e-10))
    return float(p_opt), float(p_exp), float(h_mean)


def planner_metrics_from_counts(n_closed: torch.Tensor, path_len: torch.Tensor):
    """(p_opt, p_exp, h_mean) from the per-map counts a NASTAR_FWD_PAIR launch wrote: entries [0,B) belong to the
    learned-cost search, [B,2B) to vanilla A* on the same problems (the same formulas as planner_metrics /
    reference :71-85, without reducing the B x H x W masks again)."""
    B = n_closed.numel() // 2
    exp_out, exp_ref = n_closed[:B].double(), n_closed[B:].double()
    p_opt = (path_len[:B] == path_len[B:]).double().mean()
    p_exp = ((exp_ref - exp_out) / exp_ref).clamp_min(0.0).mean()
    h_mean = 2.0 / (1.0 / (p_opt + 1e-10) + 1.0 / (p_exp + 1e-10))
    vals = torch.stack((p_opt, p_exp, h_mean)).tolist()    # one device->host read for the three scalars
    return vals[0], vals[1], vals[2]


class GraphedTrainStep:
    """One training step of scripts/train.py (forward, L1 loss on histories, backward through the search kernels
    and the encoder, RMSprop update; reference utils/training.py:52-61) captured ONCE as a CUDA graph and replayed.

    An eager step issues ~150 small launches from Python (train-mode convs + BatchNorms, their backward, the two
    search kernels, the optimiser's foreach kernels) and is CPU-launch bound at b=100 / 32x32; the replay is one
    launch.  Same arithmetic as the eager step (same kernels, same order).  Fixed batch shape; `g_ratio >= 0.5` or
    batch 1 (the batch-coupled procedure below 0.5 needs a host decision per batch).

        step = GraphedTrainStep(module, example_batch)        # module: PlannerModule, in train() mode, on CUDA
        for batch in loader:
            loss = step(batch)                                # device scalar; float(loss) synchronises
    """

    def __init__(self, module, example_batch, warmup: int = 3):
        planner = module.planner
        if not module.training:
            raise ValueError("GraphedTrainStep captures the training step: call module.train() first")
        if float(getattr(planner, "g_ratio", 0.5)) < 0.5 and example_batch[0].shape[0] > 1:
            raise ValueError("g_ratio < 0.5 needs a host decision per batch and cannot be captured")
        self.module = module
        dev = next(planner.parameters()).device
        self.optimizer = torch.optim.RMSprop(planner.parameters(), module.config.params.lr, capturable=True)
        self._batch = [torch.empty(x.shape, dtype=x.dtype, device=dev) for x in example_batch]
        for dst, src in zip(self._batch, example_batch):
            dst.copy_(torch.as_tensor(src))
        # warm-up on a side stream (cuDNN autotuning, lazy initialisation, optimiser state), then capture; the
        # parameters are restored afterwards so that constructing the step does not train the model
        saved = [p.detach().clone() for p in planner.parameters()]
        bufs = [b.detach().clone() for b in planner.buffers()]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._loss = self._one_step(zero=False)
        with torch.no_grad():
            for p, q in zip(planner.parameters(), saved):
                p.copy_(q)
            for b, q in zip(planner.buffers(), bufs):
                b.copy_(q)
        self.optimizer = self._reset_optimizer_state()
        self.replays = 0

    def _one_step(self, zero: bool = True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        loss, _ = self.module._loss(self._batch)
        loss.backward()
        self.optimizer.step()
        return loss

    def _reset_optimizer_state(self):
        # the captured graph updates the optimiser state tensors in place: zero them instead of replacing them
        for st in self.optimizer.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()
        return self.optimizer

    def __call__(self, batch) -> torch.Tensor:
        for dst, src in zip(self._batch, batch):
            dst.copy_(torch.as_tensor(src), non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self._loss


class PlannerModule(_ModuleBase):
    def __init__(self, planner, config):
        super().__init__()
        self.planner = planner
        self.vanilla_astar = VanillaAstar()
        self.config = config

    # -- Lightning hooks (same names as the reference) ---------------------------------------------
    def forward(self, map_designs, start_maps, goal_maps):
        return self.planner(map_designs, start_maps, goal_maps)

    def configure_optimizers(self) -> torch.optim.Optimizer:
        return torch.optim.RMSprop(self.planner.parameters(), self.config.params.lr)

    def training_step(self, train_batch, batch_idx):
        loss, _ = self._loss(train_batch)
        self._record("metrics/train_loss", loss)
        return loss

    def validation_step(self, val_batch, batch_idx):
        map_designs, start_maps, goal_maps, opt_trajs = val_batch
        pair = None
        if map_designs.shape[1] == 1 and hasattr(self.planner, "forward_pair") and not torch.is_grad_enabled():
            # the learned-cost search and the vanilla baseline of this batch in ONE launch, metrics
            # from the per-map counts the kernel wrote (reference :63-87 runs two planners and reduces in NumPy)
            pair = self.planner.forward_pair(map_designs, start_maps, goal_maps,
                                             vanilla_g_ratio=self.vanilla_astar.g_ratio)
        if pair is not None:
            outputs, _, counts = pair
            loss = nn.functional.l1_loss(outputs.histories, opt_trajs)
            self._record("metrics/val_loss", loss)
            metrics = planner_metrics_from_counts(*counts)
        else:
            loss, outputs = self._loss(val_batch)
            self._record("metrics/val_loss", loss)
            metrics = None
            if map_designs.shape[1] == 1:  # single-channel maps = shortest-path problems with an A* baseline
                metrics = planner_metrics(outputs, self.vanilla_astar(map_designs, start_maps, goal_maps))
        if metrics is not None:
            for name, value in zip(("p_opt", "p_exp", "h_mean"), metrics):
                self._record(f"metrics/{name}", value)
        return loss

    # -- helpers -----------------------------------------------------------------------------------
    def _loss(self, batch):
        map_designs, start_maps, goal_maps, opt_trajs = batch
        outputs = self.forward(map_designs, start_maps, goal_maps)
        return nn.functional.l1_loss(outputs.histories, opt_trajs), outputs

    def _record(self, name, value):
        if pl is not None:
            self.log(name, value)
        else:
            self.__dict__.setdefault("logged", {})[name] = float(value.detach() if torch.is_tensor(value) else value)

    def fit(self, train_loader, val_loader=None, num_epochs: int = 1, device="cuda"):
        """Plain-PyTorch stand-in for pl.Trainer.fit, driving the same step functions."""
        self.to(device)
        optimiser = self.configure_optimizers()
        curve = []
        for _ in range(num_epochs):
            self.train()
            for i, batch in enumerate(train_loader):
                optimiser.zero_grad(set_to_none=True)
                loss = self.training_step([torch.as_tensor(x).to(device) for x in batch], i)
                loss.backward()
                optimiser.step()
                curve.append(float(loss))
            if val_loader is not None:
                self.eval()
                with torch.no_grad():
                    for i, batch in enumerate(val_loader):
                        self.validation_step([torch.as_tensor(x).to(device) for x in batch], i)
        return curve
