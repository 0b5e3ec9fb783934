"""Training helpers with the reference's interface (/root/reference/src/neural_astar/utils/training.py).

`PlannerModule` is a `pytorch_lightning.LightningModule` when Lightning is installed (so
scripts/train.py runs unchanged) and a plain `nn.Module` with the same methods otherwise; a minimal
`fit()` loop is provided for environments without Lightning (this image has none).
Optimiser, loss and validation metrics follow :42-87: RMSprop(lr), L1(histories, opt_trajs),
p_opt / p_exp / h_mean against VanillaAstar.
"""
from __future__ import annotations

import random
import re
from glob import glob

import numpy as np
import torch
import torch.nn as nn
import torch.optim

from ..planner.astar import VanillaAstar

try:  # pragma: no cover - optional dependency
    import pytorch_lightning as pl

    _Base = pl.LightningModule
except ImportError:  # Lightning-free fallback keeps the same method surface
    pl = None
    _Base = nn.Module


def load_from_ptl_checkpoint(checkpoint_path: str) -> dict:
    """Newest *.ckpt under the directory -> planner state dict (keys containing 'planner', prefix stripped;
    reference :20-39)."""
    ckpt_file = sorted(glob(f"{checkpoint_path}/**/*.ckpt", recursive=True))[-1]
    print(f"load {ckpt_file}")
    state_dict = torch.load(ckpt_file, map_location="cpu", weights_only=False)["state_dict"]
    return {re.split("planner.", k)[-1]: v for k, v in state_dict.items() if "planner" in k}


def planner_metrics(outputs, va_outputs):
    """p_opt, p_exp, h_mean on device (reference :71-85 computes them in NumPy after a host copy)."""
    pathlen_astar = va_outputs.paths.sum((1, 2, 3))
    pathlen_model = outputs.paths.sum((1, 2, 3))
    p_opt = (pathlen_astar == pathlen_model).double().mean()
    exp_astar = va_outputs.histories.sum((1, 2, 3)).double()
    exp_na = outputs.histories.sum((1, 2, 3)).double()
    p_exp = torch.clamp((exp_astar - exp_na) / exp_astar, min=0.0).mean()
    h_mean = 2.0 / (1.0 / (p_opt + 1e-10) + 1.0 / (p_exp + 1e-10))
    return float(p_opt), float(p_exp), float(h_mean)


class PlannerModule(_Base):
    def __init__(self, planner, config):
        super().__init__()
        self.planner = planner
        self.vanilla_astar = VanillaAstar()
        self.config = config

    def forward(self, map_designs, start_maps, goal_maps):
        return self.planner(map_designs, start_maps, goal_maps)

    def configure_optimizers(self) -> torch.optim.Optimizer:
        return torch.optim.RMSprop(self.planner.parameters(), self.config.params.lr)

    def _log(self, name, value):
        if pl is not None:
            self.log(name, value)
        else:
            self.__dict__.setdefault("logged", {})[name] = float(value)

    def training_step(self, train_batch, batch_idx):
        map_designs, start_maps, goal_maps, opt_trajs = train_batch
        outputs = self.forward(map_designs, start_maps, goal_maps)
        loss = nn.L1Loss()(outputs.histories, opt_trajs)
        self._log("metrics/train_loss", loss)
        return loss

    def validation_step(self, val_batch, batch_idx):
        map_designs, start_maps, goal_maps, opt_trajs = val_batch
        outputs = self.forward(map_designs, start_maps, goal_maps)
        loss = nn.L1Loss()(outputs.histories, opt_trajs)
        self._log("metrics/val_loss", loss)
        if map_designs.shape[1] == 1:  # shortest-path problems
            va_outputs = self.vanilla_astar(map_designs, start_maps, goal_maps)
            p_opt, p_exp, h_mean = planner_metrics(outputs, va_outputs)
            self._log("metrics/p_opt", p_opt)
            self._log("metrics/p_exp", p_exp)
            self._log("metrics/h_mean", h_mean)
        return loss

    def fit(self, train_loader, val_loader=None, num_epochs: int = 1, device="cuda"):
        """Plain-PyTorch stand-in for pl.Trainer.fit (same step functions)."""
        self.to(device)
        opt = self.configure_optimizers()
        history = []
        for epoch in range(num_epochs):
            self.train()
            for i, batch in enumerate(train_loader):
                batch = [torch.as_tensor(x).to(device) for x in batch]
                opt.zero_grad(set_to_none=True)
                loss = self.training_step(batch, i)
                loss.backward()
                opt.step()
                history.append(float(loss))
            if val_loader is not None:
                self.eval()
                with torch.no_grad():
                    for i, batch in enumerate(val_loader):
                        self.validation_step([torch.as_tensor(x).to(device) for x in batch], i)
        return history


def set_global_seeds(seed: int) -> None:
    """reference :90-106"""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
    np.random.seed(seed)
    random.seed(seed)
