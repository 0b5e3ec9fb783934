"""Loader for the golden vectors written by tests/golden/make_golden.py (bit-packed masks)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name: str):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.meta = json.loads(str(self.z["meta"]))
        self.B, self.H, self.W = (int(v) for v in self.z["shape"])
        self.N = self.H * self.W

    def bits(self, key: str) -> np.ndarray:
        """[B,1,H,W] uint8 mask"""
        b = np.unpackbits(self.z[key], axis=1)[:, : self.N]
        return b.reshape(self.B, 1, self.H, self.W)

    def onehot(self, key: str) -> np.ndarray:
        x = np.zeros((self.B, self.N), np.float32)
        x[np.arange(self.B), self.z[key]] = 1.0
        return x.reshape(self.B, 1, self.H, self.W)

    @property
    def obst(self):
        return self.bits("obst_bits").astype(np.float32)

    @property
    def start(self):
        return self.onehot("start_idx")

    @property
    def goal(self):
        return self.onehot("goal_idx")

    @property
    def cost(self):
        """learned costs if stored, else the vanilla convention cost == map design (astar.py:93-94)"""
        if "cost" in self.z:
            return self.z["cost"].reshape(self.B, 1, self.H, self.W).astype(np.float32)
        return self.obst

    @property
    def g_ratio(self) -> float:
        return float(self.meta["g_ratio"])

    @property
    def independent(self) -> bool:
        return bool(self.meta.get("independent", False))

    def plane(self, key):
        return self.z[key].reshape(self.B, 1, self.H, self.W)


def all_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if "ckpt" not in os.path.basename(p)
                  and not os.path.basename(p).startswith(("train_curve", "inputs_", "adversarial_")))  # search vectors only
