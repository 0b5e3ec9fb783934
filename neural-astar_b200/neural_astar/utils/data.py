"""Dataset readers with the reference's interface (host-side I/O; not part of the GPU hot path).

Same public names and return conventions as /root/reference/src/neural_astar/utils/data.py:
`create_dataloader` / `MazeDataset` (:54-245), `create_warcraft_dataloader` / `WarCraftDataset`
(:248-295), `visualize_results` (:16-51).  File formats: the planning-datasets `.npz`
(arr_0..11 = {train,valid,test} x {maps, goals, optimal policies, optimal distances}, :127-150)
and the WarCraft `.npy` pairs (:273-282).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.utils.data as data

from ..planner.differentiable_astar import AstarOutput

# Moore-neighbourhood action order used by the dataset's optimal policies (reference :234-243)
_ACTION_TO_MOVE = ((0, -1, 0), (0, 0, +1), (0, 0, -1), (0, +1, 0), (0, -1, +1), (0, -1, -1), (0, +1, +1), (0, +1, -1))
_SPLIT_OFFSET = {"train": 0, "valid": 4, "test": 8}


def _make_grid(x: torch.Tensor, nrow: int = 8, padding: int = 2) -> torch.Tensor:
    """torchvision.utils.make_grid(x) with its defaults, without the torchvision dependency: a [B,C,H,W] batch
    (single-channel images repeated to RGB) tiled 8 per row with a 2-pixel zero border; a batch of ONE image is
    returned as is, unpadded — exactly like torchvision."""
    x = x.detach().cpu()
    if x.dim() == 3:
        x = x.unsqueeze(0)
    if x.shape[1] == 1:
        x = x.repeat(1, 3, 1, 1)
    B, C, H, W = x.shape
    if B == 1:
        return x[0]
    xmaps = min(nrow, B)
    ymaps = -(-B // xmaps)
    hh, ww = H + padding, W + padding
    grid = x.new_zeros((C, hh * ymaps + padding, ww * xmaps + padding))
    for k in range(B):
        yy, xx = divmod(k, xmaps)
        grid[:, yy * hh + padding:yy * hh + padding + H, xx * ww + padding:xx * ww + padding + W] = x[k]
    return grid


def visualize_results(map_designs: torch.Tensor, planner_outputs, scale: int = 1) -> np.ndarray:
    """Search results as one uint8 RGB image [Hg, Wg, 3] (reference :16-51): the maps tiled like
    torchvision's make_grid (8 per row, 2-pixel border), explored nodes painted (0.2, 0.8, 0), path cells
    (1, 0, 0), values * 255 — the format scripts/create_gif.py hands to moviepy.  `scale` > 1 enlarges with
    nearest-neighbour resampling (the reference passes (rows*scale, cols*scale) to PIL as (width, height);
    kept as is)."""
    if isinstance(planner_outputs, dict):
        histories, paths = planner_outputs["histories"], planner_outputs["paths"]
    else:
        histories, paths = planner_outputs.histories, planner_outputs.paths
    results = _make_grid(map_designs.float()).permute(1, 2, 0).clone()
    h = _make_grid(histories.float()).permute(1, 2, 0)
    p = _make_grid(paths.float()).permute(1, 2, 0)
    results[h[..., 0] == 1] = torch.tensor([0.2, 0.8, 0.0])
    results[p[..., 0] == 1] = torch.tensor([1.0, 0.0, 0.0])
    results = (results.numpy() * 255.0).astype("uint8")
    if scale > 1:
        from PIL import Image

        results = np.asarray(Image.fromarray(results).resize([x * scale for x in results.shape[:2]],
                                                             resample=Image.NEAREST))
    return results


class MazeDataset(data.Dataset):
    def __init__(self, filename: str, split: str, pct1: float = 0.55, pct2: float = 0.70, pct3: float = 0.85,
                 num_starts: int = 1):
        """Shortest-path problems from a planning-datasets .npz (reference :82-125).

        __getitem__ returns (map_design [1,W,W], start_map [num_starts,W,W], goal_map [1,W,W],
        opt_traj [num_starts,W,W]) as float32 arrays.
        """
        assert filename.endswith("npz")
        self.filename = filename
        self.dataset_type = split
        self.pcts = np.array([pct1, pct2, pct3, 1.0])
        self.num_starts = num_starts
        with np.load(filename) as f:
            k = _SPLIT_OFFSET[split]
            self.map_designs, self.goal_maps, self.opt_policies, self.opt_dists = (
                f[f"arr_{k + j}"].astype(np.float32) for j in range(4))
        label = {"train": "Train", "valid": "Validation", "test": "Test"}[split]
        print(f"Number of {label} Samples: {self.map_designs.shape[0]}")
        print(f"\tSize: {self.map_designs.shape[1]}x{self.map_designs.shape[2]}")
        self.num_actions = self.opt_policies.shape[1]
        self.num_orient = self.opt_policies.shape[2]

    def __len__(self):
        return self.map_designs.shape[0]

    def __getitem__(self, index: int):
        """Same samples and the same np.random consumption as the reference (:152-166), but the per-map work
        that does not depend on the draw (percentile bands, the policy's successor table) is computed once
        and cached — at GPU training speeds the reference's per-sample NumPy work is the bottleneck
        (SURVEY.md 8(f) rank 4)."""
        map_design = self.map_designs[index][np.newaxis]
        goal_map = self.goal_maps[index]
        bands, succ = self._prepared(index)
        H, W = goal_map.shape[-2:]
        goal_flat = int(np.argmax(goal_map.reshape(-1)))
        starts = np.zeros((self.num_starts, H, W), np.float32)
        trajs = np.zeros((self.num_starts, H, W), np.float32)
        for k in range(self.num_starts):
            r = np.random.randint(0, len(bands))          # same draw order as get_random_start_map
            loc = int(np.random.choice(bands[r]))
            starts[k].reshape(-1)[loc] = 1.0
            flat = trajs[k].reshape(-1)
            while loc != goal_flat:
                flat[loc] = 1.0
                loc = int(succ[loc])
                assert flat[loc] == 0.0, "Revisiting the same position while following the optimal policy"
        return map_design, starts, goal_map, trajs

    def _prepared(self, index: int):
        cache = self.__dict__.setdefault("_cache", {})
        hit = cache.get(index)
        if hit is None:
            od = self.opt_dists[index].flatten()
            th = np.percentile(od[od > od.min()], 100.0 * (1 - self.pcts))
            bands = [np.where((od >= th[r + 1]) & (od <= th[r]))[0] for r in range(len(th) - 1)]
            policy = self.opt_policies[index]                       # [A, 1, H, W] one-hot over actions
            H, W = policy.shape[-2:]
            act = policy.reshape(policy.shape[0], -1).argmax(0)      # first maximal action, like np.argmax
            moves = np.array([m[1] * W + m[2] for m in _ACTION_TO_MOVE])
            succ = np.arange(H * W) + moves[act]
            hit = cache[index] = (bands, succ)
        return hit

    def get_random_start_map(self, opt_dist: np.ndarray) -> np.ndarray:
        """Pick a start uniformly inside one of the 55-70 / 70-85 / 85-100 percentile bands of the optimal
        distance (reference :200-221; consumes np.random in the same order: randint then choice)."""
        od = opt_dist.flatten()
        th = np.percentile(od[od > od.min()], 100.0 * (1 - self.pcts))
        r = np.random.randint(0, len(th) - 1)
        candidates = np.where((od >= th[r + 1]) & (od <= th[r]))[0]
        start = np.zeros_like(opt_dist)
        start.ravel()[np.random.choice(candidates)] = 1.0
        return start

    def get_opt_traj(self, start_map: np.ndarray, goal_map: np.ndarray, opt_policy: np.ndarray) -> np.ndarray:
        """Roll the optimal policy out from the start until the goal (reference :171-198)."""
        traj = np.zeros_like(start_map)
        policy = opt_policy.transpose((1, 2, 3, 0))
        loc = tuple(np.array(np.nonzero(start_map)).squeeze())
        goal = tuple(np.array(np.nonzero(goal_map)).squeeze())
        while loc != goal:
            traj[loc] = 1.0
            nxt = self.next_loc(loc, policy[loc])
            assert traj[nxt] == 0.0, "Revisiting the same position while following the optimal policy"
            loc = nxt
        return traj

    def next_loc(self, current_loc: tuple, one_hot_action: np.ndarray) -> tuple:
        return tuple(np.add(current_loc, _ACTION_TO_MOVE[int(np.argmax(one_hot_action))]))


def create_dataloader(filename: str, split: str, batch_size: int, num_starts: int = 1,
                      shuffle: bool = False) -> data.DataLoader:
    """reference :54-78"""
    return data.DataLoader(MazeDataset(filename, split, num_starts=num_starts), batch_size=batch_size,
                           shuffle=shuffle, num_workers=0)


class WarCraftDataset(data.Dataset):
    def __init__(self, dirname: str, split: str):
        """WarCraft terrain maps: `<split>_maps.npy` (uint8 HWC) and `<split>_shortest_paths.npy` (reference :273-295)."""
        maps = np.load(f"{dirname}/{split}_maps.npy")
        self.map_designs = (maps.transpose(0, 3, 1, 2) / 255.0).astype(np.float32)
        self.paths = np.load(f"{dirname}/{split}_shortest_paths.npy").astype(np.float32)

    def __len__(self):
        return self.map_designs.shape[0]

    def __getitem__(self, index: int):
        opt_traj = self.paths[index][np.newaxis]
        start = np.zeros_like(opt_traj)
        start[:, 0, 0] = 1
        goal = np.zeros_like(opt_traj)
        goal[:, -1, -1] = 1
        return self.map_designs[index], start, goal, opt_traj


def create_warcraft_dataloader(dirname: str, split: str, batch_size: int, shuffle: bool = False) -> data.DataLoader:
    """reference :248-270"""
    return data.DataLoader(WarCraftDataset(dirname, split), batch_size=batch_size, shuffle=shuffle, num_workers=0)


__all__ = ["AstarOutput", "MazeDataset", "WarCraftDataset", "create_dataloader", "create_warcraft_dataloader",
           "visualize_results"]
