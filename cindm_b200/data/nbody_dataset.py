"""Reader of the reference's N-body trajectory files for `--initialization_mode 1 / 2`.

The reference driver builds a PyG `NBodyDataset` (data/nbody_dataset.py:21-222), takes the FIRST batch of an
unshuffled DataLoader (inference/inverse_design_diffusion_1d.py:182-201) and turns its `y` into the diffusion
layout with `get_item_1d` (utils.py:203-223); that tensor is `initialization_img` of `sample()` (:313-314).
This module restates exactly that indexing with numpy (memory-mapped, no torch_geometric):

* files: `<dataset_path>/nbody-2/trajectory_balls_2_simu_6000_steps_1000.npy`, `nbody-4/..._4_simu_2000_...`,
  `nbody-8/..._8_simu_200_...`, arrays `[n_simu, 1000, n_bodies, 4]` in pixel units (`:80-88`);
* only the first 800 of the 1000 stamps are used (`time_stamps = 800`, `:71`); with `t_in = max(input_steps *
  time_interval, 1)` and `t_out = max(output_steps * time_interval, 1)`, a simulation yields
  `(800 - t_in - t_out) // time_interval` samples (`:104`, `:192-194`);
* sample `idx -> (sim_id, time_id) = divmod(idx, samples_per_sim)`; evaluation splits use the LAST `n_simu`
  simulations (`:196-199`); `y = data[sim, t0 : t0 + output_steps * time_interval : time_interval]` with
  `t0 = time_id * time_interval + t_in`, `x` the `input_steps` strided stamps before `t0` (`:208-209`);
* `get_item_1d`: `[B, n_bodies, steps, 4] / 200 -> [B, steps, n_bodies * 4]`.

A file with fewer simulations than the reference's (a locally generated one) is accepted: the split sizes are
then clipped to what the file holds.
"""
import os

import numpy as np
import torch

# (total simulations, test-data split, eval split) per body count (reference :52-70)
_SPLITS = {1: (6000, 100, 100), 2: (6000, 100, 100), 4: (2000, 200, 100), 8: (200, 20, 10)}
TIME_STAMPS = 800


def trajectory_file(dataset_path, n_bodies):
    total = _SPLITS[n_bodies][0]
    return os.path.join(dataset_path, f"nbody-{n_bodies}", f"trajectory_balls_{n_bodies}_simu_{total}_steps_1000.npy")


class NBodyDataset:
    def __init__(self, dataset="nbody-2", input_steps=1, output_steps=1, time_interval=1, is_y_diff=True, is_train=True,
                 show_missing_files=False, transform=None, pre_transform=None, is_testdata=False, verbose=0,
                 dataset_path="dataset/nbody_dataset", data=None):
        if not dataset.startswith("nbody-"):
            raise ValueError(f"unknown dataset {dataset!r}")
        self.dataset = dataset
        self.n_bodies = int(dataset.split("-")[1])
        if self.n_bodies not in _SPLITS:
            raise ValueError(f"no trajectory file layout for {self.n_bodies} bodies")
        self.input_steps, self.output_steps, self.time_interval = input_steps, output_steps, time_interval
        self.is_train, self.is_testdata = is_train, is_testdata
        self.t_cushion_input = max(input_steps * time_interval, 1)
        self.t_cushion_output = max(output_steps * time_interval, 1)
        if data is None:
            path = trajectory_file(dataset_path, self.n_bodies)
            if not os.path.isfile(path):
                raise FileNotFoundError(f"{path} not found (reference layout, data/nbody_dataset.py:80-88)")
            data = np.load(path, mmap_mode="r")
        if data.ndim != 4 or data.shape[2] != self.n_bodies or data.shape[3] != 4 or data.shape[1] < TIME_STAMPS:
            raise ValueError(f"trajectory array must be [n_simu, >= {TIME_STAMPS}, {self.n_bodies}, 4], got {data.shape}")
        self.data = data
        total, n_test, n_eval = _SPLITS[self.n_bodies]
        self.total_n_simu = min(total, data.shape[0])
        if is_testdata:
            n = n_test
        else:
            n = total - 2 * n_eval if is_train else n_eval          # 6000-200 / 2000-200 / 200-20 (:55-70)
        self.n_simu = max(1, min(n, self.total_n_simu))
        self.time_stamps = TIME_STAMPS
        self.time_stamps_effective = (TIME_STAMPS - self.t_cushion_input - self.t_cushion_output) // time_interval
        self.dyn_dims = 4

    def len(self):
        return self.time_stamps_effective * self.n_simu

    __len__ = len

    def get(self, idx):
        """-> dict(x [n_bodies, input_steps, 4], y [n_bodies, output_steps, 4], sim_id, time_id), float32 pixel units."""
        if idx < 0 or idx >= self.len():
            raise IndexError(idx)
        sim_id, time_id = divmod(idx, self.time_stamps_effective)
        if not self.is_train:
            sim_id += self.total_n_simu - self.n_simu
        ti = self.time_interval
        t0 = time_id * ti + self.t_cushion_input
        x = np.asarray(self.data[sim_id, t0 - self.input_steps * ti: t0: ti], dtype=np.float32).transpose(1, 0, 2)
        y = np.asarray(self.data[sim_id, t0: t0 + self.output_steps * ti: ti], dtype=np.float32).transpose(1, 0, 2)
        return {"x": x, "y": y, "sim_id": sim_id, "time_id": time_id}

    __getitem__ = get


def get_item_1d(items, target="y"):
    """Batch of `get()` results -> [B, steps, n_bodies * 4] / 200 (utils.py:203-223)."""
    arr = np.stack([it[target] for it in items])                   # [B, n_bodies, steps, 4]
    arr = arr / np.float32(200.0)
    b, n, steps, f = arr.shape
    return torch.from_numpy(np.ascontiguousarray(arr.transpose(0, 2, 1, 3)).reshape(b, steps, n * f))


def first_batch_1d(dataset, batch_size, target="y"):
    """What the driver's `for data in dataloader: break` + get_item_1d yields: samples 0 .. batch_size-1, unshuffled."""
    n = min(batch_size, dataset.len())
    return get_item_1d([dataset.get(i) for i in range(n)], target)
