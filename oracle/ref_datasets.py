"""Synthetic stand-ins for the HDF5 loaders the reference loops import (`fno.utils_2d_ns_baseline_lie.FNODatasetMult`,
fno/train.py:9; `fno_aux.utils_2d_ns.FNODatasetMult`, fno_aux/fno_train_aux.py:9), with the constructor keywords and item
layout the loops use.  A real module file because the joint loop needs ``num_workers >= 1`` (persistent_workers=True,
fno_train_aux.py:106-111), i.e. a picklable dataset class.  TEST INFRASTRUCTURE ONLY.

Sizes come from the environment variable FNO_REF_DATA (JSON), set by oracle/run_ref_loop.py."""
from __future__ import annotations

import importlib.util
import json
import os
from pathlib import Path

import torch

_spec = importlib.util.spec_from_file_location(
    "_fno_b200_data", Path(__file__).resolve().parent.parent / "sciml-pde_b200" / "fno_b200" / "data.py")
_data = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_data)          # the synthetic generator only: no CUDA library needed


def _cfg():
    return json.loads(os.environ["FNO_REF_DATA"])


class FNODatasetMult(_data.SyntheticWindows):
    """fno/utils_2d_rd_baseline.py:59-102 layout: item -> (xx [X,Y,T0,v], yy [X,Y,rollout,v], grid [X,Y,2])."""

    def __init__(self, saved_folder=None, train_subsample=None, rollout_test=1, if_test=False, **_unused):
        c = _cfg()
        super().__init__(c["val_traj"] if if_test else c["train_traj"], c["n"], c["windows"], c["initial_step"],
                         c["num_channels"], seed=c["val_seed"] if if_test else c["train_seed"], rollout=rollout_test)


class AuxFNODatasetMult(torch.utils.data.Dataset):
    """fno_aux/utils_2d_rd.py layout: item -> (xx, yy, xx_aux [num_aux, ...], yy_aux [num_aux, ...], grid, grid_aux).
    Primary = diffusion fields; auxiliary = `num_aux` pure-diffusion fields with other diffusivities (the decomposed
    auxiliary data of config 2), all on the same grid (utils_2d_rd.py:164)."""

    def __init__(self, saved_folder=None, aux_saved_folder=None, if_test=False, if_downsample=False, train_subsample=None,
                 num_aux_samples=3, rollout_test=1, **_unused):
        c = _cfg()
        ntraj = c["val_traj"] if if_test else c["train_traj"]
        seed = c["val_seed"] if if_test else c["train_seed"]
        steps = c["initial_step"] + rollout_test + c["windows"] - 1
        traj = _data.diffusion_trajectories(ntraj, c["n"], steps, c["num_channels"], seed)
        self.xx, self.yy = _data.windows_of(traj, c["initial_step"], rollout_test)
        xa, ya = [], []
        for k in range(num_aux_samples):
            t = _data.diffusion_trajectories(ntraj, c["n"], steps, c["num_channels"], c["aux_seed"] + 17 * k + (1000 if if_test else 0),
                                             diffusivity=(3e-3 * (k + 1), 3e-2 * (k + 1)))
            a, b = _data.windows_of(t, c["initial_step"], rollout_test)
            xa.append(a)
            ya.append(b)
        self.xx_aux, self.yy_aux = torch.stack(xa, dim=1), torch.stack(ya, dim=1)
        self.grid = _data.cell_centre_grid(c["n"])

    def __len__(self):
        return self.xx.shape[0]

    def __getitem__(self, i):
        return self.xx[i], self.yy[i], self.xx_aux[i], self.yy_aux[i], self.grid, self.grid
