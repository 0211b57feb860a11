"""Synthetic PDEBench-shaped fields for benchmarks and loop-level parity.

The reference loaders (fno/utils_2d_rd_baseline.py:59-102 and friends) read HDF5 trajectories
``[t, x, y, v]`` and emit sliding windows ``xx [X, Y, initial_step, v]``, ``yy [X, Y, 1, v]`` plus a
cell-centre ``grid [X, Y, 2]``.  No datasets (and no h5py) exist here, so trajectories are
generated: a band-limited Gaussian random field per channel, evolved by exact spectral
diffusion u_hat(t) = u_hat(0) exp(-D |k|^2 t) -- the reference's own ``sim_type="diff"`` physics
(data_gen/src/sim_diff_react.py:165-167).  Generation is host-side utility code, not part of the
measured path.
"""
from __future__ import annotations

import math

import torch
import torch.utils.data


def cell_centre_grid(n: int, lo: float = -1.0, hi: float = 1.0, nd: int = 2) -> torch.Tensor:
    """meshgrid(ij) of cell centres, as the loaders build it -> [n, ..., n, nd]."""
    dx = (hi - lo) / n
    lin = torch.linspace(lo + dx / 2, hi - dx / 2, n)
    return torch.stack(torch.meshgrid(*([lin] * nd), indexing="ij"), dim=-1)


def diffusion_trajectories(n_traj: int, n: int, steps: int, channels: int, seed: int, nd: int = 2,
                           kmax: int = 16, t_end: float = 5.0, diffusivity=(1e-3, 1e-1)) -> torch.Tensor:
    """[n_traj, steps, n, (n,) n, channels] float32 trajectories with O(1) values."""
    g = torch.Generator().manual_seed(seed)
    k1 = torch.fft.fftfreq(n, d=1.0 / n)
    ks = torch.meshgrid(*([k1] * nd), indexing="ij")
    k2 = sum(k * k for k in ks)
    band = (k2 <= kmax * kmax).float() / (1.0 + k2)
    t = torch.linspace(0.0, t_end, steps)
    out = []
    for c in range(channels):
        d = diffusivity[c % len(diffusivity)]
        noise = torch.randn((n_traj,) + (n,) * nd, generator=g)
        u0 = torch.fft.fftn(noise, dim=tuple(range(1, nd + 1))) * band
        decay = torch.exp(-d * (2 * math.pi / 2.0) ** 2 * k2[None, None] * t.view(1, -1, *([1] * nd)))
        u = torch.fft.ifftn(u0[:, None] * decay, dim=tuple(range(2, nd + 2))).real
        u = u / u[:, :1].std(dim=tuple(range(2, nd + 2)), keepdim=True).clamp_min(1e-6)
        out.append(u)
    return torch.stack(out, dim=-1).float()


def windows(traj: torch.Tensor, initial_step: int, rollout: int = 1):
    """All sliding windows of a trajectory batch -> (xx [N, *sp, initial_step, v], yy [N, *sp, rollout, v])."""
    nt = traj.shape[1]
    nd = traj.dim() - 3
    xs, ys = [], []
    for s in range(nt - initial_step - rollout + 1):
        xs.append(traj[:, s:s + initial_step])
        ys.append(traj[:, s + initial_step:s + initial_step + rollout])
    xx = torch.stack(xs, dim=1).flatten(0, 1)
    yy = torch.stack(ys, dim=1).flatten(0, 1)
    perm = (0,) + tuple(range(2, 2 + nd)) + (1, 2 + nd)
    return xx.permute(*perm).contiguous(), yy.permute(*perm).contiguous()


def synthetic_batch(batch: int, n: int = 128, initial_step: int = 10, channels: int = 2, seed: int = 0, nd: int = 2):
    """One batch in the loaders' layout: (xx, yy, grid)."""
    traj = diffusion_trajectories(batch, n, initial_step + 1, channels, seed, nd=nd)
    xx, yy = windows(traj, initial_step)
    grid = cell_centre_grid(n, nd=nd).unsqueeze(0).expand(batch, *([-1] * (nd + 1))).contiguous()
    return xx, yy, grid


class SyntheticWindows(torch.utils.data.Dataset):
    """Map-style dataset with the sample layout of the reference loaders (fno/utils_2d_rd_baseline.py:
    59-102): item i -> (xx [X, Y, initial_step, v], yy [X, Y, rollout, v], grid [X, Y, 2]), sliding
    windows over synthetic diffusion trajectories.  Deterministic in (seed, sizes); used by the
    loop-level parity fixtures (oracle/make_golden_loop.py, tests/test_loop_parity*.py)."""

    def __init__(self, n_traj: int, n: int, windows: int, initial_step: int = 10, channels: int = 2, seed: int = 0,
                 rollout: int = 1):
        steps = initial_step + rollout + windows - 1
        traj = diffusion_trajectories(n_traj, n, steps, channels, seed)
        self.xx, self.yy = windows_of(traj, initial_step, rollout)
        self.grid = cell_centre_grid(n)

    def __len__(self):
        return self.xx.shape[0]

    def __getitem__(self, i):
        return self.xx[i], self.yy[i], self.grid


def windows_of(traj, initial_step, rollout=1):
    return windows(traj, initial_step, rollout)
