"""CPU/GPU-agnostic restatement of the reference training loop, `training_type="single"`
(/root/reference/pdebench/models/fno/train.py:84-129, :168-177, :217-347).

TEST INFRASTRUCTURE ONLY (see oracle/dft_oracle.py header).  It reproduces the loop's arithmetic
and its quirks -- the things per-epoch parity depends on:

* `next(iter(val_loader))` is drawn BEFORE the model is built (train.py:109 then :114): creating a
  DataLoader iterator consumes the global torch RNG, so the initial weights depend on it;
* `DataLoader(shuffle=True)` for training, `shuffle=False` for validation, `drop_last=False`;
* `CosineAnnealingLR(T_max = epochs * (len(train_data) / batch_size))` with a float T_max, stepped
  after every iteration AND once more per epoch (train.py:175, :278, :340);
* `clip_value = max(5, 0.1 * total_norm)` then `clip_grad_norm_` (train.py:273-275);
* `trainL2` / `testL2` are SUMS of per-batch mean losses (train.py:269, :316, :341-345).

Pinned by tests/golden/loop_cfg1.json: the per-epoch values the UNMODIFIED reference loop printed
for the same synthetic dataset and seed (oracle/make_golden_loop.py, run in the build container).
"""
from __future__ import annotations

import random
from typing import Callable, List, Tuple

import numpy as np
import torch


def set_seed(seed: int):
    """train.py:20-27."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def nrmse(output, tar):
    """train.py:34-40."""
    spatial_dims = tuple(range(output.ndim))[1:4]
    tar_norm = 1e-7 + tar.pow(2).mean(spatial_dims, keepdim=True)
    return (output - tar).pow(2).mean(spatial_dims, keepdim=True) / tar_norm


def run_training_port(make_model: Callable[[], torch.nn.Module], train_data, val_data, device, batch_size: int,
                      epochs: int, learning_rate: float = 1e-3, seed: int = 16) -> List[Tuple[float, float, float]]:
    """Returns [(last_val_batch_loss, trainL2, testL2)] per epoch, the numbers train.py:341-345 prints."""
    set_seed(seed)                                                       # train.py:29-30 (at import)
    train_loader = torch.utils.data.DataLoader(train_data, batch_size=batch_size, num_workers=0, shuffle=True)
    val_loader = torch.utils.data.DataLoader(val_data, batch_size=batch_size, num_workers=0, shuffle=False)
    _, _data, _ = next(iter(val_loader))                                 # train.py:109 -- consumes the global RNG
    model = make_model().to(device)                                      # train.py:113-129
    optimizer = torch.optim.Adam(model.parameters(), lr=learning_rate, weight_decay=1e-4)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=epochs * (len(train_data) / batch_size))
    out = []
    for _ep in range(epochs):
        model.train()
        train_l2_full = 0.0
        for xx, yy, grid in train_loader:
            xx, yy, grid = xx.to(device), yy.to(device), grid.to(device)
            loss = nrmse(model(xx, grid), yy).mean()
            train_l2_full += loss.item()
            optimizer.zero_grad()
            loss.backward()
            total_norm = torch.norm(torch.stack([torch.norm(p.grad.detach(), 2) for p in model.parameters()
                                                 if p.grad is not None]), 2)
            clip_value = max(5, 0.1 * total_norm)
            torch.nn.utils.clip_grad_norm_(model.parameters(), clip_value)
            optimizer.step()
            scheduler.step()
        val_l2_full = 0.0
        with torch.no_grad():
            for xx, yy, grid in val_loader:
                xx, yy, grid = xx.to(device), yy.to(device), grid.to(device)
                loss = nrmse(model(xx, grid), yy).mean()
                val_l2_full += loss.item()
        scheduler.step()                                                 # train.py:340
        out.append((float(loss.item()), train_l2_full, val_l2_full))
    return out
