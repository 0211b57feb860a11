"""CPU oracle #2 -- functional torch port of the reference FNO forward / loss / step tail.

TEST INFRASTRUCTURE ONLY (see oracle/dft_oracle.py header).  This is the "port" that
``bench.py`` times as ``cpu_baseline`` / ``--impl reference`` on the GPU box's host cores
(the Python reference under /root/reference cannot travel to the GPU box) and the
model-level checker used by ``tests/`` and ``__graft_entry__.smoke()``.

It follows the reference algorithm op for op -- same torch.fft / einsum / conv / gelu calls,
therefore the same library arithmetic the reference delegates to -- but is written as pure
functions over a parameter mapping (a ``state_dict``), so it runs in float32 *or* float64
(the reference modules are fp32-only: ``out_ft`` is hard-coded cfloat, fno.py:81):

* :func:`spectral_conv`      fno/fno.py:70-92 (2-D), :259-288 (3-D)
* :func:`fno_forward`        fno/fno.py:139-188 (FNO2d), :342-390 (FNO3d)
* :func:`fno_aux_forward`    fno_aux/fno_aux.py:123-222 (2-D), :383-475 (3-D)
* :func:`nrmse`              fno/train.py:34-40
* :func:`train_step_tail`    fno/train.py:271-278 (zero_grad/backward/clip/Adam/sched)

Pinned against the unmodified reference modules by ``oracle/make_golden.py`` ->
``tests/golden/*.npz`` -> ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import torch
import torch.nn.functional as F

Params = Mapping[str, torch.Tensor]


def _corner_slices(modes: Sequence[int]):
    """Corner index tuples in the reference's weight order (fno.py:84-89, :274-285)."""
    if len(modes) == 2:
        m1, m2 = modes
        return [
            (slice(None, m1), slice(None, m2)),
            (slice(-m1, None), slice(None, m2)),
        ]
    m1, m2, m3 = modes
    return [
        (slice(None, m1), slice(None, m2), slice(None, m3)),
        (slice(-m1, None), slice(None, m2), slice(None, m3)),
        (slice(None, m1), slice(-m2, None), slice(None, m3)),
        (slice(-m1, None), slice(-m2, None), slice(None, m3)),
    ]


def spectral_conv(x: torch.Tensor, weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """rfftn -> per-corner einsum -> zero-padded spectrum -> irfftn."""
    modes = tuple(weights[0].shape[2:])
    nd = len(modes)
    dims = tuple(range(-nd, 0))
    spatial = tuple(x.shape[-nd:])
    spec = torch.fft.rfftn(x, dim=dims)
    cdtype = spec.dtype
    full = torch.zeros(
        (x.shape[0], weights[0].shape[1]) + spatial[:-1] + (spatial[-1] // 2 + 1,),
        dtype=cdtype, device=x.device,
    )
    letters = "xyz"[:nd]
    eq = f"bi{letters},io{letters}->bo{letters}"
    for sl, w in zip(_corner_slices(modes), weights):
        idx = (slice(None), slice(None)) + sl
        full[idx] = torch.einsum(eq, spec[idx], w.to(cdtype))
    return torch.fft.irfftn(full, s=spatial, dim=dims)


def _conv_weights(p: Params, prefix: str, layer: int, nd: int):
    n = 2 if nd == 2 else 4
    return [p[f"{prefix}conv{layer}.weights{k}"] for k in range(1, n + 1)]


def trunk(p: Params, x: torch.Tensor, nd: int, prefix: str = "") -> torch.Tensor:
    """Four Fourier layers on channel-first padded activations (fno.py:161-178)."""
    for layer in range(4):
        spec = spectral_conv(x, _conv_weights(p, prefix, layer, nd))
        w = p[f"{prefix}w{layer}.weight"].reshape(p[f"{prefix}w{layer}.weight"].shape[:2])
        lin = torch.einsum("oi,bi...->bo...", w, x) + p[f"{prefix}w{layer}.bias"].reshape(
            (1, -1) + (1,) * nd)
        x = spec + lin
        if layer < 3:
            x = F.gelu(x)
    return x


def _normalise(x: torch.Tensor, nd: int):
    dims = tuple(range(1, nd + 2))
    with torch.no_grad():
        std, mean = torch.std_mean(x, dim=dims, keepdim=True)
        std = std + 1e-7
    return (x - mean) / std, std, mean


def _lift(p: Params, x: torch.Tensor, grid: torch.Tensor, nd: int, prefix: str = "") -> torch.Tensor:
    feat = torch.cat((x.reshape(*x.shape[:-2], -1), grid), dim=-1)
    h = F.linear(feat, p[f"{prefix}fc0.weight"], p[f"{prefix}fc0.bias"])
    h = h.movedim(-1, 1)
    if nd == 2:
        return F.pad(h, [0, 2, 0, 2])           # padding = 2 on both axes (fno.py:113,159)
    return F.pad(h, [0, 6])                      # padding = 6, last axis only (fno.py:314,360)


def _project(p: Params, h: torch.Tensor, nd: int, head: str, prefix: str = "") -> torch.Tensor:
    h = h[..., :-2, :-2] if nd == 2 else h[..., :-6]
    h = h.movedim(1, -1)
    h = F.gelu(F.linear(h, p[f"{prefix}fc1.weight"], p[f"{prefix}fc1.bias"]))
    return F.linear(h, p[f"{prefix}{head}.weight"], p[f"{prefix}{head}.bias"])


def fno_forward(p: Params, x: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """FNO2d / FNO3d forward.  x: [B, X, Y, (Z,) T, V]; grid: [B, X, Y, (Z,) nd]."""
    nd = grid.shape[-1]
    xn, std, mean = _normalise(x, nd)
    h = trunk(p, _lift(p, xn, grid, nd), nd)
    out = _project(p, h, nd, "fc2")
    out = out * std.squeeze(-2) + mean.squeeze(-2)
    return out.unsqueeze(-2)


def fno_aux_forward(p: Params, x, grid, x_aux, grid_aux):
    """Two-head forward of fno_aux: shared trunk run on the primary and the auxiliary batch."""
    nd = grid.shape[-1]
    outs = []
    for xi, gi, head in ((x, grid, "fc2_primary"), (x_aux, grid_aux, "fc2_auxiliary")):
        xn, std, mean = _normalise(xi, nd)
        h = trunk(p, _lift(p, xn, gi, nd), nd)
        o = _project(p, h, nd, head)
        outs.append((o * std.squeeze(-2) + mean.squeeze(-2)).unsqueeze(-2))
    return outs[0], outs[1]


def nrmse(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Per-sample, per-(t, v) normalised MSE over dims 1..3 (train.py:34-40)."""
    dims = tuple(range(output.ndim))[1:4]
    num = (output - target).pow(2).mean(dims, keepdim=True)
    den = 1e-7 + target.pow(2).mean(dims, keepdim=True)
    return num / den


def grad_global_norm(params) -> torch.Tensor:
    norms = [torch.norm(q.grad.detach(), 2) for q in params if q.grad is not None]
    return torch.norm(torch.stack(norms), 2)


def train_step_tail(loss: torch.Tensor, params, optimizer, scheduler=None):
    """zero_grad -> backward -> adaptive clip -> Adam -> (scheduler) as train.py:271-278.

    Returns (total_norm, clipped_norm)."""
    params = list(params)
    optimizer.zero_grad()
    loss.backward()
    total = grad_global_norm(params)
    clip_value = max(5, 0.1 * total)
    torch.nn.utils.clip_grad_norm_(params, clip_value)
    clipped = grad_global_norm(params)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return total, clipped


# ----------------------------------------------------------------------------
# Parameter construction in the reference's RNG order (fno.py:96-137, :292-340)
# ----------------------------------------------------------------------------
def init_params(nd: int, num_channels: int, modes: Sequence[int], width: int, initial_step: int,
                aux: bool = False) -> dict:
    """Builds a leaf-parameter dict consuming the global torch RNG in the same order as the
    reference constructors, so that ``torch.manual_seed(s); init_params(...)`` equals
    ``torch.manual_seed(s); FNO2d(...).state_dict()`` tensor for tensor."""
    from torch import nn

    p: dict = {}

    def take(name, mod):
        for k, v in mod.state_dict().items():
            p[f"{name}.{k}"] = v.detach().clone()

    take("fc0", nn.Linear(initial_step * num_channels + nd, width))
    ncorn = 2 if nd == 2 else 4
    scale = 1.0 / (width * width)
    for layer in range(4):
        for k in range(1, ncorn + 1):
            p[f"conv{layer}.weights{k}"] = scale * torch.rand(width, width, *modes, dtype=torch.cfloat)
    for layer in range(4):
        take(f"w{layer}", nn.Conv2d(width, width, 1) if nd == 2 else nn.Conv3d(width, width, 1))
    if nd == 3:
        for layer in range(4):
            take(f"bn{layer}", nn.BatchNorm3d(width))
    take("fc1", nn.Linear(width, 128))
    if aux:
        take("fc2_primary", nn.Linear(128, num_channels))
        take("fc2_auxiliary", nn.Linear(128, num_channels))
    else:
        take("fc2", nn.Linear(128, num_channels))
    return p


def as_leaves(p: Params, dtype=None) -> dict:
    """Clone into autograd leaves (float tensors optionally promoted, e.g. to float64)."""
    out = {}
    for k, v in p.items():
        if not (v.is_floating_point() or v.is_complex()):
            out[k] = v.clone()
            continue
        t = v.detach().clone()
        if dtype is torch.float64:
            t = t.to(torch.complex128 if t.is_complex() else torch.float64)
        out[k] = t.requires_grad_(True)
    return out
