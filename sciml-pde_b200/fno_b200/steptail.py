"""Device-side step tail of the reference training loop: fused nRMSE loss,
and gradient-norm + adaptive clip + Adam + cosine LR in three launches with no host round trip.

    loss = nrmse_loss(model(xx, grid), yy)          # = nrmse(...).mean() of fno/train.py:34-40,:266
    opt = FusedClipAdam(model.parameters(), lr=1e-3, weight_decay=1e-4, t_max=T)
    opt.zero_grad(); loss.backward(); opt.step()     # fno/train.py:271-278

The arithmetic is torch.optim.Adam's (coupled L2 weight decay, bias correction, eps outside the
square root) and ``clip_grad_norm_``'s (coefficient ``clip / (norm + 1e-6)`` clamped to 1); complex
parameters are updated as interleaved real pairs exactly as torch does (``view_as_real``).  Because
nothing is evaluated on the host -- the reference computes ``max(5, 0.1 * total_norm)`` in Python --
the whole training step can be captured in one CUDA graph (fno_b200.train.FusedTrainStep(graph=True)).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import lib


class _NrmseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target):
        out, target = out.contiguous(), target.contiguous()
        lib._require(out, torch.float32, "output")
        lib._require(target, torch.float32, "target")
        if out.shape != target.shape:
            raise lib.FnoError(f"nrmse: output {tuple(out.shape)} vs target {tuple(target.shape)}")
        B, V = out.shape[0], out.shape[-1]
        P = out.numel() // (B * V)
        L = lib.load()
        work = torch.empty(L.fno_nrmse_workspace_bytes(B, V) // 4, dtype=torch.float32, device=out.device)
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        with torch.cuda.device(out.device):
            lib._check(L.fno_nrmse_fwd(out.data_ptr(), target.data_ptr(), loss.data_ptr(), work.data_ptr(), B, P, V,
                                       lib._stream()), "fno_nrmse_fwd")
        ctx.save_for_backward(out, target, work)
        ctx.dims = (B, P, V)
        return loss

    @staticmethod
    def backward(ctx, g):
        out, target, work = ctx.saved_tensors
        B, P, V = ctx.dims
        g = g.contiguous().float()
        dout = torch.empty_like(out)
        with torch.cuda.device(out.device):
            lib._check(lib.load().fno_nrmse_bwd(out.data_ptr(), target.data_ptr(), work.data_ptr(), g.data_ptr(),
                                                dout.data_ptr(), B, P, V, lib._stream()), "fno_nrmse_bwd")
        return dout, None


def nrmse_loss(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """``nrmse(output, target).mean()`` (fno/train.py:34-40, :266-267) for ``[B, *spatial, 1, V]`` tensors:
    per (sample, variable) MSE over the pixels normalised by the target's mean square, averaged."""
    if output.shape[-2] != 1 and output.dim() != 5:
        raise lib.FnoError("nrmse_loss: fused form needs one output step (or the 2-D layout [B, X, Y, T, V])")
    return _NrmseFn.apply(output, target)


class FusedClipAdam:
    """Adam(weight_decay) + ``clip_grad_norm_(params, max(clip_floor, clip_frac * total_norm))`` +
    per-iteration CosineAnnealingLR, as three sync-free launches over a chunk table.

    ``params`` is an iterable of parameters or, as for torch.optim.Adam, a list of group dicts
    ``{"params": ..., "lr": ...}`` -- the joint loop's three groups (fno_aux/fno_train_aux.py:175-179: shared trunk at
    ``learning_rate_share``, the two heads at ``learning_rate_fc2``).  Groups may differ in ``lr`` only; they share
    the cosine factor, exactly as one CosineAnnealingLR over the three groups does.

    Gradients live in one flat fp32 buffer (``p.grad`` are views, created after the first backward
    so that parameters which never receive a gradient -- FNO3d's dead ``bn*`` -- are skipped exactly
    as the reference's norm code skips them).  With ``own_grads=False`` the existing ``p.grad``
    tensors are used in place (fno_b200.dp keeps them as views of its all-reduce buckets).

    ``state_dict()`` / ``load_state_dict()`` speak torch.optim.Adam's layout (per-parameter ``step``, ``exp_avg``,
    ``exp_avg_sq``; ``param_groups``), so the reference's checkpoints ``{"epoch", "model_state_dict",
    "optimizer_state_dict", "loss"}`` (fno/train.py:189-204, :319-329) load into either optimizer."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, t_max: float = 0.0, eta_min: float = 0.0, clip_floor: float = 5.0,
                 clip_frac: float = 0.1, own_grads: bool = True):
        params = list(params)
        if params and isinstance(params[0], dict):
            groups = [dict(g) for g in params]
        else:
            groups = [{"params": params}]
        self.params: List[torch.nn.Parameter] = []
        self.group_of: List[int] = []
        self.group_lr: List[float] = []
        seen = set()
        for gi, g in enumerate(groups):
            if "weight_decay" in g and float(g["weight_decay"]) != float(weight_decay):
                raise lib.FnoError("FusedClipAdam: parameter groups may differ in lr only (one weight_decay)")
            self.group_lr.append(float(g.get("lr", lr)))
            for p in g["params"]:             # de-duplicate (fno_aux shared_layers alias the trunk)
                if id(p) not in seen and p.requires_grad:
                    seen.add(id(p))
                    self.params.append(p)
                    self.group_of.append(gi)
        if not self.params:
            raise ValueError("FusedClipAdam: no parameters")
        self.device = self.params[0].device
        if self.device.type != "cuda":
            raise lib.FnoError("FusedClipAdam runs on CUDA only")
        self.own_grads = own_grads
        self.defaults = dict(lr=self.group_lr[0], betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.hparams = torch.tensor([self.group_lr[0], eta_min, float(t_max), betas[0], betas[1], eps, weight_decay,
                                     clip_floor, clip_frac, 0.0, 1.0 - betas[0], 1.0 - betas[1]], dtype=torch.float32,
                                    device=self.device)
        self.state = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._built = False
        self._pending_state = None
        self.flat_grad: Optional[torch.Tensor] = None

    # -- torch.optim-like surface ----------------------------------------------------------------
    def zero_grad(self):
        if not self._built:
            for p in self.params:
                p.grad = None
        elif self.own_grads:
            self.flat_grad.zero_()
        else:
            for p in self.live:
                p.grad.zero_()

    def step(self):
        if not self._built:
            self._build()
        L = lib.load()
        with torch.cuda.device(self.device):
            lib._check(L.fno_clip_adam_step(self.chunks.data_ptr(), self.nchunks, self.partials.data_ptr(),
                                            self.state.data_ptr(), self.hparams.data_ptr(), lib._stream()),
                       "fno_clip_adam_step")

    def scheduler_step(self):
        """One scheduler step that is not an optimizer step: the reference loops step their CosineAnnealingLR after
        every iteration AND once more per epoch (fno/train.py:278 + :340, fno_aux/fno_train_aux.py:329 + :398)."""
        self.hparams[9:10].add_(1.0)

    @property
    def total_norm(self) -> torch.Tensor:
        return self.state[1]

    @property
    def lr(self) -> torch.Tensor:
        return self.state[3]

    # -- checkpoints (torch.optim.Adam layout) ------------------------------------------------------
    def _current_group_lrs(self):
        hp = self.hparams.tolist()
        lr0, eta_min, t_max, extra = hp[0], hp[1], hp[2], hp[9]
        steps = float(self.state[0].item())
        if t_max > 0:
            import math
            f = 0.5 * (1.0 + math.cos(math.pi * (steps + extra) / t_max))
        else:
            f = 1.0
        return [eta_min + (g - eta_min) * f for g in self.group_lr]

    def state_dict(self):
        """``{"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}`` with parameter indices in group
        order -- what torch.optim.Adam(params).state_dict() holds after the same steps.  Parameters that never
        received a gradient have no entry, as in torch."""
        index = {id(p): i for i, p in enumerate(self.params)}
        state = {}
        if self._built:
            step = self.state[0].detach().clone()
            for p, off, n in self._slots:
                shape = tuple(p.shape) + ((2,) if p.is_complex() else ())
                m = self.exp_avg[off:off + n].view(shape).clone()
                v = self.exp_avg_sq[off:off + n].view(shape).clone()
                if p.is_complex():
                    m, v = torch.view_as_complex(m), torch.view_as_complex(v)
                state[index[id(p)]] = {"step": step.clone(), "exp_avg": m, "exp_avg_sq": v}
        elif self._pending_state is not None:
            state = self._pending_state["state"]
        lrs = self._current_group_lrs()
        groups = []
        for gi, lr0 in enumerate(self.group_lr):
            groups.append({"lr": lrs[gi], "betas": self.defaults["betas"], "eps": self.defaults["eps"],
                           "weight_decay": self.defaults["weight_decay"], "amsgrad": False, "maximize": False,
                           "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                           "decoupled_weight_decay": False, "initial_lr": lr0,
                           "params": [i for i, g in enumerate(self.group_of) if g == gi]})
        return {"state": state, "param_groups": groups,
                "fno_b200": {"sched_extra": float(self.hparams[9].item()), "t_max": float(self.hparams[2].item())}}

    def load_state_dict(self, sd):
        """Accepts torch.optim.Adam's state_dict (or this class's): moments and step counter are restored, group
        learning rates are taken from ``initial_lr`` when present (a CosineAnnealingLR stores it) else ``lr``."""
        groups = sd["param_groups"]
        if len(groups) != len(self.group_lr):
            raise lib.FnoError(f"load_state_dict: {len(groups)} parameter groups, optimizer has {len(self.group_lr)}")
        flat = [i for g in groups for i in g["params"]]
        if len(flat) != len(self.params):
            raise lib.FnoError(f"load_state_dict: {len(flat)} parameters, optimizer has {len(self.params)}")
        for gi, g in enumerate(groups):
            self.group_lr[gi] = float(g.get("initial_lr", g["lr"]))
        self.hparams[0] = self.group_lr[0]
        extra = sd.get("fno_b200", {})
        if "sched_extra" in extra:
            self.hparams[9] = float(extra["sched_extra"])
        # torch numbers parameters in group order; so do we
        state = {int(k): v for k, v in sd["state"].items()}
        self._pending_state = {"state": {flat.index(k) if k in flat else k: v for k, v in state.items()}}
        live = [self.params[i] for i in sorted(self._pending_state["state"])]
        if live:
            self._built = False
            self._build(live)

    # -- internals -------------------------------------------------------------------------------
    @staticmethod
    def _real(t: torch.Tensor) -> torch.Tensor:
        return torch.view_as_real(t) if t.is_complex() else t

    def _build(self, live=None):
        if live is None:
            live = [p for p in self.params if p.grad is not None]
        self.live = live
        if not self.live:
            raise lib.FnoError("FusedClipAdam.step() before any backward()")
        sizes = [self._real(p).numel() for p in self.live]
        padded = [(n + 3) & ~3 for n in sizes]          # 16-byte aligned slots (complex views need even offsets)
        total = sum(padded)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=self.device)
        if self.own_grads:
            self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        L = lib.load()
        chunk = L.fno_opt_chunk_floats()
        rec = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i4"), ("lr0", "<f4")])
        assert rec.itemsize == L.fno_opt_chunk_bytes()
        index = {id(p): i for i, p in enumerate(self.params)}
        rows = []
        self._slots = []
        off = 0
        for p, n, npad in zip(self.live, sizes, padded):
            if not p.is_contiguous():
                raise lib.FnoError("FusedClipAdam: parameters must be contiguous")
            if self.own_grads:
                view = self.flat_grad[off:off + n]
                gview = torch.view_as_complex(view.view(*p.shape, 2)) if p.is_complex() else view.view(p.shape)
                if p.grad is not None:
                    gview.copy_(p.grad)
                p.grad = gview
            elif p.grad is None:
                raise lib.FnoError("FusedClipAdam(own_grads=False): every live parameter needs a .grad")
            gptr = self._real(p.grad).data_ptr()
            if not p.grad.is_contiguous():
                raise lib.FnoError("FusedClipAdam: gradients must be contiguous")
            pptr = p.data_ptr()
            lr0 = self.group_lr[self.group_of[index[id(p)]]]
            for c0 in range(0, n, chunk):
                cn = min(chunk, n - c0)
                rows.append((pptr + 4 * c0, gptr + 4 * c0, self.exp_avg.data_ptr() + 4 * (off + c0),
                             self.exp_avg_sq.data_ptr() + 4 * (off + c0), cn, lr0))
            self._slots.append((p, off, n))
            off += npad
        table = np.array(rows, dtype=rec)
        self.nchunks = len(rows)
        self.chunks = torch.from_numpy(table.view(np.uint8).copy()).to(self.device)
        self.partials = torch.zeros(self.nchunks, dtype=torch.float32, device=self.device)
        self._built = True
        if self._pending_state is not None:             # moments / step counter from a checkpoint
            st = self._pending_state["state"]
            step = None
            for p, off, n in self._slots:
                e = st.get(index[id(p)])
                if e is None:
                    continue
                self.exp_avg[off:off + n].copy_(self._real(e["exp_avg"].to(self.device)).reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(self._real(e["exp_avg_sq"].to(self.device)).reshape(-1))
                step = e["step"]
            if step is not None:
                self.state[0] = float(step.item() if torch.is_tensor(step) else step)
            self._pending_state = None
