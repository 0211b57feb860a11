"""Batch-sharded data parallelism for the FNO training step.

The reference has no distributed code on this path; every op of ``FNO*.forward`` is per sample
and the loss is a batch mean, so replicas only need ONE exchange per step: the gradient average.
``BucketedGradAllReduce`` keeps each bucket's gradients as views into one flat fp32 buffer
(complex parameters through ``view_as_real``) and launches the bucket's all-reduce from a
post-accumulate-grad hook as soon as the last gradient of the bucket has been written, i.e.
while the backward kernels of earlier layers are still running; NCCL runs it on its own stream
over NVLink/NVSwitch.  ``finish()`` makes the compute stream wait for the reductions.

Parameters that never receive a gradient (FNO3d's dead ``bn*`` modules, fno.py:334-337) are left
with ``grad = None`` exactly as in the reference, so the optimizer skips them: the first step
discovers which parameters fire and in which order; buckets are built from that.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def fno_bucket_names(model) -> List[List[str]]:
    """Reverse-autograd groups: head, then Fourier layers 3..0, the lift joining layer 0."""
    names = [n for n, _ in model.named_parameters()]
    groups = [[n for n in names if n.startswith(("fc1", "fc2"))]]
    for layer in (3, 2, 1):
        groups.append([n for n in names if n.startswith((f"conv{layer}.", f"w{layer}."))])
    groups.append([n for n in names if n.startswith(("conv0.", "w0.", "fc0"))])
    used = {n for g in groups for n in g}
    rest = [n for n in names if n not in used]
    if rest:
        groups.append(rest)
    return [g for g in groups if g]


def _real_view(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


class _Bucket:
    @staticmethod
    def floats(params) -> int:
        return sum((_real_view(p).numel() + 3) & ~3 for p in params)

    def __init__(self, params: Sequence[torch.nn.Parameter], storage: Optional[torch.Tensor] = None):
        self.params = list(params)
        dev = self.params[0].device
        sizes = [_real_view(p).numel() for p in self.params]
        padded = [(n + 3) & ~3 for n in sizes]          # 16-byte aligned slots (complex views need even offsets)
        # `storage`: a slice of one buffer shared by all buckets, so that the inline mode can exchange everything in ONE call
        self.flat = storage if storage is not None else torch.zeros(sum(padded), dtype=torch.float32, device=dev)
        off = 0
        for p, n, npad in zip(self.params, sizes, padded):
            chunk = self.flat[off:off + n]
            if p.is_complex():
                view = torch.view_as_complex(chunk.view(*p.shape, 2))
            else:
                view = chunk.view(p.shape)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            off += npad
        self.pending = len(self.params)
        self.work = None


class BucketedGradAllReduce:
    def __init__(self, model: torch.nn.Module, bucket_names: Optional[List[List[str]]] = None,
                 group=None, average: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.average = average
        self.named = dict(model.named_parameters())      # de-duplicated (fno_aux shared_layers alias)
        self.bucket_names = bucket_names if bucket_names is not None else fno_bucket_names(model)
        self.buckets: Optional[List[_Bucket]] = None
        self._bucket_of = {}
        self._hooks = []
        backend = dist.get_backend(group)
        self._op = dist.ReduceOp.AVG if (average and backend == "nccl") else dist.ReduceOp.SUM
        self._scale_after = average and backend != "nccl"
        self.launched_in_backward = 0
        self.inline = False      # True: reductions are issued in finish(), blocking, on the current stream (CUDA graphs)

    def set_inline(self, inline: bool = True):
        """CUDA-graph mode: no autograd hooks, no async work handles -- ``finish()`` enqueues one all-reduce per bucket on
        the CURRENT stream, between the replay of the forward/backward graph and the replay of the optimizer graph
        (fno_b200.train.FusedTrainStep._capture).  At cfg 1 the whole exchange is 3.7 MB: ~0.1 ms un-overlapped, against
        the ~0.45 ms an eager step loses to the launch overhead of its ~150 kernels."""
        self.inline = inline
        if inline:
            self.remove_hooks()
        elif self.buckets is not None and not self._hooks:
            for b in self.buckets:
                for p in b.params:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def zero_grad(self):
        """Replaces optimizer.zero_grad(): keeps the flat views alive (one memset per bucket)."""
        if self.buckets is None:
            for p in self.named.values():
                p.grad = None
            return
        for b in self.buckets:
            b.flat.zero_()
            b.pending = len(b.params)
            b.work = None
        self.launched_in_backward = 0

    def finish(self):
        """Call after backward(): waits for (or, on the discovery step, performs) the reductions."""
        if self.buckets is None:
            self._discover_and_reduce()
            return
        if self.inline:
            # one collective over the buffer all buckets live in (3.7 MB at cfg 1): latency, not bandwidth, is the cost
            dist.all_reduce(self._all, op=self._op, group=self.group)
            if self._scale_after:
                self._all.mul_(1.0 / self.world)
            return
        for b in self.buckets:
            if b.work is None:          # a parameter of the bucket did not fire this step
                b.work = dist.all_reduce(b.flat, op=self._op, group=self.group, async_op=True)
        for b in self.buckets:
            b.work.wait()
            if self._scale_after:
                b.flat.mul_(1.0 / self.world)

    # -- internals ------------------------------------------------------------------------------
    def _discover_and_reduce(self):
        live = {n: p for n, p in self.named.items() if p.grad is not None}
        groups = [g for g in ([live[n] for n in names if n in live] for names in self.bucket_names) if g]
        dev = groups[0][0].device
        self._all = torch.zeros(sum(_Bucket.floats(g) for g in groups), dtype=torch.float32, device=dev)
        off = 0
        self.buckets = []
        for g in groups:
            n = _Bucket.floats(g)
            self.buckets.append(_Bucket(g, self._all[off:off + n]))
            off += n
        for bi, b in enumerate(self.buckets):
            for p in b.params:
                self._bucket_of[p] = bi
                if not self.inline:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        works = [dist.all_reduce(b.flat, op=self._op, group=self.group, async_op=True) for b in self.buckets]
        for b, w in zip(self.buckets, works):
            w.wait()
            if self._scale_after:
                b.flat.mul_(1.0 / self.world)

    def _on_grad(self, p):
        b = self.buckets[self._bucket_of[p]]
        b.pending -= 1
        if b.pending == 0 and b.work is None:
            b.work = dist.all_reduce(b.flat, op=self._op, group=self.group, async_op=True)
            self.launched_in_backward += 1

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def grad_bytes(self) -> int:
        return 0 if self.buckets is None else sum(b.flat.numel() * 4 for b in self.buckets)
