"""Step tail on the GPU: fused nRMSE loss vs the oracle port, fused clip + Adam +
cosine LR vs the reference's torch sequence (fno/train.py:251-259, :271-278), and the CUDA-graph
training step vs the eager torch step."""
import copy

import numpy as np
import pytest
import torch

from oracle import dft_oracle as O
from oracle import fno_port as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3, 16, 16, 1, 2), (2, 9, 7, 1, 3), (2, 6, 5, 4, 1, 5), (4, 8, 8, 3, 2)])
def test_nrmse_loss_vs_oracle(shape):
    from fno_b200.steptail import nrmse_loss

    g = torch.Generator().manual_seed(7)
    out = torch.randn(shape, generator=g)
    tgt = torch.randn(shape, generator=g) * 2 + 0.5
    o64 = out.double().requires_grad_()
    ref = P.nrmse(o64, tgt.double()).mean()
    ref.backward()
    oc = out.cuda().requires_grad_()
    loss = nrmse_loss(oc, tgt.cuda())
    (3.0 * loss).backward()
    assert abs(loss.item() - ref.item()) < 1e-6 * abs(ref.item())
    assert O.rel_err(oc.grad.cpu().numpy(), 3.0 * o64.grad.numpy()) < 1e-5


def _reference_tail(params, opt, sched):
    """fno/train.py:251-259, :273-278 (the host-side max() included)."""
    norms = [torch.norm(p.grad.detach(), 2) for p in params if p.grad is not None]
    total = torch.norm(torch.stack(norms), 2)
    clip_value = max(5, 0.1 * total)
    torch.nn.utils.clip_grad_norm_(params, clip_value)
    opt.step()
    sched.step()
    return float(total)


@pytest.mark.parametrize("scale", [0.01, 30.0, 400.0])   # no clipping / clip at 5 / clip at 0.1 * norm
def test_fused_clip_adam_vs_torch(scale):
    from fno_b200.steptail import FusedClipAdam

    g = torch.Generator().manual_seed(3)
    shapes = [((7, 5), False), ((4, 3, 2, 2), True), ((9,), False), ((5000,), False), ((3, 3), False)]
    ref_params, our_params = [], []
    for shp, cplx in shapes:
        t = torch.randn(shp, generator=g, dtype=torch.cfloat if cplx else torch.float32)
        ref_params.append(torch.nn.Parameter(t.clone().cuda()))
        our_params.append(torch.nn.Parameter(t.clone().cuda()))
    dead_ref = torch.nn.Parameter(torch.ones(4).cuda())      # never receives a gradient
    dead_our = torch.nn.Parameter(torch.ones(4).cuda())
    opt_ref = torch.optim.Adam(ref_params + [dead_ref], lr=1e-2, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt_ref, T_max=10)
    opt = FusedClipAdam(our_params + [dead_our], lr=1e-2, weight_decay=1e-4, t_max=10)
    for step in range(6):
        grads = [scale * torch.randn(p.shape, generator=g, dtype=p.dtype) for p in ref_params]
        opt_ref.zero_grad()
        opt.zero_grad()
        for p, q, gr in zip(ref_params, our_params, grads):
            p.grad = gr.clone().cuda()
            if q.grad is None:
                q.grad = gr.clone().cuda()
            else:
                q.grad.copy_(gr.cuda())
        total = _reference_tail(ref_params + [dead_ref], opt_ref, sched)
        opt.step()
        assert abs(float(opt.total_norm) - total) < 1e-5 * total
        for p, q in zip(ref_params, our_params):
            a, b = torch.view_as_real(p) if p.is_complex() else p, torch.view_as_real(q) if q.is_complex() else q
            assert O.rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-6, (step, tuple(p.shape))
    assert torch.equal(dead_our, dead_ref)


@pytest.mark.parametrize("graph", [False, True, "alias"])
def test_fused_train_step_matches_torch_step(graph):
    from fno_b200 import data
    from fno_b200.fno import FNO2d
    from fno_b200.train import FusedTrainStep, TrainStep

    torch.manual_seed(16)
    m_ref = FNO2d(num_channels=2, modes1=4, modes2=4, width=8, initial_step=3).cuda()
    m_our = copy.deepcopy(m_ref)
    opt = torch.optim.Adam(m_ref.parameters(), lr=1e-3, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50)
    ref_step = TrainStep(m_ref, opt, sched)
    # "alias": the graph reads the caller's input tensors in place, one captured graph per input set (two allowed
    # here, the third batch goes through the copying graph)
    our_step = FusedTrainStep(m_our, lr=1e-3, weight_decay=1e-4, t_max=50, graph=bool(graph),
                              alias_inputs=(graph == "alias"), max_graphs=2)
    batches = [tuple(t.cuda() for t in data.synthetic_batch(4, 16, 3, 2, seed=s)) for s in range(3)]
    for it in range(10 if graph == "alias" else 7):
        xx, yy, grid = batches[it % 3]
        l_ref = ref_step(xx, yy, grid)
        l_our = our_step(xx, yy, grid)
        assert abs(float(l_our) - float(l_ref)) < 2e-5 * abs(float(l_ref)), it
    for (n, p), q in zip(m_ref.named_parameters(), m_our.parameters()):
        a, b = (torch.view_as_real(t) if t.is_complex() else t for t in (p, q))
        assert O.rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 5e-5, n
