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


def _mk_params(g, shapes):
    ref, our = [], []
    for shp, cplx in shapes:
        t = torch.randn(shp, generator=g, dtype=torch.cfloat if cplx else torch.float32)
        ref.append(torch.nn.Parameter(t.clone().cuda()))
        our.append(torch.nn.Parameter(t.clone().cuda()))
    return ref, our


def _set_grads(ref_params, our_params, grads):
    for p, q, gr in zip(ref_params, our_params, grads):
        p.grad = gr.clone().cuda()
        if q.grad is None:
            q.grad = gr.clone().cuda()
        else:
            q.grad.copy_(gr.cuda())


def _close(ref_params, our_params, tol, what):
    for p, q in zip(ref_params, our_params):
        a, b = (torch.view_as_real(t) if t.is_complex() else t for t in (p, q))
        assert O.rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < tol, (what, tuple(p.shape))


def test_fused_clip_adam_param_groups_and_extra_scheduler_steps():
    """The joint loop's optimizer (fno_aux/fno_train_aux.py:175-186): three Adam groups with two learning rates under
    ONE CosineAnnealingLR that is stepped after every iteration and once more per "epoch" (:329, :398)."""
    from fno_b200.steptail import FusedClipAdam

    g = torch.Generator().manual_seed(5)
    shapes = [((6, 4), False), ((3, 3, 2, 2), True), ((11,), False), ((2, 5), False), ((7,), False)]
    ref, our = _mk_params(g, shapes)
    groups = lambda ps: [{"params": ps[:3], "lr": 3e-3, "weight_decay": 1e-4},       # noqa: E731
                         {"params": ps[3:4], "lr": 1e-2, "weight_decay": 1e-4},
                         {"params": ps[4:], "lr": 1e-2, "weight_decay": 1e-4}]
    opt_ref = torch.optim.Adam(groups(ref))
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt_ref, T_max=12.5)              # float T_max as in the loop
    opt = FusedClipAdam(groups(our), weight_decay=1e-4, t_max=12.5)
    for step in range(8):
        grads = [30.0 * torch.randn(p.shape, generator=g, dtype=p.dtype) for p in ref]
        opt_ref.zero_grad()
        opt.zero_grad()
        _set_grads(ref, our, grads)
        _reference_tail(ref, opt_ref, sched)
        opt.step()
        if step % 3 == 2:                       # the per-epoch scheduler step
            sched.step()
            opt.scheduler_step()
        _close(ref, our, 3e-6, step)
    # the learning rates a checkpoint would record
    lrs = [gr["lr"] for gr in opt.state_dict()["param_groups"]]
    for a, b in zip(lrs, [gr["lr"] for gr in opt_ref.param_groups]):
        assert abs(a - b) < 1e-6 * abs(b) + 1e-12


def test_fused_clip_adam_state_dict_round_trips_with_torch_adam():
    """Checkpoint contract of the reference loop (fno/train.py:189-204 resume, :319-329 save): the fused optimizer's
    state_dict loads into torch.optim.Adam and vice versa, and training continues identically."""
    from fno_b200.steptail import FusedClipAdam

    g = torch.Generator().manual_seed(7)
    shapes = [((5, 4), False), ((2, 3, 2, 2), True), ((9,), False)]
    ref, our = _mk_params(g, shapes)
    dead_ref, dead_our = torch.nn.Parameter(torch.ones(3).cuda()), torch.nn.Parameter(torch.ones(3).cuda())
    opt_ref = torch.optim.Adam(ref + [dead_ref], lr=2e-3, weight_decay=1e-4)
    opt = FusedClipAdam(our + [dead_our], lr=2e-3, weight_decay=1e-4)

    def run(o_ref, o_our, ps_ref, ps_our, n):
        for _ in range(n):
            grads = [torch.randn(p.shape, generator=g, dtype=p.dtype) for p in ps_ref]
            o_ref.zero_grad()
            o_our.zero_grad()
            _set_grads(ps_ref, ps_our, grads)
            total = torch.norm(torch.stack([torch.norm(p.grad.detach(), 2) for p in ps_ref]), 2)
            torch.nn.utils.clip_grad_norm_(ps_ref, max(5, 0.1 * total))        # fno/train.py:273-275
            o_ref.step()
            o_our.step()

    run(opt_ref, opt, ref, our, 3)
    _close(ref, our, 2e-6, "before")
    sd_fused, sd_torch = opt.state_dict(), opt_ref.state_dict()
    assert sorted(sd_fused["state"]) == sorted(sd_torch["state"]) == [0, 1, 2]          # the dead parameter has no state
    for i in range(3):
        for k in ("exp_avg", "exp_avg_sq"):
            a, b = sd_fused["state"][i][k], sd_torch["state"][i][k]
            assert a.shape == b.shape and a.dtype == b.dtype
            assert O.rel_err(torch.view_as_real(a).cpu().numpy() if a.is_complex() else a.cpu().numpy(),
                             torch.view_as_real(b).cpu().numpy() if b.is_complex() else b.cpu().numpy()) < 2e-6
        assert float(sd_fused["state"][i]["step"]) == float(sd_torch["state"][i]["step"]) == 3.0
    # cross-load: fused -> fresh torch Adam, torch -> fresh fused optimizer, then two more steps on both sides
    ref2 = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    our2 = [torch.nn.Parameter(p.detach().clone()) for p in our]
    dead2a, dead2b = torch.nn.Parameter(torch.ones(3).cuda()), torch.nn.Parameter(torch.ones(3).cuda())
    t2 = torch.optim.Adam(ref2 + [dead2a], lr=2e-3, weight_decay=1e-4)
    t2.load_state_dict({k: v for k, v in sd_fused.items() if k != "fno_b200"})
    f2 = FusedClipAdam(our2 + [dead2b], lr=2e-3, weight_decay=1e-4)
    f2.load_state_dict(sd_torch)
    run(t2, f2, ref2, our2, 2)
    _close(ref2, our2, 3e-6, "after cross-load")
    # ... and they continue exactly like the optimizers that were never checkpointed
    g = torch.Generator().manual_seed(7)
    ref3, our3 = _mk_params(g, shapes)
    o3 = torch.optim.Adam(ref3, lr=2e-3, weight_decay=1e-4)
    f3 = FusedClipAdam(our3, lr=2e-3, weight_decay=1e-4)
    run(o3, f3, ref3, our3, 5)
    _close(ref3, our2, 5e-6, "uninterrupted vs resumed")
