"""BASELINE.json configs[1..4] at (or near) full size on the GPU against the CPU oracle port
(oracle/fno_port.py, fp64): the two-head joint model (cfg 2), FNO2d at modes 16 / width 64 on
256x256 (cfg 3), FNO3d at modes 12 / width 20 on 64^3 x 5 channels (cfg 4), and the
autoregressive rollout used by the cfg-5 evaluation.  Tolerance: max|err| / max|ref| <= 1e-5 on
outputs, 2e-5 on gradients (fp32 mode, BASELINE north_star)."""
import pytest
import torch

from oracle import dft_oracle as O
from oracle import fno_port as P

pytestmark = pytest.mark.gpu
TOL = 1e-5


def seeded(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return scale * torch.randn(shape, generator=g)


def _compare(model, p64, out, ref, gtol=2e-5, skip_prefix="bn"):
    assert O.rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    checked = 0
    for k, q in model.named_parameters():
        if q.grad is None:
            assert k.startswith(skip_prefix), k
            continue
        a = torch.view_as_real(q.grad) if q.grad.is_complex() else q.grad
        b = torch.view_as_real(p64[k].grad) if p64[k].grad.is_complex() else p64[k].grad
        assert O.rel_err(a.cpu().numpy(), b.numpy()) < gtol, k
        checked += 1
    assert checked >= 20


def test_cfg3_fno2d_modes16_width64_256():
    from fno_b200.fno import FNO2d

    torch.manual_seed(16)
    model = FNO2d(num_channels=3, modes1=16, modes2=16, width=64, initial_step=10)
    p64 = P.as_leaves(model.state_dict(), dtype=torch.float64)
    x = seeded((1, 256, 256, 10, 3), 300)
    lin = torch.linspace(0, 1, 256)
    grid = torch.stack(torch.meshgrid(lin, lin, indexing="ij"), dim=-1).unsqueeze(0)
    yy = seeded((1, 256, 256, 1, 3), 301)
    ref = P.fno_forward(p64, x.double(), grid.double())
    P.nrmse(ref, yy.double()).mean().backward()
    model = model.cuda()
    out = model(x.cuda(), grid.cuda())
    P.nrmse(out, yy.cuda()).mean().backward()
    _compare(model, p64, out, ref)


def test_cfg4_fno3d_modes12_width20_64cubed():
    from fno_b200.fno import FNO3d

    torch.manual_seed(16)
    model = FNO3d(num_channels=5, modes1=12, modes2=12, modes3=12, width=20, initial_step=10)
    p64 = P.as_leaves(model.state_dict(), dtype=torch.float64)
    x = seeded((1, 64, 64, 64, 10, 5), 400)
    lin = torch.linspace(0, 1, 64)
    grid = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), dim=-1).unsqueeze(0)
    yy = seeded((1, 64, 64, 64, 1, 5), 401)
    ref = P.fno_forward(p64, x.double(), grid.double())
    P.nrmse(ref, yy.double()).mean().backward()
    model = model.cuda()
    out = model(x.cuda(), grid.cuda())
    P.nrmse(out, yy.cuda()).mean().backward()
    _compare(model, p64, out, ref)


def test_cfg2_fno_aux_joint_step_128():
    """fno_aux joint training step: primary batch 2 + 3 auxiliary trajectories each (config_dr.yaml),
    loss = primary + 0.7 * auxiliary (fno_train_aux.py:315)."""
    from fno_b200.fno_aux import FNO2d

    torch.manual_seed(16)
    model = FNO2d(num_channels=2, modes1=12, modes2=12, width=20, initial_step=10)
    p64 = P.as_leaves({k: v for k, v in model.state_dict().items() if not k.startswith("shared_layers")},
                      dtype=torch.float64)
    B, A = 2, 3
    x, xa = seeded((B, 128, 128, 10, 2), 500), seeded((B * A, 128, 128, 10, 2), 501, scale=0.5)
    lin = torch.linspace(-1 + 1 / 128, 1 - 1 / 128, 128)
    g1 = torch.stack(torch.meshgrid(lin, lin, indexing="ij"), dim=-1)
    grid, grid_a = g1.unsqueeze(0).repeat(B, 1, 1, 1), g1.unsqueeze(0).repeat(B * A, 1, 1, 1)
    yy, yya = seeded((B, 128, 128, 1, 2), 502), seeded((B * A, 128, 128, 1, 2), 503)
    rp, ra = P.fno_aux_forward(p64, x.double(), grid.double(), xa.double(), grid_a.double())
    (P.nrmse(rp, yy.double()).mean() + 0.7 * P.nrmse(ra, yya.double()).mean()).backward()
    model = model.cuda()
    op, oa = model(x.cuda(), grid.cuda(), xa.cuda(), grid_a.cuda())
    (P.nrmse(op, yy.cuda()).mean() + 0.7 * P.nrmse(oa, yya.cuda()).mean()).backward()
    assert O.rel_err(oa.detach().cpu().numpy(), ra.detach().numpy()) < TOL
    named = {k: q for k, q in model.named_parameters()}            # de-duplicated: 24 unique tensors
    assert len(named) == 24
    _compare(model, p64, op, rp)


def test_rollout_matches_manual_autoregression():
    from fno_b200.evaluate import evaluate_rollout, rollout
    from fno_b200.fno import FNO2d

    torch.manual_seed(16)
    model = FNO2d(num_channels=2, modes1=4, modes2=4, width=8, initial_step=3)
    p64 = P.as_leaves(model.state_dict(), dtype=torch.float64)
    model = model.cuda().eval()
    xx, grid = seeded((2, 16, 16, 3, 2), 600), torch.rand(2, 16, 16, 2, generator=torch.Generator().manual_seed(601))
    yy = seeded((2, 16, 16, 5, 2), 602)
    preds = rollout(model, xx.cuda(), grid.cuda(), 5)
    cur, ref = xx.double(), []
    with torch.no_grad():
        for _ in range(5):
            pr = P.fno_forward(p64, cur, grid.double())
            ref.append(pr)
            cur = torch.cat((cur[..., 1:, :], pr), dim=-2)
    ref = torch.cat(ref, dim=-2)
    assert preds.shape == (2, 16, 16, 5, 2)
    assert O.rel_err(preds.cpu().numpy(), ref.numpy()) < 5e-5     # five chained forwards
    res = evaluate_rollout(model, [(xx.cuda(), yy.cuda(), grid.cuda())], 5)
    err = torch.sqrt(((ref[..., -1:, :] - yy.double()[..., -1:, :]) ** 2).mean((1, 2)))
    nrm = torch.sqrt((yy.double()[..., -1:, :] ** 2).mean((1, 2)))
    assert abs(res["nrmse_last"] - float((err / nrm).mean())) < 1e-4 * float((err / nrm).mean())
    assert res["samples_x_vars"] == 4
