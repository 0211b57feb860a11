"""Data-parallel host logic on CPU: world_size-2 gloo processes, a small torch model with a complex
parameter and a never-used parameter (mirrors FNO3d's dead bn*).  Checks that the bucketed,
hook-driven all-reduce reproduces single-process gradients of the global batch and that buckets
launched from hooks give bit-identical parameters to reducing everything at the end."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sciml-pde_b200"))


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc0 = torch.nn.Linear(3, 4)
        self.weights1 = torch.nn.Parameter(torch.rand(4, 4, dtype=torch.cfloat))
        self.fc1 = torch.nn.Linear(4, 2)
        self.dead = torch.nn.Parameter(torch.ones(3))          # never used -> grad stays None

    def forward(self, x):
        h = self.fc0(x)
        h = torch.view_as_real(h.to(torch.cfloat) @ self.weights1).sum(-1)
        return self.fc1(torch.tanh(h))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fno_b200.dp import BucketedGradAllReduce

    torch.manual_seed(16)
    model = Tiny()
    buckets = [["fc1.weight", "fc1.bias"], ["weights1", "fc0.weight", "fc0.bias", "dead"]]
    dp = BucketedGradAllReduce(model, buckets)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-4)
    g = torch.Generator().manual_seed(0)
    launched = []
    for step in range(3):
        x = torch.randn(8, 3, generator=g)
        y = torch.randn(8, 2, generator=g)
        xs, ys = x[rank::world], y[rank::world]
        loss = ((model(xs) - ys) ** 2).mean()
        dp.zero_grad()
        if not overlap and dp.buckets is not None:
            dp.remove_hooks()
        loss.backward()
        dp.finish()
        launched.append(dp.launched_in_backward)
        opt.step()
    assert model.dead.grad is None
    if rank == 0:
        torch.save({"sd": model.state_dict(), "launched": launched,
                    "grads": {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}},
                   os.path.join(out_dir, f"dp_{int(overlap)}.pt"))
    dist.destroy_process_group()


def _single(steps=3):
    torch.manual_seed(16)
    model = Tiny()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-4)
    g = torch.Generator().manual_seed(0)
    for _ in range(steps):
        x = torch.randn(8, 3, generator=g)
        y = torch.randn(8, 2, generator=g)
        # mean over the global batch == mean of the two rank means (equal shard sizes)
        loss = 0.5 * (((model(x[0::2]) - y[0::2]) ** 2).mean() + ((model(x[1::2]) - y[1::2]) ** 2).mean())
        opt.zero_grad()
        loss.backward()
        opt.step()
    return model


@pytest.mark.timeout(300)
def test_bucketed_allreduce_world2_gloo(tmp_path):
    world = 2
    for overlap in (True, False):
        mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), overlap), nprocs=world, join=True)
    a = torch.load(tmp_path / "dp_1.pt")
    b = torch.load(tmp_path / "dp_0.pt")
    # hook-launched buckets vs everything reduced in finish(): bit-identical
    for k in a["sd"]:
        assert torch.equal(a["sd"][k], b["sd"][k]), k
    assert a["launched"][0] == 0 and a["launched"][1] == 2 and a["launched"][2] == 2
    assert b["launched"] == [0, 0, 0]
    ref = _single()
    for k, v in ref.state_dict().items():
        assert torch.allclose(a["sd"][k], v, rtol=1e-5, atol=1e-6), k
    assert "dead" not in a["grads"]


def test_fno_bucket_names_cover_every_parameter_once():
    from fno_b200 import fno, fno_aux
    from fno_b200.dp import fno_bucket_names

    for model in (fno.FNO2d(2, 4, 4, 8, 3), fno.FNO3d(2, 2, 2, 2, 4, 2), fno_aux.FNO2d(2, 4, 4, 8, 3)):
        groups = fno_bucket_names(model)
        flat = [n for g in groups for n in g]
        assert sorted(flat) == sorted(n for n, _ in model.named_parameters())
        assert len(flat) == len(set(flat))
        assert groups[0][0].startswith("fc1") and any(n.startswith("fc0") for n in groups[4])


# ------------------------------------------------------------------------------------------------
# ragged last global batch (DataLoader(drop_last=False), fno/train.py:95-97): unequal and EMPTY rank slices
# ------------------------------------------------------------------------------------------------
def test_epoch_plan_every_rank_every_step():
    from fno_b200.data import epoch_indices, epoch_plan

    n_items, batch, world = 21, 8, 4                       # global batches 8, 8, 5 -> tail slices 2, 1, 1, 1
    for n_items in (21, 17, 8):                            # 17: tail of 1 item < world -> three empty slices
        plans = [epoch_plan(n_items, batch, True, 16, 0, r, world) for r in range(world)]
        steps = len(plans[0])
        assert all(len(p) == steps for p in plans) and steps == -(-n_items // batch)
        for s in range(steps):
            assert all(p[s][0].numel() >= 1 for p in plans)                     # nobody skips a step
            gb = min(batch, n_items - s * batch)
            assert abs(sum(p[s][1] for p in plans) - world) < 1e-12                 # weights average to 1
            real = [p[s][0] for p in plans if p[s][1] > 0]
            assert sum(t.numel() for t in real) == gb
        # union over ranks of the weighted items = the single-process order
        single = torch.cat(epoch_indices(n_items, batch, True, 16, 0))
        multi = torch.cat([torch.cat([p[s][0] for p in plans if p[s][1] > 0]) for s in range(steps)])
        assert sorted(single.tolist()) == sorted(multi.tolist())


def _ragged_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fno_b200.data import epoch_plan
    from fno_b200.dp import BucketedGradAllReduce

    torch.manual_seed(16)
    model = Tiny()
    dp = BucketedGradAllReduce(model, [["fc1.weight", "fc1.bias"], ["weights1", "fc0.weight", "fc0.bias", "dead"]])
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(0)
    X, Y = torch.randn(9, 3, generator=g), torch.randn(9, 2, generator=g)      # 9 items, global batch 4: tail of 1 item
    for items, w in epoch_plan(9, 4, True, 16, 0, rank, world):
        loss = ((model(X[items]) - Y[items]) ** 2).mean() * w
        dp.zero_grad()
        loss.backward()
        dp.finish()
        opt.step()
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(out_dir, "ragged.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_ragged_tail_world2_gloo_matches_single_process(tmp_path):
    from fno_b200.data import epoch_indices

    mp.spawn(_ragged_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = torch.load(tmp_path / "ragged.pt")
    torch.manual_seed(16)
    model = Tiny()
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(0)
    X, Y = torch.randn(9, 3, generator=g), torch.randn(9, 2, generator=g)
    for items in epoch_indices(9, 4, True, 16, 0):
        loss = ((model(X[items]) - Y[items]) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    for k, v in model.state_dict().items():
        assert torch.allclose(got[k], v, rtol=1e-5, atol=1e-6), k
