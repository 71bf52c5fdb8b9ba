"""Data-parallel-by-video host logic on CPU with gloo, world_size 2: sharding of the schedule, the single
all-reduce of the flat gradient buffer and the SGD rule  ==  one rank accumulating both ranks' batches."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat_grad(params, x, labels):
    from oracle import torch_port as P

    for p in params.values():
        p.grad = None
    loss = P.train_step_loss(x, params, labels, train=False)
    loss.backward()
    names = sorted(n for n, p in params.items() if p.grad is not None)
    return names, torch.cat([params[n].grad.reshape(-1) for n in names]), float(loss.detach())


def _make(seed, T):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, T, 12, generator=g)
    labels = tuple((torch.rand(T, k, generator=g) < 0.1).float() for k in (6, 10, 15, 100))
    return x, labels


def _params():
    import types

    sys.path.insert(0, ROOT)
    from computervision_codes_b200.tcn import VideoNas

    torch.manual_seed(0)
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    m = VideoNas(args, 2, 2, 3, 8, 12, 100)
    return {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    params = _params()
    x, labels = _make(100 + rank, 40 + 7 * rank)  # each rank owns one video of its own length
    names, flat, _ = _flat_grad(params, x, labels)
    dist.all_reduce(flat)                          # the ONE collective of a step
    lr, wd = 0.05, 1e-5
    p0 = torch.cat([params[n].detach().reshape(-1) for n in names])
    new = p0 - lr * (flat / world + wd * p0)       # tcn_sgd_step with grad_scale = 1 / world
    if rank == 0:
        torch.save({"names": names, "new": new}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dp_equals_single_rank_accumulation(tmp_path):
    out = str(tmp_path / "dp.pt")
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    params = _params()
    acc = None
    for r in range(2):
        names, flat, _ = _flat_grad(params, *_make(100 + r, 40 + 7 * r))
        acc = flat if acc is None else acc + flat
    p0 = torch.cat([params[n].detach().reshape(-1) for n in names])
    ref = p0 - 0.05 * (acc / 2 + 1e-5 * p0)
    assert got["names"] == names
    assert torch.allclose(got["new"], ref, atol=1e-7, rtol=1e-6)


def test_lpt_assignment_balances_frames():
    sys.path.insert(0, ROOT)
    from computervision_codes_b200.trainer import lpt_assign

    lengths = [3500, 900, 1200, 3000, 2500, 1000, 2000, 1800, 950, 3100]
    for world in (1, 2, 4, 8):
        shards = lpt_assign(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)


def test_bench_schedule_is_5_folds_of_31_videos():
    sys.path.insert(0, ROOT)
    import bench

    passes, lengths = bench.fold_schedule()
    assert len(passes) == 155 and len(lengths) == 45
    assert all(900 <= t < 3600 for t in lengths.values())
    for world in (1, 2, 4, 8):
        shards = [passes[r::world] for r in range(world)]
        assert sum(len(s) for s in shards) == 155
