"""Data-parallel TemporalTrainer on 2 GPUs (NCCL all-reduce inside the captured graph) == one rank stepping on the
union of the two ranks' videos.  GPU only; skipped with fewer than two devices (run: gpurun --gpus 2)."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

LENS = [[700, 333], [512, 129]]   # per rank: two videos each (equal counts => mean over ranks == mean over videos)
LENS_UNEQUAL = [[700, 333, 512], [129]]   # unequal shares (bandwidth-weighted sharding): loss normalised by the global count
D, C, STEPS = 96, 64, 3


def _model(dev, seed=0):
    from computervision_codes_b200.tcn import VideoNas

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(seed)
    return VideoNas(args, 4, 3, 3, C, D, 100).to(dev).train()


def _data(rank, step, lens=None):
    lens = LENS if lens is None else lens
    g = torch.Generator().manual_seed(100 * step + rank)
    n = sum(lens[rank])
    x = torch.randn(n, D, generator=g)
    lab = (torch.rand(n, 132, generator=g) < 0.05).to(torch.uint8)
    lab[:, 131] = 0
    return x, lab


def _worker(rank, world, port, out_dir, use_graph, unequal=False):
    lens = LENS_UNEQUAL if unequal else LENS
    gseq = sum(len(l) for l in lens) if unequal else None
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from computervision_codes_b200.trainer import TemporalTrainer

    # rank 1 starts from DIFFERENT weights: the trainer must broadcast rank 0's
    m = _model(dev, seed=0 if rank == 0 else 123)
    tr = TemporalTrainer(m, lr=0.05, weight_decay=1e-5, process_group=torch.distributed.group.WORLD, world_size=world,
                         max_frames=2048, max_seqs=4, use_graph=use_graph, input_mask_p=0.0)
    tr.training = False   # eval-mode arithmetic: dropout masks are keyed by the packed row, which differs between layouts
    losses = []
    for s in range(STEPS):
        x, lab = _data(rank, s, lens)
        out = tr.step(x.to(dev), lab.to(dev), lens[rank], global_seqs=gseq)
        losses.append(out.clone())
    torch.cuda.synchronize()
    torch.save({"p": tr.flat_p.cpu(), "loss": torch.stack(losses).cpu()}, os.path.join(out_dir, f"rank{rank}.pt"))
    tr.close()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("use_graph", [False, True])
def test_two_rank_trainer_equals_one_rank_on_the_union(tmp_path, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from computervision_codes_b200.trainer import TemporalTrainer

    port = 29500 + (os.getpid() % 400) + (7 if use_graph else 0)
    mp.spawn(_worker, args=(2, port, str(tmp_path), use_graph), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(2))
    assert torch.equal(r0["p"], r1["p"]), "ranks diverged"
    dev = torch.device("cuda", 0)
    m = _model(dev, seed=0)
    tr = TemporalTrainer(m, lr=0.05, weight_decay=1e-5, max_frames=4096, max_seqs=8, use_graph=False, input_mask_p=0.0)
    tr.training = False
    for s in range(STEPS):
        xs, labs = zip(*[_data(r, s) for r in range(2)])
        out = tr.step(torch.cat(xs).to(dev), torch.cat(labs).to(dev), LENS[0] + LENS[1])
        both = 0.5 * (r0["loss"][s] + r1["loss"][s])   # per-rank loss = mean over its 2 videos
        assert torch.allclose(out.cpu(), both, rtol=1e-4, atol=1e-6), (s, out.cpu(), both)
    ref = tr.flat_p.cpu()
    err = float((ref - r0["p"]).abs().max()) / max(1.0, float(ref.abs().max()))
    assert err <= 2e-5, err   # different layouts sum in different orders; a wrong normalisation would be off by a factor


def test_two_ranks_with_unequal_shares_equal_one_rank_on_the_union(tmp_path):
    """3 videos on rank 0, 1 on rank 1, loss normalised by the global count (global_seqs = 4): the summed gradient is the
    mean over the four videos -- the same parameters as one rank stepping on all four."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from computervision_codes_b200.trainer import TemporalTrainer

    port = 29500 + (os.getpid() % 400) + 19
    mp.spawn(_worker, args=(2, port, str(tmp_path), True, True), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(2))
    assert torch.equal(r0["p"], r1["p"]), "ranks diverged"
    dev = torch.device("cuda", 0)
    m = _model(dev, seed=0)
    tr = TemporalTrainer(m, lr=0.05, weight_decay=1e-5, max_frames=4096, max_seqs=8, use_graph=False, input_mask_p=0.0)
    tr.training = False
    for s in range(STEPS):
        xs, labs = zip(*[_data(r, s, LENS_UNEQUAL) for r in range(2)])
        out = tr.step(torch.cat(xs).to(dev), torch.cat(labs).to(dev), LENS_UNEQUAL[0] + LENS_UNEQUAL[1])
        both = r0["loss"][s] + r1["loss"][s]   # each rank's loss is its videos' sum over the GLOBAL count
        assert torch.allclose(out.cpu(), both, rtol=1e-4, atol=1e-6), (s, out.cpu(), both)
    ref = tr.flat_p.cpu()
    err = float((ref - r0["p"]).abs().max()) / max(1.0, float(ref.abs().max()))
    assert err <= 2e-5, err   # different layouts sum in different orders; a wrong normalisation would be off by a factor
