"""CPU: host-side data-parallel logic with world_size 2 over gloo (the N>1 path of bench.py / TrainStep
that does not need a GPU): image sharding, flat-parameter views, mean all-reduce of a flat gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xmm_superres_denoise_b200.models import GeneratorRRDB_DN
from xmm_superres_denoise_b200.training import allreduce_mean_, flatten_parameters, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flatten_parameters_keeps_names_values_and_aliases_storage():
    torch.manual_seed(0)
    m = GeneratorRRDB_DN(1, 1, 32, 1)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat = flatten_parameters(m)
    assert flat.numel() == sum(p.numel() for p in m.parameters())
    after = m.state_dict()
    assert list(after.keys()) == list(before.keys())
    assert all(torch.equal(after[k], before[k]) for k in before)
    flat.zero_()
    assert all(float(p.abs().sum()) == 0 for p in m.parameters())


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = GeneratorRRDB_DN(1, 1, 32, 1)
        flat = flatten_parameters(m)
        grad = torch.full_like(flat, float(rank + 1))
        grad[rank::7] += 1.0
        want = torch.full_like(flat, 1.5)
        want[0::7] += 0.5
        want[1::7] += 0.5
        allreduce_mean_(grad)
        ok = torch.allclose(grad, want)
        lo, hi = shard_range(5, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi))
        # validation-state reduction (CompositeLoss / metric collection): the accumulated sums are summed over the
        # ranks and PSNR's running target range is the global min / max, as torchmetrics' dist_reduce_fx does --
        # not a mean of per-rank values
        from xmm_superres_denoise_b200.loss import CompositeLoss

        loss = CompositeLoss({"l1": 1.0, "psnr": 1.0})
        loss._acc.update({"abs": torch.tensor(2.0 + rank), "sq": torch.tensor(0.5 * (rank + 1)), "n": 10 * (rank + 1), "b": 1,
                          "min_t": torch.tensor(0.0), "max_t": torch.tensor(0.5 + 0.25 * rank)})
        total = float(loss.compute())
        out.put((rank, bool(ok), gathered, total))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allreduce_and_sharding():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import math

    l1 = (2.0 + 3.0) / 30.0                                    # sum |e| over both ranks / all pixels
    psnr = 10.0 * math.log10(0.75 ** 2 / ((0.5 + 1.0) / 30.0))  # global range (max over ranks), global MSE
    for rank, ok, gathered, total in results:
        assert ok, f"rank {rank}: all-reduce mean mismatch"
        assert gathered == [(0, 3), (3, 5)]
        assert abs(total - (l1 + psnr)) < 1e-4, (rank, total, l1 + psnr)
