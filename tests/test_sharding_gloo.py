"""CPU, world_size 2 over gloo: the N>1 host logic (shard bounds, counter and gradient all-reduces)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_bounds_cover_exactly():
    from neural_ldpc_decoder_torch_b200.sharding import shard_bounds
    for total in (0, 1, 7, 1 << 20, 65536 * 3 + 5):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neural_ldpc_decoder_torch_b200.sharding import allreduce_mean_grads_, allreduce_sum_, count_errors_packed, shard_bounds
    total = 1001
    b, e = shard_bounds(total, world, rank)
    rs = np.random.RandomState(0)
    bits = rs.randint(0, 2, size=(total, 832)).astype(np.uint8)
    noisy = bits.copy()
    flips = rs.randint(0, total, size=50)
    noisy[flips, rs.randint(0, 832, size=50)] ^= 1
    hard = torch.from_numpy(np.packbits(noisy[b:e], axis=1, bitorder="little"))
    exp = torch.from_numpy(np.packbits(bits[b:e], axis=1, bitorder="little"))
    cnt = allreduce_sum_(count_errors_packed(hard, exp))
    ref_bits = int((noisy != bits).sum())
    ref_frames = int(((noisy != bits).sum(axis=1) > 0).sum())
    ok = cnt.tolist() == [ref_bits, ref_frames, total * 832, total]
    # gradient all-reduce: mean over ranks of the local gradients, every rank ends with the same vector
    p1, p2 = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))
    p1.grad = torch.full((3,), float(rank + 1))
    p2.grad = torch.full((2, 2), float(10 * (rank + 1)))
    allreduce_mean_grads_([p1, p2])
    ok = ok and torch.equal(p1.grad, torch.full((3,), 1.5)) and torch.equal(p2.grad, torch.full((2, 2), 15.0))
    q.put((rank, ok, cnt.tolist()))
    dist.destroy_process_group()


def test_two_rank_counters_and_grad_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
