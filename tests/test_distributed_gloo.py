"""CPU, world_size 2 (gloo): the multi-GPU host logic -- sharding by image, loss all-reduce,
keypoint gather and the fused single-collective exchange."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from probpose_pytorch_b200 import distributed as ppd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, ragged):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = 5
        g = torch.Generator().manual_seed(0)
        records = torch.rand(B, K, 7, generator=g, dtype=torch.float64)      # the global batch, same on all ranks
        per_image_loss = torch.rand(B, generator=g)
        mine = ppd.shard_batch(records)
        lo, hi = ppd.shard_bounds(B, world, rank)
        assert mine.shape[0] == hi - lo and torch.equal(mine, records[lo:hi])
        gathered = ppd.all_gather_keypoints(mine.clone())
        assert torch.equal(gathered, records)
        mean = ppd.all_reduce_loss(per_image_loss[lo:hi].sum(), hi - lo)
        assert abs(float(mean) - float(per_image_loss.mean())) < 1e-6
        if not ragged:
            rec, loss = ppd.exchange_step_results(mine.clone(), per_image_loss[lo:hi].mean())
            assert torch.equal(rec, records)
            assert abs(float(loss) - float(per_image_loss.mean())) < 1e-6
            ex = ppd.exchange_step_results(mine.clone(), per_image_loss[lo:hi].mean(), async_op=True)   # overlapped form
            assert torch.equal(ex.wait().records, records)
            assert abs(float(ex.loss) - float(per_image_loss.mean())) < 1e-6
            # bucketed: three steps in one collective
            steps = [records * (i + 1) for i in range(3)]
            bx = ppd.exchange_bucket([ppd.shard_batch(t).clone() for t in steps],
                                     [per_image_loss[lo:hi].mean() * (i + 1) for i in range(3)], async_op=True)
            rec3, loss3 = bx
            assert rec3.shape == (3,) + tuple(records.shape) and torch.equal(rec3, torch.stack(steps))
            assert torch.allclose(loss3, per_image_loss.mean() * torch.tensor([1.0, 2.0, 3.0]), atol=1e-6)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,ragged", [(8, False), (7, True)])
def test_two_rank_exchange(B, ragged):
    mp.spawn(_worker, args=(2, _free_port(), B, ragged), nprocs=2, join=True)


def test_shard_bounds_cover_batch():
    for B in (0, 1, 7, 128, 1024):
        for ws in (1, 2, 4, 8):
            cuts = [ppd.shard_bounds(B, ws, r) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ppd.shard_bounds(8, 2, 2)


def test_single_process_is_identity():
    r = torch.rand(3, 2, 7)
    assert ppd.all_gather_keypoints(r) is r
    rec, loss = ppd.exchange_step_results(r, torch.tensor(0.5))
    assert rec is r and float(loss) == 0.5
    assert float(ppd.all_reduce_loss(torch.tensor(6.0), 3)) == 2.0
    assert np.array_equal(ppd.shard_batch(np.arange(5)), np.arange(5))
