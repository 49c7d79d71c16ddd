"""World-size-2 CPU (gloo) test of the batch-sharding host logic: shard bounds, noise slicing, padded all_gather.
The decode itself is replaced by a deterministic stand-in (a pure function of each row's tokens, noise and global index),
so the check is exactly the property the GPU path relies on: sharded result == unsharded result, bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edm_tts_b200.runner import ShardedDecoder, shard_bounds


def fake_decode(sem, ap, sp, steps=1, temperature=1.0, seed=0, batch_offset=0, cat_gumbel=None, remask_gumbel=None, **kw):
    b, T = sem.shape
    q = torch.arange(12).view(1, 12, 1)
    rows = (torch.arange(b) + batch_offset).view(b, 1, 1)
    out = (sem[:, None, :] * (q + 1) + rows * 7 + seed) % 1024
    if cat_gumbel is not None:
        out = (out + cat_gumbel.view(cat_gumbel.shape[0], b, T, -1).argmax(-1).sum(0)[:, None, :]) % 1024
    if remask_gumbel is not None:
        out = (out + (remask_gumbel.sum(0) > 0).long()[:, None, :]) % 1024
    if ap is not None:
        out = (out + ap.sum(-1)[:, :12, None] + sp.sum(-1)[:, None, None]) % 1024
    return out.long()


def _inputs(B, T):
    g = torch.Generator().manual_seed(5)
    return dict(sem=torch.randint(0, 1024, (B, T), generator=g), ap=torch.randint(0, 1024, (B, 12, 9), generator=g),
                sp=torch.randint(0, 1024, (B, 9), generator=g), cat=torch.randn(2, B * T, 16, generator=g), rem=torch.randn(2, B, T, generator=g))


def _worker(rank, world, port, B, T, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    i = _inputs(B, T)
    dec = ShardedDecoder(fake_decode)
    out = dec(i["sem"], i["ap"], i["sp"], steps=3, seed=4, cat_gumbel=i["cat"], remask_gumbel=i["rem"])
    ret[rank] = out
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("B", [5, 8, 1])
def test_sharded_equals_unsharded_world2(B):
    T, world = 11, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, T, ret), nprocs=world, join=True)
    i = _inputs(B, T)
    full = fake_decode(i["sem"], i["ap"], i["sp"], steps=3, seed=4, batch_offset=0, cat_gumbel=i["cat"], remask_gumbel=i["rem"])
    for r in range(world):
        assert torch.equal(ret[r], full), f"rank {r}"


def test_shard_bounds_cover_batch():
    for B in (1, 7, 64, 512):
        for R in (1, 2, 4, 8):
            spans = [shard_bounds(B, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[k][1] == spans[k + 1][0] for k in range(R - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_single_process_passthrough():
    i = _inputs(4, 6)
    out = ShardedDecoder(fake_decode)(i["sem"], None, None, steps=1)
    assert torch.equal(out, fake_decode(i["sem"], None, None))
