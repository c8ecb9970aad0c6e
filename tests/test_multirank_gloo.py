"""N>1 host logic on CPU: world_size-2 gloo.  Sharding assigns whole sequences to ranks; the gathered result
must be bit-identical to the single-process result in global sequence order (no cross-rank arithmetic exists)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_eval(seq_id, T=3, GH=4, GW=5):
    g = torch.Generator().manual_seed(100 + seq_id)
    return torch.randn(T, GH, GW, 4, generator=g)


def _worker(rank, world, port, nseq, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kfnet_b200.shard import gather_outputs, shard_sequences
    mine = shard_sequences(nseq, world, rank)
    local = torch.stack([_fake_eval(s) for s in mine])
    full = gather_outputs(local, nseq)
    ret[rank] = (mine, full.clone())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nseq", [2, 8])
def test_sequence_sharding_and_single_gather(nseq):
    world = 2
    mgr = mp.get_context("spawn").Manager()      # (fork from a multi-threaded pytest process is deprecated)
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), nseq, ret), nprocs=world, join=True)
    ref = torch.stack([_fake_eval(s) for s in range(nseq)])
    owned = sorted(s for r in range(world) for s in ret[r][0])
    assert owned == list(range(nseq))
    for r in range(world):
        assert torch.equal(ret[r][1], ref)


def test_shard_requires_equal_counts():
    from kfnet_b200.shard import shard_sequences
    assert shard_sequences(8, 4, 1) == [1, 5]
    with pytest.raises(ValueError):
        shard_sequences(3, 2, 0)
