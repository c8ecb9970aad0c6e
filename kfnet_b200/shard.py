"""Sequence sharding across ranks: sequences are independent units, weights are replicated, and the
only collective on the path is one all_gather of the per-rank output maps."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_sequences(num_sequences: int, world: int, rank: int) -> list[int]:
    """Rank r owns sequences r, r+world, ...  (equal counts required so a single all_gather suffices)."""
    if num_sequences % world != 0:
        raise ValueError("num_sequences must be a multiple of the world size")
    return list(range(rank, num_sequences, world))


def gather_outputs(local_out: torch.Tensor, num_sequences: int) -> torch.Tensor:
    """local_out [S_r, T, GH, GW, 4] of this rank's shard -> [num_sequences, T, GH, GW, 4] in global sequence
    order on every rank, with ONE collective (NCCL all_gather over NVLink on GPUs, gloo in the CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_out
    world = dist.get_world_size()
    buf = torch.empty((world,) + tuple(local_out.shape), dtype=local_out.dtype, device=local_out.device)
    if local_out.is_cuda:
        dist.all_gather_into_tensor(buf, local_out.contiguous())
    else:
        dist.all_gather(list(buf.unbind(0)), local_out.contiguous())
    # buf[r, k] is sequence r + k*world  ->  interleave back to global order
    return buf.transpose(0, 1).reshape((num_sequences,) + tuple(local_out.shape[1:]))
