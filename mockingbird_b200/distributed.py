"""Multi-GPU plumbing of the vocoder / synthesizer path.

The path shards by independent units (utterances; WaveRNN folds stay inside one kernel), so the only
collective is the start-up broadcast of the packed weight arenas over NCCL/NVLink; results stay on
their rank or are gathered by the host.  One process per GPU (torchrun), ``torch.distributed`` is
plumbing only.  The same functions run on the ``gloo`` backend for the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_utterances(lengths: Sequence[int], rank: int, world_size: int) -> List[int]:
    """Indices of the utterances rank `rank` processes: sort by length (longest first) and deal
    round-robin, so every rank gets the same number of items (+-1) with similar total padded work."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return order[rank::world_size]


def broadcast_weights(arenas: Sequence[torch.Tensor], src: int = 0) -> None:
    """Broadcast the packed weight arenas (uint8 tensors from ``model.packed_arena()``) from `src`."""
    rank, ws = world()
    if ws == 1:
        return
    for a in arenas:
        dist.broadcast(a, src=src)


def gather_object_lists(local: list, dst: int = 0):
    """Collect per-rank python result lists on `dst` (host gather of waveforms / spectrograms)."""
    rank, ws = world()
    if ws == 1:
        return [local]
    out = [None] * ws if rank == dst else None
    dist.gather_object(local, out, dst=dst)
    return out


def merge_sharded(results_per_rank: Sequence[Sequence], shards_per_rank: Sequence[Sequence[int]], n: int) -> list:
    """Undo shard_utterances: place rank r's k-th result at global index shards_per_rank[r][k]."""
    out = [None] * n
    for res, idx in zip(results_per_rank, shards_per_rank):
        for v, i in zip(res, idx):
            out[i] = v
    return out
