"""Multi-GPU plumbing of the vocoder / synthesizer path.

The path shards by independent units (utterances; the folds of ONE long WaveRNN utterance, which are independent
rows of the sample loop), so the only collectives are the start-up broadcast of the packed weight arenas over
NCCL/NVLink and the gather of the per-rank results (int16 fold rows, ~1 MB for a 30 s utterance).  One process per GPU (torchrun), ``torch.distributed`` is
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


def fold_range(num_folds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous folds [lo, hi) of one utterance owned by `rank` (58 folds on 8 ranks -> 8,8,7,7,7,7,7,7)."""
    base, rem = divmod(int(num_folds), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_fold_rows(local_idx, num_folds: int, dst: int = 0, device=None):
    """Gather the per-rank int16 [rows_r, steps] class indices of one utterance into [num_folds, steps] on `dst`
    (None elsewhere).  NCCL: device tensors; gloo: host tensors."""
    import numpy as np

    rank, ws = world()
    if ws == 1:
        return local_idx
    steps = int(local_idx.shape[1])
    rows_max = (num_folds + ws - 1) // ws
    use_cuda = dist.get_backend() == "nccl"
    dev = (device or torch.device("cuda", torch.cuda.current_device())) if use_cuda else torch.device("cpu")
    buf = torch.zeros(rows_max, steps, dtype=torch.int16, device=dev)
    buf[: local_idx.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_idx)).to(dev)
    wire = buf.view(torch.uint8)  # neither NCCL nor gloo transports int16: ship the bytes
    outs8 = [torch.empty_like(wire) for _ in range(ws)] if rank == dst else None
    dist.gather(wire, outs8, dst=dst)
    outs = [o.view(torch.int16) for o in outs8] if rank == dst else None
    if rank != dst:
        return None
    full = np.empty((num_folds, steps), np.int16)
    for r in range(ws):
        lo, hi = fold_range(num_folds, r, ws)
        full[lo:hi] = outs[r][: hi - lo].cpu().numpy()
    return full
