"""Host-side multi-GPU plumbing: independent audio streams shard across ranks, transcripts gather on rank 0.

The streaming path has no exchange step: stream s lives on rank s mod G for its whole life
(its K/V ring, conv state and decoder state stay on that GPU), weights replicate, and the only cross-rank traffic
is a host-side gather of token ids / transcripts. torch.distributed is used for exactly that (NCCL on the GPU box,
gloo in the CPU tests); there is no collective on the data path.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Sequence


def owner_of(stream_id: int, world_size: int) -> int:
    """Sticky placement: stream s -> rank s mod G."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    return stream_id % world_size


def local_streams(stream_ids: Iterable[int], rank: int, world_size: int) -> List[int]:
    return [s for s in stream_ids if owner_of(s, world_size) == rank]


def gather_results(local: Dict[int, object], rank: int, world_size: int, dst: int = 0):
    """Gather {stream_id: result} from every rank onto `dst` (host objects; a few bytes per chunk per stream).
    Returns the merged dict on dst, None elsewhere. Raises if two ranks claim the same stream."""
    if world_size == 1:
        return dict(local)
    import torch.distributed as dist
    parts = [None] * world_size if rank == dst else None
    dist.gather_object(local, parts, dst=dst)
    if rank != dst:
        return None
    merged: Dict[int, object] = {}
    for r, p in enumerate(parts):
        for k, v in p.items():
            if k in merged:
                raise RuntimeError(f"stream {k} reported by more than one rank")
            if owner_of(k, world_size) != r:
                raise RuntimeError(f"stream {k} reported by rank {r}, owner is {owner_of(k, world_size)}")
            merged[k] = v
    return merged


def run_sharded(stream_ids: Sequence[int], rank: int, world_size: int, process: Callable[[List[int]], Dict[int, object]]):
    """Run `process` on this rank's share of the streams and gather everything on rank 0."""
    mine = local_streams(stream_ids, rank, world_size)
    out = process(mine)
    missing = set(mine) - set(out)
    if missing:
        raise RuntimeError(f"rank {rank} produced no result for streams {sorted(missing)}")
    return gather_results(out, rank, world_size)
