"""Host-side work partitioning for data-parallel decode (no data-path collective).

The reference only ever parallelises decode by giving each rank its own decoder replica and a
disjoint set of utterances (`tts/data/data_utils.py:17-34` chunk_work,
`tts/inference/quality_validation.py:172-182`, `tts/training/rlhf/rlhf_main.py:141`). This module
is the B200 version of that: a cost-balanced partition of utterances over ranks, length-sorted
varlen buckets within a rank, and an optional final gather of the waveforms.
"""

from __future__ import annotations

import heapq
from typing import Sequence

import torch

# algorithmic FLOPs per token: 373.85 MFLOP of linear/conv work + 49 152 * T of attention
# (SURVEY.md 8a / BASELINE.md 4)
LINEAR_FLOPS_PER_TOKEN = 373_854_208
ATTN_FLOPS_PER_TOKEN_PER_T = 49_152


def utterance_cost(length: int) -> int:
    """Algorithmic FLOPs to decode one utterance of `length` tokens."""
    return length * (LINEAR_FLOPS_PER_TOKEN + ATTN_FLOPS_PER_TOKEN_PER_T * length)


def partition_utterances(lengths: Sequence[int], world_size: int) -> list[list[int]]:
    """Longest-processing-time-first greedy assignment of utterance indices to ranks. Every index
    appears exactly once; ranks are balanced by `utterance_cost`, not by count."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-utterance_cost(int(lengths[i])), i))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    shards: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + utterance_cost(int(lengths[i])), r))
    return shards


def bucket_by_length(indices: Sequence[int], lengths: Sequence[int], max_tokens: int = 16384) -> list[list[int]]:
    """Length-sorted varlen buckets of at most `max_tokens` tokens (at least one utterance each).
    Sorting keeps the attention tiles of a bucket similar in cost; packing is exact (no padding)."""
    if max_tokens < 1:
        raise ValueError("max_tokens must be >= 1")
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    buckets: list[list[int]] = []
    cur: list[int] = []
    cur_tokens = 0
    for i in order:
        n = int(lengths[i])
        if cur and cur_tokens + n > max_tokens:
            buckets.append(cur)
            cur, cur_tokens = [], 0
        cur.append(i)
        cur_tokens += n
    if cur:
        buckets.append(cur)
    return buckets


def gather_packed(flat: torch.Tensor, gathered: list[torch.Tensor] | None, dst: int = 0) -> None:
    """ONE `dist.gather` of equal-sized packed PCM buffers: `gathered` (a list of world_size tensors shaped
    like `flat`) is filled on `dst` and must be None elsewhere. Only `dst` receives data."""
    import torch.distributed as dist

    dist.gather(flat, gathered if dist.get_rank() == dst else None, dst=dst)


def gather_waveforms(local: dict[int, torch.Tensor], lengths: Sequence[int], hop: int, rank: int,
                     world_size: int, dst: int = 0, device: torch.device | str | None = None,
                     owner_lists: Sequence[Sequence[int]] | None = None):
    """The only collective of the path: the final gather of every rank's waveforms on `dst`
    (BASELINE north star: "no collective on the hot path beyond a final gather of waveforms").
    `local` maps utterance index -> (hop * T,) float32 tensor. Returns {index: tensor} on `dst`, None
    elsewhere.

    ONE `dist.gather` of each rank's packed PCM (ncclGather over NVLink on GPUs, gloo on CPU): only
    `dst` receives data -- config 3's ~7 GB of PCM lands once, not on every rank. The partition is a
    pure function of `lengths` (`partition_utterances`), so every rank can compute who owns what:
    pass `owner_lists` to skip the small `all_gather_object` of the index lists."""
    import torch.distributed as dist

    if world_size == 1:
        return dict(local)
    if owner_lists is None:
        gathered_owned: list[list[int]] = [None] * world_size  # type: ignore[list-item]
        dist.all_gather_object(gathered_owned, sorted(local.keys()))
        owner_lists = gathered_owned
    owner_lists = [sorted(int(i) for i in owned) for owned in owner_lists]
    if sorted(local.keys()) != owner_lists[rank]:
        raise ValueError("gather_waveforms: `local` does not hold exactly this rank's utterances")
    totals = [sum(int(lengths[i]) for i in owned) * hop for owned in owner_lists]
    width = max(max(totals), 1)  # gather needs equal-sized tensors: pad to the largest shard
    dev = torch.device(device) if device is not None else (next(iter(local.values())).device if local else torch.device("cpu"))
    flat = torch.zeros(width, dtype=torch.float32, device=dev)
    off = 0
    for i in owner_lists[rank]:
        n = int(lengths[i]) * hop
        flat[off:off + n] = local[i].reshape(-1).to(dev)
        off += n
    gathered = [torch.empty_like(flat) for _ in range(world_size)] if rank == dst else None
    gather_packed(flat, gathered, dst=dst)
    if rank != dst:
        return None
    out: dict[int, torch.Tensor] = {}
    for r, owned in enumerate(owner_lists):
        off = 0
        for i in owned:
            n = int(lengths[i]) * hop
            out[i] = gathered[r][off:off + n]
            off += n
    return out
