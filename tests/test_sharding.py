"""CPU: the data-parallel host logic, including a world_size-2 gloo run of the final gather."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tts_max_b200 import sharding


def test_partition_covers_everything_and_balances():
    g = torch.Generator().manual_seed(2024)
    lengths = torch.randint(100, 1001, (1000,), generator=g).tolist()
    for ws in (1, 2, 4, 8):
        shards = sharding.partition_utterances(lengths, ws)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(1000))
        loads = [sum(sharding.utterance_cost(lengths[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / ws) < 1.01  # LPT: within 1 % of perfect balance


def test_partition_edge_cases():
    assert sharding.partition_utterances([], 4) == [[], [], [], []]
    assert sorted(map(len, sharding.partition_utterances([5], 4))) == [0, 0, 0, 1]
    with pytest.raises(ValueError):
        sharding.partition_utterances([1], 0)


def test_buckets_respect_token_budget():
    g = torch.Generator().manual_seed(1)
    lengths = torch.randint(100, 1001, (300,), generator=g).tolist()
    buckets = sharding.bucket_by_length(range(300), lengths, max_tokens=8192)
    assert sorted(i for b in buckets for i in b) == list(range(300))
    for b in buckets:
        assert sum(lengths[i] for i in b) <= 8192 or len(b) == 1
        assert [lengths[i] for i in b] == sorted((lengths[i] for i in b), reverse=True)
    assert sharding.bucket_by_length([0], [50000], max_tokens=8192) == [[0]]  # oversize utterance alone


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_decode(ids: torch.Tensor, hop: int) -> torch.Tensor:
    # stand-in for the GPU decode: a deterministic function of the utterance only
    return (ids.float().repeat_interleave(hop) * 1e-3).contiguous()


def _worker(rank, world_size, port, lengths, hop, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    g = torch.Generator().manual_seed(0)
    utts = [torch.randint(0, 65536, (n,), generator=g) for n in lengths]
    mine = sharding.partition_utterances(lengths, world_size)[rank]
    local = {}
    for bucket in sharding.bucket_by_length(mine, lengths, max_tokens=64):
        for i in bucket:
            local[i] = _fake_decode(utts[i], hop)
    out = sharding.gather_waveforms(local, lengths, hop, rank, world_size, dst=0)
    if rank == 0:
        ok = sorted(out.keys()) == list(range(len(lengths))) and all(
            torch.equal(out[i], _fake_decode(utts[i], hop)) for i in range(len(lengths)))
        q.put(ok)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    lengths = [7, 3, 12, 5, 9, 1, 30]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
