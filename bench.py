#!/usr/bin/env python
"""Benchmark of the B200-native codec decode path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One "step" = one decode pass over one synthetic batch (BASELINE config 2 by default:
16 utterances x 500 tokens = 160 audio-seconds, bf16 tensor-core operands), per GPU
(weak scaling: every rank decodes its own batch; there is no data-path collective).
Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "codec decode audio-sec/sec (device-timed)"
UNIT = "audio-s/s"
TOKEN_RATE = 50
HOP = 320

WORKLOADS = {
    # name: (description, utterances, tokens per utterance)
    "b1": ("1 clip x 5 s (the shape of AudioDecoder.decode, decoding.py:84-89; not a BASELINE config)", 1, 250),
    "c1": ("4 clips x 5 s (BASELINE config 1, the reference's CPU case)", 4, 250),
    "c2": ("16 clips x 10 s (BASELINE config 2)", 16, 500),
    "c4": ("4 clips x 60 s long-form (BASELINE config 4)", 4, 3000),
    "c5": ("64 windows x (100 context + 50 new) tokens (BASELINE config 5); only the 50 new tokens of each "
           "window count as produced audio", 64, 150),
}
# tokens of each utterance that count as produced audio (streaming windows recompute their left context)
NEW_TOKENS = {"c5": 50}
# BASELINE config 3: 10 000 utterances of 2-20 s (100..1000 tokens, seed 2024), length-bucketed varlen
# packs of <= 16 384 tokens, sharded over the ranks by cost (strong scaling: total work is fixed).
C3_UTTS, C3_SEED, C3_BUCKET_TOKENS = 10_000, 2024, 16_384

LINEAR_FLOPS_PER_TOKEN = 373_854_208  # SURVEY.md 8a / BASELINE.md 4
# project_out, fc_post_a and the embed conv are folded into one lookup at load time: their
# 4 194 304 + 14 680 064 FLOP/token of the reference's algorithmic work are not executed and not counted
FOLDED_FLOPS_PER_TOKEN = 4_194_304 + 14_680_064
FRONTEND_GEMM_FLOPS_PER_TOKEN = 2 * 128 * 1024  # what is executed instead: codes im2col x [hi | lo] coefficients
GEMM_FLOPS_PER_TOKEN = FRONTEND_GEMM_FLOPS_PER_TOKEN + 50_331_648 + 301_989_888 + 2_625_536  # tcgen05 GEMM/conv kernel
ATTN_FLOPS_PER_TOKEN_PER_T = 49_152


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self._first = threading.Event()
        self._lo, self._hi = 0, None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                     "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
        except Exception:
            self._first.set()
            return
        self._proc = proc
        for line in proc.stdout:
            if self._stop.is_set():
                break
            self.rows.append([c.strip() for c in line.split(",")])
            self._first.set()
        self._first.set()

    def mark_start(self, timeout: float = 10.0):
        """Blocks until nvidia-smi has delivered its first row (its start-up can take longer than a short
        timed region), then marks the beginning of the timed region."""
        self._first.wait(timeout)
        self._lo = len(self.rows)

    def mark_end(self):
        time.sleep(0.03)  # one more sampling period, so that a short region still holds its last sample
        self._hi = len(self.rows)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        proc = getattr(self, "_proc", None)
        if proc is not None:
            proc.kill()
        self._thread.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[max(self._lo - 1, 0):self._hi]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_ids(n_utts: int, tokens: int, seed: int):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 65536, (n_utts * tokens,), generator=g, dtype=torch.int64)


def measure_encode(dev, precision: str):
    """Encode direction on this GPU: 16 clips x 10 s through Encoder.forward (acoustic + semantic
    encoders, fusion, quantise; the w2v-BERT hidden state is a synthetic input), device-timed. Random weights with
    the shapes of the reference state dict (weight-normed convs with g = 0.4 |v|, alpha = beta = 0, 12-tap
    windowed-sinc filters): throughput does not depend on the values."""
    import math
    import torch
    from tts_max_b200.codec import encoder
    g = torch.Generator().manual_seed(11)
    n = torch.arange(12, dtype=torch.float32) - 5.5
    lp = torch.sinc(n / 2) * torch.hann_window(12, periodic=False)
    lp = (lp / lp.sum()).view(1, 1, 12)
    sd = {}
    shapes = encoder.expected_state_dict_shapes()
    for key, shape in shapes.items():
        if key.endswith("filter"):
            sd[key] = lp.clone()
        elif key.endswith(("alpha", "beta")):
            sd[key] = torch.zeros(shape)
        elif key.endswith("weight_g"):
            continue
        else:
            fan_in = math.prod(shape[1:]) if len(shape) > 1 else shape[0]
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
    for key, shape in shapes.items():
        if key.endswith("weight_g"):
            v = sd[key[:-1] + "v"]
            sd[key] = 0.4 * v.flatten(1).norm(dim=1).view(shape)
    enc = encoder.Encoder(pre_bound=False, precision=precision)
    enc.load_state_dict(sd)
    enc.to(dev).eval()
    clips, S = 16, 160000
    wav = (0.3 * torch.randn(clips, 1, S, generator=g)).to(dev)
    w2v = torch.randn(clips, S // 320, 1024, generator=g).to(dev)
    for _ in range(2):
        ids = enc(wav, w2v)
    torch.cuda.synchronize(dev)
    n0 = enc.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        ids = enc(wav, w2v)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    assert ids.shape == (clips, 1, S // 320) and int(ids.min()) >= 0 and int(ids.max()) < 65536
    launches = (enc.launch_count() - n0) // steps
    del enc
    torch.cuda.empty_cache()
    return {"value": round(clips * S / 16000 / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms, 3),
            "gpu_launches_per_step": int(launches),
            "workload": "16 clips x 10 s of 16 kHz audio -> 16 x 500 ids in one launch sequence (Encoder.forward: acoustic + "
                        "semantic encoders, fusion, quantise; synthetic w2v-BERT hidden state, random weights)",
            "timing": "CUDA events around 5 back-to-back passes, inputs resident"}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (CPU restatement of the reference algorithm)
# ------------------------------------------------------------------------------------------------
CPU_FULL_TOKENS = 8000  # workloads up to this many tokens per step run in full on the CPU arm


def cpu_clips(n_utts: int, tokens: int) -> int:
    """Clips of the workload the CPU arm decodes per step: all of them when one pass takes seconds
    (c1, c2), else a bounded sample (c4: one 60 s clip; c5: 8 windows)."""
    if n_utts * tokens <= CPU_FULL_TOKENS:
        return n_utts
    return max(1, min(n_utts, CPU_FULL_TOKENS // tokens, 8))


def cpu_decode_rate(tokens: int, clips: int, steps: int, warmup: int, seed: int = 1234, new_tokens: int | None = None):
    """audio-s/s of the reference algorithm (oracle port, fp32, all host threads) on `clips` x `tokens`."""
    import torch

    from oracle import codec_oracle as O
    from oracle import weights

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = weights.make_state_dict(seed=0, perturb=False)
    ids = synthetic_ids(clips, tokens, seed).view(clips, tokens)
    for _ in range(warmup):
        O.decoder_forward(sd, ids)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.decoder_forward(sd, ids)
        times.append(time.perf_counter() - t0)
    audio_s = clips * (new_tokens or tokens) / TOKEN_RATE
    mean = sum(times) / len(times)
    return audio_s / mean, mean * 1e3, cores


def cpu_sample_text(clips: int, n_utts: int, tokens: int) -> str:
    what = "all" if clips == n_utts else f"{clips} of the"
    return (f"{what} {n_utts} x {tokens}-token utterances per step (same ids seed as the GPU arm's rank 0), "
            "oracle port of the reference algorithm, fp32, torch CPU")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    wl = "c2" if args.workload == "c3" else args.workload
    desc, n_utts, tokens = WORKLOADS[wl]
    clips = cpu_clips(n_utts, tokens)
    value, ms, cores = cpu_decode_rate(tokens, clips, args.steps, max(args.warmup, 1), new_tokens=NEW_TOKENS.get(wl))
    sample = cpu_sample_text(clips, n_utts, tokens)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{wl}: {desc}", "utterances_per_gpu": n_utts, "tokens_per_utterance": tokens,
                   "audio_seconds_per_step_per_gpu": clips * (NEW_TOKENS.get(wl) or tokens) / TOKEN_RATE,
                   "sample": sample, "weights": "random-init, reference distributions (seed 0)",
                   "timing": "time.perf_counter around each pass, mean over the timed steps"},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def measure_c3(dec, dev, rank, world, n_utts, precision, steps=1, warmup=1, e2e=True):
    """BASELINE config 3 (strong scaling): `n_utts` utterances of 2-20 s, LPT-sharded over the ranks,
    decoded in length-sorted varlen packs; then the final gather of the PCM on rank 0. Returns the block
    (rank 0) or None. Device times are CUDA events around each rank's whole shard, max over ranks."""
    import torch
    import torch.distributed as dist

    from tts_max_b200 import sharding

    g = torch.Generator().manual_seed(C3_SEED)
    lengths = torch.randint(100, 1001, (n_utts,), generator=g).tolist()
    shards = sharding.partition_utterances(lengths, world)
    mine = shards[rank]
    buckets = sharding.bucket_by_length(mine, lengths, max_tokens=C3_BUCKET_TOKENS)
    gi = torch.Generator().manual_seed(1234 + rank)
    packs = []
    for b in buckets:
        seqlens = [lengths[i] for i in b]
        packs.append((torch.randint(0, 65536, (sum(seqlens),), generator=gi, dtype=torch.int64).pin_memory(), seqlens))
    packs_dev = [(ids.to(dev), seqlens) for ids, seqlens in packs]
    hop = dec.samples_per_token
    shard_samples = [sum(lengths[i] for i in s) * hop for s in shards]
    width = max(shard_samples)                       # gather needs equal-sized buffers
    pcm = torch.empty(width, dtype=torch.float32, device=dev)   # this rank's packed PCM, bucket order

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def decode_shard(from_host):
        off = 0
        for (ids_h, seqlens), (ids_d, _) in zip(packs, packs_dev):
            n = sum(seqlens) * hop
            ids = ids_h.to(dev, non_blocking=True) if from_host else ids_d
            dec.decode_packed_device(ids, seqlens, out=pcm[off:off + n])
            off += n

    for _ in range(max(1, warmup)):
        decode_shard(False)
    barrier()
    launches0 = dec.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clocks:
        clocks.mark_start()
        e0.record()
        for _ in range(steps):
            decode_shard(False)
        e1.record()
        barrier()
        clocks.mark_end()
    my_ms = e0.elapsed_time(e1) / steps
    launches = (dec.launch_count() - launches0) // steps

    # end to end: pinned host ids -> H2D -> decode -> final gather of the PCM on rank 0 (its HBM)
    e2e_s = gather_ms = None
    if e2e:
        gathered = [torch.empty_like(pcm) for _ in range(world)] if (rank == 0 and world > 1) else None
        if world > 1:   # untimed: NCCL sets its peer connections up on the first use of a collective
            sharding.gather_packed(pcm, gathered, dst=0)
        barrier()
        t0 = time.perf_counter()
        decode_shard(True)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        if world > 1:
            sharding.gather_packed(pcm, gathered, dst=0)
        g1.record()
        barrier()
        e2e_s = time.perf_counter() - t0
        gather_ms = g0.elapsed_time(g1)
        if rank == 0 and world > 1:   # the gathered shards are what every rank decoded (spot check)
            assert all(torch.isfinite(t[:4096]).all() for t in gathered)
    assert torch.isfinite(pcm[:4096]).all()

    all_ms = [my_ms]
    if world > 1:
        t = torch.tensor([my_ms, e2e_s or 0.0, gather_ms or 0.0], dtype=torch.float64, device=dev)
        rows = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(rows, t)
        all_ms = [float(r[0]) for r in rows]
        e2e_s = max(float(r[1]) for r in rows) if e2e else None
        gather_ms = max(float(r[2]) for r in rows) if e2e else None
    if rank != 0:
        return None
    total_audio_s = sum(lengths) / TOKEN_RATE
    ms = max(all_ms)
    flops = sum(sharding.utterance_cost(t) for t in lengths)
    peaks = read_peaks()
    return {
        "workload": f"c3: {n_utts} utterances of 2-20 s (BASELINE config 3), length-bucketed varlen packs of <= "
                    f"{C3_BUCKET_TOKENS} tokens, LPT-sharded over {world} rank(s) by T*(373.85e6 + 49152*T) FLOPs",
        "scaling": "strong", "value": round(total_audio_s / (ms / 1e3), 2), "unit": UNIT, "n_gpus": world,
        "audio_seconds_total": total_audio_s, "ms_per_pass": round(ms, 3), "steps": steps,
        "rank_ms": {"min": round(min(all_ms), 3), "mean": round(sum(all_ms) / len(all_ms), 3), "max": round(ms, 3),
                    "imbalance": round(ms / (sum(all_ms) / len(all_ms)), 4)},
        "packs_on_rank0": len(packs), "gpu_launches_per_pass_rank0": int(launches),
        "e2e": None if not e2e else {
            "value": round(total_audio_s / e2e_s, 2), "unit": UNIT, "wall_s": round(e2e_s, 4),
            "gather_ms": round(gather_ms, 3), "gather_bytes_total": int(sum(shard_samples) * 4),
            "what": "pinned host ids -> H2D -> decode -> ONE dist.gather (NCCL) of each rank's packed PCM into rank 0's "
                    "HBM; wall clock between barriers, max over ranks",
            "h2d_bytes": int(sum(lengths) * 8), "d2h_bytes": 0},
        "clocks": clocks.summary(),
        "whole_step_tflops_per_gpu": round(flops / (ms / 1e3) / 1e12 / world, 1),
        "frac_of_sustained_peak": round(flops / (ms / 1e3) / 1e12 / world / peaks["tflops_sustained"], 4),
        "limiter": "per-rank tail: the LPT partition balances FLOPs, not packs -- the last (shortest-utterance) pack of a "
                   "rank is a partial one; no collective on the data path, the gather moves "
                   f"{sum(shard_samples) * 4 / 1e9:.1f} GB once",
    }


def setup_dist():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tts_max_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return rank, local_rank, world, dev


def run_c3(args):
    """`--workload c3`: the strong-scaling run of BASELINE config 3 as the headline line."""
    import torch.distributed as dist

    from tts_max_b200.codec import decoder

    rank, local_rank, world, dev = setup_dist()
    dec = decoder.Decoder(16000, HOP, None, None, precision=args.precision, init_seed=0).to(dev).eval()
    steps = max(1, min(args.steps, 5))
    blk = measure_c3(dec, dev, rank, world, args.c3_utts, args.precision, steps=steps, warmup=max(1, min(args.warmup, 2)))
    if rank == 0:
        line = {
            "metric": METRIC, "value": blk["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": blk["ms_per_pass"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": blk["workload"], "audio_seconds_total": blk["audio_seconds_total"],
                       "parallelism": f"dp{world} (independent utterances, no data-path collective)",
                       "l2": "each pack's working set (weights 374 MB + ~0.6 GB activations) exceeds the 126 MB L2",
                       "timing": "CUDA events around the rank's whole shard, max over ranks"},
            "e2e": blk["e2e"], "gpu_launches": blk["gpu_launches_per_pass_rank0"] * steps, "clocks": blk["clocks"],
            "roofline": {"kernel": "whole step (all kernels)", "bound": "tensor", "achieved": blk["whole_step_tflops_per_gpu"],
                         "peak": read_peaks()["tflops_sustained"], "unit": "TFLOP/s per GPU",
                         "frac": blk["frac_of_sustained_peak"], "traffic": None,
                         "peak_source": "measured bf16_tflops_sustained (seconds-long pass)"},
            "cpu_baseline": None, "c3_strong": blk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def latest_traffic():
    """DRAM bytes per launch of the dominant kernel from the most recent committed ncu --set full capture
    (profiles/rNN*_gemm_traffic.json, written by tools/ncu_traffic.py with the build's git hash)."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        d = json.load(f)
    return d.get("mean_dram_bytes_per_launch"), {"file": os.path.relpath(files[-1], ROOT), "git": d.get("git"), "note": d.get("note")}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from tts_max_b200.codec import decoder

    if args.no_pdl:
        from tts_max_b200 import _lib
        _lib.check(_lib.load().b200codec_set_pdl(0))
    if args.no_early_weights:
        from tts_max_b200 import _lib
        _lib.check(_lib.load().b200codec_set_gemm_early_weights(0))
    if args.chain:
        from tts_max_b200 import _lib
        _lib.check(_lib.load().b200codec_set_gemm_chain(1))
    if args.istft_tile:
        from tts_max_b200 import _lib
        _lib.check(_lib.load().b200codec_set_istft_tile(args.istft_tile))
    if args.workload == "c3":
        return run_c3(args)

    rank, local_rank, world, dev = setup_dist()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    desc, n_utts, tokens = WORKLOADS[args.workload]
    seqlens = [tokens] * n_utts
    total_tokens = n_utts * tokens
    audio_s = n_utts * NEW_TOKENS.get(args.workload, tokens) / TOKEN_RATE

    if args.model == "48k":
        dec = decoder.Decoder(48000, 160, [3, 2], [7, 6], precision=args.precision, init_seed=0)
    else:
        dec = decoder.Decoder(16000, HOP, None, None, precision=args.precision, init_seed=0)
    dec.to(dev).eval()
    ids_host = synthetic_ids(n_utts, tokens, 1234 + rank).pin_memory()
    ids_dev = ids_host.to(dev)
    wav_host = torch.empty(total_tokens * dec.samples_per_token, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") ----
    for _ in range(max(args.warmup, 3)):
        dec.decode_packed_device(ids_dev, seqlens)
    barrier()
    launches0 = dec.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        clocks.mark_start()  # waits for the sampler's first row
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()  # L2 flush between timed steps (outside the per-step event bracket)
            starts[i].record()
            wav = dec.decode_packed_device(ids_dev, seqlens)
            ends[i].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks.mark_end()
    launches = dec.launch_count() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    dev_ms_total = sum(step_ms)
    assert torch.isfinite(wav[:4096]).all()

    # ---- end to end through the public host-buffer API ("e2e") ----
    for _ in range(2):
        dec.decode_packed_host(ids_host, seqlens, out=wav_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.decode_packed_host(ids_host, seqlens, out=wav_host)  # H2D ids, decode, D2H PCM, sync
    barrier()
    e2e_s_total = time.perf_counter() - t0

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms_total, e2e_s_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_total, e2e_s_total = t.tolist()

    ms_per_step = dev_ms_total / args.steps
    value = world * audio_s / (ms_per_step / 1e3)
    e2e_value = world * audio_s * args.steps / e2e_s_total

    # ---- roofline: per-stage device times from a separate pass (never inside the timed region) ----
    roofline = None
    stage_ms = {}
    stage_ms_in_step = {}
    cpu_baseline = None
    sustained = None
    latency = None
    peaks = read_peaks()
    exec_flops_per_token = (LINEAR_FLOPS_PER_TOKEN - FOLDED_FLOPS_PER_TOKEN + FRONTEND_GEMM_FLOPS_PER_TOKEN
                            + ATTN_FLOPS_PER_TOKEN_PER_T * tokens)
    algo_flops_per_token = LINEAR_FLOPS_PER_TOKEN + ATTN_FLOPS_PER_TOKEN_PER_T * tokens
    step_flops = exec_flops_per_token * total_tokens
    if rank == 0 and args.model == "xcodec2":
        dec.profile(True)
        prof_steps = 3
        for _ in range(prof_steps):
            dec.decode_packed_device(ids_dev, seqlens)
        torch.cuda.synchronize(dev)
        stage_ms = {k: v / prof_steps for k, v in dec.stage_times().items()}
        dec.profile(False)
        # The event-fenced pass breaks the programmatic-dependent-launch overlap between kernels, so its stages
        # sum to more than the real step. In-step time per stage = its SHARE of that pass x the measured step
        # (the overlap gain is spread proportionally; the ncu launch list under profiles/ shows the same shares).
        fenced_total = max(sum(stage_ms.values()), 1e-9)
        stage_ms_in_step = {k: v / fenced_total * ms_per_step for k, v in stage_ms.items()}
        gemm_share = sum(v for k, v in stage_ms.items() if k.endswith("_gemm")) / fenced_total
        gemm_ms = gemm_share * ms_per_step
        # GEMM launches = all launches minus 12 attention, 8 GroupNorm apply, code im2col, LayerNorm, ISTFT
        # (24 with GEMM chains: front end, 8 conv3, first c_attn, 12 chains, last fc2, head; 58 without)
        n_gemm = launches // args.steps - 23
        gemm_flops = GEMM_FLOPS_PER_TOKEN * total_tokens
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        # a burst region (tens of ms at full clocks) is compared with the burst peak; the sustained block
        # below (seconds, power-capped clocks) with the sustained peak
        burst_s = dev_ms_total / 1e3
        peak = peaks["tflops_burst"]
        traffic, traffic_src = latest_traffic() if args.workload == "c2" else (None, None)
        roofline = {
            "kernel": "gemm_tc05_2cta_kernel (tcgen05 cta_group::2 GEMM / implicit conv1d; all dense layers;)",
            "bound": "tensor", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"{peaks['source']} bf16_tflops (BURST: the timed region is {burst_s * 1e3:.0f} ms of kernels at "
                           "full clocks; see roofline.sustained for the seconds-long figure against bf16_tflops_sustained)",
            "launches_per_step": n_gemm, "avg_launch_ms": round(gemm_ms / n_gemm, 4),
            "flops_per_step": gemm_flops, "flops_basis": "executed (= algorithmic for these layers; fc_post_a + embed are folded "
                                                        "into a K = 128 GEMM and counted at what is executed)",
            "share_of_step": round(gemm_share, 4),
            "time_basis": "share of the event-fenced stage pass x the PDL-on step time (stages sum to the step)",
            "whole_step": {"flops_executed": step_flops, "achieved": round(step_flops / (ms_per_step / 1e3) / 1e12, 1),
                           "frac": round(step_flops / (ms_per_step / 1e3) / 1e12 / peak, 4),
                           "achieved_algorithmic": round(algo_flops_per_token * total_tokens / (ms_per_step / 1e3) / 1e12, 1),
                           "frac_algorithmic": round(algo_flops_per_token * total_tokens / (ms_per_step / 1e3) / 1e12 / peak, 4)},
        }
        # HBM-bound kernels: algorithmic bytes per token (operand dtype as stored) / in-step stage time
        rows = total_tokens + 3 * (n_utts - 1)
        hbm = peaks["hbm_gbs"]

        def hbm_entry(name, stage, bytes_per_row, n_launch):
            ms = stage_ms_in_step.get(stage, 0.0)
            gbs = bytes_per_row * rows * n_launch / (ms / 1e3) / 1e9 if ms > 0 else 0.0
            return {"kernel": name, "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm, "unit": "GB/s",
                    "frac": round(gbs / hbm, 4), "ms_per_step": round(ms, 4), "launches_per_step": n_launch}

        roofline["hbm_kernels"] = [
            hbm_entry("fsq_im2col_kernel (ids -> codes of 7 neighbouring frames: 8 B id in, 128 x 2 B out)", "fsq_lookup", 8 + 256, 1),
            hbm_entry("groupnorm_apply_swish x8 (per apply: 4096 B in, 2048 B out; statistics come from GEMM epilogues)",
                      "groupnorm_swish", 4096 + 2048, 8),
            hbm_entry("rownorm_kernel / LayerNorm (4096 B in, 2048 B out)", "layernorm", 4096 + 2048, 1),
            hbm_entry("istft_kernel (1282 x 4 B in, 320 x 4 B out)", "istft", 5128 + 1280, 1),
        ]
        attn_ms = stage_ms_in_step.get("attention", 0.0)
        if attn_ms > 0:
            attn_flops = ATTN_FLOPS_PER_TOKEN_PER_T * tokens * total_tokens
            roofline["attention"] = {
                "kernel": "attention_tc05_kernel (tcgen05, S/P/O in TMEM)", "bound": "MUFU (16 384 exp per 128x128 tile: half the tensor roofline for d = 64)",
                "achieved": round(attn_flops / (attn_ms / 1e3) / 1e12, 1), "peak": peak / 2, "unit": "TFLOP/s",
                "frac": round(attn_flops / (attn_ms / 1e3) / 1e12 / (peak / 2), 4), "ms_per_step": round(attn_ms, 4)}

    # ---- sustained block: the same step back to back for >= args.sustain_s seconds of GPU time ----
    if rank == 0 and world == 1 and args.sustain_s > 0 and args.model == "xcodec2":
        n_loop = max(10, int(args.sustain_s * 1e3 / ms_per_step) + 1)
        torch.cuda.synchronize(dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as sclk:
            sclk.mark_start()
            s0.record()
            for _ in range(n_loop):
                dec.decode_packed_device(ids_dev, seqlens)
            s1.record()
            torch.cuda.synchronize(dev)
            sclk.mark_end()
        sus_ms = s0.elapsed_time(s1) / n_loop
        sc = sclk.summary()
        sustained = {
            "steps": n_loop, "seconds": round(sus_ms * n_loop / 1e3, 3), "ms_per_step": round(sus_ms, 4),
            "value": round(audio_s / (sus_ms / 1e3), 2), "unit": UNIT,
            "tflops": round(step_flops / (sus_ms / 1e3) / 1e12, 1), "peak": peaks["tflops_sustained"],
            "frac": round(step_flops / (sus_ms / 1e3) / 1e12 / peaks["tflops_sustained"], 4),
            "tflops_algorithmic": round(algo_flops_per_token * total_tokens / (sus_ms / 1e3) / 1e12, 1),
            "frac_algorithmic": round(algo_flops_per_token * total_tokens / (sus_ms / 1e3) / 1e12 / peaks["tflops_sustained"], 4),
            "sm_mhz_median": sc["sm_mhz"], "reasons": sc["reasons"], "samples": sc["samples"],
            "peak_source": f"{peaks['source']} bf16_tflops_sustained (cuBLAS 8192^3 back to back for 4 s); whole step, executed "
                           "FLOPs; no L2 flush (the step's 374 MB of weights + ~0.3 GB of activations exceed the 126 MB L2)",
        }
        if roofline is not None:
            roofline["sustained"] = sustained

    # ---- small-batch latency (the shape every reference caller uses: B = 1, decoding.py:84-89) ----
    if rank == 0 and world == 1 and not args.no_latency and args.model == "xcodec2":
        one = synthetic_ids(1, 250, 99).pin_memory()
        one_out = torch.empty(250 * dec.samples_per_token, dtype=torch.float32).pin_memory()
        for _ in range(10):
            dec.decode_packed_host(one, [250], out=one_out)
        lat = []
        for _ in range(200):
            t0 = time.perf_counter()
            dec.decode_packed_host(one, [250], out=one_out)
            lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        c1_ids = synthetic_ids(4, 250, 1234).to(dev)
        for _ in range(5):
            dec.decode_packed_device(c1_ids, [250] * 4)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            dec.decode_packed_device(c1_ids, [250] * 4)
        b.record()
        torch.cuda.synchronize(dev)
        c1_ms = a.elapsed_time(b) / 50
        # BASELINE config 5 (streaming, 64 streams x 50 new tokens with 100 tokens of left context), both
        # definitions in steady state; only the NEW audio of a push counts
        from tts_max_b200.codec import streaming
        c5 = {}
        g5 = torch.Generator().manual_seed(5)
        for name, sdec in (("window", streaming.StreamingDecoder(dec, 64, new_tokens=50, left_context=100, use_graph=True)),
                           ("cached", streaming.CachedStreamingDecoder(dec, 64, new_tokens=50, left_context=100, overlap=8))):
            chunks = [torch.randint(0, 65536, (64, 50), generator=g5).to(dev) for _ in range(8)]
            for k in range(6):                      # through the warm-up into steady state (and graph capture)
                sdec.push(chunks[k % 8])
            torch.cuda.synchronize(dev)
            a5, b5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a5.record()
            for k in range(40):
                sdec.push(chunks[k % 8])
            b5.record()
            torch.cuda.synchronize(dev)
            ms5 = a5.elapsed_time(b5) / 40
            c5[name] = {"ms_per_push": round(ms5, 4), "value": round(64.0 / (ms5 / 1e3), 1), "unit": UNIT}
        c5["what"] = ("64 streams x 50 new tokens per push, 100 tokens of left context, 40 steady-state pushes, CUDA events; "
                      "window = re-decode of [context | new] (reference semantics on the window), cached = key / value rings + "
                      "8 recomputed overlap rows (different function, own oracle)")
        latency = {"c5_streaming": c5, "b1_250_tokens_e2e_ms": {"p50": round(lat[len(lat) // 2], 4), "p99": round(lat[int(len(lat) * 0.99) - 1], 4),
                                            "n": len(lat), "api": "Decoder.decode_packed_host (host ids in, host PCM out, sync inside)"},
                   "c1_4x250": {"ms_per_step": round(c1_ms, 4), "value": round(20.0 / (c1_ms / 1e3), 1), "unit": UNIT,
                                "timing": "CUDA events around 50 back-to-back steps, L2-warm"}}

    encode = None
    if rank == 0 and world == 1 and not args.no_encode and args.model == "xcodec2":
        encode = measure_encode(dev, args.precision)

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "xcodec2":
        # the reference algorithm on this box's host cores, at N = 1 only (at N > 1 the other ranks would
        # spin at a barrier on the cores the CPU decode needs)
        clips = cpu_clips(n_utts, tokens)
        v, ms, cores = cpu_decode_rate(tokens, clips, steps=2, warmup=1, new_tokens=NEW_TOKENS.get(args.workload))
        cpu_baseline = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": cpu_sample_text(clips, n_utts, tokens) + f"; 2 timed passes after 1 warm-up ({ms:.0f} ms each)"}

    # ---- BASELINE config 3 (strong scaling) at this N: rides along in every line so that the driver's
    #      1/2/4/8 sweep carries a strong-scaling curve next to the weak-scaling headline ----
    c3 = None
    if not args.no_c3 and args.model == "xcodec2":
        c3 = measure_c3(dec, dev, rank, world, args.c3_utts, args.precision, steps=1, warmup=1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}" + (" [48 kHz upsampler variant, not a BASELINE config]" if args.model == "48k" else ""),
                       "utterances_per_gpu": n_utts, "tokens_per_utterance": tokens,
                       "audio_seconds_per_step_per_gpu": audio_s, "weights": "random-init, reference distributions (seed 0)",
                       "parallelism": f"dp{world} (independent utterances, no data-path collective)",
                       "l2": "256 MiB device buffer written between timed steps (outside the per-step event bracket)",
                       "timing": "CUDA events per step on the launching stream, summed; max over ranks"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": ids_host.numel() * 8 * world,
                    "d2h_bytes_per_step": wav_host.numel() * 4 * world,
                    "api": "Decoder.decode_packed_host -> b200codec_decode_host (pinned host ids in, pinned host PCM out, sync inside)"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "c3_strong": c3,
            "latency": latency,
            "encode": encode,
            "stage_ms_per_step": {k: round(v, 4) for k, v in stage_ms_in_step.items()},
            "stage_ms_event_fenced": {k: round(v, 4) for k, v in stage_ms.items()},
            "wall_s_timed_region": round(t_wall, 4),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3"])
    ap.add_argument("--c3-utts", type=int, default=C3_UTTS)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--model", default="xcodec2", choices=["xcodec2", "48k"],
                    help="xcodec2: 16 kHz, hop 320 (BASELINE configs); 48k: upsampler variant (hop 160, factors [3, 2])")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="seconds of back-to-back steps for roofline.sustained (0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the config-3 strong-scaling block")
    ap.add_argument("--no-latency", action="store_true", help="skip the B = 1 / config-1 latency block")
    ap.add_argument("--no-encode", action="store_true", help="skip the encode-direction block (SURVEY 8f-3)")
    ap.add_argument("--no-early-weights", action="store_true", help="A/B: GEMM weight loads only after griddepcontrol.wait")
    ap.add_argument("--chain", action="store_true", help="A/B: per-block GEMM chains (one persistent launch for c_proj -> fc1 -> fc2 -> next c_attn)")
    ap.add_argument("--istft-tile", type=int, default=0, help="A/B: output hops per ISTFT CTA (12 or 28)")
    ap.add_argument("--no-pdl", action="store_true", help="A/B: plain stream-ordered launches instead of programmatic dependent launch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
