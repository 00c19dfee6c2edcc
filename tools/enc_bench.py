"""Encode-direction throughput: AcousticEncoder + SemanticEncoder + fusion + quantise on one B200.

    python tools/enc_bench.py [--clips 16] [--seconds 10] [--steps 10] [--group G] [--cpu]

Prints one JSON line: audio-s/s device-timed (CUDA events around `steps` passes over the batch, inputs resident),
launches per pass (80 per launch sequence; --group: clips per launch sequence), and -- with --cpu -- the oracle port of the reference modules on the host cores for one
clip. The w2v-BERT hidden state is a synthetic input (that model stays in HuggingFace).
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import encoder_oracle as E  # noqa: E402  (weights + the CPU baseline only)
from tts_max_b200.codec import encoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=16)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--group", type=int, default=0, help="clips per launch sequence (0: the library default, all that fit)")
args = ap.parse_args()

S = int(args.seconds * 16000) // 320 * 320
T = S // 320
sd = E.make_state_dict(seed=0)
enc = encoder.Encoder(pre_bound=False, precision=args.precision)
enc.load_state_dict(sd)
enc.to("cuda").eval()
if args.group > 0:
    enc.max_batch_samples = args.group * (S + 1920)
g = torch.Generator().manual_seed(7)
wav = (0.3 * torch.randn(args.clips, 1, S, generator=g)).cuda()
w2v = torch.randn(args.clips, T, 1024, generator=g).cuda()
for _ in range(2):
    ids = enc(wav, w2v)
torch.cuda.synchronize()
n0 = enc.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    ids = enc(wav, w2v)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
audio_s = args.clips * S / 16000
# algorithmic FLOPs per audio second (2 * MAC): residual units, down-sampling convs, final conv, semantic, fusion
flops = 0
c, rate = 48, 16000
for s in (2, 2, 4, 4, 5):
    flops += rate * 3 * (2 * c * c * 7 + 2 * c * c)
    rate //= s
    flops += rate * 2 * (2 * c) * c * 2 * s
    c *= 2
flops += 50 * (2 * 1024 * 1536 * 3 + 4 * 2 * 1024 * 1024 * 3 + 2 * 2048 * 2048)
line = {"metric": "codec ENCODE audio-sec/sec (device-timed; acoustic + semantic encoders, fusion, quantise; w2v-BERT excluded)",
        "value": round(audio_s / (ms / 1e3), 1), "unit": "audio-s/s", "ms_per_step": round(ms, 3), "dtype": args.precision,
        "config": {"workload": f"{args.clips} clips x {S / 16000:.1f} s, {args.group or args.clips} clip(s) per launch sequence", "tokens_per_clip": T},
        "gpu_launches_per_step": (enc.launch_count() - n0) // args.steps,
        "algorithmic_gflop_per_audio_s": round(flops / 1e9, 2),
        "achieved_tflops": round(flops * audio_s / (ms / 1e3) / 1e12, 1)}
if args.cpu:
    torch.set_num_threads(os.cpu_count() or 1)
    w1, f1 = wav[:1].cpu(), w2v[:1].cpu()
    E.encoder_hidden(sd, w1[..., :32000], f1[:, :100])
    t0 = time.perf_counter()
    E.encoder_hidden(sd, w1, f1)
    dt = time.perf_counter() - t0
    line["cpu_baseline"] = {"value": round(S / 16000 / dt, 2), "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"1 clip of {S / 16000:.1f} s, oracle port of the reference modules, fp32, torch CPU"}
print(json.dumps(line))
