"""One decode step of a BASELINE workload, for ncu / compute-sanitizer captures.

    python tools/profile_step.py --workload c2 --warmup 1 --steps 1
Prints the number of kernel launches per step so `ncu -s/-c` can be set.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, synthetic_ids  # noqa: E402
from tts_max_b200.codec import decoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()
_, n_utts, tokens = WORKLOADS[args.workload]
dec = decoder.Decoder(16000, 320, None, None, precision=args.precision, init_seed=0).to("cuda").eval()
ids = synthetic_ids(n_utts, tokens, 1234).cuda()
seqlens = [tokens] * n_utts
for _ in range(args.warmup):
    dec.decode_packed_device(ids, seqlens)
torch.cuda.synchronize()
n0 = dec.launch_count()
for _ in range(args.steps):
    wav = dec.decode_packed_device(ids, seqlens)
torch.cuda.synchronize()
print("launches_per_step", (dec.launch_count() - n0) // args.steps, "finite", bool(torch.isfinite(wav).all()))
