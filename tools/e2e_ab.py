"""A/B of the end-to-end path (host ids in, host PCM out): zero-copy PCM output vs staged D2H copy.
    python tools/e2e_ab.py [--workload c2]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, synthetic_ids  # noqa: E402
from tts_max_b200 import _lib  # noqa: E402
from tts_max_b200.codec import decoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
_, n_utts, tokens = WORKLOADS[args.workload]
dec = decoder.Decoder(16000, 320, None, None, precision="bf16", init_seed=0).to("cuda").eval()
ids = synthetic_ids(n_utts, tokens, 1234).pin_memory()
seqlens = [tokens] * n_utts
out = torch.empty(n_utts * tokens * dec.samples_per_token, dtype=torch.float32).pin_memory()
ref = None
for mode in (1, 0, 1, 0):
    _lib.check(_lib.load().b200codec_set_zero_copy_output(mode))
    for _ in range(5):
        dec.decode_packed_host(ids, seqlens, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.decode_packed_host(ids, seqlens, out=out)
    dt = (time.perf_counter() - t0) / args.steps
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(ref, out))
    print(f"zero_copy={mode}: {dt * 1e3:.3f} ms/step  {n_utts * tokens / 50 / dt:.0f} audio-s/s  identical={same}")
