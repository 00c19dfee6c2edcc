"""A/B: one decode step replayed from a CUDA graph vs launched on the stream (development tool).
    python tools/graph_ab.py [--workload c2]
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
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
_, n_utts, tokens = WORKLOADS[args.workload]
dec = decoder.Decoder(16000, 320, None, None, precision="bf16", init_seed=0).to("cuda").eval()
ids = synthetic_ids(n_utts, tokens, 1234).cuda()
seqlens = [tokens] * n_utts
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / args.steps


for _ in range(3):
    ref = dec.decode_packed_device(ids, seqlens)
torch.cuda.synchronize()
print(f"stream launches: {timed(lambda: dec.decode_packed_device(ids, seqlens)):.4f} ms/step")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        dec.decode_packed_device(ids, seqlens)
s.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=s):
        wav = dec.decode_packed_device(ids, seqlens)
    g.replay()
    torch.cuda.synchronize()
    print("graph output identical:", bool(torch.equal(wav, ref)))
    print(f"graph replay   : {timed(g.replay):.4f} ms/step")
    print(f"stream launches: {timed(lambda: dec.decode_packed_device(ids, seqlens)):.4f} ms/step")
    print(f"graph replay   : {timed(g.replay):.4f} ms/step")
except Exception as e:  # noqa: BLE001
    print("capture failed:", repr(e)[:400])
