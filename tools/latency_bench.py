"""Single-utterance latency of the reference-facing call AudioDecoder.decode(ids) (B = 1, host ids in,
host PCM out), plus small varlen batches through decode_batch."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_max_b200.codec import decoding

cfg = decoding.DecoderConfig("", 16000, 50, 320, None, None)
dec = decoding.AudioDecoder(None, cfg, device="cuda")
g = torch.Generator().manual_seed(0)
for T in (50, 250, 500, 1000, 1792):
    ids = torch.randint(0, 65536, (T,), generator=g)
    for _ in range(5):
        dec.decode(ids)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); dec.decode(ids); ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"decode T={T:5d} ({T/50:5.1f} s audio): p50 {ts[15]*1e3:7.3f} ms  p90 {ts[27]*1e3:7.3f} ms  -> {T/50/ts[15]:8.0f} audio-s/s")
for n, T in ((8, 250), (32, 250)):
    utts = [torch.randint(0, 65536, (T,), generator=g) for _ in range(n)]
    for _ in range(3):
        dec.decode_batch(utts)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); dec.decode_batch(utts); ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"decode_batch {n} x T={T}: p50 {ts[5]*1e3:7.3f} ms -> {n*T/50/ts[5]:8.0f} audio-s/s")
