"""Dataset-sweep throughput of `batching.decode_code_store`: a synthetic code store of utterances of
2-20 s (BASELINE config 3's length mix) decoded to pageable CPU waveforms, bucket by bucket vs two buckets in flight.

    python tools/code_store_bench.py [--utts 2000]

Prints one JSON line: audio-s/s end to end (wall clock around the whole sweep, every waveform materialised on the host).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_max_b200.codec import batching, decoding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=2000)
args = ap.parse_args()
rng = np.random.default_rng(3)
lens = rng.integers(100, 1001, args.utts)
codes = rng.integers(0, 65536, int(lens.sum())).astype(np.int32)
index = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)   # start offsets
with tempfile.TemporaryDirectory() as d:
    codes.tofile(os.path.join(d, "train_codes.npy"))   # raw int32 memmap, as data_vectorizer.py writes it
    np.save(os.path.join(d, "train_codes_index.npy"), index)
    store = batching.CodeStore.open(d, "train")
    dec = decoding.AudioDecoder(None, decoding.DecoderConfig("", 16000, 50, 320, None, None), device="cuda")
    audio_s = float(lens.sum()) / 50.0
    out = {}
    for name, flag in (("sequential", False), ("pipelined", True)):
        for _ in batching.decode_code_store(dec, store, sample_ids=range(64), pipelined=flag):   # warm-up
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for _, wav in batching.decode_code_store(dec, store, pipelined=flag):
            n += wav.shape[1]
        dt = time.perf_counter() - t0
        assert n == int(lens.sum()) * 320
        out[name] = {"value": round(audio_s / dt, 1), "unit": "audio-s/s", "wall_s": round(dt, 3)}
print(json.dumps({"metric": "decode_code_store end to end (int32 code store -> pageable CPU waveforms)",
                  "utterances": args.utts, "audio_seconds": round(audio_s, 1), **out}))
