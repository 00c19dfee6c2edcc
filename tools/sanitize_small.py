"""Small decode (3 utterances, one ragged) for compute-sanitizer: exercises every kernel once."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_max_b200.codec import decoder
dec = decoder.Decoder(16000, 320, None, None, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16", init_seed=0).to("cuda").eval()
g = torch.Generator().manual_seed(0)
lens = [130, 7, 257]
ids = torch.randint(0, 65536, (sum(lens),), generator=g)
wav = dec.decode_packed_host(ids, lens)
print("ok", wav.shape, bool(torch.isfinite(wav).all()), float(wav.abs().max()))
