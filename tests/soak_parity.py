"""Randomised parity / batch-invariance soak (run by hand on a B200: `python tests/soak_parity.py`).
Lives under tests/ because it uses the oracle; not collected by pytest."""
import sys, torch, random
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_max_b200.codec import decoder
from oracle import codec_oracle as O, weights
sd = weights.make_state_dict(seed=0, perturb=True)
worst = 0.0
for prec in ("bf16", "fp16"):
    d = decoder.Decoder(16000, 320, None, None, precision=prec)
    d.load_state_dict(sd); d = d.to("cuda").eval()
    rng = random.Random(7)
    g = torch.Generator().manual_seed(7)
    for it in range(25):
        n = rng.choice([1, 2, 3, 5, 9, 17, 33])
        lens = [rng.choice([1, 2, 7, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 300, 511, 513, 700]) for _ in range(n)]
        utts = [torch.randint(0, 65536, (t,), generator=g) for t in lens]
        packed = d.decode_packed_host(torch.cat(utts), lens)
        assert torch.isfinite(packed).all(), (prec, lens)
        off = 0
        for i, ids in enumerate(utts):
            if i in (0, n - 1):
                single = d.decode_packed_host(ids, [ids.numel()])
                got = packed[off * 320:(off + ids.numel()) * 320]
                err = (got - single).abs().max().item() / max(1e-3, single.abs().max().item())
                worst = max(worst, err)
                assert err <= 1e-5, (prec, lens, i, err)
            off += ids.numel()
        if it % 8 == 0:
            j = lens.index(min(lens, key=lambda t: abs(t - 64)))
            ref = O.decoder_forward(sd, utts[j].view(1, -1))[0, 0]
            o = sum(lens[:j]) * 320
            snr = O.snr_db(ref, packed[o:o + lens[j] * 320])
            print(prec, "T", lens[j], "SNR vs oracle", round(snr, 1))
            assert snr >= (40 if prec == "bf16" else 55)
print("soak ok, worst batch-vs-single rel err", worst)
