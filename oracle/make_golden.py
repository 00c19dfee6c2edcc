"""Generates tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (it needs /root/reference, which does not exist on the GPU
box):   python oracle/make_golden.py

What it does
  1. builds the deterministic test weights (oracle/weights.py, seed 0, perturbed),
  2. saves them as checkpoints in BOTH layouts the reference accepts (decoder.py:94-119) and
     loads each through the reference's own `decoding.create(...)` / `Decoder.load_from_checkpoint`,
  3. runs the reference `Decoder.forward` / `AudioDecoder.decode` on seeded ids and stores inputs
     and outputs (plus stage tensors captured with forward hooks) as float32 npz fixtures.
The two third-party classes the reference imports are provided by oracle/ref_shims (restated;
"parity unpinned" there), everything else executes from /root/reference as is.
"""

from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")

from oracle import weights  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_CONFIG = {"model_type": "", "sample_rate": 16000, "token_rate": 50, "hop_length": 320,
                "upsample_factors": None, "kernel_sizes": None}


def main() -> None:
    from tts.core.codec import decoding as ref_decoding  # the reference, unmodified

    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN, exist_ok=True)
    sd = weights.make_state_dict(seed=0, perturb=True)

    with tempfile.TemporaryDirectory() as tmp:
        decoders = {}
        for name, ckpt in (("xcodec2", weights.to_xcodec2_checkpoint(sd)), ("ttsmax", weights.to_ttsmax_checkpoint(sd))):
            d = os.path.join(tmp, name)
            os.makedirs(d)
            with open(os.path.join(d, "model_config.json"), "w") as f:
                json.dump(MODEL_CONFIG, f)
            path = os.path.join(d, "ckpt.pt")
            torch.save(ckpt, path)
            decoders[name] = ref_decoding.create(path, device="cpu")

    ref = decoders["xcodec2"]
    ref_sd = ref._decoder.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "state-dict key order differs from the reference"
    for k in sd:
        assert torch.equal(ref_sd[k], sd[k]), k
        assert torch.equal(decoders["ttsmax"]._decoder.state_dict()[k], sd[k]), k

    g = torch.Generator().manual_seed(1234)
    out = {"weights_fingerprint": np.float64(weights.fingerprint(sd))}

    # --- K1: FSQ lookup on a handful of ids (incl. the extremes) -------------------------------
    ids_fsq = torch.cat([torch.tensor([0, 1, 2, 3, 4, 21845, 43690, 65535, 255, 256, 4095, 4096]),
                         torch.randint(0, 65536, (20,), generator=g)])
    emb = ref._decoder.decoder.quantizer.get_output_from_indices(ids_fsq.view(1, -1, 1))
    out["fsq_ids"] = ids_fsq.numpy().astype(np.int64)
    out["fsq_out"] = emb[0].detach().numpy().astype(np.float32)

    # --- whole decode, single utterances through AudioDecoder.decode (both layouts) -------------
    for name, T in (("u37", 37), ("u5", 5), ("u1", 1)):
        ids = torch.randint(0, 65536, (T,), generator=g)
        wav = ref.decode(ids)
        wav2 = decoders["ttsmax"].decode(ids)
        assert wav.shape == (1, 320 * T) and wav.dtype == torch.float32 and wav.device.type == "cpu"
        assert torch.equal(wav, wav2), "the two checkpoint layouts disagree"
        out[f"{name}_ids"] = ids.numpy().astype(np.int64)
        out[f"{name}_wav"] = wav.numpy().astype(np.float32)

    # --- batched forward with stage captures ----------------------------------------------------
    ids_b = torch.randint(0, 65536, (2, 24), generator=g)
    cap = {}
    dec = ref._decoder
    hooks = [
        dec.fc_post_a.register_forward_hook(lambda m, i, o: cap.__setitem__("fc_post_a", o.detach())),
        dec.decoder.backbone.register_forward_hook(lambda m, i, o: cap.__setitem__("backbone", o.detach())),
        dec.decoder.backbone.embed.register_forward_hook(lambda m, i, o: cap.__setitem__("embed", o.detach())),
        dec.decoder.backbone.prior_net.register_forward_hook(lambda m, i, o: cap.__setitem__("prior_net", o.detach())),
        dec.decoder.backbone.transformers[0].register_forward_hook(lambda m, i, o: cap.__setitem__("tblock0", o.detach())),
        dec.decoder.backbone.transformers.register_forward_hook(lambda m, i, o: cap.__setitem__("transformers", o.detach())),
        dec.decoder.head.out.register_forward_hook(lambda m, i, o: cap.__setitem__("head_linear", o.detach())),
    ]
    with torch.no_grad():
        wav_b = dec(ids_b)
    for h in hooks:
        h.remove()
    assert wav_b.shape == (2, 1, 320 * 24)
    out["b2x24_ids"] = ids_b.numpy().astype(np.int64)
    out["b2x24_wav"] = wav_b.numpy().astype(np.float32)
    out["b2x24_fc_post_a"] = cap["fc_post_a"].numpy().astype(np.float32)          # (2, 24, 1024)
    out["b2x24_embed"] = cap["embed"].numpy().astype(np.float32)                  # (2, 1024, 24)
    out["b2x24_prior_net"] = cap["prior_net"].numpy().astype(np.float32)          # (2, 1024, 24)
    out["b2x24_tblock0"] = cap["tblock0"].numpy().astype(np.float32)              # (2, 24, 1024)
    out["b2x24_transformers"] = cap["transformers"].numpy().astype(np.float32)    # (2, 24, 1024)
    out["b2x24_backbone"] = cap["backbone"].numpy().astype(np.float32)            # (2, 24, 1024)
    out["b2x24_head_linear"] = cap["head_linear"].numpy().astype(np.float32)      # (2, 24, 1282)

    # --- row-of-batch == single-utterance -------------------------------------
    single = ref.decode(ids_b[1])
    out["b2x24_row1_single_maxabs"] = np.float64((single[0] - wav_b[1, 0]).abs().max())

    path = os.path.join(GOLDEN, "reference_decode_seed0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    for k, v in out.items():
        print(f"  {k:28s} {getattr(v, 'shape', ())} {getattr(v, 'dtype', type(v))}")


if __name__ == "__main__":
    main()
