"""CPU: the Python mirror of the reference interface (config, construction, checkpoint layouts,
state-dict contract, error behaviour) -- everything that does not launch a kernel."""

import json
import os

import pytest
import torch

from oracle import weights
from tts_max_b200.codec import decoder, decoding


def test_decoder_config_from_json(tmp_path):
    cfg = {"sample_rate": 16000, "token_rate": 50, "hop_length": 320, "upsample_factors": None, "kernel_sizes": None}
    p = tmp_path / "model_config.json"
    p.write_text(json.dumps(cfg))  # the shipped xcodec2 example has no "model_type"
    c = decoding.DecoderConfig.from_json(p)
    assert (c.model_type, c.sample_rate, c.token_rate, c.hop_length) == ("", 16000, 50, 320)
    cfg["model_type"] = "xcodec2"
    p.write_text(json.dumps(cfg))
    assert decoding.DecoderConfig.from_json(p).model_type == "xcodec2"
    with pytest.raises(dataclasses_error()):
        c.sample_rate = 1  # frozen, like the reference


def dataclasses_error():
    import dataclasses
    return dataclasses.FrozenInstanceError


def test_create_requires_model_config(tmp_path):
    with pytest.raises(ValueError, match="No model_config.json"):
        decoding.create(str(tmp_path / "ckpt.pt"))


def test_rate_check_matches_reference():
    with pytest.raises(ValueError, match="do not match the target"):
        decoder.Decoder(16000, 300, None, None)
    with pytest.raises(ValueError, match="do not match the target"):
        decoder.Decoder(48000, 160, [3, 2, 2], [7, 6, 4])
    with pytest.raises(NotImplementedError):
        decoder.Decoder(192000, 160, [3, 2, 2, 2], [7, 6, 4, 4])  # four stages (64 channels) are not instantiated
    with pytest.raises(NotImplementedError):
        decoder.Decoder(4800, 96, None, None)  # n_fft 384 is not 64 x (a multiple of 5)
    assert decoder.Decoder(96000, 160, [3, 2, 2], [7, 6, 4], init_seed=0).samples_per_token == 1920  # three stages
    d = decoder.Decoder(48000, 160, [3, 2], [7, 6], init_seed=0)  # the 48 kHz upsampler variant
    assert d.samples_per_token == 960 and len(d.state_dict()) == 145


def test_state_dict_contract(golden, state_dict):
    d = decoder.Decoder(16000, 320, None, None, init_seed=0)
    sd = d.state_dict()
    assert list(sd.keys()) == list(state_dict.keys())  # golden generator asserted == reference order
    assert all(sd[k].shape == state_dict[k].shape and sd[k].dtype == torch.float32 for k in sd)
    assert torch.equal(sd["decoder.head.istft.window"], torch.hann_window(1280))
    conv = sd["decoder.backbone.embed.weight"]
    assert abs(conv.std().item() - 0.02) < 2e-3 and conv.abs().max().item() <= 2.0
    assert sd["decoder.backbone.embed.bias"].abs().max().item() == 0.0


def test_load_state_dict_is_strict(state_dict):
    d = decoder.Decoder(16000, 320, None, None, init_seed=0)
    d.load_state_dict(state_dict)
    assert torch.equal(d.state_dict()["fc_post_a.weight"], state_dict["fc_post_a.weight"])
    bad = dict(state_dict)
    bad.pop("fc_post_a.bias")
    with pytest.raises(RuntimeError, match="Missing key"):
        d.load_state_dict(bad)
    bad = dict(state_dict)
    bad["extra.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        d.load_state_dict(bad)
    bad = dict(state_dict)
    bad["fc_post_a.bias"] = torch.zeros(7)
    with pytest.raises(RuntimeError, match="size mismatch"):
        d.load_state_dict(bad)


@pytest.mark.parametrize("layout", ["xcodec2", "ttsmax"])
def test_load_from_checkpoint_layouts(tmp_path, state_dict, layout):
    ckpt = weights.to_xcodec2_checkpoint(state_dict) if layout == "xcodec2" else weights.to_ttsmax_checkpoint(state_dict)
    path = tmp_path / "ckpt.pt"
    torch.save(ckpt, path)
    (tmp_path / "model_config.json").write_text(json.dumps(
        {"model_type": "", "sample_rate": 16000, "token_rate": 50, "hop_length": 320,
         "upsample_factors": None, "kernel_sizes": None}))
    d = decoder.Decoder(16000, 320, None, None, checkpoint_path=str(path))
    got = d.state_dict()
    assert all(torch.equal(got[k], state_dict[k]) for k in state_dict)


def test_xcodec2_layout_missing_key_fails(tmp_path, state_dict):
    ckpt = weights.to_xcodec2_checkpoint(state_dict)
    ckpt["state_dict"].pop("generator.head.out.bias")
    path = tmp_path / "ckpt.pt"
    torch.save(ckpt, path)
    with pytest.raises(RuntimeError, match="Missing key"):
        decoder.Decoder(16000, 320, None, None, checkpoint_path=str(path))


def test_no_cpu_fallback():
    d = decoder.Decoder(16000, 320, None, None, init_seed=0)
    with pytest.raises(RuntimeError, match="no CPU"):
        d(torch.zeros(1, 4, dtype=torch.int64))


def test_input_validation_before_any_launch():
    d = decoder.Decoder(16000, 320, None, None, init_seed=0)
    with pytest.raises(ValueError):
        d(torch.zeros(1, 0, dtype=torch.int64))
    with pytest.raises(TypeError):
        d(torch.zeros(1, 4))
    with pytest.raises(ValueError):
        d(torch.zeros(2, 2, 4, dtype=torch.int64))


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "tts_max_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("name,hop,ups,ks", [("xcodec2", 320, None, None), ("48k", 160, [3, 2], [7, 6])])
def test_random_init_matches_reference_statistics(name, hop, ups, ks):
    """`random_init_state_dict` restates the reference's init DISTRIBUTIONS
    (decoder_modules.py:403-433, 13-16, 463-464; torch defaults elsewhere). Golden = per-tensor
    (mean, std, min, max) of the unmodified reference `Decoder(...)` averaged over seeds 0..2
    (oracle/make_golden_init.py). Same key order, same shapes; mean / std agree statistically, bounded
    distributions (uniform, constants) agree in their range."""
    import math

    import numpy as np

    from tts_max_b200.codec import decoder

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_init_stats.npz"))
    keys = [str(k) for k in g[f"{name}/keys"]]
    mine = [decoder.random_init_state_dict(hop, seed=s, upsample_factors=ups, kernel_sizes=ks) for s in (10, 11, 12)]
    assert list(mine[0].keys()) == keys
    for k in keys:
        ref_mean, ref_std, ref_min, ref_max = (float(v) for v in g[f"{name}/stats/{k}"])
        shape = tuple(int(d) for d in g[f"{name}/shape/{k}"])
        assert tuple(mine[0][k].shape) == shape, k
        n = math.prod(shape)
        vals = torch.stack([m[k].double().flatten() for m in mine])
        mean, std = vals.mean().item(), (vals.std(dim=1).mean().item() if n > 1 else 0.0)
        lo, hi = vals.min(dim=1).values.mean().item(), vals.max(dim=1).values.mean().item()
        if ref_std == 0.0 or k.endswith("istft.window") or k.endswith("weight_g"):
            # constants (norm weights / biases, zero conv biases), the Hann window, and weight_g = ||v||
            assert abs(mean - ref_mean) <= 2e-2 * max(abs(ref_mean), 1e-12) + 1e-12, k
            assert abs(std - ref_std) <= 3e-2 * max(ref_std, 1e-12) + 1e-12, k
            continue
        # mean of 3 seeds x n samples on both sides: allow 6 standard errors
        assert abs(mean - ref_mean) <= 6.0 * ref_std * math.sqrt(2.0 / (3 * n)) + 1e-12, (k, mean, ref_mean)
        assert abs(std - ref_std) <= ref_std * (0.02 + 4.0 / math.sqrt(3 * n)), (k, std, ref_std)
        uniform = abs(ref_max / ref_std - math.sqrt(3.0)) < 0.1      # U(-b, b): max = b = sqrt(3) std
        if uniform:
            assert abs(hi - ref_max) <= 0.03 * ref_max + 2.0 * ref_max / n, (k, hi, ref_max)
            assert abs(lo - ref_min) <= 0.03 * abs(ref_min) + 2.0 * abs(ref_min) / n, (k, lo, ref_min)
        # trunc_normal_(std=0.02) with torch's default cut at +-2 (absolute) is an unbounded-looking
        # normal: its extremes are tail events, only mean / std are compared
