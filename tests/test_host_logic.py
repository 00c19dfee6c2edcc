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
        decoder.Decoder(96000, 160, [3, 2, 2], [7, 6, 4])  # three stages are not instantiated
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
