"""The 48 kHz upsampler variant: sample_rate 48000, hop 160, upsample_factors [3, 2],
kernel_sizes [7, 6] (example/configs/codec_training_config.json:24-37). Goldens come from the unmodified
reference through its tts-max checkpoint layout (oracle/make_golden_48k.py)."""

import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import codec_oracle as O
from oracle import weights
from tts_max_b200 import _lib
from tts_max_b200.codec import decoder, decoding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UP = dict(upsample_factors=[3, 2], kernel_sizes=[7, 6])
TOL = {"bf16": (38.0, 5e-2), "fp16": (55.0, 1e-2)}


@pytest.fixture(scope="module")
def golden48():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_decode_48k_seed0.npz")))


@pytest.fixture(scope="module")
def sd48(golden48):
    sd = weights.make_state_dict(seed=0, perturb=True, hop=160, **UP)
    fp, want = weights.fingerprint(sd), float(golden48["weights_fingerprint"])
    assert abs(fp - want) <= 1e-9 * max(1.0, abs(want)), "torch CPU RNG drift: regenerate with oracle/make_golden_48k.py"
    return sd


def test_oracle_vs_reference_48k(golden48, sd48):
    assert len(sd48) == 145
    for name in ("u29", "u3", "u1"):
        ids = torch.from_numpy(golden48[f"{name}_ids"])
        wav = O.decoder_forward(sd48, ids.view(1, -1), hop=160, **UP)
        ref = torch.from_numpy(golden48[f"{name}_wav"])
        assert wav.shape == (1, 1, 960 * ids.numel())
        assert O.snr_db(ref, wav[0]) >= 90.0
    stages = {}
    wav = O.decoder_forward(sd48, torch.from_numpy(golden48["b2x16_ids"]), hop=160, stages=stages, **UP)
    assert O.snr_db(torch.from_numpy(golden48["b2x16_wav"]), wav) >= 90.0
    ref_up = torch.from_numpy(golden48["b2x16_upsampled"])
    assert (stages["upsampled"] - ref_up).abs().max().item() <= 1e-4 * max(1.0, ref_up.abs().max().item())


def test_weight_norm_restatement():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(6, 4, 7, generator=g)
    gg = torch.rand(6, 1, 1, generator=g) + 0.5
    w = O.weight_norm_weight(gg, v)
    assert torch.allclose(w.reshape(6, -1).norm(dim=1), gg.reshape(6), atol=1e-6)


def test_state_dict_contract_48k(sd48):
    d = decoder.Decoder(48000, 160, [3, 2], [7, 6], init_seed=1)
    assert list(d.state_dict().keys()) == list(sd48.keys())  # == the reference's order (asserted by the generator)
    d.load_state_dict(sd48)
    assert all(torch.equal(d.state_dict()[k], sd48[k]) for k in sd48)
    sd = d.state_dict()
    v = decoder.random_init_state_dict(160, 0, [3, 2], [7, 6])
    # weight_norm initialisation: g = ||v||
    assert torch.allclose(v["upsampler.upsample_layers.0.weight_g"].reshape(-1),
                          v["upsampler.upsample_layers.0.weight_v"].reshape(1024, -1).norm(dim=1), rtol=1e-5)
    assert sd["upsampler.resnet_blocks.1.temb_proj.weight"].shape == (256, 512)


def _check(ref, got, prec, what):
    snr_min, rel = TOL[prec]
    snr = O.snr_db(ref, got)
    peak = ref.abs().max().item()
    maxabs = (ref.double() - got.double()).abs().max().item()
    print(f"[parity 48k] {what} {prec}: SNR {snr:.1f} dB, max-abs {maxabs:.3e}, peak {peak:.3e}")
    assert torch.isfinite(got).all()
    assert snr >= snr_min and maxabs <= rel * peak, f"{what}: SNR {snr:.1f} dB, max-abs {maxabs:.3e}"


@pytest.fixture(scope="module")
def gpu48(sd48):
    out = {}
    for prec in ("bf16", "fp16"):
        d = decoder.Decoder(48000, 160, [3, 2], [7, 6], precision=prec)
        d.load_state_dict(sd48)
        out[prec] = d.to("cuda").eval()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_forward_vs_reference_golden_48k(gpu48, golden48, prec):
    d = gpu48[prec]
    for name in ("u29", "u3", "u1"):
        ids = torch.from_numpy(golden48[f"{name}_ids"])
        wav = d(ids.view(1, -1).cuda())
        assert wav.shape == (1, 1, 960 * ids.numel())
        _check(torch.from_numpy(golden48[f"{name}_wav"])[0], wav[0, 0].cpu(), prec, name)
    wav = d(torch.from_numpy(golden48["b2x16_ids"]).cuda())
    _check(torch.from_numpy(golden48["b2x16_wav"]), wav.cpu(), prec, "b2x16")


@pytest.mark.gpu
def test_create_from_ttsmax_checkpoint_48k(tmp_path, sd48, golden48):
    torch.save(weights.to_ttsmax_checkpoint(sd48), tmp_path / "ckpt.pt")
    (tmp_path / "model_config.json").write_text(json.dumps(
        {"model_type": "", "sample_rate": 48000, "token_rate": 50, "hop_length": 160,
         "upsample_factors": [3, 2], "kernel_sizes": [7, 6]}))
    dec = decoding.create(str(tmp_path / "ckpt.pt"), device="cuda")
    assert dec.sample_rate == 48000 and dec.token_rate == 50
    ids = torch.from_numpy(golden48["u29_ids"])
    wav = dec.decode(ids)
    assert wav.shape == (1, 960 * 29) and wav.device.type == "cpu"
    _check(torch.from_numpy(golden48["u29_wav"]), wav, "bf16", "decode()")


@pytest.mark.gpu
def test_varlen_equals_single_48k(gpu48):
    d = gpu48["bf16"]
    g = torch.Generator().manual_seed(5)
    lens = [1, 2, 45, 130]
    utts = [torch.randint(0, 65536, (n,), generator=g) for n in lens]
    packed = d.decode_packed_host(torch.cat(utts), lens)
    off = 0
    for ids in utts:
        single = d.decode_packed_host(ids, [ids.numel()])
        got = packed[off * 960:(off + ids.numel()) * 960]
        assert (got - single).abs().max().item() <= 1e-5 * max(1e-4, single.abs().max().item()), ids.numel()
        off += ids.numel()


@pytest.mark.gpu
def test_config_size_vs_oracle_48k(gpu48, sd48):
    ids = torch.randint(0, 65536, (2, 250), generator=torch.Generator().manual_seed(6))
    ref = O.decoder_forward(sd48, ids, hop=160, **UP)
    for prec in ("bf16", "fp16"):
        _check(ref, gpu48[prec](ids.cuda()).cpu(), prec, "2x250")


@pytest.mark.gpu
@pytest.mark.parametrize("seqlens", [[1], [5], [12], [13], [40, 3]])
def test_istft_hop160_vs_oracle(gpu48, sd48, seqlens):
    d = gpu48["bf16"]
    h = d._ensure_handle()
    g = torch.Generator().manual_seed(200 + sum(seqlens))
    rows = sum(seqlens)
    x_pred = torch.randn(rows, 642, generator=g)
    x_pred[:, :321] = x_pred[:, :321] * 1.5 - 1.0
    x_pred[:, 321:] *= 4.0
    ld = 672
    xp = torch.zeros(rows, ld)
    xp[:, :642] = x_pred
    xd = xp.cuda()
    wav = torch.full((rows * 160,), float("nan"), device="cuda")
    _lib.check(_lib.load().b200codec_istft(h, ctypes.c_void_p(xd.data_ptr()), ld, _lib.i32_array(seqlens), len(seqlens),
                                           ctypes.c_void_p(wav.data_ptr()),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    win = sd48["decoder.head.istft.window"]
    off = 0
    for T in seqlens:
        ref = O.istft_same(O.head_spectrum(x_pred[off:off + T][None]), win, 160)[0]
        got = wav[off * 160:(off + T) * 160].cpu()
        assert (got - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), f"T={T}"
        off += T
