"""Encode direction: AcousticEncoder + SemanticEncoder + fusion + quantise.

CPU: the oracle restatement (oracle/encoder_oracle.py) against golden outputs of the UNMODIFIED reference
modules (oracle/make_golden_encoder.py), and the host logic of `tts_max_b200.codec.encoder.Encoder` (state-dict
keys, both checkpoint layouts, no CPU fallback).
GPU: the CUDA path through the C ABI (b200enc_*) against the same goldens, stage by stage.

Stated tolerances (vs the fp32 reference modules, oracle weights seed 0): hidden state / encoder outputs
bf16 operands SNR >= 40 dB, fp16 >= 55 dB (some thirty 16-bit convolutions in sequence; measured 44 / 62 dB);
conv_blocks[0] (an fp32 FIR) >= 100 dB; FSQ digits of the ids equal to the reference's on >= 99 % of positions
(a digit can flip where the bounded projection sits within the hidden state's error of a rounding boundary;
measured 100 %).
"""

import collections
import os

import numpy as np
import pytest
import torch

from oracle import codec_oracle as O
from oracle import encoder_oracle as E

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_encoder_seed0.npz")
SNR_MIN = {"bf16": 40.0, "fp16": 55.0}
DIGITS_MIN = {"bf16": 0.99, "fp16": 0.99}


@pytest.fixture(scope="module")
def egolden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def esd():
    return E.make_state_dict(seed=0)


def test_oracle_matches_reference_golden(egolden, esd):
    assert np.allclose(E.kaiser_sinc_filter1d(0.25, 0.3, 12).numpy(), egolden["filter"], rtol=0, atol=1e-7)
    for name in ("b2x12", "b1x50"):
        wav, w2v = torch.from_numpy(egolden[f"{name}_wav"]), torch.from_numpy(egolden[f"{name}_w2v"])
        st = {}
        hidden = E.encoder_hidden(esd, wav, w2v, stages=st)
        for key, got in (("hidden", hidden), ("acoustic", st["acoustic"]), ("semantic", st["semantic"])):
            ref = torch.from_numpy(egolden[f"{name}_{key}"])
            assert got.shape == ref.shape
            assert O.snr_db(ref, got) >= 100.0, (name, key)
        if name == "b2x12":
            for key in ("conv0", "block1", "block3", "block5"):
                assert O.snr_db(torch.from_numpy(egolden[f"{name}_{key}"]), st[key]) >= 100.0, key


def test_encoder_host_logic(tmp_path, esd):
    from tts_max_b200.codec import encoder

    assert list(encoder.expected_state_dict_shapes().items()) == list(E.shapes().items())
    enc = encoder.Encoder(pre_bound=False)
    with pytest.raises(RuntimeError):           # strict: a missing and an unexpected key
        bad = dict(esd)
        bad.pop("fusion_layer.bias")
        bad["nope"] = torch.zeros(1)
        enc.load_state_dict(bad)
    with pytest.raises(RuntimeError):           # shape mismatch
        bad = dict(esd)
        bad["fusion_layer.bias"] = torch.zeros(7)
        enc.load_state_dict(bad)
    enc.load_state_dict({**esd, "wav2vec_model.encoder.x": torch.zeros(1)})   # a full reference state dict
    assert all(torch.equal(v, esd[k]) for k, v in enc.state_dict().items())
    # the xcodec2 checkpoint layout (encoder.py:86-110) and the tts-max one (:111-112)
    x2 = collections.OrderedDict()
    for k, v in esd.items():
        for ours, theirs in (("acoustic_encoder.", "CodecEnc."), ("semantic_encoder.", "SemanticEncoder_module."),
                             ("fusion_layer.", "fc_prior."), ("quantizer.", "generator.quantizer.")):
            if k.startswith(ours):
                x2[theirs + k[len(ours):]] = v
    x2["generator.backbone.embed.weight"] = torch.zeros(2)      # decoder keys are ignored by the encoder
    torch.save({"state_dict": x2}, tmp_path / "x2.pt")
    torch.save(dict(esd), tmp_path / "own.pt")
    for path in ("x2.pt", "own.pt"):
        e2 = encoder.Encoder(str(tmp_path / path), pre_bound=True)
        assert list(e2.state_dict().keys()) == list(esd.keys())
        assert all(torch.equal(v, esd[k]) for k, v in e2.state_dict().items())
    with pytest.raises(RuntimeError):           # no CPU fallback
        enc(torch.zeros(1, 1, 320), torch.zeros(1, 1, 1024))
    with pytest.raises(ValueError):
        enc.to("cpu")(torch.zeros(1, 1, 321), torch.zeros(1, 1, 1024))
    with pytest.raises(TypeError):
        encoder.Encoder()                       # pre_bound has no default (parity unpinned at the library boundary)


@pytest.fixture(scope="module")
def gpu_encoders(esd):
    from tts_max_b200.codec import encoder

    out = {}
    for prec in ("bf16", "fp16"):
        e = encoder.Encoder(pre_bound=False, precision=prec)
        e.load_state_dict(esd)
        out[prec] = e.to("cuda").eval()
    return out


@pytest.mark.gpu
def test_library_key_list(esd):
    from tts_max_b200.codec import encoder

    assert list(encoder.library_state_dict_shapes().items()) == list(encoder.expected_state_dict_shapes().items())


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("name", ["b2x12", "b1x50"])
def test_gpu_encoder_vs_reference_golden(gpu_encoders, egolden, esd, prec, name):
    enc = gpu_encoders[prec]
    wav, w2v = torch.from_numpy(egolden[f"{name}_wav"]), torch.from_numpy(egolden[f"{name}_w2v"])
    B, _, S = wav.shape
    T = S // 320
    if name == "b2x12":
        enc.set_stage_taps(True)
    try:
        code, aux = enc(wav.cuda(), w2v.cuda(), return_hidden=True)
        if name == "b2x12":   # taps hold every clip of the batch: (B, rows, C)
            for key, floor in (("conv0", 100.0), ("block1", SNR_MIN[prec]), ("block3", SNR_MIN[prec]), ("block5", SNR_MIN[prec])):
                got = enc.read_stage(key, S)
                ref = torch.from_numpy(egolden[f"{name}_{key}"]).transpose(1, 2)
                assert got.shape == ref.shape, (key, got.shape, ref.shape)
                snr = O.snr_db(ref, got)
                print(f"[enc stage] {key} {prec}: SNR {snr:.1f} dB")
                assert snr >= floor, (key, snr)
    finally:
        enc.set_stage_taps(False)
    assert code.shape == (B, 1, T) and code.dtype == torch.int32 and code.is_cuda
    for key in ("acoustic", "semantic", "hidden"):
        ref = torch.from_numpy(egolden[f"{name}_{key}"])
        got = aux[key].cpu()
        assert got.shape == ref.shape
        snr = O.snr_db(ref, got)
        print(f"[enc parity] {name} {key} {prec}: SNR {snr:.1f} dB")
        assert torch.isfinite(got).all() and snr >= SNR_MIN[prec], (key, snr)
    # ids: the reference's quantise of the REFERENCE hidden state (oracle restatement of ResidualFSQ.forward)
    want, _, bounded = E.quantize(esd, torch.from_numpy(egolden[f"{name}_hidden"]), pre_bound=False)
    basis = torch.tensor([4 ** d for d in range(8)])
    digits = lambda ids: (ids.long().unsqueeze(-1) // basis) % 4   # noqa: E731
    agree = (digits(code.cpu()[:, 0]) == digits(want[:, 0])).float().mean().item()
    print(f"[enc parity] {name} ids {prec}: {100 * agree:.2f} % of the FSQ digits equal the reference's")
    assert agree >= DIGITS_MIN[prec]
    # and exactly the kernel's own quantise of its own hidden state
    own, _, own_b = E.quantize(esd, aux["hidden"].cpu(), pre_bound=False)
    clear = ((own_b - own_b.round()).abs() - 0.5).abs().min(dim=-1).values > 1e-3
    assert torch.equal(code.cpu()[:, 0][clear], own[:, 0][clear].to(torch.int32))


@pytest.mark.gpu
def test_gpu_encoder_batch_equals_single_and_round_trip(gpu_encoders, gpu_decoders):
    """The clips of a batch share one launch sequence and one padded row space; every clip's ids AND hidden states
    are bit-identical to encoding it alone (also when the batch is split into groups). A 4 s utterance runs at
    full rate counts, and the ids feed the decoder: wav (1, S) -> T = ceil-ish(S / 320) tokens -> 320 T samples."""
    enc = gpu_encoders["bf16"]
    g = torch.Generator().manual_seed(5)
    wav = 0.3 * torch.randn(5, 1, 320 * 40, generator=g)
    w2v = torch.randn(5, 40, 1024, generator=g)
    batch, aux = enc(wav.cuda(), w2v.cuda(), return_hidden=True)
    for b in range(5):
        one, aux1 = enc(wav[b:b + 1].cuda(), w2v[b:b + 1].cuda(), return_hidden=True)
        assert torch.equal(one, batch[b:b + 1])
        for key in ("hidden", "acoustic", "semantic"):
            assert torch.equal(aux1[key], aux[key][b:b + 1]), (b, key)
    saved = enc.max_batch_samples
    try:
        enc.max_batch_samples = 2 * (320 * 40 + 1920)   # groups of 2, 2, 1 clips
        split, aux2 = enc(wav.cuda(), w2v.cuda(), return_hidden=True)
    finally:
        enc.max_batch_samples = saved
    assert torch.equal(split, batch) and torch.equal(aux2["hidden"], aux["hidden"])
    long_wav = 0.3 * torch.randn(1, 64000 - 77, generator=g)            # Encoder.encode pads to 64000
    ids = enc.encode(long_wav, torch.randn(1, 200, 1024, generator=g))
    assert ids.shape == (200,) and ids.min() >= 0 and ids.max() < 65536
    audio = gpu_decoders["bf16"](ids.view(1, -1).long())
    assert audio.shape == (1, 1, 64000) and torch.isfinite(audio).all()
    calls = []
    ids2 = enc.encode(long_wav, lambda audio_pad: (calls.append(tuple(audio_pad.shape)), torch.zeros(1, 200, 1024))[1])
    assert calls == [(1, 64320)] and ids2.shape == (200,)


@pytest.mark.gpu
def test_gpu_encoder_errors(gpu_encoders):
    import ctypes

    from tts_max_b200 import _lib

    enc = gpu_encoders["bf16"]
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 1, 321, device="cuda"), torch.zeros(1, 1, 1024, device="cuda"))
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 1, 640, device="cuda"), torch.zeros(1, 3, 1024, device="cuda"))
    lib = _lib.load()
    x = torch.zeros(640, device="cuda")
    rc = lib.b200enc_encode(enc._ensure_handle(), ctypes.c_void_p(x.data_ptr()), 333, ctypes.c_void_p(x.data_ptr()), None, 0, 0,
                            None, None, None, None)
    assert rc != 0 and b"multiple of 320" in lib.b200codec_last_error()
    with pytest.raises(_lib.B200CodecError):
        enc.read_stage("conv0", 640)          # taps are off
    # ragged lengths back to back (workspace regrowth) stay finite and deterministic
    g = torch.Generator().manual_seed(9)
    for T in (1, 3, 200, 7, 64):
        wav, w2v = 0.3 * torch.randn(1, 1, 320 * T, generator=g), torch.randn(1, T, 1024, generator=g)
        a = enc(wav.cuda(), w2v.cuda())
        assert a.shape == (1, 1, T) and torch.equal(a, enc(wav.cuda(), w2v.cuda()))
