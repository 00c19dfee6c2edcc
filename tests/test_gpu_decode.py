"""GPU: the whole decode path through the reference-facing API against the reference's golden
vectors and the oracle, plus the size-independent properties at BASELINE config sizes.

Stated tolerances (vs the fp32 reference, weights = oracle.weights seed 0; peak |wav| ~ 1.3e-2):
  bf16 operands: SNR >= 40 dB, max-abs <= 5e-2 * peak and <= 1e-3   (the north star's example bar; measured 44 dB)
  fp16 operands: SNR >= 55 dB, max-abs <= 1e-2 * peak and <= 1e-3   (measured 62 dB)
The FSQ lookup inside is bit-exact (tests/test_gpu_kernels.py).
"""

import json

import numpy as np
import pytest
import torch

from oracle import codec_oracle as O
from oracle import weights
from tts_max_b200.codec import decoder, decoding

pytestmark = pytest.mark.gpu

TOL = {"bf16": (40.0, 5e-2), "fp16": (55.0, 1e-2)}
MAX_ABS = 1e-3


def check_wave(ref, got, prec, what=""):
    snr_min, rel = TOL[prec]
    snr = O.snr_db(ref, got)
    peak = ref.abs().max().item()
    maxabs = (ref.double() - got.double()).abs().max().item()
    print(f"[parity] {what} {prec}: SNR {snr:.1f} dB, max-abs {maxabs:.3e}, peak {peak:.3e}")
    assert torch.isfinite(got).all()
    assert snr >= snr_min, f"{what}: SNR {snr:.1f} dB < {snr_min}"
    assert maxabs <= rel * peak, f"{what}: max-abs {maxabs:.3e} > {rel} * peak {peak:.3e}"
    assert maxabs <= MAX_ABS, f"{what}: max-abs {maxabs:.3e} > {MAX_ABS}"


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("name", ["u37", "u5", "u1"])
def test_forward_vs_reference_golden(gpu_decoders, golden, prec, name):
    d = gpu_decoders[prec]
    ids = torch.from_numpy(golden[f"{name}_ids"])
    wav = d(ids.view(1, -1).cuda())
    assert wav.shape == (1, 1, 320 * ids.numel()) and wav.dtype == torch.float32 and wav.is_cuda
    check_wave(torch.from_numpy(golden[f"{name}_wav"])[0], wav[0, 0].cpu(), prec, name)


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_batched_forward_vs_reference_golden(gpu_decoders, golden, prec):
    d = gpu_decoders[prec]
    ids = torch.from_numpy(golden["b2x24_ids"]).cuda()
    wav = d(ids)                       # (B, T)
    wav3 = d(ids.unsqueeze(1))         # (B, 1, T)
    assert wav.shape == (2, 1, 320 * 24)
    assert torch.equal(wav, wav3)
    check_wave(torch.from_numpy(golden["b2x24_wav"]), wav.cpu(), prec, "b2x24")


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_frontend_fold_equals_lookup_plus_conv7(gpu_decoders, golden, prec):
    """ids -> embed output: the folded forms (mode 1: im2col + K = 128 tensor-core GEMM with hi/lo
    coefficients, the default; mode 2: fp32 FMA lookup) against lookup + conv7 GEMM (mode 0): the same
    function, all within tolerance of the reference, the fp64-folded ones at least as close; edge
    frames (utterances of 1..7 tokens, where taps fall off the ends) included."""
    from tts_max_b200 import _lib

    d = gpu_decoders[prec]
    lib = _lib.load()
    ids = torch.from_numpy(golden["u37_ids"]).view(1, -1).cuda()
    ref = torch.from_numpy(golden["u37_wav"])[0]
    g = torch.Generator().manual_seed(321)
    lens = [1, 2, 3, 4, 5, 6, 7, 40]
    short = torch.randint(0, 65536, (sum(lens),), generator=g)
    wav, shorts = {}, {}
    try:
        for mode in (0, 2, 1):
            _lib.check(lib.b200codec_set_frontend_fold(mode))
            wav[mode] = d(ids)[0, 0].cpu()
            shorts[mode] = d.decode_packed_host(short, lens).clone()
    finally:
        _lib.check(lib.b200codec_set_frontend_fold(1))
    for mode, what in ((0, "lookup+conv7"), (2, "folded fp32 FMA"), (1, "folded tensor-core")):
        check_wave(ref, wav[mode], prec, f"u37 {what}")
    assert O.snr_db(ref, wav[1]) >= O.snr_db(ref, wav[0]) - 0.5
    assert O.snr_db(ref, wav[2]) >= O.snr_db(ref, wav[0]) - 0.5
    # the two folds differ only by the hi/lo split of the coefficients (~2^-17) and summation order
    check_wave(shorts[2], shorts[1], prec, "short utterances, tensor-core vs FMA fold")
    check_wave(shorts[0], shorts[1], prec, "short utterances, folded vs lookup+conv7")
    with pytest.raises(_lib.B200CodecError):
        _lib.check(lib.b200codec_set_frontend_fold(3))


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_config1_vs_oracle(gpu_decoders, state_dict, prec):
    """BASELINE config 1: 4 clips x 5 s against the fp32 CPU oracle."""
    d = gpu_decoders[prec]
    ids = torch.randint(0, 65536, (4, 250), generator=torch.Generator().manual_seed(1234))
    ref = O.decoder_forward(state_dict, ids)
    wav = d(ids.cuda()).cpu()
    check_wave(ref, wav, prec, "config1 4x250")


def test_audio_decoder_decode_contract(tmp_path, state_dict, golden):
    """decoding.create(...).decode(ids) -> (1, 320 T) float32 CPU, both checkpoint layouts."""
    outs = []
    for layout in ("xcodec2", "ttsmax"):
        dirp = tmp_path / layout
        dirp.mkdir()
        ckpt = weights.to_xcodec2_checkpoint(state_dict) if layout == "xcodec2" else weights.to_ttsmax_checkpoint(state_dict)
        torch.save(ckpt, dirp / "ckpt.pt")
        (dirp / "model_config.json").write_text(json.dumps(
            {"sample_rate": 16000, "token_rate": 50, "hop_length": 320, "upsample_factors": None, "kernel_sizes": None}))
        dec = decoding.create(str(dirp / "ckpt.pt"), device="cuda")
        assert dec.sample_rate == 16000 and dec.token_rate == 50
        ids = torch.from_numpy(golden["u37_ids"])
        wav = dec.decode(ids)
        assert wav.shape == (1, 320 * 37) and wav.dtype == torch.float32 and wav.device.type == "cpu"
        check_wave(torch.from_numpy(golden["u37_wav"]), wav, "bf16", f"decode {layout}")
        wav_dev_ids = dec.decode(ids.cuda())            # rewards.py passes device ids
        assert torch.equal(wav, wav_dev_ids)
        wav_i32 = dec.decode(ids.to(torch.int32))
        assert torch.equal(wav, wav_i32)
        outs.append(wav)
        sd = dec._decoder.state_dict()
        assert all(torch.equal(sd[k], state_dict[k]) for k in state_dict)
        del dec
    assert torch.equal(outs[0], outs[1])


def test_varlen_batch_equals_single_decodes(gpu_decoders):
    """Packed varlen decode reproduces each utterance's single decode (no padding semantics)."""
    d = gpu_decoders["bf16"]
    g = torch.Generator().manual_seed(77)
    lens = [1, 2, 3, 17, 64, 65, 129, 250]
    utts = [torch.randint(0, 65536, (n,), generator=g) for n in lens]
    packed = d.decode_packed_host(torch.cat(utts), lens)
    off = 0
    for ids in utts:
        single = d.decode_packed_host(ids, [ids.numel()])
        got = packed[off * 320:(off + ids.numel()) * 320]
        assert got.shape == single.shape
        # identical kernels and K order per row; only fp64 GroupNorm atomics may reorder
        assert (got - single).abs().max().item() <= 1e-5 * max(1e-3, single.abs().max().item()), ids.numel()
        off += ids.numel()


def test_row_of_batch_equals_single(gpu_decoders):
    d = gpu_decoders["bf16"]
    ids = torch.randint(0, 65536, (3, 100), generator=torch.Generator().manual_seed(8)).cuda()
    batch = d(ids)
    for b in range(3):
        single = d(ids[b:b + 1])
        assert (batch[b] - single[0]).abs().max().item() <= 1e-5 * max(1e-3, single.abs().max().item())


def test_narrow_gemm_tiles_do_not_change_the_decode(gpu_decoders):
    """B = 1 and config-1-sized decodes use 256 x 64 GEMM tiles; switching them off must give the very
    same samples (and therefore the batch invariance above is unaffected by which tiling a batch gets)."""
    from tts_max_b200 import _lib

    d = gpu_decoders["bf16"]
    lib = _lib.load()
    g = torch.Generator().manual_seed(41)
    cases = [torch.randint(0, 65536, (1, 250), generator=g), torch.randint(0, 65536, (4, 250), generator=g)]
    try:
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(0))
        wide = [d(c.cuda()).clone() for c in cases]
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(1))
        narrow = [d(c.cuda()).clone() for c in cases]
    finally:
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(1))
    for a, b in zip(wide, narrow):
        # identical GEMM results; only the fp64 GroupNorm atomics may reorder
        assert (a - b).abs().max().item() <= 1e-6 * max(1e-3, a.abs().max().item())


def test_determinism(gpu_decoders):
    d = gpu_decoders["bf16"]
    ids = torch.randint(0, 65536, (2, 300), generator=torch.Generator().manual_seed(9)).cuda()
    a = d(ids).clone()
    b = d(ids)
    assert (a - b).abs().max().item() <= 1e-6 * max(1e-3, a.abs().max().item())


def test_errors_are_python_exceptions(gpu_decoders, tmp_path):
    d = gpu_decoders["bf16"]
    with pytest.raises(ValueError):
        d.decode_packed_host(torch.tensor([5, 70000, 3]), [3])          # out of range, host path
    with pytest.raises(ValueError):
        d.decode_packed_host(torch.tensor([5, -1, 3]), [3])
    with pytest.raises(ValueError):
        d.decode_packed_host(torch.tensor([5, 1, 3]), [3, 0])           # empty utterance
    cfg = decoding.DecoderConfig("", 16000, 50, 320, None, None)
    dec = decoding.AudioDecoder(None, cfg, device="cuda")
    with pytest.raises(ValueError):
        dec.decode(torch.tensor([], dtype=torch.int64))
    with pytest.raises(ValueError):
        dec.decode(torch.tensor([1, 2, 65536]).cuda())                  # device path: flagged by the kernel
    assert dec.decode(torch.tensor([1, 2, 3])).shape == (1, 960)         # still usable afterwards


@pytest.mark.parametrize("shape", [(16, 500), (4, 3000)])
def test_full_size_properties(gpu_decoders, shape):
    """BASELINE configs 2 and 4 at full size: shape/finite, row-of-batch == single decode
    (a size-independent property; the oracle takes tens of seconds to minutes at these sizes)."""
    d = gpu_decoders["bf16"]
    B, T = shape
    ids = torch.randint(0, 65536, (B, T), generator=torch.Generator().manual_seed(B * T)).cuda()
    wav = d(ids)
    assert wav.shape == (B, 1, 320 * T) and torch.isfinite(wav).all()
    single = d(ids[B - 1:B])
    assert (wav[B - 1] - single[0]).abs().max().item() <= 1e-5 * max(1e-3, single.abs().max().item())
    assert wav.abs().max().item() < 1e3


def test_large_batch_index_math(gpu_decoders):
    """560 x 1000 tokens in ONE call: 561 677 padded rows, so rows x 4096 (the fc1 operand) passes 2^31
    elements and rows x 1024 x 4 passes 2^31 bytes -- any 32-bit row offset in a kernel shows up in the
    last utterances. Property: first / middle / last row of the batch == its single decode."""
    d = gpu_decoders["bf16"]
    B, T = 560, 1000
    ids = torch.randint(0, 65536, (B, T), generator=torch.Generator().manual_seed(560)).cuda()
    wav = d(ids)
    assert wav.shape == (B, 1, 320 * T) and torch.isfinite(wav).all()
    for b in (0, B // 2, B - 1):
        single = d(ids[b:b + 1])
        assert (wav[b] - single[0]).abs().max().item() <= 1e-5 * max(1e-3, single.abs().max().item()), b
    del wav
    torch.cuda.empty_cache()


def test_concurrent_callers_share_a_decoder(gpu_decoders):
    """Two threads on ONE decoder (ctypes releases the GIL; the handle's mutex serialises them) and on two decoders:
    every result is the one the utterance gets alone."""
    import threading
    shared = gpu_decoders["bf16"]
    other = gpu_decoders["fp16"]
    g = torch.Generator().manual_seed(77)
    utts = [torch.randint(0, 65536, (t,), generator=g) for t in (31, 250, 7, 129, 64, 300)]
    want = {id(d): [d.decode_packed_host(u, [u.numel()]) for u in utts] for d in (shared, other)}
    errors = []

    def worker(d, order):
        try:
            for _ in range(6):
                for k in order:
                    got = d.decode_packed_host(utts[k], [utts[k].numel()])
                    if not torch.equal(got, want[id(d)][k]):
                        errors.append((id(d), k))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(shared, [0, 1, 2, 3, 4, 5])),
               threading.Thread(target=worker, args=(shared, [5, 3, 1, 4, 2, 0])),
               threading.Thread(target=worker, args=(other, [2, 0, 5, 1, 3, 4]))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:4]


def test_config2_vs_oracle_subsample(gpu_decoders, state_dict):
    """BASELINE config 2 (16 x 10 s, bf16): two of the sixteen clips checked against the oracle."""
    ids = torch.randint(0, 65536, (16, 500), generator=torch.Generator().manual_seed(1234))
    for prec in ("bf16", "fp16"):
        wav = gpu_decoders[prec](ids.cuda()).cpu()
        ref = O.decoder_forward(state_dict, ids[[0, 15]])
        check_wave(ref, wav[[0, 15]], prec, "config2 rows 0,15")


def test_streaming_windows_config5(gpu_decoders, state_dict):
    """BASELINE config 5: 64 windows of (100 context + 50 new) tokens; the oracle is the reference
    forward on the same window, trimmed to the last 16000 samples."""
    d = gpu_decoders["bf16"]
    ids = torch.randint(0, 65536, (64, 150), generator=torch.Generator().manual_seed(5))
    wav = d(ids.cuda()).cpu()[:, 0, -16000:]
    ref = O.decoder_forward(state_dict, ids[:2])[:, 0, -16000:]
    check_wave(ref, wav[:2], "bf16", "config5 windows 0,1")


def test_random_varlen_batches(gpu_decoders):
    """Randomised shapes: many utterances of random length in one call; every waveform must equal its
    single decode (batch invariance) and be finite."""
    d = gpu_decoders["bf16"]
    g = torch.Generator().manual_seed(2025)
    for trial in range(3):
        n = int(torch.randint(2, 40, (1,), generator=g))
        lens = torch.randint(1, 400, (n,), generator=g).tolist()
        utts = [torch.randint(0, 65536, (t,), generator=g) for t in lens]
        packed = d.decode_packed_host(torch.cat(utts), lens)
        assert packed.numel() == 320 * sum(lens) and torch.isfinite(packed).all()
        offs = [0]
        for t in lens:
            offs.append(offs[-1] + t)
        for k in sorted(set([0, n - 1, int(torch.randint(0, n, (1,), generator=g))])):
            single = d.decode_packed_host(utts[k], [lens[k]])
            got = packed[offs[k] * 320:offs[k + 1] * 320]
            assert (got - single).abs().max().item() <= 1e-5 * max(1e-3, single.abs().max().item()), (trial, k, lens[k])


def test_many_short_utterances(gpu_decoders):
    """300 one-to-three-token utterances in one call (plan arrays, per-utterance GroupNorm, tiny tiles)."""
    d = gpu_decoders["bf16"]
    g = torch.Generator().manual_seed(7)
    lens = torch.randint(1, 4, (300,), generator=g).tolist()
    utts = [torch.randint(0, 65536, (t,), generator=g) for t in lens]
    packed = d.decode_packed_host(torch.cat(utts), lens)
    assert torch.isfinite(packed).all()
    single = d.decode_packed_host(utts[150], [lens[150]])
    off = sum(lens[:150])
    assert (packed[off * 320:(off + lens[150]) * 320] - single).abs().max().item() <= 1e-5 * max(1e-3, single.abs().max().item())


# per-stage floors (SNR dB against the reference's own stage tensors). "embed" is the folded front end:
# exact 16-bit codes x hi/lo-split fp64-folded coefficients, i.e. fp32-like; the others carry 16-bit operands.
STAGE_SNR = {"bf16": {"embed": 80.0, "default": 40.0}, "fp16": {"embed": 80.0, "default": 55.0}}


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_stage_taps_vs_reference_golden(gpu_decoders, golden, prec):
    """Stage-level parity ON THE GPU: the debug taps (b200codec_set_stage_taps / _read_stage) against the
    reference's forward-hook captures in the golden file (oracle/make_golden.py: embed, prior_net,
    transformers[0], all transformers, backbone, head.out; decoder_modules.py:390-400, :131), so
    compensating errors between stages cannot hide."""
    d = gpu_decoders[prec]
    ids = torch.from_numpy(golden["b2x24_ids"]).cuda()
    d.set_stage_taps(True)
    try:
        wav = d(ids)
        got = {name: d.read_stage(name) for name in
               ("embed", "prior_net", "tblock0", "transformers", "backbone", "head_linear")}
    finally:
        d.set_stage_taps(False)
    check_wave(torch.from_numpy(golden["b2x24_wav"]), wav.cpu(), prec, "b2x24 (taps on)")
    floors = STAGE_SNR[prec]
    for name, t in got.items():
        ref = torch.from_numpy(golden[f"b2x24_{name}"])
        if name in ("embed", "prior_net"):
            ref = ref.transpose(1, 2)          # (B, C, T) -> (B, T, C)
        ref = ref.reshape(-1, ref.shape[-1])
        assert t.shape == ref.shape, (name, t.shape, ref.shape)
        snr = O.snr_db(ref, t)
        print(f"[stage] {name} {prec}: SNR {snr:.1f} dB")
        assert snr >= floors.get(name, floors["default"]), f"{name}: {snr:.1f} dB"
    with pytest.raises(KeyError):
        d.read_stage("embed")                  # taps are off again: nothing is kept


def test_config4_long_form_vs_oracle(gpu_decoders, state_dict):
    """BASELINE config 4 (60 s clips, T = 3000: 24 key tiles per attention CTA, the longest overlap-add):
    one full-length clip against the fp32 oracle end to end, both precisions."""
    ids = torch.randint(0, 65536, (1, 3000), generator=torch.Generator().manual_seed(4))
    ref = O.decoder_forward(state_dict, ids)
    for prec in ("bf16", "fp16"):
        # decoded inside the full 4 x 3000 batch of the config (row 2), not alone
        batch = torch.randint(0, 65536, (4, 3000), generator=torch.Generator().manual_seed(44))
        batch[2] = ids[0]
        wav = gpu_decoders[prec](batch.cuda()).cpu()
        check_wave(ref[0], wav[2], prec, "config4 1 x 3000 (row 2 of 4 x 3000)")


def test_config3_bucket_vs_oracle(gpu_decoders, state_dict):
    """BASELINE config 3's unit of work: one length-sorted varlen pack (12 utterances, 100..1000 tokens,
    built by the same bucketing the sharded run uses); three members (longest, a middle one, shortest)
    against the fp32 ORACLE -- not against the GPU's own single decode."""
    from tts_max_b200 import sharding

    g = torch.Generator().manual_seed(2024)
    lengths = torch.randint(100, 1001, (12,), generator=g).tolist()
    lengths[3], lengths[7] = 1000, 100
    utts = [torch.randint(0, 65536, (n,), generator=g) for n in lengths]
    (bucket,) = sharding.bucket_by_length(range(12), lengths, max_tokens=16384)
    seqlens = [lengths[i] for i in bucket]
    offs = [0]
    for n in seqlens:
        offs.append(offs[-1] + n)
    packed_ids = torch.cat([utts[i] for i in bucket])
    for prec in ("bf16", "fp16"):
        wav = gpu_decoders[prec].decode_packed_host(packed_ids, seqlens)
        for k in (0, len(bucket) // 2, len(bucket) - 1):
            ref = O.decoder_forward(state_dict, utts[bucket[k]].view(1, -1))[0, 0]
            check_wave(ref, wav[offs[k] * 320:offs[k + 1] * 320], prec, f"config3 pack member {k} (T = {seqlens[k]})")


def test_output_guard_bands(gpu_decoders):
    """compute-sanitizer is closed on this pool, so out-of-bounds stores are hunted by hand: the PCM buffer is
    a slice of a larger tensor filled with a sentinel, for ragged shapes around every tile size (ISTFT tiles of
    12 / 28 hops, 128-row attention tiles, 256-row GEMM blocks); the guard bands must survive and every sample
    inside must have been written."""
    d = gpu_decoders["bf16"]
    g = torch.Generator().manual_seed(11)
    pad = 4096
    for lens in ([1], [11, 12, 13], [27, 28, 29, 1], [127, 128, 129], [255, 256, 257], [500] * 5 + [3], [1200, 37]):
        n = 320 * sum(lens)
        big = torch.full((n + 2 * pad,), 7777.0, device="cuda")
        ids = torch.randint(0, 65536, (sum(lens),), generator=g).cuda()
        out = d.decode_packed_device(ids, lens, out=big[pad:pad + n])
        torch.cuda.synchronize()
        assert out.data_ptr() == big[pad:].data_ptr()
        assert bool((big[:pad] == 7777.0).all()) and bool((big[pad + n:] == 7777.0).all()), lens
        assert torch.isfinite(out).all() and not bool((out == 7777.0).any()), lens
