"""Configurations the reference constructor accepts beyond the two it ships (decoder.py:17-67, upsampler.py:9-60):
any (hop_length, upsample_factors) with sample_rate // hop // prod(factors) == 50. Instantiated here: hop 320 / 240 /
160 / 80 (n_fft = 4 hop = 64 x {20, 15, 10, 5}: the ISTFT's radix-8 x radix-8 x prime-factor DFT) and up to three
upsampler stages (512 / 256 / 128 channels). The oracle is the CPU restatement of the reference modules, which is
generic in both; weights are the seeded random init of oracle/weights.py for that configuration."""

import ctypes

import pytest
import torch

from oracle import codec_oracle as O
from oracle import weights
from tts_max_b200 import _lib
from tts_max_b200.codec import decoder

CONFIGS = {
    "hop240": dict(sample_rate=12000, hop=240, factors=None, kernels=None),
    "hop80": dict(sample_rate=4000, hop=80, factors=None, kernels=None),
    "hop240_x2x2": dict(sample_rate=48000, hop=240, factors=[2, 2], kernels=[4, 4]),
    "hop80_x3x2x2": dict(sample_rate=48000, hop=80, factors=[3, 2, 2], kernels=[7, 6, 4]),
}
TOL = {"bf16": (38.0, 5e-2), "fp16": (55.0, 1e-2)}  # as for the 48 kHz variant (tests/test_upsampler.py)


def make(name, prec="bf16"):
    c = CONFIGS[name]
    sd = weights.make_state_dict(seed=0, perturb=True, hop=c["hop"], upsample_factors=c["factors"], kernel_sizes=c["kernels"])
    d = decoder.Decoder(c["sample_rate"], c["hop"], c["factors"], c["kernels"], precision=prec)
    d.load_state_dict(sd)
    return c, sd, d


def test_state_dict_shapes_follow_the_configuration():
    for name, c in CONFIGS.items():
        d = decoder.Decoder(c["sample_rate"], c["hop"], c["factors"], c["kernels"], init_seed=0)
        ref = weights.shapes(c["hop"], upsample_factors=c["factors"], kernel_sizes=c["kernels"])
        assert {k: tuple(v.shape) for k, v in d.state_dict().items()} == {k: tuple(v) for k, v in ref.items()}, name
        total = 1
        for f in c["factors"] or []:
            total *= f
        assert d.samples_per_token == c["hop"] * total


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hop240", "hop80"])
@pytest.mark.parametrize("seqlens", [[1], [2], [12], [13], [37, 5], [250]])
def test_istft_other_hops_vs_oracle(name, seqlens):
    c, sd, d = make(name)
    d = d.to("cuda").eval()
    hop = c["hop"]
    bins = 2 * hop + 1
    h = d._ensure_handle()
    g = torch.Generator().manual_seed(300 + hop + sum(seqlens))
    rows = sum(seqlens)
    x_pred = torch.randn(rows, 2 * bins, generator=g)
    x_pred[:, :bins] = x_pred[:, :bins] * 1.5 - 1.0
    x_pred[:, bins:] *= 4.0
    x_pred[0, 3] = 7.5  # clip at 100
    ld = (2 * bins + 63) // 64 * 64
    xp = torch.zeros(rows, ld)
    xp[:, :2 * bins] = x_pred
    xd = xp.cuda()
    wav = torch.full((rows * hop,), float("nan"), device="cuda")
    _lib.check(_lib.load().b200codec_istft(h, ctypes.c_void_p(xd.data_ptr()), ld, _lib.i32_array(seqlens), len(seqlens),
                                           ctypes.c_void_p(wav.data_ptr()),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    win = sd["decoder.head.istft.window"]
    off = 0
    for T in seqlens:
        ref = O.istft_same(O.head_spectrum(x_pred[off:off + T][None]), win, hop)[0]
        got = wav[off * hop:(off + T) * hop].cpu()
        assert torch.isfinite(got).all()
        assert (got - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), f"T={T}"
        off += T


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_decode_vs_oracle(name, prec):
    c, sd, d = make(name, prec)
    d = d.to("cuda").eval()
    spt = d.samples_per_token
    ids = torch.randint(0, 65536, (2, 41), generator=torch.Generator().manual_seed(41))
    ref = O.decoder_forward(sd, ids, hop=c["hop"], upsample_factors=c["factors"], kernel_sizes=c["kernels"])
    wav = d(ids.cuda()).cpu()
    assert wav.shape == ref.shape == (2, 1, spt * 41)
    snr_min, rel = TOL[prec]
    snr = O.snr_db(ref, wav)
    maxabs = (ref.double() - wav.double()).abs().max().item()
    print(f"[parity cfg] {name} {prec}: SNR {snr:.1f} dB, max-abs {maxabs:.3e}, peak {ref.abs().max().item():.3e}")
    assert torch.isfinite(wav).all() and snr >= snr_min and maxabs <= rel * ref.abs().max().item()
    # varlen batch == single decodes (ragged lengths around the ISTFT tile sizes)
    lens = [1, 13, 29]
    utts = [torch.randint(0, 65536, (t,), generator=torch.Generator().manual_seed(t)) for t in lens]
    packed = d.decode_packed_host(torch.cat(utts), lens)
    off = 0
    for u in utts:
        single = d.decode_packed_host(u, [u.numel()])
        got = packed[off * spt:(off + u.numel()) * spt]
        assert (got - single).abs().max().item() <= 1e-5 * max(1e-4, single.abs().max().item()), u.numel()
        off += u.numel()
