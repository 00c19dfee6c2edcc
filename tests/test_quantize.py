"""Encode-direction FSQ quantise: oracle vs the reference path's golden vectors
(CPU), and the CUDA kernel vs the oracle through the C ABI (GPU).

Integer parity rule: ids must be identical wherever the bounded value is not within BOUNDARY of
a rounding boundary (k + 0.5); on a boundary the 2048-long fp32 dot product (summation order)
and tanh (1-2 ulp) legitimately decide the digit. Such positions must be rare.
"""

import os

import numpy as np
import pytest
import torch

from oracle import codec_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_quantize_seed0.npz")
BOUNDARY = 2e-4


@pytest.fixture(scope="module")
def qgolden():
    return dict(np.load(GOLDEN))


def digits_of(ids: torch.Tensor) -> torch.Tensor:
    return torch.stack([(ids // 4 ** d) % 4 for d in range(8)], dim=-1)


def off_boundary(bounded: torch.Tensor) -> torch.Tensor:
    """True where every one of a token's eight bounded values is clear of a rounding boundary."""
    frac = (bounded - torch.floor(bounded) - 0.5).abs()
    return (frac > BOUNDARY).all(dim=-1)


# ------------------------------------------------------------------------------------------
# CPU: the oracle against the reference path
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pre", [0, 1])
def test_oracle_matches_reference_quantize(qgolden, state_dict, pre):
    hidden = torch.from_numpy(qgolden["hidden"])
    ids, z, bounded = O.fsq_quantize(state_dict, hidden.permute(0, 2, 1), pre_bound=bool(pre))
    assert torch.equal(ids, torch.from_numpy(qgolden[f"ids_pre{pre}"]))
    assert np.array_equal(bounded.numpy(), qgolden[f"bounded_pre{pre}"])
    assert ids.min() >= 0 and ids.max() < 65536
    # digit d is the rounded bounded value of dimension d (shifted by half_width)
    assert torch.equal(digits_of(ids), (bounded.round() + 2).to(torch.int64))
    # all four levels occur in the fixture
    assert set(digits_of(ids).flatten().tolist()) == {0, 1, 2, 3}


def test_encode_decode_round_trip_oracle(qgolden, state_dict):
    """lookup(quantise(x)) == project_out(codes): what ResidualFSQ.forward returns as quantized_out."""
    hidden = torch.from_numpy(qgolden["hidden"])
    ids, _, bounded = O.fsq_quantize(state_dict, hidden.permute(0, 2, 1))
    out = O.fsq_lookup(state_dict, ids)
    codes = bounded.round() / 2
    ref = torch.nn.functional.linear(codes, state_dict["decoder.quantizer.project_out.weight"],
                                     state_dict["decoder.quantizer.project_out.bias"])
    assert torch.equal(out, ref)


def test_bound_range_and_fixed_points():
    z = torch.linspace(-20, 20, 4001)
    b = O.fsq_bound(z.unsqueeze(-1).expand(-1, 8))[:, 0]
    assert b.min() > -2.01 and b.max() < 1.01          # (-half_l - 0.5, half_l - 0.5)
    assert torch.all(b[1:] >= b[:-1])                  # monotone
    assert abs(float(O.fsq_bound(torch.zeros(1, 8))[0, 0])) < 1e-6   # bound(0) = tanh(shift) * half_l - 0.5 = 0


# ------------------------------------------------------------------------------------------
# GPU: the kernel through the C ABI
# ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("pre", [0, 1])
def test_gpu_quantize_vs_golden(qgolden, gpu_decoders, pre):
    from tts_max_b200.codec import encoder

    dec = gpu_decoders["bf16"]
    hidden = torch.from_numpy(qgolden["hidden"]).cuda()
    code = encoder.FSQQuantizer(dec, pre_bound=bool(pre)).quantize(hidden)
    assert code.shape == (hidden.shape[0], 1, hidden.shape[2]) and code.dtype == torch.int32   # FSQ.codes_to_indices
    ids = code[:, 0, :].cpu().long()
    want = torch.from_numpy(qgolden[f"ids_pre{pre}"])
    clear = off_boundary(torch.from_numpy(qgolden[f"bounded_pre{pre}"]))
    assert clear.float().mean() > 0.99
    assert torch.equal(ids[clear], want[clear])
    assert (ids != want).float().mean() < 0.01
    # the projection itself
    tok = hidden.permute(0, 2, 1).reshape(-1, 2048).contiguous()
    ids2, z = dec.quantize_features(tok, pre_bound=bool(pre), return_projection=True)
    assert torch.equal(ids2.cpu().long(), ids.reshape(-1))
    np.testing.assert_allclose(z.cpu().numpy(), qgolden["z"].reshape(-1, 8), rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_gpu_quantize_round_trip_full_size(gpu_decoders, state_dict):
    """BASELINE config-2 size (16 x 500 tokens): ids are consistent with the kernel's own projection,
    encode -> decode-side lookup reproduces project_out(codes) bit for bit, and a token's id does
    not depend on its neighbours."""
    import ctypes

    from tts_max_b200 import _lib

    dec = gpu_decoders["bf16"]
    g = torch.Generator().manual_seed(99)
    n = 16 * 500
    feats = (torch.randn(n, 2048, generator=g) * 2.0).cuda()
    ids, z = dec.quantize_features(feats, pre_bound=False, return_projection=True, id_dtype=torch.int64)
    assert ids.min() >= 0 and ids.max() < 65536
    bounded = O.fsq_bound(z.cpu())
    clear = off_boundary(bounded)
    assert clear.float().mean() > 0.99
    want_digits = (bounded.round() + 2).to(torch.int64)
    assert torch.equal(digits_of(ids.cpu())[clear], want_digits[clear])
    assert set(digits_of(ids.cpu()).flatten().tolist()) == {0, 1, 2, 3}
    # against the oracle's own projection (different summation order -> only off-boundary tokens)
    o_ids, o_z, o_bounded = O.fsq_quantize(state_dict, feats.cpu())
    both = clear & off_boundary(o_bounded)
    assert torch.equal(ids.cpu()[both], o_ids[both])
    np.testing.assert_allclose(z.cpu().numpy(), o_z.numpy(), rtol=0, atol=5e-5)
    # round trip through K1
    out = torch.empty(n, 2048, device="cuda")
    _lib.check(_lib.load().b200codec_fsq_lookup(dec._ensure_handle(), ctypes.c_void_p(ids.data_ptr()), _lib.IDS_I64, n,
                                                ctypes.c_void_p(out.data_ptr()),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    codes = (digits_of(ids.cpu()).float() - 2) / 2
    ref = torch.nn.functional.linear(codes, state_dict["decoder.quantizer.project_out.weight"],
                                     state_dict["decoder.quantizer.project_out.bias"])
    assert torch.equal(out.cpu(), ref)
    # batch invariance, ragged tail (n not a multiple of the 4 tokens a warp handles)
    for lo, hi in ((0, 1), (5, 8), (123, 130), (n - 3, n)):
        assert torch.equal(dec.quantize_features(feats[lo:hi].clone(), pre_bound=False, id_dtype=torch.int64), ids[lo:hi])


@pytest.mark.gpu
def test_gpu_quantize_rejects_bad_input(gpu_decoders):
    dec = gpu_decoders["bf16"]
    with pytest.raises(ValueError):
        dec.quantize_features(torch.zeros(4, 2048), pre_bound=False)                     # CPU tensor
    with pytest.raises(ValueError):
        dec.quantize_features(torch.zeros(4, 2048, device="cuda", dtype=torch.float16), pre_bound=False)
    assert dec.quantize_features(torch.zeros(0, 2048, device="cuda"), pre_bound=False).numel() == 0
    with pytest.raises(TypeError):
        dec.quantize_features(torch.zeros(4, 2048, device="cuda"))   # pre_bound has no default (parity unpinned)
