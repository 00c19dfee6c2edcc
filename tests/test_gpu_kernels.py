"""GPU: per-stage parity of the sm_100a kernels against the oracle, called through the C ABI."""

import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import codec_oracle as O
from tts_max_b200 import _lib

pytestmark = pytest.mark.gpu

PREC = {"bf16": (0, torch.bfloat16), "fp16": (1, torch.float16)}


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(prec, a, w, taps=1, out_fp32=True, bias=None, residual=None, act=0, ldc=None):
    code, dt = PREC[prec]
    M, Cin = a.shape
    N = w.shape[0]
    n64 = (N + 63) // 64 * 64   # tiles are 64 columns wide at least; ragged N: weight rows past N read as zeros
    ldc = ldc or n64
    out = torch.full((M, ldc), float("nan"), device="cuda", dtype=torch.float32 if out_fp32 else dt)
    _lib.check(_lib.load().b200codec_gemm(code, ptr(a), ptr(w), M, N, Cin, taps, ptr(out), 0 if out_fp32 else 1,
                                          ldc, ptr(bias), ptr(residual), residual.shape[1] if residual is not None else 0,
                                          act, stream()))
    torch.cuda.synchronize()
    return out


# ----------------------------------------------------------------------------------------------
# K1 FSQ lookup: bit-exact
# ----------------------------------------------------------------------------------------------
def test_fsq_lookup_bit_exact_all_codes(gpu_decoders, state_dict, golden):
    d = gpu_decoders["bf16"]
    h = d._ensure_handle()
    lib = _lib.load()
    for dtype, id_type in ((torch.int64, _lib.IDS_I64), (torch.int32, _lib.IDS_I32)):
        ids = torch.arange(65536, dtype=dtype, device="cuda")
        out = torch.empty(65536, 2048, device="cuda")
        _lib.check(lib.b200codec_fsq_lookup(h, ptr(ids), id_type, 65536, ptr(out), stream()))
        torch.cuda.synchronize()
        ref = O.fsq_lookup(state_dict, torch.arange(65536).view(1, -1))[0]
        assert torch.equal(out.cpu(), ref), "FSQ lookup is not bit-exact"
    # the reference's own vectors
    ids = torch.from_numpy(golden["fsq_ids"]).cuda()
    out = torch.empty(ids.numel(), 2048, device="cuda")
    _lib.check(lib.b200codec_fsq_lookup(h, ptr(ids), _lib.IDS_I64, ids.numel(), ptr(out), stream()))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), golden["fsq_out"])


def test_fsq_lookup_bit_exact_random_million(gpu_decoders, state_dict):
    d = gpu_decoders["bf16"]
    h = d._ensure_handle()
    ids = torch.randint(0, 65536, (1 << 20,), generator=torch.Generator().manual_seed(11))
    ref = O.fsq_lookup(state_dict, ids.view(1, -1))[0]
    ids_d = ids.cuda()
    out = torch.empty(ids.numel(), 2048, device="cuda")
    _lib.check(_lib.load().b200codec_fsq_lookup(h, ptr(ids_d), _lib.IDS_I64, ids.numel(), ptr(out), stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)


# ----------------------------------------------------------------------------------------------
# tcgen05 GEMM / implicit conv
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K", [(128, 1024, 1024), (300, 1024, 2048), (1000, 3072, 1024), (77, 4096, 1024),
                                   (513, 1024, 4096), (260, 1282, 1024), (8050, 1024, 1024)])
def test_gemm_linear(prec, M, N, K):
    _, dt = PREC[prec]
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dt)
    out = gemm(prec, a, w)
    ref = a.double() @ w.double().t()
    err = (out[:, :N].double() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max err {err}"
    # 16-bit output path
    out16 = gemm(prec, a, w, out_fp32=False)
    assert (out16[:, :N].double() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_gemm_epilogue_bias_silu_residual(prec):
    _, dt = PREC[prec]
    M, N, K = 391, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dt)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    lin = a.double() @ w.double().t() + bias.double()
    out = gemm(prec, a, w, bias=bias)
    assert (out.double() - lin).abs().max().item() <= 3e-3 * lin.abs().max().item()
    out = gemm(prec, a, w, bias=bias, act=1)
    assert (out.double() - F.silu(lin)).abs().max().item() <= 3e-3 * lin.abs().max().item()
    out = gemm(prec, a, w, bias=bias, act=1, residual=res)
    assert (out.double() - (F.silu(lin) + res.double())).abs().max().item() <= 3e-3 * lin.abs().max().item()


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("taps,M", [(3, 500), (7, 250), (3, 129), (7, 1)])
def test_gemm_implicit_conv(prec, taps, M):
    """Conv1d(k, padding='same') over one utterance of M frames == taps row-shifted K-slabs."""
    _, dt = PREC[prec]
    C = 1024
    g = torch.Generator(device="cuda").manual_seed(taps * 100 + M)
    x = torch.randn(M, C, device="cuda", generator=g).to(dt)                    # token-major
    w = (torch.randn(C, C, taps, device="cuda", generator=g) * 0.02).to(dt)     # torch Conv1d layout
    bias = torch.randn(C, device="cuda", generator=g) * 0.1
    w_packed = w.permute(0, 2, 1).reshape(C, taps * C).contiguous()             # [Cout, tap*Cin + c]
    out = gemm(prec, x, w_packed, taps=taps, bias=bias)
    ref = F.conv1d(x.double().t().unsqueeze(0), w.double(), bias.double(), padding=taps // 2)[0].t()
    err = (out.double() - ref).abs().max().item()
    assert err <= 3e-3 * max(1.0, ref.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("M,N,K,taps", [(1, 1024, 1024, 1), (250, 1024, 4096, 1), (253, 3072, 1024, 1),
                                         (500, 1024, 1024, 3), (1009, 1024, 1024, 1)])
def test_gemm_narrow_tiles_are_invisible(M, N, K, taps):
    """Small M runs on 256 x 64 tiles; every output element sees the same K order, so the result is
    bit-identical to the 256-wide tiling (fp32 + residual and 16-bit + SiLU epilogues)."""
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, taps * K, device="cuda", generator=g) / (taps * K) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    try:
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(0))
        wide32 = gemm("bf16", a, w, taps=taps, bias=bias, residual=res)
        wide16 = gemm("bf16", a, w, taps=taps, out_fp32=False, bias=bias, act=1)
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(1))
        narrow32 = gemm("bf16", a, w, taps=taps, bias=bias, residual=res)
        narrow16 = gemm("bf16", a, w, taps=taps, out_fp32=False, bias=bias, act=1)
    finally:
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(1))
    assert torch.isfinite(narrow32).all()
    assert torch.equal(wide32, narrow32) and torch.equal(wide16, narrow16)


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K,taps", [(1000, 128, 128, 7), (517, 192, 192, 7), (2300, 384, 384, 7), (300, 384, 768, 3),
                                         (700, 128, 64, 1), (40000, 192, 192, 1)])
def test_gemm_encoder_tile_widths(prec, M, N, K, taps):
    """N % 256 != 0 (the encoder's 128 / 192 / 384-channel convs) runs on 256 x 128 / 256 x 192 tiles; the
    result matches the fp64 product and is bit-identical to the 256 x 64 tiling (same K order per element)."""
    lib = _lib.load()
    _, dt = PREC[prec]
    g = torch.Generator(device="cuda").manual_seed(M + N + K + taps)
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, taps * K, device="cuda", generator=g) / (taps * K) ** 0.5).to(dt)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    try:
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(2))
        n32 = gemm(prec, a, w, taps=taps, bias=bias, residual=res)
        n16 = gemm(prec, a, w, taps=taps, out_fp32=False, bias=bias, act=1)
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(1))
        w32 = gemm(prec, a, w, taps=taps, bias=bias, residual=res)
        w16 = gemm(prec, a, w, taps=taps, out_fp32=False, bias=bias, act=1)
    finally:
        _lib.check(lib.b200codec_set_gemm_narrow_tiles(1))
    ap = F.pad(a.double(), (0, 0, taps // 2, taps // 2))
    cols = torch.cat([ap[t:t + M] for t in range(taps)], dim=1)  # [M, taps * K], tap-major like the weight
    ref = cols @ w.double().t() + bias.double() + res.double()
    assert (w32.double() - ref).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item())
    assert torch.equal(n32, w32) and torch.equal(n16, w16)


# ----------------------------------------------------------------------------------------------
# norms
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_rmsnorm_layernorm(prec):
    code, dt = PREC[prec]
    g = torch.Generator(device="cuda").manual_seed(1)
    rows = 1037
    x = torch.randn(rows, 1024, device="cuda", generator=g) * 3 + 0.5
    w = torch.randn(1024, device="cuda", generator=g)
    b = torch.randn(1024, device="cuda", generator=g)
    out = torch.empty(rows, 1024, device="cuda", dtype=dt)
    lib = _lib.load()
    _lib.check(lib.b200codec_rmsnorm(code, ptr(x), ptr(w), rows, 1024, 1e-6, ptr(out), stream()))
    torch.cuda.synchronize()
    ref = O.rms_norm(x.cpu(), w.cpu())
    tol = 1e-2 if prec == "bf16" else 2e-3
    assert (out.float().cpu() - ref).abs().max().item() <= tol * ref.abs().max().item()
    _lib.check(lib.b200codec_layernorm(code, ptr(x), ptr(w), ptr(b), rows, 1024, 1e-6, ptr(out), stream()))
    torch.cuda.synchronize()
    ref = F.layer_norm(x.cpu(), (1024,), w.cpu(), b.cpu(), 1e-6)
    assert (out.float().cpu() - ref).abs().max().item() <= tol * ref.abs().max().item()


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_groupnorm_swish_varlen(prec):
    code, dt = PREC[prec]
    g = torch.Generator(device="cuda").manual_seed(2)
    seqlens = [5, 130, 1, 64]
    rows = sum(seqlens)
    x = torch.randn(rows, 1024, device="cuda", generator=g) * 2 + 1.0
    gamma = torch.randn(1024, device="cuda", generator=g)
    beta = torch.randn(1024, device="cuda", generator=g)
    out = torch.empty(rows, 1024, device="cuda", dtype=dt)
    _lib.check(_lib.load().b200codec_groupnorm_swish(code, ptr(x), ptr(gamma), ptr(beta), _lib.i32_array(seqlens),
                                                     len(seqlens), 1024, 1e-6, ptr(out), stream()))
    torch.cuda.synchronize()
    off = 0
    tol = 1e-2 if prec == "bf16" else 2e-3
    for T in seqlens:
        xu = x[off:off + T].cpu().t().unsqueeze(0)  # (1, C, T): statistics over 32 ch x T per utterance
        ref = O.swish(F.group_norm(xu, 32, gamma.cpu(), beta.cpu(), 1e-6))[0].t()
        got = out[off:off + T].float().cpu()
        assert (got - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
        off += T


# ----------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("seqlens", [[64], [1], [37, 250], [65, 128, 500, 3], [129, 127, 256, 257]])
def test_attention_varlen(prec, seqlens):
    code, dt = PREC[prec]
    g = torch.Generator(device="cuda").manual_seed(sum(seqlens))
    rows = sum(seqlens)
    qkv = (torch.randn(rows, 3072, device="cuda", generator=g) * 1.5).to(dt)
    out = torch.full((rows, 1024), float("nan"), device="cuda", dtype=dt)
    _lib.check(_lib.load().b200codec_attention(code, ptr(qkv), _lib.i32_array(seqlens), len(seqlens), 16, ptr(out),
                                               stream()))
    torch.cuda.synchronize()
    off = 0
    for T in seqlens:
        blk = qkv[off:off + T].float().cpu().view(T, 3, 16, 64).permute(1, 2, 0, 3)  # r h t d
        ref = F.scaled_dot_product_attention(blk[0][None], blk[1][None], blk[2][None])[0]  # h t d
        ref = ref.permute(1, 0, 2).reshape(T, 1024)
        got = out[off:off + T].float().cpu()
        assert torch.isfinite(got).all()
        tol = 2e-2 if prec == "bf16" else 4e-3
        assert (got - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
        off += T


def test_attention_long_form():
    code, dt = PREC["bf16"]
    T = 3000
    g = torch.Generator(device="cuda").manual_seed(9)
    qkv = torch.randn(T, 3072, device="cuda", generator=g).to(dt)
    out = torch.empty(T, 1024, device="cuda", dtype=dt)
    _lib.check(_lib.load().b200codec_attention(code, ptr(qkv), _lib.i32_array([T]), 1, 16, ptr(out), stream()))
    torch.cuda.synchronize()
    blk = qkv.float().view(T, 3, 16, 64).permute(1, 2, 0, 3)
    ref = F.scaled_dot_product_attention(blk[0][None], blk[1][None], blk[2][None])[0].permute(1, 0, 2).reshape(T, 1024)
    assert (out.float() - ref).abs().max().item() <= 2e-2


# ----------------------------------------------------------------------------------------------
# K13 + K14 ISTFT
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seqlens", [[1], [2], [3], [12], [13], [37], [250, 5, 24], [500]])
def test_istft_vs_oracle(gpu_decoders, state_dict, seqlens):
    d = gpu_decoders["bf16"]
    h = d._ensure_handle()
    g = torch.Generator().manual_seed(100 + sum(seqlens))
    rows = sum(seqlens)
    x_pred = torch.randn(rows, 1282, generator=g)
    x_pred[:, :641] = x_pred[:, :641] * 1.5 - 1.0
    x_pred[:, 641:] *= 4.0
    x_pred[0, 5] = 7.5  # clip at 100
    ld = 1344
    xp = torch.zeros(rows, ld)
    xp[:, :1282] = x_pred
    xd = xp.cuda()
    wav = torch.full((rows * 320,), float("nan"), device="cuda")
    _lib.check(_lib.load().b200codec_istft(h, ptr(xd), ld, _lib.i32_array(seqlens), len(seqlens), ptr(wav), stream()))
    torch.cuda.synchronize()
    win = state_dict["decoder.head.istft.window"]
    off = 0
    for T in seqlens:
        ref = O.istft_same(O.head_spectrum(x_pred[off:off + T][None]), win, 320)[0]
        got = wav[off * 320:(off + T) * 320].cpu()
        assert torch.isfinite(got).all()
        scale = max(1.0, ref.abs().max().item())
        assert (got - ref).abs().max().item() <= 2e-5 * scale, f"T={T}"
        off += T
