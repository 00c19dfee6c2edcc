"""Stateful streaming decode: `StreamingDecoder.push` equals the defined semantics
-- the reference forward on cat(context, new), trimmed to the new tokens' samples -- step by step,
with and without the CUDA-graph replay; and the quality against the one-shot decode of the whole
utterance is reported as a function of the context length."""

import pytest
import torch

from oracle import codec_oracle as O


def test_streaming_requires_cuda_decoder():
    from tts_max_b200.codec import decoder, streaming

    dec = decoder.Decoder(16000, 320, None, None, init_seed=0)   # stays on the CPU: no handle is created
    with pytest.raises(RuntimeError):
        streaming.StreamingDecoder(dec, n_streams=1)
    with pytest.raises(ValueError):
        streaming.StreamingDecoder(dec, n_streams=0)


@pytest.mark.parametrize("ctx_len,new", [(20, 10), (25, 10), (5, 10), (0, 7), (30, 30)])
def test_window_slide_logic(ctx_len, new):
    """The device-side token history: after every push the window holds exactly the last
    min(ctx_len, pushed) tokens (right-aligned) followed by the new ones."""
    from tts_max_b200.codec import streaming

    g = torch.Generator().manual_seed(ctx_len * 100 + new)
    ids = torch.randint(0, 65536, (2, new * 9), generator=g)
    window = torch.zeros(2, ctx_len + new, dtype=torch.int64)
    filled = 0
    for k in range(9):
        chunk = ids[:, k * new:(k + 1) * new]
        nxt = streaming.slide_window(window, filled, ctx_len, chunk)
        lo = max(0, k * new - ctx_len)
        want = ids[:, lo:(k + 1) * new]
        assert torch.equal(window[:, ctx_len + new - want.shape[1]:], want), k
        assert filled == min(ctx_len, k * new)
        filled = nxt


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_push_equals_window_decode(gpu_decoders, state_dict, use_graph):
    from tts_max_b200.codec import streaming

    d = gpu_decoders["bf16"]
    n_streams, new, ctx_len, n_push = 3, 10, 20, 7
    g = torch.Generator().manual_seed(17)
    ids = torch.randint(0, 65536, (n_streams, new * n_push), generator=g)
    sd = streaming.StreamingDecoder(d, n_streams, new_tokens=new, left_context=ctx_len, use_graph=use_graph)
    for k in range(n_push):
        out = sd.push(ids[:, k * new:(k + 1) * new])
        assert out.shape == (n_streams, new * 320) and out.is_cuda
        lo = max(0, (k + 1) * new - (ctx_len + new))
        window = ids[:, lo:(k + 1) * new]
        want = d(window.cuda())[:, 0, -new * 320:]
        scale = max(1e-3, want.abs().max().item())
        assert (out - want).abs().max().item() <= 1e-5 * scale, (k, use_graph)
        assert sd.context_tokens == min(ctx_len, (k + 1) * new)
        if k in (0, n_push - 1):   # pin the definition itself on the oracle: first and a steady-state step
            ref = O.decoder_forward(state_dict, window[:1])[:, 0, -new * 320:]
            assert O.snr_db(ref, out[:1].cpu()) >= 40.0
    # a new stream after reset() sees no history
    sd.reset()
    first = sd.push(ids[:, :new])
    want = d(ids[:, :new].cuda())[:, 0, :]
    assert (first - want).abs().max().item() <= 1e-5 * max(1e-3, want.abs().max().item())
    with pytest.raises(ValueError):
        sd.push(ids[:, :new + 1])


@pytest.mark.gpu
def test_graph_survives_reset_and_foreign_decodes(gpu_decoders):
    """The captured CUDA graph bakes the decoder handle's plan and workspace in. reset() is followed by
    warm-up pushes of other shapes, and callers may decode unrelated batches (here: a LARGER one, which
    reallocates the workspace) between steady pushes: every steady push must still equal the eager
    window decode (the graph is re-captured when the handle's plan generation moved)."""
    from tts_max_b200.codec import streaming

    d = gpu_decoders["bf16"]
    n_streams, new, ctx_len = 2, 8, 16
    g = torch.Generator().manual_seed(99)
    ids = torch.randint(0, 65536, (n_streams, new * 12), generator=g)
    other = torch.randint(0, 65536, (5, 211), generator=g).cuda()
    bigger = torch.randint(0, 65536, (9, 1200), generator=g).cuda()
    sd = streaming.StreamingDecoder(d, n_streams, new_tokens=new, left_context=ctx_len, use_graph=True)

    def run(n_push, foreign):
        for k in range(n_push):
            out = sd.push(ids[:, k * new:(k + 1) * new])
            lo = max(0, (k + 1) * new - (ctx_len + new))
            want = d(ids[:, lo:(k + 1) * new].cuda())[:, 0, -new * 320:]
            assert (out - want).abs().max().item() <= 1e-5 * max(1e-3, want.abs().max().item()), k
            if foreign and k % 2 == 1:
                d(other if k % 4 == 1 else bigger)       # another shape on the same decoder

    run(6, foreign=False)     # warm-up, capture, replay
    sd.reset()
    run(8, foreign=False)     # warm-up pushes after reset() invalidate the old graph
    sd.reset()
    run(12, foreign=True)     # unrelated decodes between steady pushes


@pytest.mark.gpu
def test_streaming_quality_vs_one_shot(gpu_decoders):
    """Chunking changes GroupNorm / attention statistics, so streaming != one-shot by construction; more
    context must not make it worse. (Random-init weights: the numbers characterise the method, not a
    trained codec.)"""
    from tts_max_b200.codec import streaming

    d = gpu_decoders["bf16"]
    ids = torch.randint(0, 65536, (2, 500), generator=torch.Generator().manual_seed(23))
    full = d(ids.cuda())[:, 0, :]
    snr = {}
    for ctx_len in (0, 100, 450):
        sd = streaming.StreamingDecoder(d, 2, new_tokens=50, left_context=ctx_len)
        chunks = [sd.push(ids[:, k:k + 50]) for k in range(0, 500, 50)]
        got = torch.cat(chunks, dim=1)
        assert got.shape == full.shape and torch.isfinite(got).all()
        snr[ctx_len] = O.snr_db(full.cpu(), got.cpu())
    print(f"[streaming] SNR vs one-shot decode, 50-token chunks: {snr}")
    assert snr[450] >= snr[100] - 1.0 and snr[100] >= snr[0] - 1.0
    assert snr[450] >= 10.0     # 450 + 50 = the whole utterance from the last chunk on


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_cached_streaming_vs_its_oracle(gpu_decoders, state_dict, prec):
    """The cached streaming path (key / value rings, [overlap | new] rows per push) against the CPU restatement
    of exactly that algorithm (oracle/streaming_oracle.py), push by push: through the warm-up, the ring
    wrap-around (capacity 30 tokens, 9 pushes of 10) and a reset. The first push has no history: it is the plain
    decode of the chunk."""
    from oracle import streaming_oracle
    from tts_max_b200.codec import streaming

    d = gpu_decoders[prec]
    n_streams, new, ctx_len, overlap, n_push = 3, 10, 20, 8, 9
    floor = 40.0 if prec == "bf16" else 55.0
    ids = torch.randint(0, 65536, (n_streams, new * n_push), generator=torch.Generator().manual_seed(31))
    sd = streaming.CachedStreamingDecoder(d, n_streams, new_tokens=new, left_context=ctx_len, overlap=overlap)
    assert sd.capacity == 30
    for round_ in range(2):
        ref = streaming_oracle.CachedStreamOracle(state_dict, n_streams, new, ctx_len)
        for k in range(n_push):
            ov = min(overlap, k * new)
            out = sd.push(ids[:, k * new:(k + 1) * new])
            want = ref.push(ids[:, k * new - ov:(k + 1) * new], ov)
            assert out.shape == (n_streams, new * 320) and out.is_cuda
            snr = O.snr_db(want, out.cpu())
            assert torch.isfinite(out).all() and snr >= floor, (round_, k, snr)
            if k == 0:
                plain = d(ids[:, :new].cuda())[:, 0, :]
                assert (out - plain).abs().max().item() <= 1e-5 * max(1e-3, plain.abs().max().item())
        sd.reset()


@pytest.mark.gpu
def test_cached_streaming_quality_and_cost(gpu_decoders):
    """Quality of the two streaming definitions against the one-shot decode (random-init weights: the numbers
    characterise the methods, not a trained codec) and their cost per push."""
    from tts_max_b200.codec import streaming

    d = gpu_decoders["bf16"]
    ids = torch.randint(0, 65536, (4, 500), generator=torch.Generator().manual_seed(23))
    full = d(ids.cuda())[:, 0, :].cpu()
    snr = {}
    for name, make in (("window L=100", lambda: streaming.StreamingDecoder(d, 4, new_tokens=50, left_context=100, use_graph=False)),
                       ("cached L=100 overlap=8", lambda: streaming.CachedStreamingDecoder(d, 4, 50, 100, overlap=8)),
                       ("cached L=450 overlap=8", lambda: streaming.CachedStreamingDecoder(d, 4, 50, 450, overlap=8)),
                       ("cached L=100 overlap=24", lambda: streaming.CachedStreamingDecoder(d, 4, 50, 100, overlap=24))):
        sdec = make()
        got = torch.cat([sdec.push(ids[:, k:k + 50]) for k in range(0, 500, 50)], dim=1).cpu()
        assert got.shape == full.shape and torch.isfinite(got).all()
        snr[name] = round(O.snr_db(full, got), 2)
    print(f"[streaming] SNR vs one-shot decode, 50-token pushes: {snr}")
    # the cached method must be in the same quality class as the window method it replaces on the fast path
    assert snr["cached L=100 overlap=8"] >= snr["window L=100"] - 6.0


@pytest.mark.gpu
def test_cached_streaming_errors_and_isolation(gpu_decoders):
    """Error paths of the stream ABI are Python exceptions, a state belongs to its decoder, and ordinary decodes
    between pushes do not disturb the rings (they live outside the decode workspace)."""
    import ctypes

    from tts_max_b200 import _lib
    from tts_max_b200.codec import streaming

    d, other = gpu_decoders["bf16"], gpu_decoders["fp16"]
    with pytest.raises(ValueError):
        streaming.CachedStreamingDecoder(d, 0)
    sd = streaming.CachedStreamingDecoder(d, 2, new_tokens=10, left_context=20, overlap=8)
    ids = torch.randint(0, 65536, (2, 60), generator=torch.Generator().manual_seed(3)).cuda()
    with pytest.raises(ValueError):
        sd.push(ids[:, :11])
    lib = _lib.load()
    wav = torch.empty(2, 18 * 320, device="cuda")
    # overlap larger than the tokens pushed so far; a state used with another decoder's handle
    rc = lib.b200codec_stream_push(d._ensure_handle(), sd._state, ctypes.c_void_p(ids.data_ptr()), _lib.IDS_I64, 8,
                                   ctypes.c_void_p(wav.data_ptr()), None)
    assert rc != 0 and b"overlap" in lib.b200codec_last_error()
    rc = lib.b200codec_stream_push(other._ensure_handle(), sd._state, ctypes.c_void_p(ids.data_ptr()), _lib.IDS_I64, 0,
                                   ctypes.c_void_p(wav.data_ptr()), None)
    assert rc != 0 and b"another decoder" in lib.b200codec_last_error()
    ref = streaming.CachedStreamingDecoder(d, 2, new_tokens=10, left_context=20, overlap=8)
    for k in range(6):
        a = sd.push(ids[:, 10 * k:10 * k + 10])
        d(torch.randint(0, 65536, (3, 77 + k), generator=torch.Generator().manual_seed(k)).cuda())   # foreign decode
        b = ref.push(ids[:, 10 * k:10 * k + 10])
        assert (a - b).abs().max().item() <= 1e-5 * max(1e-3, b.abs().max().item()), k
