"""Caller-side batching: code-store reader on the CPU, batched decode on the GPU."""

import numpy as np
import pytest
import torch

from tts_max_b200.codec import batching, decoding


def write_store(tmp_path, utts, split="train"):
    """Writes codes / index exactly like tools/data/data_vectorizer.py:122-146 (raw int32 memmap + .npy index)."""
    codes = np.concatenate(utts).astype(np.int32)
    index = np.cumsum([0] + [len(u) for u in utts[:-1]]).astype(np.int64)
    arr = np.memmap(tmp_path / f"{split}_codes.npy", dtype=np.int32, mode="w+", shape=(codes.shape[0],))
    arr[:] = codes
    arr.flush()
    np.save(tmp_path / f"{split}_codes_index.npy", index)


def test_code_store_spans(tmp_path):
    rng = np.random.default_rng(0)
    utts = [rng.integers(0, 65536, n) for n in (5, 1, 12, 7)]
    write_store(tmp_path, utts)
    store = batching.CodeStore.open(str(tmp_path), "train")
    assert len(store) == 4
    assert [store.length(i) for i in range(4)] == [5, 1, 12, 7]
    for i, u in enumerate(utts):
        left, right = store.span(i)
        assert np.array_equal(store.codes[left:right], u.astype(np.int32))


def test_extract_speech_ids_mirror():
    assert batching.extract_speech_ids(["<|s_0|>", "<|s_65535|>", "<|text|>", "<|s_12|>"]) == [0, 65535, 12]


@pytest.mark.gpu
def test_decode_code_store_equals_single_decodes(tmp_path):
    rng = np.random.default_rng(1)
    utts = [rng.integers(0, 65536, n) for n in (40, 3, 130, 64)]
    write_store(tmp_path, utts)
    store = batching.CodeStore.open(str(tmp_path), "train")
    dec = decoding.AudioDecoder(None, decoding.DecoderConfig("", 16000, 50, 320, None, None), device="cuda")
    got = dict(batching.decode_code_store(dec, store, max_tokens=150))  # forces several buckets
    assert sorted(got) == [0, 1, 2, 3]
    for i, u in enumerate(utts):
        single = dec.decode(torch.from_numpy(u.astype(np.int64)))
        assert got[i].shape == single.shape == (1, 320 * len(u))
        assert (got[i] - single).abs().max().item() <= 1e-5 * max(1e-3, single.abs().max().item())


@pytest.mark.gpu
def test_decode_completions_matches_reward_loop():
    g = torch.Generator().manual_seed(3)
    dec = decoding.AudioDecoder(None, decoding.DecoderConfig("", 16000, 50, 320, None, None), device="cuda")
    prompts = [torch.randint(0, 65536, (n,), generator=g) for n in (20, 20, 35, 10)]
    gens = [torch.randint(0, 65536, (n,), generator=g) for n in (50, 0, 7, 120)]
    out = batching.decode_completions(dec, prompts, gens)
    for p, gen, wav in zip(prompts, gens, out):
        if gen.numel() == 0:
            assert wav.shape == (1, 0)  # rewards.py:76-82
            continue
        ref = dec.decode(torch.cat([p, gen]))  # rewards.py:84-93
        ref = ref[:, int(len(p) / dec.token_rate * dec.sample_rate):]
        assert wav.shape == ref.shape == (1, 320 * gen.numel())
        assert (wav - ref).abs().max().item() <= 1e-5 * max(1e-3, ref.abs().max().item())


@pytest.mark.gpu
def test_decode_stream_windows_trims_to_new_audio():
    g = torch.Generator().manual_seed(4)
    dec = decoding.AudioDecoder(None, decoding.DecoderConfig("", 16000, 50, 320, None, None), device="cuda")
    windows = [torch.randint(0, 65536, (n,), generator=g) for n in (150, 150, 80)]
    out = batching.decode_stream_windows(dec, windows, [50, 50, 30])
    for win, n_new, wav in zip(windows, (50, 50, 30), out):
        ref = dec.decode(win)[:, -n_new * 320:]  # the reference forward on the same window, trimmed
        assert wav.shape == ref.shape == (1, n_new * 320)
        assert (wav - ref).abs().max().item() <= 1e-5 * max(1e-3, ref.abs().max().item())
    with pytest.raises(ValueError):
        batching.decode_stream_windows(dec, windows, [50, 50, 500])
