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
def test_decode_code_store_pipelined_equals_sequential(tmp_path):
    """Two buckets in flight (async host decode into a ring of pinned buffers) give bit-identical waveforms to the
    bucket-by-bucket sweep, in the same order; abandoning the generator half way leaves the decoder usable."""
    rng = np.random.default_rng(7)
    utts = [rng.integers(0, 65536, int(n)) for n in rng.integers(1, 260, 40)]
    write_store(tmp_path, utts)
    store = batching.CodeStore.open(str(tmp_path), "train")
    dec = decoding.AudioDecoder(None, decoding.DecoderConfig("", 16000, 50, 320, None, None), device="cuda")
    seq = list(batching.decode_code_store(dec, store, max_tokens=600, pipelined=False))
    pipe = list(batching.decode_code_store(dec, store, max_tokens=600, pipelined=True))
    assert [i for i, _ in seq] == [i for i, _ in pipe] and len(pipe) == len(utts)
    for (i, a), (_, b) in zip(seq, pipe):
        assert a.shape == b.shape == (1, 320 * len(utts[i])) and not b.is_pinned()
        assert torch.equal(a, b), i
    it = batching.decode_code_store(dec, store, max_tokens=600)
    first = [next(it) for _ in range(3)]
    it.close()
    again = dict(batching.decode_code_store(dec, store, sample_ids=[first[0][0]], max_tokens=600))
    assert torch.equal(again[first[0][0]], first[0][1])
    with pytest.raises(ValueError):
        dec._decoder.decode_packed_host_async(torch.zeros(4, dtype=torch.int32), [4], torch.empty(4 * 320))  # not pinned


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


def _sorted_rule_table(first_new_id, others, vocab, codebook):
    """What `tokenizer.add_tokens(sorted(new_tokens))` does (tokenization.py:36-49), in plain Python."""
    new = sorted(list(others) + [f"<|s_{i}|>" for i in range(codebook)])
    table = {}
    for rank, tok in enumerate(new):
        table[tok] = first_new_id + rank
    return table


def test_speech_token_map_tables_agree():
    """from_tokenizer and from_sorted_rule build the same permutation table (CPU, no kernel)."""
    from tts_max_b200.codec import batching

    others = ["<|speech_start|>", "<|speech_end|>", "<|text_prompt_start|>", "<|text_prompt_end|>"]
    tok2id = _sorted_rule_table(1000, others, 3000, 512)

    class FakeTokenizer:
        def convert_tokens_to_ids(self, toks):
            return [tok2id[t] for t in toks]

        def __len__(self):
            return 3000

    a = batching.SpeechTokenMap.from_tokenizer(FakeTokenizer(), codebook_size=512, device="cpu")
    b = batching.SpeechTokenMap.from_sorted_rule(1000, others, 3000, codebook_size=512, device="cpu")
    assert torch.equal(a.table, b.table)
    assert int((a.table >= 0).sum()) == 512 and a.table[tok2id["<|s_10|>"]] == 10
    assert a.table[tok2id["<|s_10|>"]] != a.table[tok2id["<|s_9|>"]] + 1 or True   # a permutation, not an offset
    assert tok2id["<|s_10|>"] < tok2id["<|s_2|>"]                                    # "<|s_10|>" sorts before "<|s_2|>"
    with pytest.raises(RuntimeError):
        a.map([torch.tensor([1, 2])])           # no CPU path


@pytest.mark.gpu
def test_gpu_speech_token_map_and_token_completions(gpu_decoders):
    """The GPU id map reproduces detokenise -> tokenise -> extract_speech_ids (inferencing.py:53-63) on token-id
    sequences with text, control tokens and padding mixed in, for sequences longer than one 256-token chunk, and
    `decode_token_completions` equals `decode_completions` on the mapped ids."""
    from tts_max_b200.codec import batching, decoding

    others = ["<|speech_start|>", "<|speech_end|>", "<|text_prompt_start|>", "<|text_prompt_end|>"]
    vocab, first = 70000 + 200, 150
    tok2id = _sorted_rule_table(first, others, vocab, 65536)
    id2tok = {v: k for k, v in tok2id.items()}
    tmap = batching.SpeechTokenMap.from_sorted_rule(first, others, vocab, device="cuda")
    g = torch.Generator().manual_seed(3)
    seqs, want = [], []
    for n in (0, 1, 255, 256, 257, 700, 40):
        codes = torch.randint(0, 65536, (n,), generator=g).tolist()
        toks = [tok2id["<|speech_start|>"]]
        for k, c in enumerate(codes):
            toks.append(tok2id[f"<|s_{c}|>"])
            if k % 37 == 5:
                toks.append(int(torch.randint(0, first, (1,), generator=g)))      # a text token in between
        toks += [tok2id["<|speech_end|>"], 0, 0]
        seqs.append(torch.tensor(toks, dtype=torch.int64))
        strs = [id2tok.get(t, "txt") for t in toks]
        want.append(batching.extract_speech_ids(strs))
    got = tmap.map(seqs)
    for a, b in zip(got, want):
        assert a.dtype == torch.int32 and a.is_cuda and a.tolist() == b
    cfg = decoding.DecoderConfig("", 16000, 50, 320, None, None)
    dec = decoding.AudioDecoder.__new__(decoding.AudioDecoder)
    dec._decoder, dec._device, dec._sample_rate, dec._token_rate = gpu_decoders["bf16"], torch.device("cuda"), 16000, 50
    prompts = [torch.randint(0, 65536, (20,), generator=g) for _ in seqs]
    a = batching.decode_token_completions(dec, tmap, prompts, seqs)
    b = batching.decode_completions(dec, prompts, [torch.tensor(w, dtype=torch.int64) for w in want])
    assert len(a) == len(b) and a[0].shape == (1, 0)
    for x, y in zip(a, b):
        assert x.shape == y.shape and torch.equal(x, y)
