"""Caller-side batching for the decode path.

The reference's two high-volume consumers of `AudioDecoder.decode` feed it one utterance at a time:
  * RLHF rewards: `RewardFunc._decode_audio` (tts/training/rlhf/rewards.py:67-98) concatenates prompt
    and generated ids, decodes, and strips the prompt's samples -- in a Python loop over
    `num_generations x batch` completions;
  * dataset tooling: codes live on disk as one flat int32 array plus an offsets array
    (`{split}_codes.npy` / `{split}_codes_index.npy`, tools/data/data_vectorizer.py:122-146, read back
    by tts/data/data_utils.py:98-152), which already IS a packed varlen layout.
Both map onto one varlen launch sequence per bucket of <= max_tokens tokens; every waveform equals the
single-utterance `decode` of the same ids.
"""

from __future__ import annotations

import ctypes
import logging
import os
from typing import Iterable, Iterator, Sequence

import numpy as np
import torch

from tts_max_b200 import _lib, sharding
from tts_max_b200.codec.decoding import AudioDecoder

_LOG = logging.getLogger(__name__)


def extract_speech_ids(speech_tokens_str: Sequence[str]) -> list[int]:
    """`<|s_N|>` strings -> N (mirror of tts/inference/inferencing.py:53-63; unexpected tokens are
    skipped there with an error log, and skipped here)."""
    speech_ids = []
    for token_str in speech_tokens_str:
        if token_str.startswith("<|s_") and token_str.endswith("|>"):
            speech_ids.append(int(token_str[4:-2]))
    return speech_ids


SPEECH_TOKEN_PATTERN = "<|s_{}|>"          # tts/core/constants.py:5


class SpeechTokenMap:
    """LLM vocabulary id -> FSQ code id, applied on the GPU (`b200codec_map_speech_tokens`).

    The reference turns a completion into speech ids by detokenising it, tokenising the string again and parsing
    `<|s_N|>` (rewards.py:70-73, inferencing.py:53-63). The speech tokens are added to the tokenizer with
    `tokenizer.add_tokens(sorted(new_tokens))` (tokenization.py:36-49) -- lexicographic order -- so id -> N is a
    permutation of a contiguous id range, not an offset; this class holds it as a device table
    (`table[v] = N`, -1 for every other token)."""

    def __init__(self, table: torch.Tensor):
        if table.dtype != torch.int32 or table.dim() != 1:
            raise ValueError("table must be a 1-D int32 tensor indexed by vocabulary id")
        self.table = table
        self.vocab = int(table.numel())

    @staticmethod
    def from_tokenizer(tokenizer, codebook_size: int = 65536, device: torch.device | str = "cuda") -> "SpeechTokenMap":
        """`tokenizer`: the extended tokenizer of tts/core/tokenization.py (anything with `convert_tokens_to_ids`
        and `__len__`)."""
        ids = tokenizer.convert_tokens_to_ids([SPEECH_TOKEN_PATTERN.format(i) for i in range(codebook_size)])
        table = torch.full((len(tokenizer),), -1, dtype=torch.int32)
        table[torch.tensor(ids, dtype=torch.int64)] = torch.arange(codebook_size, dtype=torch.int32)
        return SpeechTokenMap(table.to(device))

    @staticmethod
    def from_sorted_rule(first_new_id: int, other_new_tokens: Sequence[str], vocab_size: int, codebook_size: int = 65536,
                         device: torch.device | str = "cuda") -> "SpeechTokenMap":
        """The same table without a tokenizer object: the new tokens (`other_new_tokens` + the speech tokens) take
        the ids `first_new_id ...` in sorted order (tokenization.py:36-49)."""
        new_tokens = sorted(list(other_new_tokens) + [SPEECH_TOKEN_PATTERN.format(i) for i in range(codebook_size)])
        table = torch.full((vocab_size,), -1, dtype=torch.int32)
        for rank, tok in enumerate(new_tokens):
            if tok.startswith("<|s_") and tok.endswith("|>") and tok[4:-2].isdigit():
                table[first_new_id + rank] = int(tok[4:-2])
        return SpeechTokenMap(table.to(device))

    @torch.no_grad()
    def map(self, token_ids: Sequence[torch.Tensor]) -> list[torch.Tensor]:
        """token id sequences (1-D integer tensors, any device) -> int32 code-id tensors on the table's device:
        the speech tokens of each sequence, in order; everything else (text, `<|speech_end|>`, padding) dropped."""
        if len(token_ids) == 0:
            return []
        dev = self.table.device
        if dev.type != "cuda":
            raise RuntimeError("SpeechTokenMap.map runs on a CUDA device (there is no CPU path)")
        lens = [int(t.numel()) for t in token_ids]
        off = torch.zeros(len(lens) + 1, dtype=torch.int32)
        off[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0).to(torch.int32)
        packed = torch.cat([t.reshape(-1).to(dev, torch.int64) for t in token_ids]) if sum(lens) else torch.zeros(0, dtype=torch.int64, device=dev)
        codes = torch.empty(max(sum(lens), 1), dtype=torch.int32, device=dev)
        out_len = torch.zeros(len(lens), dtype=torch.int32, device=dev)
        off_dev = off.to(dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.load().b200codec_map_speech_tokens(
                ctypes.c_void_p(self.table.data_ptr()), self.vocab, ctypes.c_void_p(packed.data_ptr()),
                ctypes.c_void_p(off_dev.data_ptr()), len(lens), ctypes.c_void_p(codes.data_ptr()),
                ctypes.c_void_p(out_len.data_ptr()), ctypes.c_void_p(stream)))
        n_out = out_len.tolist()
        starts = off.tolist()
        return [codes[starts[i]:starts[i] + n_out[i]] for i in range(len(lens))]


def decode_token_completions(audio_decoder: AudioDecoder, token_map: SpeechTokenMap,
                             prompt_speech_ids: Sequence[torch.Tensor], completion_token_ids: Sequence[torch.Tensor],
                             max_tokens: int = 16384) -> list[torch.Tensor]:
    """`decode_completions` fed with the LLM's completion TOKEN IDS (what generation returns) instead of speech
    ids: the vocabulary-id -> code-id map and the dropping of non-speech tokens run on the GPU."""
    return decode_completions(audio_decoder, prompt_speech_ids, token_map.map(completion_token_ids), max_tokens=max_tokens)


def decode_completions(
    audio_decoder: AudioDecoder,
    prompt_speech_ids: Sequence[torch.Tensor],
    generated_speech_ids: Sequence[torch.Tensor],
    max_tokens: int = 16384,
) -> list[torch.Tensor]:
    """Batched `RewardFunc._decode_audio`: for every (prompt, completion) pair decode
    `cat([prompt, completion])` and drop the first `int(len(prompt) / token_rate * sample_rate)` samples
    (rewards.py:84-93). An empty completion yields `zeros((1, 0))` like rewards.py:76-82. Results are
    float32 CPU tensors of shape (1, L), in input order."""
    if len(prompt_speech_ids) != len(generated_speech_ids):
        raise ValueError("prompt_speech_ids and generated_speech_ids must have the same length")
    out: list[torch.Tensor | None] = [None] * len(generated_speech_ids)
    todo, utts = [], []
    for i, (p, g) in enumerate(zip(prompt_speech_ids, generated_speech_ids)):
        if g.numel() == 0:
            out[i] = torch.zeros((1, 0))
            continue
        todo.append(i)
        utts.append(torch.cat([p.detach().to("cpu", torch.int64).reshape(-1), g.detach().to("cpu", torch.int64).reshape(-1)]))
    lengths = [int(u.numel()) for u in utts]
    # One bad completion must not fail the others: `_decode_audio` wraps each decode in try/except and
    # returns zeros((1, 0)) for the one that failed (rewards.py:86-97). Ids are validated per utterance
    # BEFORE packing, so a bucket never aborts on someone else's out-of-range id.
    bad = {k for k, u in enumerate(utts) if int(u.min()) < 0 or int(u.max()) > 65535}
    for k in bad:
        _LOG.error("Error decoding audio: speech id outside [0, 65535] in completion %d", todo[k])
        out[todo[k]] = torch.zeros((1, 0))
    good = [k for k in range(len(utts)) if k not in bad]
    for bucket in sharding.bucket_by_length(good, lengths, max_tokens=max_tokens):
        try:
            wavs = audio_decoder.decode_batch([utts[k] for k in bucket])
        except Exception as e:  # the reference catches everything per completion
            _LOG.error("Error decoding a bucket of %d completions (%s); retrying one by one", len(bucket), e)
            wavs = []
            for k in bucket:
                try:
                    wavs.append(audio_decoder.decode_batch([utts[k]])[0])
                except Exception as e1:
                    _LOG.error("Error decoding audio: %s", e1)
                    wavs.append(None)
        for k, wav in zip(bucket, wavs):
            if wav is None:
                out[todo[k]] = torch.zeros((1, 0))
                continue
            i = todo[k]
            prompt_wav_length = int(prompt_speech_ids[i].numel() / audio_decoder.token_rate * audio_decoder.sample_rate)
            out[i] = wav[:, prompt_wav_length:]
    return out  # type: ignore[return-value]


def decode_stream_windows(
    audio_decoder: AudioDecoder, windows: Sequence[torch.Tensor], new_tokens: Sequence[int] | int
) -> list[torch.Tensor]:
    """Chunked / streaming decode (BASELINE config 5): every element of `windows` is
    `cat([left_context_ids, new_ids])`; all windows are decoded in one varlen batch and only the samples
    of the trailing `new_tokens` tokens are returned. The reference has no streaming decode and chunking
    changes GroupNorm / attention statistics, so the defined result is "the reference
    forward on exactly this window, trimmed" -- which is what this returns."""
    if isinstance(new_tokens, int):
        new_tokens = [new_tokens] * len(windows)
    if len(new_tokens) != len(windows):
        raise ValueError("new_tokens must match windows")
    hop = audio_decoder._decoder.samples_per_token
    out = []
    for wav, win, n_new in zip(audio_decoder.decode_batch(list(windows)), windows, new_tokens):
        if not 0 < int(n_new) <= win.numel():
            raise ValueError("new_tokens must be in (0, len(window)]")
        out.append(wav[:, wav.shape[1] - int(n_new) * hop:])
    return out


class CodeStore:
    """Read-only view of the vectorizer's on-disk codes: flat int32 `codes` (memmap) + `index` offsets."""

    def __init__(self, codes: np.ndarray, index: np.ndarray):
        if codes.dtype != np.int32 or codes.ndim != 1:
            raise ValueError("codes must be a flat int32 array")
        self.codes = codes
        self.index = np.asarray(index, dtype=np.int64)

    @staticmethod
    def open(dataset_dir: str, split: str) -> "CodeStore":
        """`{split}_codes.npy` is a raw int32 memmap (no .npy header), `{split}_codes_index.npy` a real
        .npy file -- exactly as tools/data/data_vectorizer.py:122-146 writes them."""
        codes = np.memmap(os.path.join(dataset_dir, f"{split}_codes.npy"), dtype=np.int32, mode="r")
        index = np.load(os.path.join(dataset_dir, f"{split}_codes_index.npy"))
        return CodeStore(codes, index)

    def __len__(self) -> int:
        return int(self.index.shape[0])

    def span(self, i: int) -> tuple[int, int]:
        """[left, right) of sample i (tts/data/data_utils.py:143-147)."""
        left = int(self.index[i])
        right = int(self.index[i + 1]) if i < len(self.index) - 1 else int(self.codes.shape[0])
        return left, right

    def length(self, i: int) -> int:
        left, right = self.span(i)
        return right - left


def decode_code_store(
    audio_decoder: AudioDecoder, store: CodeStore, sample_ids: Iterable[int] | None = None, max_tokens: int = 16384,
    pipelined: bool = True,
) -> Iterator[tuple[int, torch.Tensor]]:
    """Decodes samples of a CodeStore in length-sorted varlen buckets; yields (sample id, (1, L) float32
    CPU waveform). The int32 codes go to the GPU as stored (no int64 widening). `pipelined`: two buckets are
    kept in flight (two pinned id / PCM buffers), so packing bucket k + 1 and copying bucket k's waveforms out of
    the pinned buffer overlap the GPU's work; results are the same either way."""
    ids = list(range(len(store))) if sample_ids is None else [int(i) for i in sample_ids]
    lengths = {i: store.length(i) for i in ids}
    for i in ids:
        if lengths[i] <= 0:
            raise ValueError(f"sample {i} has no codes")
    dec = audio_decoder._decoder
    hop = dec.samples_per_token
    buckets = sharding.bucket_by_length(ids, lengths, max_tokens=max_tokens)
    if not pipelined:
        for bucket in buckets:
            seqlens = [lengths[i] for i in bucket]
            packed = np.concatenate([store.codes[slice(*store.span(i))] for i in bucket]).astype(np.int32, copy=False)
            wav = dec.decode_packed_host(torch.from_numpy(np.ascontiguousarray(packed)), seqlens)
            off = 0
            for i, n in zip(bucket, seqlens):
                yield i, wav[off * hop:(off + n) * hop].view(1, -1)
                off += n
        return
    ring_ids: list[torch.Tensor | None] = [None, None]
    ring_wav: list[torch.Tensor | None] = [None, None]

    def finish(pending):
        bucket, seqlens, wav, event = pending
        event.synchronize()
        off = 0
        for i, n in zip(bucket, seqlens):
            yield i, wav[off * hop:(off + n) * hop].clone().view(1, -1)   # pageable copy: the pinned buffer is reused
            off += n

    pending = None
    k = 0
    for bucket in buckets:
        seqlens = [lengths[i] for i in bucket]
        total = sum(seqlens)
        if ring_ids[k] is None or ring_ids[k].numel() < total:
            ring_ids[k] = torch.empty(total + total // 4, dtype=torch.int32).pin_memory()
            ring_wav[k] = torch.empty((total + total // 4) * hop, dtype=torch.float32).pin_memory()
        ids_k, off = ring_ids[k][:total], 0
        ids_np = ids_k.numpy()
        for i, n in zip(bucket, seqlens):
            ids_np[off:off + n] = store.codes[slice(*store.span(i))]
            off += n
        wav_k = ring_wav[k][:total * hop]
        event = dec.decode_packed_host_async(ids_k, seqlens, wav_k)
        if pending is not None:
            yield from finish(pending)
        pending = (bucket, seqlens, wav_k, event)
        k ^= 1
    if pending is not None:
        yield from finish(pending)
