"""Stateful chunked decode for serving (BASELINE config 5's shape).

The reference has no streaming decoder (`tools/serving/inference.py:155-170` is one-shot), and
chunking changes what the model computes: GroupNorm statistics and the unmasked attention see the
window, not the utterance. The semantics are therefore DEFINED as

    audio(new tokens) = last `new * samples_per_token` samples of  Decoder.forward(cat(context, new))

with `context` = the stream's previous `left_context` tokens (fewer at the start of a stream). That is
exactly `batching.decode_stream_windows` on the window the stream has reached, so the parity oracle is
the reference forward on the same window, trimmed -- and the quality against the one-shot decode of the
whole utterance is a property of `left_context`, measured in tests/test_streaming.py.

What this class adds is the state (per-stream token history on the device) and, once every stream has a
full context, a FIXED window shape: the launch chain of one step is then captured once in a CUDA graph
and replayed (tools/graph_ab.py: about 5 % at small sizes, identical samples).
"""

from __future__ import annotations

import ctypes

import torch

from tts_max_b200 import _lib
from tts_max_b200.codec import decoder as decoder_lib


def slide_window(window: torch.Tensor, filled: int, left_context: int, new_ids: torch.Tensor) -> int:
    """In-place update of the (n_streams, left_context + new) window `[context | new]`: the context
    becomes the last `filled` tokens of what the window held (right-aligned), the new tokens go in
    behind it. Returns the context length valid for the NEXT push."""
    n_new = new_ids.shape[1]
    if filled > 0:
        keep = min(filled, left_context)
        window[:, left_context - keep:left_context] = \
            window[:, left_context + n_new - keep:left_context + n_new].clone()
    window[:, left_context:] = new_ids
    return min(left_context, filled + n_new)


class StreamingDecoder:
    """`n_streams` independent streams advancing in lock step, `new_tokens` tokens per `push`."""

    def __init__(self, decoder: decoder_lib.Decoder, n_streams: int, new_tokens: int = 50,
                 left_context: int = 100, use_graph: bool = True):
        if n_streams <= 0 or new_tokens <= 0 or left_context < 0:
            raise ValueError("n_streams and new_tokens must be positive, left_context non-negative")
        self._dec = decoder
        self._device = decoder.device
        if self._device.type != "cuda":
            raise RuntimeError("StreamingDecoder needs a decoder on a CUDA device (there is no CPU path)")
        self.n_streams = int(n_streams)
        self.new_tokens = int(new_tokens)
        self.left_context = int(left_context)
        self.samples_per_token = decoder.samples_per_token
        self._use_graph = bool(use_graph)
        win = self.left_context + self.new_tokens
        # the window of every stream, packed back to back: [context (oldest first) | new]
        self._window = torch.zeros(self.n_streams, win, dtype=torch.int64, device=self._device)
        self._filled = 0  # valid context tokens (identical for all streams: they advance together)
        self._graph: torch.cuda.CUDAGraph | None = None
        self._graph_wav: torch.Tensor | None = None
        self._graph_generation = -1  # Decoder.plan_generation() right after capture
        self._stale_captures = 0     # consecutive captures that were invalidated before their first replay

    # ------------------------------------------------------------------------------------------
    def reset(self) -> None:
        """Start new streams (drops the token history). The warm-up pushes that follow decode other
        shapes on the same decoder, which rebuilds its plan: the captured graph is detected as stale
        (plan generation) and re-captured at the first steady push."""
        self._filled = 0
        self._window.zero_()

    @property
    def context_tokens(self) -> int:
        return self._filled

    @torch.no_grad()
    def push(self, new_ids: torch.Tensor) -> torch.Tensor:
        """new_ids (n_streams, new_tokens) integer FSQ ids -> (n_streams, new_tokens * samples_per_token)
        float32 on the decoder's device: the audio of exactly these tokens, given each stream's context."""
        if new_ids.shape != (self.n_streams, self.new_tokens):
            raise ValueError(f"new_ids must be ({self.n_streams}, {self.new_tokens})")
        new_ids = new_ids.to(device=self._device, dtype=torch.int64)
        n_new, ctx = self.new_tokens, self._filled
        spt = self.samples_per_token
        next_filled = slide_window(self._window, ctx, self.left_context, new_ids)
        steady = ctx >= self.left_context
        if steady:
            wav = self._decode_steady()
        else:
            # warm-up of a stream: the window is still growing, shapes change from step to step
            cur = self._window[:, self.left_context - ctx:].contiguous()
            wav = self._dec.decode_packed_device(cur.view(-1), [ctx + n_new] * self.n_streams)
            wav = wav.view(self.n_streams, (ctx + n_new) * spt)
        out = wav[:, -n_new * spt:].clone()
        self._filled = next_filled
        return out

    # ------------------------------------------------------------------------------------------
    def _decode_steady(self) -> torch.Tensor:
        win = self.left_context + self.new_tokens
        seqlens = [win] * self.n_streams
        if not self._use_graph:
            return self._dec.decode_packed_device(self._window.view(-1), seqlens).view(self.n_streams, -1)
        # The graph bakes in the decoder handle's plan / workspace / statistics pointers and the plan
        # CONTENTS. Any decode of another shape on this decoder (a stream's own warm-up pushes after
        # reset(), or an unrelated decode between pushes) rewrites or reallocates them, so the graph is
        # replayed only while the handle's plan generation is the one seen right after capture.
        if self._graph is not None and self._dec.plan_generation() == self._graph_generation:
            self._stale_captures = 0
            self._graph.replay()
            return self._graph_wav.view(self.n_streams, -1)
        if self._graph is not None:
            self._stale_captures += 1
        self._graph = None
        if self._stale_captures >= 3:
            # something decodes other shapes on this decoder between every two pushes: capturing again each time
            # would cost more than the replay saves
            return self._dec.decode_packed_device(self._window.view(-1), seqlens).view(self.n_streams, -1)
        # eager decode of this push: the result, and the plan / workspace / tensor maps of this shape
        wav = self._dec.decode_packed_device(self._window.view(-1), seqlens).view(self.n_streams, -1)
        # capture for the NEXT push (the capture itself executes nothing)
        stream = torch.cuda.Stream(device=self._device)
        stream.wait_stream(torch.cuda.current_stream(self._device))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            self._graph_wav = self._dec.decode_packed_device(self._window.view(-1), seqlens)
        torch.cuda.current_stream(self._device).wait_stream(stream)
        self._graph = graph
        self._graph_generation = self._dec.plan_generation()
        return wav


class CachedStreamingDecoder:
    """Stateful chunked decode with CACHED attention state (`b200codec_stream_*`, include/b200codec.h): a push
    runs the model on [overlap | new] rows only (about (overlap + new) / new of the FLOPs of the new audio,
    against 3x for the window method at 100 + 50 tokens). `overlap` previous tokens are recomputed as the left
    halo of the convolutions and of the ISTFT overlap-add; attention reads the keys / values of the last
    `left_context` tokens from a per-layer ring, as they were computed when those tokens were new.

    Different function from `StreamingDecoder` (which is the reference forward on the window, trimmed); its
    oracle is oracle/streaming_oracle.py, and tests/test_streaming.py reports both methods' SNR against the
    one-shot decode."""

    def __init__(self, decoder: decoder_lib.Decoder, n_streams: int, new_tokens: int = 50,
                 left_context: int = 100, overlap: int = 8):
        if n_streams <= 0 or new_tokens <= 0 or left_context < 0 or overlap < 0:
            raise ValueError("n_streams and new_tokens must be positive, left_context and overlap non-negative")
        if decoder.device.type != "cuda":
            raise RuntimeError("CachedStreamingDecoder needs a decoder on a CUDA device (there is no CPU path)")
        self._dec = decoder
        self._device = decoder.device
        self.n_streams, self.new_tokens, self.left_context = int(n_streams), int(new_tokens), int(left_context)
        self.samples_per_token = decoder.samples_per_token
        lib = _lib.load()
        st = ctypes.c_void_p()
        _lib.check(lib.b200codec_stream_create(decoder._ensure_handle(), self.n_streams, self.new_tokens,
                                               self.left_context, ctypes.byref(st)))
        self._state = st
        self.capacity = int(lib.b200codec_stream_capacity(st))
        self.overlap = min(int(overlap), self.capacity - self.new_tokens)
        self._hist = torch.zeros(self.n_streams, max(self.overlap, 1), dtype=torch.int64, device=self._device)
        self._seen = 0

    def __del__(self) -> None:
        try:
            if getattr(self, "_state", None) is not None:
                _lib.load().b200codec_stream_destroy(self._state)
                self._state = None
        except Exception:  # interpreter shutdown
            pass

    def reset(self) -> None:
        """Start new streams: the rings are treated as empty again."""
        _lib.check(_lib.load().b200codec_stream_reset(self._state))
        self._seen = 0

    @property
    def context_tokens(self) -> int:
        return min(self._seen, self.capacity - self.new_tokens)

    @torch.no_grad()
    def push(self, new_ids: torch.Tensor) -> torch.Tensor:
        """new_ids (n_streams, new_tokens) -> (n_streams, new_tokens * samples_per_token) float32 on the device."""
        if new_ids.shape != (self.n_streams, self.new_tokens):
            raise ValueError(f"new_ids must be ({self.n_streams}, {self.new_tokens})")
        new_ids = new_ids.to(device=self._device, dtype=torch.int64)
        ov = min(self.overlap, self._seen)
        ids = torch.cat([self._hist[:, self._hist.shape[1] - ov:], new_ids], dim=1).contiguous() if ov else new_ids.contiguous()
        spt = self.samples_per_token
        lib = _lib.load()
        with torch.cuda.device(self._device):
            wav = torch.empty(self.n_streams, (ov + self.new_tokens) * spt, dtype=torch.float32, device=self._device)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(lib.b200codec_stream_push(self._dec._ensure_handle(), self._state, ctypes.c_void_p(ids.data_ptr()),
                                                 _lib.IDS_I64, ov, ctypes.c_void_p(wav.data_ptr()), ctypes.c_void_p(stream)))
        if self.overlap:
            keep = torch.cat([self._hist, new_ids], dim=1)[:, -self._hist.shape[1]:]
            self._hist.copy_(keep)
        self._seen += self.new_tokens
        return wav[:, ov * spt:]
