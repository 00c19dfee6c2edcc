"""Encode-direction FSQ quantise: the part of `tts.core.codec.encoder.Encoder` that mirrors the
decoder's K1. The acoustic / semantic encoders are out of scope; this module
takes the fused hidden states they produce.

Reference: `Encoder.quantize` (tts/core/codec/encoder.py:73-78)

    hidden_states = hidden_states.permute(0, 2, 1)      # (B, 2048, T) -> (B, T, 2048)
    _, vq_code = self.quantizer(hidden_states)          # ResidualFSQ.forward -> indices (B, T, 1)
    return vq_code.permute(0, 2, 1)                     # (B, 1, T)

The quantizer weights are the `generator.quantizer.*` tensors of the codec checkpoint
(encoder.py:98-111), i.e. the ones a `Decoder` has already loaded.
"""

import torch

from tts_max_b200.codec import decoder as decoder_lib


class FSQQuantizer:
    """`quantize(hidden_states)` with the reference's shapes, backed by b200codec_fsq_quantize."""

    def __init__(self, decoder: decoder_lib.Decoder, *, pre_bound: bool):
        """`pre_bound` is REQUIRED: releases of vector-quantize-pytorch differ in whether
        `ResidualFSQ.forward` applies `layers[0].bound` to the projected input before the layer loop,
        the pinned 1.17.8 wheel is not available offline, and the two variants give different ids for
        most inputs (bound(bound(z)) != bound(z)). Encode parity is therefore UNVERIFIED at that library
        boundary; pass what the wheel you deploy against does."""
        self._decoder = decoder
        self._pre_bound = bool(pre_bound)

    @torch.no_grad()
    def quantize(self, hidden_states: torch.Tensor) -> torch.Tensor:
        """(B, 2048, T) float32 -> (B, 1, T) int32 FSQ ids in [0, 65536) (`FSQ.codes_to_indices`
        returns int32)."""
        if hidden_states.dim() != 3:
            raise ValueError("hidden_states must be (batch, channels, frames)")
        b, c, t = hidden_states.shape
        # the reference's fusion layer hands over a transposed view of a (B, T, 2048) buffer, so this
        # permute is normally free
        tok_major = hidden_states.permute(0, 2, 1).reshape(b * t, c)
        ids = self._decoder.quantize_features(tok_major.float(), pre_bound=self._pre_bound)
        return ids.view(b, t, 1).permute(0, 2, 1)
