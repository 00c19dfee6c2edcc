"""Codec encoder -- B200-native mirror of `tts/core/codec/encoder.py` (class `Encoder`) up to the FSQ ids,
minus the w2v-BERT model.

What runs on the B200 (libb200codec.so, `b200enc_*`): the acoustic encoder
(`encoder_modules.AcousticEncoder`, encoder_modules.py:128-191), the semantic encoder
(`encoder_modules.SemanticEncoder`, :72-125), the fusion layer and the FSQ quantise
(encoder.py:42, 66-78). What stays outside: `transformers.Wav2Vec2BertModel` (encoder.py:50-55, 63); its
`hidden_states[16]` is an INPUT here (a tensor, or a callable that maps the padded audio to it).

Same state-dict keys as the reference `Encoder` for `semantic_encoder.*`, `acoustic_encoder.*`,
`fusion_layer.*`, `quantizer.project_{in,out}.*`, and the same two checkpoint layouts in
`load_from_checkpoint` (encoder.py:80-112). There is no CPU fallback.

`FSQQuantizer` (the quantise step alone, over a Decoder's quantizer weights) is kept from round 1.
"""

from __future__ import annotations

import collections
import ctypes
import logging
from typing import Any, Callable, Sequence

import torch

from tts_max_b200 import _lib
from tts_max_b200.codec import decoder as decoder_lib

_LOG = logging.getLogger(__name__)

_HOP_LENGTH = 320
_HALF_HOP_LENGTH = _HOP_LENGTH // 2
CODEC_SAMPLE_RATE = 16000   # tts/core/constants.py
CODEC_TOKENS_RATE = 50


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


def library_state_dict_shapes(device: int = 0) -> "collections.OrderedDict[str, tuple[int, ...]]":
    """Keys / shapes the encoder handle expects, read from the library (needs a GPU; tests check that
    `expected_state_dict_shapes` restates exactly this)."""
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.b200enc_create(0, device, ctypes.byref(h)))
    try:
        out: "collections.OrderedDict[str, tuple[int, ...]]" = collections.OrderedDict()
        shape = (ctypes.c_int64 * 4)()
        for i in range(lib.b200enc_num_tensors(h)):
            nd = lib.b200enc_tensor_shape(h, i, shape)
            out[lib.b200enc_tensor_key(h, i).decode()] = tuple(int(shape[d]) for d in range(nd))
        return out
    finally:
        lib.b200enc_destroy(h)


def expected_state_dict_shapes() -> "collections.OrderedDict[str, tuple[int, ...]]":
    """Keys / shapes of `Encoder.state_dict()` (encoder.py:28-47) without the `wav2vec_model.*` entries, in
    module order; the Activation1d filters are registered buffers and therefore state-dict entries."""
    sd: "collections.OrderedDict[str, tuple[int, ...]]" = collections.OrderedDict()

    def act(p: str, c: int) -> None:
        sd[p + "act.alpha"] = (c,)
        sd[p + "act.beta"] = (c,)
        sd[p + "upsample.filter"] = (1, 1, 12)
        sd[p + "downsample.lowpass.filter"] = (1, 1, 12)

    def wn(p: str, cout: int, cin: int, k: int) -> None:
        sd[p + "bias"] = (cout,)
        sd[p + "weight_g"] = (cout, 1, 1)
        sd[p + "weight_v"] = (cout, cin, k)

    s = "semantic_encoder."
    sd[s + "initial_conv.weight"] = (1024, 1024, 3)
    sd[s + "residual_blocks.1.weight"] = (1024, 1024, 3)
    sd[s + "residual_blocks.1.bias"] = (1024,)
    sd[s + "residual_blocks.3.weight"] = (1024, 1024, 3)
    sd[s + "residual_blocks.3.bias"] = (1024,)
    sd[s + "final_conv.weight"] = (1024, 1024, 3)
    a = "acoustic_encoder."
    wn(a + "conv_blocks.0.", 48, 1, 7)
    d = 48
    for i, stride in enumerate((2, 2, 4, 4, 5)):
        p = f"{a}conv_blocks.{i + 1}."
        for u in range(3):
            q = f"{p}block.{u}."
            act(q + "block.0.", d)
            wn(q + "block.1.", d, d, 7)
            act(q + "block.2.", d)
            wn(q + "block.3.", d, d, 1)
        act(p + "block.3.", d)
        wn(p + "block.4.", 2 * d, d, 2 * stride)
        d *= 2
    act(a + "conv_final_block.0.", d)
    wn(a + "conv_final_block.1.", 1024, d, 3)
    sd["fusion_layer.weight"] = (2048, 2048)
    sd["fusion_layer.bias"] = (2048,)
    sd["quantizer.project_in.weight"] = (8, 2048)
    sd["quantizer.project_in.bias"] = (8,)
    sd["quantizer.project_out.weight"] = (2048, 8)
    sd["quantizer.project_out.bias"] = (2048,)
    return sd


class Encoder(torch.nn.Module):
    """The audio encoder model (B200-native), reference `Encoder` (encoder.py:17-128) minus w2v-BERT.

    `forward(wavs, w2v_hidden)`: wavs (B, 1, S) float32 with S a multiple of 320, w2v_hidden (B, T, 1024) float32
    = `wav2vec_model(feats).hidden_states[16]` of the same audio (encoder.py:63) -> (B, 1, T) int32 ids.
    `pre_bound` is required for the same reason as in `FSQQuantizer`."""

    def __init__(self, model_path: str | None = None, *, pre_bound: bool, precision: str = "bf16"):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.sample_rate = CODEC_SAMPLE_RATE
        self.token_rate = CODEC_TOKENS_RATE
        self.precision = precision
        self._pre_bound = bool(pre_bound)
        self._shapes = expected_state_dict_shapes()
        self._host_state: "collections.OrderedDict[str, torch.Tensor]" = collections.OrderedDict()
        self._handle: ctypes.c_void_p | None = None
        self._device = torch.device("cpu")
        self._dirty = True
        self._tap_clips = 0
        # clips of one forward() are encoded in groups of at most this many samples (~3 KB of workspace per sample)
        self.max_batch_samples = 4_000_000
        if model_path is not None:
            self.load_from_checkpoint(model_path)

    # ------------------------------------------------------------------ module surface
    def to(self, device: Any = None, *args: Any, **kwargs: Any) -> "Encoder":  # type: ignore[override]
        if device is None:
            return self
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if dev != self._device:
            self._release()
            self._device = dev
        return self

    def state_dict(self, *args: Any, **kwargs: Any) -> "collections.OrderedDict[str, torch.Tensor]":  # type: ignore[override]
        return collections.OrderedDict((k, v.clone()) for k, v in self._host_state.items())

    def load_state_dict(self, state_dict: Any, strict: bool = True, assign: bool = False) -> Any:  # type: ignore[override]
        # `wav2vec_model.*` entries of a full reference state dict are not this module's business
        state_dict = {k: v for k, v in state_dict.items() if not k.startswith("wav2vec_model.")}
        missing = [k for k in self._shapes if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._shapes]
        errors = []
        if strict and unexpected:
            errors.append("Unexpected key(s) in state_dict: " + ", ".join(f'"{k}"' for k in unexpected) + ". ")
        if strict and missing:
            errors.append("Missing key(s) in state_dict: " + ", ".join(f'"{k}"' for k in missing) + ". ")
        for k, shape in self._shapes.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shape):
                errors.append(f"size mismatch for {k}: copying a param with shape {tuple(state_dict[k].shape)} "
                              f"from checkpoint, the shape in current model is {tuple(shape)}.")
        if errors:
            raise RuntimeError("Error(s) in loading state_dict for Encoder:\n\t" + "\n\t".join(errors))
        for k in self._shapes:
            if k in state_dict:
                self._host_state[k] = state_dict[k].detach().to("cpu", torch.float32).contiguous().clone()
        self._dirty = True
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def load_from_checkpoint(self, checkpoint_path: str) -> None:
        """Both layouts of the reference (encoder.py:80-112)."""
        _LOG.info("Loading encoder checkpoint from %s", checkpoint_path)
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        if "state_dict" in ckpt.keys():
            # https://huggingface.co/HKUSTAudio/xcodec2/tree/main/ckpt layout
            ckpt = ckpt["state_dict"]
            merged: "collections.OrderedDict[str, torch.Tensor]" = collections.OrderedDict()
            for key, value in ckpt.items():
                if key.startswith("CodecEnc."):
                    merged["acoustic_encoder." + key[len("CodecEnc."):]] = value
                elif key.startswith("generator.quantizer."):
                    merged["quantizer." + key[len("generator.quantizer."):]] = value
                elif key.startswith("SemanticEncoder_module."):
                    merged["semantic_encoder." + key[len("SemanticEncoder_module."):]] = value
                elif key.startswith("fc_prior."):
                    merged["fusion_layer." + key[len("fc_prior."):]] = value
            self.load_state_dict(merged, strict=True)
        else:
            self.load_state_dict(ckpt, strict=True)

    # ------------------------------------------------------------------ native handle
    def _release(self) -> None:
        if self._handle is not None:
            _lib.load().b200enc_destroy(self._handle)
            self._handle = None
        self._dirty = True

    def __del__(self) -> None:
        try:
            self._release()
        except Exception:  # interpreter shutdown
            pass

    def _ensure_handle(self) -> ctypes.c_void_p:
        if self._device.type != "cuda":
            raise RuntimeError("tts_max_b200 Encoder runs only on a CUDA (sm_100a / B200) device and has no CPU "
                               f"fallback; call .to('cuda') first (current device: {self._device})")
        lib = _lib.load()
        if self._handle is None:
            h = ctypes.c_void_p()
            _lib.check(lib.b200enc_create(_lib.PRECISIONS[self.precision], self._device.index or 0, ctypes.byref(h)))
            self._handle = h
            self._dirty = True
        if self._dirty:
            missing = [k for k in self._shapes if k not in self._host_state]
            if missing:
                raise RuntimeError("Encoder weights have not been loaded: missing " + ", ".join(missing[:4]) + " ...")
            for key, t in self._host_state.items():
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.b200enc_load_tensor(self._handle, key.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()))
            with torch.cuda.device(self._device):
                stream = torch.cuda.current_stream(self._device).cuda_stream
                _lib.check(lib.b200enc_finalize_weights(self._handle, ctypes.c_void_p(stream)))
            self._dirty = False
        return self._handle

    # ------------------------------------------------------------------ compute
    @torch.no_grad()
    def forward(self, wavs: torch.Tensor, w2v_hidden: torch.Tensor, return_hidden: bool = False):
        """Computes VQ codes for a batch of audio (reference: encoder.py:58-71 with `wav2vec_model(...)
        .hidden_states[16]` passed in). The clips of a batch share one launch sequence (80 kernels): they sit
        in one padded row space, separated by zero gap rows that act as the convolutions' zero padding."""
        if wavs.dim() != 3 or wavs.shape[1] != 1:
            raise ValueError(f"wavs must be (batch, 1, samples), got {tuple(wavs.shape)}")
        b, _, s = wavs.shape
        if s == 0 or s % _HOP_LENGTH != 0:
            raise ValueError(f"the number of samples ({s}) must be a positive multiple of {_HOP_LENGTH}")
        t = s // _HOP_LENGTH
        if tuple(w2v_hidden.shape) != (b, t, 1024):
            raise ValueError(f"w2v_hidden must be ({b}, {t}, 1024), got {tuple(w2v_hidden.shape)}")
        handle = self._ensure_handle()
        lib = _lib.load()
        wavs = wavs.to(self._device, torch.float32).contiguous()
        w2v_hidden = w2v_hidden.to(self._device, torch.float32).contiguous()
        with torch.cuda.device(self._device):
            ids = torch.empty(b, t, dtype=torch.int32, device=self._device)
            hidden = torch.empty(b, t, 2048, dtype=torch.float32, device=self._device) if return_hidden else None
            acoustic = torch.empty(b, t, 1024, dtype=torch.float32, device=self._device) if return_hidden else None
            semantic = torch.empty(b, t, 1024, dtype=torch.float32, device=self._device) if return_hidden else None
            stream = torch.cuda.current_stream(self._device).cuda_stream
            # one launch sequence per group of clips; a group is bounded by `max_batch_samples` of workspace
            group = max(1, min(b, self.max_batch_samples // (s + 6 * _HOP_LENGTH)))
            for i in range(0, b, group):
                n = min(group, b - i)
                _lib.check(lib.b200enc_encode_batch(
                    handle, ctypes.c_void_p(wavs[i].data_ptr()), n, s, ctypes.c_void_p(w2v_hidden[i].data_ptr()),
                    ctypes.c_void_p(ids[i].data_ptr()), _lib.IDS_I32, 1 if self._pre_bound else 0,
                    ctypes.c_void_p(hidden[i].data_ptr()) if hidden is not None else None,
                    ctypes.c_void_p(acoustic[i].data_ptr()) if acoustic is not None else None,
                    ctypes.c_void_p(semantic[i].data_ptr()) if semantic is not None else None,
                    ctypes.c_void_p(stream)))
                self._tap_clips = n
        vq_code = ids.view(b, t, 1).permute(0, 2, 1)     # (B, 1, T), like Encoder.quantize (encoder.py:73-78)
        if return_hidden:
            return vq_code, {"hidden": hidden.permute(0, 2, 1), "acoustic": acoustic.permute(0, 2, 1),
                             "semantic": semantic.permute(0, 2, 1)}
        return vq_code

    @torch.no_grad()
    def encode(self, wav: torch.Tensor, w2v_hidden: torch.Tensor | Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        """Encodes a waveform (1, S) into a sequence of tokens (reference: encoder.py:114-128). The padding is
        the reference's: right-pad to the next multiple of 320 (a whole extra hop when S already is one), and
        the w2v-BERT feature extractor sees that audio padded by 160 samples on both sides. `w2v_hidden`: the
        (1, T, 1024) hidden state, or a callable `audio_pad (1, S' + 320) -> (1, T, 1024)` wrapping the
        HuggingFace feature extractor + `Wav2Vec2BertModel(...).hidden_states[16]`."""
        if wav.dim() != 2 or wav.shape[0] != 1:
            raise ValueError(f"wav must be (1, samples), got {tuple(wav.shape)}")
        audio = torch.nn.functional.pad(wav.cpu(), (0, _HOP_LENGTH - (wav.shape[1] % _HOP_LENGTH)))
        if callable(w2v_hidden):
            audio_pad = torch.nn.functional.pad(audio, (_HALF_HOP_LENGTH, _HALF_HOP_LENGTH))
            w2v_hidden = w2v_hidden(audio_pad)
        return self.forward(audio.unsqueeze(0), w2v_hidden).squeeze()

    def set_stage_taps(self, on: bool) -> None:
        _lib.check(_lib.load().b200enc_set_stage_taps(self._ensure_handle(), 1 if on else 0))

    def read_stage(self, name: str, n_samples: int) -> torch.Tensor:
        """fp32 copy of conv_blocks[i]'s output of the LAST launch sequence ("conv0", "block1".."block5"): (rows, C)
        when it encoded one clip, (clips, rows, C) for a group of clips."""
        idx = 0 if name == "conv0" else int(name[len("block"):])
        rows, c = n_samples, 48
        for stride in (2, 2, 4, 4, 5)[:idx]:
            rows //= stride
            c *= 2
        clips = max(1, self._tap_clips)
        out = torch.empty(clips, rows, c, dtype=torch.float32)
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(_lib.load().b200enc_read_stage(self._ensure_handle(), name.encode(), n_samples,
                                                      ctypes.c_void_p(out.data_ptr()), out.numel(), ctypes.c_void_p(stream)))
        return out[0] if clips == 1 else out

    def launch_count(self) -> int:
        return 0 if self._handle is None else int(_lib.load().b200enc_launch_count(self._handle))
