"""Audio encoder entry point -- mirror of `tts/core/codec/encoding.py` (`AudioEncoderInterface`, `AudioEncoder`,
`CachingAudioEncoder`, `create`) over the B200 encoder (`tts_max_b200/codec/encoder.py`).

The reference `Encoder` owns a HuggingFace w2v-BERT feature extractor and model (encoder.py:50-55) and calls them
inside `encode` (:114-128). Here that model stays a HuggingFace object outside the CUDA library: `AudioEncoder`
builds it exactly like the reference unless the caller hands over `w2v_hidden_fn`, a callable
`audio_pad (1, S' + 320) float32 -> (1, T, 1024)` (`hidden_states[16]`), e.g. a model that is already loaded.
"""

from __future__ import annotations

import abc
from typing import Callable

import torch

from tts_max_b200.codec import encoder


class AudioEncoderInterface(metaclass=abc.ABCMeta):
    """Abstract interface class for audio encoders (reference: encoding.py:8-27)."""

    @abc.abstractmethod
    def encode(self, wav: torch.Tensor) -> torch.Tensor:
        """Encodes a waveform into a sequence of tokens."""
        raise NotImplementedError("Subclasses must implement this encode method.")

    @property
    @abc.abstractmethod
    def sample_rate(self) -> int:
        """Returns the input sample rate of the audio decoder."""
        raise NotImplementedError("Subclasses must implement this property.")

    @property
    @abc.abstractmethod
    def token_rate(self) -> int:
        """Returns the output token rate of the audio encoder."""
        raise NotImplementedError("Subclasses must implement this property.")


def hf_w2v_bert_hidden_fn(device: torch.device | str) -> Callable[[torch.Tensor], torch.Tensor]:
    """The reference's semantic front end (encoder.py:50-55, 121-123, 63): `facebook/w2v-bert-2.0` feature extractor
    + `Wav2Vec2BertModel(...).hidden_states[16]`. Needs the HuggingFace weights (hub or local cache)."""
    import transformers

    extractor = transformers.AutoFeatureExtractor.from_pretrained("facebook/w2v-bert-2.0")
    model = transformers.Wav2Vec2BertModel.from_pretrained("facebook/w2v-bert-2.0", output_hidden_states=True).to(device).eval()

    @torch.no_grad()
    def fn(audio_pad: torch.Tensor) -> torch.Tensor:
        feat = extractor(audio_pad, sampling_rate=encoder.CODEC_SAMPLE_RATE, return_tensors="pt").data["input_features"]
        return model(feat.to(device)).hidden_states[16]

    return fn


class AudioEncoder(AudioEncoderInterface):
    """Audio encoder class (reference: encoding.py:30-54)."""

    def __init__(self, model_path: str, device: torch.device | str, *, pre_bound: bool, precision: str = "bf16",
                 w2v_hidden_fn: Callable[[torch.Tensor], torch.Tensor] | None = None):
        super().__init__()
        self._device = torch.device(device)
        self._encoder = encoder.Encoder(model_path=model_path, pre_bound=pre_bound, precision=precision)
        self._encoder.to(self._device)
        self._encoder.eval()
        self._w2v_hidden_fn = w2v_hidden_fn

    @torch.no_grad()
    def encode(self, wav: torch.Tensor) -> torch.Tensor:
        """Encodes a waveform (1, S) into a sequence of tokens (T,)."""
        if self._w2v_hidden_fn is None:
            self._w2v_hidden_fn = hf_w2v_bert_hidden_fn(self._device)
        return self._encoder.encode(wav, self._w2v_hidden_fn)

    @property
    def sample_rate(self) -> int:
        return self._encoder.sample_rate

    @property
    def token_rate(self) -> int:
        return self._encoder.token_rate


class CachingAudioEncoder:
    """Encodes audios and caches the results (reference: encoding.py:57-72)."""

    def __init__(self, model_path: str, device: torch.device | str, **kwargs):
        super().__init__()
        self._encoder = create(model_path=model_path, device=device, **kwargs)
        self._prompt_encoding_cache: dict[str, list[int]] = {}

    @torch.no_grad()
    def encode(self, prompt_id: str, prompt_wav: torch.Tensor) -> list[int]:
        if prompt_id in self._prompt_encoding_cache:
            return self._prompt_encoding_cache[prompt_id]
        codes = self._encoder.encode(prompt_wav).cpu().tolist()
        self._prompt_encoding_cache[prompt_id] = codes
        return codes


def create(model_path: str, device: torch.device | str | None = "cuda", **kwargs) -> AudioEncoderInterface:
    """Create audio encoder with model path (reference: encoding.py:75-80; `device` defaults to "cuda" because
    there is no CPU path; `pre_bound=` is required, see `encoder.FSQQuantizer`)."""
    return AudioEncoder(model_path, device=device if device is not None else "cuda", **kwargs)
