"""Decoding tokens to audio -- B200-native mirror of `tts/core/codec/decoding.py` (reference).

Same public surface: `DecoderConfig`, `AudioDecoderInterface`, `AudioDecoder`, `create`.
`AudioDecoder.decode` keeps the reference contract (1-D ids -> `(1, hop*T)` float32 CPU tensor,
synchronous; decoding.py:84-89) and `decode_batch` adds the varlen batch entry the high-volume
callers (rewards.py:113-159) can use instead of a serial B=1 loop.
"""

from __future__ import annotations

import abc
import dataclasses
import json
import os
from typing import Sequence

import torch

from tts_max_b200.codec import decoder


@dataclasses.dataclass(frozen=True)
class DecoderConfig:
    """Model config for the codec decoder model (reference: decoding.py:13-35)."""

    model_type: str
    sample_rate: int
    token_rate: int
    hop_length: int
    upsample_factors: list[int] | None
    kernel_sizes: list[int] | None

    @staticmethod
    def from_json(file: str | os.PathLike) -> "DecoderConfig":
        with open(file) as f:
            config = json.load(f)
        return DecoderConfig(
            # the xcodec2 example config ships without "model_type"
            # (example/codec/model_config.json); the trainer writes it (train_codec.py:65-72)
            model_type=config.get("model_type", ""),
            sample_rate=config["sample_rate"],
            token_rate=config["token_rate"],
            hop_length=config["hop_length"],
            upsample_factors=config["upsample_factors"],
            kernel_sizes=config["kernel_sizes"],
        )


class AudioDecoderInterface(metaclass=abc.ABCMeta):
    """Abstract interface class for audio decoders (reference: decoding.py:38-56)."""

    @abc.abstractmethod
    def decode(self, speech_ids: torch.Tensor) -> torch.Tensor:
        """Decodes a batch of speech IDs into audio waveforms."""
        raise NotImplementedError("Subclasses must implement this decode method.")

    @property
    @abc.abstractmethod
    def sample_rate(self) -> int:
        """Returns the sample rate of the audio decoder."""
        raise NotImplementedError("Subclasses must implement this property.")

    @property
    @abc.abstractmethod
    def token_rate(self) -> int:
        """Returns the input token rate of the audio decoder."""
        raise NotImplementedError("Subclasses must implement this property.")


class AudioDecoder(AudioDecoderInterface):
    """A wrapper around the audio decoder model for batch decoding (reference: decoding.py:59-97)."""

    def __init__(
        self,
        model_path: str | None,
        config: DecoderConfig,
        device: torch.device | str | None = "cuda",
        precision: str = "bf16",
    ):
        super().__init__()
        if device is None:
            device = "cuda"
        self._device = torch.device(device)
        self._decoder = decoder.Decoder(
            sample_rate=config.sample_rate,
            hop_length=config.hop_length,
            upsample_factors=config.upsample_factors,
            kernel_sizes=config.kernel_sizes,
            checkpoint_path=model_path,
            precision=precision,
        )
        self._decoder.to(self._device)
        self._decoder.eval()

        self._sample_rate = config.sample_rate
        self._token_rate = config.token_rate

    @torch.no_grad()
    def decode(self, speech_ids: torch.Tensor) -> torch.Tensor:
        """Decodes one utterance: 1-D integer ids -> (1, hop_length * T) float32 on the CPU."""
        if speech_ids.dim() != 1:
            raise ValueError(f"speech_ids must be 1-D, got shape {tuple(speech_ids.shape)}")
        if speech_ids.numel() == 0:
            raise ValueError("decode: empty speech_ids")
        if speech_ids.device.type == "cpu":
            ids = speech_ids if speech_ids.dtype in (torch.int32, torch.int64) else speech_ids.to(torch.int64)
            wav = self._decoder.decode_packed_host(ids, [ids.numel()])
            return wav.view(1, -1)
        vq_codes = speech_ids.unsqueeze(0).unsqueeze(0)
        vq_codes = vq_codes.to(self._device)
        wav = self._decoder(vq_codes).detach().cpu().squeeze(0)
        if self._decoder.take_id_error():
            raise ValueError("speech id outside [0, 65535] in speech_ids")
        return wav

    @torch.no_grad()
    def decode_batch(self, speech_ids: Sequence[torch.Tensor]) -> list[torch.Tensor]:
        """Decodes several utterances of different lengths in one varlen launch sequence. Each
        result equals `decode(ids)` of that utterance (per-utterance GroupNorm / attention /
        conv-edge semantics; no padding is involved)."""
        if len(speech_ids) == 0:
            return []
        seqlens = [int(t.numel()) for t in speech_ids]
        for t in speech_ids:
            if t.dim() != 1 or t.numel() == 0:
                raise ValueError("decode_batch: every element must be a non-empty 1-D id tensor")
        packed = torch.cat([t.detach().to("cpu", torch.int64) for t in speech_ids])
        wav = self._decoder.decode_packed_host(packed, seqlens)
        hop = self._decoder.samples_per_token
        out, off = [], 0
        for n in seqlens:
            out.append(wav[off * hop:(off + n) * hop].view(1, -1))
            off += n
        return out

    @property
    def sample_rate(self) -> int:
        return self._sample_rate

    @property
    def token_rate(self) -> int:
        return self._token_rate


def create(
    model_path: str, device: torch.device | str | None = "cuda", precision: str = "bf16"
) -> AudioDecoderInterface:
    """Create audio decoder with model path and optionally a config file
    (reference: decoding.py:100-112; `device` defaults to "cuda" because there is no CPU path)."""

    ckpt_dir = os.path.dirname(model_path)
    config_path = os.path.join(ckpt_dir, "model_config.json")

    if not os.path.exists(config_path):
        raise ValueError("No model_config.json found in the provided path.")

    model_config = DecoderConfig.from_json(config_path)
    return AudioDecoder(model_path, model_config, device=device, precision=precision)
