"""B200-native (sm_100a) implementation of tts-max's xcodec2-compatible codec DECODE path.

The package mirrors the reference's `tts.core.codec.decoding` / `tts.core.codec.decoder`
surface (same names, arguments and error behaviour) over a C-ABI shared library of
hand-written CUDA kernels (`include/b200codec.h`, `tts_max_b200/csrc`). There is no CPU
fallback: compute entry points raise if the CUDA extension or a B200 is missing.
"""

from tts_max_b200.codec import decoder, decoding  # noqa: F401
from tts_max_b200.codec.decoding import (  # noqa: F401
    AudioDecoder,
    AudioDecoderInterface,
    DecoderConfig,
    create,
)

__all__ = ["decoder", "decoding", "AudioDecoder", "AudioDecoderInterface", "DecoderConfig", "create"]
