"""Decoder model for audio waveform generation -- B200-native mirror of
`tts/core/codec/decoder.py:14-119` (class `Decoder`) of the reference.

Same constructor arguments, same `forward(vq_codes)` contract, same two checkpoint layouts in
`load_from_checkpoint`, same `state_dict()` key set. The compute is not
PyTorch: weights are handed to libb200codec.so tensor by tensor and `forward` is one C-ABI
call (`b200codec_decode_varlen`) that runs the hand-written sm_100a kernels. PyTorch is used
for device memory, streams and checkpoint I/O only. There is no CPU fallback.
"""

from __future__ import annotations

import collections
import ctypes
import logging
import math
import threading
from typing import Any, Iterable, Sequence

import torch

from tts_max_b200 import _lib

_LOG = logging.getLogger(__name__)

HIDDEN_DIM = 1024
DEPTH = 12
HEADS = 16
VQ_DIM = 2048


def expected_state_dict_shapes(
    hop_length: int = 320, depth: int = DEPTH, upsample_factors: Sequence[int] | None = None,
    kernel_sizes: Sequence[int] | None = None,
) -> "collections.OrderedDict[str, tuple[int, ...]]":
    """Keys and shapes of `Decoder.state_dict()`, in module order (reference: `Generator.__init__`
    decoder_modules.py:403-433, `VocosBackbone.__init__` :330-388, `UpSamplerBlock.__init__`
    upsampler.py:12-60, `Decoder.__init__` decoder.py:46-63). 117 tensors for the xcodec2 config, 145 for
    the 48 kHz config with upsample_factors [3, 2]."""
    C, V, n_fft = HIDDEN_DIM, VQ_DIM, 4 * hop_length
    sd: "collections.OrderedDict[str, tuple[int, ...]]" = collections.OrderedDict()
    g = "decoder."
    sd[g + "quantizer.project_in.weight"] = (8, V)
    sd[g + "quantizer.project_in.bias"] = (8,)
    sd[g + "quantizer.project_out.weight"] = (V, 8)
    sd[g + "quantizer.project_out.bias"] = (V,)
    sd[g + "backbone.embed.weight"] = (C, C, 7)
    sd[g + "backbone.embed.bias"] = (C,)

    def resnet(prefix: str) -> None:
        for n in ("1", "2"):
            sd[f"{prefix}norm{n}.weight"] = (C,)
            sd[f"{prefix}norm{n}.bias"] = (C,)
            sd[f"{prefix}conv{n}.weight"] = (C, C, 3)
            sd[f"{prefix}conv{n}.bias"] = (C,)

    resnet(g + "backbone.prior_net.0.")
    resnet(g + "backbone.prior_net.1.")
    for layer in range(depth):
        p = f"{g}backbone.transformers.{layer}."
        sd[p + "att_norm.weight"] = (C,)
        sd[p + "ffn_norm.weight"] = (C,)
        sd[p + "att.c_attn.weight"] = (3 * C, C)
        sd[p + "att.c_proj.weight"] = (C, C)
        sd[p + "mlp.fc1.weight"] = (4 * C, C)
        sd[p + "mlp.fc2.weight"] = (C, 4 * C)
    sd[g + "backbone.final_layer_norm.weight"] = (C,)
    sd[g + "backbone.final_layer_norm.bias"] = (C,)
    resnet(g + "backbone.post_net.0.")
    resnet(g + "backbone.post_net.1.")
    sd[g + "head.out.weight"] = (n_fft + 2, C)
    sd[g + "head.out.bias"] = (n_fft + 2,)
    sd[g + "head.istft.window"] = (n_fft,)
    if upsample_factors:
        n = len(upsample_factors)
        for i, k in enumerate(kernel_sizes):
            cin, cout = C // (2 ** i), C // (2 ** (i + 1))
            sd[f"upsampler.upsample_layers.{i}.bias"] = (cout,)
            sd[f"upsampler.upsample_layers.{i}.weight_g"] = (cin, 1, 1)
            sd[f"upsampler.upsample_layers.{i}.weight_v"] = (cin, cout, k)
        for i in range(n):
            c = C // (2 ** (i + 1))
            p = f"upsampler.resnet_blocks.{i}."
            sd[p + "norm1.weight"] = (c,)
            sd[p + "norm1.bias"] = (c,)
            sd[p + "conv1.weight"] = (c, c, 3)
            sd[p + "conv1.bias"] = (c,)
            sd[p + "temb_proj.weight"] = (c, 512)  # built with the default temb_channels; unused (temb=None)
            sd[p + "temb_proj.bias"] = (c,)
            sd[p + "norm2.weight"] = (c,)
            sd[p + "norm2.bias"] = (c,)
            sd[p + "conv2.weight"] = (c, c, 3)
            sd[p + "conv2.bias"] = (c,)
        sd["upsampler.out_proj.weight"] = (C, C // (2 ** n))
        sd["upsampler.out_proj.bias"] = (C,)
    sd["fc_post_a.weight"] = (C, V)
    sd["fc_post_a.bias"] = (C,)
    return sd


def random_init_state_dict(
    hop_length: int = 320, seed: int | None = None, upsample_factors: Sequence[int] | None = None,
    kernel_sizes: Sequence[int] | None = None,
) -> "collections.OrderedDict[str, torch.Tensor]":
    """Random initialisation with the reference's distributions (not its RNG stream):
    Conv1d weight trunc_normal(std=0.02) / bias 0 inside `Generator` (decoder_modules.py:13-16, 463-464:
    `self.apply(init_weights)` covers the backbone only), Linear = torch default
    (kaiming_uniform(a=sqrt(5)) -> U(+-1/sqrt(fan_in)) for weight and bias), norm weight 1 / bias 0,
    `window` = periodic hann (decoder_modules.py:32-33). The UpSamplerBlock sits outside `Generator`
    (decoder.py:48-61), so ITS Conv1d / ConvTranspose1d / Linear layers keep torch's defaults
    (U(+-1/sqrt(fan_in)) for weight and bias). Checked against the reference's own init statistics in
    tests/test_host_logic.py::test_random_init_matches_reference_statistics."""
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    out: "collections.OrderedDict[str, torch.Tensor]" = collections.OrderedDict()
    all_shapes = expected_state_dict_shapes(hop_length, DEPTH, upsample_factors, kernel_sizes)
    pending_g: dict[str, torch.Tensor] = {}
    for key, shape in all_shapes.items():
        if key.endswith("istft.window"):
            t = torch.hann_window(shape[0])
        elif key.endswith("weight_g"):
            t = torch.ones(shape)  # replaced by ||v|| below (torch.nn.utils.weight_norm initialisation)
        elif key.endswith("weight_v"):
            # ConvTranspose1d default init: kaiming_uniform(a=sqrt(5)) with fan_in = Cout * k
            bound = 1.0 / math.sqrt(shape[1] * shape[2])
            t = (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound
            pending_g[key[: -len("weight_v")] + "weight_g"] = t.reshape(shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
        elif "norm" in key:
            t = torch.ones(shape) if key.endswith("weight") else torch.zeros(shape)
        elif key.startswith("upsampler.resnet_blocks.") and ("conv" in key):
            # torch Conv1d default (no init_weights outside Generator): fan_in = Cin * k
            wshape = shape if len(shape) == 3 else all_shapes[key[: -len("bias")] + "weight"]
            bound = 1.0 / math.sqrt(wshape[1] * wshape[2])
            t = (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound
        elif len(shape) == 3:  # Conv1d weight
            t = torch.empty(shape)
            torch.nn.init.trunc_normal_(t, std=0.02, generator=gen)
        elif key.endswith("bias") and ("embed" in key or "conv" in key):
            t = torch.zeros(shape)
        elif key.endswith("bias") and "upsample_layers" in key:
            wv = all_shapes[key[: -len("bias")] + "weight_v"]
            bound = 1.0 / math.sqrt(wv[1] * wv[2])
            t = (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound
        else:  # Linear weight / bias: bound = 1 / sqrt(fan_in)
            wshape = shape if len(shape) == 2 else all_shapes[key[: -len("bias")] + "weight"]
            bound = 1.0 / math.sqrt(wshape[1])
            t = (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound
        out[key] = t.to(torch.float32).contiguous()
    for key, g in pending_g.items():
        out[key] = g.to(torch.float32).contiguous()
    return out


_DTYPE_TAGS = {torch.float32: _lib.DT_F32, torch.float16: _lib.DT_F16, torch.bfloat16: _lib.DT_BF16,
               torch.float64: _lib.DT_F64}


class Decoder(torch.nn.Module):
    """The decoder model for audio waveform generation (B200-native).

    Args mirror the reference (`tts/core/codec/decoder.py:17-24`); `precision` ("bf16" or
    "fp16") selects the tensor-core operand type and is the only addition.
    """

    def __init__(
        self,
        sample_rate: int,
        hop_length: int,
        upsample_factors: list[int] | None,
        kernel_sizes: list[int] | None,
        checkpoint_path: str | None = None,
        precision: str = "bf16",
        init_seed: int | None = None,
    ):
        super().__init__()
        self.sample_rate = sample_rate
        self.hop_length = hop_length
        self.upsample_factors = upsample_factors
        self.kernel_sizes = kernel_sizes
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.precision = precision

        total_ups = math.prod(self.upsample_factors) if self.upsample_factors else 1
        if self.sample_rate // self.hop_length // total_ups != 50:
            raise ValueError(  # reference: decoder.py:31-37
                f"Current hop length {self.hop_length} and upsample "
                f"factors {self.upsample_factors} do not match the target "
                f"sample rate {self.sample_rate}."
            )
        if self.upsample_factors:
            if not self.kernel_sizes or len(self.kernel_sizes) != len(self.upsample_factors):
                raise ValueError("kernel_sizes must match upsample_factors")
            if len(self.upsample_factors) > 3 or any((k - u) % 2 or k < u for k, u in zip(self.kernel_sizes, self.upsample_factors)):
                raise NotImplementedError(
                    f"upsampler configuration factors={self.upsample_factors} kernels={self.kernel_sizes} is not "
                    "instantiated (at most three stages: 512 / 256 / 128 channels; kernel - factor even)")
        if self.hop_length not in (320, 240, 160, 80):
            raise NotImplementedError("hop_length must be 320, 240, 160 or 80 (n_fft = 4 hop = 64 x {20, 15, 10, 5})")
        self.samples_per_token = self.hop_length * total_ups

        self._shapes = expected_state_dict_shapes(hop_length, DEPTH, self.upsample_factors, self.kernel_sizes)
        # host fp32 copy of the weights; the library holds the device copies once .to(cuda) ran
        self._host_state: "collections.OrderedDict[str, torch.Tensor]" = random_init_state_dict(
            hop_length, init_seed, self.upsample_factors, self.kernel_sizes)
        self._handle: ctypes.c_void_p | None = None
        self._pinned_out: torch.Tensor | None = None
        self._pinned_lock = threading.Lock()  # the staging buffer is shared by the callers of decode_packed_host
        self._device: torch.device = torch.device("cpu")
        self._dirty = True

        if checkpoint_path is not None:
            _LOG.info("Loading codec checkpoint from %s", checkpoint_path)
            self.load_from_checkpoint(checkpoint_path)

    # ------------------------------------------------------------------ module surface
    def to(self, device: Any = None, *args: Any, **kwargs: Any) -> "Decoder":  # type: ignore[override]
        if device is None:
            return self
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if dev != self._device:
            self._release()
            self._device = dev
        return self

    def cuda(self, device: Any = None) -> "Decoder":  # type: ignore[override]
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    @property
    def device(self) -> torch.device:
        return self._device

    def state_dict(self, *args: Any, **kwargs: Any) -> "collections.OrderedDict[str, torch.Tensor]":  # type: ignore[override]
        prefix = kwargs.get("prefix", "")
        return collections.OrderedDict((prefix + k, v.clone()) for k, v in self._host_state.items())

    def load_state_dict(self, state_dict: Any, strict: bool = True, assign: bool = False) -> Any:  # type: ignore[override]
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
            raise RuntimeError("Error(s) in loading state_dict for Decoder:\n\t" + "\n\t".join(errors))
        for k in self._shapes:
            if k in state_dict:
                self._host_state[k] = state_dict[k].detach().to("cpu", torch.float32).contiguous().clone()
        self._dirty = True
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def load_from_checkpoint(self, checkpoint_path: str) -> None:
        """Both layouts of the reference (decoder.py:91-119)."""
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        if "state_dict" in ckpt.keys():
            # https://huggingface.co/HKUSTAudio/xcodec2/tree/main/ckpt layout
            ckpt = ckpt["state_dict"]
            generator_sd = collections.OrderedDict()
            fc_post_sd = collections.OrderedDict()
            for key, value in ckpt.items():
                if key.startswith("generator."):
                    generator_sd["decoder." + key[len("generator."):]] = value
                elif key.startswith("fc_post_a."):
                    fc_post_sd[key] = value
            # the reference loads the two sub-modules strictly and independently
            want_gen = [k for k in self._shapes if k.startswith("decoder.")]
            want_fc = [k for k in self._shapes if k.startswith("fc_post_a.")]
            self._strict_subset(generator_sd, want_gen, "Generator")
            self._strict_subset(fc_post_sd, want_fc, "Linear")
            merged = collections.OrderedDict(generator_sd)
            merged.update(fc_post_sd)
            self.load_state_dict(merged, strict=True)
        else:
            ckpt = ckpt["model"]
            ckpt = {k.replace("generator.", ""): v for k, v in ckpt.items() if k.startswith("generator.")}
            self.load_state_dict(ckpt, strict=True)

    @staticmethod
    def _strict_subset(got: dict, want: Sequence[str], name: str) -> None:
        missing = [k for k in want if k not in got]
        unexpected = [k for k in got if k not in want]
        if missing or unexpected:
            msg = f"Error(s) in loading state_dict for {name}:"
            if missing:
                msg += "\n\tMissing key(s) in state_dict: " + ", ".join(f'"{k}"' for k in missing) + ". "
            if unexpected:
                msg += "\n\tUnexpected key(s) in state_dict: " + ", ".join(f'"{k}"' for k in unexpected) + ". "
            raise RuntimeError(msg)

    # ------------------------------------------------------------------ native handle
    def _release(self) -> None:
        if self._handle is not None:
            _lib.load().b200codec_destroy(self._handle)
            self._handle = None
        self._dirty = True

    def __del__(self) -> None:
        try:
            self._release()
        except Exception:  # interpreter shutdown
            pass

    def _ensure_handle(self) -> ctypes.c_void_p:
        if self._device.type != "cuda":
            raise RuntimeError(
                "tts_max_b200.Decoder runs only on a CUDA (sm_100a / B200) device and has no CPU "
                f"fallback; call .to('cuda') first (current device: {self._device})"
            )
        lib = _lib.load()
        if self._handle is not None and not self._dirty:
            return self._handle
        with self._pinned_lock:  # first use from several threads: create / load once
            return self._ensure_handle_locked(lib)

    def _ensure_handle_locked(self, lib) -> ctypes.c_void_p:
        if self._handle is None:
            ups = list(self.upsample_factors or [])
            ks = list(self.kernel_sizes or []) if ups else []
            cfg = _lib.B200CodecConfig(
                abi_version=_lib.ABI_VERSION, sample_rate=self.sample_rate, hop_length=self.hop_length,
                n_upsample=len(ups), precision=_lib.PRECISIONS[self.precision],
                device=self._device.index or 0, hidden_dim=HIDDEN_DIM, depth=DEPTH, heads=HEADS, vq_dim=VQ_DIM,
                upsample_factors=(ctypes.c_int32 * 3)(*(ups + [0] * (3 - len(ups)))),
                kernel_sizes=(ctypes.c_int32 * 3)(*(ks + [0] * (3 - len(ks)))),
            )
            handle = ctypes.c_void_p()
            _lib.check(lib.b200codec_create(ctypes.byref(cfg), ctypes.byref(handle)))
            self._handle = handle
            self._dirty = True
        if self._dirty:
            for key, t in self._host_state.items():
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.b200codec_load_tensor(self._handle, key.encode(), ctypes.c_void_p(t.data_ptr()),
                                                     _DTYPE_TAGS[t.dtype], shape, t.dim()))
            with torch.cuda.device(self._device):
                stream = torch.cuda.current_stream(self._device).cuda_stream
                _lib.check(lib.b200codec_finalize_weights(self._handle, ctypes.c_void_p(stream)))
            self._dirty = False
        return self._handle

    # ------------------------------------------------------------------ compute
    @torch.no_grad()
    def forward(self, vq_codes: torch.Tensor) -> torch.Tensor:
        """vq_codes: (batch, codes_length) or (batch, 1, codes_length) integer ids ->
        (batch, 1, hop_length * prod(upsample_factors) * codes_length) float32 on the decoder's device
        (reference: decoder.py:69-89)."""
        if vq_codes.dim() == 2:
            vq_codes = vq_codes.unsqueeze(1)
        if vq_codes.dim() != 3 or vq_codes.shape[1] != 1:
            raise ValueError(f"vq_codes must be (B, T) or (B, 1, T), got {tuple(vq_codes.shape)}")
        if vq_codes.dtype not in (torch.int32, torch.int64):
            if vq_codes.is_floating_point() or vq_codes.dtype == torch.bool:
                raise TypeError(f"vq_codes must be an integer tensor, got {vq_codes.dtype}")
            vq_codes = vq_codes.to(torch.int64)
        batch, _, length = vq_codes.shape
        if batch == 0 or length == 0:
            raise ValueError(f"decode: empty batch or empty utterance (shape {tuple(vq_codes.shape)})")
        ids = vq_codes.to(self._device).reshape(batch * length).contiguous()
        wavs = self.decode_packed_device(ids, [length] * batch)
        return wavs.view(batch, 1, self.samples_per_token * length)

    @torch.no_grad()
    def decode_packed_device(self, ids: torch.Tensor, seqlens: Sequence[int],
                             out: torch.Tensor | None = None) -> torch.Tensor:
        """Varlen decode of packed device ids (sum(seqlens),) -> packed device waveform
        (hop_length * sum(seqlens),); asynchronous on the current stream. `out` (optional): a
        contiguous float32 device tensor of exactly that many samples to write into (e.g. a slice of a
        shard-wide PCM buffer that is gathered afterwards)."""
        handle = self._ensure_handle()
        lib = _lib.load()
        total = int(sum(int(t) for t in seqlens))
        if ids.device != self._device or ids.dim() != 1 or ids.numel() != total:
            raise ValueError("ids must be a packed 1-D tensor on the decoder's device matching seqlens")
        id_type = _lib.IDS_I64 if ids.dtype == torch.int64 else _lib.IDS_I32
        n_out = total * self.samples_per_token
        if out is not None and (out.device != self._device or out.dtype != torch.float32 or out.numel() != n_out
                                or not out.is_contiguous()):
            raise ValueError("out must be a contiguous float32 tensor of hop * sum(seqlens) samples on the decoder's device")
        with torch.cuda.device(self._device):
            wav = out if out is not None else torch.empty(n_out, dtype=torch.float32, device=self._device)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(lib.b200codec_decode_varlen(handle, ctypes.c_void_p(ids.data_ptr()), id_type,
                                                   _lib.i32_array(seqlens), len(seqlens),
                                                   ctypes.c_void_p(wav.data_ptr()), ctypes.c_void_p(stream)))
        return wav

    @torch.no_grad()
    def decode_packed_host(self, ids: torch.Tensor, seqlens: Sequence[int], out: torch.Tensor | None = None) -> torch.Tensor:
        """Varlen decode with HOST buffers (ids CPU int32/int64, result CPU float32): H2D, the
        decode, D2H and the synchronisation happen inside one C-ABI call."""
        handle = self._ensure_handle()
        lib = _lib.load()
        total = int(sum(int(t) for t in seqlens))
        if ids.device.type != "cpu" or ids.dim() != 1 or ids.numel() != total:
            raise ValueError("ids must be a packed 1-D CPU tensor matching seqlens")
        if ids.dtype not in (torch.int32, torch.int64):
            raise TypeError(f"ids must be int32 or int64, got {ids.dtype}")
        ids = ids.contiguous()
        id_type = _lib.IDS_I64 if ids.dtype == torch.int64 else _lib.IDS_I32
        n_out = total * self.samples_per_token
        if out is not None:
            if out.device.type != "cpu" or out.dtype != torch.float32 or out.numel() != n_out or not out.is_contiguous():
                raise ValueError("out must be a contiguous float32 CPU tensor of hop * sum(seqlens) samples")
            dst = out
            with torch.cuda.device(self._device):
                stream = torch.cuda.current_stream(self._device).cuda_stream
                _lib.check(lib.b200codec_decode_host(handle, ctypes.c_void_p(ids.data_ptr()), id_type,
                                                     _lib.i32_array(seqlens), len(seqlens),
                                                     ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(stream)))
            return dst
        # One page-locked staging buffer per decoder (grow-only): the last kernel stores the PCM
        # straight into it (zero-copy). Callers get a pageable copy, so accumulating results (a
        # dataset sweep) never pins an unbounded amount of host memory; pass `out=` (ideally a
        # pinned tensor the caller reuses) to skip the copy. The lock covers decode + copy: the C call
        # releases the GIL, and another thread's decode would overwrite the staging buffer before the copy.
        with self._pinned_lock:
            if self._pinned_out is None or self._pinned_out.numel() < n_out:
                self._pinned_out = None
                self._pinned_out = torch.empty(n_out + n_out // 4, dtype=torch.float32, pin_memory=True)
            dst = self._pinned_out[:n_out]
            with torch.cuda.device(self._device):
                stream = torch.cuda.current_stream(self._device).cuda_stream
                _lib.check(lib.b200codec_decode_host(handle, ctypes.c_void_p(ids.data_ptr()), id_type,
                                                     _lib.i32_array(seqlens), len(seqlens),
                                                     ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(stream)))
            return dst.clone()

    def decode_packed_host_async(self, ids: torch.Tensor, seqlens: Sequence[int], out: torch.Tensor) -> "torch.cuda.Event":
        """`decode_packed_host` without the final synchronisation, for callers that keep several batches in
        flight: `out` must be a PINNED float32 CPU tensor (the last kernel stores the PCM straight into it), `ids`
        should be pinned too and must stay alive until the returned event has completed. Returns an event
        recorded behind the decode on the current stream; `event.synchronize()` before reading `out`."""
        handle = self._ensure_handle()
        lib = _lib.load()
        total = int(sum(int(t) for t in seqlens))
        if ids.device.type != "cpu" or ids.dim() != 1 or ids.numel() != total or not ids.is_contiguous():
            raise ValueError("ids must be a packed, contiguous 1-D CPU tensor matching seqlens")
        if ids.dtype not in (torch.int32, torch.int64):
            raise TypeError(f"ids must be int32 or int64, got {ids.dtype}")
        n_out = total * self.samples_per_token
        if (out.device.type != "cpu" or out.dtype != torch.float32 or out.numel() != n_out or not out.is_contiguous()
                or not out.is_pinned()):
            raise ValueError("out must be a pinned, contiguous float32 CPU tensor of hop * sum(seqlens) samples")
        id_type = _lib.IDS_I64 if ids.dtype == torch.int64 else _lib.IDS_I32
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device)
            _lib.check(lib.b200codec_decode_host_async(handle, ctypes.c_void_p(ids.data_ptr()), id_type,
                                                       _lib.i32_array(seqlens), len(seqlens),
                                                       ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream.cuda_stream)))
            event = torch.cuda.Event()
            event.record(stream)
        return event

    def take_id_error(self) -> bool:
        """True if a device-side decode since the last call saw an id outside [0, 65535]."""
        if self._handle is None:
            return False
        return bool(_lib.load().b200codec_take_id_error(self._handle))

    @torch.no_grad()
    def quantize_features(self, feats: torch.Tensor, *, pre_bound: bool,
                          return_projection: bool = False, id_dtype: torch.dtype = torch.int32):
        """Encode-direction FSQ with this checkpoint's `quantizer.project_in`: token-major
        features (n_tokens, 2048) fp32 on the device -> ids (n_tokens,) [+ the (n_tokens, 8) projected
        values]. ids are int32 like `FSQ.codes_to_indices` returns them (int64 on request).
        `pre_bound` has NO default: whether vector-quantize-pytorch 1.17.8 applies `FSQ.bound` to the
        projected input once more before the layer loop cannot be checked offline (parity unpinned at
        that library boundary), so the caller must say which release behaviour it wants.
        See `codec.encoder.FSQQuantizer`."""
        handle = self._ensure_handle()
        lib = _lib.load()
        if feats.device != self._device or feats.dim() != 2 or feats.dtype != torch.float32:
            raise ValueError("features must be a 2-D float32 tensor on the decoder's device")
        if feats.stride(1) != 1 or feats.stride(0) % 4 != 0 or feats.data_ptr() % 16 != 0:
            feats = feats.contiguous()
        n = feats.shape[0]
        with torch.cuda.device(self._device):
            if id_dtype not in (torch.int32, torch.int64):
                raise TypeError("id_dtype must be torch.int32 or torch.int64")
            ids = torch.empty(n, dtype=id_dtype, device=self._device)
            z = torch.empty(n, 8, dtype=torch.float32, device=self._device) if return_projection else None
            if n == 0:
                return (ids, z) if return_projection else ids
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(lib.b200codec_fsq_quantize(
                handle, ctypes.c_void_p(feats.data_ptr()), int(feats.stride(0)) if n > 1 else feats.shape[1], n,
                ctypes.c_void_p(ids.data_ptr()), _lib.IDS_I64 if id_dtype == torch.int64 else _lib.IDS_I32,
                ctypes.c_void_p(z.data_ptr()) if z is not None else None, 1 if pre_bound else 0,
                ctypes.c_void_p(stream)))
        return (ids, z) if return_projection else ids

    def plan_generation(self) -> int:
        """Changes whenever a decode rebuilt the handle's plan / workspace (a CUDA graph captured from a
        decode is only valid while this value stays what it was right after capture)."""
        return 0 if self._handle is None else int(_lib.load().b200codec_plan_generation(self._handle))

    def set_stage_taps(self, on: bool) -> None:
        """Debug: keep copies of named stage tensors of every decode (see b200codec.h)."""
        _lib.check(_lib.load().b200codec_set_stage_taps(self._ensure_handle(), 1 if on else 0))

    def read_stage(self, name: str) -> torch.Tensor:
        """Packed token-major fp32 (rows, width) copy of a tapped stage of the LAST decode."""
        lib = _lib.load()
        handle = self._ensure_handle()
        width = int(lib.b200codec_stage_width(handle, name.encode()))
        rows = int(lib.b200codec_stage_rows(handle, name.encode()))
        if width <= 0 or rows < 0:
            raise KeyError(f"no stage tap named {name!r} (enable set_stage_taps(True) and decode first)")
        out = torch.empty(rows, width, dtype=torch.float32)
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(lib.b200codec_read_stage(handle, name.encode(), ctypes.c_void_p(out.data_ptr()),
                                                out.numel(), ctypes.c_void_p(stream)))
        return out

    def launch_count(self) -> int:
        return 0 if self._handle is None else int(_lib.load().b200codec_launch_count(self._handle))

    def profile(self, on: bool) -> None:
        _lib.check(_lib.load().b200codec_profile(self._ensure_handle(), 1 if on else 0))

    def stage_times(self) -> "collections.OrderedDict[str, float]":
        lib = _lib.load()
        names = (ctypes.c_char_p * 64)()
        ms = (ctypes.c_float * 64)()
        n = ctypes.c_int(0)
        _lib.check(lib.b200codec_stage_times(self._ensure_handle(), 64, names, ms, ctypes.byref(n)))
        return collections.OrderedDict((names[i].decode(), float(ms[i])) for i in range(n.value))
