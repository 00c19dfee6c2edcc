"""CPU fp32 restatement of the reference ENCODE path up to the FSQ ids (TEST INFRASTRUCTURE, see
oracle/__init__.py): AcousticEncoder, SemanticEncoder, fusion layer, quantise. The w2v-BERT model that
produces the semantic features stays in HuggingFace; its hidden states are an input here.

One function per reference function, each citing the reference file:line (paths relative to the reference
repo root). State-dict keys are those of `tts.core.codec.encoder.Encoder.state_dict()` restricted to
`acoustic_encoder.*`, `semantic_encoder.*`, `fusion_layer.*` (and `quantizer.*`, which the decoder oracle
spells `decoder.quantizer.*`).
"""

from __future__ import annotations

import collections
import math
import zlib
from typing import Mapping

import torch
import torch.nn.functional as F

SD = Mapping[str, torch.Tensor]

UP_RATIOS = (2, 2, 4, 4, 5)
DILATIONS = (1, 3, 9)
GEN_FEATURES = 48
OUT_DIM = 1024


# ---------------------------------------------------------------------------------------------
# filters.py
# ---------------------------------------------------------------------------------------------
def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """tts/core/codec/filters.py:16-46 -> (kernel_size,) float32."""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = (torch.arange(-half_size, half_size) + 0.5) if even else (torch.arange(kernel_size) - half_size)
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return filt / filt.sum()


def upsample2(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """UpSample1d(ratio=2, kernel_size=12).forward (filters.py:87-113); x (B, C, T) -> (B, C, 2T)."""
    ratio, k = 2, filt.numel()
    pad = k // ratio - 1
    pad_left = pad * ratio + (k - ratio) // 2
    pad_right = pad * ratio + (k - ratio + 1) // 2
    C = x.shape[1]
    x = F.pad(x, (pad, pad), mode="replicate")
    x = ratio * F.conv_transpose1d(x, filt.view(1, 1, k).expand(C, -1, -1), stride=ratio, groups=C)
    return x[..., pad_left:-pad_right]


def downsample2(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """DownSample1d(ratio=2, kernel_size=12) -> LowPassFilter1d(stride=2).forward (filters.py:49-84, 116-135)."""
    k = filt.numel()
    even = k % 2 == 0
    pad_left, pad_right = k // 2 - int(even), k // 2
    C = x.shape[1]
    x = F.pad(x, (pad_left, pad_right), mode="replicate")
    return F.conv1d(x, filt.view(1, 1, k).expand(C, -1, -1), stride=2, groups=C)


# ---------------------------------------------------------------------------------------------
# activations.py
# ---------------------------------------------------------------------------------------------
def snake_beta(x: torch.Tensor, alpha: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """SnakeBeta(alpha_logscale=True).forward (activations.py:70-87): x + sin^2(x e^alpha) / (e^beta + 1e-9)."""
    a = torch.exp(alpha).view(1, -1, 1)
    b = torch.exp(beta).view(1, -1, 1)
    return x + (1.0 / (b + 1e-9)) * torch.pow(torch.sin(x * a), 2)


def activation1d(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Activation1d.forward (activations.py:90-110): 2x up (anti-imaging FIR) -> SnakeBeta -> 2x down."""
    up = sd[prefix + "upsample.filter"].reshape(-1)
    down = sd[prefix + "downsample.lowpass.filter"].reshape(-1)
    x = upsample2(x, up)
    x = snake_beta(x, sd[prefix + "act.alpha"], sd[prefix + "act.beta"])
    return downsample2(x, down)


# ---------------------------------------------------------------------------------------------
# encoder_modules.py
# ---------------------------------------------------------------------------------------------
def wn_weight(sd: SD, prefix: str) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, the norm over all dims but 0."""
    g, v = sd[prefix + "weight_g"], sd[prefix + "weight_v"]
    return g * v / v.reshape(v.shape[0], -1).norm(dim=1).view(-1, 1, 1)


def residual_unit(sd: SD, prefix: str, x: torch.Tensor, dilation: int) -> torch.Tensor:
    """ResidualUnit.forward (encoder_modules.py:20-43): x + conv1x1(act(conv7_dilated(act(x))))."""
    h = activation1d(sd, prefix + "block.0.", x)
    h = F.conv1d(h, wn_weight(sd, prefix + "block.1."), sd[prefix + "block.1.bias"], dilation=dilation,
                 padding=((7 - 1) * dilation) // 2)
    h = activation1d(sd, prefix + "block.2.", h)
    h = F.conv1d(h, wn_weight(sd, prefix + "block.3."), sd[prefix + "block.3.bias"])
    return x + h


def encoder_block(sd: SD, prefix: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """EncoderBlock.forward (encoder_modules.py:46-69): 3 ResidualUnits, act, Conv1d(k=2s, stride s)."""
    for i, d in enumerate(DILATIONS):
        x = residual_unit(sd, f"{prefix}block.{i}.", x, d)
    x = activation1d(sd, prefix + "block.3.", x)
    return F.conv1d(x, wn_weight(sd, prefix + "block.4."), sd[prefix + "block.4.bias"], stride=stride,
                    padding=stride // 2 + stride % 2)


def acoustic_encoder(sd: SD, wav: torch.Tensor, prefix: str = "acoustic_encoder.", stages: dict | None = None) -> torch.Tensor:
    """AcousticEncoder.forward (encoder_modules.py:187-191): wav (B, 1, S), S % 320 == 0 -> (B, S / 320, 1024)."""
    x = F.conv1d(wav, wn_weight(sd, prefix + "conv_blocks.0."), sd[prefix + "conv_blocks.0.bias"], padding=3)
    if stages is not None:
        stages["conv0"] = x
    for i, s in enumerate(UP_RATIOS):
        x = encoder_block(sd, f"{prefix}conv_blocks.{i + 1}.", x, s)
        if stages is not None:
            stages[f"block{i + 1}"] = x
    x = activation1d(sd, prefix + "conv_final_block.0.", x)
    x = F.conv1d(x, wn_weight(sd, prefix + "conv_final_block.1."), sd[prefix + "conv_final_block.1.bias"], padding=1)
    return x.permute(0, 2, 1)


def semantic_encoder(sd: SD, feats: torch.Tensor, prefix: str = "semantic_encoder.") -> torch.Tensor:
    """SemanticEncoder.forward (encoder_modules.py:121-125); feats (B, 1024, T) -> (B, 1024, T).
    `residual_blocks` starts with ReLU(inplace=True), which rewrites x before `+ x` is evaluated, so the
    skip connection carries relu(x) (encoder_modules.py:92-93, 123)."""
    x = F.conv1d(feats, sd[prefix + "initial_conv.weight"], None, padding=1)
    r = F.relu(x)
    h = F.conv1d(r, sd[prefix + "residual_blocks.1.weight"], sd[prefix + "residual_blocks.1.bias"], padding=1)
    h = F.conv1d(F.relu(h), sd[prefix + "residual_blocks.3.weight"], sd[prefix + "residual_blocks.3.bias"], padding=1)
    x = h + r
    return F.conv1d(x, sd[prefix + "final_conv.weight"], None, padding=1)


def quantize(sd: SD, hidden: torch.Tensor, pre_bound: bool):
    """Encoder.quantize (encoder.py:73-78) -> ids (B, 1, T) via the decoder oracle's ResidualFSQ restatement."""
    from oracle import codec_oracle as O

    qsd = {"decoder.quantizer.project_in.weight": sd["quantizer.project_in.weight"],
           "decoder.quantizer.project_in.bias": sd["quantizer.project_in.bias"]}
    ids, z, bounded = O.fsq_quantize(qsd, hidden.permute(0, 2, 1), pre_bound=pre_bound)
    return ids.unsqueeze(1), z, bounded


def encoder_hidden(sd: SD, wavs: torch.Tensor, w2v_hidden: torch.Tensor, stages: dict | None = None) -> torch.Tensor:
    """Encoder.forward up to the quantiser (tts/core/codec/encoder.py:58-71) with the w2v-BERT hidden state
    (`hidden_states[16]`, (B, T, 1024)) as an input: -> hidden_states (B, 2048, T)."""
    acoustic = acoustic_encoder(sd, wavs, stages=stages).transpose(1, 2)
    semantic = semantic_encoder(sd, w2v_hidden.transpose(1, 2))
    hidden = torch.cat([semantic, acoustic], dim=1)
    if stages is not None:
        stages["acoustic"], stages["semantic"] = acoustic, semantic
    return F.linear(hidden.transpose(1, 2), sd["fusion_layer.weight"], sd["fusion_layer.bias"]).transpose(1, 2)


# ---------------------------------------------------------------------------------------------
# deterministic weights
# ---------------------------------------------------------------------------------------------
def shapes() -> "collections.OrderedDict[str, tuple[int, ...]]":
    """Keys / shapes of Encoder.state_dict() for acoustic_encoder.*, semantic_encoder.*, fusion_layer.*
    in module order (encoder.py:28-44, encoder_modules.py:20-69, 72-119, 128-185; the Activation1d filters are
    registered buffers and therefore state-dict entries)."""
    sd: "collections.OrderedDict[str, tuple[int, ...]]" = collections.OrderedDict()

    def act(p: str, c: int) -> None:
        sd[p + "act.alpha"] = (c,)
        sd[p + "act.beta"] = (c,)
        sd[p + "upsample.filter"] = (1, 1, 12)
        sd[p + "downsample.lowpass.filter"] = (1, 1, 12)

    def wn_conv(p: str, cout: int, cin: int, k: int) -> None:
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
    wn_conv(a + "conv_blocks.0.", GEN_FEATURES, 1, 7)
    d = GEN_FEATURES
    for i, stride in enumerate(UP_RATIOS):
        p = f"{a}conv_blocks.{i + 1}."
        for u in range(3):
            q = f"{p}block.{u}."
            act(q + "block.0.", d)
            wn_conv(q + "block.1.", d, d, 7)
            act(q + "block.2.", d)
            wn_conv(q + "block.3.", d, d, 1)
        act(p + "block.3.", d)
        wn_conv(p + "block.4.", 2 * d, d, 2 * stride)
        d *= 2
    act(a + "conv_final_block.0.", d)
    wn_conv(a + "conv_final_block.1.", OUT_DIM, d, 3)
    sd["fusion_layer.weight"] = (2048, 2048)
    sd["fusion_layer.bias"] = (2048,)
    # ResidualFSQ: project_in / project_out are its only persistent tensors
    sd["quantizer.project_in.weight"] = (8, 2048)
    sd["quantizer.project_in.bias"] = (8,)
    sd["quantizer.project_out.weight"] = (2048, 8)
    sd["quantizer.project_out.bias"] = (2048,)
    return sd


def make_state_dict(seed: int = 0) -> "collections.OrderedDict[str, torch.Tensor]":
    """Deterministic test weights: every tensor from its own generator (seed, crc32(key)); alpha / beta and
    the biases are perturbed away from their zero initialisation so that they matter."""
    out: "collections.OrderedDict[str, torch.Tensor]" = collections.OrderedDict()
    filt = kaiser_sinc_filter1d(0.25, 0.3, 12).view(1, 1, 12)
    all_shapes = shapes()
    for key, shape in all_shapes.items():
        g = torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2 ** 31))
        if key.endswith("filter"):
            t = filt.clone()
        elif key.endswith("act.alpha") or key.endswith("act.beta"):
            t = 0.3 * torch.randn(shape, generator=g)
        elif key.endswith("weight_g"):
            # ||w|| per output channel = g. Inside the 15 residual units a gain near 1 would double the
            # variance per unit (x + block(x)); 0.3-0.5 keeps activations O(1) from the waveform to the
            # quantiser, like a trained codec, so that the stated dB tolerances mean something
            inside_unit = ".block.1." in key or ".block.3." in key
            t = (0.3 + 0.2 * torch.rand(shape, generator=g)) if inside_unit and "conv_blocks" in key and key.count(".block.") == 2 \
                else (0.9 + 0.2 * torch.rand(shape, generator=g))
        elif key.endswith("weight_v"):
            t = torch.randn(shape, generator=g) / math.sqrt(shape[1] * shape[2])
        elif key.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 3:   # semantic encoder Conv1d
            t = torch.randn(shape, generator=g) / math.sqrt(shape[1] * shape[2])
        else:                   # fusion / quantizer Linear weights
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) / math.sqrt(shape[-1] if len(shape) == 2 else 2048)
        out[key] = t.to(torch.float32).contiguous()
    return out
