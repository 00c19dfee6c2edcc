"""Deterministic test weights for the codec decoder (TEST INFRASTRUCTURE).

`make_state_dict(seed)` returns a `Decoder.state_dict()`-keyed dict of fp32 CPU tensors drawn
with the reference's init distributions (Conv1d trunc_normal(std=0.02) + zero bias,
decoder_modules.py:13-16; Linear = torch default; `window` = hann). With `perturb=True` the
tensors the reference initialises to constants (norm weights / biases, conv biases) are
randomised too, so a kernel that ignores them cannot pass, and the head's log-magnitude bias is
shifted so magnitudes are not all ~1. Each tensor is seeded from its own key, so the result
does not depend on construction order.
"""

from __future__ import annotations

import collections
import math
import zlib

import torch

C, V, DEPTH = 1024, 2048, 12


def shapes(hop: int = 320, depth: int = DEPTH, upsample_factors=None, kernel_sizes=None):
    n_fft = 4 * hop
    sd = collections.OrderedDict()
    g = "decoder."
    sd[g + "quantizer.project_in.weight"] = (8, V)
    sd[g + "quantizer.project_in.bias"] = (8,)
    sd[g + "quantizer.project_out.weight"] = (V, 8)
    sd[g + "quantizer.project_out.bias"] = (V,)
    sd[g + "backbone.embed.weight"] = (C, C, 7)
    sd[g + "backbone.embed.bias"] = (C,)

    def resnet(p):
        for n in ("1", "2"):
            sd[f"{p}norm{n}.weight"] = (C,)
            sd[f"{p}norm{n}.bias"] = (C,)
            sd[f"{p}conv{n}.weight"] = (C, C, 3)
            sd[f"{p}conv{n}.bias"] = (C,)

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
        # registration order of UpSamplerBlock.__init__ (upsampler.py:27-60): upsample_layers,
        # resnet_blocks, out_proj; Decoder registers `upsampler` before `fc_post_a` (decoder.py:48-63)
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
            sd[p + "temb_proj.weight"] = (c, 512)
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


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2 ** 31))


def make_state_dict(seed: int = 0, perturb: bool = True, hop: int = 320, depth: int = DEPTH,
                    upsample_factors=None, kernel_sizes=None):
    all_shapes = shapes(hop, depth, upsample_factors, kernel_sizes)
    out = collections.OrderedDict()
    for key, shape in all_shapes.items():
        g = _gen(seed, key)
        if key.endswith("istft.window"):
            t = torch.hann_window(shape[0])
        elif key.endswith("weight_g"):
            # weight_norm initialises g to ||v||; perturbed so that g actually matters
            t = 0.35 + 0.1 * torch.rand(shape, generator=g)
        elif key.endswith("weight_v"):
            fan = shape[1] * shape[2]
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) / math.sqrt(fan)
        elif "norm" in key:
            if key.endswith("weight"):
                t = 1.0 + 0.2 * torch.randn(shape, generator=g) if perturb else torch.ones(shape)
            else:
                t = 0.1 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        elif len(shape) == 3:
            t = torch.empty(shape)
            torch.nn.init.trunc_normal_(t, std=0.02, generator=g)
        elif key.endswith("bias") and ("embed" in key or "conv" in key or "upsample_layers" in key):
            t = 0.05 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        else:
            wshape = shape if len(shape) == 2 else all_shapes[key[: -len("bias")] + "weight"]
            bound = 1.0 / math.sqrt(wshape[1])
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
            if perturb and key == "decoder.head.out.bias":
                t = t.clone()
                t[: shape[0] // 2] -= 2.0  # log-magnitudes around -2: spectrum magnitudes ~0.1
        out[key] = t.to(torch.float32).contiguous()
    return out


def fingerprint(sd) -> float:
    """Order-independent float64 checksum used to detect RNG drift between torch builds."""
    acc = 0.0
    for k, v in sd.items():
        acc += float(v.double().sum()) + 1e-3 * float(v.double().abs().sum())
    return acc


def to_xcodec2_checkpoint(sd) -> dict:
    """{"state_dict": {"generator.*", "fc_post_a.*", + keys the decoder must ignore}} (decoder.py:94-110)."""
    st = collections.OrderedDict()
    for k, v in sd.items():
        if k.startswith("decoder."):
            st["generator." + k[len("decoder."):]] = v
        else:
            st[k] = v
    st["CodecEnc.conv_blocks.0.weight"] = torch.zeros(3)        # ignored by the decoder
    st["fc_prior.weight"] = torch.zeros(2, 2)                    # ignored by the decoder
    return {"state_dict": st}


def to_ttsmax_checkpoint(sd) -> dict:
    """{"model": {"generator.<Decoder key>", + discriminator keys}} (decoder.py:112-119, checkpointing.py:46-52)."""
    st = collections.OrderedDict(("generator." + k, v) for k, v in sd.items())
    st["discriminator.mpd.0.weight"] = torch.zeros(3)            # filtered out by the prefix test
    return {"model": st, "optimizer": {}, "config": {}}
