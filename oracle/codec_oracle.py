"""CPU fp32 restatement of the reference decode path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional PyTorch-on-CPU code, one function per reference function, each citing the reference
file:line it follows (paths relative to the reference repo root). The state dict uses the keys
of `Decoder.state_dict()`.
"""

from __future__ import annotations

import math
from typing import Mapping

import torch
import torch.nn.functional as F

SD = Mapping[str, torch.Tensor]

HEADS = 16
DEPTH = 12
GROUPS = 32
EPS = 1e-6


# ---------------------------------------------------------------------------------------------
# K1: vector_quantize_pytorch.ResidualFSQ(dim=2048, levels=[4]*8, num_quantizers=1)
#     .get_output_from_indices   (library 1.17.8, restated; built at decoder_modules.py:418-420,
#     called at decoder.py:77)
# ---------------------------------------------------------------------------------------------
def fsq_codes(ids: torch.Tensor) -> torch.Tensor:
    """ids (...,) int -> codes (..., 8) float32: digit_d = (id // 4^d) % 4; code = (digit - 2) / 2."""
    basis = torch.tensor([4 ** d for d in range(8)], dtype=torch.int64)
    digits = (ids.to(torch.int64).unsqueeze(-1) // basis) % 4
    half_width = 2  # levels // 2
    return (digits.to(torch.float32) - half_width) / half_width


def fsq_bound(z: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
    """FSQ.bound (vector-quantize-pytorch 1.17.8, finite_scalar_quantization.py) for levels [4]*8:
    half_l = (levels - 1) * (1 + eps) / 2; offset = 0.5 (even levels); shift = atanh(offset / half_l);
    tanh(z + shift) * half_l - offset."""
    levels = torch.full((8,), 4, dtype=torch.int32)
    half_l = (levels - 1) * (1 + eps) / 2
    offset = torch.where(levels % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    return (z + shift).tanh() * half_l - offset


def fsq_quantize(sd: SD, feats: torch.Tensor, pre_bound: bool = False):
    """ResidualFSQ.forward as called by Encoder.quantize (tts/core/codec/encoder.py:73-78), one
    quantizer: feats (..., 2048) -> (ids int64 (...,), z (..., 8), bounded (..., 8)).
        z = project_in(x); codes = round(bound(z)) / half_width;
        ids = sum((codes * half_width + half_width) * basis), basis = cumprod([1, 4, ...]) (int32)
    `pre_bound`: some releases run `residual = layers[0].bound(x)` before the layer loop (the
    wheel is absent here, so which one 1.17.8 does is unpinned); the flag selects that variant.
    "parity unpinned": restated from the library's published source, not executed upstream code."""
    z = F.linear(feats, sd["decoder.quantizer.project_in.weight"], sd["decoder.quantizer.project_in.bias"])
    r = fsq_bound(z) if pre_bound else z
    bounded = fsq_bound(r)
    half_width = 2
    codes = bounded.round() / half_width                      # round_ste forward value
    basis = torch.cumprod(torch.tensor([1] + [4] * 7), dim=0).to(torch.int32)
    ids = ((codes * half_width + half_width) * basis).sum(dim=-1).to(torch.int32)
    return ids.to(torch.int64), z, bounded


def fsq_lookup(sd: SD, ids: torch.Tensor) -> torch.Tensor:
    """ids (B, T) -> (B, T, 2048): project_out(codes) (scales == 1 for the single quantizer)."""
    codes = fsq_codes(ids)
    return F.linear(codes, sd["decoder.quantizer.project_out.weight"], sd["decoder.quantizer.project_out.bias"])


# ---------------------------------------------------------------------------------------------
# decoder_modules.py:151-223  nonlinearity / Normalize / ResnetBlock (temb_channels=0, eval mode)
# ---------------------------------------------------------------------------------------------
def swish(x: torch.Tensor) -> torch.Tensor:  # decoder_modules.py:151-153
    return x * torch.sigmoid(x)


def resnet_block(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """x (B, C, T). decoder_modules.py:201-223; dropout is inactive in eval (decoding.py:79)."""
    h = F.group_norm(x, GROUPS, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], EPS)
    h = swish(h)
    h = F.conv1d(h, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1)
    h = F.group_norm(h, GROUPS, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], EPS)
    h = swish(h)
    h = F.conv1d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], padding=1)
    return x + h


# ---------------------------------------------------------------------------------------------
# decoder_modules.py:226-314  RMSNorm / MLP / Attention / TransformerBlock
# ---------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:  # decoder_modules.py:233-236
    norm_x = torch.mean(x ** 2, dim=-1, keepdim=True)
    return x * torch.rsqrt(norm_x + EPS) * weight


def rope_cache(max_seq_len: int = 4096, dim: int = 64, base: int = 10000) -> torch.Tensor:
    """torchtune.modules.RotaryPositionalEmbeddings._rope_init/build_rope_cache (0.6.1, restated)."""
    theta = 1.0 / (base ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    seq_idx = torch.arange(max_seq_len, dtype=theta.dtype)
    idx_theta = torch.einsum("i, j -> ij", seq_idx, theta).float()
    return torch.stack([torch.cos(idx_theta), torch.sin(idx_theta)], dim=-1)  # [S, dim/2, 2]


_ROPE_CACHE = None


def rope_torchtune(x: torch.Tensor) -> torch.Tensor:
    """torchtune RotaryPositionalEmbeddings.forward(x) with input_pos=None. The library contract is
    x = [b, s, n_h, h_d]; the reference passes [b, h, t, d] (decoder_modules.py:276-281), so the
    "position" is the HEAD index and the rotation is constant over time."""
    global _ROPE_CACHE
    if _ROPE_CACHE is None:
        _ROPE_CACHE = rope_cache()
    seq_len = x.size(1)
    rc = _ROPE_CACHE[:seq_len]
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    rc = rc.view(-1, xs.size(1), 1, xs.size(3), 2)
    out = torch.stack(
        [xs[..., 0] * rc[..., 0] - xs[..., 1] * rc[..., 1], xs[..., 1] * rc[..., 0] + xs[..., 0] * rc[..., 1]], -1
    )
    return out.flatten(3).type_as(x)


def attention(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """x (B, T, C). decoder_modules.py:275-290."""
    B, T, C = x.shape
    qkv = F.linear(x, sd[prefix + "c_attn.weight"])
    # einops "b t (r h d) -> r b h t d"
    qkv = qkv.view(B, T, 3, HEADS, C // HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = rope_torchtune(q)
    k = rope_torchtune(k)
    y = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0, is_causal=False)
    y = y.permute(0, 2, 1, 3).reshape(B, T, C)  # "b h t d -> b t (h d)"
    return F.linear(y, sd[prefix + "c_proj.weight"])


def mlp(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:  # decoder_modules.py:247-251
    return F.linear(F.silu(F.linear(x, sd[prefix + "fc1.weight"])), sd[prefix + "fc2.weight"])


def transformer_block(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:  # decoder_modules.py:311-314
    x = x + attention(sd, prefix + "att.", rms_norm(x, sd[prefix + "att_norm.weight"]))
    x = x + mlp(sd, prefix + "mlp.", rms_norm(x, sd[prefix + "ffn_norm.weight"]))
    return x


# ---------------------------------------------------------------------------------------------
# decoder_modules.py:390-400  VocosBackbone.forward
# ---------------------------------------------------------------------------------------------
def backbone(sd: SD, x: torch.Tensor, depth: int = DEPTH) -> torch.Tensor:
    """x (B, T, C) -> (B, T, C)."""
    p = "decoder.backbone."
    x = x.transpose(1, 2)
    x = F.conv1d(x, sd[p + "embed.weight"], sd[p + "embed.bias"], padding=3)
    x = resnet_block(sd, p + "prior_net.0.", x)
    x = resnet_block(sd, p + "prior_net.1.", x)
    x = x.transpose(1, 2)
    for layer in range(depth):
        x = transformer_block(sd, f"{p}transformers.{layer}.", x)
    x = x.transpose(1, 2)
    x = resnet_block(sd, p + "post_net.0.", x)
    x = resnet_block(sd, p + "post_net.1.", x)
    x = x.transpose(1, 2)
    return F.layer_norm(x, (x.shape[-1],), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], EPS)


# ---------------------------------------------------------------------------------------------
# decoder_modules.py:35-93  ISTFT.forward (padding == "same");  :118-148 ISTFTHead.forward
# ---------------------------------------------------------------------------------------------
def istft_same(spec: torch.Tensor, window: torch.Tensor, hop: int) -> torch.Tensor:
    """spec (B, N/2+1, T) complex -> (B, hop * T)."""
    n_fft = window.numel()
    pad = (n_fft - hop) // 2
    B, N, T = spec.shape
    ifft = torch.fft.irfft(spec, n_fft, dim=1, norm="backward")
    ifft = ifft * window[None, :, None]
    output_size = (T - 1) * hop + n_fft
    y = F.fold(ifft, output_size=(1, output_size), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, pad:-pad]
    window_sq = window.square().expand(1, T, -1).transpose(1, 2)
    env = F.fold(window_sq, output_size=(1, output_size), kernel_size=(1, n_fft), stride=(1, hop)).squeeze()[pad:-pad]
    assert (env > 1e-11).all()
    return y / env


def head_spectrum(x_pred: torch.Tensor) -> torch.Tensor:
    """x_pred (B, T, n_fft + 2) -> complex spectrum (B, n_fft/2 + 1, T). decoder_modules.py:131-146."""
    x_pred = x_pred.transpose(1, 2)
    mag, p = x_pred.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    return mag * (torch.cos(p) + 1j * torch.sin(p))


def istft_head(sd: SD, x: torch.Tensor, hop: int) -> torch.Tensor:
    """x (B, T, C) -> (B, 1, hop * T)."""
    x_pred = F.linear(x, sd["decoder.head.out.weight"], sd["decoder.head.out.bias"])
    audio = istft_same(head_spectrum(x_pred), sd["decoder.head.istft.window"], hop)
    return audio.unsqueeze(1)


# ---------------------------------------------------------------------------------------------
# upsampler.py:9-69  UpSamplerBlock (48 kHz variant): weight-normed ConvTranspose1d + ResnetBlock per
# factor, then out_proj + swish. ResnetBlock is built with the default temb_channels=512, so its
# state dict carries an unused temb_proj (forward is called with temb=None, decoder_modules.py:209).
# ---------------------------------------------------------------------------------------------
def weight_norm_weight(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, the norm taken over all dims but 0."""
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def upsampler(sd: SD, x: torch.Tensor, factors, kernels) -> torch.Tensor:
    """x (B, C, T) -> (B, T * prod(factors), C). upsampler.py:62-69."""
    for i, (k, u) in enumerate(zip(kernels, factors)):
        p = f"upsampler.upsample_layers.{i}."
        w = weight_norm_weight(sd[p + "weight_g"], sd[p + "weight_v"])
        x = F.conv_transpose1d(x, w, sd[p + "bias"], stride=u, padding=(k - u) // 2)
        x = resnet_block(sd, f"upsampler.resnet_blocks.{i}.", x)
    x = F.linear(x.transpose(1, 2), sd["upsampler.out_proj.weight"], sd["upsampler.out_proj.bias"])
    return swish(x)


# ---------------------------------------------------------------------------------------------
# decoder.py:69-89  Decoder.forward
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def decoder_forward(sd: SD, vq_codes: torch.Tensor, hop: int = 320, depth: int = DEPTH, stages: dict | None = None,
                    upsample_factors=None, kernel_sizes=None) -> torch.Tensor:
    """vq_codes (B, T) or (B, 1, T) -> (B, 1, hop * prod(upsample_factors) * T) float32. `stages`, if
    given, receives the intermediate tensors (K1 output, fc_post_a output, backbone output, upsampler
    output, head Linear output)."""
    if vq_codes.dim() == 2:
        vq_codes = vq_codes.unsqueeze(1)
    ids = vq_codes.transpose(1, 2)[..., 0]  # (B, T)
    emb = fsq_lookup(sd, ids)
    x = F.linear(emb, sd["fc_post_a.weight"], sd["fc_post_a.bias"])
    hidden = backbone(sd, x, depth)
    up = upsampler(sd, hidden.transpose(1, 2), upsample_factors, kernel_sizes) if upsample_factors else hidden
    if stages is not None:
        stages["fsq"] = emb
        stages["fc_post_a"] = x
        stages["backbone"] = hidden
        stages["upsampled"] = up
        stages["head_linear"] = F.linear(up, sd["decoder.head.out.weight"], sd["decoder.head.out.bias"])
    return istft_head(sd, up, hop)


def snr_db(ref: torch.Tensor, test: torch.Tensor) -> float:
    ref = ref.double().flatten()
    err = (test.double().flatten() - ref)
    den = float((err ** 2).sum())
    if den == 0.0:
        return math.inf
    return 10.0 * math.log10(float((ref ** 2).sum()) / den)
