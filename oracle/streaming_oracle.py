"""CPU fp32 restatement of the CACHED streaming decode (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference has no streaming decoder; this states, with the reference's own layer functions
(oracle/codec_oracle.py, each citing its reference file:line), exactly the algorithm of
`b200codec_stream_push` (include/b200codec.h): a push runs the model on [overlap | new] rows, and attention
(decoder_modules.py:275-290) reads the keys / values of the last `cap` tokens as they were computed when those
tokens were new. Everything else -- convolutions, GroupNorm, RMSNorm, MLP, LayerNorm, head, ISTFT -- sees the
rows of the push only.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import codec_oracle as O


class CachedStreamOracle:
    def __init__(self, sd, n_streams: int, new_tokens: int, left_context: int, depth: int = O.DEPTH):
        self.sd, self.n, self.depth = sd, new_tokens, depth
        self.cap = (left_context + new_tokens - 1) // new_tokens * new_tokens + new_tokens
        self.k = [None] * depth   # (B, H, t <= cap, d) per layer, oldest first
        self.v = [None] * depth
        self.n_streams = n_streams

    @torch.no_grad()
    def push(self, ids: torch.Tensor, overlap: int) -> torch.Tensor:
        """ids (B, overlap + new) -> (B, new * 320): the audio of the new tokens."""
        sd = self.sd
        B, T = ids.shape
        assert T == overlap + self.n
        C, H = 1024, O.HEADS
        x = F.linear(O.fsq_lookup(sd, ids), sd["fc_post_a.weight"], sd["fc_post_a.bias"])
        p = "decoder.backbone."
        x = x.transpose(1, 2)
        x = F.conv1d(x, sd[p + "embed.weight"], sd[p + "embed.bias"], padding=3)
        x = O.resnet_block(sd, p + "prior_net.0.", x)
        x = O.resnet_block(sd, p + "prior_net.1.", x)
        x = x.transpose(1, 2)
        for layer in range(self.depth):
            q_ = f"{p}transformers.{layer}."
            hn = O.rms_norm(x, sd[q_ + "att_norm.weight"])
            qkv = F.linear(hn, sd[q_ + "att.c_attn.weight"]).view(B, T, 3, H, C // H).permute(2, 0, 3, 1, 4)
            q, k, v = O.rope_torchtune(qkv[0]), O.rope_torchtune(qkv[1]), qkv[2]
            k_new, v_new = k[:, :, overlap:], v[:, :, overlap:]
            self.k[layer] = k_new if self.k[layer] is None else torch.cat([self.k[layer], k_new], dim=2)[:, :, -self.cap:]
            self.v[layer] = v_new if self.v[layer] is None else torch.cat([self.v[layer], v_new], dim=2)[:, :, -self.cap:]
            y = F.scaled_dot_product_attention(q, self.k[layer], self.v[layer], attn_mask=None, dropout_p=0, is_causal=False)
            y = y.permute(0, 2, 1, 3).reshape(B, T, C)
            x = x + F.linear(y, sd[q_ + "att.c_proj.weight"])
            x = x + O.mlp(sd, q_ + "mlp.", O.rms_norm(x, sd[q_ + "ffn_norm.weight"]))
        x = x.transpose(1, 2)
        x = O.resnet_block(sd, p + "post_net.0.", x)
        x = O.resnet_block(sd, p + "post_net.1.", x)
        x = x.transpose(1, 2)
        x = F.layer_norm(x, (C,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], O.EPS)
        wav = O.istft_head(sd, x, 320)[:, 0]
        return wav[:, overlap * 320:]
