"""TEST-ONLY shim for `vector_quantize_pytorch` (pinned 1.17.8 in the reference's uv.lock:4889-4890).

The wheel is not installed here and not vendored in /root/reference. This restates, from the
library's published source, exactly the part of `ResidualFSQ` the reference decoder touches
(construction at decoder_modules.py:418-420, `get_output_from_indices` at decoder.py:77) and the
reference encoder's quantise step (`forward`, called at encoder.py:76), so the UNMODIFIED
reference modules can be imported by oracle/make_golden*.py. "parity unpinned": this
file is a restatement, not the upstream code.
"""

import torch


class ResidualFSQ(torch.nn.Module):
    def __init__(self, *, dim, levels, num_quantizers, **kwargs):
        super().__init__()
        assert num_quantizers == 1, "shim covers the single-quantizer configuration only"
        codebook_dim = len(levels)
        # persistent state of the real class: project_in / project_out only
        self.project_in = torch.nn.Linear(dim, codebook_dim)
        self.project_out = torch.nn.Linear(codebook_dim, dim)
        levels_t = torch.tensor(levels, dtype=torch.int32)
        self.register_buffer("_levels", levels_t, persistent=False)
        basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0, dtype=torch.int32)
        self.register_buffer("_basis", basis, persistent=False)
        # scales = (levels - 1) ** -q for quantizer q; q = 0 -> 1
        self.register_buffer("scales", torch.ones(1, codebook_dim), persistent=False)
        n = int(torch.prod(levels_t.long()))
        codebook = self._indices_to_codes(torch.arange(n))
        self.register_buffer("implicit_codebook", codebook, persistent=False)

    def _indices_to_codes(self, indices):
        level_indices = (indices.unsqueeze(-1) // self._basis) % self._levels
        half_width = self._levels // 2
        return (level_indices - half_width) / half_width

    def get_codes_from_indices(self, indices):
        # indices [b, n, q]; -1 marks a dropped-out quantizer -> zero code
        mask = indices == -1
        idx = indices.masked_fill(mask, 0)
        codes = self.implicit_codebook[idx[..., 0]]            # [b, n, d]
        codes = codes.masked_fill(mask[..., :1], 0.0)
        return (codes * self.scales).unsqueeze(0)              # [q, b, n, d]

    def get_output_from_indices(self, indices):
        codes = self.get_codes_from_indices(indices)
        return self.project_out(codes.sum(dim=0))

    # ---- encode direction (Encoder.quantize, encoder.py:73-78) ----
    pre_bound = False  # class switch: `residual = layers[0].bound(x)` before the layer loop

    def _bound(self, z, eps=1e-3):
        half_l = (self._levels - 1) * (1 + eps) / 2
        offset = torch.where(self._levels % 2 == 0, 0.5, 0.0)
        shift = (offset / half_l).atanh()
        return (z + shift).tanh() * half_l - offset

    def forward(self, x):
        """x [b, n, dim] -> (quantized_out [b, n, dim], indices [b, n, 1] int32)."""
        z = self.project_in(x)
        residual = self._bound(z) if self.pre_bound else z
        half_width = self._levels // 2
        codes = self._bound(residual / self.scales).round() / half_width       # FSQ.quantize
        indices = ((codes * half_width + half_width) * self._basis).sum(dim=-1).to(torch.int32)
        quantized_out = self.project_out(codes * self.scales)
        return quantized_out, indices.unsqueeze(-1)
