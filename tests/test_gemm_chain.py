"""GPU: GEMM chains (c_proj -> fc1 -> fc2 -> next c_attn as one persistent launch with per-m-block
dependencies, csrc/gemm_tc05_2cta.cuh) against one launch per GEMM: the same arithmetic per element, so the
decode must come out the same -- at every batch shape that changes the tile walk (one m-block, a few, more
tiles than CTA pairs; narrow 256 x 64 tiles and wide 256 x 256 tiles), repeatedly (a dependency race would
show up as run-to-run differences)."""

import pytest
import torch

from tts_max_b200 import _lib

pytestmark = pytest.mark.gpu


def _decode(d, ids, seqlens, chain):
    _lib.check(_lib.load().b200codec_set_gemm_chain(1 if chain else 0))
    try:
        return d.decode_packed_device(ids, seqlens).clone()
    finally:
        _lib.check(_lib.load().b200codec_set_gemm_chain(0))


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("seqlens", [[7], [250], [250] * 4, [300, 1, 129, 64, 511], [500] * 16, [3000] * 2,
                                     [1000, 900, 800, 700, 600, 500, 400, 300, 200, 100] * 3])
def test_chain_equals_separate_launches(gpu_decoders, prec, seqlens):
    d = gpu_decoders[prec]
    g = torch.Generator().manual_seed(sum(seqlens) + len(seqlens))
    ids = torch.randint(0, 65536, (sum(seqlens),), generator=g).cuda()
    ref = _decode(d, ids, seqlens, chain=False)
    assert torch.isfinite(ref).all()
    scale = max(1e-3, ref.abs().max().item())
    for rep in range(3):
        got = _decode(d, ids, seqlens, chain=True)
        # identical GEMM arithmetic; only the fp64 GroupNorm atomics may reorder
        assert (got - ref).abs().max().item() <= 1e-6 * scale, (rep, seqlens[:4])


def test_chain_cuts_launches(gpu_decoders):
    """81 launches per decode by default (one per GEMM); 47 with chains (12 chains replace 46 GEMM launches)."""
    d = gpu_decoders["bf16"]
    ids = torch.randint(0, 65536, (500,), generator=torch.Generator().manual_seed(1)).cuda()
    n0 = d.launch_count()
    d.decode_packed_device(ids, [500])
    default = d.launch_count() - n0
    _decode(d, ids, [500], chain=True)
    with_chain = d.launch_count() - n0 - default
    assert with_chain == 47 and default == 81, (with_chain, default)
