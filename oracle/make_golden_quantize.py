"""Generates tests/golden/reference_quantize_seed0.npz (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):   python oracle/make_golden_quantize.py

Runs the UNMODIFIED reference `Encoder.quantize` (tts/core/codec/encoder.py:73-78) -- unbound, on a
stand-in `self` that carries only `.quantizer`, because `Encoder.__init__` downloads w2v-BERT --
over the oracle/ref_shims restatement of `ResidualFSQ` (the wheel is absent: "parity unpinned"),
with the deterministic test weights' `quantizer.project_in`. Stores the features, the ids for
both `pre_bound` variants, and the projected / bounded values so a test can tell a genuine
mismatch from a value that sits on a rounding boundary.
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")

from oracle import codec_oracle, weights  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main() -> None:
    import vector_quantize_pytorch as vq  # the shim
    from tts.core.codec import encoder as ref_encoder  # the reference, unmodified

    torch.manual_seed(0)
    sd = weights.make_state_dict(seed=0, perturb=True)
    q = vq.ResidualFSQ(dim=2048, levels=[4] * 8, num_quantizers=1)
    q.load_state_dict({k[len("decoder.quantizer."):]: v for k, v in sd.items() if k.startswith("decoder.quantizer.")})
    g = torch.Generator().manual_seed(4321)
    b, t = 3, 41
    # spread the projected values over all four levels: project_in has default-init scale ~1/sqrt(2048)
    hidden = torch.randn(b, 2048, t, generator=g) * 2.0
    out = {"hidden": hidden.numpy()}
    for pre in (False, True):
        vq.ResidualFSQ.pre_bound = pre
        stand_in = types.SimpleNamespace(quantizer=q)
        with torch.no_grad():
            code = ref_encoder.Encoder.quantize(stand_in, hidden)      # (B, 1, T) int32
        assert code.shape == (b, 1, t)
        ids, z, bounded = codec_oracle.fsq_quantize(sd, hidden.permute(0, 2, 1), pre_bound=pre)
        assert torch.equal(ids, code[:, 0, :].to(torch.int64)), "oracle restatement != reference path over the shim"
        out[f"ids_pre{int(pre)}"] = ids.numpy()
        out[f"bounded_pre{int(pre)}"] = bounded.numpy()
        out["z"] = z.numpy()
        digits = [(ids // 4 ** d) % 4 for d in range(8)]
        print("pre_bound", pre, "digit histogram", torch.bincount(torch.stack(digits).flatten(), minlength=4).tolist())
    vq.ResidualFSQ.pre_bound = False
    np.savez_compressed(os.path.join(GOLDEN, "reference_quantize_seed0.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "reference_quantize_seed0.npz"))


if __name__ == "__main__":
    main()
