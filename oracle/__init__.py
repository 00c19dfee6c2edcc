"""ORACLE -- TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's codec-decode algorithm (tts-max,
`tts/core/codec/{decoding,decoder,decoder_modules}.py`), used as the checker in `tests/`,
`__graft_entry__.smoke()` and as the CPU baseline leg of `bench.py`. Nothing under
`tts_max_b200/` may import, call, link or execute anything in this directory: the product
path is the CUDA library and fails loudly when it is missing.

PARITY STATUS: "parity unpinned" at the third-party boundary. The reference ships no tests,
golden vectors or fixtures for this path, and two classes on the path --
`vector_quantize_pytorch.ResidualFSQ` (1.17.8) and
`torchtune.modules.RotaryPositionalEmbeddings` (0.6.1) -- are neither vendored in the
reference nor installed here; their published algorithms are restated in
`oracle/ref_shims/`. Everything else is pinned against the reference's own module code:
`oracle/make_golden.py` imports the unmodified reference from /root/reference (with those
two shims), loads our deterministic weights through the reference's own
`load_from_checkpoint` in both checkpoint layouts and stores its outputs under
`tests/golden/`; `tests/test_oracle.py` checks this restatement against those vectors.
"""
