"""Independent numpy second opinions for the two HBM-bound stages (TEST INFRASTRUCTURE).

They share no code with codec_oracle.py / torch: explicit loops and np.fft only.
"""

from __future__ import annotations

import numpy as np


def fsq_lookup_np(ids: np.ndarray, w_out: np.ndarray, b_out: np.ndarray) -> np.ndarray:
    """K1 in the exact fp32 evaluation order the CUDA kernel promises:
    acc = 0; for d in 0..7: acc = fl(acc + code_d * W[c, d]); out = fl(acc + b[c]).
    ids (n,), w_out (C, 8) f32, b_out (C,) f32 -> (n, C) f32."""
    ids = np.asarray(ids).astype(np.int64)
    w = np.asarray(w_out, dtype=np.float32)
    b = np.asarray(b_out, dtype=np.float32)
    out = np.zeros((ids.shape[0], w.shape[0]), dtype=np.float32)
    acc = np.zeros_like(out)
    for d in range(8):
        digit = (ids // (4 ** d)) % 4
        code = ((digit - 2) / 2).astype(np.float32)  # {-1, -0.5, 0, 0.5}: products are exact
        acc = (acc + code[:, None] * w[None, :, d]).astype(np.float32)
    out[:] = (acc + b[None, :]).astype(np.float32)
    return out


def istft_same_np(x_pred: np.ndarray, window: np.ndarray, hop: int) -> np.ndarray:
    """K13 + K14 for one utterance in float64. x_pred (T, n_fft + 2): columns [0, n_bins) are
    log-magnitudes, [n_bins, 2 n_bins) phases (decoder_modules.py:131-146, 59-93)."""
    x_pred = np.asarray(x_pred, dtype=np.float64)
    w = np.asarray(window, dtype=np.float64)
    n_fft = w.shape[0]
    n_bins = n_fft // 2 + 1
    T = x_pred.shape[0]
    mag = np.minimum(np.exp(x_pred[:, :n_bins]), 100.0)
    ph = x_pred[:, n_bins:2 * n_bins]
    spec = mag * (np.cos(ph) + 1j * np.sin(ph))
    frames = np.fft.irfft(spec, n=n_fft, axis=1) * w[None, :]
    total = (T - 1) * hop + n_fft
    y = np.zeros(total)
    env = np.zeros(total)
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[t]
        env[t * hop:t * hop + n_fft] += w * w
    pad = (n_fft - hop) // 2
    return (y[pad:total - pad] / env[pad:total - pad]).astype(np.float32)
