"""Micro-benchmark of the tcgen05 GEMM kernel through the C ABI vs torch.matmul (cuBLAS) on the
decode path's shapes. CUDA events, L2 flushed between iterations."""
import ctypes
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_max_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def timeit(fn, iters=20, do_flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


clock_rows = []
stop = False


def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, text=True)
    while not stop:
        line = p.stdout.readline()
        if line:
            clock_rows.append(line.strip())
    p.kill()


th = threading.Thread(target=sampler, daemon=True)
th.start()

shapes = [  # (name, M, N, K, taps, out_fp32, residual, act)
    ("qkv", 8045, 3072, 1024, 1, False, False, 0),
    ("proj+res", 8045, 1024, 1024, 1, True, True, 0),
    ("fc1+silu", 8045, 4096, 1024, 1, False, False, 1),
    ("fc2+res", 8045, 1024, 4096, 1, True, True, 0),
    ("conv3", 8045, 1024, 1024, 3, True, False, 0),
    ("conv7", 8045, 1024, 1024, 7, True, False, 0),
    ("big-K", 8192, 4096, 8192, 1, False, False, 0),
    ("square", 8192, 8192, 8192, 1, False, False, 0),
    ("qkv-16k", 16384, 3072, 1024, 1, False, False, 0),
]
print(f"{'name':10s} {'M':>6s} {'N':>5s} {'K':>5s} {'ours us':>9s} {'TF/s':>7s} {'cublas us':>9s} {'TF/s':>7s}")
for name, M, N, K, taps, out_fp32, res, act in shapes:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K * taps, device=dev) * 0.02).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    resid = torch.randn(M, N, device=dev) if res else None
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def ours():
        _lib.check(lib.b200codec_gemm(0, ptr(a), ptr(w), M, N, K, taps, ptr(out), 0 if out_fp32 else 1, N, None,
                                      ptr(resid), N if res else 0, act, s))

    a2 = torch.randn(M, K * taps, device=dev).bfloat16()
    wt = w.t().contiguous()

    def cublas():
        torch.matmul(a2, w.t())

    t1 = timeit(ours)
    t2 = timeit(cublas)
    fl = 2.0 * M * N * K * taps
    print(f"{name:10s} {M:6d} {N:5d} {K * taps:5d} {t1 * 1e3:9.1f} {fl / t1 / 1e9:7.0f} {t2 * 1e3:9.1f} {fl / t2 / 1e9:7.0f}")

# sustained: run the qkv GEMM back to back for ~2 s and report clocks
a = torch.randn(8045, 1024, device=dev).bfloat16()
w = (torch.randn(3072, 1024, device=dev) * 0.02).bfloat16()
out = torch.empty(8045, 3072, device=dev, dtype=torch.bfloat16)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
n0 = len(clock_rows)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 20000
for _ in range(iters):
    lib.b200codec_gemm(0, ptr(a), ptr(w), 8045, 3072, 1024, 1, ptr(out), 1, 3072, None, None, 0, 0, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"sustained qkv x{iters}: {ms / iters * 1e3:.1f} us each, {2.0 * 8045 * 3072 * 1024 * iters / ms / 1e9:.0f} TF/s")
print("clocks during sustained:", clock_rows[n0 + 2:][:: max(1, (len(clock_rows) - n0) // 10)])
stop = True
