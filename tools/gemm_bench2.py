"""Epilogue variants of the N=1024,K=1024 GEMM (c_proj shape), warm and cold L2."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_max_b200 import _lib
lib = _lib.load()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def ptr(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else None
def timeit(fn, iters=20, do_flush=True):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print(f"{'variant':28s} {'M':>6s} {'N':>5s} {'K':>5s} {'cold us':>8s} {'warm us':>8s} {'warm TF/s':>9s}")
for M in (8045, 2048, 512, 16384):
    for N, K in ((1024, 1024), (1024, 4096), (3072, 1024)):
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        resid = torch.randn(M, N, device=dev)
        bias = torch.randn(N, device=dev)
        for name, out_fp32, res, b, inplace in (("bf16 out", False, False, False, False), ("fp32 out", True, False, False, False),
                                       ("fp32 out + residual", True, True, False, False), ("fp32 + residual in place", True, True, False, True),
                                       ("fp32 out + bias", True, False, True, False)):
            out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
            o = resid if inplace else out
            def ours():
                _lib.check(lib.b200codec_gemm(0, ptr(a), ptr(w), M, N, K, 1, ptr(o), 0 if out_fp32 else 1, N, ptr(bias) if b else None,
                                              ptr(resid) if res else None, N if res else 0, 0, s))
            tc, tw = timeit(ours), timeit(ours, do_flush=False)
            print(f"{name:28s} {M:6d} {N:5d} {K:5d} {tc*1e3:8.1f} {tw*1e3:8.1f} {2.0*M*N*K/tw/1e9:9.0f}")
        def cublas(): torch.matmul(a, w.t())
        tc, tw = timeit(cublas), timeit(cublas, do_flush=False)
        print(f"{'cuBLAS bf16 out':28s} {M:6d} {N:5d} {K:5d} {tc*1e3:8.1f} {tw*1e3:8.1f} {2.0*M*N*K/tw/1e9:9.0f}")
