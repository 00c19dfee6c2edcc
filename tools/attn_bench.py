"""Attention kernel micro-benchmark: time vs number of CTAs (co-residency check) and TFLOP/s."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_max_b200 import _lib
lib = _lib.load()
def ptr(t): return ctypes.c_void_p(t.data_ptr())
def run(seqlens, impl, iters=10):
    lib.b200codec_set_attention_impl(impl)
    rows = sum(seqlens)
    qkv = torch.randn(rows, 3072, device="cuda").bfloat16()
    out = torch.empty(rows, 1024, device="cuda", dtype=torch.bfloat16)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    arr = _lib.i32_array(seqlens)
    for _ in range(2):
        lib.b200codec_attention(0, ptr(qkv), arr, len(seqlens), 16, ptr(out), s)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lib.b200codec_attention(0, ptr(qkv), arr, len(seqlens), 16, ptr(out), s); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts)//2]
    flops = sum(4.0 * T * T * 64 * 16 for T in seqlens)
    return t, flops / t / 1e9
for name, sl in (("1x1024 (128 CTAs)", [1024]), ("2x1024 (256 CTAs)", [1024]*2), ("4x1024 (512 CTAs)", [1024]*4),
                 ("16x500", [500]*16), ("4x3000", [3000]*4), ("1x3000", [3000]), ("64x150", [150]*64)):
    for impl, iname in ((0, "tcgen05"), (1, "mma.sync")):
        t, tf = run(sl, impl)
        print(f"{name:20s} {iname:9s} {t*1e3:9.1f} us (incl. plan upload + sync) {tf:7.0f} TF/s")
