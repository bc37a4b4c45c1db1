"""Times stac_mha_bf16 alone at the benchmark shape (64 x 751 frames, d 256, 4 heads) for every .so given.
  python tools/bench_mha.py lib1.so [lib2.so ...]      (timing experiment; variants built with -DMHA_* give wrong results)"""
import ctypes, os, sys, torch
from ctypes import c_void_p, c_int64
b, t, d, h = 64, 751, 256, 4
qkv = (torch.randn(b * t, 3 * d, device="cuda") * 1.0).to(torch.bfloat16)
kv = torch.full((b,), t, dtype=torch.int32, device="cuda")
ctx = torch.empty(b * t, d, device="cuda", dtype=torch.bfloat16)
st = c_void_p(torch.cuda.current_stream().cuda_stream)
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    f = lib.stac_mha_bf16
    f.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p]
    call = lambda: f(qkv.data_ptr(), None, kv.data_ptr(), b, t, 752, d, h, ctx.data_ptr(), st)
    for _ in range(5):
        rc = call()
    torch.cuda.synchronize()
    assert rc == 0, rc
    calls = {"stac_mha_bf16": call}
    if hasattr(lib, "stac_mha_bf16_v2"):
        f2 = lib.stac_mha_bf16_v2
        f2.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p]
        calls["stac_mha_bf16_v2"] = lambda: f2(qkv.data_ptr(), kv.data_ptr(), b, t, d, h, ctx.data_ptr(), st)
    for name, fn in calls.items():
        for _ in range(5):
            rc = fn()
        torch.cuda.synchronize()
        assert rc == 0, rc
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        flops = 4.0 * b * t * t * d
        print(f"{path:40s} {name:18s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")
