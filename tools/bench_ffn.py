"""Times stac_ffn_fused_bf16 alone at the benchmark shape for every .so given (timing experiments: variants built with
-DFFN_NOLOAD / -DFFN_NOGELU give wrong results).  python tools/bench_ffn.py lib1.so [lib2.so ...]"""
import ctypes, sys, torch
from ctypes import c_void_p, c_int64
m, d, dffn = 48064, 256, 1024
h = torch.randn(m, d, device="cuda").to(torch.bfloat16)
w1 = (torch.randn(dffn, d, device="cuda") / 16).to(torch.bfloat16)
w2 = (torch.randn(d, dffn, device="cuda") / 32).to(torch.bfloat16)
b1, b2 = torch.randn(dffn, device="cuda"), torch.randn(d, device="cuda")
x = torch.zeros(m, d, device="cuda")
st = c_void_p(torch.cuda.current_stream().cuda_stream)
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    f = lib.stac_ffn_fused_bf16
    f.argtypes = [c_void_p] * 6 + [c_int64] * 3 + [c_void_p]
    call = lambda: f(h.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), x.data_ptr(), m, d, dffn, st)
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    print(f"{path:40s} {us:8.1f} us  {4.0 * m * d * dffn / us / 1e6:7.1f} TFLOP/s")
