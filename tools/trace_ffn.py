"""Timeline of CTA 0 of the fused feed-forward kernel (lib built with -DFFN_TRACE) at the benchmark shape."""
import ctypes, sys, torch
from ctypes import c_void_p, c_int64
m, d, dffn = 48064, 256, 1024
h = torch.randn(m, d, device="cuda").to(torch.bfloat16)
w1 = (torch.randn(dffn, d, device="cuda") / 16).to(torch.bfloat16)
w2 = (torch.randn(d, dffn, device="cuda") / 32).to(torch.bfloat16)
b1, b2 = torch.randn(dffn, device="cuda"), torch.randn(d, device="cuda")
x = torch.zeros(m, d, device="cuda")
lib = ctypes.CDLL(sys.argv[1])
f = lib.stac_ffn_fused_bf16
f.argtypes = [c_void_p] * 6 + [c_int64] * 3 + [c_void_p]
st = c_void_p(torch.cuda.current_stream().cuda_stream)
call = lambda: f(h.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), x.data_ptr(), m, d, dffn, st)
for _ in range(3):
    assert call() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    call()
e1.record(); torch.cuda.synchronize()
print(f"fused ffn m={m}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
buf = torch.zeros(2 * 32 * 8, dtype=torch.int32, device="cuda")
lib.stac_ffn_trace.argtypes = [c_void_p]
lib.stac_ffn_trace(buf.data_ptr())
call(); torch.cuda.synchronize()
tr = (buf.cpu().long() & 0xffffffff).view(2, 32, 8)
base = int(tr[0, 0, 3])
names = ["mma: 0 O-start 1 p_full passed 2 first W2 unit there | 3 S-start 4 s_free passed 5 first W1 unit there",
         "epilogue warp 0: 0 start 1 s_full passed 2 gelu done 3 p_free passed 4 p_full arrived"]
for role in range(2):
    print(names[role])
    for t in range(24):
        print(f"{t:3d} " + " ".join(f"{(int(v) - base) & 0xffffffff:7d}" if v else "      -" for v in tr[role, t, :6]))
