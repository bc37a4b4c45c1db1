"""conv block 0 (stac_conv0_ln_lrelu, bf16 output) alone at the benchmark shape (64 x 3001 frames) for every .so given.
  python tools/bench_conv0.py lib1.so [lib2.so ...]     (timing-variant builds -DC0_NO_* give wrong results)"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stac_speech_translation_b200 import _lib
b, t = 64, 3001
t1 = (t - 1) // 2 + 1
feats = torch.randn(b, t, 80, device="cuda")
w0 = torch.randn(256, 9, device="cuda") * 0.3; b0 = torch.randn(256, device="cuda") * 0.1
g = torch.rand(40 * 256, device="cuda") + 0.5; be = torch.randn(40 * 256, device="cuda") * 0.1
tp2 = (t1 + 3) // 2
out = torch.empty(b * 4 * tp2 * 21 * 256, device="cuda", dtype=torch.bfloat16)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    f = lib.stac_conv0_ln_lrelu
    res, args = _lib._SIGNATURES["stac_conv0_ln_lrelu"]
    f.restype, f.argtypes = res, args
    call = lambda: f(feats.data_ptr(), w0.data_ptr(), b0.data_ptr(), g.data_ptr(), be.data_ptr(), b, t, out.data_ptr(), _lib.DT_BF16, st)
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    gb = out.numel() * 2 / 1e9
    print(f"{os.path.basename(path):36s} {ts[len(ts)//2]*1e3:7.1f} us   {gb / ts[len(ts)//2] * 1e3:6.0f} GB/s written")
