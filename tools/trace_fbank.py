"""Timeline of CTA 0 of the tensor-core Fbank kernel (lib built with -DFBANK_TRACE) at the benchmark shape."""
import ctypes, os, sys, torch
from ctypes import c_void_p, c_int64
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stac_speech_translation_b200 import ops
b, n = 64, 480000
wavs = torch.randn(b, n, device="cuda") * 0.1
tab, tw = ops.build_fbank_tc_tables("cuda")
t = 1 + n // 160
db = torch.empty(b, t, 80, device="cuda"); umax = torch.empty(b, dtype=torch.int32, device="cuda")
lib = ctypes.CDLL(sys.argv[1])
f = lib.stac_fbank_logmel_tc
f.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
st = c_void_p(torch.cuda.current_stream().cuda_stream)
call = lambda: f(wavs.data_ptr(), b, n, n, tab.data_ptr(), tw.data_ptr(), db.data_ptr(), umax.data_ptr(), st)
for _ in range(3):
    assert call() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    call()
e1.record(); torch.cuda.synchronize()
print(f"fbank tc: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
buf = torch.zeros(3 * 8 * 16, dtype=torch.int32, device="cuda")
lib.stac_fbank_trace.argtypes = [c_void_p]
lib.stac_fbank_trace(buf.data_ptr())
call(); torch.cuda.synchronize()
tr = (buf.cpu().long() & 0xffffffff).view(3, 8, 16)
base = int(tr[0, 0, 0])
names = ["producer: 0 start 1 pcm loaded 2..9 k-block built", "mma: 0 tempty 1..8 full passed", "epilogue: 0 start 1 tfull 2 mel done 3 staged 4 copied"]
for role in range(3):
    print(names[role])
    for tnum in range(4):
        print(f"{tnum:3d} " + " ".join(f"{(int(v) - base) & 0xffffffff:7d}" if v else "      -" for v in tr[role, tnum, :10]))
