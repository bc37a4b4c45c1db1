"""Timeline of CTA 0 of the second tensor-core Fbank kernel (lib built with -DFBANK2_TRACE) at the benchmark shape.
usage: python tools/trace_fbank2.py libstac_b200_fbtrace.so [pair]"""
import ctypes, os, sys, torch
from ctypes import c_void_p, c_int64, c_int
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stac_speech_translation_b200 import ops
b, n = 64, 480000
pair = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wavs = torch.randn(b, n, device="cuda") * 0.1
tabs = ops.build_fbank_tc_tables("cuda")
t = 1 + n // 160
db = torch.empty(b, t, 80, device="cuda"); umax = torch.empty(b, dtype=torch.int32, device="cuda")
lib = ctypes.CDLL(sys.argv[1])
f = lib.stac_fbank_logmel_tc2
f.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
st = c_void_p(torch.cuda.current_stream().cuda_stream)
call = lambda: f(wavs.data_ptr(), b, n, n, tabs.tab2.data_ptr(), tabs.tw2.data_ptr(), db.data_ptr(), umax.data_ptr(), pair, st)
for _ in range(3):
    assert call() == 0
torch.cuda.synchronize()
buf = torch.zeros(4 * 8 * 16, dtype=torch.int32, device="cuda")
lib.stac_fbank2_trace.argtypes = [c_void_p]
lib.stac_fbank2_trace(buf.data_ptr())
call(); torch.cuda.synchronize()
tr = (buf.cpu().long() & 0xffffffff).view(4, 8, 16)
base = int(tr[3, 0, 0]) or int(tr[0, 0, 0])
names = ["producer (first producer warp): 0 pcm_full passed, 1..7 stage i stored; stage 2: 8 folded, 9 slot free", "mma: 0 tempty passed, 1..7 stage i inputs ready",
         "epilogue: 0 loop top, 1 tfull passed, 2 TMEM released, 3 dB + max done, 4 rows stored",
         "pcm loader: 0 loop top, 1 pcm_free passed, 2 zero fill done, 3 copies issued"]
print(f"pair={pair}")
for role in range(4):
    print(names[role])
    for tnum in range(6):
        print(f"{tnum:3d} " + " ".join(f"{(int(v) - base) & 0xffffffff:7d}" if v else "      -" for v in tr[role, tnum, :(12 if role == 0 else 9)]))
