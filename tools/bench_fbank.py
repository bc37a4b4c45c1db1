"""Tensor-core Fbank kernels at the benchmark shape (64 x 30 s): time per launch, fraction of the HBM roofline, and
agreement between the variants.  usage: python tools/bench_fbank.py [lib.so]"""
import ctypes, json, os, sys, torch
from ctypes import c_void_p, c_int64, c_int
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stac_speech_translation_b200 import ops, _lib

b, n = int(os.environ.get("FB_BATCH", 64)), int(os.environ.get("FB_SAMPLES", 480000))
wavs = torch.randn(b, n, device="cuda") * 0.1
tabs = ops.build_fbank_tc_tables("cuda")
t = 1 + n // 160
lib = ctypes.CDLL(sys.argv[1]) if len(sys.argv) > 1 else _lib.lib()
st = c_void_p(torch.cuda.current_stream().cuda_stream)
f1, f2 = lib.stac_fbank_logmel_tc, lib.stac_fbank_logmel_tc2
f1.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
f2.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
outs = {}
peak = 6550.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in sys.argv[2:] or ["v1", "v2_single", "v2_pair"]:
    db = torch.full((b, t, 80), float("nan"), device="cuda")
    umax = torch.empty(b, dtype=torch.int32, device="cuda")
    if name == "v1":
        call = lambda: f1(wavs.data_ptr(), b, n, n, tabs.tab.data_ptr(), tabs.tw.data_ptr(), db.data_ptr(), umax.data_ptr(), st)
    else:
        pair = int(name == "v2_pair")
        call = lambda: f2(wavs.data_ptr(), b, n, n, tabs.tab2.data_ptr(), tabs.tw2.data_ptr(), db.data_ptr(), umax.data_ptr(), pair, st)
    for _ in range(3):
        rc = call()
        assert rc == 0, rc
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    gb = b * (4 * n + 4 * 80 * t) / 1e9
    print(f"{name:10s} {us:8.1f} us  (min {ts[0]:.1f})  {gb / us * 1e6:7.0f} GB/s = {gb / us * 1e6 / peak:.3f} of the HBM peak", flush=True)
    outs[name] = (db.clone(), umax.clone())
names = list(outs)
for a in names[1:]:
    d0, d1 = outs[names[0]][0], outs[a][0]
    rel = ((d0 - d1).norm() / d0.norm()).item()
    print(f"{a} vs {names[0]}: rel-L2 {rel:.2e}, max-key equal {bool(torch.equal(outs[names[0]][1], outs[a][1]))}, finite {bool(torch.isfinite(d1).all())}")
