"""Per-step timeline of CTA 0 of the attention kernel (lib built with -DMHA_TRACE)."""
import ctypes, sys, torch, collections
from ctypes import c_void_p, c_int64
b, t, d, h = 64, 751, 256, 4
qkv = torch.randn(b * t, 3 * d, device="cuda").to(torch.bfloat16)
kv = torch.full((b,), t, dtype=torch.int32, device="cuda")
ctx = torch.empty(b * t, d, device="cuda", dtype=torch.bfloat16)
lib = ctypes.CDLL(sys.argv[1])
f = lib.stac_mha_bf16
f.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p]
st = c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    f(qkv.data_ptr(), None, kv.data_ptr(), b, t, 752, d, h, ctx.data_ptr(), st)
torch.cuda.synchronize()
buf = torch.zeros(5 * 64 * 8, dtype=torch.int32, device="cuda")
lib.stac_mha_trace.argtypes = [c_void_p]
lib.stac_mha_trace(buf.data_ptr())
f(qkv.data_ptr(), None, kv.data_ptr(), b, t, 752, d, h, ctx.data_ptr(), st)
torch.cuda.synchronize()
tr = (buf.cpu().long() & 0xffffffff).view(5, 64, 8)
import json, os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(tr.tolist(), open(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/mha_trace_raw.json", "w"))
t0 = int(tr[tr > 0].min())
names = {0: "producer: kv_empty passed", 1: "mma0: 0 S-start 1 kv_full 2 s_free 3 S-issued 4 PV-start 5 p_full 6 PV-issued",
         3: "softmax0: 0 start 1 s_full 2 ld+s_free 3 p_free 4 exp/store done 5 p_full arrived"}
names[4] = 'softmax1 (same events)'
names[2] = 'mma1 (same events)'
for role in (1, 3):
    print(names[role])
    for step in range(20, 28):
        row = tr[role, step]
        base = int(tr[1, 14, 0])
        print(f"{step:3d} " + " ".join(f"{(int(x) - base) & 0xffffffff:7d}" for x in row[:8]))
