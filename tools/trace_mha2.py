"""Per-step timeline of CTA 0 of the second attention kernel (stac_mha_bf16_v2); needs a variant build with the hooks:
  python -m stac_speech_translation_b200.build --variant mha2trace -- -DMHA2_TRACE
  python tools/trace_mha2.py stac_speech_translation_b200/libstac_b200_mha2trace.so [out.json]
Roles / events: see the TRACE2 comment in csrc/attention_tc2.cu.  Prints clocks relative to the first event for a few
steady-state steps of both MMA issuers and both softmax groups, and the mean step period."""
import ctypes
import json
import os
import sys

import torch
from ctypes import c_int64, c_void_p

b, t, d, h = 64, 751, 256, 4
qkv = torch.randn(b * t, 3 * d, device="cuda").to(torch.bfloat16)
kv = torch.full((b,), t, dtype=torch.int32, device="cuda")
ctx = torch.empty(b * t, d, device="cuda", dtype=torch.bfloat16)
lib = ctypes.CDLL(sys.argv[1])
f = lib.stac_mha_bf16_v2
f.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p]
st = c_void_p(torch.cuda.current_stream().cuda_stream)
call = lambda: f(qkv.data_ptr(), kv.data_ptr(), b, t, d, h, ctx.data_ptr(), st)
for _ in range(3):
    assert call() == 0
torch.cuda.synchronize()
buf = torch.zeros(5 * 64 * 8, dtype=torch.int32, device="cuda")
lib.stac_mha2_trace.argtypes = [c_void_p]
lib.stac_mha2_trace(buf.data_ptr())
assert call() == 0
torch.cuda.synchronize()
tr = (buf.cpu().long() & 0xffffffff).view(5, 64, 8)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(tr.tolist(), open(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/mha2_trace_raw.json", "w"))
base = int(tr[tr > 0].min())
names = {0: "mma0: 0 S-start 1 operands ready 2 S issued 3 PV-start 4 p_full passed 5 PV issued", 1: "mma1 (same events)",
         2: "softmax0: 0 start 1 s_full 2 scores in regs 3 max done 4 exp + P stores 5 p_full arrived",
         3: "softmax1 (same events)", 4: "epilogue (index = item * 2 + group): 0 start 1 l_full 2 o_full 3 O read 4 staged 5 store"}
for role in range(5):
    print(names[role])
    for step in range(12, 20):
        print(f"{step:3d} " + " ".join(f"{(int(x) - base) & 0xffffffff:8d}" for x in tr[role, step, :6]))
for role in (2, 3):
    starts = [int(tr[role, s, 1]) for s in range(8, 40) if int(tr[role, s, 1]) and int(tr[role, s + 1, 1])]
    ends = [int(tr[role, s + 1, 1]) for s in range(8, 40) if int(tr[role, s, 1]) and int(tr[role, s + 1, 1])]
    if starts:
        per = sum((e - s) & 0xffffffff for s, e in zip(starts, ends)) / len(starts)
        print(f"softmax group {role - 2}: mean period of a 96-key step {per:.0f} clk (MUFU floor for both groups: 1536)")
