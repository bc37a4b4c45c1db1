"""Times the two d_model = 256 projections of an encoder layer at the benchmark shape (48064 rows) with the
weight-resident kernel and with the general one (STAC_WRES=0 in a child process).  python tools/bench_gemm.py"""
import os, subprocess, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stac_speech_translation_b200 import ops, _lib
if os.environ.get("STAC_LIB"):                      # timing experiments: a variant build of the library
    from pathlib import Path
    _lib.LIB_PATH = Path(os.environ["STAC_LIB"]).resolve()

def time_it(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

m = 64 * 751
a = torch.randn(m, 256, device="cuda").to(torch.bfloat16)
wq = (torch.randn(768, 256, device="cuda") / 16).to(torch.bfloat16)
wo = (torch.randn(256, 256, device="cuda") / 16).to(torch.bfloat16)
bq, bo = torch.randn(768, device="cuda"), torch.randn(256, device="cuda")
qkv = torch.empty(m, 768, device="cuda", dtype=torch.bfloat16)
x = torch.zeros(m, 256, device="cuda")
big = torch.empty(1 << 28, device="cuda")          # 1 GiB: flush L2 between variants
print(os.environ.get("STAC_LIB", "in-tree"), "STAC_WRES", os.environ.get("STAC_WRES", "1"),
      "qkv %.1f us" % time_it(lambda: ops._gemm(a, wq, bq, qkv, "bf16")),
      "out_proj %.1f us" % time_it(lambda: ops._gemm(a, wo, bo, x, "bf16", resid=x)))
if os.environ.get("STAC_WRES") is None and not os.environ.get("STAC_LIB"):
    subprocess.run([sys.executable, __file__], env={**os.environ, "STAC_WRES": "0"})
