#!/usr/bin/env python
"""Top stall sites of one kernel from an ncu report's source page.
  python tools/ncu_hot.py gpurun_out/prof.ncu-rep [--top 40] [--kernel-index 0]"""
import argparse, csv, io, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--top", type=int, default=40); ap.add_argument("--kernel-index", type=int, default=0)
ap.add_argument("--range", default=None, help="a:b instruction index range to print in full")
a = ap.parse_args()
txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# split per kernel
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[a.kernel_index]; e = starts[a.kernel_index + 1] if a.kernel_index + 1 < len(starts) else len(rows)
hdr = rows[s + 1]; data = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("kernel", rows[s][1][:80], "instructions", len(data), "samples", tot)
agg = {}
for r in data:
    for c in stall:
        if r[c] not in ("", "0"): agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
if a.range:
    lo, hi = map(int, a.range.split(":")); sel = range(lo, min(hi, len(data)))
else:
    sel = sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:a.top])
for i in sel:
    r = data[i]
    st = sorted([(int(r[c]), hdr[c][6:]) for c in stall if r[c] not in ("", "0")], reverse=True)[:3]
    print(f"{i:5d} {r[isrc].strip()[:64]:64s} samp {int(r[isamp]):6d} exec {r[iex]:>9s} {st}")
