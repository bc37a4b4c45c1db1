"""Per-kernel counts of the Blackwell-specific SASS opcodes in the built library (evidence that the hot path is
tcgen05 / TMEM / TMA code, /opt/skills/guides/B200_PROFILING.md "What proves a Blackwell-native kernel").
  python tools/sass_opcodes.py [lib.so] > profiles/sass_opcodes.txt
Columns: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store /
reduce, UBLKCP = cp.async.bulk, MUFU.EX2, HMMA (legacy mma.sync: must be 0)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "stac_speech_translation_b200/libstac_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cols = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "MUFU.EX2", "HMMA", "instructions"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur is None or not m:
        continue
    op = m.group(1)
    cur["instructions"] += 1
    if re.match(r"UTC[A-Z]*MMA", op):
        cur["UTC*MMA"] += 1
    for c in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA"):
        if op.startswith(c):
            cur[c] += 1
    if op.startswith("MUFU.EX2"):
        cur["MUFU.EX2"] += 1
print(f"# {lib}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
print(f"{'kernel':60s} " + " ".join(f"{c:>9s}" for c in cols))
tot = collections.Counter()
for k, c in per.items():
    print(f"{k[:60]:60s} " + " ".join(f"{c[x]:9d}" for x in cols))
    tot.update(c)
print(f"{'TOTAL':60s} " + " ".join(f"{tot[x]:9d}" for x in cols))
