#!/usr/bin/env python
"""Summarise an ncu report for profiles/: per-kernel duration, DRAM bytes, pipe utilisation.

  python tools/ncu_extract.py gpurun_out/prof.ncu-rep --tag r1h [--map 'conv0_ln_lrelu_bf16=stac_conv0_ln_lrelu' ...]

Reads the raw page (`ncu -i <rep> --page raw --csv`), normalises units, writes
profiles/<tag>_ncu_summary.csv and merges DRAM bytes per launch into profiles/ncu_traffic.json
(the file bench.py reads for `roofline.traffic`).  Numbers under ncu are cold-cache and serialised:
they document traffic and pipe shares, never a bench value.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
COLS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("dram__bytes_read.sum", "dram_read_bytes"),
    ("dram__bytes_write.sum", "dram_write_bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor_hmma_inst_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]


def raw_rows(path):
    if path.endswith(".csv"):
        text = open(path).read()
    else:
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True,
                              check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    return rows[0], rows[1], rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--map", action="append", default=[], help="regex=trace_key (first match wins)")
    ap.add_argument("--no-traffic", action="store_true")
    args = ap.parse_args()
    hdr, units, rows = raw_rows(args.report)
    col = {h: i for i, h in enumerate(hdr)}
    maps = [(re.compile(m.split("=", 1)[0]), m.split("=", 1)[1]) for m in args.map]
    out_rows, traffic = [], {}
    for r in rows:
        name = r[col["Kernel Name"]]
        rec = {"kernel": re.sub(r"\(.*", "", name).replace("<unnamed>::", "").replace("void ", "")}
        for metric, short in COLS:
            if metric in col and r[col[metric]] != "":
                try:
                    v = float(r[col[metric]].replace(",", ""))
                except ValueError:
                    continue
                rec[short] = v * UNIT_SCALE.get(units[col[metric]], 1.0)
        out_rows.append(rec)
        for rx, key in maps:
            if rx.search(name):
                t = traffic.setdefault(key, {"n": 0, "bytes": 0.0, "us": 0.0})
                t["n"] += 1
                t["bytes"] += rec.get("dram_read_bytes", 0.0) + rec.get("dram_write_bytes", 0.0)
                t["us"] += rec.get("duration_us", 0.0)
                break
    fields = ["kernel"] + [s for _, s in COLS]
    dst = os.path.join(ROOT, "profiles", f"{args.tag}_ncu_summary.csv")
    with open(dst, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=fields)
        w.writeheader()
        for rec in out_rows:
            w.writerow({k: (f"{rec[k]:.6g}" if isinstance(rec.get(k), float) else rec.get(k, "")) for k in fields})
    print("wrote", dst, len(out_rows), "launches")
    if traffic and not args.no_traffic:
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        cur = json.load(open(tp)) if os.path.exists(tp) else {}
        for key, t in traffic.items():
            cur[key] = {"dram_bytes_per_launch": int(t["bytes"] / t["n"]), "launches": t["n"],
                        "ncu_avg_us": round(t["us"] / t["n"], 2), "source": f"profiles/{args.tag}_ncu_summary.csv"}
        json.dump(cur, open(tp, "w"), indent=1, sort_keys=True)
        print("updated", tp, sorted(traffic))


if __name__ == "__main__":
    sys.exit(main())
