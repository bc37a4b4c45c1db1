"""Kernel timeline of one CUDA-graph replay of the configs[1] step (CUPTI through torch.profiler): per-kernel warm device
times inside the graph and the idle gaps between consecutive kernels.  A breakdown, not a bench value.
  python tools/profile_gaps.py [--replays 3]"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stac_speech_translation_b200 as sb  # noqa: E402
from stac_speech_translation_b200 import synth  # noqa: E402
from stac_speech_translation_b200.pipeline import GraphedPipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replays", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hp = sb.HParams.for_size("S")
mods = sb.build_modules(hp, precision="bf16", device=dev)
wavs, wl = synth.fast_synth_batch(64, 30.0, seed=1234)
wavs, wl = wavs.to(dev), wl.to(dev)
calib = wavs[:8, : 16000 * 4].contiguous()
mods["normalize"].calibrate(mods["compute_features"](calib), torch.ones(8, device=dev))
g = GraphedPipeline(sb.EncoderPipeline(mods), wavs, wl)
for _ in range(3):
    g.graph.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.replays):
        g.graph.replay()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() > 0),
            key=lambda e: e.time_range.start)
per = len(ev) // a.replays
ev = ev[-per:]                                    # the last replay
busy = sum(e.time_range.elapsed_us() for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
gaps = [ev[i + 1].time_range.start - ev[i].time_range.end for i in range(len(ev) - 1)]
print(f"kernels {len(ev)}  span {span:.1f} us  busy {busy:.1f} us  gaps {sum(gaps):.1f} us "
      f"(mean {sum(gaps) / max(len(gaps), 1):.2f}, max {max(gaps):.2f})")
agg = collections.OrderedDict()
for i, e in enumerate(ev):
    k = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]
    d = agg.setdefault(k, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += e.time_range.elapsed_us()
    d[2] += gaps[i] if i < len(gaps) else 0.0
print("kernel, launches, us per launch, us per step, gap behind it (us per launch)")
for k, (n, t, gp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:60s} {n:4d} {t / n:9.2f} {t:9.1f} {gp / n:7.2f}")
