"""configs[1] batches (S model, 64 x 30 s, bf16) replayed as CUDA graphs on 1..N streams with alternating batches: does
overlapping one batch's memory-bound front end and kernel tails with another batch's tensor-bound layers raise the
whole-job rate?  (run on a B200)   python tools/bench_streams.py [--streams 1 2 3] [--steps 12]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stac_speech_translation_b200 as sb  # noqa: E402
from stac_speech_translation_b200.pipeline import GraphedPipeline  # noqa: E402
from stac_speech_translation_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, nargs="+", default=[1, 2, 3])
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=30.0)
a = ap.parse_args()
dev = torch.device("cuda:0")
hp = sb.HParams.for_size("S")
mods = sb.build_modules(hp, precision="bf16", device=dev)
wavs, wl = synth.fast_synth_batch(a.batch, a.seconds, seed=1234)
wavs, wl = wavs.to(dev), wl.to(dev)
calib = wavs[: min(8, a.batch), : 16000 * 4].contiguous()
mods["normalize"].calibrate(mods["compute_features"](calib), torch.ones(calib.shape[0], device=dev))
n_max = max(a.streams)
pipes = [sb.EncoderPipeline(mods) for _ in range(n_max)]
streams = [torch.cuda.Stream(device=dev) for _ in range(n_max)]
graphs = []
for p, s in zip(pipes, streams):
    with torch.cuda.stream(s):
        graphs.append(GraphedPipeline(p, wavs.clone(), wl.clone()))
torch.cuda.synchronize()
graphs[0].graph.replay()                          # (capture does not execute: the result buffers are filled by a replay)
torch.cuda.synchronize()
key = [k for k in graphs[0].out if "id" in k or "greedy" in k]
key = key[0] if key else list(graphs[0].out)[-1]
ref = graphs[0].out[key].clone()
audio = a.batch * a.seconds
for n in a.streams:
    def run(k):
        for i in range(k):
            with torch.cuda.stream(streams[i % n]):
                graphs[i % n].graph.replay()
    run(3 * n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(dev)
    e0.record(main)
    for s in streams[:n]:
        s.wait_event(e0)
    run(a.steps)
    for s in streams[:n]:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    same = all(torch.equal(g.out[key], ref) for g in graphs[:n])
    print(f"streams {n}: {ms:.3f} ms per batch  {audio / ms * 1e3 / 1e3:.1f} k audio-s/s   {key} identical across graphs: {same}")
