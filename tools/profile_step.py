"""One warm-up step + one profiled step of the bf16 pipeline at the benchmark workload (for ncu)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stac_speech_translation_b200 as sb  # noqa: E402
from stac_speech_translation_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="S")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--stop-after", default=None, choices=[None, "cnn"])
ap.add_argument("--layers", type=int, default=None, help="encoder layers (default: the size's own; 1 keeps an ncu capture short)")
args = ap.parse_args()

dev = torch.device("cuda", 0)
hp = sb.HParams.for_size(args.size, **({"num_encoder_layers": args.layers} if args.layers else {}))
mods = sb.build_modules(hp, precision="bf16", device=dev)
wavs, wl = synth.fast_synth_batch(args.batch, args.seconds, seed=1234)
wavs, wl = wavs.to(dev), wl.to(dev)
calib = wavs[: min(8, args.batch), : 16000 * 4].contiguous()
mods["normalize"].calibrate(mods["compute_features"](calib), torch.ones(calib.shape[0], device=dev))
pipe = sb.EncoderPipeline(mods)
for _ in range(args.steps):
    res = pipe(wavs, wl, stop_after=args.stop_after)
torch.cuda.synchronize()
print("ok", float(next(iter(res.values())).float().abs().mean()))
