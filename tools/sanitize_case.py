"""Smallest end-to-end case that touches every bf16-mode kernel (for compute-sanitizer)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stac_speech_translation_b200 as sb
from stac_speech_translation_b200 import synth

dev = torch.device("cuda", 0)
mods = sb.build_modules(sb.HParams.for_size("S", num_encoder_layers=2), precision="bf16", device=dev)
wavs, wl = synth.synth_batch([1.31, 0.9, 0.47], seed=3)
wavs, wl = wavs.to(dev), wl.to(dev)
mods["normalize"].calibrate(mods["compute_features"](wavs), wl)
res = sb.EncoderPipeline(mods)(wavs, wl)
torch.cuda.synchronize()
print("ok", tuple(res["p_ctc"].shape), float(res["p_ctc"].exp().sum(-1).mean()))
