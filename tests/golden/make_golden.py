"""Generates tests/golden/path_tiny.npz: seeded inputs -> oracle outputs at every stage boundary.

SpeechBrain cannot be imported in this environment, so these vectors come
from the oracle restatement itself: they pin it against regressions and give the GPU tests a
file-based target, but they are NOT reference-generated ("parity unpinned", see oracle/__init__.py).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from stac_speech_translation_b200 import synth  # noqa: E402
from util import TINY, oracle_modules  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "path_tiny.npz")
SECONDS = [1.2, 0.83]
SEED_AUDIO, SEED_WEIGHTS, VOCAB = 4321, 8886, 64


def generate():
    torch.set_num_threads(1)          # fixed reduction order
    omods = oracle_modules(TINY, seed=SEED_WEIGHTS, vocab=VOCAB)
    wavs, wl = synth.synth_batch(SECONDS, seed=SEED_AUDIO)
    out = oracle.reference_compute_forward(omods, wavs, wl)
    out_t = oracle.reference_compute_forward(omods, wavs, wl, train_mask=True)
    return wavs, wl, out, out_t


if __name__ == "__main__":
    wavs, wl, out, out_t = generate()
    np.savez_compressed(
        GOLDEN, wav_lens=wl.numpy(), fbank=out["fbank"].numpy().astype(np.float32),
        feats=out["feats"].numpy(), cnn_sample=out["cnn"][:, ::7, ::3, ::16].numpy(),
        enc_out=out["enc_out"].numpy(), p_ctc=out["p_ctc"].numpy(),
        enc_out_train_mask=out_t["enc_out"].numpy(),
        wav_checksum=np.array([float(wavs.double().sum()), float(wavs.double().abs().sum())]))
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
