"""Shared helpers for the parity tests: build the oracle and the product object graphs with
identical weights and normalisation statistics, and the error metrics the tolerances use."""
import torch

import oracle
import stac_speech_translation_b200 as sb
from stac_speech_translation_b200 import synth

FP32_TOL = 1e-4   # north_star: fp32 mode within 1e-4 relative
BF16_TOL = 2e-2   # north_star: bf16 mode within 2e-2 relative

TINY = dict(d_model=128, nhead=2, num_encoder_layers=2, d_ffn=512)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def oracle_modules(size="S", seed=8886, calib=None, **over):
    mods = oracle.build_reference_modules(size, seed=seed, **over)
    wavs, wl = calib if calib is not None else synth.synth_batch([2.0, 1.4, 0.9], seed=77)
    norm = mods["normalize"]
    norm.train()
    norm(mods["compute_features"](wavs), wl)   # SpeechBrain train-mode statistics step
    norm.eval()
    return mods


def product_from_oracle(omods, precision, device="cuda"):
    """Product object graph carrying the oracle's weights (loaded through state_dict, as a
    reference checkpoint would be) and its normaliser statistics."""
    tr = omods["Transformer"]
    d = tr.encoder.norm.norm.weight.shape[0]
    layer0 = tr.encoder.layers[0]
    hp = sb.HParams(d_model=d, nhead=layer0.self_att.att.num_heads, num_encoder_layers=len(tr.encoder.layers),
                    d_ffn=layer0.pos_ffn.ffn[0].weight.shape[0],
                    output_neurons=omods["ctc_lin"].w.weight.shape[0])
    mods = sb.build_modules(hp, precision=precision, device=device)
    for k in ("CNN", "Transformer", "ctc_lin"):
        missing = mods[k].load_state_dict(omods[k].state_dict(), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
    mods["normalize"]._load_statistics_dict(omods["normalize"]._statistics_dict())
    return mods
