"""End-to-end bf16-mode parity against the fp32 oracle (GPU only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import BF16_TOL, FP32_TOL, oracle_modules, product_from_oracle, rel_l2  # noqa: E402
import oracle  # noqa: E402
import stac_speech_translation_b200 as sb  # noqa: E402
from stac_speech_translation_b200 import synth  # noqa: E402


@pytest.mark.parametrize("size,secs,layers", [("S", [4.0, 3.3, 2.05, 1.0], 12), ("M", [3.0, 1.2], 4), ("L", [2.5, 2.0], 2)])
def test_bf16_pipeline_vs_oracle(size, secs, layers):
    omods = oracle_modules(size, num_encoder_layers=layers)
    wavs, wl = synth.synth_batch(secs, seed=21)
    ref = oracle.reference_compute_forward(omods, wavs, wl)
    mods = product_from_oracle(omods, "bf16")
    got = sb.compute_forward(mods, wavs.cuda(), wl.cuda())
    assert rel_l2(got["feats"], ref["feats"]) < FP32_TOL           # features are fp32 in both modes
    for k in ("cnn", "enc_out", "p_ctc"):
        assert got[k].dtype == torch.float32
        assert rel_l2(got[k], ref[k]) < BF16_TOL, (k, rel_l2(got[k], ref[k]))
    res = sb.EncoderPipeline(mods)(wavs.cuda(), wl.cuda())
    assert rel_l2(res["enc_out"], ref["enc_out"]) < BF16_TOL
    assert rel_l2(res["p_ctc"], ref["p_ctc"]) < BF16_TOL
