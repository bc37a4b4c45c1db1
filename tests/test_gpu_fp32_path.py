"""End-to-end fp32-mode parity of the drop-in call sequence against the oracle (GPU only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import FP32_TOL, oracle_modules, product_from_oracle, rel_l2, rel_max  # noqa: E402
import oracle  # noqa: E402
import stac_speech_translation_b200 as sb  # noqa: E402
from stac_speech_translation_b200 import synth  # noqa: E402


@pytest.mark.parametrize("train_mask", [False, True])
def test_small_model_all_stages(train_mask):
    omods = oracle_modules("S")
    wavs, wl = synth.synth_batch([4.0, 3.3, 2.05, 1.0], seed=11)
    ref = oracle.reference_compute_forward(omods, wavs, wl, train_mask=train_mask)
    mods = product_from_oracle(omods, "fp32")
    got = sb.compute_forward(mods, wavs.cuda(), wl.cuda(), train_mask=train_mask)
    for k in ("fbank", "feats", "cnn", "enc_out", "logits", "p_ctc"):
        assert got[k].shape == ref[k].shape, k
        assert got[k].dtype == torch.float32 and got[k].is_contiguous()
        assert rel_l2(got[k], ref[k]) < FP32_TOL, (k, rel_l2(got[k], ref[k]))
    assert rel_max(got["enc_out"], ref["enc_out"]) < 1e-3
    # fused pipeline gives the same answer as the staged drop-ins
    pipe = sb.EncoderPipeline(mods, train_mask=train_mask)
    res = pipe(wavs.cuda(), wl.cuda())
    assert rel_l2(res["enc_out"], ref["enc_out"]) < FP32_TOL
    assert rel_l2(res["p_ctc"], ref["p_ctc"]) < FP32_TOL
    agree = (res["greedy"].cpu().long() == ref["p_ctc"].argmax(-1)).float().mean()
    assert agree > 0.99


@pytest.mark.parametrize("size,secs", [("M", [2.0, 1.2]), ("L", [1.5])])
def test_medium_large_sizes(size, secs):
    omods = oracle_modules(size, num_encoder_layers=2)
    wavs, wl = synth.synth_batch(secs, seed=3)
    ref = oracle.reference_compute_forward(omods, wavs, wl)
    mods = product_from_oracle(omods, "fp32")
    got = sb.compute_forward(mods, wavs.cuda(), wl.cuda())
    assert rel_l2(got["enc_out"], ref["enc_out"]) < FP32_TOL
    assert rel_l2(got["p_ctc"], ref["p_ctc"]) < FP32_TOL
