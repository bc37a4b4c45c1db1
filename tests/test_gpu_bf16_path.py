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


def test_bf16_pipeline_at_the_benchmark_shape():
    """8 x 30 s multi-turn segments through the full 12-layer S model in bf16 mode - a slice of BASELINE.json configs[1]
    (64 x 30 s), the shape bench.py times - against the oracle at the north-star tolerance, per utterance, plus the
    size-independent properties (posteriors normalised, greedy ids = arg-max of the posteriors)."""
    import os
    torch.set_num_threads(os.cpu_count() or 1)
    omods = oracle_modules("S")
    wavs, wl = synth.fast_synth_batch(8, 30.0, seed=1234)
    ref = oracle.reference_compute_forward(omods, wavs, wl)
    mods = product_from_oracle(omods, "bf16")
    res = sb.EncoderPipeline(mods)(wavs.cuda(), wl.cuda())
    torch.cuda.synchronize()
    assert res["enc_out"].shape == ref["enc_out"].shape == (8, 751, 256)
    for i in range(8):
        assert rel_l2(res["enc_out"][i], ref["enc_out"][i]) < BF16_TOL, i
        assert rel_l2(res["p_ctc"][i], ref["p_ctc"][i]) < BF16_TOL, i
    p = res["p_ctc"].float()
    assert torch.allclose(p.exp().sum(-1), torch.ones_like(p[..., 0]), atol=1e-3)
    assert torch.equal(res["greedy"].long(), p.argmax(-1))
