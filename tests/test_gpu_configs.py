"""Parity cases shaped like BASELINE.json configs[2..4] (reduced so the CPU oracle finishes in seconds), the
CTC-greedy agreement criterion of the north star, and the valid-length kernel (GPU only)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import BF16_TOL, FP32_TOL, oracle_modules, product_from_oracle, rel_l2  # noqa: E402
import oracle  # noqa: E402
import stac_speech_translation_b200 as sb  # noqa: E402
from stac_speech_translation_b200 import ops, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _valid_frames(wl, t2):
    """Frames the reference's encode() mask keeps: j <= floor(wav_len * T2) (TransformerMultiTask.py:289-294)."""
    return (torch.floor(wl * t2) + 1).clamp(1, t2).long().tolist()


def _cat_valid(x, n_list):
    return torch.cat([x[i, :n] for i, n in enumerate(n_list)])


def test_kv_lengths_kernel_matches_reference_expressions():
    for t2 in (1, 26, 37, 251, 751, 1501, 2500):
        g = torch.Generator().manual_seed(t2)
        wl = torch.rand(4096, generator=g)
        wl[:64] = (torch.arange(64).float() + 0.5) / t2          # products that land on .5 (round-half-to-even)
        wl[64] = 1.0
        wl[65] = 0.0
        for train_mask in (False, True):
            want = ops.kv_lengths(wl, 4096, t2, "cpu", train_mask)            # the torch expressions of the reference
            got = ops.kv_lengths(wl.cuda(), 4096, t2, "cuda", train_mask)
            assert torch.equal(got.cpu(), want), (t2, train_mask)
    assert ops.kv_lengths(None, 3, 10, "cuda", False).tolist() == [10, 10, 10]


def test_config3_medium_model_ragged_bucketed_batches():
    """M-size model, length-bucketed ragged batches (reduced: 4 layers, short utterances), both precisions."""
    omods = oracle_modules("M", num_encoder_layers=4)
    durs = synth.lognormal_durations(24, seed=3, median_s=2.0, sigma=0.6, lo=0.6, hi=5.0)
    bucketed = synth.bucket_batches(durs, max_batch_len=16.0, num_buckets=4, max_batch_ex=8)
    mods32, mods16 = product_from_oracle(omods, "fp32"), product_from_oracle(omods, "bf16")
    assert sum(len(b) for b in bucketed.batches) == 24
    for bi, idx in enumerate(bucketed.batches[:4]):
        wavs, wl = synth.synth_batch([float(durs[i]) for i in idx], seed=100 + bi)
        ref = oracle.reference_compute_forward(omods, wavs, wl)
        t2 = ref["enc_out"].shape[1]
        nv = _valid_frames(wl, t2)
        for mods, tol in ((mods32, FP32_TOL), (mods16, BF16_TOL)):
            res = sb.EncoderPipeline(mods)(wavs.cuda(), wl.cuda())
            assert rel_l2(_cat_valid(res["enc_out"].cpu(), nv), _cat_valid(ref["enc_out"], nv)) < tol
            assert rel_l2(_cat_valid(res["p_ctc"].cpu(), nv), _cat_valid(ref["p_ctc"], nv)) < tol
            assert rel_l2(res["enc_out"], ref["enc_out"]) < tol             # padded query rows as well


def test_config4_large_model_60s_long_masks():
    """L-size model on 60 s inputs (T'' = 1501, the longest BASELINE config), ragged key-padding masks."""
    omods = oracle_modules("L", num_encoder_layers=2)
    wavs, wl = synth.synth_batch([60.0, 47.3], seed=44, turns=4)
    ref = oracle.reference_compute_forward(omods, wavs, wl)
    assert ref["enc_out"].shape[1] == 1501
    nv = _valid_frames(wl, 1501)
    for precision, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        res = sb.EncoderPipeline(product_from_oracle(omods, precision))(wavs.cuda(), wl.cuda())
        assert rel_l2(_cat_valid(res["enc_out"].cpu(), nv), _cat_valid(ref["enc_out"], nv)) < tol
        assert rel_l2(_cat_valid(res["p_ctc"].cpu(), nv), _cat_valid(ref["p_ctc"], nv)) < tol
        assert float((res["p_ctc"].exp().sum(-1) - 1).abs().max()) < 1e-3


def test_config5_frontend_only_sweep():
    """Fbank + InputNormalization + ConvolutionFrontEnd over variable-length utterances (front-end-only sweep)."""
    omods = oracle_modules("S", num_encoder_layers=1)
    durs = synth.uniform_durations(40, seed=5, lo=1.0, hi=30.0)
    bucketed = synth.bucket_batches(durs, max_batch_len=60.0, num_buckets=8, max_batch_ex=16)
    mods32, mods16 = product_from_oracle(omods, "fp32"), product_from_oracle(omods, "bf16")
    for bi in (0, len(bucketed.batches) // 2, len(bucketed.batches) - 1):       # shortest, middle, longest bucket
        idx = bucketed.batches[bi]
        wavs, wl = synth.synth_batch([float(durs[i]) for i in idx], seed=200 + bi)
        ref = oracle.reference_compute_forward(omods, wavs, wl, stages="frontend")
        for mods, tol in ((mods32, FP32_TOL), (mods16, BF16_TOL)):
            got = sb.EncoderPipeline(mods)(wavs.cuda(), wl.cuda(), stop_after="cnn")["cnn"]
            assert got.shape == (len(idx), ref["cnn"].shape[1], 5120)
            assert rel_l2(got.float().cpu().view_as(ref["cnn"]), ref["cnn"]) < tol


def test_ctc_greedy_agreement_bf16_vs_oracle():
    """North star: identical CTC-greedy tokens in bf16 mode.  With random (untrained) weights most frames have a
    top-2 logit margin far below bf16 resolution, so the hard assertion is on frames whose oracle margin is clear;
    the raw agreement and the sequence-level agreement are recorded under gpurun_out/ for the round's notes."""
    omods = oracle_modules("S")
    torch.manual_seed(5)
    with torch.no_grad():                      # peakier posteriors: a trained CTC head is far from uniform
        omods["ctc_lin"].w.weight.mul_(6.0)
    secs = [6.0, 5.5, 5.0, 4.2, 3.7, 3.1, 2.4, 1.6]
    wavs, wl = synth.synth_batch(secs, seed=31, turns=2)
    ref = oracle.reference_compute_forward(omods, wavs, wl)
    t2 = ref["p_ctc"].shape[1]
    nv = _valid_frames(wl, t2)
    res = sb.EncoderPipeline(product_from_oracle(omods, "bf16"))(wavs.cuda(), wl.cuda())
    ids = res["greedy"].cpu().long()
    assert torch.equal(ids, res["p_ctc"].argmax(-1).cpu())                     # fused argmax == argmax of its posteriors
    want = ref["p_ctc"].argmax(-1)
    top2 = ref["logits"].topk(2, -1).values
    margin = top2[..., 0] - top2[..., 1]
    err = float((res["p_ctc"].cpu() - ref["p_ctc"]).abs().max())
    valid = torch.zeros_like(want, dtype=torch.bool)
    for i, n in enumerate(nv):
        valid[i, :n] = True
    clear = valid & (margin > 4 * err)
    assert bool((ids[clear] == want[clear]).all())
    frame_agree = float((ids[valid] == want[valid]).float().mean())
    seq_got = sb.ctc_greedy_collapse(ids, nv)
    seq_ref = sb.ctc_greedy_collapse(want, nv)
    seq_agree = float(np.mean([a == b for a, b in zip(seq_got, seq_ref)]))
    os.makedirs(OUT, exist_ok=True)
    json.dump({"frames": int(valid.sum()), "clear_margin_frames": int(clear.sum()), "frame_agreement": frame_agree,
               "sequence_agreement": seq_agree, "max_abs_logprob_error": err,
               "median_top2_margin": float(margin[valid].median())},
              open(os.path.join(OUT, "ctc_greedy_agreement.json"), "w"), indent=1)
    assert frame_agree > 0.9


def test_graphed_pipeline_replays_the_same_results():
    """sb.GraphedPipeline (one CUDA graph of the whole path) against the eager pipeline, on new audio per replay."""
    omods = oracle_modules("S", num_encoder_layers=2)
    mods = product_from_oracle(omods, "bf16")
    pipe = sb.EncoderPipeline(mods)
    wavs, wl = synth.synth_batch([2.0, 1.5, 1.1], seed=8)
    wavs2, _ = synth.synth_batch([2.0, 1.5, 1.1], seed=9)
    buf = wavs.cuda().clone()
    graphed = sb.GraphedPipeline(pipe, buf, wl.cuda())
    for w in (wavs, wavs2, wavs):
        got = graphed(w.cuda())
        torch.cuda.synchronize()
        want = pipe(w.cuda(), wl.cuda())
        torch.cuda.synchronize()
        assert torch.equal(got["greedy"], want["greedy"])
        assert torch.equal(got["enc_out"], want["enc_out"])
        assert torch.equal(got["p_ctc"], want["p_ctc"])
