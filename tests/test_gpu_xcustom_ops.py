"""The optional torch custom-op layer (custom_ops.py) on the device: every op equals the direct call it wraps."""
import os

import pytest
import torch

# First run on a B200 in round 2 (profiles/r3/r3a_first_call_verification.log): part of the default GPU suite since.
pytestmark = [pytest.mark.gpu]

import stac_speech_translation_b200.custom_ops  # noqa: E402,F401
from stac_speech_translation_b200 import ingest, ops, turns  # noqa: E402


def test_custom_ops_equal_direct_calls():
    ns = torch.ops.stac_b200
    g = torch.Generator().manual_seed(0)
    wavs = (torch.randn(2, 8000, generator=g) * 0.1).cuda()
    feats = ns.fbank(wavs, 80.0, True)
    assert torch.equal(feats, ops.fbank(wavs, ops.build_fbank_tables(wavs.device), 80.0, True))
    mean, std = torch.randn(80, generator=g).cuda(), (torch.rand(80, generator=g) + 0.5).cuda()
    assert torch.equal(ns.input_norm(feats, mean, std), ops.input_norm(feats, mean, std))
    x = torch.randn(3, 7, 128, generator=g).cuda()
    w, b = torch.randn(50, 128, generator=g).cuda(), torch.randn(50, generator=g).cuda()
    y = ns.linear(x, w, b, "fp32")
    assert torch.equal(y, ops.linear(x, w, b, "fp32"))
    assert torch.allclose(y, x @ w.T + b, atol=1e-3, rtol=1e-4)
    out, ids = ns.log_softmax_greedy(y)
    assert torch.equal(out, ns.log_softmax(y)) and torch.equal(ids.long(), y.argmax(-1))
    assert torch.equal(ns.argmax_rows(out), turns.greedy_ids(out))
    pcm = torch.randint(-32768, 32768, (4, 999), generator=g, dtype=torch.int16).cuda()
    assert torch.equal(ns.pcm_to_float(pcm), ingest.pcm_to_float(pcm))
