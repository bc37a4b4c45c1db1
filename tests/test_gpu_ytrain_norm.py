"""Train-mode InputNormalization on the device (stac_utt_mean_std + SpeechBrain's running update) against the oracle."""
import os

import pytest
import torch

# First run on a B200 in round 2 (profiles/r3/r3a_first_call_verification.log): part of the default GPU suite since.
pytestmark = [pytest.mark.gpu]

import stac_speech_translation_b200 as sb  # noqa: E402
from oracle.speechbrain_path import InputNormalization as OracleNorm  # noqa: E402
from util import rel_l2  # noqa: E402


def test_train_mode_statistics_and_running_update():
    g = torch.Generator().manual_seed(4)
    ours = sb.InputNormalization(norm_type="global", update_until_epoch=2).train()
    ref = OracleNorm(norm_type="global", update_until_epoch=2).train()
    for step, epoch in enumerate([0, 0, 1, 2, 3]):
        x = torch.randn(5, 301, 80, generator=g) * (1 + step) + step
        wl = torch.tensor([1.0, 0.73, 0.41, 0.0101, 0.5])
        got, want = ours(x.cuda(), wl.cuda(), epoch=epoch), ref(x, wl, epoch=epoch)
        assert rel_l2(got, want) < 1e-5, (step, epoch)
        assert ours.count == ref.count
        assert rel_l2(ours.glob_mean, ref.glob_mean) < 1e-5 and rel_l2(ours.glob_std, ref.glob_std) < 1e-5
    ours.eval(), ref.eval()
    x = torch.randn(2, 30, 80, generator=g)
    assert rel_l2(ours(x.cuda(), torch.ones(2).cuda()), ref(x, torch.ones(2))) < 1e-5
