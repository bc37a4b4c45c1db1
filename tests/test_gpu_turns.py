"""Turn detection on the device (stac_argmax_rows, stac_ctc_spikes, turns.append_speaker_turns) against the output of the
reference's own append_speaker_turns (tests/golden/turns_reference.json) and against oracle/turns.py on seeded inputs.
Index work: the comparison is exact."""
import json
import os

import numpy as np
import pytest
import torch

# First run on a B200 in round 2 (profiles/r3/r3a_first_call_verification.log): part of the default GPU suite since.
pytestmark = [pytest.mark.gpu]

from oracle import turns as oturns  # noqa: E402
from stac_speech_translation_b200 import turns  # noqa: E402
from stac_speech_translation_b200._lib import StacB200Error  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "turns_reference.json")


@pytest.mark.parametrize("as_posteriors", [False, True])
def test_rttm_lines_equal_the_reference_function(as_posteriors):
    for c in json.load(open(GOLDEN)):
        ids = torch.tensor(c["ids"], dtype=torch.int32)
        x = ids.cuda()
        if as_posteriors:
            p = torch.full(ids.shape + (c["vocab"],), -20.0)
            p.scatter_(2, ids.long()[..., None], -0.1)
            x = p.cuda()
        turn, xt = [], []
        turns.append_speaker_turns(c["utt"], x, 7, 8, turn, xt)
        assert turn == c["turn_rttm"]
        assert xt == c["xt_rttm"]


@pytest.mark.parametrize("b,t2,p", [(1, 1, 0.5), (64, 751, 0.03), (300, 257, 0.2), (700, 33, 0.0), (5, 1501, 1.0)])
def test_spikes_equal_the_oracle(b, t2, p):
    g = torch.Generator().manual_seed(b * 1000 + t2)
    ids = torch.randint(9, 5000, (b, t2), generator=g, dtype=torch.int32)
    r = torch.rand(b, t2, generator=g)
    ids[r < p / 2] = 7
    ids[(r >= p / 2) & (r < p)] = 8
    s_turn, s_xt = turns.ctc_spikes(ids.cuda(), 7, 8)
    flat = ids.reshape(-1).numpy()
    assert np.array_equal(s_turn.numpy(), np.nonzero(flat == 7)[0])
    assert np.array_equal(s_xt.numpy(), np.nonzero(flat == 8)[0])
    if b <= 64:
        utt = [f"s-r-{100 * i:07d}-x" for i in range(b)]
        want_t, want_x, got_t, got_x = [], [], [], []
        oturns.append_speaker_turns(utt, ids.numpy(), 7, 8, want_t, want_x)
        turns.append_speaker_turns(utt, ids.cuda(), 7, 8, got_t, got_x)
        assert got_t == want_t and got_x == want_x


def test_argmax_rows_first_index_on_ties():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(513, 5000, generator=g)
    x[7, 100] = x[7, 4000] = 50.0          # tie: the first index wins
    x[9, 4999] = 60.0                      # maximum in the last column
    x[11, 0] = 60.0
    small = torch.randn(40, 3, generator=g)      # fewer columns than threads
    for t in (x, small):
        ids = turns.greedy_ids(t.cuda().view(1, *t.shape))
        assert ids.dtype == torch.int32
        assert torch.equal(ids.cpu().view(-1).long(), t.argmax(-1))
    assert int(turns.greedy_ids(x.cuda().view(1, 513, 5000))[0, 7]) == 100


def test_cpu_tensors_raise():
    with pytest.raises(StacB200Error):
        turns.append_speaker_turns(["a-b-0000100-c"], torch.zeros(1, 4, dtype=torch.int32), 7, 8, [], [])
