"""TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else): CPU restatement of the reference's turn detection.

Follows ``append_speaker_turns``, /root/reference/stac-st/inference.py:54-84, line by line (argmax :58, the two masks
:59-63, the per-sample / per-frame loop :65-84, the RTTM formatting :73-80).  Pinned by the reference itself:
tests/golden/make_turns_golden.py executes the reference's own function body (extracted from inference.py with `ast`,
because importing that script needs speechbrain) and tests/test_oracle.py compares this restatement with its output.
"""
from __future__ import annotations

import numpy as np

DOWNSAMPLING = 25          # inference.py:46-48


def append_speaker_turns(batch_ids, model_ctc_outputs, turn, xt, turn_rttm, xt_rttm):
    ids = np.asarray(model_ctc_outputs)
    if ids.ndim == 3:
        ids = ids.argmax(-1)                               # :58
    pred_turn = (ids == turn).astype(int)                  # :59, :62
    pred_xt = (ids == xt).astype(int)                      # :60, :63
    for sample_idx in range(len(ids)):                     # :65
        cnt = 0
        utt_id = batch_ids[sample_idx]
        abs_start = int(utt_id.split("-")[2]) / 100.0      # :69
        for turn_sample, xt_sample in zip(pred_turn[sample_idx], pred_xt[sample_idx]):
            start = cnt * (1 / DOWNSAMPLING)               # :72
            if turn_sample == 1:
                turn_rttm.append(
                    f"SPEAKER {utt_id} 1 {abs_start + start:.3f} {(1/DOWNSAMPLING)} <NA> <NA> SPK1 <NA> <NA>")
            if xt_sample == 1:
                xt_rttm.append(
                    f"SPEAKER {utt_id} 1 {abs_start + start:.3f} {(1/DOWNSAMPLING)} <NA> <NA> SPK1 <NA> <NA>")
            cnt += 1
