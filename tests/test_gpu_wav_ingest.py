"""int16 PCM ingest on the device (stac_pcm_i16_to_f32, ingest.PcmStager) against the decode rule of a 16-bit file
(sample / 32768): exact."""
import numpy as np
import os

import pytest
import torch

# First run on a B200 in round 2 (profiles/r3/r3a_first_call_verification.log): part of the default GPU suite since.
pytestmark = [pytest.mark.gpu]

from stac_speech_translation_b200 import ingest  # noqa: E402


@pytest.mark.parametrize("n", [1, 7, 8, 9, 65536, 64 * 480000 + 3])
def test_pcm_to_float_is_exact(n):
    g = torch.Generator().manual_seed(n)
    pcm = torch.randint(-32768, 32768, (n,), generator=g, dtype=torch.int16)
    if n >= 4:
        pcm[:4] = torch.tensor([-32768, 32767, 0, -1], dtype=torch.int16)
    guard = torch.full((64,), 3.0, device="cuda")
    got = ingest.pcm_to_float(pcm.cuda())
    assert got.dtype == torch.float32 and torch.equal(got.cpu(), pcm.float() / 32768.0)
    assert (guard == 3.0).all()


def test_stager_double_buffering_delivers_every_batch():
    rng = np.random.default_rng(0)
    batches = [[rng.integers(-32768, 32768, int(n), dtype=np.int16) for n in rng.integers(100, 5000, size=b)]
               for b in (4, 1, 7, 3, 5)]
    st = ingest.PcmStager(max_elems=7 * 5000)
    st.put(batches[0])
    for i, batch in enumerate(batches):
        if i + 1 < len(batches):
            st.put(batches[i + 1])                      # next batch is staged while this one is consumed
        wavs, wl = st.get()
        pcm, want_wl = ingest.collate(batch)
        assert torch.equal(wavs.cpu(), pcm.float() / 32768.0) and torch.equal(wl.cpu(), want_wl)
