"""north_star's greedy criterion: "bf16 mode ... identical CTC-greedy token sequences on >= 99.5 % of utterances".

Needs posteriors that are peaky the way a trained model's are, so the oracle carries the fitted CTC head of
tests/golden/ctc_peaky.npz (tests/golden/make_ctc_peaky_fixture.py: tone-coded synthetic audio, top of the S model
fitted with CTC on the CPU).  The same weights go into the product through state_dict; both run the same 240
utterances in the same 10 length-sorted batches; CTC-greedy sequences (merge repeats, drop blanks, valid frames only)
must be identical on >= 99.5 % of the utterances.  Frame flip rate and the oracle's top-2 margin histogram are
written to gpurun_out/ctc_peaky_agreement.json (copied to profiles/ when the numbers are quoted)."""
import json
import os
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu]

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import oracle  # noqa: E402
import stac_speech_translation_b200 as sb  # noqa: E402
from make_ctc_peaky_fixture import load_peaky, synth_tone_batches  # noqa: E402
from stac_speech_translation_b200.pipeline import ctc_greedy_collapse  # noqa: E402
from util import BF16_TOL, FP32_TOL, oracle_modules, product_from_oracle, rel_l2  # noqa: E402


def _run(precision):
    torch.set_num_threads(os.cpu_count() or 1)
    omods = load_peaky(oracle_modules("S"))
    mods = product_from_oracle(omods, precision)
    pipe = sb.EncoderPipeline(mods)
    stats = {"utterances": 0, "identical": 0, "frames": 0, "flipped": 0, "decoded_is_target": 0,
             "enc_rel_l2_max": 0.0, "pctc_rel_l2_max": 0.0, "max_margin_of_flipped_frame": 0.0}
    edges = [0.0, 0.05, 0.1, 0.2, 0.5, 1.0, 2.0, 5.0, 1e9]
    hist = [0] * (len(edges) - 1)
    for wavs, wl, targets, _ in synth_tone_batches():
        ref = oracle.reference_compute_forward(omods, wavs, wl)
        res = pipe(wavs.cuda(), wl.cuda())
        torch.cuda.synchronize()
        t2 = ref["p_ctc"].shape[1]
        n_valid = (torch.floor(wl * t2) + 1).clamp(max=t2).long().tolist()     # frames encode() keeps
        ref_ids = ref["p_ctc"].argmax(-1)
        ids = res["greedy"].cpu().long()
        top2 = ref["p_ctc"].topk(2, dim=-1).values
        margin = top2[..., 0] - top2[..., 1]
        a, b = ctc_greedy_collapse(ids, n_valid), ctc_greedy_collapse(ref_ids, n_valid)
        for i, n in enumerate(n_valid):
            stats["utterances"] += 1
            stats["identical"] += a[i] == b[i]
            stats["decoded_is_target"] += b[i] == targets[i]
            stats["frames"] += n
            flips = ids[i, :n] != ref_ids[i, :n]
            stats["flipped"] += int(flips.sum())
            if flips.any():
                stats["max_margin_of_flipped_frame"] = max(stats["max_margin_of_flipped_frame"],
                                                           float(margin[i, :n][flips].max()))
            h = torch.histogram(margin[i, :n].float(), torch.tensor(edges))[0]
            hist = [x + int(y) for x, y in zip(hist, h)]
            stats["enc_rel_l2_max"] = max(stats["enc_rel_l2_max"], rel_l2(res["enc_out"][i, :n], ref["enc_out"][i, :n]))
            stats["pctc_rel_l2_max"] = max(stats["pctc_rel_l2_max"], rel_l2(res["p_ctc"][i, :n], ref["p_ctc"][i, :n]))
    stats["sequence_agreement"] = stats["identical"] / stats["utterances"]
    stats["frame_flip_rate"] = stats["flipped"] / stats["frames"]
    stats["oracle_margin_histogram"] = {f"[{edges[k]}, {edges[k + 1]})": hist[k] for k in range(len(hist))}
    stats["precision"] = precision
    return stats


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_ctc_greedy_sequences_on_peaky_posteriors(precision):
    stats = _run(precision)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/ctc_peaky_agreement_{precision}.json", "w") as f:
        json.dump(stats, f, indent=1)
    print(json.dumps(stats))
    assert stats["utterances"] >= 200
    # the fixture is meaningful: the oracle decodes what the audio encodes
    assert stats["decoded_is_target"] >= 0.95 * stats["utterances"], stats
    tol = BF16_TOL if precision == "bf16" else FP32_TOL
    assert stats["enc_rel_l2_max"] < tol and stats["pctc_rel_l2_max"] < tol, stats
    assert stats["sequence_agreement"] >= 0.995, stats          # north_star: >= 99.5 % of utterances
