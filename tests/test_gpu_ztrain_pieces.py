"""Training-stage pieces on the device: SpecAugment (stac_spec_augment) and the CTC loss value
(stac_ctc_loss) against the oracle's restatement of SpeechBrain's functions over this image's torch."""
import pytest
import torch

pytestmark = [pytest.mark.gpu]

import stac_speech_translation_b200 as sb  # noqa: E402
from oracle.train_pieces import SpecAugment as OracleAug, ctc_loss as oracle_ctc  # noqa: E402

REF_CFG = dict(time_warp=True, time_warp_window=5, time_warp_mode="bicubic", freq_mask=True, n_freq_mask=2, time_mask=True,
               n_time_mask=2, freq_mask_width=30, time_mask_width=40)                  # transformer_multitask.yaml:283-293


@pytest.mark.parametrize("shape", [(3, 97, 80), (8, 3001, 80), (2, 12, 80), (5, 200, 23)])
def test_spec_augment_matches_the_oracle(shape):
    b, t, f = shape
    g = torch.Generator().manual_seed(t)
    x = torch.randn(b, t, f, generator=g) * 2 - 0.5
    cfg = dict(REF_CFG, freq_mask_width=min(30, f - 1))
    ours, ref = sb.SpecAugment(**cfg), OracleAug(**cfg)
    xd = x.cuda()
    for step in range(3):
        torch.manual_seed(7 * t + step)
        want = ref(x.clone())
        torch.manual_seed(7 * t + step)
        got = ours(xd)
        assert got.is_cuda and got.shape == want.shape
        assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5), (got.cpu() - want).abs().max()
        assert torch.equal(got.cpu() == 0, want == 0)
    assert torch.equal(xd.cpu(), x)                                # the input is not touched (SpeechBrain works in place)
    for part in ("time_warp", "freq_mask", "time_mask"):
        one = dict(cfg, time_warp=False, freq_mask=False, time_mask=False)
        one[part] = True
        torch.manual_seed(11)
        want = OracleAug(**one)(x.clone())
        torch.manual_seed(11)
        assert torch.allclose(sb.SpecAugment(**one)(xd).cpu(), want, rtol=1e-5, atol=1e-5), part


@pytest.mark.parametrize("reduction", ["mean", "sum", "batchmean", "batch", "none"])
def test_ctc_loss_matches_torch(reduction):
    g = torch.Generator().manual_seed(5)
    b, t, v, lmax = 9, 751, 5000, 60                                # the S model's frame rate and vocabulary
    lp = (torch.randn(b, t, v, generator=g) * 3).log_softmax(-1)
    tg = torch.randint(1, v, (b, lmax), generator=g)
    tg[2, :6] = torch.tensor([7, 7, 7, 9, 9, 7])                    # repeated labels
    in_rel = torch.tensor([1.0, 0.93, 0.8, 0.66, 0.5, 0.31, 0.12, 0.05, 1.0])
    tg_rel = torch.tensor([1.0, 0.9, 0.1, 0.5, 0.35, 0.2, 0.1, 1.0, 0.02])     # row 7: 60 tokens in 38 frames (infeasible)
    want = oracle_ctc(lp, tg, in_rel, tg_rel, 0, reduction)
    got = sb.ctc_loss(lp.cuda(), tg.cuda(), in_rel.cuda(), tg_rel.cuda(), 0, reduction)
    assert got.is_cuda and got.shape == want.shape
    assert torch.allclose(got.cpu(), want, rtol=2e-5, atol=1e-4), (got.cpu(), want)
    per = sb.ctc_loss(lp.cuda(), tg.cuda(), in_rel.cuda(), tg_rel.cuda(), 0, "none").cpu()
    assert float(per[7]) == 0.0 and bool(torch.isfinite(per).all())   # zero_infinity


def test_ctc_loss_on_the_path_posteriors():
    """The loss of the path's own p_ctc (bf16 pipeline) equals torch's on the same tensor."""
    from stac_speech_translation_b200 import synth
    hp = sb.HParams.for_size("S", num_encoder_layers=2, num_decoder_layers=0, output_neurons=96)
    mods = sb.build_modules(hp, precision="bf16", device=torch.device("cuda", 0))
    wavs, wl = synth.synth_batch([2.0, 1.3, 0.7], seed=9)
    wavs, wl = wavs.cuda(), wl.cuda()
    mods["normalize"].calibrate(mods["compute_features"](wavs), torch.ones(3, device=wavs.device))
    res = sb.EncoderPipeline(mods)(wavs, wl)
    p = res["p_ctc"]
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(1, 96, (3, 8), generator=g)
    tl = torch.tensor([1.0, 0.75, 0.5])
    got = sb.ctc_loss(p, tokens.cuda(), wl, tl.cuda(), 0, "batchmean")
    want = oracle_ctc(p.float().cpu(), tokens, wl.cpu(), tl, 0, "batchmean")
    assert torch.allclose(got.cpu(), want, rtol=2e-5, atol=1e-4)
