"""The oracle against independent formulations, the reference's pinned constants and the golden file.

The reference ships no tests or vectors, so each stage of the restatement is
cross-checked here against a second, differently-written formulation (numpy DFT, explicit loops,
hand-written attention), plus the three constants the reference does pin:
5120-wide CNN output (transformer_multitask.yaml:184), 25 frames/s (inference.py:46-48),
2500-entry positional table (TransformerMultiTask.py:108)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

import oracle
from oracle import speechbrain_path as sp
from stac_speech_translation_b200 import synth
from util import TINY, oracle_modules, rel_l2


@pytest.fixture(scope="module")
def tiny():
    return oracle_modules(TINY, vocab=64)


def test_pinned_shape_constants(tiny):
    for secs, (t, t1, t2) in {10: (1001, 501, 251), 30: (3001, 1501, 751), 60: (6001, 3001, 1501)}.items():
        assert synth.frames_of(secs * 16000) == (t, t1, t2)
        assert round(t2 / secs) == 25                           # DOWNSAMPLING = 25 (inference.py:48)
    wavs, wl = synth.synth_batch([1.0], seed=0)
    out = oracle.reference_compute_forward(tiny, wavs, wl, stages="frontend")
    assert out["fbank"].shape == (1, 101, 80)
    assert out["cnn"].shape == (1, 26, 20, 256) and 20 * 256 == 5120
    assert tiny["Transformer"].positional_encoding.pe.shape[1] == 2500


def test_stft_against_numpy_dft():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 1600, generator=g)
    stft = sp.STFT(16000)(x)                                        # [B, T, 201, 2]
    assert stft.shape == (2, 11, 201, 2)
    n = np.arange(400)
    win = 0.54 - 0.46 * np.cos(2 * np.pi * n / 400)                 # periodic hamming
    xp = np.pad(x.numpy().astype(np.float64), ((0, 0), (200, 200)))
    for t in (0, 1, 5, 10):
        frame = xp[:, 160 * t:160 * t + 400] * win
        ref = np.fft.rfft(frame, axis=-1)
        got = stft[:, t, :, 0].numpy() + 1j * stft[:, t, :, 1].numpy()
        assert np.abs(got - ref).max() < 1e-3 * np.abs(ref).max()


def test_mel_matrix_structure():
    fb = sp.Filterbank(n_mels=80).fbank_matrix()                     # [201, 80]
    assert fb.shape == (201, 80)
    nnz = (fb > 0).sum(0)
    assert int((fb > 0).sum()) == 387 and int(nnz.min()) >= 1 and int(nnz.max()) == 13
    assert int(torch.nonzero(fb.sum(1) > 0).max()) == 199
    # symmetric triangle with the LEFT band as half width, peak <= 1 at the centre frequency
    mel = np.linspace(0, 2595 * math.log10(1 + 8000 / 700), 82)
    hz = 700 * (10 ** (mel / 2595) - 1)
    freqs = np.linspace(0, 8000, 201)
    for m in (0, 17, 79):
        tri = np.maximum(0, 1 - np.abs(freqs - hz[m + 1]) / (hz[m + 1] - hz[m]))
        assert np.abs(fb[:, m].numpy() - tri).max() < 1e-4


def test_fbank_db_and_topdb():
    wavs, _ = synth.synth_batch([1.0, 0.5], seed=3)
    fb = oracle.Fbank(n_mels=80)
    x = fb(wavs)
    power = sp.spectral_magnitude(fb.compute_STFT(wavs))
    raw = 10 * torch.log10(torch.clamp(power @ fb.compute_fbanks.fbank_matrix(), min=1e-10))
    for b in range(2):
        assert torch.allclose(x[b], torch.maximum(raw[b], raw[b].max() - 80))
    # padded tail of the short utterance is silence: -100 dB clamped to (utterance max - 80)
    assert torch.allclose(x[1, 60:], (raw[1].max() - 80).expand_as(x[1, 60:]))
    fb_g = oracle.Fbank(n_mels=80, top_db_per_utterance=False)
    assert torch.allclose(fb_g(wavs), torch.maximum(raw, raw.max() - 80))


def test_input_normalization_semantics():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 50, 80, generator=g) * 3 + 1
    lens = torch.tensor([1.0, 0.5, 0.26])
    norm = oracle.InputNormalization(norm_type="global", update_until_epoch=4)
    norm.train()
    y = norm(x.clone(), lens)
    ns = [50, 25, 13]
    mean = torch.stack([x[i, :n].mean(0) for i, n in enumerate(ns)]).mean(0)
    std = torch.stack([x[i, :n].std(0) for i, n in enumerate(ns)]).mean(0)
    assert torch.allclose(norm.glob_mean, mean) and torch.allclose(norm.glob_std, std) and norm.count == 1
    assert torch.allclose(y, (x - mean) / std, atol=1e-6)
    x2 = torch.randn(3, 50, 80, generator=g)
    norm(x2.clone(), lens, epoch=1)                                  # running average, weight 1/2
    m2 = torch.stack([x2[i, :n].mean(0) for i, n in enumerate(ns)]).mean(0)
    assert torch.allclose(norm.glob_mean, 0.5 * mean + 0.5 * m2, atol=1e-6)
    norm.eval()
    before = norm.glob_mean.clone()
    y3 = norm(x.clone(), lens)
    assert torch.equal(norm.glob_mean, before)                       # eval never updates
    assert torch.allclose(y3, (x - norm.glob_mean) / norm.glob_std, atol=1e-6)
    norm(x.clone(), lens, epoch=7)                                   # still eval
    fresh = oracle.InputNormalization()
    assert fresh.glob_std.item() == 0                                # eval before any statistics divides by zero


def test_conv_block_against_explicit_loops(tiny):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 9, 80, generator=g)
    blk = tiny["CNN"].convblock_0
    got = blk.convs.conv_0(x)                                        # [1, 5, 40, 256]
    assert got.shape == (1, 5, 40, 256)
    w, b = blk.convs.conv_0.conv.weight, blk.convs.conv_0.conv.bias  # [256, 1, kF, kT]
    refl = lambda i, n: -i if i < 0 else (2 * (n - 1) - i if i >= n else i)
    for (t1, f1, c) in [(0, 0, 0), (4, 39, 255), (2, 17, 100), (4, 0, 7)]:
        acc = b[c].item()
        for kf in range(3):
            for kt in range(3):
                acc += w[c, 0, kf, kt].item() * x[0, refl(2 * t1 + kt - 1, 9), refl(2 * f1 + kf - 1, 80)].item()
        assert abs(got[0, t1, f1, c].item() - acc) < 1e-4
    y = blk.convs.norm_0(got)
    flat = got.reshape(1, 5, -1)
    ref = (flat - flat.mean(-1, keepdim=True)) / torch.sqrt(flat.var(-1, unbiased=False, keepdim=True) + 1e-5)
    ref = ref.view_as(got) * blk.convs.norm_0.norm.weight + blk.convs.norm_0.norm.bias
    assert torch.allclose(y, ref, atol=1e-5)
    out = tiny["CNN"](x)
    assert out.shape == (1, 3, 20, 256)


def test_mask_rules_floor_vs_round():
    t2 = 26
    wl = torch.tensor([1.0, 0.52, 0.5, 0.1])
    keep_encode = ~(torch.arange(t2)[None, :].float() > torch.floor(wl * t2)[:, None])
    keep_train = sp.length_to_mask(torch.round(wl * t2), max_len=t2)
    assert keep_encode.sum(1).tolist() == [26, 14, 14, 3]            # floor(.)+1 frames, capped at T
    assert keep_train.sum(1).tolist() == [26, 14, 13, 3]             # round(.) frames (13.52->14, 13.0->13, 2.6->3)


def test_encoder_layer_against_handwritten_math(tiny):
    tr = tiny["Transformer"]
    layer = tr.encoder.layers[0]
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 11, 128, generator=g)
    kpm = torch.zeros(2, 11, dtype=torch.bool)
    kpm[1, 6:] = True
    got, w = layer(x, src_key_padding_mask=kpm)
    att = layer.self_att.att
    ln = lambda v, n: torch.nn.functional.layer_norm(v, (128,), n.norm.weight, n.norm.bias, 1e-6)
    h = ln(x, layer.norm1)
    q, k, v = (h @ att.in_proj_weight.T + att.in_proj_bias).split(128, -1)
    heads = lambda z: z.view(2, 11, 2, 64).transpose(1, 2)
    s = (heads(q) / 8.0) @ heads(k).transpose(-1, -2)
    s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    ctx = (p @ heads(v)).transpose(1, 2).reshape(2, 11, 128)
    y = x + ctx @ att.out_proj.weight.T + att.out_proj.bias
    f = layer.pos_ffn.ffn
    h2 = ln(y, layer.norm2)
    z = torch.nn.functional.gelu(h2 @ f[0].weight.T + f[0].bias) @ f[3].weight.T + f[3].bias
    assert rel_l2(got, y + z) < 1e-5
    assert torch.allclose(w, p.mean(1), atol=1e-5)                   # head-averaged weights (discarded by encode)


def test_init_distribution(tiny):
    # _init_params: xavier_normal_ on Transformer params with dim > 1; CNN / ctc_lin keep nn defaults
    w = tiny["Transformer"].encoder.layers[0].pos_ffn.ffn[0].weight
    assert abs(w.std().item() - math.sqrt(2.0 / (128 + 512))) < 0.1 * math.sqrt(2.0 / (128 + 512))
    c = tiny["ctc_lin"].w.weight
    assert c.abs().max().item() <= 1 / math.sqrt(128) + 1e-6


def test_golden_vectors():
    import make_golden
    gold = np.load(make_golden.GOLDEN)
    wavs, wl, out, out_t = make_golden.generate()
    assert np.allclose(gold["wav_checksum"], [float(wavs.double().sum()), float(wavs.double().abs().sum())], rtol=1e-9)
    assert np.array_equal(gold["wav_lens"], wl.numpy())
    for k in ("fbank", "feats", "enc_out", "p_ctc"):
        assert rel_l2(out[k], torch.from_numpy(gold[k])) < 1e-5, k
    assert rel_l2(out["cnn"][:, ::7, ::3, ::16], torch.from_numpy(gold["cnn_sample"])) < 1e-5
    assert rel_l2(out_t["enc_out"], torch.from_numpy(gold["enc_out_train_mask"])) < 1e-5
    assert rel_l2(out["enc_out"], out_t["enc_out"]) > 1e-6           # the two mask rules differ on this batch


# --------------------------------------------------------------------------
# Reference-generated fixture: tests/golden/glue_reference.npz was produced by executing the
# reference's own stac-st/modules/TransformerMultiTask.py (tests/golden/make_glue_golden.py).
# --------------------------------------------------------------------------
def _glue_fixture():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "glue_reference.npz"))
    state = {k[len("state/"):]: torch.from_numpy(d[k].astype(np.float32)) for k in d.files if k.startswith("state/")}
    return d, state


def test_oracle_glue_against_reference_generated_fixture():
    d, state = _glue_fixture()
    d_model = state["encoder.norm.norm.weight"].shape[0]
    tr = sp.TransformerMultiTask(tgt_vocab=64, input_size=state["custom_src_module.layers.0.w.weight"].shape[1],
                                 d_model=d_model, nhead=d_model // 64,
                                 num_encoder_layers=2, d_ffn=state["encoder.layers.0.pos_ffn.ffn.0.weight"].shape[0],
                                 activation=torch.nn.GELU, normalize_before=True).eval()
    res = tr.load_state_dict(state, strict=False)
    # (the fixture holds the encoder side only; the decoder side has its own, decoder_reference.npz)
    assert not res.unexpected_keys and all(
        k == "positional_encoding.pe" or k.startswith(("decoder.", "custom_tgt_module.")) for k in res.missing_keys)
    src = torch.from_numpy(d["src"].astype(np.float32))
    wl = torch.from_numpy(d["wav_lens"])
    torch.set_num_threads(1)
    with torch.no_grad():
        assert rel_l2(tr.encode(src, wl), torch.from_numpy(d["enc_encode"])) < 1e-6
        assert rel_l2(tr.encode(src.reshape(src.shape[0], src.shape[1], -1)), torch.from_numpy(d["enc_encode_nolen"])) < 1e-6
        assert rel_l2(tr.forward_encoder(src, wl), torch.from_numpy(d["enc_forward"])) < 1e-6
        assert rel_l2(sp.EncoderWrapper(tr)(src, wl), torch.from_numpy(d["enc_encode"])) < 1e-6
    # the two mask rules really differ on this fixture (0.5 * 37 = 18.5: floor+1 = 19 keys vs round = 18)
    assert rel_l2(torch.from_numpy(d["enc_forward"])[2], torch.from_numpy(d["enc_encode"])[2]) > 1e-4


def test_turn_detection_oracle_matches_the_reference_function():
    """oracle/turns.py against the output of the reference's own append_speaker_turns (inference.py:54-84), recorded by
    tests/golden/make_turns_golden.py; ids and one-hot-like posteriors give the same lines."""
    import json
    from oracle import turns as oturns
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "turns_reference.json")))
    assert len(cases) >= 4
    for c in cases:
        ids = np.asarray(c["ids"])
        for as_posteriors in (False, True):
            x = ids
            if as_posteriors:
                x = np.full(ids.shape + (c["vocab"],), -20.0, dtype=np.float32)
                np.put_along_axis(x, ids[..., None], -0.1, axis=2)
            turn, xt = [], []
            oturns.append_speaker_turns(c["utt"], x, 7, 8, turn, xt)
            assert turn == c["turn_rttm"] and xt == c["xt_rttm"]


def test_ctc_peaky_fixture_decodes():
    """tests/golden/ctc_peaky.npz (make_ctc_peaky_fixture.py): loaded into the oracle, the fitted CTC head decodes the
    tone-coded utterances to the symbols they encode, with posteriors that are peaky (clear top-2 margins) - the
    precondition of the >= 99.5 % greedy-sequence criterion tested on the GPU (tests/test_gpu_ctc_peaky.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_ctc_peaky_fixture import load_peaky, synth_tone_batches, SYMBOLS
    from stac_speech_translation_b200.pipeline import ctc_greedy_collapse
    from util import oracle_modules
    omods = load_peaky(oracle_modules("S"))
    wavs, wl, targets, _ = synth_tone_batches()[0]                  # the batch of the 24 shortest utterances
    ref = oracle.reference_compute_forward(omods, wavs, wl)
    t2 = ref["p_ctc"].shape[1]
    n_valid = (torch.floor(wl * t2) + 1).clamp(max=t2).long().tolist()
    ids = ref["p_ctc"].argmax(-1)
    assert ctc_greedy_collapse(ids, n_valid) == targets
    assert set(ids.unique().tolist()) <= set(SYMBOLS)
    top2 = ref["p_ctc"].topk(2, dim=-1).values
    margin = torch.cat([(top2[i, :n, 0] - top2[i, :n, 1]) for i, n in enumerate(n_valid)])
    assert float((margin > 1.0).float().mean()) > 0.9          # peaky: nine frames in ten decided by > 1 nat
