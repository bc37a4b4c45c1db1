"""The library's plain SIMT kernels, compiled from their real source for a CPU emulation of the CUDA execution model
(tests/simt_cpu: blocks one after another, threads of a block as OS threads, barriers for __syncthreads and the
warp-synchronous primitives), driven through the same extern "C" entry points on host memory.

This is how the kernels that were written after the round's GPU budget was spent (turn detection, decoder building
blocks, PCM ingest, train-mode normalisation statistics) have been EXECUTED so far; their first run on a B200 is
tests/test_gpu_{turns,decoder,wav_ingest,ytrain_norm}.py.  Test infrastructure: nothing of it ships."""
import ctypes
import json
import os
import sys
from ctypes import c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt_cpu"))
import build as simt_build  # noqa: E402

from stac_speech_translation_b200 import _lib  # noqa: E402


@pytest.fixture(scope="module")
def simt():
    lib = ctypes.CDLL(str(simt_build.build()))
    for name in ("stac_argmax_rows", "stac_ctc_spikes", "stac_embed_scale_pe", "stac_attention_f32",
                 "stac_pcm_i16_to_f32", "stac_utt_mean_std", "stac_layernorm", "stac_mha_f32", "stac_log_softmax",
                 "stac_kv_lengths", "stac_cast_bf16", "stac_gemm_f32", "stac_spec_augment", "stac_ctc_loss",
                 "stac_attention_beam_f32", "stac_attention_step_f32", "stac_embed_step"):
        res, args = _lib._SIGNATURES[name]
        getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    return lib


def P(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def test_turn_detection_kernels(simt):
    for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "turns_reference.json"))):
        ids = torch.tensor(c["ids"], dtype=torch.int32)
        b, t2 = ids.shape
        counts = torch.empty(b, 2, dtype=torch.int32)
        spikes = torch.full((2, b * t2), -1, dtype=torch.int32)
        n_out = torch.empty(2, dtype=torch.int32)
        assert simt.stac_ctc_spikes(P(ids), b, t2, 7, 8, P(counts), P(spikes[0]), P(spikes[1]), P(n_out), None) == 0
        flat = ids.reshape(-1).numpy()
        for k, tok in enumerate((7, 8)):
            want = np.nonzero(flat == tok)[0]
            assert int(n_out[k]) == len(want) and np.array_equal(spikes[k, :len(want)].numpy(), want)
            assert (spikes[k, len(want):] == -1).all()                      # nothing written past the count
    g = torch.Generator().manual_seed(5)
    x = torch.randn(70, 300, generator=g)
    x[7, 100] = x[7, 250] = 50.0                                            # tie: the first index wins
    x[9, 299] = 60.0
    out = torch.empty(70, dtype=torch.int32)
    assert simt.stac_argmax_rows(P(x), 70, 300, P(out), None) == 0
    assert torch.equal(out.long(), x.argmax(-1)) and int(out[7]) == 100
    small = torch.randn(5, 3, generator=g)                                  # fewer columns than threads
    out = torch.empty(5, dtype=torch.int32)
    assert simt.stac_argmax_rows(P(small), 5, 3, P(out), None) == 0 and torch.equal(out.long(), small.argmax(-1))


def test_spike_compaction_across_many_rows(simt):
    g = torch.Generator().manual_seed(2)
    b, t2 = 300, 70                                                         # row offsets need more than one pass of 256
    ids = torch.randint(9, 50, (b, t2), generator=g, dtype=torch.int32)
    r = torch.rand(b, t2, generator=g)
    ids[r < 0.1] = 7
    ids[(r >= 0.1) & (r < 0.2)] = 8
    counts, n_out = torch.empty(b, 2, dtype=torch.int32), torch.empty(2, dtype=torch.int32)
    spikes = torch.empty(2, b * t2, dtype=torch.int32)
    assert simt.stac_ctc_spikes(P(ids), b, t2, 7, 8, P(counts), P(spikes[0]), P(spikes[1]), P(n_out), None) == 0
    flat = ids.reshape(-1).numpy()
    assert np.array_equal(spikes[0, :int(n_out[0])].numpy(), np.nonzero(flat == 7)[0])
    assert np.array_equal(spikes[1, :int(n_out[1])].numpy(), np.nonzero(flat == 8)[0])


@pytest.mark.parametrize("rows,lq,lk,h,div,causal", [(1, 1, 1, 1, 1, 0), (3, 4, 4, 2, 1, 1), (4, 2, 37, 2, 2, 0),
                                                     (2, 1, 300, 1, 1, 0)])
def test_attention_f32_kernel(simt, rows, lq, lk, h, div, causal):
    g = torch.Generator().manual_seed(rows * 100 + lk)
    d = 64 * h
    n_mem = rows // div
    q = torch.randn(rows * lq, d, generator=g)
    kv = torch.randn(n_mem * lk, 2 * d, generator=g)
    kv_len = torch.randint(1, lk + 1, (rows,), generator=g, dtype=torch.int32)
    tok = torch.randint(0, 3, (rows, lk), generator=g)
    tok[:, 0] = 1
    use_tok = causal == 1
    ctx = torch.full((rows * lq, d), float("nan"))
    w = torch.full((rows, lq, lk), float("nan"))
    rc = simt.stac_attention_f32(P(q), d, P(kv), c_void_p(kv.data_ptr() + 4 * d), lk * 2 * d, 2 * d, rows, lq, lk, h, div,
                                 causal, P(kv_len), P(tok) if use_tok else None, 0, P(ctx), d, P(w), None)
    assert rc == 0
    qq = q.view(rows, lq, h, 64).permute(0, 2, 1, 3).double()
    k = kv[:, :d].reshape(n_mem, lk, h, 64).permute(0, 2, 1, 3).double().repeat_interleave(div, 0)
    v = kv[:, d:].reshape(n_mem, lk, h, 64).permute(0, 2, 1, 3).double().repeat_interleave(div, 0)
    mask = torch.arange(lk)[None, None, :] >= kv_len[:, None, None]
    if causal:
        mask = mask | (torch.arange(lk)[None, None, :] > torch.arange(lq)[None, :, None])
    if use_tok:
        mask = mask | (tok == 0)[:, None, :]
    p = torch.softmax((qq @ k.transpose(-1, -2)).masked_fill(mask[:, None], float("-inf")), -1)
    want = (p @ v).permute(0, 2, 1, 3).reshape(rows * lq, d)
    assert (ctx.double() - want).norm() / want.norm() < 1e-5
    assert (w.double() - p.mean(1)).norm() / p.mean(1).norm() < 1e-5


@pytest.mark.parametrize("rows,lk,h,div,with_len", [(5, 1, 1, 1, False), (6, 33, 2, 2, True), (3, 256, 4, 1, True),
                                                    (9, 70, 4, 3, False)])
def test_attention_step_warp_kernel(simt, rows, lk, h, div, with_len):
    """lq = 1 without weights / token mask and at most 256 keys: stac_attention_f32 runs the warp-per-(row, head) kernel
    (rows * h not a multiple of the eight warps of a CTA: idle warps leave)."""
    g = torch.Generator().manual_seed(rows * 1000 + lk)
    d = 64 * h
    n_mem = rows // div
    q = torch.randn(rows, d, generator=g)
    kv = torch.randn(n_mem * lk, 2 * d, generator=g)
    kv_len = torch.randint(1, lk + 1, (rows,), generator=g, dtype=torch.int32)
    ctx = torch.full((rows, d), float("nan"))
    rc = simt.stac_attention_f32(P(q), d, P(kv), c_void_p(kv.data_ptr() + 4 * d), lk * 2 * d, 2 * d, rows, 1, lk, h, div,
                                 0, P(kv_len) if with_len else None, None, 0, P(ctx), d, None, None)
    assert rc == 0
    qq = q.view(rows, 1, h, 64).permute(0, 2, 1, 3).double()
    k = kv[:, :d].reshape(n_mem, lk, h, 64).permute(0, 2, 1, 3).double().repeat_interleave(div, 0)
    v = kv[:, d:].reshape(n_mem, lk, h, 64).permute(0, 2, 1, 3).double().repeat_interleave(div, 0)
    sc = qq @ k.transpose(-1, -2)
    if with_len:
        sc = sc.masked_fill((torch.arange(lk)[None, :] >= kv_len[:, None])[:, None, None, :], float("-inf"))
    want = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(rows, d)
    assert (ctx.double() - want).norm() / want.norm() < 1e-5


@pytest.mark.parametrize("rows,lk,h,with_map", [(3, 5, 2, False), (5, 40, 2, True), (2, 300, 1, True), (3, 530, 2, False)])
def test_attention_step_kernel_with_row_map(simt, rows, lk, h, with_map):
    """stac_attention_step_f32: time-major cache [max_len][rows][2 d], hypothesis rows found through the row map (lazy
    beam re-ordering); caches longer than 256 keys run chunk by chunk with a running maximum / sum."""
    g = torch.Generator().manual_seed(rows * 1000 + lk)
    d = 64 * h
    q = torch.randn(rows, d, generator=g)
    cache = torch.randn(lk + 2, rows, 2 * d, generator=g)                   # (two unused positions behind the prefix)
    rmap = torch.randint(0, rows, (lk, rows), generator=g, dtype=torch.int32) if with_map else None
    ctx = torch.full((rows, d), float("nan"))
    rc = simt.stac_attention_step_f32(P(q), d, P(cache), c_void_p(cache.data_ptr() + 4 * d), 2 * d, rows * 2 * d, rows, lk,
                                      h, P(rmap), None, None, 0, None, P(ctx), d, None)
    assert rc == 0
    # append mode: the last position comes from the projection's rows, the counter from memory; same result, and the
    # kernel has stored the new keys / values into slab lk - 1
    cache2 = cache.clone()
    kv_new = cache[lk - 1].clone()
    cache2[lk - 1] = float("nan")
    if with_map:
        rmap[lk - 1] = torch.arange(rows, dtype=torch.int32)        # a new position always lives in its own row
    t_dev = torch.tensor([lk - 1], dtype=torch.int32)
    ctx_b = torch.full((rows, d), float("nan"))
    rc = simt.stac_attention_step_f32(P(q), d, P(cache2), c_void_p(cache2.data_ptr() + 4 * d), 2 * d, rows * 2 * d, rows,
                                      lk + 2, h, P(rmap), P(kv_new), c_void_p(kv_new.data_ptr() + 4 * d), 2 * d, P(t_dev),
                                      P(ctx_b), d, None)
    assert rc == 0 and torch.equal(cache2[lk - 1], kv_new)
    if with_map:                                                    # (the first call read slab lk - 1 through the old map)
        ctx = torch.full((rows, d), float("nan"))
        assert simt.stac_attention_step_f32(P(q), d, P(cache), c_void_p(cache.data_ptr() + 4 * d), 2 * d, rows * 2 * d,
                                            rows, lk, h, P(rmap), None, None, 0, None, P(ctx), d, None) == 0
    assert torch.allclose(ctx_b, ctx, rtol=1e-5, atol=1e-6)       # (the new key opens the running softmax: other order)
    src = rmap.long() if with_map else torch.arange(rows).repeat(lk, 1)
    kv = cache[:lk].gather(1, src[:, :, None].expand(lk, rows, 2 * d))          # [lk, rows, 2d] as each row sees it
    kk = kv[:, :, :d].permute(1, 0, 2).reshape(rows, lk, h, 64).permute(0, 2, 1, 3).double()
    vv = kv[:, :, d:].permute(1, 0, 2).reshape(rows, lk, h, 64).permute(0, 2, 1, 3).double()
    p = torch.softmax(q.view(rows, 1, h, 64).permute(0, 2, 1, 3).double() @ kk.transpose(-1, -2), -1)
    want = (p @ vv).permute(0, 2, 1, 3).reshape(rows, d)
    assert (ctx.double() - want).norm() / want.norm() < 1e-5


def test_attention_time_major_cache_and_embedding(simt):
    """The addressing DecoderCache uses: keys / values in a time-major cache [lk][rows][2 d]; and the embedding kernel
    with a position offset into the table."""
    g = torch.Generator().manual_seed(9)
    rows, lk, h = 3, 5, 2
    d = 64 * h
    q = torch.randn(rows, d, generator=g)
    cache = torch.randn(8, rows, 2 * d, generator=g)                        # max_len 8, only lk rows are valid
    ctx = torch.empty(rows, d)
    rc = simt.stac_attention_f32(P(q), d, P(cache), c_void_p(cache.data_ptr() + 4 * d), 2 * d, rows * 2 * d, rows, 1, lk,
                                 h, 1, 0, None, None, 0, P(ctx), d, None, None)
    assert rc == 0
    kk = cache[:lk, :, :d].permute(1, 0, 2).reshape(rows, lk, h, 64).permute(0, 2, 1, 3).double()
    vv = cache[:lk, :, d:].permute(1, 0, 2).reshape(rows, lk, h, 64).permute(0, 2, 1, 3).double()
    p = torch.softmax(q.view(rows, 1, h, 64).permute(0, 2, 1, 3).double() @ kk.transpose(-1, -2), -1)
    want = (p @ vv).permute(0, 2, 1, 3).reshape(rows, d)
    assert (ctx.double() - want).norm() / want.norm() < 1e-5
    vocab, L, r = 97, 5, 4
    emb, pe = torch.randn(vocab, d, generator=g), torch.randn(40, d, generator=g)
    tok = torch.randint(0, vocab, (r, L), generator=g)
    out = torch.empty(r * L, d)
    assert simt.stac_embed_scale_pe(P(tok), P(emb), P(pe), r * L, L, d, vocab, float(np.sqrt(d)), P(out), None) == 0
    assert torch.allclose(out.view(r, L, d), emb[tok] * np.float32(np.sqrt(d)) + pe[:L][None], atol=1e-5)
    out1 = torch.empty(r, d)                                                # one position (t = 3) for every row
    tok3 = tok[:, 3].contiguous()                                           # (kept alive across the call)
    rc = simt.stac_embed_scale_pe(P(tok3), P(emb), c_void_p(pe.data_ptr() + 3 * d * 4), r, 1, d, vocab,
                                  float(np.sqrt(d)), P(out1), None)
    assert rc == 0 and torch.allclose(out1, out.view(r, L, d)[:, 3], atol=1e-6)
    out2, pos = torch.empty(r, d), torch.tensor([3], dtype=torch.int32)     # the same with the position read from memory
    assert simt.stac_embed_step(P(tok3), P(emb), P(pe), r, d, vocab, float(np.sqrt(d)), P(pos), P(out2), None) == 0
    assert torch.equal(out2, out1)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 4099])
def test_pcm_to_float(simt, n):
    pcm = torch.randint(-32768, 32768, (n,), dtype=torch.int16)
    if n >= 4:
        pcm[:4] = torch.tensor([-32768, 32767, 0, -1], dtype=torch.int16)
    out = torch.full((n + 8,), 3.0)
    assert simt.stac_pcm_i16_to_f32(P(pcm), n, P(out), None) == 0
    assert torch.equal(out[:n], pcm.float() / 32768.0) and (out[n:] == 3.0).all()


def test_utt_mean_std(simt):
    g = torch.Generator().manual_seed(4)
    b, t, f = 4, 57, 80
    x = torch.randn(b, t, f, generator=g) * 3 + 1
    wl = torch.tensor([1.0, 0.73, 0.41, 0.5])
    mean, std = torch.empty(b, f), torch.empty(b, f)
    assert simt.stac_utt_mean_std(P(x), P(wl), b, t, f, 1e-10, P(mean), P(std), None) == 0
    for i in range(b):
        n = int(torch.round(wl[i] * t))
        assert torch.allclose(mean[i], x[i, :n].mean(0), atol=1e-5)
        assert torch.allclose(std[i], x[i, :n].std(0), atol=1e-5)


def test_fp32_encoder_kernels(simt):
    """The fp32-mode encoder kernels (already parity-green on the B200) under the same CPU emulation: a regression net
    for sessions without a GPU, and a check of the emulation itself against kernels whose behaviour is known."""
    g = torch.Generator().manual_seed(1)
    rows, dim = 37, 256
    x, gam, bet = torch.randn(rows, dim, generator=g) * 2 + 1, torch.rand(dim, generator=g) + 0.5, torch.randn(dim, generator=g)
    out = torch.empty(rows, dim)
    outb = torch.empty(rows, dim, dtype=torch.bfloat16)
    assert simt.stac_layernorm(P(x), rows, dim, P(gam), P(bet), 1e-6, P(out), P(outb), None) == 0
    want = torch.nn.functional.layer_norm(x, (dim,), gam, bet, 1e-6)
    assert torch.allclose(out, want, atol=2e-5) and torch.allclose(outb.float(), want, atol=4e-2, rtol=1e-2)
    # attention over a packed qkv projection with key padding by length
    b, t, h = 2, 70, 2
    d = 64 * h
    qkv = torch.randn(b * t, 3 * d, generator=g)
    kv_len = torch.tensor([70, 33], dtype=torch.int32)
    ctx = torch.empty(b * t, d)
    assert simt.stac_mha_f32(P(qkv), P(kv_len), b, t, d, h, P(ctx), None) == 0
    q, k, v = (z.view(b, t, h, 64).transpose(1, 2).double() for z in qkv.split(d, dim=-1))
    mask = torch.arange(t)[None, :] >= kv_len[:, None]
    p = torch.softmax((q @ k.transpose(-1, -2)).masked_fill(mask[:, None, None, :], float("-inf")), -1)
    want = (p @ v).transpose(1, 2).reshape(b * t, d)
    assert (ctx.double() - want).norm() / want.norm() < 1e-5
    # log-softmax with the greedy id
    logits = torch.randn(9, 501, generator=g)
    lp, ids = torch.empty(9, 501), torch.empty(9, dtype=torch.int32)
    assert simt.stac_log_softmax(P(logits), 9, 501, P(lp), P(ids), None) == 0
    assert torch.allclose(lp, torch.log_softmax(logits, -1), atol=1e-5) and torch.equal(ids.long(), logits.argmax(-1))
    # valid-length rules of encode() / forward() in fp32 (TransformerMultiTask.py:289-294 / :225-226)
    wl = torch.tensor([1.0, 0.7304, 0.5, 0.013])
    for rule, fn in ((0, lambda z: torch.floor(z) + 1), (1, torch.round)):
        n = torch.empty(4, dtype=torch.int32)
        assert simt.stac_kv_lengths(P(wl), 4, 37, rule, P(n), None) == 0
        assert torch.equal(n.long(), fn(wl * 37).clamp(1, 37).long())
    # the CUDA-core GEMM with bias, exact-erf GELU and a periodic residual
    m, n_, k_ = 45, 24, 64
    a, w, bias = torch.randn(m, k_, generator=g), torch.randn(n_, k_, generator=g), torch.randn(n_, generator=g)
    res, c = torch.randn(9, n_, generator=g), torch.empty(m, n_)
    assert simt.stac_gemm_f32(P(a), P(w), P(bias), P(res), 9, _lib.ACT_GELU_ERF, P(c), m, n_, k_, None) == 0
    want = torch.nn.functional.gelu(a @ w.T + bias) + res[torch.arange(m) % 9]
    assert torch.allclose(c, want, atol=1e-4)


def test_fp32_path_on_the_emulated_kernels(simt, monkeypatch):
    """The whole fp32 path - Fbank (FFT kernel), normalisation, both convolution blocks, encoder, CTC head - with every
    kernel running from its real source under the CPU emulation, against the oracle at the fp32 tolerance."""
    import abi_emulator
    import oracle
    import stac_speech_translation_b200 as sb
    from stac_speech_translation_b200 import synth
    from util import FP32_TOL, TINY, oracle_modules, product_from_oracle, rel_l2
    emu = abi_emulator.install_simt(monkeypatch, simt)
    omods = oracle_modules(TINY, vocab=64, num_encoder_layers=1)
    mods = product_from_oracle(omods, "fp32", device="cpu")
    wavs, wl = synth.synth_batch([0.33, 0.2], seed=35)
    with torch.no_grad():
        want = oracle.reference_compute_forward(omods, wavs, wl)
    got = sb.compute_forward(mods, wavs, wl)
    for key in ("fbank", "feats", "cnn", "enc_out", "logits", "p_ctc"):
        assert rel_l2(got[key], want[key]) < FP32_TOL, (key, rel_l2(got[key], want[key]))
    assert set(emu.routed) >= {"stac_fbank_logmel", "stac_fbank_topdb_norm", "stac_input_norm", "stac_conv0_ln_lrelu",
                               "stac_conv1_f32", "stac_group_ln_lrelu", "stac_gemm_f32", "stac_layernorm", "stac_mha_f32",
                               "stac_log_softmax"}
    assert set(emu.calls) == set(emu.routed)                  # nothing fell back to the numpy stand-ins


def test_decoder_on_the_emulated_kernels_against_reference_vectors(simt, monkeypatch):
    """TransformerMultiTask.decode() and the KV-cached step with every kernel running from its real source under the CPU
    emulation, against the vectors the REFERENCE'S OWN decode() produced (tests/golden/decoder_reference.npz)."""
    import abi_emulator
    import stac_speech_translation_b200 as sb
    from test_host_emulated import build, fixture
    from util import FP32_TOL, rel_l2
    d, state = fixture()
    emu = abi_emulator.install_simt(monkeypatch, simt)
    tr = build(sb.TransformerMultiTask, state, precision="fp32")
    prefix, enc_out = torch.from_numpy(d["prefix"]), torch.from_numpy(d["enc_out"])
    pred, attn = tr.decode(prefix, enc_out)
    assert rel_l2(pred, torch.from_numpy(d["pred"])) < FP32_TOL and rel_l2(attn, torch.from_numpy(d["attn"])) < FP32_TOL
    pred_len, attn_len = tr.decode(prefix, enc_out, torch.from_numpy(d["enc_len"]))
    assert rel_l2(pred_len, torch.from_numpy(d["pred_len"])) < FP32_TOL
    assert rel_l2(attn_len, torch.from_numpy(d["attn_len"])) < FP32_TOL
    cache = tr.decoder_cache(enc_out, rows=prefix.shape[0], max_len=6)
    for t in range(prefix.shape[1]):
        out, w = cache.step(prefix[:, t].contiguous())
    assert rel_l2(out, torch.from_numpy(d["pred"])[:, -1]) < FP32_TOL
    assert rel_l2(w, torch.from_numpy(d["attn"])[:, -1]) < FP32_TOL
    assert set(emu.calls) == set(emu.routed) >= {"stac_embed_scale_pe", "stac_attention_f32", "stac_gemm_f32"}


def test_turns_ingest_and_train_norm_drop_ins_on_the_emulated_kernels(simt, monkeypatch):
    """The three small drop-ins end to end (Python side + kernels from source under the CPU emulation): RTTM lines equal
    to the reference function's, the PCM decode rule, SpeechBrain's train-mode normalisation updates."""
    import abi_emulator
    import stac_speech_translation_b200 as sb
    from oracle.speechbrain_path import InputNormalization as OracleNorm
    from stac_speech_translation_b200 import ingest, turns
    from util import rel_l2
    abi_emulator.install_simt(monkeypatch, simt)
    for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "turns_reference.json")))[:3]:
        ids = torch.tensor(c["ids"], dtype=torch.int32)
        p = torch.full(ids.shape + (min(c["vocab"], 60),), -20.0)
        p.scatter_(2, ids.long().clamp(max=p.shape[-1] - 1)[..., None], -0.1)
        for x in ((ids, p) if c["vocab"] <= 60 else (ids,)):
            turn, xt = [], []
            turns.append_speaker_turns(c["utt"], x, 7, 8, turn, xt)
            assert turn == c["turn_rttm"] and xt == c["xt_rttm"]
    pcm = torch.randint(-32768, 32768, (3, 1001), dtype=torch.int16)
    assert torch.equal(ingest.pcm_to_float(pcm), pcm.float() / 32768.0)
    g = torch.Generator().manual_seed(4)
    ours, ref = sb.InputNormalization(update_until_epoch=2).train(), OracleNorm(update_until_epoch=2).train()
    for step, epoch in enumerate([0, 0, 1, 2]):
        x = torch.randn(3, 50, 80, generator=g) * (1 + step) + step
        wl = torch.tensor([1.0, 0.73, 0.41])
        assert rel_l2(ours(x, wl, epoch=epoch), ref(x, wl, epoch=epoch)) < 1e-5
        assert rel_l2(ours.glob_mean, ref.glob_mean) < 1e-5 and rel_l2(ours.glob_std, ref.glob_std) < 1e-5


def _spec_augment_case(simt, seed, b, t, f, **kw):
    """One SpecAugment call: the oracle (SpeechBrain's control flow over torch's interpolate / masked_fill_) and the
    kernel fed with the parameters the product class draws from the SAME generator state."""
    from oracle.train_pieces import SpecAugment as OracleAug
    from stac_speech_translation_b200.augment import SpecAugment
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, t, f, generator=g) * 3 + 1
    torch.manual_seed(seed)
    want = OracleAug(**kw)(x.clone())
    torch.manual_seed(seed)
    c, w, (fpos, flen), (tpos, tlen) = SpecAugment(**kw).draw(b, t, f)
    out = torch.full_like(x, float("nan"))
    rc = simt.stac_spec_augment(P(x), b, t, f, c, w, P(fpos), P(flen), 0 if fpos is None else fpos.shape[1], P(tpos),
                                P(tlen), 0 if tpos is None else tpos.shape[1], 0.0, P(out), None)
    assert rc == 0
    return out, want, (c, w)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_spec_augment_kernel(simt, seed):
    """stac_spec_augment from source against the oracle's SpecAugment with the reference's yaml settings
    (transformer_multitask.yaml:283-293), and with each of its three parts alone."""
    ref_cfg = dict(time_warp=True, time_warp_window=5, freq_mask=True, n_freq_mask=2, time_mask=True, n_time_mask=2,
                   freq_mask_width=30, time_mask_width=40)
    out, want, (c, w) = _spec_augment_case(simt, seed, 3, 97 + 13 * seed, 80, **ref_cfg)
    assert c > 0 and w > 0
    assert torch.allclose(out, want, rtol=1e-5, atol=1e-5), (out - want).abs().max()
    assert torch.equal(out == 0, want == 0)                        # the masks are exact
    for part in ("time_warp", "freq_mask", "time_mask"):
        cfg = dict(ref_cfg, time_warp=False, freq_mask=False, time_mask=False)
        cfg[part] = True
        out, want, _ = _spec_augment_case(simt, 10 + seed, 2, 64, 40, **cfg)
        assert torch.allclose(out, want, rtol=1e-5, atol=1e-5), part
    # a sequence too short to warp (time - window <= window) passes through the warp untouched
    out, want, (c, w) = _spec_augment_case(simt, seed, 2, 10, 16, **dict(ref_cfg, freq_mask=False, time_mask=False))
    assert w < 0 and torch.equal(out, want)


@pytest.mark.parametrize("reduction", ["none", "sum", "mean", "batchmean", "batch"])
def test_ctc_loss_kernel(simt, reduction):
    """stac_ctc_loss from source against torch.nn.functional.ctc_loss through SpeechBrain's wrapper (oracle): ragged
    input and target lengths, repeated labels (no skip transition), an empty target, an infeasible utterance (more
    tokens than frames: infinite loss, zeroed), every reduction."""
    from oracle.train_pieces import ctc_loss as oracle_ctc
    g = torch.Generator().manual_seed(21)
    b, t, v, lmax = 6, 40, 12, 9
    lp = torch.randn(b, t, v, generator=g).log_softmax(-1)
    tg = torch.randint(1, v, (b, lmax), generator=g)
    tg[1, :4] = torch.tensor([3, 3, 3, 5])                         # repeats
    in_rel = torch.tensor([1.0, 0.9, 0.55, 0.31, 0.1, 1.0])
    tg_rel = torch.tensor([1.0, 0.45, 0.7, 0.2, 1.0, 0.0])         # row 4: 9 tokens in 4 frames; row 5: empty target
    want = oracle_ctc(lp, tg, in_rel, tg_rel, 0, reduction)
    il = (in_rel * t).round().int()
    tl = (tg_rel * lmax).round().int()
    mode = {"none": 0, "sum": 1, "mean": 2, "batchmean": 3, "batch": 4}[reduction]
    nll = torch.full((b,), float("nan"))
    out = torch.full((b if mode == 4 else 1,), float("nan"))
    tg32 = tg.int().contiguous()
    assert simt.stac_ctc_loss(P(lp), P(tg32), P(il), P(tl), b, t, v, lmax, 0, mode, P(nll), P(out), None) == 0
    per_utt = torch.nn.functional.ctc_loss(lp.transpose(0, 1), tg, il, tl, 0, reduction="none", zero_infinity=True)
    assert torch.allclose(nll, per_utt, rtol=1e-5, atol=1e-5), (nll, per_utt)
    assert float(nll[4]) == 0.0                                    # infeasible -> inf -> zero_infinity
    got = nll if mode == 0 else (out if mode == 4 else out[0])
    if reduction == "batch":                                       # row 5 divides by a zero target length in both
        assert torch.allclose(got[:5], want[:5], rtol=1e-5, atol=1e-5)
    else:
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), (got, want)


@pytest.mark.parametrize("group,lk,h,with_len,with_w", [(1, 37, 2, False, True), (3, 70, 4, True, True), (10, 129, 4, True, False),
                                                         (16, 50, 1, False, True)])
def test_attention_beam_kernel(simt, group, lk, h, with_len, with_w):
    """stac_attention_beam_f32 from source (all hypothesis rows of an utterance in one CTA) against stac_attention_f32 from
    source with lq = 1 - the kernel it replaces in the cached decoding step - and against fp64 softmax attention."""
    g = torch.Generator().manual_seed(group * 100 + lk)
    bm, d = 3, h * 64
    rows = bm * group
    q = torch.randn(rows, d, generator=g) * 0.3
    kv = torch.randn(bm * lk, 2 * d, generator=g)
    kl = torch.randint(1, lk + 1, (rows,), generator=g).int() if with_len else None
    ctx_a, ctx_b = torch.full((rows, d), float("nan")), torch.full((rows, d), float("nan"))
    w_a = torch.full((rows, lk), float("nan")) if with_w else None
    w_b = torch.full((rows, lk), float("nan")) if with_w else None
    k_ptr, v_ptr = c_void_p(kv.data_ptr()), c_void_p(kv.data_ptr() + d * 4)
    assert simt.stac_attention_beam_f32(P(q), d, k_ptr, v_ptr, lk * 2 * d, 2 * d, rows, group, lk, h, P(kl), P(ctx_a), d,
                                        P(w_a), None, None) == 0
    if with_w:                  # the same with the heads on separate CTAs (scratch given): bit-identical weights
        ctx_c, w_c = torch.full((rows, d), float("nan")), torch.full((rows, lk), float("nan"))
        scratch = torch.full((h, rows, lk), float("nan"))
        assert simt.stac_attention_beam_f32(P(q), d, k_ptr, v_ptr, lk * 2 * d, 2 * d, rows, group, lk, h, P(kl), P(ctx_c),
                                            d, P(w_c), P(scratch), None) == 0
        assert torch.equal(ctx_c, ctx_a) and torch.equal(w_c, w_a)
    assert simt.stac_attention_f32(P(q), d, k_ptr, v_ptr, lk * 2 * d, 2 * d, rows, 1, lk, h, group, 0, P(kl), None, 0,
                                   P(ctx_b), d, P(w_b), None) == 0
    assert torch.allclose(ctx_a, ctx_b, rtol=1e-5, atol=1e-6)
    if with_w:
        assert torch.allclose(w_a, w_b, rtol=1e-5, atol=1e-7)
    # fp64 reference
    qq = q.double().view(bm, group, h, 64)
    kk = kv[:, :d].double().view(bm, lk, h, 64)
    vv = kv[:, d:].double().view(bm, lk, h, 64)
    sc = torch.einsum("bghd,bkhd->bghk", qq, kk)
    if kl is not None:
        mask = torch.arange(lk)[None, None, :] >= kl.view(bm, group, 1)
        sc = sc.masked_fill(mask[:, :, None, :], float("-inf"))
    pr = sc.softmax(-1)
    want = torch.einsum("bghk,bkhd->bghd", pr, vv).reshape(rows, d)
    assert torch.allclose(ctx_a.double(), want, rtol=1e-4, atol=1e-5)
    if with_w:
        assert torch.allclose(w_a.double(), pr.mean(2).reshape(rows, lk), rtol=1e-4, atol=1e-6)
