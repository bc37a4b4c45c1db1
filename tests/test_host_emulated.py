"""Host-side orchestration of the decoder (decoder.py, transformer.decode / forward) on the CPU, against the vectors the
REFERENCE'S OWN decode() / forward() produced (tests/golden/decoder_reference.npz, make_decoder_golden.py).

The kernels cannot run here; tests/abi_emulator.py stands in for the five entry points the fp32 path uses, taking the
raw addresses and sizes the host code passes across the C ABI.  What this pins without a GPU: weight packing (q rows
pre-scaled, cross-attention in_proj split), argument order, pointer offsets into the packed projections, leading
dimensions, which masks reach which attention, the returned attention weights.  The kernels themselves are compared with
the same vectors in tests/test_gpu_decoder.py."""
import os

import numpy as np
import pytest
import torch

import abi_emulator
import stac_speech_translation_b200 as sb
from oracle import speechbrain_path as sp
from util import rel_l2

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "decoder_reference.npz")


def fixture():
    d = np.load(GOLDEN)
    state = {k[len("state/"):]: torch.from_numpy(d[k].astype(np.float32)) for k in d.files if k.startswith("state/")}
    return d, state


def build(cls, state, **extra):
    d_model = state["encoder.norm.norm.weight"].shape[0]
    n_dec = 1 + max(int(k.split(".")[2]) for k in state if k.startswith("decoder.layers."))
    tr = cls(tgt_vocab=state["custom_tgt_module.layers.0.emb.Embedding.weight"].shape[0],
             input_size=state["custom_src_module.layers.0.w.weight"].shape[1], d_model=d_model, nhead=d_model // 64,
             num_encoder_layers=1, num_decoder_layers=n_dec,
             d_ffn=state["encoder.layers.0.pos_ffn.ffn.0.weight"].shape[0], dropout=0.1, activation=torch.nn.GELU,
             normalize_before=True, causal=False, **extra)
    res = tr.load_state_dict(state, strict=False)
    assert res.missing_keys == ["positional_encoding.pe"] and not res.unexpected_keys     # every decoder key present
    return tr.eval()


def test_oracle_decoder_against_reference_generated_fixture():
    d, state = fixture()
    tr = build(sp.TransformerMultiTask, state)
    src, wl = torch.from_numpy(d["src"].astype(np.float32)), torch.from_numpy(d["wav_lens"])
    tgt, prefix = torch.from_numpy(d["tgt"]), torch.from_numpy(d["prefix"])
    enc_out = torch.from_numpy(d["enc_out"])
    torch.set_num_threads(1)
    with torch.no_grad():
        enc_f, dec_f = tr(src, tgt, wl, pad_idx=0)
        pred, attn = tr.decode(prefix, enc_out)
        pred_len, attn_len = tr.decode(prefix, enc_out, torch.from_numpy(d["enc_len"]))
        pred1, attn1 = tr.decode(prefix[:, :1], enc_out)
    for got, key in [(enc_f, "enc_forward"), (dec_f, "dec_forward"), (pred, "pred"), (attn, "attn"),
                     (pred_len, "pred_len"), (attn_len, "attn_len"), (pred1, "pred1"), (attn1, "attn1")]:
        assert rel_l2(got, torch.from_numpy(d[key])) < 1e-6, key
    # the fixture separates the cases: memory padding changes decode(), and forward() pads both sides
    assert rel_l2(torch.from_numpy(d["pred_len"]), torch.from_numpy(d["pred"])) > 1e-4
    assert np.abs(d["attn_len"][1, :, 15:]).max() == 0 and np.abs(d["attn"][1, :, 15:]).max() > 0


def test_host_decoder_orchestration_against_reference_vectors(monkeypatch):
    d, state = fixture()
    emu = abi_emulator.install(monkeypatch)
    tr = build(sb.TransformerMultiTask, state, precision="fp32")
    prefix = torch.from_numpy(d["prefix"])
    enc_out = torch.from_numpy(d["enc_out"])
    pred, attn = tr.decode(prefix, enc_out)                                    # mutitask_decoder.py:126
    assert pred.dtype == torch.float32 and pred.shape == d["pred"].shape and attn.shape == d["attn"].shape
    assert rel_l2(pred, torch.from_numpy(d["pred"])) < 1e-5
    assert rel_l2(attn, torch.from_numpy(d["attn"])) < 1e-5
    assert emu.calls.count("stac_attention_f32") == 2 * len(tr.decoder.layers)
    pred_len, attn_len = tr.decode(prefix, enc_out, torch.from_numpy(d["enc_len"]))
    assert rel_l2(pred_len, torch.from_numpy(d["pred_len"])) < 1e-5
    assert rel_l2(attn_len, torch.from_numpy(d["attn_len"])) < 1e-5
    pred1, attn1 = tr.decode(prefix[:, :1], enc_out)
    assert rel_l2(pred1, torch.from_numpy(d["pred1"])) < 1e-5 and rel_l2(attn1, torch.from_numpy(d["attn1"])) < 1e-5
    # forward(): encoder (round-rule lengths) + decoder with look-ahead, tgt == pad and memory padding
    src, wl = torch.from_numpy(d["src"].astype(np.float32)), torch.from_numpy(d["wav_lens"])
    enc_f, dec_f = tr(src, torch.from_numpy(d["tgt"]), wl, pad_idx=0)
    assert rel_l2(enc_f, torch.from_numpy(d["enc_forward"])) < 1e-5
    assert rel_l2(dec_f, torch.from_numpy(d["dec_forward"])) < 1e-5
    # beam-inflated rows over a shared (not inflated) memory: row r reads memory[r // beam]
    from stac_speech_translation_b200 import decoder as dec
    beam = 2
    out, w = dec.decoder_stack(prefix.repeat_interleave(beam, 0), enc_out, tr.packed_decoder())
    assert rel_l2(out[::beam], torch.from_numpy(d["pred"])) < 1e-5 and torch.equal(out[::beam], out[1::beam])
    assert rel_l2(w[1::beam], torch.from_numpy(d["attn"])) < 1e-5


def test_decoder_weights_follow_checkpoint_updates(monkeypatch):
    d, state = fixture()
    abi_emulator.install(monkeypatch)
    tr = build(sb.TransformerMultiTask, state, precision="fp32")
    prefix, enc_out = torch.from_numpy(d["prefix"]), torch.from_numpy(d["enc_out"])
    a, _ = tr.decode(prefix, enc_out)
    sd = tr.state_dict()
    sd["decoder.layers.1.pos_ffn.ffn.3.weight"] = sd["decoder.layers.1.pos_ffn.ffn.3.weight"] * 2.0
    tr.load_state_dict(sd)                                   # e.g. checkpoint averaging, inference.py:228-233
    b, _ = tr.decode(prefix, enc_out)
    assert rel_l2(b, a) > 1e-3


def test_kv_cached_steps_equal_the_full_prefix_decode(monkeypatch):
    """DecoderCache.step, token by token, against the reference-generated decode() vectors (last position of every
    prefix) and against the full-prefix device path; then a beam re-ordering."""
    d, state = fixture()
    abi_emulator.install(monkeypatch)
    tr = build(sb.TransformerMultiTask, state, precision="fp32")
    prefix, enc_out = torch.from_numpy(d["prefix"]), torch.from_numpy(d["enc_out"])
    b, L = prefix.shape
    cache = tr.decoder_cache(enc_out, rows=b, max_len=8)
    for t in range(L):
        out, w = cache.step(prefix[:, t])
        full, full_w = tr.decode(prefix[:, :t + 1].contiguous(), enc_out)
        assert rel_l2(out, full[:, -1]) < 1e-5 and rel_l2(w, full_w[:, -1]) < 1e-5
        if t == 0:
            assert rel_l2(out, torch.from_numpy(d["pred1"])[:, 0]) < 1e-5
    assert rel_l2(out, torch.from_numpy(d["pred"])[:, -1]) < 1e-5
    assert rel_l2(w, torch.from_numpy(d["attn"])[:, -1]) < 1e-5
    # with encoder lengths, beam rows over the un-inflated memory, and a re-ordering in the middle
    beam = 2
    enc_len = torch.from_numpy(d["enc_len"])
    rows_prefix = prefix.repeat_interleave(beam, 0).clone()
    rows_prefix[1::beam, 1:] = (rows_prefix[1::beam, 1:] + 7) % 60 + 1            # the second beam diverges after step 0
    cache = tr.decoder_cache(enc_out, rows=b * beam, max_len=L, enc_len=enc_len.repeat_interleave(beam, 0))
    index = torch.tensor([1, 1, 2, 3, 5, 4])                                       # row 0 <- row 1, rows 4 / 5 swapped
    cur = rows_prefix.clone()
    for t in range(L):
        if t == 2:
            cache.reorder(index)
            cur = cur[index]
            cur[:, t:] = rows_prefix[:, t:]
        out, w = cache.step(cur[:, t].contiguous())
    want, want_w = tr.decode(cur, enc_out.repeat_interleave(beam, 0), enc_len.repeat_interleave(beam, 0))
    assert rel_l2(out, want[:, -1]) < 1e-5 and rel_l2(w, want_w[:, -1]) < 1e-5
    with pytest.raises(sb.StacB200Error, match="full"):
        cache.step(cur[:, 0].contiguous())


def test_kv_cached_steps_in_bf16_mode(monkeypatch):
    """DecoderCache(precision="bf16"): the step's projections and feed-forward block through stac_gemm_bf16 (bf16 operand
    copies of the weights, bf16 LayerNorm outputs, fp32 caches and residual stream): every step within the bf16
    tolerance of the fp32 cache and of the reference-generated vectors, same call structure (7 GEMMs per layer and step)."""
    from util import BF16_TOL
    d, state = fixture()
    emu = abi_emulator.install(monkeypatch)
    tr = build(sb.TransformerMultiTask, state, precision="fp32")
    prefix, enc_out = torch.from_numpy(d["prefix"]), torch.from_numpy(d["enc_out"])
    b, L = prefix.shape
    c32 = tr.decoder_cache(enc_out, rows=b, max_len=L)
    emu.calls.clear()
    c16 = tr.decoder_cache(enc_out, rows=b, max_len=L, precision="bf16")
    n_layers = len(tr.packed_decoder().layers)
    assert emu.calls.count("stac_gemm_bf16") == n_layers                       # cross keys / values, once per layer
    for t in range(L):
        o32, w32 = c32.step(prefix[:, t])
        emu.calls.clear()
        o16, w16 = c16.step(prefix[:, t])
        assert emu.calls.count("stac_gemm_bf16") == 6 * n_layers and "stac_gemm_f32" not in emu.calls   # fused qkv, out, q2, out2, ffn1, ffn2
        assert rel_l2(o16, o32) < BF16_TOL and rel_l2(w16, w32) < BF16_TOL, t
    assert rel_l2(o16, torch.from_numpy(d["pred"])[:, -1]) < BF16_TOL
    with pytest.raises(sb.StacB200Error):
        tr.decoder_cache(enc_out, rows=b, max_len=L, precision="fp16")


def test_turns_and_ingest_host_code_through_the_emulated_abi(monkeypatch):
    """The Python side of turns.py / ingest.py (argument order, slicing of the compacted spikes, shapes) with the
    emulator behind the C ABI: same golden lines as the reference function, same decode rule."""
    import json
    from stac_speech_translation_b200 import ingest, turns
    abi_emulator.install(monkeypatch)
    for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "turns_reference.json"))):
        ids = torch.tensor(c["ids"], dtype=torch.int32)
        p = torch.full(ids.shape + (c["vocab"],), -20.0)
        p.scatter_(2, ids.long()[..., None], -0.1)
        for x in (ids, p):
            turn, xt = [], []
            turns.append_speaker_turns(c["utt"], x, 7, 8, turn, xt)
            assert turn == c["turn_rttm"] and xt == c["xt_rttm"]
    pcm = torch.randint(-32768, 32768, (3, 1001), dtype=torch.int16)
    assert torch.equal(ingest.pcm_to_float(pcm), pcm.float() / 32768.0)


class _StubBeamSearcher:
    """A minimal beam search with the calling convention of SpeechBrain's S2SBeamSearcher as the reference's subclass
    sees it (reset_mem -> per step: forward_step, top-k over beam x vocab, permute_mem): enough to drive both
    forward_step implementations through identical control flow.  NOT a restatement of SpeechBrain's searcher."""

    def __init__(self, model, fc, beam_size, bos_index, prefix, temperature=1.15):
        self.model, self.fc, self.beam_size, self.bos_index = model, fc, beam_size, bos_index
        self.decoder_input_tokens, self.temperature = prefix, temperature
        self.softmax = torch.nn.LogSoftmax(dim=-1)

    def search(self, enc_states, steps):
        b = enc_states.shape[0]
        beam = self.beam_size
        enc = enc_states.repeat_interleave(beam, 0)
        memory = self.reset_mem(b * beam, enc.device)
        inp = torch.full((b * beam,), self.bos_index, dtype=torch.long)
        scores = torch.zeros(b, beam)
        scores[:, 1:] = -1e9
        trace = []
        for _ in range(steps):
            logp, memory, attn = self.forward_step(inp, memory, enc, None)
            vocab = logp.shape[-1]
            cand = (scores.view(b * beam, 1) + logp).view(b, beam * vocab)
            scores, idx = cand.topk(beam, dim=-1)
            pred = (idx // vocab + torch.arange(b)[:, None] * beam).view(-1)
            inp = (idx % vocab).view(-1)
            memory = self.permute_mem(memory, pred)
            trace.append((scores.clone(), inp.clone(), pred.clone(), attn[:, -1].clone()))
        return memory, trace


def test_cached_forward_step_drives_the_same_search_as_the_reference_forward_step(monkeypatch):
    """searcher.CachedStepMixin against the reference's forward_step / permute_mem / reset_mem bodies
    (mutitask_decoder.py:101-128, restated in the baseline class below) under the same beam-search control flow: same
    hypotheses, scores and back-pointers at every step."""
    from stac_speech_translation_b200.searcher import CachedStepMixin, _update_mem
    d, state = fixture()
    abi_emulator.install(monkeypatch)
    tr = build(sb.TransformerMultiTask, state, precision="fp32")
    torch.manual_seed(5)
    fc = torch.nn.Linear(tr.d_model, tr.tgt_vocab)

    class ReferenceStep(_StubBeamSearcher):           # mutitask_decoder.py:101-128, verbatim semantics
        def reset_mem(self, batch_size, device):
            return torch.tensor([self.decoder_input_tokens] * batch_size).to(device)

        def permute_mem(self, memory, index):
            return torch.index_select(memory, dim=0, index=index)

        def forward_step(self, inp_tokens, memory, enc_states, enc_lens):
            if not torch.all(inp_tokens == self.bos_index):
                memory = _update_mem(inp_tokens, memory)
            pred, attn = self.model.decode(memory, enc_states)
            prob_dist = self.softmax(self.fc(pred) / self.temperature)
            return prob_dist[:, -1, :], memory, attn

    class CachedStep(CachedStepMixin, _StubBeamSearcher):
        pass

    enc_out = torch.from_numpy(d["enc_out"])
    args = dict(model=tr, fc=fc, beam_size=3, bos_index=1, prefix=[1, 9, 12])
    with torch.no_grad():
        mem_ref, trace_ref = ReferenceStep(**args).search(enc_out, steps=5)
        mem_new, trace_new = CachedStep(**args).search(enc_out, steps=5)
    assert torch.equal(mem_ref, mem_new)
    for (s0, i0, p0, a0), (s1, i1, p1, a1) in zip(trace_ref, trace_new):
        assert torch.equal(i0, i1) and torch.equal(p0, p1)
        assert rel_l2(s1, s0) < 1e-5 and rel_l2(a1, a0) < 1e-4


def test_train_mode_input_normalization_follows_speechbrain_updates(monkeypatch):
    """features.InputNormalization in training mode (device statistics through the emulated ABI + SpeechBrain's running
    update on the host) against the oracle's restatement over several steps, epochs and the eval call that follows."""
    from oracle.speechbrain_path import InputNormalization as OracleNorm
    abi_emulator.install(monkeypatch)
    g = torch.Generator().manual_seed(4)
    ours = sb.InputNormalization(norm_type="global", update_until_epoch=2).train()
    ref = OracleNorm(norm_type="global", update_until_epoch=2).train()
    for step, epoch in enumerate([0, 0, 1, 2, 3]):
        x = torch.randn(3, 50, 80, generator=g) * (1 + step) + step
        wl = torch.tensor([1.0, 0.73, 0.41])
        got, want = ours(x, wl, epoch=epoch), ref(x, wl, epoch=epoch)
        assert rel_l2(got, want) < 1e-5, (step, epoch)
        assert ours.count == ref.count and rel_l2(ours.glob_mean, ref.glob_mean) < 1e-6
        assert rel_l2(ours.glob_std, ref.glob_std) < 1e-6
    ours.eval(), ref.eval()
    x = torch.randn(2, 30, 80, generator=g)
    assert rel_l2(ours(x, torch.ones(2)), ref(x, torch.ones(2))) < 1e-5
    assert ours.count == ref.count == 5


def test_fp32_path_host_orchestration_against_the_oracle(monkeypatch):
    """The whole six-call sequence (pipeline.compute_forward: Fbank, InputNormalization, ConvolutionFrontEnd, encode,
    ctc_lin, log_softmax) in fp32 mode on the CPU, with the emulated C ABI standing in for the kernels, against the
    oracle at every stage boundary: weight packing, table layouts, shapes, strides, mask rules and call order of the host
    side are exercised without a GPU (the kernels themselves: tests/test_gpu_fp32_path.py)."""
    import oracle
    from stac_speech_translation_b200 import synth
    from util import TINY, oracle_modules, product_from_oracle
    abi_emulator.install(monkeypatch)
    omods = oracle_modules(TINY, vocab=64)
    mods = product_from_oracle(omods, "fp32", device="cpu")
    wavs, wl = synth.synth_batch([0.9, 0.62], seed=31)
    for train_mask in (False, True):
        with torch.no_grad():
            want = oracle.reference_compute_forward(omods, wavs, wl, train_mask=train_mask)
        got = sb.compute_forward(mods, wavs, wl, train_mask=train_mask)
        for key in ("fbank", "feats", "cnn", "enc_out", "logits", "p_ctc"):
            assert got[key].shape == want[key].shape, key
            assert rel_l2(got[key], want[key]) < 1e-4, (key, train_mask, rel_l2(got[key], want[key]))
        # the fused call (what bench.py and the multi-GPU driver use): same results, plus the greedy ids
        res = sb.EncoderPipeline(mods, train_mask=train_mask)(wavs, wl)
        assert rel_l2(res["enc_out"], want["enc_out"]) < 1e-4 and rel_l2(res["p_ctc"], want["p_ctc"]) < 1e-4
        assert torch.equal(res["greedy"].long(), want["p_ctc"].argmax(-1))


@pytest.mark.parametrize("mha_v2,tiny", [(False, False), (True, False), (False, True)])
def test_bf16_path_host_orchestration_against_the_oracle(monkeypatch, mha_v2, tiny):
    """The benchmark path's host side (fused EncoderPipeline in bf16 mode: tensor-core Fbank tables, padded parity-split
    conv0 layout, packed conv1 / projection weights with the pre-scaled q rows, fused FFN and CTC head calls, greedy
    ids) on the CPU through the emulated C ABI, against the fp32 oracle at the bf16 tolerance; S width so that the fused
    feed-forward call is taken (tiny: d_model 128, the two-GEMM feed-forward that the M / L sizes use).
    mha_v2: the STAC_MHA_V2=1 branch of ops.encoder_stack."""
    import oracle
    from stac_speech_translation_b200 import ops, synth
    from util import BF16_TOL, oracle_modules, product_from_oracle
    emu = abi_emulator.install(monkeypatch)
    monkeypatch.setattr(ops, "MHA_V2", mha_v2)
    from util import TINY
    omods = oracle_modules(TINY, vocab=64) if tiny else oracle_modules("S", num_encoder_layers=2, vocab=64)
    mods = product_from_oracle(omods, "bf16", device="cpu")
    wavs, wl = synth.synth_batch([0.7, 0.45], seed=32)
    with torch.no_grad():
        want = oracle.reference_compute_forward(omods, wavs, wl)
    res = sb.EncoderPipeline(mods)(wavs, wl)
    assert res["enc_out"].dtype == torch.float32 and res["p_ctc"].dtype == torch.float32
    assert rel_l2(res["enc_out"], want["enc_out"]) < BF16_TOL and rel_l2(res["p_ctc"], want["p_ctc"]) < BF16_TOL
    assert (res["greedy"].long() == res["p_ctc"].argmax(-1)).all()
    assert ("stac_mha_bf16_v2" if mha_v2 else "stac_mha_bf16") in emu.calls
    assert {"stac_fbank_logmel_tc2", "stac_conv1_bf16", "stac_ctc_head_bf16"} <= set(emu.calls)
    assert ("stac_ffn_fused_bf16" in emu.calls) == (not tiny)
    n_layers = len(mods["Transformer"].packed().layers)
    assert emu.calls.count("stac_layernorm") == 2 * n_layers + 1
    if not tiny:
        # the out-proj + residual + LayerNorm 2 kernel (off by default: measured not faster) gives the same result
        monkeypatch.setattr(ops, "FUSED_OUTPROJ_LN", True)
        emu.calls.clear()
        res2 = sb.EncoderPipeline(mods)(wavs, wl)
        assert emu.calls.count("stac_outproj_ln_bf16") == n_layers and emu.calls.count("stac_layernorm") == n_layers + 1
        assert rel_l2(res2["enc_out"], res["enc_out"]) < 5e-3
        monkeypatch.setattr(ops, "FUSED_OUTPROJ_LN", False)
    # the six-call sequence in bf16 mode (fp32 tensors at every stage boundary, as the reference's callers expect)
    got = sb.compute_forward(mods, wavs, wl)
    assert all(got[k].dtype == torch.float32 for k in ("fbank", "feats", "cnn", "enc_out", "logits", "p_ctc"))
    assert rel_l2(got["cnn"], want["cnn"]) < BF16_TOL and rel_l2(got["enc_out"], want["enc_out"]) < BF16_TOL
    assert rel_l2(got["p_ctc"], want["p_ctc"]) < BF16_TOL


def test_structured_custom_ops_equal_the_pipeline(monkeypatch):
    """torch.ops.stac_b200.{fbank-free front-end, encoder, ctc_head} with flat weight lists give what the fused
    EncoderPipeline gives (bf16 mode, emulated ABI), i.e. the custom-op layer is a faithful view of the same calls."""
    import stac_speech_translation_b200.custom_ops as co
    from stac_speech_translation_b200 import ops, synth
    from util import oracle_modules, product_from_oracle
    abi_emulator.install(monkeypatch)
    omods = oracle_modules("S", num_encoder_layers=1, vocab=64)
    mods = product_from_oracle(omods, "bf16", device="cpu")
    wavs, wl = synth.synth_batch([0.5, 0.33], seed=33)
    pipe = sb.EncoderPipeline(mods)
    res = pipe(wavs, wl)
    ns = torch.ops.stac_b200
    fb, norm = mods["compute_features"], mods["normalize"]
    mean, std = norm.device_stats(wavs.device, 80)
    feats = ops.fbank_tc(wavs, fb.tc_tables(wavs.device), fb.top_db, fb.top_db_per_utterance, mean, std)
    src = ns.conv_frontend(feats, co.frontend_weight_list(mods["CNN"].packed()), "bf16", True)
    tr = mods["Transformer"]
    enc = ns.encoder(src, res["kv_len"], co.encoder_weight_list(tr.packed()), tr.d_model, tr.nhead, "bf16")
    assert torch.equal(enc, res["enc_out"])
    ctc = mods["ctc_lin"]
    p, ids = ns.ctc_head(enc.to(torch.bfloat16), ctc.packed_weight(), ctc.w.bias.detach().float().contiguous())
    assert rel_l2(p, res["p_ctc"]) < 2e-2 and p.shape == res["p_ctc"].shape and ids.shape == res["greedy"].shape


@pytest.mark.parametrize("n", [160 * 3, 16000 + 32, 160 * 129 + 96])
def test_fbank_tc2_tables_against_the_oracle(monkeypatch, n):
    """Tables of the second tensor-core Fbank kernel (fold factors wa / wb, the 7 twiddle tiles [208 x (cos 32 | sin 32)])
    through an emulation that does the kernel's arithmetic with exactly those tables: log-mel equal to the oracle's Fbank
    and to the first kernel's formulation at the fp16-operand tolerance (1e-3, as the GPU test asserts)."""
    from stac_speech_translation_b200 import ops, synth
    from util import oracle_modules
    abi_emulator.install(monkeypatch)
    omods = oracle_modules("S", num_encoder_layers=1, vocab=64)
    wavs, _ = synth.synth_batch([n / 16000.0, max(0.03, 0.61 * n / 16000.0)], seed=n)
    wavs = wavs[:, :n].contiguous()
    ref = omods["compute_features"](wavs)
    tabs = ops.build_fbank_tc_tables("cpu")
    assert tabs.tab2.numel() == 416 and tuple(tabs.tw2.shape) == (7 * 208, 64)
    got2 = ops.fbank_tc(wavs, tabs)
    monkeypatch.setenv("STAC_FBANK_V2", "0")
    got1 = ops.fbank_tc(wavs, tabs)
    assert got2.shape == ref.shape
    assert rel_l2(got2, ref) < 1e-3 and rel_l2(got1, ref) < 1e-3 and rel_l2(got2, got1) < 1e-3


def test_training_stage_pieces_host_side(monkeypatch):
    """SpecAugment and ctc_loss drop-ins on the CPU through the emulated C ABI: the parameter draws follow
    SpeechBrain's order (same seed -> same augmentation as the oracle), lengths are rounded the way SpeechBrain rounds
    them, every reduction is passed through, unsupported settings raise."""
    from oracle.train_pieces import SpecAugment as OracleAug, ctc_loss as oracle_ctc
    emu = abi_emulator.install(monkeypatch)
    cfg = dict(time_warp=True, time_warp_window=5, time_warp_mode="bicubic", freq_mask=True, n_freq_mask=2, time_mask=True,
               n_time_mask=2, freq_mask_width=30, time_mask_width=40)                 # transformer_multitask.yaml:283-293
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 131, 80, generator=g)
    ours, ref = sb.SpecAugment(**cfg), OracleAug(**cfg)
    for step in range(3):
        torch.manual_seed(100 + step)
        want = ref(x.clone())
        torch.manual_seed(100 + step)
        got = ours(x)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5) and torch.equal(got == 0, want == 0)
    assert emu.calls.count("stac_spec_augment") == 3
    with pytest.raises(sb.StacB200Error):
        sb.SpecAugment(time_warp_mode="nearest")
    with pytest.raises(sb.StacB200Error):
        sb.SpecAugment(replace_with_zero=False)
    lp = torch.randn(5, 60, 30, generator=g).log_softmax(-1)
    tg = torch.randint(1, 30, (5, 11), generator=g)
    in_rel, tg_rel = torch.tensor([1.0, 0.83, 0.5, 0.25, 0.61]), torch.tensor([1.0, 0.5, 0.37, 0.1, 0.9])
    for reduction in ("mean", "sum", "batchmean", "batch", "none"):
        got = sb.ctc_loss(lp, tg, in_rel, tg_rel, 0, reduction)
        want = oracle_ctc(lp, tg, in_rel, tg_rel, 0, reduction)
        assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-5, atol=1e-5), reduction
    with pytest.raises(sb.StacB200Error):
        sb.ctc_loss(lp, tg, in_rel, tg_rel, 0, "median")
