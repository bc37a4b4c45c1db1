"""Decoder side on the device (csrc/decoder_f32.cu + decoder.py) against the vectors the REFERENCE'S OWN decode() /
forward() produced (tests/golden/decoder_reference.npz) and against the oracle at S-model width.  fp32 path: 1e-4.

Host side also covered on the CPU (tests/test_host_emulated.py).  First run on a B200 in round 2
(profiles/r3/r3a_first_call_verification.log); part of the default GPU suite since."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]

import stac_speech_translation_b200 as sb  # noqa: E402
from oracle import speechbrain_path as sp  # noqa: E402
from stac_speech_translation_b200 import decoder as dec, ops  # noqa: E402
from test_host_emulated import build, fixture  # noqa: E402
from util import FP32_TOL, rel_l2  # noqa: E402


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_and_forward_against_reference_vectors(precision):
    d, state = fixture()
    tr = build(sb.TransformerMultiTask, state, precision=precision).cuda()
    prefix = torch.from_numpy(d["prefix"]).cuda()
    enc_out = torch.from_numpy(d["enc_out"]).cuda()
    pred, attn = tr.decode(prefix, enc_out)
    assert pred.dtype == torch.float32 and pred.shape == d["pred"].shape and attn.shape == d["attn"].shape
    assert rel_l2(pred, torch.from_numpy(d["pred"])) < FP32_TOL
    assert rel_l2(attn, torch.from_numpy(d["attn"])) < FP32_TOL
    pred_len, attn_len = tr.decode(prefix, enc_out, torch.from_numpy(d["enc_len"]).cuda())
    assert rel_l2(pred_len, torch.from_numpy(d["pred_len"])) < FP32_TOL
    assert rel_l2(attn_len, torch.from_numpy(d["attn_len"])) < FP32_TOL
    pred1, attn1 = tr.decode(prefix[:, :1].contiguous(), enc_out)
    assert rel_l2(pred1, torch.from_numpy(d["pred1"])) < FP32_TOL and rel_l2(attn1, torch.from_numpy(d["attn1"])) < FP32_TOL
    if precision == "fp32":        # forward() runs the encoder too; its bf16 mode has its own (looser) tolerance
        src, wl = torch.from_numpy(d["src"].astype(np.float32)).cuda(), torch.from_numpy(d["wav_lens"]).cuda()
        enc_f, dec_f = tr(src, torch.from_numpy(d["tgt"]).cuda(), wl, pad_idx=0)
        assert rel_l2(enc_f, torch.from_numpy(d["enc_forward"])) < FP32_TOL
        assert rel_l2(dec_f, torch.from_numpy(d["dec_forward"])) < FP32_TOL


def test_decoder_s_width_against_oracle_with_beam_rows():
    torch.manual_seed(3)
    o = sp.TransformerMultiTask(tgt_vocab=5000, input_size=5120, d_model=256, nhead=4, num_encoder_layers=1,
                                num_decoder_layers=6, d_ffn=1024, activation=torch.nn.GELU, normalize_before=True).eval()
    p = sb.TransformerMultiTask(tgt_vocab=5000, input_size=5120, d_model=256, nhead=4, num_encoder_layers=1,
                                num_decoder_layers=6, d_ffn=1024, activation=torch.nn.GELU, normalize_before=True,
                                precision="bf16").eval()
    p.load_state_dict(o.state_dict(), strict=True)
    p = p.cuda()
    g = torch.Generator().manual_seed(11)
    b, beam, t2, L = 3, 4, 251, 9
    enc = torch.randn(b, t2, 256, generator=g)
    tok = torch.randint(1, 5000, (b * beam, L), generator=g)
    with torch.no_grad():
        want, want_w = o.decode(tok, enc.repeat_interleave(beam, 0))          # the searcher's inflated memory
    got, got_w = p.decode(tok.cuda(), enc.repeat_interleave(beam, 0).cuda())
    assert rel_l2(got, want) < FP32_TOL and rel_l2(got_w, want_w) < FP32_TOL
    got2, got_w2 = dec.decoder_stack(tok.cuda(), enc.cuda(), p.packed_decoder())  # shared memory, row r -> r // beam
    assert rel_l2(got2, want) < FP32_TOL and rel_l2(got_w2, want_w) < FP32_TOL
    assert (got_w.sum(-1) - 1).abs().max() < 1e-4


@pytest.mark.parametrize("rows,lq,lk,h,div,causal", [(1, 1, 1, 1, 1, 0), (5, 7, 7, 2, 1, 1), (6, 3, 300, 4, 2, 0),
                                                     (4, 2, 2500, 1, 1, 0), (3, 33, 33, 8, 1, 1)])
def test_attention_f32_kernel(rows, lq, lk, h, div, causal):
    g = torch.Generator().manual_seed(rows * 100 + lk)
    d = 64 * h
    n_mem = rows // div
    q = torch.randn(rows * lq, d, generator=g)
    kv = torch.randn(n_mem * lk, 2 * d, generator=g)
    kv_len = torch.randint(1, lk + 1, (rows,), generator=g, dtype=torch.int32)
    tok = torch.randint(0, 3, (rows, lk), generator=g)
    tok[:, 0] = 1                                            # never a fully masked row
    use_tok = causal == 1
    qd, kvd, kld, tokd = q.cuda(), kv.cuda(), kv_len.cuda(), tok.cuda()
    ctx = torch.full((rows * lq, d), float("nan"), device="cuda")
    w = torch.full((rows, lq, lk), float("nan"), device="cuda")
    ops._call("stac_attention_f32", ops.ptr(qd), d, ops.ptr(kvd), dec._off(kvd, d), lk * 2 * d, 2 * d, rows, lq, lk, h, div,
              causal,
              ops.ptr(kld), ops.ptr(tokd) if use_tok else ops.ptr(None), 0, ops.ptr(ctx), d, ops.ptr(w), ops.stream())
    torch.cuda.synchronize()
    qq = q.view(rows, lq, h, 64).permute(0, 2, 1, 3).double()
    k = kv[:, :d].reshape(n_mem, lk, h, 64).permute(0, 2, 1, 3).double().repeat_interleave(div, 0)
    v = kv[:, d:].reshape(n_mem, lk, h, 64).permute(0, 2, 1, 3).double().repeat_interleave(div, 0)
    mask = torch.arange(lk)[None, None, :] >= kv_len[:, None, None]
    if causal:
        mask = mask | (torch.arange(lk)[None, None, :] > torch.arange(lq)[None, :, None])
    if use_tok:
        mask = mask | (tok == 0)[:, None, :]
    s = (qq @ k.transpose(-1, -2)).masked_fill(mask[:, None], float("-inf"))
    p = torch.softmax(s, -1)
    want = (p @ v).permute(0, 2, 1, 3).reshape(rows * lq, d)
    assert rel_l2(ctx, want) < 1e-5
    assert rel_l2(w, p.mean(1)) < 1e-5


def test_embed_scale_pe_kernel():
    g = torch.Generator().manual_seed(1)
    vocab, d, L, rows = 97, 128, 5, 4
    emb, pe = torch.randn(vocab, d, generator=g), torch.randn(40, d, generator=g)
    tok = torch.randint(0, vocab, (rows, L), generator=g)
    out = torch.empty(rows * L, d, device="cuda")
    embd, ped, tokd = emb.cuda(), pe.cuda(), tok.cuda()
    ops._call("stac_embed_scale_pe", ops.ptr(tokd, torch.int64), ops.ptr(embd), ops.ptr(ped), rows * L, L, d, vocab,
              float(np.sqrt(d)), ops.ptr(out), ops.stream())
    want = emb[tok] * np.float32(np.sqrt(d)) + pe[:L][None]
    assert rel_l2(out.view(rows, L, d), want) < 1e-6


def test_kv_cached_steps_equal_full_prefix_decode_on_device():
    d, state = fixture()
    tr = build(sb.TransformerMultiTask, state, precision="bf16").cuda()
    prefix, enc_out = torch.from_numpy(d["prefix"]).cuda(), torch.from_numpy(d["enc_out"]).cuda()
    beam = 3
    rows = prefix.repeat_interleave(beam, 0)
    cache = tr.decoder_cache(enc_out, rows=rows.shape[0], max_len=rows.shape[1])
    for t in range(rows.shape[1]):
        out, w = cache.step(rows[:, t].contiguous())
        if t == 1:
            cache.reorder(torch.arange(rows.shape[0], device="cuda"))      # identity re-ordering: nothing may change
    assert rel_l2(out[::beam], torch.from_numpy(d["pred"])[:, -1]) < FP32_TOL
    assert rel_l2(w[::beam], torch.from_numpy(d["attn"])[:, -1]) < FP32_TOL
    full, full_w = tr.decode(rows, enc_out.repeat_interleave(beam, 0))
    assert rel_l2(out, full[:, -1]) < 1e-5 and rel_l2(w, full_w[:, -1]) < 1e-5


def test_kv_cached_steps_bf16_gemms_on_device():
    """DecoderCache(precision="bf16") on the device: tensor-core GEMMs for every projection of the step, against the
    fp32 cache step by step (bf16 tolerance) and against the reference-generated vectors; with beam rows over the
    un-inflated memory and a re-ordering."""
    from util import BF16_TOL
    d, state = fixture()
    tr = build(sb.TransformerMultiTask, state, precision="bf16").cuda()
    prefix, enc_out = torch.from_numpy(d["prefix"]).cuda(), torch.from_numpy(d["enc_out"]).cuda()
    beam = 3
    rows = prefix.repeat_interleave(beam, 0)
    c32 = tr.decoder_cache(enc_out, rows=rows.shape[0], max_len=rows.shape[1])
    c16 = tr.decoder_cache(enc_out, rows=rows.shape[0], max_len=rows.shape[1], precision="bf16")
    index = torch.arange(rows.shape[0], device="cuda").flip(0)
    for t in range(rows.shape[1]):
        if t == 2:
            c32.reorder(index)
            c16.reorder(index)
        o32, w32 = c32.step(rows[:, t].contiguous())
        o16, w16 = c16.step(rows[:, t].contiguous())
        assert rel_l2(o16, o32) < BF16_TOL and rel_l2(w16, w32) < BF16_TOL, t
    c16b = tr.decoder_cache(enc_out, rows=rows.shape[0], max_len=rows.shape[1], precision="bf16")
    for t in range(rows.shape[1]):
        out, w = c16b.step(rows[:, t].contiguous())
    assert rel_l2(out[::beam], torch.from_numpy(d["pred"])[:, -1]) < BF16_TOL
    assert rel_l2(w[::beam], torch.from_numpy(d["attn"])[:, -1]) < BF16_TOL


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_kv_cached_steps_as_one_graph_replay(precision):
    """DecoderCache(graph=True): from the second step on a step is one CUDA-graph replay (the position counter lives on
    the device, the self-attention kernel appends to the cache itself).  Bit-identical to the eager steps, including a
    beam re-ordering after the capture and a rewind."""
    d, state = fixture()
    tr = build(sb.TransformerMultiTask, state, precision="bf16").cuda()
    prefix, enc_out = torch.from_numpy(d["prefix"]).cuda(), torch.from_numpy(d["enc_out"]).cuda()
    rows = prefix.repeat_interleave(3, 0)
    eager = tr.decoder_cache(enc_out, rows=rows.shape[0], max_len=rows.shape[1] + 2, precision=precision)
    graph = tr.decoder_cache(enc_out, rows=rows.shape[0], max_len=rows.shape[1] + 2, precision=precision, graph=True)
    index = torch.arange(rows.shape[0], device="cuda").roll(1)
    for t in range(rows.shape[1]):
        if t == 3:
            eager.reorder(index)
            graph.reorder(index)
        oe, we = eager.step(rows[:, t].contiguous())
        og, wg = graph.step(rows[:, t].contiguous())
        assert torch.equal(oe, og) and torch.equal(we, wg), t
    assert graph._graph is not None
    eager.rewind(2)
    graph.rewind(2)
    oe, we = eager.step(rows[:, 2].contiguous())
    og, wg = graph.step(rows[:, 2].contiguous())
    assert torch.equal(oe, og) and torch.equal(we, wg)


@pytest.mark.parametrize("graph", [False, True])
def test_cached_forward_step_drives_the_same_search_on_the_device(graph):
    """searcher.CachedStepMixin on the device (row-map re-ordering, optionally one graph replay per step) against the
    reference's forward_step / permute_mem bodies (mutitask_decoder.py:101-128: whole-prefix decode, index_select of
    the token memory) under the same beam-search control flow: same hypotheses and back-pointers at every step."""
    from stac_speech_translation_b200.searcher import CachedStepMixin, _update_mem
    from test_host_emulated import _StubBeamSearcher
    d, state = fixture()
    tr = build(sb.TransformerMultiTask, state, precision="bf16").cuda()
    torch.manual_seed(5)
    fc = torch.nn.Linear(tr.d_model, tr.tgt_vocab).cuda()

    class DeviceSearch(_StubBeamSearcher):
        def search(self, enc_states, steps):
            dev = enc_states.device
            b, beam = enc_states.shape[0], self.beam_size
            enc = enc_states.repeat_interleave(beam, 0)
            memory = self.reset_mem(b * beam, dev)
            inp = torch.full((b * beam,), self.bos_index, dtype=torch.long, device=dev)
            scores = torch.zeros(b, beam, device=dev)
            scores[:, 1:] = -1e9
            trace = []
            for _ in range(steps):
                logp, memory, attn = self.forward_step(inp, memory, enc, None)
                vocab = logp.shape[-1]
                cand = (scores.view(b * beam, 1) + logp).view(b, beam * vocab)
                scores, idx = cand.topk(beam, dim=-1)
                pred = (idx // vocab + torch.arange(b, device=dev)[:, None] * beam).view(-1)
                inp = (idx % vocab).view(-1)
                memory = self.permute_mem(memory, pred)
                trace.append((scores.clone(), inp.clone(), pred.clone(), attn[:, -1].clone()))
            return memory, trace

    class ReferenceStep(DeviceSearch):
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

    class CachedStep(CachedStepMixin, DeviceSearch):
        decoder_graph = graph

    enc_out = torch.from_numpy(d["enc_out"]).cuda()
    args = dict(model=tr, fc=fc, beam_size=3, bos_index=1, prefix=[1, 9, 12])
    with torch.no_grad():
        mem_ref, trace_ref = ReferenceStep(**args).search(enc_out, steps=7)
        mem_new, trace_new = CachedStep(**args).search(enc_out, steps=7)
    assert torch.equal(mem_ref, mem_new)
    for (s0, i0, p0, a0), (s1, i1, p1, a1) in zip(trace_ref, trace_new):
        assert torch.equal(i0, i1) and torch.equal(p0, p1)
        assert rel_l2(s1, s0) < 1e-5 and rel_l2(a1, a0) < 1e-4
