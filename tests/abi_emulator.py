"""TEST INFRASTRUCTURE ONLY: a numpy stand-in for a few libstac_b200 entry points, so that the HOST-side orchestration
(argument order, pointer offsets into packed projections, leading dimensions, mask plumbing, weight packing) can be
exercised by the CPU test-suite, which has no GPU.  It is not a fallback: it lives under tests/, is patched in by the
``emulated_abi`` fixture only, and the product raises on CPU tensors without it.

Every emulated function takes exactly the arguments of its C prototype (include/stac_b200.h) - raw addresses and
sizes - and is checked against the ctypes signature table before it runs."""
import ctypes
from ctypes import c_void_p

import numpy as np
from scipy.special import erf

from stac_speech_translation_b200 import _lib


def _arr(addr, n, dtype=np.float32):
    if not addr:
        return None
    nbytes = int(n) * np.dtype(dtype).itemsize
    return np.frombuffer((ctypes.c_char * nbytes).from_address(addr), dtype=dtype)


class Emulator:
    def __init__(self):
        self.calls = []

    def call(self, name, *args):
        res, sig = _lib._SIGNATURES[name]
        assert len(args) == len(sig), (name, len(args), len(sig))
        vals = []
        for a, t in zip(args, sig):
            if t is c_void_p:
                assert a is None or isinstance(a, c_void_p), (name, a)
                vals.append((a.value or 0) if a is not None else 0)
            elif t is ctypes.c_float:
                vals.append(float(a))
            else:
                assert isinstance(a, int) and not isinstance(a, bool), (name, a)
                vals.append(int(a))
        self.calls.append(name)
        getattr(self, name)(*vals)

    # ---- entry points (arguments exactly as in include/stac_b200.h) ----
    def stac_embed_scale_pe(self, tokens, emb, pe, rows, seq_len, d_model, vocab, scale, out, stream):
        tok = _arr(tokens, rows, np.int64)
        e = _arr(emb, vocab * d_model).reshape(vocab, d_model)
        p = _arr(pe, seq_len * d_model).reshape(seq_len, d_model)
        o = _arr(out, rows * d_model).reshape(rows, d_model)
        o[:] = e[tok] * np.float32(scale) + p[np.arange(rows) % seq_len]

    def stac_layernorm(self, x, rows, dim, gamma, beta, eps, out_f32, out_bf16, stream):
        assert out_bf16 == 0, "emulator: fp32 output only"
        xx = _arr(x, rows * dim).reshape(rows, dim).astype(np.float64)
        mu = xx.mean(1, keepdims=True)
        var = xx.var(1, keepdims=True)
        y = (xx - mu) / np.sqrt(var + eps) * _arr(gamma, dim) + _arr(beta, dim)
        _arr(out_f32, rows * dim).reshape(rows, dim)[:] = y.astype(np.float32)

    def stac_gemm_f32(self, a, w, bias, resid, resid_period, act, c, m, n, k, stream):
        y = _arr(a, m * k).reshape(m, k).astype(np.float64) @ _arr(w, n * k).reshape(n, k).astype(np.float64).T
        if bias:
            y = y + _arr(bias, n)
        if act == _lib.ACT_GELU_ERF:
            y = 0.5 * y * (1.0 + erf(y / np.sqrt(2.0)))
        if resid:
            if resid_period:
                y = y + _arr(resid, resid_period * n).reshape(resid_period, n)[np.arange(m) % resid_period]
            else:
                y = y + _arr(resid, m * n).reshape(m, n)
        _arr(c, m * n).reshape(m, n)[:] = y.astype(np.float32)

    def stac_mha_f32(self, qkv, kv_len, batch, seq_len, d_model, n_head, ctx, stream):
        x = _arr(qkv, batch * seq_len * 3 * d_model).reshape(batch, seq_len, 3, n_head, 64).astype(np.float64)
        n = _arr(kv_len, batch, np.int32)
        out = _arr(ctx, batch * seq_len * d_model).reshape(batch, seq_len, n_head, 64)
        for b in range(batch):
            nk = min(max(int(n[b]), 1), seq_len)
            for h in range(n_head):
                s = x[b, :, 0, h] @ x[b, :nk, 1, h].T
                p = np.exp(s - s.max(1, keepdims=True))
                out[b, :, h] = (p / p.sum(1, keepdims=True)) @ x[b, :nk, 2, h]

    def stac_attention_f32(self, q, ldq, k, v, kv_bs, kv_rs, rows, lq, lk, n_head, mem_rows_div, causal, kv_len,
                           key_tokens, pad_idx, ctx, ldctx, weights, stream):
        n_mem = (rows + mem_rows_div - 1) // mem_rows_div
        d = n_head * 64
        qq = _arr(q, (rows * lq - 1) * ldq + d)
        kk = _arr(k, (n_mem - 1) * kv_bs + (lk - 1) * kv_rs + d)
        vv = _arr(v, (n_mem - 1) * kv_bs + (lk - 1) * kv_rs + d)
        cc = _arr(ctx, (rows * lq - 1) * ldctx + d)
        kl = _arr(kv_len, rows, np.int32)
        kt = _arr(key_tokens, rows * lk, np.int64)
        ww = _arr(weights, rows * lq * lk)
        for r in range(rows):
            rb = r // mem_rows_div
            for i in range(lq):
                row = r * lq + i
                nk = lk if kl is None else min(max(int(kl[r]), 0), lk)
                if causal:
                    nk = min(nk, i + 1)
                wacc = np.zeros(lk)
                for h in range(n_head):
                    qv = qq[row * ldq + h * 64: row * ldq + h * 64 + 64].astype(np.float64)
                    s = np.full(lk, -np.inf)
                    for j in range(nk):
                        if kt is not None and kt[r * lk + j] == pad_idx:
                            continue
                        o = rb * kv_bs + j * kv_rs + h * 64
                        s[j] = qv @ kk[o:o + 64]
                    p = np.exp(s - s.max())
                    p /= p.sum()
                    acc = np.zeros(64)
                    for j in range(nk):
                        o = rb * kv_bs + j * kv_rs + h * 64
                        acc += p[j] * vv[o:o + 64]
                    cc[row * ldctx + h * 64: row * ldctx + h * 64 + 64] = acc.astype(np.float32)
                    wacc += p / n_head
                if ww is not None:
                    ww[row * lk:(row + 1) * lk] = wacc.astype(np.float32)


    def stac_utt_mean_std(self, x, wav_len, batch, frames, dim, eps, mean, std, stream):
        xx = _arr(x, batch * frames * dim).reshape(batch, frames, dim).astype(np.float64)
        wl = _arr(wav_len, batch)
        for b in range(batch):
            n = int(np.rint(np.float32(wl[b]) * np.float32(frames)))
            seg = xx[b, :n]
            _arr(mean, batch * dim).reshape(batch, dim)[b] = seg.mean(0)
            _arr(std, batch * dim).reshape(batch, dim)[b] = np.maximum(seg.std(0, ddof=1), eps)

    def stac_input_norm(self, x, mean, std, rows, dim, out, stream):
        _arr(out, rows * dim).reshape(rows, dim)[:] = (_arr(x, rows * dim).reshape(rows, dim) - _arr(mean, dim)) / _arr(std, dim)

    def stac_argmax_rows(self, x, rows, cols, out, stream):
        _arr(out, rows, np.int32)[:] = _arr(x, rows * cols).reshape(rows, cols).argmax(1)

    def stac_ctc_spikes(self, ids, batch, t2, turn_id, xt_id, row_counts, spikes_turn, spikes_xt, n_out, stream):
        flat = _arr(ids, batch * t2, np.int32)
        for k, (tok, dst) in enumerate(((turn_id, spikes_turn), (xt_id, spikes_xt))):
            pos = np.nonzero(flat == tok)[0]
            _arr(dst, batch * t2, np.int32)[:len(pos)] = pos
            _arr(n_out, 2, np.int32)[k] = len(pos)
            _arr(row_counts, batch * 2, np.int32).reshape(batch, 2)[:, k] = (flat.reshape(batch, t2) == tok).sum(1)

    def stac_pcm_i16_to_f32(self, pcm, n, out, stream):
        _arr(out, n)[:] = _arr(pcm, n, np.int16).astype(np.float32) * np.float32(1 / 32768)


class EmuLib:
    """Stands in for the ctypes handle (`_lib.lib()`): every entry point returns STAC_OK after the emulator ran."""

    def __init__(self, emu):
        self._emu = emu

    def __getattr__(self, name):
        def fn(*args):
            self._emu.call(name, *args)
            return 0
        return fn


def cpu_ptr(t, dtype=None):
    """`_lib.ptr` without the is_cuda requirement (same contiguity / dtype checks)."""
    if t is None:
        return c_void_p(0)
    assert t.is_contiguous()
    if dtype is not None and t.dtype != dtype:
        raise _lib.StacB200Error(f"expected {dtype}, got {t.dtype}")
    return c_void_p(t.data_ptr())


def install(monkeypatch):
    """Route the host orchestration's kernel calls to the emulator (CPU tensors)."""
    from stac_speech_translation_b200 import decoder, ingest, ops, turns
    emu = Emulator()
    for mod in (ops, decoder, turns, ingest):
        monkeypatch.setattr(mod, "ptr", cpu_ptr)
        monkeypatch.setattr(mod, "stream", lambda: c_void_p(0))
    for mod in (turns, ingest):
        monkeypatch.setattr(mod, "lib", lambda: EmuLib(emu))
    monkeypatch.setattr(ops, "_call", emu.call)
    return emu
