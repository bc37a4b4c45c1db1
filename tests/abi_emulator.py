"""TEST INFRASTRUCTURE ONLY: a numpy stand-in for a few libstac_b200 entry points, so that the HOST-side orchestration
(argument order, pointer offsets into packed projections, leading dimensions, mask plumbing, weight packing) can be
exercised by the CPU test-suite, which has no GPU.  It is not a fallback: it lives under tests/, is patched in by the
``emulated_abi`` fixture only, and the product raises on CPU tensors without it.

Every emulated function takes exactly the arguments of its C prototype (include/stac_b200.h) - raw addresses and
sizes - and is checked against the ctypes signature table before it runs."""
import ctypes
from ctypes import c_void_p

import numpy as np
import torch
from scipy.special import erf

from stac_speech_translation_b200 import _lib


def _arr(addr, n, dtype=np.float32):
    if not addr:
        return None
    nbytes = int(n) * np.dtype(dtype).itemsize
    return np.frombuffer((ctypes.c_char * nbytes).from_address(addr), dtype=dtype)


def _tarr(addr, n, dtype):
    """torch view of raw memory (for bf16 / fp16, which numpy does not have)."""
    if not addr:
        return None
    nbytes = int(n) * torch.empty(0, dtype=dtype).element_size()
    return torch.frombuffer((ctypes.c_char * nbytes).from_address(addr), dtype=dtype)


def _gelu(y):
    return 0.5 * y * (1.0 + erf(y / np.sqrt(2.0)))


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
        xx = _arr(x, rows * dim).reshape(rows, dim).astype(np.float64)
        mu = xx.mean(1, keepdims=True)
        var = xx.var(1, keepdims=True)
        y = (xx - mu) / np.sqrt(var + eps) * _arr(gamma, dim) + _arr(beta, dim)
        if out_f32:
            _arr(out_f32, rows * dim).reshape(rows, dim)[:] = y.astype(np.float32)
        if out_bf16:
            _tarr(out_bf16, rows * dim, torch.bfloat16)[:] = torch.from_numpy(y).to(torch.bfloat16).flatten()

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

    def stac_embed_step(self, tokens, emb, pe, rows, d_model, vocab, scale, pos_dev, out, stream):
        pos = int(_arr(pos_dev, 1, np.int32)[0])
        self.stac_embed_scale_pe(tokens, emb, pe + pos * d_model * 4, rows, 1, d_model, vocab, scale, out, stream)

    def stac_attention_step_f32(self, q, ldq, k, v, kv_rs, kv_ts, rows, lk, n_head, row_map, k_new, v_new, ld_new, t_dev,
                                ctx, ldctx, stream):
        """Self-attention of one decoding step over a time-major cache, key j of row r in cache row row_map[j, r]; in
        append mode the position counter is read from the device and the new keys / values are stored into slab t."""
        d = n_head * 64
        if t_dev:
            t = min(int(_arr(t_dev, 1, np.int32)[0]), lk - 1)
            lk = t + 1
        qq = _arr(q, (rows - 1) * ldq + d)
        rm = _arr(row_map, lk * rows, np.int32)
        span = (rows - 1) * kv_rs + (lk - 1) * kv_ts + d
        kk, vv = _arr(k, span), _arr(v, span)
        if t_dev:
            kn, vn = _arr(k_new, (rows - 1) * ld_new + d), _arr(v_new, (rows - 1) * ld_new + d)
            for r in range(rows):
                o = r * kv_rs + t * kv_ts
                kk[o:o + d] = kn[r * ld_new:r * ld_new + d]
                vv[o:o + d] = vn[r * ld_new:r * ld_new + d]
        cc = _arr(ctx, (rows - 1) * ldctx + d)
        for r in range(rows):
            for h in range(n_head):
                qv = qq[r * ldq + h * 64: r * ldq + h * 64 + 64].astype(np.float64)
                offs = [(int(rm[j * rows + r]) if rm is not None else r) * kv_rs + j * kv_ts + h * 64 for j in range(lk)]
                s = np.array([qv @ kk[o:o + 64] for o in offs])
                pr = np.exp(s - s.max())
                pr /= pr.sum()
                acc = np.zeros(64)
                for j, o in enumerate(offs):
                    acc += pr[j] * vv[o:o + 64]
                cc[r * ldctx + h * 64: r * ldctx + h * 64 + 64] = acc.astype(np.float32)

    def stac_attention_beam_f32(self, q, ldq, k, v, kv_bs, kv_rs, rows, group, lk, n_head, kv_len, ctx, ldctx, weights,
                                head_scratch, stream):
        """One query per row, `group` rows per memory block: by its documentation the general attention with lq = 1."""
        assert 0 < group <= 16 and rows % group == 0
        self.stac_attention_f32(q, ldq, k, v, kv_bs, kv_rs, rows, 1, lk, n_head, group, 0, kv_len, 0, 0, ctx, ldctx,
                                weights, stream)

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


    # ---- bf16 mode (the benchmark path): same arithmetic with bf16 roundings at the documented hand-offs ----
    def stac_cast_bf16(self, x, n, out, stream):
        _tarr(out, n, torch.bfloat16)[:] = torch.from_numpy(_arr(x, n).copy()).to(torch.bfloat16)

    def stac_outproj_ln_bf16(self, a, w, bias, x, ln_g, ln_b, eps, h, m, stream):
        y = _tarr(a, m * 256, torch.bfloat16).double().view(m, 256).numpy() @ \
            _tarr(w, 256 * 256, torch.bfloat16).double().view(256, 256).numpy().T
        if bias:
            y = y + _arr(bias, 256)
        xs = _arr(x, m * 256).reshape(m, 256)
        xn = (xs.astype(np.float64) + y).astype(np.float32)
        xs[:] = xn
        xd = xn.astype(np.float64)
        mu, var = xd.mean(-1, keepdims=True), xd.var(-1, keepdims=True)
        out = (xd - mu) / np.sqrt(var + eps) * _arr(ln_g, 256) + _arr(ln_b, 256)
        _tarr(h, m * 256, torch.bfloat16).view(m, 256)[:] = torch.from_numpy(out).to(torch.bfloat16)

    def stac_cast_f32(self, x, n, out, stream):
        _arr(out, n)[:] = _tarr(x, n, torch.bfloat16).float().numpy()

    def stac_gemm_bf16(self, a, w, bias, resid, resid_period, act, c, c_dtype, m, n, k, vt_out, vt_cols, seq_len, t_pad,
                       stream):
        assert vt_out == 0, "emulator: no transposed V output"
        y = _tarr(a, m * k, torch.bfloat16).double().view(m, k).numpy() @ \
            _tarr(w, n * k, torch.bfloat16).double().view(n, k).numpy().T
        if bias:
            y = y + _arr(bias, n)
        if act == _lib.ACT_GELU_ERF:
            y = _gelu(y)
        if resid:
            if resid_period:
                y = y + _arr(resid, resid_period * n).reshape(resid_period, n)[np.arange(m) % resid_period]
            else:
                y = y + _arr(resid, m * n).reshape(m, n)
        if c_dtype == _lib.DT_BF16:
            _tarr(c, m * n, torch.bfloat16)[:] = torch.from_numpy(y).to(torch.bfloat16).flatten()
        else:
            _arr(c, m * n).reshape(m, n)[:] = y.astype(np.float32)

    def stac_ffn_fused_bf16(self, h, w1, b1, w2, b2, x, m, d_model, d_ffn, stream):
        hh = _tarr(h, m * d_model, torch.bfloat16).double().view(m, d_model).numpy()
        hid = _gelu(hh @ _tarr(w1, d_ffn * d_model, torch.bfloat16).double().view(d_ffn, d_model).numpy().T
                    + _arr(b1, d_ffn))
        hid = torch.from_numpy(hid).to(torch.bfloat16).double().numpy()          # P is bf16 in shared memory
        y = hid @ _tarr(w2, d_model * d_ffn, torch.bfloat16).double().view(d_model, d_ffn).numpy().T + _arr(b2, d_model)
        xx = _arr(x, m * d_model).reshape(m, d_model)
        xx[:] = (xx + y).astype(np.float32)

    def _mha_bf16(self, qkv, kv_len, batch, seq_len, d_model, n_head, ctx):
        x = _tarr(qkv, batch * seq_len * 3 * d_model, torch.bfloat16).double().view(batch, seq_len, 3, n_head, 64).numpy()
        n = _arr(kv_len, batch, np.int32)
        out = np.zeros((batch, seq_len, n_head, 64))
        for b in range(batch):
            nk = min(max(int(n[b]), 1), seq_len)
            for h in range(n_head):
                s = x[b, :, 0, h] @ x[b, :nk, 1, h].T
                p = np.exp(s - s.max(1, keepdims=True))
                out[b, :, h] = (p / p.sum(1, keepdims=True)) @ x[b, :nk, 2, h]
        _tarr(ctx, batch * seq_len * d_model, torch.bfloat16)[:] = torch.from_numpy(out).to(torch.bfloat16).flatten()

    def stac_mha_bf16(self, qkv, v_t, kv_len, batch, seq_len, t_pad, d_model, n_head, ctx, stream):
        assert v_t == 0 and t_pad >= seq_len
        self._mha_bf16(qkv, kv_len, batch, seq_len, d_model, n_head, ctx)

    def stac_mha_bf16_v2(self, qkv, kv_len, batch, seq_len, d_model, n_head, ctx, stream):
        self._mha_bf16(qkv, kv_len, batch, seq_len, d_model, n_head, ctx)

    def stac_ctc_head_bf16(self, enc, w, bias, m, vocab, d_model, workspace, log_probs, out_dtype, argmax, stream):
        assert workspace
        x = _tarr(enc, m * d_model, torch.bfloat16).double().view(m, d_model).numpy() @ \
            _tarr(w, vocab * d_model, torch.bfloat16).double().view(vocab, d_model).numpy().T
        if bias:
            x = x + _arr(bias, vocab)
        mx = x.max(1, keepdims=True)
        y = x - (mx + np.log(np.exp(x - mx).sum(1, keepdims=True)))
        if out_dtype == _lib.DT_BF16:
            _tarr(log_probs, m * vocab, torch.bfloat16)[:] = torch.from_numpy(y).to(torch.bfloat16).flatten()
        else:
            _arr(log_probs, m * vocab).reshape(m, vocab)[:] = y.astype(np.float32)
        if argmax:
            _arr(argmax, m, np.int32)[:] = x.argmax(1)

    def stac_fbank_logmel_tc(self, pcm, batch, n_samples, row_stride, tables, twiddles, logmel_db, utt_max_ordered,
                             stream):
        from stac_speech_translation_b200 import ops
        tab = _arr(tables, 400 + 208 * 2)
        window, wbin = tab[:400].astype(np.float64), tab[400:].reshape(208, 2).astype(np.float64)
        assert _tarr(twiddles, 2 * 208 * 256, torch.float16).abs().max() <= 1.0
        fb = ops.mel_filter_matrix().numpy()                      # only its sparsity structure is used (compiled into the
        frames = 1 + n_samples // 160                             # kernel); the weights come from the table
        x = _arr(pcm, (batch - 1) * row_stride + n_samples)
        db = _arr(logmel_db, batch * frames * 80).reshape(batch, frames, 80)
        umax = _arr(utt_max_ordered, batch, np.uint32)
        for b in range(batch):
            sig = np.concatenate([np.zeros(200), x[b * row_stride: b * row_stride + n_samples].astype(np.float64),
                                  np.zeros(200)])
            idx = np.arange(frames)[:, None] * 160 + np.arange(400)[None, :]
            spec = np.fft.rfft(sig[idx] * window, axis=1)
            power = spec.real ** 2 + spec.imag ** 2
            mel = np.zeros((frames, 80))
            for k in range(201):
                for j, m in enumerate(np.nonzero(fb[:, k] > 0)[0]):
                    mel[:, m] += power[:, k] * wbin[k, j]
            db[b] = (10.0 * np.log10(np.maximum(mel, 1e-10))).astype(np.float32)
            u = np.float32(db[b].max()).view(np.uint32)
            umax[b] = (~u & 0xffffffff) if (u & 0x80000000) else (u | 0x80000000)

    def stac_fbank_logmel_tc2(self, pcm, batch, n_samples, row_stride, tables, twiddles, logmel_db, utt_max_ordered,
                              pair, stream):
        """The second tensor-core kernel, arithmetic as the kernel does it: the frame folded on its symmetry, rounded to
        fp16, multiplied stage by stage with the windowed twiddle tiles the host built (so their layout is what is under
        test), fp32-like accumulation, power, mel with the per-bin weights."""
        from stac_speech_translation_b200 import ops
        assert pair in (0, 1)
        wbin = _arr(tables, 208 * 2).reshape(208, 2).astype(np.float64)
        tw = _tarr(twiddles, 7 * 208 * 64, torch.float16).double().view(7, 208, 64).numpy()
        fb = ops.mel_filter_matrix().numpy()
        frames = 1 + n_samples // 160
        x = _arr(pcm, (batch - 1) * row_stride + n_samples)
        db = _arr(logmel_db, batch * frames * 80).reshape(batch, frames, 80)
        umax = _arr(utt_max_ordered, batch, np.uint32)
        n = np.arange(224)
        for b in range(batch):
            sig = np.concatenate([np.zeros(200, np.float32), x[b * row_stride: b * row_stride + n_samples],
                                  np.zeros(1024, np.float32)])
            start = np.arange(frames)[:, None] * 160
            a = sig[start + n[None, :]]
            m = sig[start + (400 - n)[None, :]]
            lone = (n == 0) | (n == 200)                        # no mirror partner
            e = np.where(lone[None, :], a, a + m).astype(np.float16).astype(np.float64)
            o = (a - m).astype(np.float16).astype(np.float64)
            re = np.zeros((frames, 208))
            im = np.zeros((frames, 208))
            for i in range(7):
                re += e[:, 32 * i: 32 * i + 32] @ tw[i, :, :32].T
                im += o[:, 32 * i: 32 * i + 32] @ tw[i, :, 32:].T
            power = re ** 2 + im ** 2
            mel = np.zeros((frames, 80))
            for k in range(201):
                for j, mm in enumerate(np.nonzero(fb[:, k] > 0)[0]):
                    mel[:, mm] += power[:, k] * wbin[k, j]
            db[b] = (10.0 * np.log10(np.maximum(mel, 1e-10))).astype(np.float32)
            u = np.float32(db[b].max()).view(np.uint32)
            umax[b] = (~u & 0xffffffff) if (u & 0x80000000) else (u | 0x80000000)

    def stac_conv1_bf16(self, xpad, w1, b1, ln_g, ln_b, batch, t1, out, stream):
        t2, tp2 = (t1 - 1) // 2 + 1, (t1 + 3) // 2
        planes = _tarr(xpad, batch * 4 * tp2 * 21 * 256, torch.bfloat16).double().view(batch, 4, tp2, 21, 256).numpy()
        pad = np.zeros((batch, 2 * tp2, 42, 256))                 # padded coordinates tp = t1 + 1, fp = f1 + 1
        for par in range(4):
            pad[:, (par >> 1)::2, (par & 1)::2] = planes[:, par]
        w = _tarr(w1, 9 * 256 * 256, torch.bfloat16).double().view(3, 3, 256, 256).numpy()      # [kf][kt][out][in]
        y = np.zeros((batch, t2, 20, 256))
        for kf in range(3):
            for kt in range(3):
                patch = pad[:, kt: kt + 2 * t2: 2, kf: kf + 40: 2]                                   # [B, T2, 20, in]
                y += patch @ w[kf, kt].T
        y = (y + _arr(b1, 256)).reshape(batch, t2, 20 * 256)
        y = self._ln_lrelu(y, _arr(ln_g, 20 * 256), _arr(ln_b, 20 * 256), 1e-5, 0.01)
        _tarr(out, batch * t2 * 5120, torch.bfloat16)[:] = torch.from_numpy(y).to(torch.bfloat16).flatten()

    # ---- a2 - a4, a8: the fp32 front-end and CTC head, from the C-ABI documentation (include/stac_b200.h) ----
    def stac_fbank_logmel(self, pcm, batch, n_samples, row_stride, tables, logmel_db, utt_max_ordered, stream):
        tab = _arr(tables, 2692)
        window = tab[:400].astype(np.float64)
        start, count = tab[1252:1332].astype(int), tab[1332:1412].astype(int)
        weights = tab[1412:].reshape(80, 16).astype(np.float64)
        frames = 1 + n_samples // 160
        x = _arr(pcm, (batch - 1) * row_stride + n_samples)
        db = _arr(logmel_db, batch * frames * 80).reshape(batch, frames, 80)
        umax = _arr(utt_max_ordered, batch, np.uint32)
        for b in range(batch):
            sig = np.concatenate([np.zeros(200), x[b * row_stride: b * row_stride + n_samples].astype(np.float64),
                                  np.zeros(200)])                                   # centre = True, zero padding
            idx = np.arange(frames)[:, None] * 160 + np.arange(400)[None, :]
            spec = np.fft.rfft(sig[idx] * window, axis=1)
            power = spec.real ** 2 + spec.imag ** 2                                   # [T, 201]
            mel = np.stack([(power[:, start[m]: start[m] + count[m]] * weights[m, :count[m]]).sum(1)
                            for m in range(80)], 1)
            db[b] = (10.0 * np.log10(np.maximum(mel, 1e-10))).astype(np.float32)
            u = np.float32(db[b].max()).view(np.uint32)
            umax[b] = (~u & 0xffffffff) if (u & 0x80000000) else (u | 0x80000000)      # order-preserving key

    def stac_fbank_topdb_norm(self, logmel_db, utt_max_ordered, per_utterance, top_db, mean, std, batch, frames, n_mels,
                              out, stream):
        x = _arr(logmel_db, batch * frames * n_mels).reshape(batch, frames, n_mels).copy()
        keys = _arr(utt_max_ordered, batch, np.uint32)
        mx = np.array([np.uint32((k & 0x7fffffff) if (k & 0x80000000) else (~k & 0xffffffff)).view(np.float32)
                       for k in keys])
        if not per_utterance:
            mx[:] = mx.max()
        y = np.maximum(x, (mx - np.float32(top_db))[:, None, None])
        if mean:
            y = (y - _arr(mean, n_mels)) / _arr(std, n_mels)
        _arr(out, batch * frames * n_mels).reshape(batch, frames, n_mels)[:] = y

    @staticmethod
    def _conv_block(x_btfc, weight_oifk, bias, stride=2):
        """[B, T, F, C_in] -> [B, T', F', C_out]: reflect pad 1 in time and frequency, 3 x 3 conv, kernel [kf][kt]."""
        import torch
        import torch.nn.functional as F
        x = torch.from_numpy(np.ascontiguousarray(x_btfc)).double().permute(0, 3, 2, 1)      # [B, C, F, T]
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        y = F.conv2d(x, torch.from_numpy(weight_oifk).double(), torch.from_numpy(bias).double(), stride=stride)
        return y.permute(0, 3, 2, 1).numpy()                                                  # [B, T', F', C_out]

    @staticmethod
    def _ln_lrelu(x, gamma, beta, eps, slope):
        mu = x.mean(-1, keepdims=True)
        var = x.var(-1, keepdims=True)
        y = (x - mu) / np.sqrt(var + eps) * gamma + beta
        return np.where(y >= 0, y, y * slope)

    def stac_conv0_ln_lrelu(self, feats, w0, b0, ln_g, ln_b, batch, frames, out, out_mode, stream):
        t1 = (frames - 1) // 2 + 1
        x = _arr(feats, batch * frames * 80).reshape(batch, frames, 80, 1)
        w = _arr(w0, 256 * 9).reshape(256, 1, 3, 3).astype(np.float64)
        y = self._conv_block(x, w, _arr(b0, 256).astype(np.float64)).reshape(batch, t1, 40 * 256)
        y = self._ln_lrelu(y, _arr(ln_g, 40 * 256), _arr(ln_b, 40 * 256), 1e-5, 0.01)
        if out_mode == _lib.DT_F32:
            _arr(out, batch * t1 * 40 * 256).reshape(batch, t1, 40 * 256)[:] = y
            return
        # bf16: reflect-padded, parity-split planes [B][4 = (tp & 1) * 2 + (fp & 1)][Tp2][21][256], tp = t1 + 1, fp = f1 + 1
        tp2 = (t1 + 3) // 2
        y = y.reshape(batch, t1, 40, 256)
        tidx = np.abs(np.arange(-1, t1 + 1))
        tidx = np.where(tidx >= t1, 2 * (t1 - 1) - tidx, tidx)
        fidx = np.abs(np.arange(-1, 41))
        fidx = np.where(fidx >= 40, 2 * 39 - fidx, fidx)
        pad = np.zeros((batch, 2 * tp2, 42, 256))
        pad[:, : t1 + 2] = y[:, tidx][:, :, fidx]
        planes = _tarr(out, batch * 4 * tp2 * 21 * 256, torch.bfloat16).view(batch, 4, tp2, 21, 256)
        for par in range(4):
            planes[:, par] = torch.from_numpy(np.ascontiguousarray(pad[:, (par >> 1)::2, (par & 1)::2])).to(torch.bfloat16)

    def stac_conv0_topdb_norm_bf16(self, logmel_db, utt_max_ordered, per_utterance, top_db, mean, std, w0, b0, ln_g, ln_b,
                                   batch, frames, out, stream):
        tmp = np.zeros(batch * frames * 80, np.float32)
        self.stac_fbank_topdb_norm(logmel_db, utt_max_ordered, per_utterance, top_db, mean, std, batch, frames, 80,
                                   tmp.ctypes.data, stream)
        self.stac_conv0_ln_lrelu(tmp.ctypes.data, w0, b0, ln_g, ln_b, batch, frames, out, _lib.DT_BF16, stream)

    def stac_conv1_f32(self, x, w1, b1, batch, t1, out, stream):
        t2 = (t1 - 1) // 2 + 1
        xx = _arr(x, batch * t1 * 40 * 256).reshape(batch, t1, 40, 256)
        w = _arr(w1, 256 * 9 * 256).reshape(256, 3, 3, 256).transpose(0, 3, 1, 2).astype(np.float64)   # [out][in][kf][kt]
        y = self._conv_block(xx, np.ascontiguousarray(w), _arr(b1, 256).astype(np.float64))
        _arr(out, batch * t2 * 20 * 256).reshape(batch, t2, 20, 256)[:] = y

    def stac_group_ln_lrelu(self, x, rows, dim, gamma, beta, eps, slope, out, out_dtype, stream):
        assert out_dtype == _lib.DT_F32, "emulator: fp32 output only"
        y = self._ln_lrelu(_arr(x, rows * dim).reshape(rows, dim).astype(np.float64), _arr(gamma, dim), _arr(beta, dim),
                           eps, slope)
        _arr(out, rows * dim).reshape(rows, dim)[:] = y

    def stac_log_softmax(self, logits, rows, vocab, out, argmax, stream):
        x = _arr(logits, rows * vocab).reshape(rows, vocab).astype(np.float64)
        m = x.max(1, keepdims=True)
        _arr(out, rows * vocab).reshape(rows, vocab)[:] = x - (m + np.log(np.exp(x - m).sum(1, keepdims=True)))
        if argmax:
            _arr(argmax, rows, np.int32)[:] = x.argmax(1)

    def stac_utt_mean_std(self, x, wav_len, batch, frames, dim, eps, mean, std, stream):
        xx = _arr(x, batch * frames * dim).reshape(batch, frames, dim).astype(np.float64)
        wl = _arr(wav_len, batch)
        for b in range(batch):
            n = int(np.rint(np.float32(wl[b]) * np.float32(frames)))
            seg = xx[b, :n]
            _arr(mean, batch * dim).reshape(batch, dim)[b] = seg.mean(0)
            _arr(std, batch * dim).reshape(batch, dim)[b] = np.maximum(seg.std(0, ddof=1), eps)

    def stac_spec_augment(self, x, batch, frames, dim, warp_center, warp_width, freq_pos, freq_len, n_freq, time_pos,
                          time_len, n_time, fill, out, stream):
        """From the C-ABI documentation: rows [0, w) = rows [0, c) resampled, rows [w, T) = rows [c, T) resampled (bicubic,
        align_corners), then the [pos, pos + len) masks - with torch's own interpolate as the resampler."""
        import torch.nn.functional as F
        xx = _tarr(x, batch * frames * dim, torch.float32).view(batch, 1, frames, dim).clone()
        if warp_width >= 0:
            c, w = warp_center, warp_width
            xx = torch.cat([F.interpolate(xx[:, :, :c], (w, dim), mode="bicubic", align_corners=True),
                            F.interpolate(xx[:, :, c:], (frames - w, dim), mode="bicubic", align_corners=True)], 2)
        xx = xx.view(batch, frames, dim)
        for pos, ln, n, axis in ((freq_pos, freq_len, n_freq, 2), (time_pos, time_len, n_time, 1)):
            if n == 0:
                continue
            pp, ll = _arr(pos, batch * n, np.int32).reshape(batch, n), _arr(ln, batch * n, np.int32).reshape(batch, n)
            for b in range(batch):
                for j in range(n):
                    if axis == 2:
                        xx[b, :, pp[b, j]: pp[b, j] + ll[b, j]] = fill
                    else:
                        xx[b, pp[b, j]: pp[b, j] + ll[b, j], :] = fill
        _tarr(out, batch * frames * dim, torch.float32)[:] = xx.flatten()

    def stac_ctc_loss(self, log_probs, targets, input_len, target_len, batch, frames, vocab, max_targets, blank,
                      reduction, nll, loss, stream):
        import torch.nn.functional as F
        lp = _tarr(log_probs, batch * frames * vocab, torch.float32).view(batch, frames, vocab)
        tg = _tarr(targets, batch * max_targets, torch.int32).view(batch, max_targets).long()
        il, tl = _tarr(input_len, batch, torch.int32).long(), _tarr(target_len, batch, torch.int32).long()
        per = F.ctc_loss(lp.transpose(0, 1), tg, il, tl, blank, reduction="none", zero_infinity=True)
        _tarr(nll, batch, torch.float32)[:] = per
        if reduction == 1:
            _tarr(loss, 1, torch.float32)[0] = per.sum()
        elif reduction == 2:
            _tarr(loss, 1, torch.float32)[0] = (per / tl.clamp(min=1)).mean()
        elif reduction == 3:
            _tarr(loss, 1, torch.float32)[0] = per.sum() / batch
        elif reduction == 4:
            _tarr(loss, batch, torch.float32)[:] = per / tl

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
    from stac_speech_translation_b200 import augment, decoder, ingest, losses, ops, turns
    emu = Emulator()
    for mod in (ops, decoder, turns, ingest, augment, losses):
        monkeypatch.setattr(mod, "ptr", cpu_ptr)
        monkeypatch.setattr(mod, "stream", lambda: c_void_p(0))
    for mod in (turns, ingest, augment, losses):
        monkeypatch.setattr(mod, "lib", lambda: EmuLib(emu))
    monkeypatch.setattr(ops, "_call", emu.call)
    return emu


def install_simt(monkeypatch, simt_lib):
    """Like install(), but every entry point the CPU build of the SIMT kernels exports (tests/simt_cpu) runs THAT - the
    kernel source itself under the CPU emulation of the CUDA execution model - instead of the numpy stand-in."""
    emu = install(monkeypatch)
    from stac_speech_translation_b200 import ops
    routed = []

    def call(name, *args):
        fn = getattr(simt_lib, name, None)
        if fn is None:
            return emu.call(name, *args)
        res, sig = _lib._SIGNATURES[name]
        fn.restype, fn.argtypes = res, sig
        emu.calls.append(name)
        routed.append(name)
        rc = fn(*args)
        assert rc == 0, (name, rc)

    monkeypatch.setattr(ops, "_call", call)
    # turns.py / ingest.py call the handle directly: give them the CPU build with the ctypes signatures applied
    from stac_speech_translation_b200 import augment, ingest, losses, turns
    for name, (res, sig) in _lib._SIGNATURES.items():
        fn = getattr(simt_lib, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = res, sig
    for mod in (turns, ingest, augment, losses):
        monkeypatch.setattr(mod, "lib", lambda: simt_lib)
    emu.routed = routed
    return emu
