"""Generator of tests/golden/ctc_peaky.npz: a PEAKY CTC head for the greedy-agreement criterion of north_star
("identical CTC-greedy token sequences on >= 99.5 % of utterances", bf16 mode).

With untrained weights the oracle's own top-2 margin over 5000 classes is below bf16 resolution on half the frames, so
sequence agreement measures noise (hard part 5).  No checkpoint is published and no corpus can be
fetched, so this script makes the posteriors peaky the way training would: a brief, deterministic CTC fit of the top of
the ORACLE (plain PyTorch, CPU) on tone-coded synthetic audio -

  * audio: every utterance is a sequence of 4..14 "words"; a word is a 0.20-0.36 s two-partial tone whose pitch encodes
    its token (24 tone tokens + the two turn symbols the reference's RTTM code looks for, ids 7 and 8,
    /root/reference/stac-st/inference.py:48-63), separated by 60-160 ms of near-silence (synth_tone_utterance below);
  * model: the S model with its seeded random weights (oracle.build_reference_modules, seed 8886); everything up to and
    including encoder layer 10 stays as drawn; encoder layer 11, the final LayerNorm and the rows of ``ctc_lin`` that
    belong to the 27 symbols in use are fitted (Adam, full batch, fixed seed) first with a frame-wise cross-entropy on the synthesis alignment (the
    all-blank plateau of CTC is not worth the CPU minutes), then with torch.nn.functional.ctc_loss plus a top-2 margin
    term; the
    biases of the 4973 unused symbols are lowered by a constant so that they stay in the softmax without competing;
  * fixture: the fitted tensors (fp16 storage for the layer's matrices) + the utterance recipe (seed) - the audio itself
    is regenerated from the seed by the tests (synth_tone_batches), not stored.

Run from the repo root:  python tests/golden/make_ctc_peaky_fixture.py   (about 10 minutes on 8 cores)
The tests (tests/test_oracle.py::test_ctc_peaky_fixture_decodes, tests/test_gpu_ctc_peaky.py) load the fixture into the
oracle with `load_peaky`, and from there into the product through state_dict as a checkpoint would be.
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))

SR = 16000
TONE_IDS = list(range(9, 33))           # 24 tone tokens
TURN, XT = 7, 8                         # transformer_multitask.yaml:138-149
SYMBOLS = [0, TURN, XT] + TONE_IDS      # blank first
UNUSED_BIAS_SHIFT = 10.0
N_UTT, BATCH, SEED = 240, 24, 4242
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ctc_peaky.npz")


def _pitch(sym_index: int) -> float:
    """Pitch of symbol number `sym_index` (0..25 over [turn, xt] + tone tokens): a sixth of an octave apart."""
    return 280.0 * 2.0 ** (sym_index / 6.0)


def synth_tone_utterance(gen: torch.Generator):
    """(waveform fp32, token list, [(first sample, end sample)] of every word).  Words: raised-cosine-edged two-partial
    tones; gaps: -60 dB noise."""
    n_words = int(torch.randint(4, 15, (), generator=gen))
    toks, pieces, spans = [], [], []
    pieces.append(0.001 * torch.randn(int(0.1 * SR), generator=gen))
    pos = pieces[0].numel()
    prev = -1
    for _ in range(n_words):
        k = int(torch.randint(0, 26, (), generator=gen))
        if k == prev:                                   # CTC needs a blank between repeats: keep neighbours distinct
            k = (k + 1 + int(torch.randint(0, 25, (), generator=gen))) % 26
        prev = k
        toks.append(([TURN, XT] + TONE_IDS)[k])
        n = int((0.20 + 0.16 * float(torch.rand((), generator=gen))) * SR)
        t = torch.arange(n, dtype=torch.float64) / SR
        f = _pitch(k)
        sig = torch.sin(2 * math.pi * f * t) + (0.5 * torch.sin(2 * math.pi * 2 * f * t + 1.0) if 2 * f < 7600 else 0.0)
        edge = int(0.02 * SR)
        env = torch.ones(n, dtype=torch.float64)
        ramp = 0.5 - 0.5 * torch.cos(math.pi * torch.arange(edge, dtype=torch.float64) / edge)
        env[:edge], env[-edge:] = ramp, ramp.flip(0)
        pieces.append((0.08 * sig * env).float() + 0.001 * torch.randn(n, generator=gen))
        spans.append((pos, pos + n))
        g = int((0.06 + 0.10 * float(torch.rand((), generator=gen))) * SR)
        pieces.append(0.001 * torch.randn(g, generator=gen))
        pos += n + g
    pieces.append(0.001 * torch.randn(int(0.1 * SR), generator=gen))
    return torch.cat(pieces), toks, spans


def synth_tone_batches(n_utt=N_UTT, batch=BATCH, seed=SEED):
    """Length-sorted batches [(wavs [B, Lmax] zero-padded right, wav_lens [B] = len / Lmax, targets, word spans)] - the
    PaddedBatch contract of the reference's compute_forward (inference.py:91-92)."""
    gen = torch.Generator().manual_seed(seed)
    utts = [synth_tone_utterance(gen) for _ in range(n_utt)]
    utts.sort(key=lambda u: u[0].numel())
    out = []
    for i in range(0, n_utt, batch):
        chunk = utts[i:i + batch]
        lmax = max(u[0].numel() for u in chunk)
        lmax = (lmax + 3) // 4 * 4
        wavs = torch.zeros(len(chunk), lmax)
        for j, u in enumerate(chunk):
            wavs[j, : u[0].numel()] = u[0]
        wl = torch.tensor([u[0].numel() / lmax for u in chunk], dtype=torch.float32)
        out.append((wavs, wl, [u[1] for u in chunk], [u[2] for u in chunk]))
    return out


def load_peaky(omods, path=OUT):
    """Put the fitted tensors of the fixture into an oracle module graph built with the fixture's seed."""
    z = np.load(path)
    tr, ctc = omods["Transformer"], omods["ctc_lin"]
    layer = tr.encoder.layers[-1]
    sd = layer.state_dict()
    for k in sd:
        sd[k] = torch.from_numpy(z["layer." + k].astype(np.float32))
    layer.load_state_dict(sd)
    tr.encoder.norm.norm.weight.data = torch.from_numpy(z["norm.weight"])
    tr.encoder.norm.norm.bias.data = torch.from_numpy(z["norm.bias"])
    sym = torch.from_numpy(z["symbols"]).long()
    w, b = ctc.w.weight.data, ctc.w.bias.data
    b -= float(z["unused_bias_shift"])
    w[sym] = torch.from_numpy(z["ctc.weight_rows"])
    b[sym] = torch.from_numpy(z["ctc.bias_rows"])
    return omods


def main():
    from util import oracle_modules
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    omods = oracle_modules("S")
    tr, ctc = omods["Transformer"], omods["ctc_lin"]
    batches = synth_tone_batches()
    # ---- frozen part, once: everything below the last encoder layer ----
    cached = []
    with torch.no_grad():
        for wavs, wl, tg, spans in batches:
            feats = omods["normalize"](omods["compute_features"](wavs), wl)
            src = tr._embed(omods["CNN"](feats))
            abs_len = torch.floor(wl * src.shape[1])
            kpm = torch.arange(src.shape[1])[None, :].to(abs_len) > abs_len[:, None]
            x = tr.custom_src_module(src)
            x = x + tr.positional_encoding(x)
            for lyr in tr.encoder.layers[:-1]:
                x, _ = lyr(x, src_key_padding_mask=kpm)
            n_frames = (abs_len + 1).clamp(max=src.shape[1]).long()
            # frame labels for the warm start: the middle half of every word carries its symbol, everything else blank
            # (one encoder frame = 640 samples, frame j is centred on sample 640 j)
            lab = torch.zeros(x.shape[0], x.shape[1], dtype=torch.long)
            for i, (seq, sp) in enumerate(zip(tg, spans)):
                for tok, (s0, s1) in zip(seq, sp):
                    q = (s1 - s0) // 4
                    j0, j1 = math.ceil((s0 + q) / 640), (s1 - q) // 640
                    lab[i, j0:max(j1, j0) + 1] = SYMBOLS.index(tok)
            cached.append((x, kpm, n_frames, tg, lab))
            print(f"frozen pass: batch of {wavs.shape[0]} x {wavs.shape[1] / SR:.2f} s -> {tuple(x.shape)}", flush=True)
    # ---- the fit ----
    layer, norm = tr.encoder.layers[-1], tr.encoder.norm
    sym = torch.tensor(SYMBOLS)
    w_rows = ctc.w.weight.data[sym].clone().requires_grad_(True)
    b_rows = ctc.w.bias.data[sym].clone().requires_grad_(True)
    with torch.no_grad():
        ctc.w.bias.data -= UNUSED_BIAS_SHIFT
    mask_unused = torch.ones(ctc.w.weight.shape[0], dtype=torch.bool)
    mask_unused[sym] = False
    w_un, b_un = ctc.w.weight.data[mask_unused], ctc.w.bias.data[mask_unused]
    remap = {s: i for i, s in enumerate(SYMBOLS)}
    params = list(layer.parameters()) + list(norm.parameters()) + [w_rows, b_rows]
    for p in params:
        p.requires_grad_(True)
    layer.eval(); norm.eval()                       # dropout off: the fit is of the inference function
    opt = torch.optim.Adam(params, lr=2e-3)
    steps = int(os.environ.get("PEAKY_STEPS", "400"))
    for step in range(steps):
        if step == int(steps * 0.7):
            for g in opt.param_groups:
                g["lr"] = 5e-4
        opt.zero_grad()
        total, n_ok, n_all = 0.0, 0, 0
        warm = step < steps // 2                     # first half: frame-wise cross-entropy on the synthesis alignment
        for x, kpm, n_frames, tg, lab in cached:
            y, _ = layer(x, src_key_padding_mask=kpm)
            y = norm(y)
            used = y @ w_rows.t() + b_rows                                    # [B, T, 27]
            lse_un = torch.logsumexp(y.detach() @ w_un.t() + b_un, dim=-1, keepdim=True)   # unused symbols: constants
            logp = used - torch.logsumexp(torch.cat([used, lse_un], -1), dim=-1, keepdim=True)
            tgt = torch.tensor([remap[t] for seq in tg for t in seq])
            tl = torch.tensor([len(seq) for seq in tg])
            valid = torch.arange(logp.shape[1])[None, :] < n_frames[:, None]
            if warm:
                loss = (F.nll_loss(logp.reshape(-1, logp.shape[-1]), lab.reshape(-1), reduction="none")
                        * valid.reshape(-1)).sum()
                loss.backward()
            else:
                loss = F.ctc_loss(logp.transpose(0, 1), tgt, n_frames, tl, blank=0, reduction="sum",
                                  zero_infinity=True)
                # margin term: push the winning symbol of every frame away from the runner-up (sharp spikes)
                top2 = logp.topk(2, dim=-1).values
                margin_pen = (F.relu(3.0 - (top2[..., 0] - top2[..., 1])) * valid).sum()
                (loss + 0.2 * margin_pen).backward()
            total += float(loss)
            with torch.no_grad():
                ids = logp.argmax(-1)
                for i, seq in enumerate(tg):
                    prev, dec = None, []
                    for tkn in ids[i, : int(n_frames[i])].tolist():
                        if tkn != prev and tkn != 0:
                            dec.append(SYMBOLS[tkn])
                        prev = tkn
                    n_ok += dec == seq
                    n_all += 1
        opt.step()
        if step % 10 == 0 or step == steps - 1:
            print(f"step {step:4d}  {'ce ' if warm else 'ctc'} loss {total / N_UTT:8.3f}  greedy == target on {n_ok}/{n_all} utterances", flush=True)
    out = {"layer." + k: (v.detach().numpy().astype(np.float16) if v.dim() > 1 else v.detach().numpy())
           for k, v in layer.state_dict().items()}
    out["norm.weight"] = norm.norm.weight.detach().numpy()
    out["norm.bias"] = norm.norm.bias.detach().numpy()
    out["ctc.weight_rows"] = w_rows.detach().numpy()
    out["ctc.bias_rows"] = b_rows.detach().numpy()
    out["symbols"] = np.array(SYMBOLS, dtype=np.int64)
    out["unused_bias_shift"] = np.float32(UNUSED_BIAS_SHIFT)
    out["recipe"] = np.array([N_UTT, BATCH, SEED], dtype=np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
