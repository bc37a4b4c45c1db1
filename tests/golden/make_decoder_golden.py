"""Generates tests/golden/decoder_reference.npz by running the REFERENCE'S OWN decoder-side glue.

Like make_glue_golden.py: /root/reference/stac-st/modules/TransformerMultiTask.py is imported unmodified and its
``forward()`` (:144-209, decoder half included) and ``decode()`` (:234-271) are executed on seeded inputs; the
SpeechBrain names the file imports are served from a stub package made of the oracle's restatements - this time with the
real (restated) TransformerDecoder, NormalizedEmbedding, get_lookahead_mask and get_key_padding_mask instead of a
pass-through decoder.  What the fixture pins is therefore the in-repo part: which masks reach the decoder in ``forward``
(look-ahead + ``tgt == pad_idx`` + the round-rule key padding of the memory) and in ``decode`` (look-ahead only; memory
padding only when ``enc_len`` is given - the beam searcher's ``forward_step`` gives none, mutitask_decoder.py:126), the
sqrt(d) embedding scale, the positional-encoding add, and that ``decode`` returns the LAST layer's head-averaged
cross-attention weights.  SpeechBrain's arithmetic underneath stays "parity unpinned" (oracle/__init__.py).

Only this script reads /root/reference; the tests read the committed .npz.
Run from the repo root:  python tests/golden/make_decoder_golden.py
"""
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import speechbrain_path as sp  # noqa: E402
import make_glue_golden as glue  # noqa: E402

GOLDEN = os.path.join(HERE, "decoder_reference.npz")
CFG = dict(tgt_vocab=64, input_size=5 * 16, d_model=128, nhead=2, num_encoder_layers=1, num_decoder_layers=2,
           d_ffn=128, dropout=0.1, activation=nn.GELU, encoder_module="transformer", attention_type="regularMHA",
           normalize_before=True, causal=False)
SEED_W, SEED_X = 8886, 515
B, T2, L = 3, 21, 6
WAV_LENS = [1.0, 0.7304, 0.5]


class _TransformerInterface(nn.Module):
    """TransformerInterface with both halves built from the oracle's restated SpeechBrain classes, in SpeechBrain's
    construction order (positional encoding, encoder, decoder)."""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, d_ffn=2048, dropout=0.1,
                 activation=nn.ReLU, custom_src_module=None, custom_tgt_module=None,
                 positional_encoding="fixed_abs_sine", normalize_before=True, kernel_size=31, bias=True,
                 encoder_module="transformer", conformer_activation=None, attention_type="regularMHA",
                 max_length=2500, causal=False, **_unused):
        super().__init__()
        assert encoder_module == "transformer" and attention_type == "regularMHA" and not causal
        self.causal = causal
        self.attention_type = attention_type
        self.positional_encoding_type = positional_encoding
        self.positional_encoding = sp.PositionalEncoding(d_model, max_length)
        self.encoder = sp.TransformerEncoder(num_encoder_layers, nhead, d_ffn, d_model, dropout, activation,
                                             normalize_before)
        self.decoder = sp.TransformerDecoder(num_decoder_layers, nhead, d_ffn, d_model, dropout, activation,
                                             normalize_before)


def load_reference_module():
    glue._install_stub()
    m = sys.modules["speechbrain.lobes.models.transformer.Transformer"]
    m.TransformerInterface = _TransformerInterface
    m.NormalizedEmbedding = sp.NormalizedEmbedding
    m.get_key_padding_mask = sp.get_key_padding_mask
    m.get_lookahead_mask = sp.get_lookahead_mask
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_TransformerMultiTask_dec", glue.REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def generate():
    torch.set_num_threads(1)
    ref = load_reference_module()
    torch.manual_seed(SEED_W)
    model = ref.TransformerMultiTask(**CFG).eval()          # the reference class, its own _init_params
    with torch.no_grad():                                    # fp16-representable values: half-size fixture
        for p in model.parameters():
            p.copy_(p.half().float())
    g = torch.Generator().manual_seed(SEED_X)
    src = torch.randn(B, T2, 5, 16, generator=g).half().float()
    wl = torch.tensor(WAV_LENS)
    tgt = torch.randint(1, CFG["tgt_vocab"], (B, L), generator=g)
    tgt[1, 4:] = 0                                          # padded targets (pad_idx 0) for forward()
    tgt[2, 5:] = 0
    prefix = torch.randint(1, CFG["tgt_vocab"], (B, 4), generator=g)     # beam-search style prefix, no padding
    enc_len = torch.tensor([21, 15, 11])
    with torch.no_grad():
        enc_forward, dec_forward = model(src, tgt, wl, pad_idx=0)        # train_multitask.py:70-72
        enc_out = model.encode(src, wl)                                   # inference.py:100
        pred, attn = model.decode(prefix, enc_out)                        # mutitask_decoder.py:126 (no enc_len)
        pred_len, attn_len = model.decode(prefix, enc_out, enc_len)       # decode() with lengths
        pred1, attn1 = model.decode(prefix[:, :1], enc_out)               # first step: a single token
    state = {k: v.numpy() for k, v in model.state_dict().items()}
    out = dict(src=src.numpy(), wav_lens=wl.numpy(), tgt=tgt.numpy(), prefix=prefix.numpy(), enc_len=enc_len.numpy(),
               enc_forward=enc_forward.numpy(), dec_forward=dec_forward.numpy(), enc_out=enc_out.numpy(),
               pred=pred.numpy(), attn=attn.numpy(), pred_len=pred_len.numpy(), attn_len=attn_len.numpy(),
               pred1=pred1.numpy(), attn1=attn1.numpy())
    return out, state


if __name__ == "__main__":
    out, state = generate()
    half = lambda v: v.astype(np.float16)
    np.savez_compressed(GOLDEN, **{k: half(v) if k == "src" else v for k, v in out.items()},
                        **{"state/" + k: half(v) for k, v in state.items() if not k.endswith(".pe")})
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
    print({k: v.shape for k, v in out.items()})
