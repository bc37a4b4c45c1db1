"""Generates tests/golden/glue_reference.npz by running the REFERENCE'S OWN glue code.

/root/reference/stac-st/modules/TransformerMultiTask.py is imported unmodified from where it lies
and its ``encode()`` (:273-309), ``forward()`` encoder half (:144-183) with ``make_masks()``
(:211-232), ``_init_params()`` (:311-314) and ``EncoderWrapper`` (:317-349) are executed on seeded
inputs.  The file imports ``speechbrain`` (un-vendored, not installable here), so the seven
SpeechBrain names it needs are served from a stub package whose classes are the oracle's
restatements (``oracle/speechbrain_path.py``): what this fixture pins is therefore the in-repo part
of the path - reshape order, both key-padding mask rules in fp32, src-linear, positional-encoding
add, encoder call contract, xavier re-initialisation - executed by the reference itself; the
SpeechBrain arithmetic underneath stays "parity unpinned" (oracle/__init__.py).

Only this script reads /root/reference; the tests read the committed .npz.
Run from the repo root:  python tests/golden/make_glue_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import speechbrain_path as sp  # noqa: E402

REF_FILE = "/root/reference/stac-st/modules/TransformerMultiTask.py"
GOLDEN = os.path.join(ROOT, "tests", "golden", "glue_reference.npz")
CFG = dict(tgt_vocab=64, input_size=20 * 16, d_model=128, nhead=2, num_encoder_layers=2,
           num_decoder_layers=1, d_ffn=256, dropout=0.1, activation=nn.GELU,
           encoder_module="transformer", attention_type="regularMHA", normalize_before=True, causal=False)
SEED_W, SEED_X = 8886, 2024
B, T2 = 3, 37
WAV_LENS = [1.0, 0.7304, 0.5]       # 0.5 * 37 = 18.5: floor/round/half-to-even all differ


class _PassThroughDecoder(nn.Module):
    """The decoder is out of scope for the path; forward() only needs a callable."""

    def forward(self, tgt, memory, **kwargs):
        return tgt, [None], [None]


class _TransformerInterface(nn.Module):
    """speechbrain.lobes.models.transformer.Transformer.TransformerInterface, encoder side, built
    from the oracle's restated SpeechBrain classes."""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, d_ffn=2048,
                 dropout=0.1, activation=nn.ReLU, custom_src_module=None, custom_tgt_module=None,
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
        self.decoder = _PassThroughDecoder()


class _NormalizedEmbedding(nn.Module):
    def __init__(self, d_model, vocab):
        super().__init__()
        self.emb = nn.Embedding(vocab, d_model, padding_idx=0)
        self.d_model = d_model

    def forward(self, x):
        return self.emb(x) * self.d_model ** 0.5


def _install_stub():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("speechbrain")
    mod("speechbrain.dataio")
    mod("speechbrain.dataio.dataio", length_to_mask=sp.length_to_mask)
    mod("speechbrain.lobes")
    mod("speechbrain.lobes.models")
    mod("speechbrain.lobes.models.transformer")
    mod("speechbrain.lobes.models.transformer.Transformer",
        NormalizedEmbedding=_NormalizedEmbedding, TransformerInterface=_TransformerInterface,
        get_key_padding_mask=lambda tgt, pad_idx=0: tgt == pad_idx,
        get_lookahead_mask=lambda tgt: torch.triu(torch.ones(tgt.shape[1], tgt.shape[1]), 1).bool())
    mod("speechbrain.nnet")
    mod("speechbrain.nnet.activations", Swish=nn.SiLU)
    mod("speechbrain.nnet.containers", ModuleList=sp._Layers)
    mod("speechbrain.nnet.linear", Linear=sp.Linear)


def load_reference_module():
    _install_stub()
    spec = importlib.util.spec_from_file_location("ref_TransformerMultiTask", REF_FILE)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def generate():
    torch.set_num_threads(1)
    ref = load_reference_module()
    torch.manual_seed(SEED_W)
    model = ref.TransformerMultiTask(**CFG).eval()          # the reference class, its own _init_params
    with torch.no_grad():                                    # fp16-representable values: half-size fixture
        for p in model.parameters():
            p.copy_(p.half().float())
    g = torch.Generator().manual_seed(SEED_X)
    src = torch.randn(B, T2, 20, 16, generator=g).half().float()   # CNN-shaped input [B, T'', F, C]
    wl = torch.tensor(WAV_LENS)
    tgt = torch.randint(1, CFG["tgt_vocab"], (B, 5), generator=g)
    with torch.no_grad():
        enc_encode = model.encode(src, wl)                               # inference.py:100
        enc_encode_nolen = model.encode(src.reshape(B, T2, 320))       # 3-D input, no lengths
        enc_forward, _ = model(src, tgt, wl, pad_idx=0)                  # train_multitask.py:70-72
        enc_wrapped = ref.EncoderWrapper(model)(src, wl)
    assert torch.equal(enc_wrapped, enc_encode)
    state = {k: v.numpy() for k, v in model.state_dict().items()
             if not k.startswith(("decoder.", "custom_tgt_module."))}
    return dict(src=src.numpy(), wav_lens=wl.numpy(), enc_encode=enc_encode.numpy(),
                enc_encode_nolen=enc_encode_nolen.numpy(), enc_forward=enc_forward.numpy()), state


if __name__ == "__main__":
    out, state = generate()
    half = lambda k, v: v.astype(np.float16) if v.dtype == np.float32 and not k.endswith(".pe") else v
    np.savez_compressed(GOLDEN, **{k: half(k, v) if k == "src" else v for k, v in out.items()},
                        **{"state/" + k: half(k, v) for k, v in state.items() if not k.endswith(".pe")})
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
