"""The CUDA encoder against outputs of the reference's own glue code (GPU only).

tests/golden/glue_reference.npz holds what /root/reference/stac-st/modules/TransformerMultiTask.py
itself returned from encode() (:273-309), forward() (:144-183, make_masks :211-232) on seeded inputs
(generator: tests/golden/make_glue_golden.py).  Weights travel through load_state_dict with the
reference's own key names, as a checkpoint would."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import BF16_TOL, FP32_TOL, rel_l2  # noqa: E402
import stac_speech_translation_b200 as sb  # noqa: E402


def _fixture():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "glue_reference.npz"))
    state = {k[len("state/"):]: torch.from_numpy(d[k].astype(np.float32)) for k in d.files if k.startswith("state/")}
    return d, state


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_encoder_against_reference_glue_outputs(precision, tol):
    d, state = _fixture()
    d_model = state["encoder.norm.norm.weight"].shape[0]
    tr = sb.TransformerMultiTask(
        tgt_vocab=64, input_size=state["custom_src_module.layers.0.w.weight"].shape[1], d_model=d_model,
        nhead=d_model // 64, num_encoder_layers=2, num_decoder_layers=1,
        d_ffn=state["encoder.layers.0.pos_ffn.ffn.0.weight"].shape[0], dropout=0.1, activation=torch.nn.GELU,
        encoder_module="transformer", attention_type="regularMHA", normalize_before=True, causal=False,
        precision=precision)
    res = tr.load_state_dict(state, strict=False)
    # (the fixture holds the encoder side only; the decoder side has its own, decoder_reference.npz)
    assert not res.unexpected_keys and all(
        k == "positional_encoding.pe" or k.startswith(("decoder.", "custom_tgt_module.")) for k in res.missing_keys)
    tr = tr.cuda().eval()
    src = torch.from_numpy(d["src"].astype(np.float32)).cuda()
    wl = torch.from_numpy(d["wav_lens"]).cuda()
    n_valid = [int(np.floor(np.float32(w) * np.float32(src.shape[1]))) + 1 for w in d["wav_lens"]]

    def valid(x, n_list):      # compare the frames the reference itself attends to
        return torch.cat([x[i, :min(n, x.shape[1])] for i, n in enumerate(n_list)])

    enc = tr.encode(src, wl)
    assert enc.dtype == torch.float32 and enc.shape == d["enc_encode"].shape
    assert rel_l2(valid(enc.cpu(), n_valid), valid(torch.from_numpy(d["enc_encode"]), n_valid)) < tol
    assert rel_l2(enc, torch.from_numpy(d["enc_encode"])) < tol          # padded query rows too
    enc3 = tr.encode(src.reshape(src.shape[0], src.shape[1], -1))
    assert rel_l2(enc3, torch.from_numpy(d["enc_encode_nolen"])) < tol
    fwd = tr.forward_encoder(src, wl)
    assert rel_l2(fwd, torch.from_numpy(d["enc_forward"])) < tol
    assert rel_l2(sb.EncoderWrapper(tr)(src, wl), torch.from_numpy(d["enc_encode"])) < tol
