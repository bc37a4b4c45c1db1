"""Smallest case that touches every kernel of the KV-cached decoder step (written for compute-sanitizer memcheck /
racecheck, which this pool refuses to run - kept as a quick stand-alone check: attention rows sum to one):
fp32 and bf16 caches, beam rows over un-inflated memory, a re-ordering, the weights-returning layer."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stac_speech_translation_b200 as sb  # noqa: E402

torch.manual_seed(0)
tr = sb.TransformerMultiTask(tgt_vocab=97, input_size=5120, d_model=256, nhead=4, num_encoder_layers=1,
                             num_decoder_layers=2, d_ffn=512, activation=torch.nn.GELU, normalize_before=True,
                             precision="bf16").eval().cuda()
bm, beam, frames = 2, 3, 70
rows = bm * beam
enc = torch.randn(bm, frames, 256, device="cuda")
tok = torch.randint(1, 97, (rows, 6), device="cuda")
for precision in ("fp32", "bf16"):
    cache = tr.decoder_cache(enc, rows=rows, max_len=8, precision=precision)
    for t in range(6):
        if t == 3:
            cache.reorder(torch.arange(rows, device="cuda").roll(1))
        out, w = cache.step(tok[:, t].contiguous())
    torch.cuda.synchronize()
    print(precision, "ok", float(out.abs().mean()), float(w.sum(-1).mean()))
pred, attn = tr.decode(tok, enc.repeat_interleave(beam, 0))
torch.cuda.synchronize()
print("decode ok", float(pred.abs().mean()))
