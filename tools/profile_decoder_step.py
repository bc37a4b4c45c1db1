import os, sys, torch
sys.path.insert(0, "/root/repo")
import stac_speech_translation_b200 as sb
torch.manual_seed(0)
tr = sb.TransformerMultiTask(tgt_vocab=5000, input_size=5120, d_model=256, nhead=4, num_encoder_layers=1, num_decoder_layers=6, d_ffn=1024, activation=torch.nn.GELU, normalize_before=True, precision="bf16").eval().cuda()
rows, beam, frames, prefix = 640, 10, 751, 32
enc = torch.randn(rows // beam, frames, 256, device="cuda")
tok = torch.randint(1, 5000, (rows, prefix), device="cuda")
cache = tr.decoder_cache(enc, rows=rows, max_len=prefix + 8, precision="bf16")
for t in range(prefix):
    cache.step(tok[:, t].contiguous())
torch.cuda.synchronize()
