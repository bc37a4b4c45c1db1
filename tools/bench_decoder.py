"""Times one beam-search decoding step of the S model on the device, both ways (run on a B200):
  * the reference's way, mutitask_decoder.py:119-128: the whole decoder over the whole prefix (TransformerMultiTask.decode),
  * KV-cached (decoder.DecoderCache.step): one token per row over cached keys / values.
  python tools/bench_decoder.py [--batch 64] [--beam 10] [--frames 751] [--prefix 32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stac_speech_translation_b200 as sb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--beam", type=int, default=10)
ap.add_argument("--frames", type=int, default=751)
ap.add_argument("--prefix", type=int, default=32)
a = ap.parse_args()
torch.manual_seed(0)
tr = sb.TransformerMultiTask(tgt_vocab=5000, input_size=5120, d_model=256, nhead=4, num_encoder_layers=1,
                             num_decoder_layers=6, d_ffn=1024, activation=torch.nn.GELU, normalize_before=True,
                             precision="bf16").eval().cuda()
rows = a.batch * a.beam
enc = torch.randn(a.batch, a.frames, 256, device="cuda")
tok = torch.randint(1, 5000, (rows, a.prefix), device="cuda")


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


inflated = enc.repeat_interleave(a.beam, 0)
ms_full = timed(lambda: tr.decode(tok, inflated))
cache = tr.decoder_cache(enc, rows=rows, max_len=a.prefix + 8)
for t in range(a.prefix - 1):
    cache.step(tok[:, t].contiguous())
t0 = cache.t


def one_step():
    cache.rewind(t0)
    cache.step(tok[:, t0].contiguous())


ms_step = timed(one_step)
cache16 = tr.decoder_cache(enc, rows=rows, max_len=a.prefix + 8, precision="bf16")
for t in range(a.prefix - 1):
    cache16.step(tok[:, t].contiguous())


def one_step16():
    cache16.rewind(t0)
    cache16.step(tok[:, t0].contiguous())


ms_step16 = timed(one_step16)
# the same bf16 step with graph=True: DecoderCache replays one captured launch sequence per step (the position counter
# lives on the device), so this is what a searcher's forward_step costs, tokens in / prediction out
cache_g = tr.decoder_cache(enc, rows=rows, max_len=a.prefix + 8, precision="bf16", graph=True)
for t in range(a.prefix - 1):
    cache_g.step(tok[:, t].contiguous())


def one_step_graph():
    cache_g.rewind(t0)
    cache_g.step(tok[:, t0].contiguous())


ms_graph16 = timed(one_step_graph)
ref_o, ref_w = cache16.step(tok[:, t0].contiguous()) if cache16.rewind(t0) is None else None
got_o, got_w = cache_g.step(tok[:, t0].contiguous()) if cache_g.rewind(t0) is None else None
assert torch.equal(ref_o, got_o) and torch.equal(ref_w, got_w), "graph replay differs from the eager step"
print(f"rows {rows}  prefix {a.prefix}  frames {a.frames}:  full-prefix decode {ms_full:.3f} ms   cached step {ms_step:.3f} ms"
      f"   cached step, bf16 GEMMs {ms_step16:.3f} ms   the same with graph=True (one replay per step) {ms_graph16:.3f} ms")
# beam re-ordering: the row map (what reorder() does) against gathering the cached prefix (what it did at first)
idx = torch.randint(0, rows, (rows,), device="cuda")
ms_reorder = timed(lambda: cache16.reorder(idx))
ms_step_mapped = timed(one_step16)
ms_gather = timed(lambda: cache16.self_kv[:, :t0].copy_(cache16.self_kv[:, :t0].index_select(2, idx)))
print(f"beam re-ordering at prefix {t0}: row map {ms_reorder:.3f} ms (then a step through the map {ms_step_mapped:.3f} ms)"
      f"   gather of the cached prefix {ms_gather:.3f} ms")
if os.environ.get("STAC_DECODER_PROFILE", "1") != "0":
    # warm per-kernel device times of the graph-less bf16 step (CUPTI through torch.profiler; a breakdown, not a bench value)
    try:
        from torch.profiler import ProfilerActivity, profile
        n_prof = 5
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(n_prof):
                one_step16()
            torch.cuda.synchronize()
        rows_ = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        print("kernel, launches per step, us per launch, us per step")
        for e in rows_[:12]:
            print(f"  {e.key[:70]:70s} {e.count / n_prof:6.1f} {e.device_time_total / max(e.count, 1):9.2f} "
                  f"{e.device_time_total / n_prof:9.1f}")
    except Exception as exc:  # noqa: BLE001 - the breakdown is optional
        print("profiler breakdown unavailable:", exc)
