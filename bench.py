#!/usr/bin/env python
"""Benchmark of the STAC-ST encoder-side hot path on B200 (see BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the whole path (PCM -> Fbank -> normalise -> CNN -> encoder -> CTC
log-posteriors + greedy ids) over one synthetic batch of the workload BASELINE.json's metric is
quoted on that fits one GPU: configs[1], default ("S") model, 64 x 30 s multi-turn segments, bf16.
Prints ONE JSON line (rank 0).  `value` is measured with the batch resident in HBM; `e2e` goes
through the same public call with pinned HOST buffers (H2D of the PCM and D2H of the greedy token
ids inside the timed region).  `--impl reference` times the CPU oracle (plain-PyTorch restatement
of the reference's SpeechBrain path, oracle/) on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encoder audio-sec/sec (RTFx)"
UNIT = "audio-s/s"
VOCAB = 5000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="index into BASELINE.json configs: 1 = S model, 64 x 30 s (the metric's configuration, default); "
                         "2 = M model, length-bucketed ragged batches sharded over the ranks; 3 = L model, 16 x 45-60 s "
                         "per GPU; 4 = front-end only (Fbank + normalise + CNN) sweep over 10 k utterances of 1-30 s")
    ap.add_argument("--size", default=None, choices=["S", "M", "L"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--seconds", type=float, default=None)
    ap.add_argument("--min-seconds", type=float, default=None,
                    help="ragged batch: utterance lengths uniform in [min-seconds, seconds], zero-padded to --seconds")
    ap.add_argument("--utterances", type=int, default=None, help="configs 2 / 4: utterances in the synthetic corpus")
    ap.add_argument("--max-batch-len", type=float, default=200.0,
                    help="configs 2 / 4: seconds of audio per length bucket batch (DynamicBatchSampler's max_batch_len; "
                         "200 = the reference's M-size evaluation value, ablations/run_m_and_l_size.sh:87-88)")
    ap.add_argument("--streams", type=int, default=4,
                    help="configs 2 / 4: CUDA streams the batches of a rank are replayed on (a 200 s batch fills a "
                         "fraction of the GPU: independent batches run side by side)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--gather", default="bf16", choices=["ids", "bf16", "fp32"],
                    help="what travels to rank 0 besides enc_out when N > 1: greedy ids only, bf16 or fp32 posteriors")
    ap.add_argument("--gather-transport", default="push", choices=["push", "peer", "nccl"],
                    help="N > 1: every rank pushes its results into rank 0's peer-mapped buffers with its own copy "
                         "engines (default), rank 0 pulls them (peer), or NCCL send/recv")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="SMs left free for the NCCL transfer kernels when N > 1 (default 0: measured no gain at N=8, "
                         "the gather is bound by NCCL's per-peer point-to-point bandwidth, not by SM contention)")
    ap.add_argument("--pcm", default="int16", choices=["int16", "fp32"],
                    help="what the end-to-end leg ships from the host: 16-bit PCM as a wav file holds it (decoded on the "
                         "device, sample / 32768, SURVEY.md 8f-2) or the fp32 waveform the reference's DataLoader builds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=8, help="utterances in the CPU-baseline sample")
    ap.add_argument("--trace-out", default=None, help="write the per-kernel timing table (json) here")
    args = ap.parse_args()
    defaults = {1: ("S", 64, 30.0, None), 2: ("M", None, 30.0, None), 3: ("L", 16, 60.0, 45.0), 4: ("S", None, 30.0, None)}
    size, batch, seconds, min_s = defaults[args.config]
    args.size = args.size or size
    args.batch = args.batch or batch
    args.seconds = args.seconds or seconds
    args.min_seconds = args.min_seconds if args.min_seconds is not None else min_s
    if args.config == 3 and args.cpu_batch == 8:
        args.cpu_batch = 2                          # L model, 60 s: two utterances are ~20 s of CPU work per pass
    if args.utterances is None:
        args.utterances = {2: 4096, 4: 10000}.get(args.config)
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------
# algorithmic work (SURVEY.md 8d / BASELINE.md section 2)
# --------------------------------------------------------------------------
def work_table(size_cfg, batch, n_samples, kv_len_sum_sq_like):
    d, layers, dffn = size_cfg["d_model"], size_cfg["num_encoder_layers"], size_cfg["d_ffn"]
    t = 1 + n_samples // 160
    t1 = (t - 1) // 2 + 1
    t2 = (t1 - 1) // 2 + 1
    m = batch * t2
    w = {
        "stac_fbank_logmel": ("hbm", batch * (4 * n_samples + 4 * 80 * t), 1),
        "stac_fbank_logmel_tc": ("hbm", batch * (4 * n_samples + 4 * 80 * t), 1),
        "stac_fbank_logmel_tc2": ("hbm", batch * (4 * n_samples + 4 * 80 * t), 1),
        "stac_fbank_topdb_norm": ("hbm", batch * 2 * 4 * 80 * t, 1),
        "stac_conv0_ln_lrelu": ("hbm", batch * (4 * 80 * t + 2 * 40 * 256 * t1), 1),
        "stac_conv0_topdb_norm_bf16": ("hbm", batch * (4 * 80 * t + 2 * 40 * 256 * t1), 1),
        "stac_conv1_bf16": ("tensor", 2 * 9 * 256 * 256 * 20 * m, 1),
        "stac_group_ln_lrelu": ("hbm", m * 5120 * (4 + 2), 1),
        "stac_gemm_bf16:src_linear": ("tensor", 2 * 5120 * d * m, 1),
        "stac_layernorm": ("hbm", m * d * (4 + 2), 2 * layers + 1),
        "stac_gemm_bf16:qkv": ("tensor", 2 * d * 3 * d * m, layers),
        "stac_mha_bf16": ("tensor", 4 * d * kv_len_sum_sq_like, layers),
        "stac_mha_bf16_v2": ("tensor", 4 * d * kv_len_sum_sq_like, layers),     # STAC_MHA_V2=1 (experimental kernel)
        "stac_gemm_bf16:out_proj": ("tensor", 2 * d * d * m, layers),
        # fused out-proj + residual + LayerNorm: HBM-bound (ctx bf16 in, x fp32 read + write, LN(x) bf16 out)
        "stac_outproj_ln_bf16": ("hbm", m * d * (2 + 4 + 4 + 2), layers),
        "stac_gemm_bf16:ffn1": ("tensor", 2 * d * dffn * m, layers),
        "stac_gemm_bf16:ffn2": ("tensor", 2 * d * dffn * m, layers),
        "stac_ffn_fused_bf16": ("tensor", 4 * d * dffn * m, layers),
        "stac_gemm_bf16:ctc_lin": ("tensor", 2 * d * VOCAB * m, 1),
        "stac_log_softmax": ("hbm", m * VOCAB * 4 * 2, 1),
        # fused CTC head: the fp32 posteriors are written once (bf16 states in); its two GEMM passes are 2x the FLOPs
        "stac_ctc_head_bf16": ("hbm", m * VOCAB * 4 + m * d * 2, 1),
    }
    return w


# --------------------------------------------------------------------------
def workload_config(args, t2, world):
    cfg_idx = {("S", 64, 30.0, None): "configs[1]", ("L", 16, 60.0, 45.0): "configs[3]"}.get(
        (args.size, args.batch, args.seconds, args.min_seconds), "custom")
    ragged = "" if args.min_seconds is None else f" (ragged: {args.min_seconds:g}-{args.seconds:g} s, key-padding masks)"
    return {"workload": f"{cfg_idx}: STAC-ST {args.size} encoder + CTC head, batch {args.batch} x "
                        f"{args.seconds:g} s multi-turn synthetic 16 kHz segments per GPU{ragged}",
            "precision": args.precision, "frames_25hz": t2,
            "arithmetic": ("bf16 mode: bf16 tensor-core GEMMs / attention with fp32 accumulation and an fp32 residual "
                           "stream; two stated deviations from the reference's fp32 formulas, both inside the 2e-2 "
                           "tolerance (see `parity`): the STFT is an fp16-operand tensor-core GEMM on the folded frame "
                           "(log-mel 4-6e-4 relative), GELU is a tanh-form minimax fit of erf-GELU"
                           if args.precision == "bf16" else "fp32 mode: CUDA-core kernels, exact formulas"),
            "l2": "per-step working set (~6 GB of activations) >> 126 MB L2, no explicit flush",
            "multi_gpu": ("whole batches per rank; enc_out + greedy ids"
                          + ("" if args.gather == "ids" else f" + {args.gather} posteriors")
                          + ({"peer": " pulled by rank 0 over NVLink peer memory (copy engines; torch.distributed gloo "
                                      "control messages, NCCL barrier)",
                              "push": " pushed by every rank's own copy engines into rank 0's peer-mapped buffers over "
                                      "NVLink (no communication kernel, one working context per GPU; gloo control "
                                      "messages, NCCL barrier)"}.get(getattr(args, "gather_transport", "push"),
                                                                     " sent to rank 0 with NCCL point-to-point"))
                          + " inside the timed region") if world > 1 else "single GPU"}


def run_reference(args, rank, world):
    """CPU arm: the reference's path on the host cores.  SpeechBrain cannot be installed here, so this is
    the oracle port of it (oracle/, plain PyTorch fp32, the reference's six eager stages and slow-path MHA)
    on all host threads; each step is a bounded sample (a few utterances) of the same workload."""
    if rank != 0:
        return
    import oracle
    from stac_speech_translation_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    omods = oracle.build_reference_modules(args.size)
    b = max(1, min(args.cpu_batch, args.batch))
    wavs, wl = synth.fast_synth_batch(b, args.seconds, seed=1234)
    norm = omods["normalize"]
    norm.train(); norm(omods["compute_features"](wavs[:, :32000]), wl); norm.eval()
    t0 = time.perf_counter()
    oracle.reference_compute_forward(omods, wavs, wl)
    first = time.perf_counter() - t0
    budget = 150.0                                   # seconds for warm-up + timed steps
    if first * (args.steps + args.warmup) > budget and b > 1:
        b = max(1, int(b * budget / (first * (args.steps + args.warmup))))
        wavs, wl = wavs[:b].contiguous(), wl[:b].contiguous()
    for _ in range(max(0, args.warmup - 1)):
        oracle.reference_compute_forward(omods, wavs, wl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.reference_compute_forward(omods, wavs, wl)
    dt = (time.perf_counter() - t0) / args.steps
    val = b * args.seconds / dt
    t2 = ((1 + int(args.seconds * 16000) // 160 - 1) // 2 + 1 - 1) // 2 + 1
    sample = (f"{b} x {args.seconds:g} s utterances per step ({args.steps} timed steps) of the {args.batch} x "
              f"{args.seconds:g} s workload; torch {torch.__version__} fp32, {cores} threads")
    cfg = workload_config(args, t2, 1)
    cfg["precision"] = "fp32"
    cfg["cpu_sample"] = sample
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline(args, mods=None, wavs=None, wl=None):
    """The oracle port timed on the host cores on a bounded sample of the workload.  When the product's module graph
    and the timed batch are given, the oracle is loaded with the PRODUCT's weights and normaliser statistics (through
    state_dict, as a checkpoint would be) and run on the first utterances of that very batch, and what it returns is
    handed back as the parity reference for the timed outputs (the oracle is the checker here, nothing else)."""
    import oracle
    from stac_speech_translation_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    omods = oracle.build_reference_modules(args.size)
    b = max(1, min(args.cpu_batch, args.batch))
    if mods is not None:
        for k in ("CNN", "Transformer", "ctc_lin"):
            sd = {k_: v.detach().float().cpu() for k_, v in mods[k].state_dict().items()}
            res = omods[k].load_state_dict(sd, strict=True)
            assert not res.missing_keys and not res.unexpected_keys
        st = mods["normalize"]._statistics_dict()
        omods["normalize"]._load_statistics_dict(
            {k_: (v.detach().float().cpu() if torch.is_tensor(v) else v) for k_, v in st.items()})
        omods["normalize"].eval()
        wavs, wl = wavs[:b].contiguous(), wl[:b].contiguous()
    else:
        wavs, wl = synth.fast_synth_batch(b, args.seconds, seed=1234)
        norm = omods["normalize"]
        norm.train(); norm(omods["compute_features"](wavs[:, :32000]), wl); norm.eval()
    t0 = time.perf_counter()
    ref = oracle.reference_compute_forward(omods, wavs, wl)
    warm = time.perf_counter() - t0
    reps = max(3, min(20, int(12.0 / max(warm, 1e-3))))          # about 10-15 s of CPU work
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.reference_compute_forward(omods, wavs, wl)
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(b * args.seconds / dt, 2), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{b} x {args.seconds:g} s utterances, {reps} timed passes after 1 warm-up, torch fp32, "
                      f"{cores} threads"}, ref


PARITY_TOL = {"bf16": 2e-2, "fp32": 1e-4}      # north_star: relative tolerance on encoder states / posteriors


def parity_block(res, ref, n, precision, wl=None):
    """The timed batch's outputs (first n utterances) against the oracle's on the same audio and weights, over the
    frames encode() keeps (j <= floor(wav_len * T2), TransformerMultiTask.py:289-294)."""
    from stac_speech_translation_b200.pipeline import ctc_greedy_collapse

    def rel(a, b_):
        a, b_ = a.detach().double().cpu(), b_.detach().double().cpu()
        return float((a - b_).norm() / b_.norm().clamp_min(1e-30))

    enc, p = res["enc_out"][:n].float().cpu(), res["p_ctc"][:n].float().cpu()
    ids = res["greedy"][:n].cpu().long()
    ref_ids = ref["p_ctc"].argmax(-1)
    t2 = ids.shape[1]
    keep = [t2] * n if wl is None else (torch.floor(wl[:n].float() * t2) + 1).clamp(max=t2).long().tolist()
    mask = torch.arange(t2)[None, :] < torch.tensor(keep)[:, None]
    seq = ctc_greedy_collapse(ids, keep)
    ref_seq = ctc_greedy_collapse(ref_ids, keep)
    top2 = ref["p_ctc"].topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    flipped = (ids != ref_ids) & mask
    tol = PARITY_TOL[precision]
    enc, p = enc * mask[..., None], p * mask[..., None]
    ref = {"enc_out": ref["enc_out"] * mask[..., None], "p_ctc": ref["p_ctc"] * mask[..., None]}
    out = {"utterances": n, "enc_rel_l2": round(rel(enc, ref["enc_out"]), 6),
           "pctc_rel_l2": round(rel(p, ref["p_ctc"]), 6),
           "pctc_max_abs": round(float((p - ref["p_ctc"]).abs().max()), 5),
           "greedy_frame": round(1.0 - float(flipped.sum()) / float(mask.sum()), 5),
           "greedy_seq": round(sum(a == b_ for a, b_ in zip(seq, ref_seq)) / n, 4),
           "max_margin_of_flipped_frame": round(float(margin[flipped].max()) if flipped.any() else 0.0, 5),
           "tolerance": tol,
           "note": "oracle (fp32 CPU restatement, product's weights) vs the timed batch; synthetic weights are untrained, "
                   "so greedy ids flip wherever the oracle's own top-2 margin is below the posterior error "
                   "(the trained-weight criterion is tests/test_gpu_ctc_peaky.py)"}
    out["ok"] = bool(out["enc_rel_l2"] <= tol and out["pctc_rel_l2"] <= tol)
    return out


def kernel_table(trace, n_steps, wt, pk):
    per = {}
    for name, label, e0, e1 in trace:
        k = ops_mod().trace_key(name, label)
        t, n = per.get(k, (0.0, 0))
        per[k] = (t + e0.elapsed_time(e1), n + 1)
    table = []
    for k, (t, n) in sorted(per.items(), key=lambda kv_: -kv_[1][0]):
        bound, work, _ = wt.get(k, ("hbm", 0, 1))
        avg_ms = t / n
        ach = work / (avg_ms * 1e-3) / (1e12 if bound == "tensor" else 1e9) if avg_ms > 0 else 0.0
        peak = pk["tflops_sustained"] if bound == "tensor" else pk["hbm_gbs"]
        table.append({"kernel": k, "launches_per_step": n // n_steps, "avg_ms": round(avg_ms, 4),
                      "ms_per_step": round(t / n_steps, 4), "bound": bound, "work_per_launch": work,
                      "achieved": round(ach, 1), "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                      "frac": round(ach / peak, 4)})
    return table


def timeline_table(sym, dur_us, wt, pk):
    """Kernel symbols of one graph replay, in launch order, with their device times -> rows labelled like the `kernels`
    table (work_table keys) with achieved rates.  Symbols that several entry points share are told apart by launch
    order: the weight-resident GEMM alternates QKV / out-proj layer by layer; of the three general linear GEMMs of the
    S path the first is the src-linear and the other two, with the row reduction between them, are one CTC head."""
    n_general = sum(1 for s_ in sym if s_.startswith("gemm_bf16_kernel<false>") or s_.startswith("gemm_bf16_kernel<0>"))
    fixed = {"gemm_bf16_kernel<true>": "stac_conv1_bf16", "gemm_bf16_kernel<1>": "stac_conv1_bf16",
             "mha2_bf16_kernel": "stac_mha_bf16_v2", "mha_bf16_kernel": "stac_mha_bf16",
             "ffn_fused_kernel": "stac_ffn_fused_bf16", "conv0_tc_kernel": "stac_conv0_ln_lrelu",
             "layernorm_kernel": "stac_layernorm", "fbank_tc2_kernel": "stac_fbank_logmel_tc2",
             "fbank_tc_kernel": "stac_fbank_logmel_tc", "topdb_norm_kernel": "stac_fbank_topdb_norm"}
    rows, n_wres, n_gen = {}, 0, 0
    for s_, us_ in zip(sym, dur_us):
        key, part = None, 1
        for prefix, label in fixed.items():
            if s_ == prefix or s_.startswith(prefix + "<"):
                key = label
        if s_.startswith("gemm_wres_kernel"):                       # QKV and out-proj alternate, layer by layer
            key = "stac_gemm_bf16:qkv" if n_wres % 2 == 0 else "stac_gemm_bf16:out_proj"
            n_wres += 1
        elif (s_.startswith("gemm_bf16_kernel<false>") or s_.startswith("gemm_bf16_kernel<0>")) and n_general == 3:
            key, part = ("stac_gemm_bf16:src_linear", 1) if n_gen == 0 else ("stac_ctc_head_bf16", 3)
            n_gen += 1
        elif s_.startswith("ctc_reduce_kernel") and n_general == 3:
            key, part = "stac_ctc_head_bf16", 3                     # (pass 1, reduce, pass 2 = one launch of the head)
        r = rows.setdefault(key or s_, [0.0, 0, part])
        r[0] += us_
        r[1] += 1
    table = []
    for k, (us, n, part) in sorted(rows.items(), key=lambda kv_: -kv_[1][0]):
        launches = max(n // part, 1)
        row = {"kernel": k, "launches_per_step": launches, "avg_us": round(us / launches, 2), "us_per_step": round(us, 1)}
        if k in wt:
            bound, work, _ = wt[k]
            ach = work / (us / launches * 1e-6) / (1e12 if bound == "tensor" else 1e9)
            peak = pk["tflops_sustained"] if bound == "tensor" else pk["hbm_gbs"]
            row.update({"bound": bound, "achieved": round(ach, 1), "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                        "frac": round(ach / peak, 4)})
        table.append(row)
    return table


def graph_timeline(graphed, wt, pk, n_rep=3):
    """Warm duration of every kernel INSIDE one replay of the path's CUDA graph, and the idle time between kernels (CUPTI
    through torch.profiler).  The `kernels` table and `roofline` are CUDA events around eager launches (events cannot sit
    inside a graph) and carry a few microseconds of launch overhead per launch; this is the same kernels as the timed
    replays run them.  A breakdown that explains `value`, never a bench value; absent when CUPTI is not available."""
    from torch.profiler import ProfilerActivity, profile
    g = graphed.graph
    g.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_rep):
            g.replay()
        torch.cuda.synchronize()
    ev = sorted((e for e in prof.events()
                 if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() > 0),
                key=lambda e: e.time_range.start)
    per = len(ev) // n_rep
    if per == 0 or len(ev) != per * n_rep:
        raise RuntimeError(f"{len(ev)} device events for {n_rep} replays")
    ev = ev[-per:]                                                  # the last replay
    sym = [e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0] for e in ev]
    table = timeline_table(sym, [e.time_range.elapsed_us() for e in ev], wt, pk)
    busy = sum(e.time_range.elapsed_us() for e in ev)
    span = ev[-1].time_range.end - ev[0].time_range.start
    return {"how": "torch.profiler (CUPTI) over graph replays run after the timed region; device time of every kernel of "
                   "the last replay, labelled by kernel symbol and launch order",
            "kernels_per_replay": per, "span_us": round(span, 1), "busy_us": round(busy, 1),
            "idle_between_kernels_us": round(span - busy, 1), "kernels": table[:12]}


def ops_mod():
    from stac_speech_translation_b200 import ops
    return ops


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of `kernel_key` at this workload from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by tools/ncu_extract.py), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(kernel_key, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import stac_speech_translation_b200 as sb
    from stac_speech_translation_b200 import ops, synth

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        # the gather's NCCL send/recv kernels hold whole SMs for the length of the transfer: persistent kernels must not
        # count on them (a CTA that waits for a held SM costs a full kernel duration)
        ops.check(ops.lib().stac_set_reserved_sms(max(0, args.reserve_sms)), "stac_set_reserved_sms")
    hp = sb.HParams.for_size(args.size)
    mods = sb.build_modules(hp, precision=args.precision, device=dev)
    wavs_cpu, wl_cpu = synth.fast_synth_batch(args.batch, args.seconds, seed=1234 + rank)
    # the synthetic audio is what a 16-bit wav file holds (sample / 32768): the device-resident batch, the oracle's input and
    # the int16 PCM of the end-to-end leg are the same signal
    pcm16_cpu = (wavs_cpu * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
    wavs_cpu = pcm16_cpu.float() / 32768.0
    if args.min_seconds is not None:
        # ragged batch (configs[3]): lengths uniform in [min, seconds]; right zero padding and wav_lens = len / Lmax as
        # the reference's PaddedBatch builds them.  The longest utterance keeps the full length.
        g_len = torch.Generator().manual_seed(4321 + rank)
        frac = args.min_seconds / args.seconds
        wl_cpu = frac + (1.0 - frac) * torch.rand(args.batch, generator=g_len)
        wl_cpu[0] = 1.0
        n_valid = torch.round(wl_cpu * wavs_cpu.shape[1]).long()
        for i in range(args.batch):
            wavs_cpu[i, int(n_valid[i]):] = 0.0
            pcm16_cpu[i, int(n_valid[i]):] = 0
        wl_cpu = (n_valid.double() / wavs_cpu.shape[1]).float()
    # normaliser statistics: one SpeechBrain-style statistics step on a calibration slice
    calib = wavs_cpu[: min(8, args.batch), : 16000 * 4].to(dev)
    mods["normalize"].calibrate(mods["compute_features"](calib), torch.ones(calib.shape[0], device=dev))
    # non-root ranks produce the posteriors directly in the dtype that travels to rank 0
    ship_bf16 = world > 1 and args.gather == "bf16" and args.precision == "bf16"   # wire dtype of the posteriors
    pipe = sb.EncoderPipeline(mods, posterior_dtype=torch.bfloat16 if (ship_bf16 and rank != 0) else torch.float32)
    pinned = wavs_cpu.pin_memory()
    wavs = pinned.to(dev, non_blocking=True)
    wl = wl_cpu.to(dev)
    n_samples = wavs.shape[1]
    audio_s = float(wl_cpu.double().sum()) * n_samples / 16000.0        # valid (unpadded) audio of this rank's batch
    padded_audio_s = args.batch * n_samples / 16000.0
    t2 = ops.frames_of(n_samples)[2]

    # ---- multi-GPU: results of every rank to rank 0 ----
    gather_bufs, peer = None, None
    if world > 1:
        d = hp.d_model
        pdt = torch.bfloat16 if (args.gather == "bf16" and args.precision == "bf16") else torch.float32
        if args.gather_transport in ("peer", "push"):
            from stac_speech_translation_b200.distributed import PeerGather, PeerGatherUnavailable, PushGather
            specs = {"enc_out": ((args.batch, t2, d), torch.float32), "greedy": ((args.batch, t2), torch.int32)}
            if args.gather != "ids":
                specs["p_ctc"] = ((args.batch, t2, VOCAB), pdt)
            try:
                peer = (PushGather if args.gather_transport == "push" else PeerGather)(specs, dev)
            except PeerGatherUnavailable as e:     # raised on every rank together: all fall back to NCCL p2p
                if rank == 0:
                    print(f"bench: peer transport unavailable ({e}); using NCCL point-to-point", file=sys.stderr)
                args.gather_transport = "nccl"
        if args.gather_transport == "nccl" and rank == 0:
            gather_bufs = {"enc": [torch.empty(args.batch, t2, d, device=dev) for _ in range(world)],
                           "ids": [torch.empty(args.batch, t2, device=dev, dtype=torch.int32) for _ in range(world)]}
            if args.gather != "ids":
                gather_bufs["p"] = [torch.empty(args.batch, t2, VOCAB, device=dev, dtype=pdt) for _ in range(world)]

    pending = []          # NCCL transport: gathers in flight (works, tensors kept alive); at most one step behind
    gstep = [0]           # steps issued so far (selects the result slot of the peer transport)

    def drain(keep=0):
        if peer is not None:
            if keep == 0:
                peer.finish(gstep[0] - 1)
            return
        while len(pending) > keep:
            works, _ = pending.pop(0)
            for w in works:
                w.wait()

    def gather_nccl(res):
        """enc_out, greedy ids (and posteriors) of this step to rank 0 over NCCL point-to-point, asynchronously:
        the transfer of step i overlaps the compute of step i+1 (the references keep the buffers alive); rank 0's
        own results stay where they are."""
        keys = ["enc", "ids"] + ([] if args.gather == "ids" else ["p"])
        if rank == 0:
            ops_ = [dist.P2POp(dist.irecv, gather_bufs[k][r], r) for r in range(1, world) for k in keys]
            keep = [res]
        else:
            payload = {"enc": res["enc_out"], "ids": res["greedy"]}
            if args.gather != "ids":
                p = res["p_ctc"]
                want = torch.bfloat16 if args.gather == "bf16" else torch.float32
                if p.dtype != want:           # only when the posteriors were not produced in the wire dtype
                    pb = torch.empty(p.shape, device=dev, dtype=want)
                    ops._call("stac_cast_bf16", ops.ptr(p), p.numel(), ops.ptr(pb), ops.stream())
                    p = pb
                payload["p"] = p
            ops_ = [dist.P2POp(dist.isend, payload[k], 0) for k in keys]
            keep = [res, payload]
        pending.append((dist.batch_isend_irecv(ops_), keep))
        drain(keep=1)

    def step(x, graphs=None):
        """One step of this rank: the path over one batch + shipping its results to rank 0.  `graphs`: per-slot CUDA
        graphs of the path (the end-to-end loop); eager launches otherwise."""
        i = gstep[0]
        gstep[0] += 1
        if peer is not None and rank != 0:
            peer.begin_write(i)               # rank 0 has pulled the step that used this slot last
            res = graphs[i % 2]() if graphs else pipe(x, wl, outputs=peer.slot(i))
            peer.end_write(i)
            return res
        res = graphs[i % 2]() if graphs else pipe(x, wl)
        if peer is not None:
            peer.collect(i)                   # copy-engine pull of every peer's step i, on a side stream
        elif world > 1:
            gather_nccl(res)
        return res

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step(wavs)
    barrier()

    # ---- untimed pre-pass: CUDA events around EVERY launch -> per-kernel table; names the dominant kernel ----
    kv = ops.kv_lengths(wl, args.batch, t2, dev, False).long()
    attn_pairs = int((kv * t2).sum())                # sum over utterances of T2 * S_valid
    wt = work_table(sb.MODEL_SIZES[args.size], args.batch, n_samples, attn_pairs)
    pk = peaks()
    ops.TRACE = []
    for _ in range(2):
        pipe(wavs, wl)
    torch.cuda.synchronize()
    trace, ops.TRACE = ops.TRACE, None
    table = kernel_table(trace, 2, wt, pk)
    top_key = table[0]["kernel"]
    barrier()

    # The public call is a CUDA graph of the path per input slot (sb.GraphedPipeline): the host's share of a step is
    # two copies and a replay, so the rate does not depend on how fast this box's CPU runs Python.
    copy_stream = torch.cuda.Stream()
    dev_in = [torch.empty_like(wavs) for _ in range(2)]
    for d_ in dev_in:
        d_.copy_(wavs)
    graphed = [sb.GraphedPipeline(pipe, dev_in[s], wl,
                                  **({"outputs": peer.slot(s)} if (peer is not None and rank != 0) else {}))
               for s in range(2)]
    barrier()

    # ---- timed region: device-resident inputs, K steps = K replays of the path's CUDA graph ----
    # (the product's public call; with eager launches the number measured this box's Python - the same kernels took
    # 4.7 ms per step on one box and 7.4 ms on another whose host was slower, and at 8 GPUs rank 0's launches plus the
    # gather's control messages were what the step waited for: 6.8 ms against 5.6 ms, round 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step(None, graphs=graphed)
        drain()                  # the compute stream waits for the last gather: it is inside the timed region
        e1.record()
        barrier()
        # same K steps again with eager launches and CUDA events around every launch of the dominant kernel (events
        # cannot sit inside a graph): the kernel's live duration, and the eager step time next to the graph's
        ops.TRACE, ops.TRACE_FILTER = [], {top_key}
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(args.steps):
            step(wavs)
        drain()
        x1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    eager_ms = max_over_ranks(x0.elapsed_time(x1) / args.steps)
    top_trace, ops.TRACE, ops.TRACE_FILTER = ops.TRACE, None, None
    launches = (len(trace) // 2) * args.steps            # kernels inside the K replayed graphs
    top_live = kernel_table(top_trace, args.steps, wt, pk)[0]

    # ---- end to end: pinned host PCM in (H2D inside the timed region), greedy ids out (D2H) ----
    ids_host = [torch.empty(args.batch, t2, dtype=torch.int32).pin_memory() for _ in range(2)]
    use_i16 = args.pcm == "int16"
    if use_i16:
        pinned_i16 = pcm16_cpu.pin_memory()
        dev_i16 = [torch.empty_like(pinned_i16, device=dev) for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_done = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream()

    def e2e_loop(n):
        k0 = gstep[0]                       # input / graph / result slot of a step = its global index & 1
        for i in range(n + 1):
            if i < n:                       # stage PCM of step i (overlaps compute of step i-1)
                s = (k0 + i) % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(in_free[s])
                    if use_i16:                # 16-bit PCM over the link, decoded on the device (stac_pcm_i16_to_f32)
                        dev_i16[s].copy_(pinned_i16, non_blocking=True)
                        ops.check(ops.lib().stac_pcm_i16_to_f32(ops.ptr(dev_i16[s], torch.int16), dev_i16[s].numel(),
                                                                ops.ptr(dev_in[s]), ops.stream()), "stac_pcm_i16_to_f32")
                    else:
                        dev_in[s].copy_(pinned, non_blocking=True)
                    in_ready[s].record(copy_stream)
            if i > 0:                       # compute step i-1, ship its greedy ids to the host
                s = (k0 + i - 1) % 2
                main.wait_event(in_ready[s])
                res = step(None, graphs=graphed)
                in_free[s].record(main)
                ids_host[s].copy_(res["greedy"], non_blocking=True)
                out_done[s].record(main)
            if i > 1:                       # the consumer reads step i-2's ids
                out_done[(k0 + i) % 2].synchronize()
                _ = int(ids_host[(k0 + i) % 2][0, 0])
        drain()
        torch.cuda.synchronize()

    for s in range(2):
        in_free[s].record(main)
    e2e_loop(2)
    # plain pinned H2D bandwidth of this box (explains e2e when the link, not the GPU, is the limit)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(3):
        dev_in[0].copy_(pinned, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * pinned.numel() * 4 / (h0.elapsed_time(h1) * 1e-3) / 1e9
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)

    # ---- multi-GPU: what rank 0 holds must be what the ranks produced (checked on the last step) ----
    gather_check = None
    if peer is not None:
        last = gstep[0] - 1
        barrier()
        mine = None
        if rank != 0:
            sl = peer.slot(last)
            mine = {k: float(v.double().sum()) for k, v in sl.items()}
        sums = [None] * world
        dist.all_gather_object(sums, mine)
        if rank == 0:
            for r in range(1, world):
                for k, want in sums[r].items():
                    got = float(peer.gathered_view(r, last)[k].double().sum())
                    if abs(got - want) > 1e-6 * max(1.0, abs(want)):
                        raise RuntimeError(f"peer gather mismatch: rank {r} {k}: {got} != {want}")
            gather_check = "checksums of every rank's last step match what rank 0 holds"

    if world > 1:                                  # valid audio differs per rank when the batch is ragged
        t_a = torch.tensor([audio_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t_a)
        total_audio = float(t_a)
    else:
        total_audio = audio_s
    if rank != 0:
        return

    ingress = None
    if world > 1:
        per_rank = args.batch * t2 * (hp.d_model * 4 + 4)
        if args.gather != "ids":
            per_rank += args.batch * t2 * VOCAB * (2 if (args.gather == "bf16" and args.precision == "bf16") else 4)
        ingress = {"bytes_per_step": int(per_rank * (world - 1)),
                   "achieved_gbs": round(per_rank * (world - 1) / (ms * 1e-3) / 1e9, 1),
                   "peer_copy_peak_gbs": 770.0,
                   "note": "what rank 0 receives per step over NVLink (enc_out fp32, greedy ids"
                           + ("" if args.gather == "ids" else f", {args.gather} posteriors") + ") / step time; the "
                           "measured peer-copy peak of this pool is 770 GB/s per direction (B200_PROFILING.md)"}
    step_sum = sum(r["ms_per_step"] for r in table)
    peak = pk["tflops_sustained"] if top_live["bound"] == "tensor" else pk["hbm_gbs"]
    roofline = {"kernel": top_key, "bound": top_live["bound"], "achieved": top_live["achieved"], "peak": peak,
                "unit": top_live["unit"], "frac": top_live["frac"], "traffic": ncu_traffic(top_key),
                "peak_source": pk["source"] + (" (sustained bf16 GEMM)" if top_live["bound"] == "tensor" else " (copy)"),
                "algorithmic_per_launch": top_live["work_per_launch"],
                "avg_launch_ms": top_live["avg_ms"], "launches_timed": len(top_trace),
                "share_of_step": round(table[0]["ms_per_step"] / step_sum, 4),
                "how": "CUDA events around every launch of this kernel in K eager steps run right behind the K timed graph "
                       "replays, inside the same clock-sampled region (events cannot be recorded inside a graph)",
                "eager_ms_per_step": round(eager_ms, 3)}
    if args.trace_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.trace_out)), exist_ok=True)
        json.dump({"table": table, "traced_step_ms": round(step_sum, 3)}, open(args.trace_out, "w"), indent=1)

    cpu, parity = None, None
    if not args.no_cpu_baseline:
        cpu, ref = cpu_baseline(args, mods, wavs_cpu, wl_cpu)
        res = pipe(wavs, wl)                       # the batch every timed step ran on, same kernels
        torch.cuda.synchronize()
        parity = parity_block(res, ref, ref["enc_out"].shape[0], args.precision, wl_cpu)
    timeline = None
    if world == 1 and args.precision == "bf16":
        try:
            timeline = graph_timeline(graphed[0], wt, pk)
        except Exception as exc:  # noqa: BLE001 - an optional breakdown must never cost the bench line
            timeline = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    line = {
        "metric": METRIC, "value": round(total_audio / (ms * 1e-3), 1), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {**workload_config(args, t2, world),
                   "audio": "value counts valid (unpadded) audio seconds; padded rate "
                            f"{padded_audio_s * world / (ms * 1e-3):.0f} audio-s/s"},
        "e2e": {"value": round(total_audio / (e2e_ms * 1e-3), 1), "unit": UNIT,
                "h2d_bytes_per_step": int(pinned.numel() * (2 if use_i16 else 4)),
                "d2h_bytes_per_step": int(args.batch * t2 * 4),
                "ms_per_step": round(e2e_ms, 3), "pinned_h2d_gbs": round(h2d_gbs, 1),
                "note": ("pinned 16-bit PCM in, decoded to fp32 on the device (sample / 32768)" if use_i16 else
                         "pinned fp32 PCM in") + " (double-buffered on a copy stream), one CUDA-graph replay of the path "
                        "(GraphedPipeline) per step, greedy CTC ids out; enc_out and p_ctc stay on the device as in "
                        "the reference's compute_forward"},
        "gpu_launches": launches,
        **({"gather_check": gather_check} if gather_check else {}),
        **({"rank0_ingress": ingress} if ingress else {}),
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "kernels": [{k: v for k, v in r.items() if k != "work_per_launch"} for r in table[:8]],
        **({"graph_timeline": timeline} if timeline else {}),
    }
    emit(line)
    if parity is not None and not parity["ok"]:
        print(f"bench: PARITY FAILURE against the oracle: {parity}", file=sys.stderr)
        sys.exit(3)


# --------------------------------------------------------------------------
# configs[2] / configs[4]: a corpus of length-bucketed ragged batches, whole batches sharded over the ranks
# --------------------------------------------------------------------------
def run_bucketed(args, rank, world, local_rank):
    """configs[2]: M model over `--utterances` utterances with LogNormal(median 8 s, sigma 0.7) durations clipped to
    [1, 30] s; configs[4]: front-end only (Fbank + normalise + ConvolutionFrontEnd) over 10 k utterances of 1-30 s.
    Batches are built the way the reference's DynamicBatchSampler set-up builds them (synth.bucket_batches,
    /root/reference/stac-st/dataio_and_utils.py:203-231) and WHOLE batches go to ranks longest-processing-time-first
    (distributed.plan): outputs depend on batch membership (wav_lens = len / Lmax, reflect padding at the batch edge), so
    a batch is never split.  A step = one pass over the corpus: every rank replays one CUDA graph per batch of its share,
    round-robin over `--streams` streams, then the ranks' results (encoder states fp32, greedy ids, bf16 posteriors) go
    to rank 0 with one NCCL send per tensor and rank.  value = valid audio seconds of the corpus / max-over-ranks time:
    strong scaling (the corpus is fixed, N ranks share it)."""
    import numpy as np
    import torch.distributed as dist
    import stac_speech_translation_b200 as sb
    from stac_speech_translation_b200 import distributed, ops, synth

    frontend_only = args.config == 4
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    hp = sb.HParams.for_size(args.size)
    mods = sb.build_modules(hp, precision=args.precision, device=dev)
    dur = (synth.uniform_durations(args.utterances, 1234) if frontend_only
           else synth.lognormal_durations(args.utterances, 1234))
    bucketed, per_rank = distributed.plan(dur, world, max_batch_len=args.max_batch_len)
    tmpl = synth.fast_synth_batch(8, 30.0, seed=1234)[0]
    calib = tmpl[:, : 16000 * 4].to(dev)
    mods["normalize"].calibrate(mods["compute_features"](calib), torch.ones(calib.shape[0], device=dev))
    tmpl_dev = tmpl.to(dev)
    d = hp.d_model
    my_ids = list(per_rank[rank])

    # ---- this rank's batches, PCM resident on the device (and a pinned host copy for the end-to-end leg) ----
    batches, frames_total = [], 0
    for bid in my_ids:
        idx = bucketed.batches[bid]
        n = [max(640, int(round(float(dur[i]) * 16000))) for i in idx]
        lmax = (max(n) + 31) // 32 * 32       # collation pads the batch to whole 32-sample rows (the Fbank kernel's PCM tiles)
        wav = torch.zeros(len(idx), lmax, device=dev)
        for k, (i, ni) in enumerate(zip(idx, n)):
            off = (i * 7919) % (tmpl_dev.shape[1] - ni + 1)
            wav[k, :ni] = tmpl_dev[i % 8, off:off + ni] * (0.5 + (i % 10) / 10.0)
        wl = torch.tensor([ni / lmax for ni in n], dtype=torch.float64).float().to(dev)
        t2 = ops.frames_of(lmax)[2]
        batches.append({"id": bid, "wavs": wav, "wl": wl, "n": len(idx), "lmax": lmax, "t2": t2, "row0": frames_total,
                        "valid_s": sum(n) / 16000.0, "padded_s": len(idx) * lmax / 16000.0})
        frames_total += len(idx) * t2
    valid_s = sum(b["valid_s"] for b in batches)
    padded_s = sum(b["padded_s"] for b in batches)

    # ---- result buffers of the rank: every batch's kernels write their slice; ONE transfer per tensor to rank 0 ----
    pdt = torch.float32 if world == 1 else torch.bfloat16
    flat = {}
    if not frontend_only:
        flat = {"enc_out": torch.empty(max(frames_total, 1), d, device=dev),
                "greedy": torch.empty(max(frames_total, 1), device=dev, dtype=torch.int32)}
        if args.gather != "ids":
            flat["p_ctc"] = torch.empty(max(frames_total, 1), VOCAB, device=dev, dtype=pdt)
    rank_frames = [0] * world
    if world > 1:
        t_f = torch.zeros(world, device=dev, dtype=torch.int64)
        t_f[rank] = frames_total
        dist.all_reduce(t_f)
        rank_frames = t_f.tolist()
    gathered = None
    if world > 1 and rank == 0 and not frontend_only:
        gathered = {r: {k: torch.empty((max(rank_frames[r], 1),) + tuple(v.shape[1:]), device=dev, dtype=v.dtype)
                        for k, v in flat.items()} for r in range(1, world)}

    pipe = sb.EncoderPipeline(mods, posterior_dtype=pdt)

    def outputs_of(b):
        if frontend_only:
            return {}
        r0, r1 = b["row0"], b["row0"] + b["n"] * b["t2"]
        out = {"enc_out": flat["enc_out"][r0:r1].view(b["n"], b["t2"], d),
               "greedy": flat["greedy"][r0:r1].view(b["n"], b["t2"])}
        if "p_ctc" in flat:
            out["p_ctc"] = flat["p_ctc"][r0:r1].view(b["n"], b["t2"], VOCAB)
        return out

    call_kw = {"stop_after": "cnn"} if frontend_only else {}

    # ---- untimed traced pass (eager): per-kernel table with the work of every batch shape ----
    pk = peaks()
    work_sum, launches_sum = {}, {}
    ops.TRACE = []
    for b in batches:
        kv = ops.kv_lengths(b["wl"].cpu(), b["n"], b["t2"], "cpu", False).long()      # (host twin: keeps the trace clean)
        wt = work_table(sb.MODEL_SIZES[args.size], b["n"], b["lmax"], int((kv * b["t2"]).sum()))
        for k, (bound, work, n_l) in wt.items():
            work_sum[k] = (bound, work_sum.get(k, (bound, 0))[1] + work * n_l)
        pipe(b["wavs"], b["wl"], outputs=outputs_of(b), **call_kw)
    torch.cuda.synchronize()
    trace, ops.TRACE = ops.TRACE, None
    per = {}
    for name, label, e0, e1 in trace:
        k = ops.trace_key(name, label)
        t_, n_ = per.get(k, (0.0, 0))
        per[k] = (t_ + e0.elapsed_time(e1), n_ + 1)
    table = []
    for k, (t_, n_) in sorted(per.items(), key=lambda kv_: -kv_[1][0]):
        bound, work = work_sum.get(k, ("hbm", 0))
        ach = work / (t_ * 1e-3) / (1e12 if bound == "tensor" else 1e9) if t_ > 0 else 0.0
        peak = pk["tflops_sustained"] if bound == "tensor" else pk["hbm_gbs"]
        table.append({"kernel": k, "launches_per_step": n_, "ms_per_step": round(t_, 4), "bound": bound,
                      "achieved": round(ach, 1), "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                      "frac": round(ach / peak, 4), "work_per_step": work})

    # ---- one CUDA graph per batch; graphs replayed on the same stream share that stream's memory pool ----
    n_streams = max(1, min(args.streams, max(1, len(batches))))
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    pools = [torch.cuda.graph_pool_handle() for _ in range(n_streams)]
    graphs = []
    for i, b in enumerate(batches):
        graphs.append(sb.GraphedPipeline(pipe, b["wavs"], b["wl"], warmup=1, pool=pools[i % n_streams],
                                         outputs=outputs_of(b), **call_kw))
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    done = [torch.cuda.Event() for _ in range(n_streams)]

    # ---- transport of the results to rank 0 ----
    # push (default): rank 0's per-rank flat buffers are peer-mapped into every producer (CUDA IPC, once); right behind
    # each batch's graph the producer's own copy engines push that batch's slices over NVLink, so the transfer overlaps
    # the batches that follow (one NCCL send per tensor at the end of the pass delivered ~75 GB/s per peer and cost the
    # 8-GPU pass 20 ms of its 49).  nccl: the end-of-pass sends.
    push = None
    if world > 1 and not frontend_only and args.gather_transport != "nccl":
        import ctypes
        from torch.multiprocessing.reductions import reduce_tensor
        from stac_speech_translation_b200.distributed import _IpcEvent
        ctl = dist.new_group(backend="gloo")
        ok = True
        handles = [None] * world
        try:
            if rank == 0:
                for r in range(1, world):
                    handles[r] = {"device": dev.index, "tensors": {k: reduce_tensor(v) for k, v in gathered[r].items()}}
        except Exception as e:                       # noqa: BLE001
            ok = False
        mine = [None]
        dist.scatter_object_list(mine, handles if rank == 0 else None, src=0, group=ctl)
        remote = None
        try:
            if rank != 0:
                root_dev = mine[0]["device"]
                if not torch.cuda.can_device_access_peer(dev.index, root_dev):
                    raise RuntimeError("no peer access")
                remote = {k: fn(*a) for k, (fn, a) in mine[0]["tensors"].items()}
                flat["greedy"][:1].copy_(remote["greedy"][:1])      # makes torch enable peer access dev -> root
                torch.cuda.synchronize()
        except Exception as e:                       # noqa: BLE001
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=ctl)
        if int(flag) == 1:
            push = {"rt": _IpcEvent.rt(), "remote": remote, "root": mine[0]["device"] if rank != 0 else dev.index,
                    "streams": [torch.cuda.Stream(device=dev) for _ in range(n_streams)],
                    "computed": [torch.cuda.Event() for _ in batches]}
        elif rank == 0:
            print("bench: peer transport unavailable; using NCCL sends at the end of the pass", file=sys.stderr)

    def push_batch(i, st):
        """Producer: enqueue the push of batch i's result slices behind its graph (stream st)."""
        b = batches[i]
        r0, r1 = b["row0"], b["row0"] + b["n"] * b["t2"]
        ps = push["streams"][i % n_streams]
        push["computed"][i].record(st)
        ps.wait_event(push["computed"][i])
        for k in flat:
            src, dst = flat[k][r0:r1], push["remote"][k][r0:r1]
            rc = push["rt"].cudaMemcpyPeerAsync(ctypes.c_void_p(dst.data_ptr()), push["root"], ctypes.c_void_p(src.data_ptr()),
                                               dev.index, src.numel() * src.element_size(), ctypes.c_void_p(ps.cuda_stream))
            if rc != 0:
                raise RuntimeError(f"cudaMemcpyPeerAsync failed with cudaError {rc}")

    def gather():
        if world == 1 or frontend_only:
            return
        if push is not None:
            if rank != 0:                                  # the pass is over when its pushes are
                for ps in push["streams"]:
                    ev = torch.cuda.Event()
                    ev.record(ps)
                    main.wait_event(ev)
            return
        if rank == 0:
            ops_ = [dist.P2POp(dist.irecv, gathered[r][k], r) for r in range(1, world) for k in flat if rank_frames[r]]
        else:
            ops_ = [dist.P2POp(dist.isend, flat[k], 0) for k in flat] if frames_total else []
        for wk in (dist.batch_isend_irecv(ops_) if ops_ else []):
            wk.wait()

    def step(h2d=None):
        start = torch.cuda.Event()
        start.record(main)
        for i, g in enumerate(graphs):
            st = streams[i % n_streams]
            if i < n_streams:
                st.wait_event(start)
            with torch.cuda.stream(st):
                if h2d is not None:
                    g.wavs.copy_(h2d[i], non_blocking=True)
                g.graph.replay()
            if push is not None and rank != 0:
                push_batch(i, st)
        for k, st in enumerate(streams):
            done[k].record(st)
            main.wait_event(done[k])
        gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = sum(r["launches_per_step"] for r in table) * args.steps

    # ---- end to end: pinned host PCM of every batch in, greedy ids (front-end sweep: nothing) out, wall clock ----
    pinned = [b["wavs"].cpu().pin_memory() for b in batches]
    ids_host = None if frontend_only else torch.empty(max(sum(rank_frames) if world > 1 else frames_total, 1),
                                                       dtype=torch.int32).pin_memory()

    def e2e_step():
        step(h2d=pinned)
        if ids_host is not None and rank == 0:
            ids_host[:frames_total].copy_(flat["greedy"][:frames_total], non_blocking=True)
            off = frames_total
            for r in range(1, world):
                ids_host[off:off + rank_frames[r]].copy_(gathered[r]["greedy"][:rank_frames[r]], non_blocking=True)
                off += rank_frames[r]
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)

    tot = torch.tensor([valid_s, padded_s, float(len(batches))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    # shard/gather check: what rank 0 holds of every other rank equals what that rank produced (checksums)
    gather_check = None
    if world > 1 and not frontend_only:
        mine = {k: float(v[:frames_total].double().sum()) for k, v in flat.items()} if rank else None
        sums = [None] * world
        dist.all_gather_object(sums, mine)
        if rank == 0:
            for r in range(1, world):
                for k, want in sums[r].items():
                    got = float(gathered[r][k][:rank_frames[r]].double().sum())
                    if abs(got - want) > 1e-6 * max(1.0, abs(want)):
                        raise RuntimeError(f"gather mismatch: rank {r} {k}: {got} != {want}")
            gather_check = "checksums of every rank's results match what rank 0 received"
    if rank != 0:
        return
    total_valid, total_padded, n_batches = [float(x) for x in tot]

    # ---- parity on a bounded sample: the oracle (product's weights) on this rank's shortest batch ----
    cpu, parity = None, None
    if not args.no_cpu_baseline and batches:
        b = min(batches, key=lambda x: x["padded_s"])
        keep = min(b["n"], args.cpu_batch)
        sub = dict(args=args)
        import oracle
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        omods = oracle.build_reference_modules(args.size)
        for k in ("CNN", "Transformer", "ctc_lin"):
            omods[k].load_state_dict({k_: v.detach().float().cpu() for k_, v in mods[k].state_dict().items()}, strict=True)
        omods["normalize"]._load_statistics_dict({k_: (v.detach().float().cpu() if torch.is_tensor(v) else v)
                                                  for k_, v in mods["normalize"]._statistics_dict().items()})
        omods["normalize"].eval()
        w_cpu, wl_c = b["wavs"].cpu(), b["wl"].cpu()
        t0 = time.perf_counter()
        ref = oracle.reference_compute_forward(omods, w_cpu, wl_c, stages="frontend" if frontend_only else None)
        dt = time.perf_counter() - t0
        cpu = {"value": round(b["valid_s"] / dt, 2), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"one pass over the shortest batch of the corpus ({b['n']} utterances, {b['padded_s']:.0f} s "
                         f"padded), torch fp32, {cores} threads"}
        res = pipe(b["wavs"], b["wl"], **call_kw)
        torch.cuda.synchronize()
        if frontend_only:
            a_, r_ = res["cnn"].float().cpu().reshape(ref["cnn"].shape[0], ref["cnn"].shape[1], -1), \
                ref["cnn"].reshape(ref["cnn"].shape[0], ref["cnn"].shape[1], -1)
            err = float((a_ - r_).double().norm() / r_.double().norm())
            parity = {"utterances": b["n"], "cnn_rel_l2": round(err, 6), "tolerance": PARITY_TOL[args.precision],
                      "ok": bool(err <= PARITY_TOL[args.precision])}
        else:
            parity = parity_block(res, ref, b["n"], args.precision, wl_c)

    top = table[0]
    peak = pk["tflops_sustained"] if top["bound"] == "tensor" else pk["hbm_gbs"]
    roofline = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": peak,
                "unit": top["unit"], "frac": top["frac"], "traffic": None,
                "peak_source": pk["source"] + (" (sustained bf16 GEMM)" if top["bound"] == "tensor" else " (copy)"),
                "algorithmic_per_step": top["work_per_step"], "ms_per_step": top["ms_per_step"],
                "how": "CUDA events around every launch of an untimed eager pass over rank 0's batches: sum of the "
                       "algorithmic work of every batch shape / sum of the launch durations"}
    cfg = {"workload": (f"configs[{args.config}]: " + (
               f"front-end only (Fbank + InputNormalization + ConvolutionFrontEnd), {args.utterances} utterances "
               f"U[1, 30] s" if frontend_only else
               f"STAC-ST {args.size} encoder + CTC head, {args.utterances} utterances LogNormal(median 8 s, sigma 0.7) "
               f"clipped to [1, 30] s")
               + f", length-bucketed batches of <= {args.max_batch_len:g} s (DynamicBatchSampler rule), whole batches "
                 f"sharded over {world} rank(s) longest-processing-time-first"),
           "precision": args.precision, "batches": int(n_batches), "streams": n_streams,
           "valid_audio_s": round(total_valid, 1), "padded_audio_s": round(total_padded, 1),
           "l2": "every batch has its own buffers (corpus PCM resident: no reuse between steps of the same data in L2 "
                 "beyond what a real pass over a corpus has)",
           "collation": "batches zero-padded to a multiple of 32 samples (2 ms)",
           "multi_gpu": "single GPU" if world == 1 else
                        ("enc_out fp32 + greedy ids" + ("" if args.gather == "ids" else " + bf16 posteriors")
                         + " of every rank to rank 0, "
                         + ("pushed batch by batch by the producers' copy engines into rank 0's peer-mapped buffers"
                            if push is not None else "one NCCL send per tensor and rank at the end of the pass")
                         + ", inside the timed region")}
    line = {
        "metric": METRIC, "value": round(total_valid / (ms * 1e-3), 1), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": cfg,
        "e2e": {"value": round(total_valid / (e2e_ms * 1e-3), 1), "unit": UNIT,
                "h2d_bytes_per_step": int(sum(p.numel() for p in pinned) * 4),
                "d2h_bytes_per_step": 0 if ids_host is None else int(ids_host.numel() * 4),
                "ms_per_step": round(e2e_ms, 3),
                "note": "pinned fp32 PCM of every batch copied in on the batch's stream before its graph replay; greedy "
                        "ids of the whole corpus read back by rank 0; wall clock, max over ranks"},
        "gpu_launches": launches,
        **({"gather_check": gather_check} if gather_check else {}),
        "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "kernels": [{k: v for k, v in r.items() if k != "work_per_step"} for r in table[:10]],
    }
    if frontend_only:
        def agg(keys):
            t_ = sum(per[k][0] for k in keys if k in per)
            w_ = sum(work_sum[k][1] for k in keys if k in per)
            return t_, w_
        t_f, w_f = agg(["stac_fbank_logmel_tc2", "stac_fbank_logmel_tc", "stac_fbank_logmel", "stac_fbank_topdb_norm"])
        t_c, w_c = agg(["stac_conv1_bf16"])
        line["frontend_fractions"] = {
            "fbank_norm_hbm_frac": round(w_f / (t_f * 1e-3) / 1e9 / pk["hbm_gbs"], 4) if t_f else None,
            "conv1_tensor_frac": round(w_c / (t_c * 1e-3) / 1e12 / pk["tflops_sustained"], 4) if t_c else None,
            "note": "a2-a3 are HBM-bound (algorithmic bytes 4 L + 4*80 T per utterance), a4's conv1 is tensor-bound "
                    "(SURVEY.md 8d): reported separately"}
    emit(line)
    if parity is not None and not parity["ok"]:
        print(f"bench: PARITY FAILURE against the oracle: {parity}", file=sys.stderr)
        sys.exit(3)


_RESULT_FD = None


def emit(line):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    import faulthandler
    global _RESULT_FD
    faulthandler.enable()
    args = parse_args()
    # stdout carries exactly one line: libraries that print there from C (NCCL's version banner on the first
    # communicator) are sent to stderr for the whole run, the result goes to a duplicate of the original descriptor
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.config in (2, 4):
            run_bucketed(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
