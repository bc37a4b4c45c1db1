"""Host-side logic and the C-ABI boundary, without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import oracle
import stac_speech_translation_b200 as sb
from stac_speech_translation_b200 import _lib, ops, synth
from util import TINY, oracle_modules

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "stac_b200.h")).read()
    declared = set(re.findall(r"\b(stac_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.exported_symbols())                 # ctypes table mirrors the header
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(handle, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (stac_\w+)", nm))
    assert exported == declared
    assert _lib.lib().stac_version() == 100
    assert b"invalid argument" in _lib.lib().stac_error_string(-1)
    assert _lib.lib().stac_fbank_tables_floats() == 2692
    assert _lib.lib().stac_conv0_padded_elems(2, 501) == 2 * 4 * 252 * 21 * 256


def test_ctypes_signatures_mirror_the_header():
    """Every prototype of include/stac_b200.h, parameter by parameter, against the ctypes table the Python side binds
    with (a mismatch would corrupt arguments silently on the GPU box)."""
    from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
    header = open(os.path.join(ROOT, "include", "stac_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"\b(int64_t|int|const char\s*\*)\s+(stac_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header)
    assert len(protos) == len(_lib.exported_symbols())

    def ctype_of(param):
        param = param.strip()
        if "*" in param:
            return c_void_p
        base = param.split()[0] if not param.startswith("const") else param.split()[1]
        return {"int64_t": c_int64, "int32_t": c_int32, "int": c_int, "float": c_float}[base]

    for ret, name, params in protos:
        res, args = _lib._SIGNATURES[name]
        want = [] if params.strip() in ("", "void") else [ctype_of(p) for p in params.split(",")]
        assert args == want, (name, args, want)
        assert res == {"int": c_int, "int64_t": c_int64}.get(ret, c_char_p), name


def test_state_dict_layout_matches_speechbrain_keys():
    mods = sb.build_modules(sb.HParams(**TINY, output_neurons=64), "bf16", device="cpu")
    model = torch.nn.ModuleList([mods["CNN"], mods["Transformer"], mods["ctc_lin"], mods["ctc_lin"]])
    keys = set(model.state_dict())
    for k in ["0.convblock_0.convs.conv_0.conv.weight", "0.convblock_1.convs.conv_0.conv.bias",
              "0.convblock_0.convs.norm_0.norm.weight", "0.convblock_1.convs.norm_0.norm.bias",
              "1.custom_src_module.layers.0.w.weight", "1.custom_src_module.layers.0.w.bias",
              "1.positional_encoding.pe", "1.encoder.layers.0.self_att.att.in_proj_weight",
              "1.encoder.layers.1.self_att.att.out_proj.bias", "1.encoder.layers.0.pos_ffn.ffn.0.weight",
              "1.encoder.layers.0.pos_ffn.ffn.3.bias", "1.encoder.layers.1.norm1.norm.weight",
              "1.encoder.layers.1.norm2.norm.bias", "1.encoder.norm.norm.weight", "2.w.weight", "3.w.bias"]:
        assert k in keys, k
    sd = model.state_dict()
    assert sd["0.convblock_0.convs.conv_0.conv.weight"].shape == (256, 1, 3, 3)
    assert sd["0.convblock_1.convs.conv_0.conv.weight"].shape == (256, 256, 3, 3)
    assert sd["0.convblock_0.convs.norm_0.norm.weight"].shape == (40, 256)
    assert sd["0.convblock_1.convs.norm_0.norm.weight"].shape == (20, 256)
    assert sd["1.custom_src_module.layers.0.w.weight"].shape == (128, 5120)
    assert sd["1.positional_encoding.pe"].shape == (1, 2500, 128)
    # an oracle (SpeechBrain-layout) checkpoint loads unchanged, and vice versa
    omods = oracle_modules(TINY, vocab=64)
    for k in ("CNN", "Transformer", "ctc_lin"):
        r = mods[k].load_state_dict(omods[k].state_dict(), strict=True)
        assert not r.missing_keys and not r.unexpected_keys
        omods[k].load_state_dict(mods[k].state_dict(), strict=True)
    assert torch.equal(mods["Transformer"].positional_encoding.pe, omods["Transformer"].positional_encoding.pe)


def test_weight_packing_follows_checkpoint_updates():
    mods = sb.build_modules(sb.HParams(**TINY, output_neurons=64), "fp32", device="cpu")
    tr = mods["Transformer"]
    p1 = tr.packed()
    assert tr.packed() is p1                                        # cached while parameters are untouched
    d = 128
    att = tr.encoder.layers[0].self_att.att
    assert torch.allclose(p1.layers[0].w_qkv[:d], att.in_proj_weight[:d] / 8)      # 1/sqrt(64) folded into W_q
    assert torch.equal(p1.layers[0].w_qkv[d:], att.in_proj_weight[d:].detach())
    w_src_before = p1.w_src.clone()      # fp32 packs alias the parameters, so keep a copy
    sd = {k: v * 0.5 if v.is_floating_point() else v for k, v in tr.state_dict().items()}
    tr.load_state_dict(sd)                                          # e.g. checkpoint averaging, inference.py:228-233
    p2 = tr.packed()
    assert p2 is not p1 and torch.allclose(p2.w_src, w_src_before * 0.5)
    cnn = mods["CNN"]
    w = cnn.packed()
    assert w.w1.shape == (256, 9, 256)
    c1 = cnn.convblock_1.convs.conv_0.conv.weight
    assert torch.equal(w.w1[5, 2 * 3 + 1, 77], c1[5, 77, 2, 1].detach())           # tap = kf*3 + kt
    mods16 = sb.build_modules(sb.HParams(**TINY, output_neurons=64), "bf16", device="cpu")
    wb = mods16["CNN"].packed()
    assert wb.w1.shape == (9, 256, 256) and wb.w1.dtype == torch.bfloat16


def test_no_cpu_fallback_and_clear_errors():
    mods = sb.build_modules(sb.HParams(**TINY, output_neurons=64), "bf16", device="cpu")
    with pytest.raises(sb.StacB200Error, match="no CPU fallback"):
        mods["compute_features"](torch.zeros(1, 16000))
    with pytest.raises(sb.StacB200Error, match="no CPU fallback"):
        mods["log_softmax"](torch.zeros(2, 3, 64))
    mods["Transformer"].train()
    with pytest.raises(sb.StacB200Error, match="inference-only"):
        mods["Transformer"].encode(torch.zeros(1, 4, 5120))
    mods["normalize"].train()            # train-mode statistics are a device path as well
    with pytest.raises(sb.StacB200Error, match="no CPU fallback"):
        mods["normalize"](torch.zeros(1, 4, 80), torch.ones(1))
    mods["normalize"].eval()
    with pytest.raises(sb.StacB200Error, match="no statistics"):
        mods["normalize"](torch.zeros(1, 4, 80), torch.ones(1))
    with pytest.raises(sb.StacB200Error):
        sb.Fbank(n_mels=40)
    with pytest.raises(sb.StacB200Error):
        sb.TransformerMultiTask(5000, 5120, d_model=256, nhead=4, normalize_before=False, activation=torch.nn.GELU)
    # the built-in decoder path is a device path too
    with pytest.raises(sb.StacB200Error, match="no CPU fallback"):
        mods["Transformer"].eval().decode(torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 4, 128))
    no_dec = sb.TransformerMultiTask(64, 5120, d_model=128, nhead=2, num_encoder_layers=1, num_decoder_layers=0,
                                     d_ffn=256, normalize_before=True, activation=torch.nn.GELU).eval()
    with pytest.raises(sb.StacB200Error, match="decoder"):
        no_dec.decode(torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 4, 128))


def test_kv_lengths_match_reference_masks():
    for t2 in (26, 251, 751, 1501):
        wl = torch.rand(64, generator=torch.Generator().manual_seed(t2)).clamp_min(0.05)
        wl[0] = 1.0
        enc = ops.kv_lengths(wl, 64, t2, "cpu", train_mask=False)
        trn = ops.kv_lengths(wl, 64, t2, "cpu", train_mask=True)
        keep_e = ~(torch.arange(t2)[None, :].to(wl) > torch.floor(wl * t2)[:, None])          # encode() :289-294
        keep_t = oracle.speechbrain_path.length_to_mask(torch.round(wl * t2), max_len=t2)     # make_masks :225-226
        assert torch.equal(enc.long(), keep_e.sum(1))
        assert torch.equal(trn.long(), keep_t.sum(1).clamp_min(1))
    assert ops.kv_lengths(None, 3, 10, "cpu", False).tolist() == [10, 10, 10]


def test_normalizer_checkpoint_roundtrip(tmp_path):
    omods = oracle_modules(TINY, vocab=64)
    n = sb.InputNormalization(norm_type="global", update_until_epoch=4)
    n._load_statistics_dict(omods["normalize"]._statistics_dict())
    n._save(tmp_path / "normalizer.ckpt")
    m = sb.InputNormalization(norm_type="global", update_until_epoch=4)
    m._load(tmp_path / "normalizer.ckpt")
    assert torch.equal(m.glob_mean, omods["normalize"].glob_mean) and m.count == omods["normalize"].count
    # calibrate() is SpeechBrain's first train-mode statistics step
    wavs, wl = synth.synth_batch([1.0, 0.6], seed=5)
    feats = omods["compute_features"](wavs)
    ref = oracle.InputNormalization(norm_type="global")
    ref.train(); ref(feats.clone(), wl)
    c = sb.InputNormalization(norm_type="global")
    c.calibrate(feats, wl)
    assert torch.allclose(c.glob_mean, ref.glob_mean) and torch.allclose(c.glob_std, ref.glob_std)


def test_synthetic_audio_and_bucketing():
    wavs, wl = synth.synth_batch([2.0, 1.0], seed=1)
    wavs2, _ = synth.synth_batch([2.0, 1.0], seed=1)
    assert torch.equal(wavs, wavs2) and wavs.shape == (2, 32000)
    assert wl.tolist() == [1.0, 0.5] and float(wavs[1, 16000:].abs().max()) == 0.0
    assert float(wavs.abs().max()) <= 1.0
    d = synth.lognormal_durations(4096, seed=0)
    assert d.min() >= 1.0 and d.max() <= 30.0
    bk = synth.bucket_batches(d, max_batch_len=200.0, num_buckets=50, max_batch_ex=128)
    seen = sorted(i for b in bk.batches for i in b)
    assert seen == list(range(4096))                                 # every utterance exactly once
    for b in bk.batches:
        assert len(b) <= 128 and len(b) * d[b].max() <= 200.0 + 30.0
    for world in (2, 4, 8):
        shards = synth.shard_batches(bk, world)
        assert sorted(i for s in shards for i in s) == list(range(len(bk.batches)))   # whole batches, once each
        loads = [sum(synth.batch_cost(d, bk.batches[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.05             # LPT keeps ranks within 5 %


def test_ctc_greedy_collapse():
    ids = torch.tensor([[0, 7, 7, 0, 7, 8, 8, 0], [5, 5, 5, 0, 0, 3, 9, 9]])
    assert sb.ctc_greedy_collapse(ids, [8, 6]) == [[7, 7, 8], [5, 3]]


def test_fbank_tables_match_oracle_filterbank():
    # build_fbank_tables needs the library only for its size check
    tab = ops.build_fbank_tables("cpu")
    fb = oracle.speechbrain_path.Filterbank(n_mels=80).fbank_matrix()      # [201, 80]
    assert torch.equal(tab[:400], torch.hamming_window(400))
    start, count, w = tab[1252:1332].long(), tab[1332:1412].long(), tab[1412:].view(80, 16)
    dense = torch.zeros(201, 80)
    for m in range(80):
        dense[start[m]:start[m] + count[m], m] = w[m, :count[m]]
    assert torch.equal(dense, fb)                                    # bit-identical sparse copy of the matrix


def test_turn_detection_host_side():
    """turns.rttm_line formats a spike exactly as the reference's append_speaker_turns does (golden lines produced by the
    reference function itself), and the device entry points refuse CPU tensors (no CPU fallback)."""
    import json
    import os
    from stac_speech_translation_b200 import turns
    from stac_speech_translation_b200._lib import StacB200Error
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "turns_reference.json")))
    for c in cases:
        ids = np.asarray(c["ids"])
        t2 = ids.shape[1]
        flat = ids.reshape(-1)
        assert [turns.rttm_line(c["utt"][f // t2], f % t2) for f in np.nonzero(flat == 7)[0]] == c["turn_rttm"]
        assert [turns.rttm_line(c["utt"][f // t2], f % t2) for f in np.nonzero(flat == 8)[0]] == c["xt_rttm"]
    with pytest.raises(StacB200Error):
        turns.append_speaker_turns(["a-b-0000100-c"], torch.zeros(1, 4, dtype=torch.int32), 7, 8, [], [])
    with pytest.raises(StacB200Error):
        turns.greedy_ids(torch.zeros(1, 4, 9))


def test_ingest_collation_matches_padded_batch_semantics():
    """ingest.collate against SpeechBrain's PaddedBatch behaviour for ``sig`` (right zero padding, wav_lens = len / max
    as a double division stored in fp32) and the int16 -> fp32 decode rule (sample / 32768, exact)."""
    from stac_speech_translation_b200 import ingest
    from stac_speech_translation_b200._lib import StacB200Error
    rng = np.random.default_rng(3)
    turns = [rng.integers(-32768, 32768, n, dtype=np.int16) for n in (1600, 37, 4801)]
    utt = ingest.concat_turns(turns)
    assert utt.dtype == np.int16 and utt.shape == (1600 + 37 + 4801,) and np.array_equal(utt[1600:1637], turns[1])
    batch = [utt, turns[0], turns[2][:4799]]
    pcm, wl = ingest.collate(batch)
    assert pcm.dtype == torch.int16 and pcm.shape == (3, utt.shape[0]) and wl.dtype == torch.float32
    for i, u in enumerate(batch):
        assert np.array_equal(pcm[i, :len(u)].numpy(), u) and not pcm[i, len(u):].any()
        assert float(wl[i]) == float(np.float32(len(u) / len(utt)))
    flat = torch.full((3 * utt.shape[0] + 5,), 9, dtype=torch.int16)
    pcm2, wl2 = ingest.collate(batch, out=flat)
    assert torch.equal(pcm2, pcm) and torch.equal(wl2, wl) and pcm2.data_ptr() == flat.data_ptr()
    # the decode rule the device kernel applies is exact in fp32 and equals what a 16-bit file loads as
    every = np.arange(-32768, 32768, dtype=np.int16)
    assert np.array_equal((every.astype(np.float32) * np.float32(1 / 32768)).astype(np.float64), every / 32768.0)
    with pytest.raises(StacB200Error):
        ingest.collate([])
    with pytest.raises(StacB200Error):
        ingest.collate([np.zeros(4, np.float32)])
    with pytest.raises(StacB200Error, match="no CPU fallback"):
        ingest.pcm_to_float(torch.zeros(16, dtype=torch.int16))


def test_attention_v2_protocol_model_check():
    """tools/model_check_mha2.py: the barrier protocol of csrc/attention_tc2.cu under random schedules (no deadlock, no
    parity aliasing, no data hazard), and the checker itself catches the race DESIGN.md section 4 describes (a consumer
    that jumps over uses of a parity-tracked mbarrier)."""
    import importlib.util
    path = os.path.join(ROOT, "tools", "model_check_mha2.py")
    spec = importlib.util.spec_from_file_location("model_check_mha2", path)
    mc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mc)
    assert mc.check(runs=80, seed=11) == 80
    assert mc.check(runs=60, seed=12, kv_stages=3) == 60            # the shallowest ring the kernel allows
    src = open(path).read()
    walk = src[src.index("                for _ in range(pc.n_kt):"):src.index("                sc = copy(pc)")]
    jump = ("                for _ in range(pc.n_kt):\n                    self.kv_empty[pc.stage].arrive()\n"
            "                    advance(pc)\n                    yield\n")
    ns = {"__name__": "mutant"}
    exec(compile(src.replace(walk, jump, 1), "mutant", "exec"), ns)
    with pytest.raises(AssertionError):
        ns["check"](200, 7)
    # scores running THREE tiles ahead of P.V (one more than there are buffers) must be caught as well
    ns = {"__name__": "mutant"}
    exec(compile(src.replace("g_s < g_p + 2", "g_s < g_p + 3", 1), "mutant", "exec"), ns)
    with pytest.raises(AssertionError):
        ns["check"](200, 7)


def test_default_attention_protocol_model_check():
    """tools/model_check_mha1.py on the default attention kernel's protocol.  With one o_staged barrier per (Q buffer,
    group) - what the kernel has - no random schedule deadlocks or aliases; with one per group (the round-1 build) a
    group that runs two short items ahead of the store warp does - the finding recorded in DESIGN.md section 9."""
    import importlib.util
    sys_path = os.path.join(ROOT, "tools")
    import sys
    sys.path.insert(0, sys_path)
    try:
        spec = importlib.util.spec_from_file_location("model_check_mha1", os.path.join(sys_path, "model_check_mha1.py"))
        mc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mc)
        assert mc.check(runs=150, seed=5, per_buffer=True) == 150
        with pytest.raises(AssertionError, match="deadlock|meant completion"):
            mc.check(runs=300, seed=3, per_buffer=False)
    finally:
        sys.path.remove(sys_path)


def test_fused_ffn_protocol_model_check():
    """tools/model_check_ffn.py: the fused feed-forward kernel as built (one issuer warp) has no deadlock / aliasing /
    hazard under random schedules with heavy-tailed TMA latencies, and the checker reproduces the race of the earlier
    two-issuer design that only ever fired on an 8-GPU run."""
    import importlib.util
    import sys
    tools = os.path.join(ROOT, "tools")
    sys.path.insert(0, tools)
    try:
        spec = importlib.util.spec_from_file_location("model_check_ffn", os.path.join(tools, "model_check_ffn.py"))
        mc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mc)
        assert mc.check(runs=100, seed=2) == 100
        with pytest.raises(AssertionError, match="meant completion|holds unit"):
            mc.check(runs=400, seed=1, two_issuers=True)
    finally:
        sys.path.remove(tools)


def test_custom_op_layer_registers_and_propagates_shapes():
    """custom_ops.py: the tensor-in / tensor-out entry points as torch.ops.stac_b200.* with fake-tensor shape functions
    (what torch.compile / export need); still no CPU implementation."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    import stac_speech_translation_b200.custom_ops  # noqa: F401
    ns = torch.ops.stac_b200
    with FakeTensorMode():
        f = ns.fbank(torch.empty(4, 16000), 80.0, True)
        assert f.shape == (4, 101, 80) and f.dtype == torch.float32
        assert ns.input_norm(f, torch.empty(80), torch.empty(80)).shape == f.shape
        y = ns.linear(torch.empty(4, 26, 256), torch.empty(5000, 256), None, "fp32")
        assert y.shape == (4, 26, 5000) and ns.log_softmax(y).shape == y.shape
        out, ids = ns.log_softmax_greedy(y)
        assert out.shape == y.shape and ids.shape == (4, 26) and ids.dtype == torch.int32
        assert ns.argmax_rows(out).dtype == torch.int32
        assert ns.pcm_to_float(torch.empty(3, 100, dtype=torch.int16)).dtype == torch.float32
        src = ns.conv_frontend(f, [torch.empty(1)] * 8, "bf16", True)
        assert src.shape == (4, 26, 5120) and src.dtype == torch.bfloat16
        enc = ns.encoder(src, torch.empty(4, dtype=torch.int32), [torch.empty(1)] * 29, 256, 4, "bf16")
        assert enc.shape == (4, 26, 256) and enc.dtype == torch.float32
        p, ids = ns.ctc_head(enc.to(torch.bfloat16), torch.empty(5000, 256, dtype=torch.bfloat16), None)
        assert p.shape == (4, 26, 5000) and ids.shape == (4, 26)
    with pytest.raises(sb.StacB200Error, match="no CPU fallback"):
        ns.input_norm(torch.zeros(1, 4, 80), torch.zeros(80), torch.ones(80))


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): one JSON line on stdout with the contract's
    keys, its own cpu_baseline description and an e2e object without device copies."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-batch", "1", "--seconds", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("encoder audio-sec/sec") and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_custom_ops_survive_torch_export():
    """What the custom-op layer is for: a module written over torch.ops.stac_b200.* exports to one graph whose nodes are
    those ops (fake tensors, no kernel runs)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    import stac_speech_translation_b200.custom_ops  # noqa: F401

    class M(torch.nn.Module):
        def forward(self, feats, mean, std, w, b):
            x = torch.ops.stac_b200.input_norm(feats, mean, std)
            y = torch.ops.stac_b200.linear(x, w, b, "fp32")
            return torch.ops.stac_b200.log_softmax_greedy(y)

    with FakeTensorMode(allow_non_fake_inputs=True):
        args = (torch.empty(2, 50, 80), torch.empty(80), torch.empty(80), torch.empty(64, 80), torch.empty(64))
        ep = torch.export.export(M(), args)
    targets = [str(n.target) for n in ep.graph.nodes if n.op == "call_function"]
    assert [t for t in targets if t.startswith("stac_b200.")] == [
        "stac_b200.input_norm.default", "stac_b200.linear.default", "stac_b200.log_softmax_greedy.default"]


def test_bench_configs_and_parity_block():
    """bench.py: --config N selects BASELINE.json configs[N] (model size, batch shape, raggedness), and the in-run parity
    block compares only the frames encode() keeps and fails above the north-star tolerance."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    old = sys.argv
    try:
        want = {1: ("S", 64, 30.0, None, None), 2: ("M", None, 30.0, None, 4096), 3: ("L", 16, 60.0, 45.0, None),
                4: ("S", None, 30.0, None, 10000)}
        for cfg, (size, batch, secs, min_s, n_utt) in want.items():
            sys.argv = ["bench.py", "--config", str(cfg)]
            a = bench.parse_args()
            assert (a.size, a.batch, a.seconds, a.min_seconds, a.utterances) == (size, batch, secs, min_s, n_utt)
    finally:
        sys.argv = old
    g = torch.Generator().manual_seed(0)
    p = torch.log_softmax(torch.randn(2, 10, 50, generator=g) * 4, -1)
    enc = torch.randn(2, 10, 8, generator=g)
    ref = {"enc_out": enc, "p_ctc": p}
    wl = torch.tensor([1.0, 0.55])                       # second utterance keeps floor(0.55 * 10) + 1 = 6 frames
    res = {"enc_out": enc.clone(), "p_ctc": p.clone(), "greedy": p.argmax(-1).int()}
    res["enc_out"][1, 6:] += 100.0                       # garbage in the padded frames must not count
    res["greedy"][1, 6:] = 0
    out = bench.parity_block(res, ref, 2, "bf16", wl)
    assert out["ok"] and out["enc_rel_l2"] == 0 and out["greedy_frame"] == 1.0 and out["greedy_seq"] == 1.0
    res["enc_out"][0] *= 1.05
    assert not bench.parity_block(res, ref, 2, "bf16", wl)["ok"]


def test_bench_graph_timeline_labels_the_launch_sequence_of_the_s_path():
    """bench.timeline_table: the kernel symbols of one graph replay of configs[1] (83 launches, as CUPTI reports them)
    are labelled with the work-table keys by symbol and launch order, and carry the same roofline arithmetic as the
    `kernels` table."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    layer = ["layernorm_kernel<2, 4>", "gemm_wres_kernel<256>", "mha2_bf16_kernel", "gemm_wres_kernel<256>",
             "layernorm_kernel<2, 4>", "ffn_fused_kernel"]
    sym = (["Memset", "fbank_tc2_kernel<true>", "topdb_norm_kernel", "conv0_tc_kernel", "gemm_bf16_kernel<true>",
            "kv_lengths_kernel", "gemm_bf16_kernel<false>"] + layer * 12 +
           ["layernorm_kernel<2, 4>", "gemm_bf16_kernel<false>", "ctc_reduce_kernel", "gemm_bf16_kernel<false>"])
    assert len(sym) == 83
    dur = [10.0] * len(sym)
    wt = bench.work_table(sb.MODEL_SIZES["S"], 64, 480000, 64 * 751 * 751)
    pk = {"tflops_sustained": 1387.0, "hbm_gbs": 6550.0}
    rows = {r["kernel"]: r for r in bench.timeline_table(sym, dur, wt, pk)}
    assert rows["stac_gemm_bf16:qkv"]["launches_per_step"] == 12 and rows["stac_gemm_bf16:out_proj"]["launches_per_step"] == 12
    assert rows["stac_layernorm"]["launches_per_step"] == 25 and rows["stac_mha_bf16_v2"]["launches_per_step"] == 12
    assert rows["stac_gemm_bf16:src_linear"]["launches_per_step"] == 1
    head = rows["stac_ctc_head_bf16"]                              # pass 1 + reduce + pass 2 = one launch of 30 us
    assert head["launches_per_step"] == 1 and head["avg_us"] == 30.0
    conv1 = rows["stac_conv1_bf16"]
    assert conv1["bound"] == "tensor" and abs(conv1["achieved"] - wt["stac_conv1_bf16"][1] / 10e-6 / 1e12) < 0.1
    assert abs(conv1["frac"] - conv1["achieved"] / 1387.0) < 1e-3
    assert "frac" not in rows["Memset"] and "frac" not in rows["kv_lengths_kernel"]
    # another model size (general GEMMs in the layers): no guessing, the shared symbol stays unlabelled
    other = bench.timeline_table(["gemm_bf16_kernel<false>"] * 5, [1.0] * 5, wt, pk)
    assert [r["kernel"] for r in other] == ["gemm_bf16_kernel<false>"]
