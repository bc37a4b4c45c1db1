"""fp32 CUDA-core kernels against the oracle / plain torch fp32, through the C ABI (GPU only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import FP32_TOL, oracle_modules, product_from_oracle, rel_l2, rel_max  # noqa: E402
from stac_speech_translation_b200 import ops, synth  # noqa: E402


@pytest.fixture(scope="module")
def omods():
    return oracle_modules("S", num_encoder_layers=2)


@pytest.fixture(scope="module")
def batch():
    return synth.synth_batch([3.0, 2.2, 1.51, 0.7], seed=5)


def test_fbank_matches_oracle(omods, batch):
    wavs, _ = batch
    ref = omods["compute_features"](wavs)
    tab = ops.build_fbank_tables("cuda")
    got = ops.fbank(wavs.cuda(), tab)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < FP32_TOL
    assert rel_max(got, ref) < 5e-4
    # batch-global top-dB variant
    omods["compute_features"].compute_fbanks.top_db_per_utterance = False
    try:
        ref_g = omods["compute_features"](wavs)
    finally:
        omods["compute_features"].compute_fbanks.top_db_per_utterance = True
    got_g = ops.fbank(wavs.cuda(), tab, per_utterance=False)
    assert rel_l2(got_g, ref_g) < FP32_TOL


@pytest.mark.parametrize("n", [160 * 3, 16000, 16000 + 159, 5 * 16000 + 1, 160 * 64 - 1])
def test_fbank_ragged_lengths(omods, n):
    g = torch.Generator().manual_seed(n)
    wavs = torch.randn(2, n, generator=g) * 0.1
    ref = omods["compute_features"](wavs)
    got = ops.fbank(wavs.cuda(), ops.build_fbank_tables("cuda"))
    assert got.shape == ref.shape == (2, 1 + n // 160, 80)
    assert rel_l2(got, ref) < FP32_TOL


def test_fbank_silence_and_fused_norm(omods, batch):
    wavs, wl = batch
    z = torch.zeros(2, 8000)
    ref = omods["compute_features"](z)
    got = ops.fbank(z.cuda(), ops.build_fbank_tables("cuda"))
    assert torch.allclose(got.cpu(), ref, atol=1e-4)   # -100 dB everywhere, clamp is a no-op
    norm = omods["normalize"]
    ref_n = norm(omods["compute_features"](wavs), wl)
    got_n = ops.fbank(wavs.cuda(), ops.build_fbank_tables("cuda"), mean=norm.glob_mean.cuda(),
                      std=norm.glob_std.cuda())
    assert rel_l2(got_n, ref_n) < FP32_TOL
    got_s = ops.input_norm(omods["compute_features"](wavs).cuda(), norm.glob_mean.cuda(), norm.glob_std.cuda())
    assert rel_l2(got_s, ref_n) < 1e-6


def test_conv_frontend_fp32(omods, batch):
    wavs, wl = batch
    feats = omods["normalize"](omods["compute_features"](wavs), wl)
    ref = omods["CNN"](feats)
    mods = product_from_oracle(omods, "fp32")
    got = mods["CNN"](feats.cuda())
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert rel_l2(got, ref) < FP32_TOL
    assert rel_max(got, ref) < 5e-4


def test_conv0_padded_bf16_layout(omods, batch):
    wavs, wl = batch
    feats = omods["normalize"](omods["compute_features"](wavs), wl)
    ref = omods["CNN"].convblock_0(feats)                      # [B, T1, 40, 256]
    mods = product_from_oracle(omods, "bf16")
    w = mods["CNN"].packed()
    b, t, _ = feats.shape
    t1 = (t - 1) // 2 + 1
    n = ops.lib().stac_conv0_padded_elems(b, t1)
    x0 = torch.full((n,), float("nan"), device="cuda", dtype=torch.bfloat16)
    f = feats.cuda().contiguous()
    ops.check(ops.lib().stac_conv0_ln_lrelu(ops.ptr(f), ops.ptr(w.w0), ops.ptr(w.b0), ops.ptr(w.g0), ops.ptr(w.be0),
                                            b, t, ops.ptr(x0), ops.DT_BF16, ops.stream()))
    tp2 = (t1 + 3) // 2
    planes = x0.view(b, 2, 2, tp2, 21, 256).float().cpu()
    # rebuild the reflect-padded tensor [B, T1+2, 42, 256] from the parity planes
    pad = torch.full((b, 2 * tp2, 42, 256), float("nan"))
    for pt in range(2):
        for pf in range(2):
            pad[:, pt::2, pf::2] = planes[:, pt, pf]
    want = torch.nn.functional.pad(ref.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)
    got = pad[:, : t1 + 2, :41]          # column 41 is never read by the stride-2 conv
    want = want[:, :, :41]
    assert not torch.isnan(got).any()
    assert rel_l2(got, want) < 4e-3      # bf16 rounding of the stored activations


def test_layernorm_gemm_softmax_fp32():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(333, 512, generator=g) * 3 + 0.5
    gam, bet = torch.randn(512, generator=g), torch.randn(512, generator=g)
    ref = torch.nn.functional.layer_norm(x, (512,), gam, bet, 1e-6)
    o32 = torch.empty(333, 512, device="cuda")
    o16 = torch.empty(333, 512, device="cuda", dtype=torch.bfloat16)
    ops._layernorm(x.cuda(), gam.cuda(), bet.cuda(), 1e-6, out_f32=o32, out_bf16=o16)
    assert rel_l2(o32, ref) < 1e-5
    assert rel_l2(o16.float(), ref) < 4e-3

    a, w, bias = torch.randn(301, 1024, generator=g), torch.randn(257 * 4, 1024, generator=g), torch.randn(257 * 4, generator=g)
    res = torch.randn(301, 257 * 4, generator=g)
    c = torch.empty(301, 257 * 4, device="cuda")
    ops._gemm(a.cuda(), w.cuda(), bias.cuda(), c, "fp32", resid=res.cuda())
    assert rel_l2(c, a @ w.T + bias + res) < 1e-5
    ops._gemm(a.cuda(), w.cuda(), bias.cuda(), c, "fp32", act=ops.ACT_GELU_ERF)
    assert rel_l2(c, torch.nn.functional.gelu(a @ w.T + bias)) < 1e-5
    pe = torch.randn(43, 257 * 4, generator=g)
    ops._gemm(a[:43 * 7].cuda(), w.cuda(), bias.cuda(), c[:43 * 7], "fp32", resid=pe.cuda(), resid_period=43)
    assert rel_l2(c[:43 * 7], a[:43 * 7] @ w.T + bias + pe.repeat(7, 1)) < 1e-5

    logits = torch.randn(77, 5000, generator=g) * 4
    lp, ids = ops.log_softmax(logits.cuda(), want_argmax=True)
    assert rel_l2(lp, torch.log_softmax(logits, -1)) < 1e-6
    assert torch.equal(ids.cpu().long(), logits.argmax(-1))


@pytest.mark.parametrize("t,lens", [(251, [251, 100, 1]), (64, [64, 33]), (130, [129, 130, 5])])
def test_mha_fp32(t, lens):
    g = torch.Generator().manual_seed(t)
    b, d, h = len(lens), 256, 4
    qkv = torch.randn(b * t, 3 * d, generator=g)
    kv = torch.tensor(lens, dtype=torch.int32)
    ctx = torch.empty(b * t, d, device="cuda")
    qkv_d, kv_d = qkv.cuda(), kv.cuda()      # keep the device tensors alive across the async launch
    ops.check(ops.lib().stac_mha_f32(ops.ptr(qkv_d), ops.ptr(kv_d), b, t, d, h, ops.ptr(ctx), ops.stream()))
    torch.cuda.synchronize()
    q, k, v = (x.view(b, t, h, 64).transpose(1, 2) for x in qkv.split(d, dim=-1))
    mask = torch.arange(t)[None, :] >= kv[:, None]
    s = (q @ k.transpose(-1, -2)).masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, d)
    assert rel_l2(ctx, ref) < 1e-5


@pytest.mark.parametrize("n", [160 * 3, 16000, 16000 + 159, 5 * 16000 + 1, 160 * 127, 160 * 128, 160 * 129 + 8, 160 * 200 + 4,
                               30 * 16000])
def test_fbank_tensor_core_stft(omods, n):
    """stac_fbank_logmel_tc (STFT as an fp16 tcgen05 GEMM, folded DFT) against the oracle Fbank: bf16-mode feature
    extraction, log-mel within 1e-3 relative (measured 3.6e-4; fp32 mode keeps the exact FFT kernel)."""
    wavs, _ = synth.synth_batch([n / 16000.0, max(0.03, 0.61 * n / 16000.0)], seed=n)
    wavs = wavs[:, :n].contiguous()
    ref = omods["compute_features"](wavs)
    tabs = ops.build_fbank_tc_tables("cuda")
    got = ops.fbank_tc(wavs.cuda(), tabs)
    assert got.shape == ref.shape == (2, 1 + n // 160, 80)
    assert rel_l2(got, ref) < 1e-3, rel_l2(got, ref)
    exact = ops.fbank(wavs.cuda(), ops.build_fbank_tables("cuda"))
    assert rel_l2(got, exact) < 1e-3
    # fused normalisation path and batch-global clamp share the second kernel with the FFT version
    norm = omods["normalize"]
    got_n = ops.fbank_tc(wavs.cuda(), tabs, mean=norm.glob_mean.cuda(), std=norm.glob_std.cuda())
    ref_n = norm(ref, torch.ones(2))
    assert rel_l2(got_n, ref_n) < 2e-3
