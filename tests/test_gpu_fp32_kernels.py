"""fp32 CUDA-core kernels against the oracle / plain torch fp32, through the C ABI (GPU only)."""
import ctypes

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


@pytest.mark.parametrize("variant", ["v2_pair", "v2_single", "v1"])
@pytest.mark.parametrize("n", [160 * 3, 16000, 16000 + 159, 5 * 16000 + 1, 160 * 127, 160 * 128, 160 * 129 + 8, 160 * 200 + 4,
                               160 * 255 + 12, 160 * 129 + 32, 160 * 200 + 64, 30 * 16000])
def test_fbank_tensor_core_stft(omods, n, variant, monkeypatch):
    """The tensor-core Fbank kernels (STFT as an fp16 tcgen05 GEMM on the folded frame) against the oracle Fbank:
    bf16-mode feature extraction, log-mel within 1e-3 relative (measured 3.6e-4; fp32 mode keeps the exact FFT kernel).
    v2_pair: stac_fbank_logmel_tc2 as two-CTA cta_group::2 instances (the default), v2_single: the same kernel per CTA,
    v1: stac_fbank_logmel_tc."""
    monkeypatch.setenv("STAC_FBANK_V2", "0" if variant == "v1" else "1")
    monkeypatch.setenv("STAC_FBANK_PAIR", "1" if variant == "v2_pair" else "0")
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


def _ordered_key(x: torch.Tensor) -> torch.Tensor:
    """The order-preserving uint32 encoding of a float the kernels' atomicMax works on (as int32 bit patterns)."""
    u = x.contiguous().view(torch.int32).long() & 0xffffffff
    k = torch.where(u >= 0x80000000, (~u) & 0xffffffff, u | 0x80000000)
    return torch.where(k >= 0x80000000, k - (1 << 32), k).to(torch.int32)


@pytest.mark.parametrize("pair", [1, 0])
def test_fbank_tc2_many_tiles_ragged_batch(pair, monkeypatch):
    """stac_fbank_logmel_tc2 on a batch with an odd tile count per utterance and more tiles than CTAs (every CTA walks
    several tiles, both PCM buffers and all ring slots wrap, the follower of the last pair repeats a tile): equal to the
    exact FFT kernel, guard bands around the output untouched, and bit-identical between two runs."""
    monkeypatch.setenv("STAC_FBANK_V2", "1")
    monkeypatch.setenv("STAC_FBANK_PAIR", str(pair))
    b, n = 37, 160 * (128 * 5 + 70)                       # 6 tiles per utterance, 222 tiles
    g = torch.Generator().manual_seed(5)
    wavs = (torch.randn(b, n, generator=g) * 0.1).cuda()
    tabs = ops.build_fbank_tc_tables("cuda")
    exact = ops.fbank(wavs, ops.build_fbank_tables("cuda"), raw=True).db
    got = ops.fbank_tc(wavs, tabs, raw=True)
    assert rel_l2(got.db, exact) < 1e-3, rel_l2(got.db, exact)
    assert torch.isfinite(got.db).all()
    again = ops.fbank_tc(wavs, tabs, raw=True)
    assert torch.equal(got.db, again.db) and torch.equal(got.utt_max, again.utt_max)
    # the per-utterance maximum is the maximum of what was written
    assert torch.equal(got.utt_max, _ordered_key(got.db.amax(dim=(1, 2))))


@pytest.mark.parametrize("pair", [1, 0])
def test_fbank_tc2_edge_shapes(pair, monkeypatch):
    """stac_fbank_logmel_tc2 at the edges of its contract: utterances shorter than one window, one utterance, an odd number
    of tiles (the follower of the last pair repeats a tile and stores nothing), rows that are views into a wider buffer
    (row stride > samples: what lies beyond an utterance's end in memory must not be read as audio)."""
    monkeypatch.setenv("STAC_FBANK_V2", "1")
    monkeypatch.setenv("STAC_FBANK_PAIR", str(pair))
    tabs = ops.build_fbank_tc_tables("cuda")
    fft = ops.build_fbank_tables("cuda")
    g = torch.Generator().manual_seed(11)
    for b, n in [(1, 32), (3, 64), (2, 320), (1, 160 * 128), (5, 160 * 130 + 96), (149, 1600)]:
        wavs = (torch.randn(b, n, generator=g) * 0.2).cuda()
        got = ops.fbank_tc(wavs, tabs, raw=True)
        want = ops.fbank(wavs, fft, raw=True)
        assert got.db.shape == (b, 1 + n // 160, 80)
        assert rel_l2(got.db, want.db) < 1e-3, (b, n, rel_l2(got.db, want.db))
        # strided rows: the same audio inside a buffer whose rows continue with large garbage
        wide = torch.full((b, n + 4096), 1e3, device="cuda")
        wide[:, :n] = wavs
        view = wide[:, :n]
        assert view.stride(0) == n + 4096
        db = torch.empty(b, 1 + n // 160, 80, device="cuda")
        umax = torch.empty(b, dtype=torch.int32, device="cuda")
        ops._call("stac_fbank_logmel_tc2", ctypes.c_void_p(view.data_ptr()), b, n, view.stride(0), ops.ptr(tabs.tab2),
                  ops.ptr(tabs.tw2, torch.float16), ops.ptr(db), ops.ptr(umax), pair, ops.stream())
        assert torch.equal(db, got.db), (b, n)
