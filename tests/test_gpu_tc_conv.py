"""tcgen05 implicit-GEMM conv1 (stac_conv1_bf16) against torch conv2d on the same bf16-rounded data."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import rel_l2  # noqa: E402
from stac_speech_translation_b200 import ops  # noqa: E402


def _pack_padded(x0):
    """[B, T1, 40, 256] -> reflect-padded parity-split planes (the layout conv0 writes in bf16 mode)."""
    b, t1 = x0.shape[:2]
    tp2 = (t1 + 3) // 2
    pad = torch.nn.functional.pad(x0.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)
    full = torch.zeros(b, 2 * tp2, 42, 256)
    full[:, : t1 + 2] = pad
    planes = torch.zeros(b, 2, 2, tp2, 21, 256)
    for pt in range(2):
        for pf in range(2):
            planes[:, pt, pf] = full[:, pt::2, pf::2]
    return planes.to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("b,t1", [(1, 12), (2, 13), (3, 126), (2, 501), (1, 7), (5, 1501)])
def test_conv1_bf16_fused_block(b, t1):
    g = torch.Generator().manual_seed(t1)
    x0 = torch.randn(b, t1, 40, 256, generator=g).to(torch.bfloat16).float()
    w = (torch.randn(256, 256, 3, 3, generator=g) / 48).to(torch.bfloat16).float()
    bias = torch.randn(256, generator=g)
    gam = 1.0 + 0.2 * torch.randn(20, 256, generator=g)
    bet = 0.3 * torch.randn(20, 256, generator=g)
    xin = torch.nn.functional.pad(x0.permute(0, 3, 2, 1), (1, 1, 1, 1), mode="reflect")   # [B, C, F, T]
    pre = torch.nn.functional.conv2d(xin, w, bias, stride=2).permute(0, 3, 2, 1)         # [B, T2, 20, 256]
    t2 = (t1 - 1) // 2 + 1
    assert pre.shape == (b, t2, 20, 256)
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.layer_norm(pre, (20, 256), gam, bet, 1e-5), 0.01)
    wp = w.permute(2, 3, 0, 1).reshape(9, 256, 256).to(torch.bfloat16).contiguous()
    out = torch.full((b, t2, 20 * 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    guard = torch.full((4096,), 7.0, device="cuda", dtype=torch.bfloat16)        # allocated right after `out`
    x_d, w_d, b_d = _pack_padded(x0).cuda(), wp.cuda(), bias.cuda()     # keep alive across the async launch
    g_d, be_d = gam.flatten().cuda(), bet.flatten().cuda()
    ops.check(ops.lib().stac_conv1_bf16(ops.ptr(x_d), ops.ptr(w_d), ops.ptr(b_d), ops.ptr(g_d), ops.ptr(be_d), b, t1,
                                        ops.ptr(out), ops.stream()))
    torch.cuda.synchronize()
    got = out.float().cpu().view(b, t2, 20, 256)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 4e-3, rel_l2(got, ref)          # bf16 rounding of the stored activations
    assert float((guard.float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("b,t", [(1, 3), (2, 4), (1, 5), (3, 37), (2, 300), (1, 1001), (4, 3001)])
def test_conv0_tc_block_padded_layout(b, t):
    """Tensor-core block 0 (stac_conv0_ln_lrelu, bf16 mode) against torch: conv + LayerNorm(40,256) + LeakyReLU,
    every element of the reflect-padded parity-split layout that conv1 reads, ragged / odd / tiny lengths."""
    g = torch.Generator().manual_seed(1000 + t)
    feats = torch.randn(b, t, 80, generator=g) * 1.3 + 0.2
    w = torch.randn(256, 1, 3, 3, generator=g) / 3
    bias = torch.randn(256, generator=g) * 0.5
    gam = 1.0 + 0.2 * torch.randn(40, 256, generator=g)
    bet = 0.3 * torch.randn(40, 256, generator=g)
    xin = torch.nn.functional.pad(feats.transpose(1, 2).unsqueeze(1), (1, 1, 1, 1), mode="reflect")   # [B,1,F,T]
    pre = torch.nn.functional.conv2d(xin.double(), w.double(), bias.double(), stride=2).permute(0, 3, 2, 1)
    t1 = (t - 1) // 2 + 1
    assert pre.shape == (b, t1, 40, 256)
    ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.layer_norm(pre, (40, 256), gam.double(), bet.double(), 1e-5), 0.01).float()
    n = ops.lib().stac_conv0_padded_elems(b, t1)
    x0 = torch.full((n,), float("nan"), device="cuda", dtype=torch.bfloat16)
    guard = torch.full((4096,), 7.0, device="cuda", dtype=torch.bfloat16)
    f_d, w_d, b_d = feats.cuda().contiguous(), w.reshape(256, 3, 3).cuda().contiguous(), bias.cuda()
    g_d, be_d = gam.flatten().cuda(), bet.flatten().cuda()
    ops.check(ops.lib().stac_conv0_ln_lrelu(ops.ptr(f_d), ops.ptr(w_d), ops.ptr(b_d), ops.ptr(g_d), ops.ptr(be_d),
                                            b, t, ops.ptr(x0), ops.DT_BF16, ops.stream()))
    torch.cuda.synchronize()
    tp2 = (t1 + 3) // 2
    planes = x0.view(b, 2, 2, tp2, 21, 256).float().cpu()
    pad = torch.full((b, 2 * tp2, 42, 256), float("nan"))
    for pt in range(2):
        for pf in range(2):
            pad[:, pt::2, pf::2] = planes[:, pt, pf]
    want = torch.nn.functional.pad(ref.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)
    got = pad[:, : t1 + 2, :41]
    want = want[:, :, :41]
    assert not torch.isnan(got).any()
    assert rel_l2(got, want) < 2.5e-3, rel_l2(got, want)      # bf16 rounding of the stored value only
    # the value before rounding is fp32-accurate: positive outputs are the correctly rounded bf16 almost always;
    # negative ones go through one more bf16 rounding (LeakyReLU runs on packed bf16 pairs): within 1 ulp (2^-7 rel)
    wb = want.to(torch.bfloat16).float()
    pos = want > 0
    exact = (got[pos] == wb[pos]).float().mean()
    assert exact > 0.97, float(exact)
    assert float(((got - want).abs() <= want.abs() * 2.0 ** -7 + 1e-4).float().mean()) > 0.9999
    assert float((guard.float() - 7.0).abs().max()) == 0.0


def test_conv0_with_fused_topdb_norm_equals_the_two_kernel_path():
    """stac_conv0_topdb_norm_bf16 (clamp + normalisation inside block 0's loader) must give what stac_fbank_topdb_norm
    followed by stac_conv0_ln_lrelu gives (the loader multiplies by 1 / std where the separate kernel divides: one fp32
    ulp on the features, far below the bf16 output's resolution) - for the per-utterance and the batch-global top-dB
    rule, with and without statistics, on a ragged batch."""
    import stac_speech_translation_b200 as sb
    from stac_speech_translation_b200 import ops, synth
    mods = sb.build_modules(sb.HParams.for_size("S", num_encoder_layers=1), precision="bf16", device="cuda")
    w = mods["CNN"].packed()
    wavs, _ = synth.synth_batch([2.0, 0.7, 1.31], seed=5)
    wavs = wavs[:, : wavs.shape[1] // 4 * 4].contiguous().cuda()
    wavs[1] *= 0.01                                        # a quiet utterance: its own maximum matters
    g = torch.Generator().manual_seed(3)
    mean = (torch.randn(80, generator=g) * 5 - 20).cuda()
    std = (torch.rand(80, generator=g) * 10 + 5).cuda()
    tabs = ops.build_fbank_tc_tables(wavs.device)
    for per_utt in (True, False):
        for stats in ((mean, std), (None, None)):
            feats = ops.fbank_tc(wavs, tabs, 80.0, per_utt, *stats)
            raw = ops.fbank_tc(wavs, tabs, 80.0, per_utt, *stats, raw=True)
            a = ops.conv_frontend(feats, w, torch.bfloat16)
            b = ops.conv_frontend(raw, w, torch.bfloat16)
            torch.cuda.synchronize()
            assert rel_l2(a, b) < 1e-3, (per_utt, stats[0] is not None, rel_l2(a, b))
            if stats[0] is None:
                assert torch.equal(a, b)                # clamp only: no arithmetic differs
