"""tcgen05 GEMM (stac_gemm_bf16) against torch fp32 matmul of the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import rel_l2, rel_max  # noqa: E402
from stac_speech_translation_b200 import ops  # noqa: E402


def _run(m, n, k, bias=True, resid=False, period=0, act=ops.ACT_NONE, out_bf16=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(m, k, generator=g)).to(torch.bfloat16)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(n, generator=g) if bias else None
    rows_r = period if period else m
    r = torch.randn(rows_r, n, generator=g) if resid else None
    c = torch.full((m, n), float("nan"), device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ops._gemm(a.cuda(), w.cuda(), None if b is None else b.cuda(), c, "bf16",
              resid=None if r is None else r.cuda(), resid_period=period, act=act)
    ref = a.float() @ w.float().T
    if b is not None:
        ref = ref + b
    if act == ops.ACT_GELU_ERF:
        ref = torch.nn.functional.gelu(ref)
    if r is not None:
        ref = ref + (r.repeat(m // period, 1) if period else r)
    return c.float().cpu(), ref


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 256, 256), (300, 256, 256), (1000, 768, 256),
                                    (257, 1024, 256), (129, 256, 1024), (515, 5000, 256), (200, 512, 512),
                                    (4096 + 17, 256, 5120)])
def test_shapes_block_n_256(m, n, k):
    got, ref = _run(m, n, k)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 1e-5, rel_l2(got, ref)


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (333, 384, 256), (130, 640, 128)])
def test_shapes_block_n_128(m, n, k):
    got, ref = _run(m, n, k)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 1e-5


def test_epilogues():
    got, ref = _run(300, 1024, 256, act=ops.ACT_GELU_ERF, out_bf16=True)
    assert rel_l2(got, ref) < 4e-3
    got, ref = _run(300, 256, 1024, resid=True)
    assert rel_l2(got, ref) < 1e-5
    got, ref = _run(43 * 6, 256, 512, resid=True, period=43)
    assert rel_l2(got, ref) < 1e-5
    got, ref = _run(300, 256, 256, bias=False)
    assert rel_l2(got, ref) < 1e-5


def test_many_tiles_persistent_loop():
    # more tiles than SMs so every CTA loops, both TMEM accumulator stages and all smem stages wrap
    got, ref = _run(128 * 200 + 5, 1024, 256, seed=3)
    assert rel_l2(got, ref) < 1e-5
    assert rel_max(got, ref) < 1e-4


@pytest.mark.parametrize("m,n", [(128 * 100 + 37, 768), (128 * 300 + 5, 256), (128 * 40, 2048), (128 * 149, 512)])
def test_weight_resident_projection(m, n):
    """K = 256 with enough row tiles: stac_gemm_bf16 routes to the weight-resident kernel (csrc/gemm_wres.cu).
    bf16 store with bias (QKV), fp32 store without bias, fp32 in-place residual through TMA reduce-add (out-proj)."""
    got, ref = _run(m, n, 256, out_bf16=True, seed=5)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 4e-3
    got, ref = _run(m, n, 256, bias=False, seed=6)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 1e-5 and rel_max(got, ref) < 1e-4
    g = torch.Generator().manual_seed(7)
    a = torch.randn(m, 256, generator=g).to(torch.bfloat16)
    w = (torch.randn(n, 256, generator=g) / 16).to(torch.bfloat16)
    b = torch.randn(n, generator=g)
    x = torch.randn(m, n, generator=g)
    c = x.cuda()
    ops._gemm(a.cuda(), w.cuda(), b.cuda(), c, "bf16", resid=c)
    ref = x + a.float() @ w.float().T + b
    assert rel_l2(c.cpu(), ref) < 1e-5 and rel_max(c.cpu(), ref) < 1e-4


def test_qkv_with_transposed_v():
    g = torch.Generator().manual_seed(1)
    b, t, d, h = 3, 77, 256, 4
    m = b * t
    a = torch.randn(m, d, generator=g).to(torch.bfloat16)
    w = (torch.randn(3 * d, d, generator=g) / 16).to(torch.bfloat16)
    bias = torch.randn(3 * d, generator=g)
    t_pad = (t + 7) // 8 * 8
    qkv = torch.zeros(m, 3 * d, device="cuda", dtype=torch.bfloat16)
    vt = torch.zeros(b * h * 64 * t_pad, device="cuda", dtype=torch.bfloat16)
    ops._gemm(a.cuda(), w.cuda(), bias.cuda(), qkv, "bf16", vt=vt, vt_cols=d, seq_len=t, t_pad=t_pad)
    ref = a.float() @ w.float().T + bias
    assert rel_l2(qkv[:, : 2 * d].float(), ref[:, : 2 * d]) < 4e-3
    v_ref = ref[:, 2 * d:].view(b, t, h, 64).permute(0, 2, 3, 1)          # [B, H, 64, T]
    v_got = vt.view(b, h, 64, t_pad).float().cpu()
    assert rel_l2(v_got[..., :t], v_ref) < 4e-3
    assert float(v_got[..., t:].abs().max()) == 0.0


@pytest.mark.parametrize("m,v,d", [(300, 5000, 256), (77, 64, 128), (1000, 5000, 512), (129, 5000, 1024), (5, 136, 256)])
def test_fused_ctc_head(m, v, d):
    """stac_ctc_head_bf16 (two-pass GEMM + log-softmax + argmax) against torch on the same bf16-rounded operands."""
    g = torch.Generator().manual_seed(m + v)
    x = torch.randn(m, d, generator=g).to(torch.bfloat16)
    w = (torch.randn(v, d, generator=g) / d ** 0.5 * 3).to(torch.bfloat16)
    b = torch.randn(v, generator=g)
    logits = x.float() @ w.float().t() + b
    ref = torch.log_softmax(logits, -1)
    got, ids = ops.ctc_head_bf16(x.cuda(), w.cuda(), b.cuda())
    torch.cuda.synchronize()
    assert got.shape == (m, v) and ids.shape == (m,)
    assert float((got.cpu() - ref).abs().max()) < 2e-4
    assert float((got.cpu().exp().sum(-1) - 1).abs().max()) < 1e-4        # posteriors sum to one
    # greedy ids: equal to torch's argmax except on near ties (fp32 summation order differs)
    top2 = logits.topk(2, -1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert bool((ids.cpu().long()[clear] == logits.argmax(-1)[clear]).all())
    assert bool((logits.gather(1, ids.cpu().long()[:, None])[:, 0] >= top2[:, 0] - 1e-4).all())


@pytest.mark.parametrize("m,dffn", [(128, 1024), (300, 1024), (1000, 1024), (77, 128), (20000, 1024), (513, 256), (129, 384)])
def test_fused_ffn_block(m, dffn):
    """stac_ffn_fused_bf16: x += W2 GELU(W1 h + b1) + b2 in one kernel against torch (exact-erf GELU) on the same
    bf16-rounded operands, and against the two-GEMM path it replaces."""
    g = torch.Generator().manual_seed(m + dffn)
    h = torch.randn(m, 256, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(dffn, 256, generator=g) / 16).to(torch.bfloat16)
    w2 = (torch.randn(256, dffn, generator=g) / dffn ** 0.5).to(torch.bfloat16)
    b1, b2 = torch.randn(dffn, generator=g) * 0.3, torch.randn(256, generator=g) * 0.3
    x0 = torch.randn(m, 256, generator=g)
    hid = torch.nn.functional.gelu(h.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    ref = x0 + hid @ w2.float().t() + b2
    x = x0.cuda().clone()
    guard = torch.full((4096,), 7.0, device="cuda")
    h_d, w1_d, w2_d, b1_d, b2_d = h.cuda(), w1.cuda(), w2.cuda(), b1.cuda(), b2.cuda()
    ops.check(ops.lib().stac_ffn_fused_bf16(ops.ptr(h_d), ops.ptr(w1_d), ops.ptr(b1_d), ops.ptr(w2_d), ops.ptr(b2_d),
                                            ops.ptr(x), m, 256, dffn, ops.stream()))
    torch.cuda.synchronize()
    assert rel_l2(x, ref) < 2e-3, rel_l2(x, ref)
    assert rel_max(x, ref) < 1e-2
    assert float((guard - 7.0).abs().max()) == 0.0
    # unsupported widths are refused, not mis-computed
    assert ops.lib().stac_ffn_fused_bf16(ops.ptr(h_d), ops.ptr(w1_d), ops.ptr(b1_d), ops.ptr(w2_d), ops.ptr(b2_d),
                                         ops.ptr(x), m, 512, dffn, ops.stream()) == -2


@pytest.mark.parametrize("m", [1, 127, 128, 129, 4000, 48064])
def test_outproj_residual_layernorm_fused(m):
    """stac_outproj_ln_bf16 against fp64 arithmetic on the same bf16 operands: x += ctx W^T + b (fp32 residual stream, in
    place), h = LayerNorm(x) in bf16; rows past m untouched (guard band), including the ragged last tile."""
    from stac_speech_translation_b200 import ops
    g = torch.Generator().manual_seed(m)
    a = (torch.randn(m, 256, generator=g)).to(torch.bfloat16)
    w = (torch.randn(256, 256, generator=g) / 16).to(torch.bfloat16)
    bias = torch.randn(256, generator=g)
    x0 = torch.randn(m, 256, generator=g) * 3 + 0.5
    gam, bet = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    xg = torch.full((m + 64, 256), 7.0)
    xg[:m] = x0
    xg = xg.cuda()
    hg = torch.full((m + 64, 256), 3.0, dtype=torch.bfloat16).cuda()
    a_d, w_d, b_d, g_d, be_d = a.cuda(), w.cuda(), bias.cuda(), gam.cuda(), bet.cuda()     # (kept alive: ptr() of a
    ops._call("stac_outproj_ln_bf16", ops.ptr(a_d), ops.ptr(w_d), ops.ptr(b_d), ops.ptr(xg),  # temporary dangles)
              ops.ptr(g_d), ops.ptr(be_d), 1e-6, ops.ptr(hg), m, ops.stream())
    torch.cuda.synchronize()
    x_ref = x0.double() + a.double() @ w.double().T + bias.double()
    mu, var = x_ref.mean(-1, keepdim=True), x_ref.var(-1, unbiased=False, keepdim=True)
    h_ref = (x_ref - mu) / torch.sqrt(var + 1e-6) * gam.double() + bet.double()
    assert rel_l2(xg[:m].cpu(), x_ref) < 1e-5
    assert rel_l2(hg[:m].float().cpu(), h_ref) < 4e-3                 # bf16 rounding of the output
    assert torch.all(xg[m:] == 7.0) and torch.all(hg[m:].float() == 3.0)
