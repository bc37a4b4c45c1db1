"""tcgen05 flash attention (stac_mha_bf16) against an fp32 softmax-attention with -inf key padding."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from util import rel_l2  # noqa: E402
from stac_speech_translation_b200 import ops  # noqa: E402


@pytest.mark.parametrize("t,lens,d,h", [(128, [128], 64, 1), (251, [251, 100, 1], 256, 4), (64, [64, 33], 128, 2),
                                         (130, [129, 130, 5], 256, 4), (751, [751, 400], 256, 4),
                                         (300, [300, 299], 512, 8), (300, [300 - 3 * i for i in range(40)], 256, 4),
                                         (100, [100 - i for i in range(90)], 128, 2), (1501, [1501, 1200], 128, 2)])
@pytest.mark.parametrize("use_vt", [False, True])
def test_mha_bf16(t, lens, d, h, use_vt):
    g = torch.Generator().manual_seed(t + d)
    b = len(lens)
    qkv = (torch.randn(b * t, 3 * d, generator=g)).to(torch.bfloat16)
    kv = torch.tensor(lens, dtype=torch.int32)
    t_pad = (t + 7) // 8 * 8
    vt = torch.zeros(b, h, 64, t_pad, dtype=torch.bfloat16)
    vt[..., :t] = qkv[:, 2 * d:].view(b, t, h, 64).permute(0, 2, 3, 1)
    ctx = torch.full((b * t, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    qkv_d, vt_d, kv_d = qkv.cuda(), vt.cuda().contiguous(), kv.cuda()   # keep alive across the async launch
    ops.check(ops.lib().stac_mha_bf16(ops.ptr(qkv_d), ops.ptr(vt_d if use_vt else None), ops.ptr(kv_d),
                                      b, t, t_pad, d, h, ops.ptr(ctx), ops.stream()))
    torch.cuda.synchronize()
    q, k, v = (x.float().view(b, t, h, 64).transpose(1, 2) for x in qkv.split(d, dim=-1))
    mask = torch.arange(t)[None, :] >= kv[:, None]
    s = (q @ k.transpose(-1, -2)).masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, d)
    got = ctx.float().cpu()
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 1e-2, rel_l2(got, ref)


# stac_mha_bf16_v2 (csrc/attention_tc2.cu) was written after the round-1 GPU budget was spent: it compiles for sm_100a but
# has not run on a B200 yet, so it is not part of the default GPU suite.  STAC_EXPERIMENTAL=1 enables it
# (tools/gpu_v2_check.sh runs it under a time limit, then times it against stac_mha_bf16).
@pytest.mark.skipif(os.environ.get("STAC_EXPERIMENTAL") != "1", reason="attention v2 not yet run on a B200")
@pytest.mark.parametrize("t,lens,d,h", [(128, [128], 64, 1), (256, [256], 64, 1), (251, [251, 100, 1], 256, 4),
                                         (64, [64, 33], 128, 2), (130, [129, 130, 5], 256, 4),
                                         (751, [751, 400], 256, 4), (300, [300, 299], 512, 8),
                                         (300, [300 - 3 * i for i in range(40)], 256, 4),
                                         (100, [100 - i for i in range(90)], 128, 2), (1501, [1501, 1200], 128, 2)])
def test_mha_bf16_v2(t, lens, d, h):
    g = torch.Generator().manual_seed(t + d)
    b = len(lens)
    qkv = (torch.randn(b * t, 3 * d, generator=g)).to(torch.bfloat16)
    kv = torch.tensor(lens, dtype=torch.int32)
    ctx = torch.full((b * t, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    guard = torch.full((4096,), 7.0, device="cuda", dtype=torch.bfloat16)      # allocated right behind ctx
    qkv_d, kv_d = qkv.cuda(), kv.cuda()
    ops.check(ops.lib().stac_mha_bf16_v2(ops.ptr(qkv_d), ops.ptr(kv_d), b, t, d, h, ops.ptr(ctx), ops.stream()))
    torch.cuda.synchronize()
    q, k, v = (x.float().view(b, t, h, 64).transpose(1, 2) for x in qkv.split(d, dim=-1))
    mask = torch.arange(t)[None, :] >= kv[:, None]
    s = (q @ k.transpose(-1, -2)).masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, d)
    got = ctx.float().cpu()
    assert not torch.isnan(got).any()
    assert (guard == 7.0).all()
    assert rel_l2(got, ref) < 1e-2, rel_l2(got, ref)
