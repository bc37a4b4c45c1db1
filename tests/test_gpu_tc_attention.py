"""tcgen05 flash attention (stac_mha_bf16) against an fp32 softmax-attention with -inf key padding."""
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
