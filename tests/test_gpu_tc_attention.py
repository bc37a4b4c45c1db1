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


# stac_mha_bf16_v2 (csrc/attention_tc2.cu): P in TMEM, one thread per query row, 128-key tiles.
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


@pytest.mark.parametrize("jump_at", [40, 70, 100, 130, 170, 300, 500])
def test_mha_bf16_v2_reference_raise(jump_at):
    """The lazy-rescaling path of stac_mha_bf16_v2: keys from position `jump_at` on score far above everything before
    them (more than 2^40 above the running reference), so the reference maximum must be raised in the middle of an item -
    in the second / third 32-key chunk of the first tile (only this tile's partial sums and P chunks are rescaled), and in
    later tiles (O in tensor memory is rescaled too) - and again a second time further on."""
    g = torch.Generator().manual_seed(jump_at)
    b, t, d, h = 2, 600, 128, 2
    qkv = torch.randn(b * t, 3 * d, generator=g) * 0.5
    q = qkv[:, :d].view(b, t, h, 64)
    k = qkv[:, d:2 * d].view(b, t, h, 64)
    # the keys from jump_at on are aligned with the mean query direction and large; a second, larger step later
    direction = q.mean(1, keepdim=True) / q.mean(1, keepdim=True).norm(dim=-1, keepdim=True)
    k[:, jump_at:] += 30.0 * direction
    k[:, min(t - 1, jump_at + 150):] += 30.0 * direction
    q += 4.0 * direction                                   # every query has a positive component along it
    qkv = qkv.to(torch.bfloat16)
    kv = torch.tensor([t, t - 37], dtype=torch.int32)
    ctx = torch.full((b * t, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    qkv_d, kv_d = qkv.cuda(), kv.cuda()
    ops.check(ops.lib().stac_mha_bf16_v2(ops.ptr(qkv_d), ops.ptr(kv_d), b, t, d, h, ops.ptr(ctx), ops.stream()))
    torch.cuda.synchronize()
    qf, kf, vf = (x.float().view(b, t, h, 64).transpose(1, 2) for x in qkv.split(d, dim=-1))
    mask = torch.arange(t)[None, :] >= kv[:, None]
    s = (qf @ kf.transpose(-1, -2)).masked_fill(mask[:, None, None, :], float("-inf"))
    assert float((s[..., jump_at:].amax(-1) - s[..., :jump_at].amax(-1)).min()) > 40.0      # the jump is there, in nats
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(b * t, d)
    got = ctx.float().cpu()
    assert not torch.isnan(got).any() and bool(torch.isfinite(got).all())
    assert rel_l2(got, ref) < 1e-2, rel_l2(got, ref)


# Stress shape for the finding in DESIGN.md section 9: many consecutive one-tile work items per CTA in which both query
# groups are active (short utterances in a batch padded beyond 128 frames), which is what lets one softmax group run
# two items ahead of the store warp.  The kernel now has one o_staged barrier per (Q buffer, group), which cannot run
# ahead (tools/model_check_mha1.py); a regression would show as the bounded mbarrier wait trapping.
def test_mha_bf16_many_short_items_per_cta():
    g = torch.Generator().manual_seed(7)
    b, t, d, h = 600, 300, 256, 4
    lens = torch.randint(1, 65, (b,), generator=g, dtype=torch.int32)
    lens[::7] = 300                                         # a few long utterances in between
    qkv = torch.randn(b * t, 3 * d, generator=g).to(torch.bfloat16).cuda()
    kv = lens.cuda()
    ctx = torch.empty(b * t, d, device="cuda", dtype=torch.bfloat16)
    for _ in range(20):
        ops.check(ops.lib().stac_mha_bf16(ops.ptr(qkv), ops.ptr(None), ops.ptr(kv), b, t, (t + 7) // 8 * 8, d, h,
                                          ops.ptr(ctx), ops.stream()))
    torch.cuda.synchronize()
    # spot-check a few utterances against fp32 attention
    for i in (0, 1, 7, 599):
        q, k, v = (x.float().view(t, h, 64).transpose(0, 1) for x in qkv[i * t:(i + 1) * t].cpu().split(d, dim=-1))
        n = int(lens[i])
        p = torch.softmax(q @ k[:, :n].transpose(-1, -2), -1)
        ref = (p @ v[:, :n]).transpose(0, 1).reshape(t, d)
        assert rel_l2(ctx[i * t:(i + 1) * t].float().cpu(), ref) < 1e-2
