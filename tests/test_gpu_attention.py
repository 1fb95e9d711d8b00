"""Fused attention (C ABI) vs the oracle's materialised softmax attention (fm_utils.py:160-180 restated) in fp32."""
import pytest
import torch

from oracle import fourm_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(B, H, Nq, Nk, seed, packed):
    g = torch.Generator().manual_seed(seed)
    D = H * 64
    if packed and Nq == Nk:
        qkv = (torch.randn(B * Nq, 3 * D, generator=g)).to(torch.bfloat16)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q = torch.randn(B * Nq, D, generator=g).to(torch.bfloat16)
        kv = torch.randn(B * Nk, 2 * D, generator=g).to(torch.bfloat16)
        k, v = kv[:, :D], kv[:, D:]
    return q, k, v


def _mask(kind, B, Nq, Nk, seed):
    g = torch.Generator().manual_seed(seed + 100)
    if kind == "none":
        return None
    if kind == "key":        # encoder key-padding mask [B,1,Nk]; make one sample fully masked
        m = torch.rand(B, 1, Nk, generator=g) < 0.3
        m[0] = True
        return m
    m = torch.rand(B, Nq, Nk, generator=g) < 0.5      # dense decoder-style mask [B,Nq,Nk] with some fully masked rows
    m[:, 0, :] = True
    return m


def _ref(q, k, v, B, H, Nq, Nk, mask):
    qf = q.float().reshape(B, Nq, H, 64).permute(0, 2, 1, 3)
    kf = k.float().reshape(B, Nk, H, 64).permute(0, 2, 1, 3)
    vf = v.float().reshape(B, Nk, H, 64).permute(0, 2, 1, 3)
    o = O._sdpa(qf, kf, vf, None if mask is None else mask[:, None], 64 ** -0.5)
    return o.permute(0, 2, 1, 3).reshape(B * Nq, H * 64)


CASES = [(2, 2, 128, 128), (3, 6, 128, 128), (2, 3, 100, 77), (1, 2, 24, 24), (2, 2, 256, 256), (2, 4, 200, 130), (1, 1, 130, 128),
         (2, 2, 128, 256)]


@pytest.mark.parametrize("B,H,Nq,Nk", CASES)
@pytest.mark.parametrize("mk", ["none", "key", "dense"])
def test_attention_fwd(B, H, Nq, Nk, mk):
    from b200fm import ops
    q, k, v = _inputs(B, H, Nq, Nk, 1, packed=True)
    mask = _mask(mk, B, Nq, Nk, 2)
    out, stats = ops.attention_fwd(q.cuda(), k.cuda(), v.cuda(), B, H, Nq, Nk, None if mask is None else mask.cuda())
    torch.cuda.synchronize()
    ref = _ref(q, k, v, B, H, Nq, Nk, mask)
    # P is rounded to bf16 before P.V (as in the reference's autocast attn@v) and the output is bf16
    torch.testing.assert_close(out.float().cpu(), ref, rtol=2e-2, atol=2e-2)


LONG_CASES = [(1, 2, 196, 300), (2, 3, 64, 513), (1, 4, 290, 1900), (2, 2, 257, 257), (1, 32, 256, 1024)]


@pytest.mark.parametrize("B,H,Nq,Nk", LONG_CASES)
@pytest.mark.parametrize("mk", ["none", "key", "dense", "causal"])
def test_attention_fwd_streamed_keys(B, H, Nq, Nk, mk):
    """Nk > 256 (generation callers, generate.py:407-445 / 886-913): keys streamed in 128-key tiles with the running-max rescale.
    'key' has a fully masked sample, 'dense' fully masked rows and tiles whose keys are all masked for a row, 'causal' is the AR
    decoder mask."""
    from b200fm import ops
    q, k, v = _inputs(B, H, Nq, Nk, 11, packed=False)
    if mk == "causal":
        mask = torch.triu(torch.ones(Nq, Nk, dtype=torch.bool), diagonal=1 + max(0, Nk - Nq))[None].expand(B, Nq, Nk).contiguous()
    else:
        mask = _mask(mk, B, Nq, Nk, 12)
        if mk == "dense":
            mask[:, 1, : min(Nk, 256)] = True           # first two key tiles fully masked for row 1, later tiles decide
    out, stats = ops.attention_fwd(q.cuda(), k.cuda(), v.cuda(), B, H, Nq, Nk, None if mask is None else mask.cuda())
    torch.cuda.synchronize()
    ref = _ref(q, k, v, B, H, Nq, Nk, mask)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=2e-2, atol=2e-2)
    # saved statistics: 1/sum and the log2-domain row max reproduce the oracle's log-sum-exp
    qf = q.float().reshape(B, Nq, H, 64).permute(0, 2, 1, 3)
    kf = k.float().reshape(B, Nk, H, 64).permute(0, 2, 1, 3)
    s = (qf @ kf.transpose(-1, -2)) * (64 ** -0.5)
    if mask is not None:
        s = s.masked_fill(mask[:, None], -torch.finfo(torch.float32).max)
    lse_ref = torch.logsumexp(s, dim=-1)
    st = stats.cpu()
    lse = (st[..., 0] - torch.log2(st[..., 1])) * 0.6931471805599453
    ok = lse_ref > -1e30                                   # fully masked rows: the reference's lse is -finfo.max + log(Nk)
    torch.testing.assert_close(lse[ok], lse_ref[ok], rtol=1e-3, atol=1e-2)


def test_attention_fwd_full_size_cfg2():
    """cfg-2 size (B=128, h=12, N=128) -- linearity property in V and agreement on a strided subsample."""
    from b200fm import ops
    B, H, N = 128, 12, 128
    q, k, v = _inputs(B, H, N, N, 3, packed=True)
    mask = _mask("key", B, N, N, 4)
    qc, kc, vc, mc = q.cuda(), k.cuda(), v.cuda(), mask.cuda()
    o1, _ = ops.attention_fwd(qc, kc, vc, B, H, N, N, mc)
    o2, _ = ops.attention_fwd(qc, kc, (vc.float() * 2).to(torch.bfloat16), B, H, N, N, mc)
    torch.testing.assert_close(o2.float(), o1.float() * 2, rtol=2e-2, atol=2e-2)
    sub = slice(0, 4)
    ref = _ref(q[: 4 * N], k[: 4 * N], v[: 4 * N], 4, H, N, N, mask[sub])
    torch.testing.assert_close(o1[: 4 * N].float().cpu(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 2, 128, 128), (3, 6, 128, 128), (2, 3, 100, 77), (1, 2, 24, 24), (2, 2, 256, 256), (2, 4, 200, 130)])
@pytest.mark.parametrize("mk", ["none", "key", "dense"])
@pytest.mark.parametrize("math_warps", [8, 16])
def test_attention_bwd(B, H, Nq, Nk, mk, math_warps):
    """vs fp32 autograd of the materialised attention; both CTA shapes (option "attn_bwd_warps": 8 or 16 softmax-backward warps)."""
    from b200fm import lib, ops
    lib.set_option("attn_bwd_warps", math_warps)
    try:
        _attention_bwd_case(B, H, Nq, Nk, mk)
    finally:
        lib.set_option("attn_bwd_warps", 8)


def _attention_bwd_case(B, H, Nq, Nk, mk):
    from b200fm import ops
    q, k, v = _inputs(B, H, Nq, Nk, 5, packed=False)
    mask = _mask(mk, B, Nq, Nk, 6)
    g = torch.Generator().manual_seed(7)
    dout = torch.randn(B * Nq, H * 64, generator=g).to(torch.bfloat16)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    _ref(qf, kf, vf, B, H, Nq, Nk, mask).backward(dout.float())
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    mc = None if mask is None else mask.cuda()
    out, stats = ops.attention_fwd(qc, kc, vc, B, H, Nq, Nk, mc)
    dq, dk, dv = ops.attention_bwd(qc, kc, vc, out, dout.cuda(), stats, B, H, Nq, Nk, mc)
    torch.cuda.synchronize()
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        err = (got.float().cpu() - ref).abs().max().item()
        scale = ref.abs().max().item() + 1e-6
        assert err <= 3e-2 * scale + 2e-2, f"{name}: max err {err} vs scale {scale}"


@pytest.mark.parametrize("B,H,Nk", [(2, 4, 1), (2, 32, 256), (3, 6, 77), (1, 12, 1500), (2, 2, 13000)])
@pytest.mark.parametrize("mk", ["none", "key", "all"])
def test_attention_decode_single_query(B, H, Nk, mk):
    """b200fm_attention_decode (one query row per sequence: the attention of the K/V-cached decode step) vs fp32 torch on the same bf16
    operands, K / V as the two column halves of a [B, Nk, 2D] cache, key masks incl. a fully masked row (uniform weights)."""
    from b200fm import ops
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Nk)
    D = H * 64
    q = torch.randn(B, D, device="cuda", generator=g).bfloat16()
    cache = torch.randn(B, Nk, 2 * D, device="cuda", generator=g).bfloat16()
    mask = None
    if mk == "key":
        mask = torch.rand(B, 1, Nk, device="cuda", generator=g) < 0.3
        mask[:, :, 0] = False
    elif mk == "all":
        mask = torch.ones(B, 1, Nk, dtype=torch.bool, device="cuda")
        mask[0] = torch.rand(1, Nk, device="cuda", generator=g) < 0.5
    c2 = cache.view(B * Nk, 2 * D)
    out = ops.attention_decode(q, c2[:, :D], c2[:, D:], B, H, Nk, mask, 0.125)
    qf = q.float().view(B, H, 1, 64)
    kf = cache[:, :, :D].float().view(B, Nk, H, 64).permute(0, 2, 1, 3)
    vf = cache[:, :, D:].float().view(B, Nk, H, 64).permute(0, 2, 1, 3)
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if mask is not None:
        s = s.masked_fill(mask[:, None], -torch.finfo(torch.float32).max)
    ref = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B, D)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
