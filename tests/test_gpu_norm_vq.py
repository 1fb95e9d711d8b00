"""LayerNorm fwd/bwd and the VQ codebook scan vs the oracle / torch fp32 on identical inputs."""
import pytest
import torch
import torch.nn.functional as F

from oracle import vq_oracle as V
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D", [(1, 128), (37, 384), (1000, 768), (16384, 768), (513, 1024), (64, 2048)])
@pytest.mark.parametrize("out_bf16", [True, False])
def test_layernorm_fwd(rows, D, out_bf16):
    from b200fm import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(rows, D, generator=g) * 2 + 0.3
    w = torch.randn(D, generator=g) * 0.1 + 1
    b = torch.randn(D, generator=g) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-6, out_bf16=out_bf16)
    ref = F.layer_norm(x, (D,), w, b, 1e-6)
    if out_bf16:
        torch.testing.assert_close(y.float().cpu(), ref.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
    else:
        torch.testing.assert_close(y.cpu(), ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(mean.cpu(), x.mean(-1), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rstd.cpu(), (x.var(-1, unbiased=False) + 1e-6).rsqrt(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("rows,D", [(37, 384), (4096, 768), (300, 1024)])
@pytest.mark.parametrize("dy_bf16", [True, False])
def test_layernorm_bwd(rows, D, dy_bf16):
    from b200fm import ops
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.3).requires_grad_(True)
    w = (torch.randn(D, generator=g) * 0.1 + 1).requires_grad_(True)
    b = (torch.randn(D, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(rows, D, generator=g)
    if dy_bf16:
        dy = dy.to(torch.bfloat16).float()
    dres = torch.randn(rows, D, generator=g)
    F.layer_norm(x, (D,), w, b, 1e-6).backward(dy)
    _, mean, rstd = ops.layernorm_fwd(x.detach().cuda(), w.detach().cuda(), b.detach().cuda(), 1e-6)
    dgamma = torch.zeros(D, device="cuda"); dbeta = torch.zeros(D, device="cuda")
    dyc = dy.cuda().to(torch.bfloat16) if dy_bf16 else dy.cuda()
    dx, dxb = ops.layernorm_bwd(dyc, x.detach().cuda(), w.detach().cuda(), mean, rstd, dres=dres.cuda(), want_bf16=True,
                                dgamma=dgamma, dbeta=dbeta)
    torch.testing.assert_close(dx.cpu(), x.grad + dres, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dxb.float().cpu(), (x.grad + dres).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dgamma.cpu(), w.grad, rtol=1e-3, atol=1e-3 * rows ** 0.5)
    torch.testing.assert_close(dbeta.cpu(), b.grad, rtol=1e-3, atol=1e-3 * rows ** 0.5)


def _check_scan(idx, ref, scores, tol):
    """index-exact except on genuine fp32 near-ties (documented parity rule)."""
    bad = idx != ref
    if bad.any():
        gap = (scores.gather(1, idx[:, None]) - scores.gather(1, ref[:, None])).abs()[bad]
        assert float(gap.max()) <= tol, f"{int(bad.sum())} mismatches, worst score gap {float(gap.max())}"
        assert bad.float().mean() < 1e-3


def test_vq_scan_golden_kats():
    from b200fm import ops
    s = H.load_golden("vq_golden.pt")["scan"]
    idx, quant = ops.vq_argmax(s["z"].cuda(), s["cos_embed"].cuda(), cosine=True, want_quant=True)
    _check_scan(idx.cpu(), s["cos_idx"], V.scan_scores(s["z"], s["cos_embed"], True), 1e-6)
    assert torch.equal(quant.cpu(), s["cos_embed"][idx.cpu()])
    idx = ops.vq_argmax(s["z"].cuda(), s["l2_embed"].cuda(), cosine=False)
    _check_scan(idx.cpu(), s["l2_idx"], V.scan_scores(s["z"], s["l2_embed"], False), 1e-4)


@pytest.mark.parametrize("n,K,d,cosine", [(1, 16, 32, True), (129, 1000, 32, True), (5000, 16384, 32, True), (777, 8192, 16, False),
                                          (300, 4096, 64, True), (4096, 1024, 8, False), (0, 128, 32, True)])
def test_vq_scan_vs_oracle(n, K, d, cosine):
    from b200fm import ops
    g = torch.Generator().manual_seed(n + K)
    z = torch.randn(n, d, generator=g)
    cb = torch.randn(K, d, generator=g)
    idx = ops.vq_argmax(z.cuda(), cb.cuda(), cosine=cosine).cpu()
    if n == 0:
        assert idx.numel() == 0
        return
    ref = V.cosine_scan(z, cb) if cosine else V.euclidean_scan(z, cb)
    _check_scan(idx, ref, V.scan_scores(z, cb, cosine), 1e-6 if cosine else 1e-4)


def test_vq_scan_ties_pick_lowest_index():
    from b200fm import ops
    cb = torch.randn(256, 32, generator=torch.Generator().manual_seed(3))
    cb[200] = cb[7]; cb[131] = cb[7]          # exact duplicates
    z = cb[[7, 131, 200, 9]].clone()
    idx = ops.vq_argmax(z.cuda(), cb.cuda(), cosine=True).cpu()
    assert idx.tolist() == [7, 7, 7, 9]


def test_vq_scan_full_size_properties():
    """BASELINE cfg-5 micro-benchmark size (n=131072, K=16384, d=32): idempotence of quantisation and a strided
    subsample check against the oracle."""
    from b200fm import ops
    g = torch.Generator().manual_seed(0)
    z = F.normalize(torch.randn(131072, 32, generator=g), dim=-1)
    cb = F.normalize(torch.empty(16384, 32).uniform_(-1, 1, generator=g), dim=-1)
    idx, quant = ops.vq_argmax(z.cuda(), cb.cuda(), cosine=True, want_quant=True)
    idx2 = ops.vq_argmax(quant, cb.cuda(), cosine=True)           # codes are fixed points
    assert torch.equal(idx, idx2)
    sub = torch.arange(0, 131072, 257)
    _check_scan(idx.cpu()[sub], V.cosine_scan(z[sub], cb), V.scan_scores(z[sub], cb, True), 1e-6)


@pytest.mark.parametrize("rows,D", [(37, 384), (16384, 768), (300, 1024), (9, 768)])
def test_layernorm_bwd_v2_matches_v1(rows, D):
    """The two LayerNorm-backward kernels (option "ln_bwd_v2": 0 = plain, 1 = all loads of a row hoisted ahead of the reductions -- the
    default) agree to fp32 rounding (the compiler contracts the expressions differently, so not bit for bit)."""
    from b200fm import lib, ops
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.3).cuda()
    w = (torch.randn(D, generator=g) * 0.1 + 1).cuda()
    dy = torch.randn(rows, D, generator=g).to(torch.bfloat16).cuda()
    dres = torch.randn(rows, D, generator=g).cuda()
    _, mean, rstd = ops.layernorm_fwd(x, w, None, 1e-6)
    outs = []
    for v in (0, 1):
        lib.set_option("ln_bwd_v2", v)
        try:
            dgamma = torch.zeros(D, device="cuda")
            dx, dxb = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_bf16=True, dgamma=dgamma)
            outs.append((dx.clone(), dxb.clone(), dgamma.clone()))
        finally:
            lib.set_option("ln_bwd_v2", 1)
    for v in (1,):
        torch.testing.assert_close(outs[0][0], outs[v][0], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(outs[0][1].float(), outs[v][1].float(), rtol=1e-2, atol=1e-2)
    for o in outs[1:]:
        torch.testing.assert_close(outs[0][2], o[2], rtol=1e-4, atol=1e-4 * rows ** 0.5)
