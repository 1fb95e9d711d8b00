"""tcgen05 GEMM (through the C ABI) vs a plain fp32 torch reference on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _ref(a, b, layout):
    a32, b32 = a.float(), b.float()
    if layout == 0:
        return a32 @ b32.t()
    if layout == 1:
        return a32 @ b32
    return a32.t() @ b32


SHAPES = [(128, 128, 64), (128, 256, 64), (256, 256, 128), (384, 768, 768), (1000, 520, 200), (136, 72, 72), (16384, 768, 768)]


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_f32_out(layout, M, N, K):
    from b200fm import ops
    if layout == 0:
        a, b = _mk((M, K), 1), _mk((N, K), 2)
    elif layout == 1:
        a, b = _mk((M, K), 1), _mk((K, N), 2)
    else:
        a, b = _mk((K, M), 1), _mk((K, N), 2)
    out = ops.gemm(a, b, layout=layout, epilogue=ops.EPI_F32)
    torch.cuda.synchronize()
    ref = _ref(a, b, layout)
    # bf16 products are exact in fp32; only the accumulation order differs
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-3 * (K ** 0.5) / 8)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1000, 520, 200), (16384, 2304, 768)])
def test_gemm_bf16_bias_alpha(M, N, K):
    from b200fm import ops
    a, b = _mk((M, K), 3), _mk((N, K), 4)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(5)).cuda()
    adev = torch.tensor([0.5], device="cuda")
    out = ops.gemm(a, b, epilogue=ops.EPI_BF16, bias=bias, alpha=2.0, alpha_dev=adev)
    ref = ((a.float() @ b.float().t() + bias) * 1.0).to(torch.bfloat16)
    torch.testing.assert_close(out.float(), ref.float(), rtol=1e-2, atol=1e-2 * (K ** 0.5) / 8)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (16384, 768, 2048), (300, 392, 136)])
def test_gemm_residual(M, N, K):
    from b200fm import ops
    a, b = _mk((M, K), 6, 0.5), _mk((N, K), 7, 0.5)
    resid = torch.randn(M, N, generator=torch.Generator().manual_seed(8)).cuda()
    out = ops.gemm(a, b, epilogue=ops.EPI_RESID, resid=resid)
    ref = resid + (a.float() @ b.float().t()).to(torch.bfloat16).float()
    torch.testing.assert_close(out, ref, rtol=1e-2, atol=2e-2 * (K ** 0.5) / 8)


@pytest.mark.parametrize("M,H,K", [(256, 128, 128), (512, 1024, 384), (16384, 2048, 768), (200, 344, 128)])
def test_gemm_swiglu(M, H, K):
    from b200fm import ops
    x = _mk((M, K), 9, 0.5)
    w13 = _mk((2 * H, K), 10, 0.2)
    ab, g = ops.gemm(x, w13, epilogue=ops.EPI_SWIGLU)
    ref = x.float() @ w13.float().t()
    a_ref, b_ref = ref[:, :H].to(torch.bfloat16), ref[:, H:].to(torch.bfloat16)
    torch.testing.assert_close(ab[:, :H].float(), a_ref.float(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(ab[:, H:].float(), b_ref.float(), rtol=1e-2, atol=1e-2)
    # gate computed from the kernel's own (bf16-rounded) a, b exactly as the reference's autocast chain
    g_ref = (torch.nn.functional.silu(ab[:, :H]) * ab[:, H:])
    torch.testing.assert_close(g.float(), g_ref.float(), rtol=2e-2, atol=2e-3)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (4096, 3072, 768)])
def test_gemm_gelu(M, N, K):
    from b200fm import ops
    x, w = _mk((M, K), 11, 0.5), _mk((N, K), 12, 0.2)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(13)).cuda() * 0.1
    pre, act = ops.gemm(x, w, epilogue=ops.EPI_GELU, bias=bias)
    ref = (x.float() @ w.float().t() + bias).to(torch.bfloat16)
    torch.testing.assert_close(pre.float(), ref.float(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(act.float(), torch.nn.functional.gelu(pre.float()).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-3)


def test_gemm_strided_views():
    """q/k/v style column slices of a packed buffer as operands (row stride != K)."""
    from b200fm import ops
    big = _mk((512, 768), 14)
    a = big[:, 256:512]                   # [512, 256] with row stride 768
    b = _mk((384, 256), 15)
    out = ops.gemm(a, b, epilogue=ops.EPI_F32)
    torch.testing.assert_close(out, a.float() @ b.float().t(), rtol=2e-4, atol=5e-3)


@pytest.mark.parametrize("M,N,K", [(768, 768, 16384), (2304, 768, 16384), (4096, 768, 8192), (200, 328, 4096)])
def test_gemm_wgrad_split_k(M, N, K):
    """Weight-gradient shapes (few output tiles, long K) take the split-K path (fp32 atomics into a pre-zeroed output)."""
    from b200fm import ops
    a, b = _mk((K, M), 21, 0.5), _mk((K, N), 22, 0.5)
    out = ops.gemm(a, b, layout=2, epilogue=ops.EPI_F32)
    ref = a.float().t() @ b.float()
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-2 * (K ** 0.5) / 8)
    out2 = ops.gemm(a, b, layout=2, epilogue=ops.EPI_F32, alpha=0.5)
    torch.testing.assert_close(out2, ref * 0.5, rtol=2e-3, atol=2e-2 * (K ** 0.5) / 8)


def test_gemm_degenerate_shapes():
    """A modality without target rows gives M = 0 in the head's forward GEMM and K = 0 in its wgrad (fm.py:589-597): empty
    result / all-zero gradient, no launch."""
    from b200fm import ops
    w = _mk((512, 128), 31)
    x0 = torch.empty(0, 128, dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(x0, w, epilogue=ops.EPI_F32)
    assert out.shape == (0, 512) and out.dtype == torch.float32
    dy0 = torch.empty(0, 512, dtype=torch.bfloat16, device="cuda")
    dw = ops.gemm(dy0, x0, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32)               # [512, 128] from zero rows
    assert dw.shape == (512, 128) and float(dw.abs().sum()) == 0.0


@pytest.mark.parametrize("M", [1, 2, 3, 5, 8])
def test_small_m_weight_streaming_kernel(M):
    """NT problems with <= 8 rows (the linears of the K/V-cached decode loop) run on the weight-streaming kernel (csrc/gemv.cu):
    bf16 / fp32 / SwiGLU epilogues vs fp32 torch on the same bf16 operands, and vs the tcgen05 tile path (option gemv = 0)."""
    from b200fm import lib, ops
    g = torch.Generator(device="cuda").manual_seed(M)
    for N, K in ((2048, 2048), (30000, 768), (768, 5464), (104, 64)):
        x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        ref = x.float() @ w.float().t() + bias
        y32 = ops.gemm(x, w, epilogue=ops.EPI_F32, bias=bias)
        torch.testing.assert_close(y32, ref, rtol=1e-3, atol=1e-3)
        y16 = ops.gemm(x, w, epilogue=ops.EPI_BF16, bias=bias, alpha=0.5)
        torch.testing.assert_close(y16.float(), 0.5 * ref, rtol=1e-2, atol=1e-2)
        res = torch.randn(M, N, device="cuda", generator=g)
        yr = ops.gemm(x, w, epilogue=ops.EPI_RESID, bias=bias, resid=res)                   # resid + bf16(acc + bias)
        torch.testing.assert_close(yr, res + ref.bfloat16().float(), rtol=1e-2, atol=1e-2)
        lib.set_option("gemv", 0)
        try:
            t32 = ops.gemm(x, w, epilogue=ops.EPI_F32, bias=bias)
        finally:
            lib.set_option("gemv", 1)
        torch.testing.assert_close(y32, t32, rtol=1e-3, atol=1e-3)
    H, K = 5464, 2048
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w13 = (torch.randn(2 * H, K, device="cuda", generator=g) * 0.05).bfloat16()
    ab, gate = ops.gemm(x, w13, epilogue=ops.EPI_SWIGLU)
    a, b = x.float() @ w13[:H].float().t(), x.float() @ w13[H:].float().t()
    torch.testing.assert_close(ab.float(), torch.cat([a, b], 1), rtol=1e-2, atol=1e-2)
    ar, br = a.bfloat16().float(), b.bfloat16().float()
    torch.testing.assert_close(gate.float(), torch.nn.functional.silu(ar).bfloat16().float() * br, rtol=2e-2, atol=2e-2)
    lib.set_option("gemv", 0)
    try:
        ab2, gate2 = ops.gemm(x, w13, epilogue=ops.EPI_SWIGLU)
    finally:
        lib.set_option("gemv", 1)
    torch.testing.assert_close(gate.float(), gate2.float(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("n,V,D", [(1, 1000, 384), (130, 1000, 384), (517, 16384, 768), (3000, 30000, 768), (257, 133, 64), (4096, 8192, 1024)])
def test_fused_head_cross_entropy(n, V, D):
    """b200fm_head_ce (logits stay in the accumulator: statistics epilogue -> row reduction -> gradient epilogue) vs torch fp32
    cross-entropy on the same bf16 operands, and vs the two-kernel path (fp32 logits GEMM + cross-entropy kernel).  Covers V not a
    multiple of the 256-column tile / 128-column half and a device-side row count (rows behind it: loss 0, gradient rows zero up to
    the next multiple of 64, later rows untouched)."""
    from b200fm import ops
    g = torch.Generator(device="cuda").manual_seed(n + V)
    h = torch.randn(n, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(V, D, device="cuda", generator=g) * (3.0 / D ** 0.5)).bfloat16()       # logits ~ N(0, 9): peaked softmax rows
    t = torch.randint(0, V, (n,), device="cuda", generator=g)
    logits = h.float() @ w.float().t()
    ref_loss = torch.nn.functional.cross_entropy(logits, t, reduction="none")
    ref_grad = torch.softmax(logits, -1)
    ref_grad[torch.arange(n), t] -= 1.0
    loss, dl = ops.head_ce(h, w, V, t)
    torch.testing.assert_close(loss, ref_loss, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(dl.float(), ref_grad, rtol=1e-2, atol=2e-4)
    l2, d2 = ops.cross_entropy(ops.gemm(h, w, epilogue=ops.EPI_F32), t)
    torch.testing.assert_close(loss, l2, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(dl.float(), d2.float(), rtol=1e-2, atol=1e-5)
    for m in sorted({0, 1, n // 2, n}):
        nd = torch.tensor([m], dtype=torch.int32, device="cuda")
        loss, dl = ops.head_ce(h, w, V, t, nd)
        torch.testing.assert_close(loss[:m], ref_loss[:m], rtol=1e-4, atol=2e-4)
        assert float(loss[m:].abs().sum()) == 0.0
        torch.testing.assert_close(dl[:m].float(), ref_grad[:m], rtol=1e-2, atol=2e-4)
        m64 = min((m + 63) // 64 * 64, n)
        assert float(dl[m:m64].float().abs().sum()) == 0.0


def test_kv_append_writes_one_cache_row():
    """b200fm_kv_append: cache[b, *pos, col0:col0+w] = src[b] with the position in device memory; everything else untouched; an
    out-of-range position is a no-op."""
    from b200fm import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    for B, L, W, w, col0 in ((2, 7, 256, 256, 0), (3, 5, 512, 256, 256), (1, 4, 24, 7, 3)):
        cache = torch.randn(B, L, W, device="cuda", generator=g).bfloat16()
        src_full = torch.randn(B, w + 16, device="cuda", generator=g).bfloat16()
        src = src_full[:, 8:8 + w] if w % 8 == 0 else src_full[:, 1:1 + w]
        for pos in (0, L - 1, L):
            ref = cache.clone()
            if pos < L:
                ref[:, pos, col0:col0 + w] = src
            out = ops.kv_append(src, cache.clone(), torch.tensor([pos], device="cuda"), col0)
            assert torch.equal(out, ref)


@pytest.mark.parametrize("terms,tol", [(3, 5e-5), (6, 5e-5)])      # 6 terms: the accumulator's own rounding (K up to 18432 products) is what is left
def test_fp32_faithful_linear_and_attention(terms, tol):
    """b200fm.functional.linear_f32 (bf16 limb products on the tcgen05 GEMM; vectorised and scalar limb-split kernels) and
    ops.attention_f32 against fp64: relative error of an fp32 matmul, not of a bf16 one (which is ~4e-3)."""
    from b200fm import functional as BF
    from b200fm import ops
    g = torch.Generator(device="cuda").manual_seed(terms)
    for M, K, N in ((300, 768, 512), (64, 104, 72), (1000, 3072, 768)):
        x = torch.randn(M, K, device="cuda", generator=g)
        if K == 104:                                            # a misaligned row view: the scalar limb-split kernel
            x = torch.randn(M, K + 3, device="cuda", generator=g)[:, 1:1 + K]
        w = torch.randn(N, K, device="cuda", generator=g) * 0.1
        b = torch.randn(N, device="cuda", generator=g)
        with BF.precise(terms):
            y = BF.linear_f32(x, w, b, cache=False)
        ref = x.double() @ w.double().t() + b.double()
        err = float((y.double() - ref).norm() / ref.norm())
        bf16_err = float(((x.bfloat16().double() @ w.bfloat16().double().t() + b.double()) - ref).norm() / ref.norm())
        assert err <= tol and err < bf16_err / 50, (M, K, N, err, bf16_err)
    B, H, N, D = 3, 4, 197, 64
    qkv = torch.randn(B * N, 3 * H * D, device="cuda", generator=g)
    C = H * D
    o = ops.attention_f32(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, None, D ** -0.5)
    q, k, v = (t.view(B, N, H, D).permute(0, 2, 1, 3).double() for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * D ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(B * N, C)
    assert float((o.double() - ref).norm() / ref.norm()) <= 2e-6
