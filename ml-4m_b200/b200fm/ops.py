"""Tensor-level wrappers over the C ABI (device pointers + current CUDA stream).  CUDA tensors only."""
import torch

from . import lib

LAYOUT_NT, LAYOUT_NN, LAYOUT_TN = 0, 1, 2
EPI_BF16, EPI_F32, EPI_RESID, EPI_SWIGLU, EPI_GELU = 0, 1, 2, 3, 4


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise lib.B200FMError("b200fm ops need CUDA tensors (there is no CPU fallback)")


def _rowmajor2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} / {t.stride()}")
    return t.stride(0)


def gemm(a, b, layout=LAYOUT_NT, epilogue=EPI_BF16, out=None, out1=None, bias=None, resid=None, alpha=1.0, alpha_dev=None,
         n_out=None):
    """C = op(A, B) with a fused epilogue; see include/b200fm.h.  a, b bf16 2-D (row stride free, inner stride 1).
    Returns out (and out1 for SWIGLU / GELU)."""
    _need_cuda(a, b, out, out1, bias, resid, alpha_dev)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    lda, ldb = _rowmajor2d(a, "a"), _rowmajor2d(b, "b")
    if layout == LAYOUT_NT:
        M, K = a.shape
        N = b.shape[0] if epilogue != EPI_SWIGLU else b.shape[0] // 2
        assert b.shape[1] == K
    elif layout == LAYOUT_NN:
        M, K = a.shape
        N = b.shape[1]
        assert b.shape[0] == K
    else:
        K, M = a.shape
        N = b.shape[1]
        assert b.shape[0] == K
    if n_out is not None:
        N = n_out
    dev = a.device
    if out is None:
        if epilogue in (EPI_F32, EPI_RESID):
            out = torch.empty(M, N, device=dev, dtype=torch.float32)
        elif epilogue == EPI_SWIGLU:
            out = torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16)
        else:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if out1 is None and epilogue in (EPI_SWIGLU, EPI_GELU):
        out1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ld0 = _rowmajor2d(out, "out")
    ld1 = _rowmajor2d(out1, "out1") if out1 is not None else 0
    ldr = _rowmajor2d(resid, "resid") if resid is not None else 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if resid is not None:
        assert resid.dtype == torch.float32
    lib.call("b200fm_gemm_bf16", layout, epilogue, M, N, K, _ptr(a), lda, _ptr(b), ldb, _ptr(out), ld0, _ptr(out1), ld1,
             _ptr(bias), _ptr(resid), ldr, float(alpha), _ptr(alpha_dev), _stream())
    return (out, out1) if epilogue in (EPI_SWIGLU, EPI_GELU) else out


def layernorm_fwd(x, gamma, beta, eps, out_bf16=True, save_stats=True):
    """x fp32 [..., D] -> (y [..., D] bf16|fp32, mean [rows], rstd [rows])."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous()
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    lib.call("b200fm_layernorm_fwd", _ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), int(out_bf16), _ptr(mean), _ptr(rstd), rows, D,
             float(eps), _stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, want_bf16=False, dgamma=None, dbeta=None):
    """Returns (dx fp32, dx_bf16 | None).  dgamma / dbeta (fp32 [D]) are accumulated into when given."""
    _need_cuda(dy, x, gamma, mean, rstd, dres, dgamma, dbeta)
    assert dy.is_contiguous() and x.is_contiguous() and x.dtype == torch.float32
    D = x.shape[-1]
    rows = x.numel() // D
    dx = torch.empty_like(x)
    dxb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    if dres is not None:
        assert dres.dtype == torch.float32 and dres.is_contiguous()
    lib.call("b200fm_layernorm_bwd", _ptr(dy), int(dy.dtype == torch.bfloat16), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd),
             _ptr(dres), _ptr(dx), _ptr(dxb), _ptr(dgamma), _ptr(dbeta), rows, D, _stream())
    return dx, dxb


def vq_argmax(z, codebook, cosine=True, want_quant=False):
    """z fp32 [n, d], codebook fp32 [K, d] -> int64 [n] (and fp32 [n, d] codebook rows)."""
    _need_cuda(z, codebook)
    assert z.dtype == torch.float32 and codebook.dtype == torch.float32 and z.is_contiguous() and codebook.is_contiguous()
    n, d = z.shape
    K = codebook.shape[0]
    idx = torch.empty(n, device=z.device, dtype=torch.int64)
    quant = torch.empty(n, d, device=z.device, dtype=torch.float32) if want_quant else None
    lib.call("b200fm_vq_argmax", _ptr(z), _ptr(codebook), _ptr(idx), _ptr(quant), n, K, d, int(cosine), _stream())
    return (idx, quant) if want_quant else idx


def _mask_args(mask, B, Nq, Nk):
    """mask: None or bool/uint8 broadcastable [B, 1|Nq, Nk] (True = masked) -> (tensor kept alive, ptr, b_stride, q_stride)."""
    if mask is None:
        return None, 0, 0, 0
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    assert mask.dtype == torch.uint8 and mask.dim() == 3 and mask.shape[2] == Nk and mask.shape[0] in (1, B) and mask.shape[1] in (1, Nq)
    if mask.stride(2) != 1:
        mask = mask.contiguous()
    bs = 0 if mask.shape[0] == 1 else mask.stride(0)
    qs = 0 if mask.shape[1] == 1 else mask.stride(1)
    return mask, mask.data_ptr(), bs, qs


def attention_fwd(q, k, v, B, H, Nq, Nk, mask=None, scale=None):
    """q [B*Nq, >=H*64], k/v [B*Nk, >=H*64] bf16 views with unit inner stride (may be column slices of a packed qkv).
    Returns (out bf16 [B*Nq, H*64], stats fp32 [B, H, Nq, 2])."""
    _need_cuda(q, k, v, mask)
    for t in (q, k, v):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    scale = 64 ** -0.5 if scale is None else scale
    out = torch.empty(B * Nq, H * 64, device=q.device, dtype=torch.bfloat16)
    stats = torch.empty(B, H, Nq, 2, device=q.device, dtype=torch.float32)
    mk, mp, mbs, mqs = _mask_args(mask, B, Nq, Nk)
    lib.call("b200fm_attention_fwd", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), mp, mbs, mqs, _ptr(out),
             out.stride(0), _ptr(stats), B, H, Nq, Nk, float(scale), _stream())
    return out, stats


def attention_bwd(q, k, v, out, dout, stats, B, H, Nq, Nk, mask=None, scale=None, dq=None, dk=None, dv=None):
    """Gradients w.r.t. q, k, v (bf16).  dq/dk/dv may be pre-allocated views (e.g. column slices of a packed dqkv)."""
    _need_cuda(q, k, v, out, dout, stats, mask)
    scale = 64 ** -0.5 if scale is None else scale
    dev = q.device
    dq = torch.empty(B * Nq, H * 64, device=dev, dtype=torch.bfloat16) if dq is None else dq
    dk = torch.empty(B * Nk, H * 64, device=dev, dtype=torch.bfloat16) if dk is None else dk
    dv = torch.empty(B * Nk, H * 64, device=dev, dtype=torch.bfloat16) if dv is None else dv
    mk, mp, mbs, mqs = _mask_args(mask, B, Nq, Nk)
    lib.call("b200fm_attention_bwd", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), mp, mbs, mqs, _ptr(out),
             out.stride(0), _ptr(dout), dout.stride(0), _ptr(stats), _ptr(dq), dq.stride(0), _ptr(dk), dk.stride(0), _ptr(dv),
             dv.stride(0), B, H, Nq, Nk, float(scale), _stream())
    return dq, dk, dv
