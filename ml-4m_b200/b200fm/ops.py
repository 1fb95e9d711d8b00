"""Tensor-level wrappers over the C ABI (device pointers + current CUDA stream).  CUDA tensors only."""
import torch

from . import lib

LAYOUT_NT, LAYOUT_NN, LAYOUT_TN = 0, 1, 2
EPI_BF16, EPI_F32, EPI_RESID, EPI_SWIGLU, EPI_GELU, EPI_TANH = 0, 1, 2, 3, 4, 5


PROFILE = None     # bench.py sets this to a list: (start_event, end_event, flops) per GEMM launch
RECORD = None      # bench.py sets this to a list: the keyword arguments of every GEMM launch of a step (tensors kept alive), to replay
                   # exactly that launch sequence back to back as a CUDA graph (GEMM-family throughput under the step's launch conditions)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
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
         n_out=None, dyn=None):
    """C = op(A, B) with a fused epilogue; see include/b200fm.h.  a, b bf16 2-D (row stride free, inner stride 1).
    Returns out (and out1 for SWIGLU / GELU / TANH).  (Hot wrapper: argument validation lives in the C entry point.)
    dyn: optional device int32 [1] problem size (rows of NT / NN, contraction length of TN), see b200fm_gemm_bf16_dyn."""
    if not a.is_cuda:
        raise lib.B200FMError("b200fm ops need CUDA tensors (there is no CPU fallback)")
    if a.dtype is not torch.bfloat16 or b.dtype is not torch.bfloat16 or a.stride(1) != 1 or b.stride(1) != 1:
        raise ValueError("gemm: operands must be bf16 2-D tensors with unit inner stride")
    ash, bsh = a.shape, b.shape
    if layout == LAYOUT_NT:
        M, K = ash
        N = bsh[0] if epilogue != EPI_SWIGLU else bsh[0] // 2
    elif layout == LAYOUT_NN:
        M, K = ash
        N = bsh[1]
    else:
        K, M = ash
        N = bsh[1]
    if n_out is not None:
        N = n_out
    two = epilogue in (EPI_SWIGLU, EPI_GELU, EPI_TANH)
    if out is None:
        dev = a.device
        if epilogue == EPI_F32 or epilogue == EPI_RESID:
            out = torch.empty((M, N), device=dev, dtype=torch.float32)
        elif epilogue == EPI_SWIGLU:
            out = torch.empty((M, 2 * N), device=dev, dtype=torch.bfloat16)
        else:
            out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    if two and out1 is None:
        out1 = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    if M == 0 or N == 0 or K == 0:
        # degenerate problems (a modality without target rows: fm.py:589 `y[mod_mask == id]` can be empty): nothing to launch;
        # an empty contraction dimension leaves the (bias-free) output at zero, like the reference's matmul
        if K == 0 and M and N:
            if epilogue == EPI_RESID and resid is not None:
                out.copy_(resid)
            else:
                out.zero_()
            if out1 is not None:
                out1.zero_()
        return (out, out1) if two else out
    if RECORD is not None:
        RECORD.append((dict(a=a, b=b, layout=layout, epilogue=epilogue, out=out, out1=out1, bias=bias, resid=resid, alpha=alpha,
                            alpha_dev=alpha_dev, n_out=n_out, dyn=dyn), 2.0 * M * K * (2 * N if epilogue == EPI_SWIGLU else N)))
    if PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    if dyn is None:
        lib.call("b200fm_gemm_bf16", layout, epilogue, M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                 out.stride(0), 0 if out1 is None else out1.data_ptr(), 0 if out1 is None else out1.stride(0),
                 0 if bias is None else bias.data_ptr(), 0 if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0),
                 float(alpha), 0 if alpha_dev is None else alpha_dev.data_ptr(), _stream())
    else:
        lib.call("b200fm_gemm_bf16_dyn", layout, epilogue, M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                 out.stride(0), 0 if out1 is None else out1.data_ptr(), 0 if out1 is None else out1.stride(0),
                 0 if bias is None else bias.data_ptr(), 0 if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0),
                 float(alpha), 0 if alpha_dev is None else alpha_dev.data_ptr(), dyn.data_ptr(), 2 if layout == LAYOUT_TN else 1, _stream())
    if PROFILE is not None:
        ev1.record()
        PROFILE.append((ev0, ev1, 2.0 * M * K * (2 * N if epilogue == EPI_SWIGLU else N), (layout, epilogue, M, N, K)))
    return (out, out1) if two else out


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


def add_layernorm_fwd(x, add, gamma, beta, eps, out_bf16=True):
    """s = x + add (fp32 stream + bf16 pending branch; add may be None -> s is x), y = LN(s).  -> (s, y, mean, rstd)."""
    _need_cuda(x, add, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous()
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    if add is None:
        s = x
        lib.call("b200fm_layernorm_fwd", _ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), int(out_bf16), _ptr(mean), _ptr(rstd), rows, D,
                 float(eps), _stream())
    else:
        assert add.dtype == torch.bfloat16 and add.is_contiguous() and add.numel() == x.numel()
        s = torch.empty_like(x)
        lib.call("b200fm_add_layernorm_fwd", _ptr(x), _ptr(add), _ptr(s), _ptr(gamma), _ptr(beta), _ptr(y), int(out_bf16), _ptr(mean),
                 _ptr(rstd), rows, D, float(eps), _stream())
    return s, y, mean, rstd


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


def headnorm_fwd(x, H, gamma, beta, eps):
    """Per-head LayerNorm over head_dim 64: x bf16 [R, >=H*64] view (unit inner stride) -> (y bf16 [R, H*64], stats [R, H, 2])."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] == H * 64
    R = x.shape[0]
    y = torch.empty(R, H * 64, device=x.device, dtype=torch.bfloat16)
    stats = torch.empty(R, H, 2, device=x.device, dtype=torch.float32)
    lib.call("b200fm_headnorm_fwd", _ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), _ptr(y), y.stride(0), _ptr(stats), R, H, float(eps),
             _stream())
    return y, stats


def headnorm_bwd(dy, x, gamma, stats, H, dgamma=None, dbeta=None):
    """dx bf16 [R, H*64]; dgamma / dbeta fp32 [64] are accumulated into when given."""
    _need_cuda(dy, x, gamma, stats)
    assert dy.dtype == torch.bfloat16 and dy.stride(1) == 1 and x.stride(1) == 1
    R = x.shape[0]
    dx = torch.empty(R, H * 64, device=x.device, dtype=torch.bfloat16)
    lib.call("b200fm_headnorm_bwd", _ptr(dy), dy.stride(0), _ptr(x), x.stride(0), _ptr(gamma), _ptr(stats), _ptr(dx), dx.stride(0),
             _ptr(dgamma), _ptr(dbeta), R, H, _stream())
    return dx


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


def attention_decode(q, k, v, B, H, Nk, mask=None, scale=None):
    """One query row per sequence: q bf16 [B, >=H*64], k / v bf16 [B*Nk, >=H*64] row views, mask bool / uint8 [B|1, 1, Nk] (True =
    masked) -> out bf16 [B, H*64] (b200fm_attention_decode: plain-load kernel, exact fp32 softmax)."""
    _need_cuda(q, k, v, mask)
    for t in (q, k, v):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    scale = 64 ** -0.5 if scale is None else scale
    out = torch.empty(B, H * 64, device=q.device, dtype=torch.bfloat16)
    mk, mp, mbs, _ = _mask_args(mask, B, 1, Nk)
    lib.call("b200fm_attention_decode", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), mp, mbs, _ptr(out), out.stride(0),
             B, H, Nk, float(scale), _stream())
    return out


def attention_bwd(q, k, v, out, dout, stats, B, H, Nq, Nk, mask=None, scale=None, dq=None, dk=None, dv=None):
    """Gradients w.r.t. q, k, v (bf16).  dq/dk/dv may be pre-allocated views (e.g. column slices of a packed dqkv)."""
    _need_cuda(q, k, v, out, dout, stats, mask)
    scale = 64 ** -0.5 if scale is None else scale
    dev = q.device
    dq = torch.empty(B * Nq, H * 64, device=dev, dtype=torch.bfloat16) if dq is None else dq
    dk = torch.empty(B * Nk, H * 64, device=dev, dtype=torch.bfloat16) if dk is None else dk
    dv = torch.empty(B * Nk, H * 64, device=dev, dtype=torch.bfloat16) if dv is None else dv
    mk, mp, mbs, mqs = _mask_args(mask, B, Nq, Nk)
    dsum = torch.empty(B, H, Nq, device=dev, dtype=torch.float32)
    lib.call("b200fm_attention_bwd", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), mp, mbs, mqs, _ptr(out),
             out.stride(0), _ptr(dout), dout.stride(0), _ptr(stats), _ptr(dsum), _ptr(dq), dq.stride(0), _ptr(dk), dk.stride(0),
             _ptr(dv), dv.stride(0), B, H, Nq, Nk, float(scale), _stream())
    return dq, dk, dv


# ----------------------------------------------------------------------------------------------------------------------
# element-wise / row-wise
# ----------------------------------------------------------------------------------------------------------------------

def swiglu_bwd(ab, dg):
    _need_cuda(ab, dg)
    R, H2 = ab.shape
    H = H2 // 2
    dab = torch.empty_like(ab)
    lib.call("b200fm_swiglu_bwd", _ptr(ab), ab.stride(0), _ptr(dg), dg.stride(0), _ptr(dab), dab.stride(0), R, H, _stream())
    return dab


def act_bwd(pre, dact, act):
    _need_cuda(pre, dact)
    assert pre.is_contiguous() and dact.is_contiguous()
    out = torch.empty_like(pre)
    lib.call("b200fm_act_bwd", {"gelu": 0, "tanh": 1}[act], _ptr(pre), _ptr(dact), _ptr(out), pre.numel(), _stream())
    return out


def cross_entropy(logits, targets, want_grad=True):
    """logits fp32 [n, V], targets int64 [n] -> (loss_rows fp32 [n], dlogits bf16 [n, V] = softmax - onehot | None)."""
    _need_cuda(logits, targets)
    n, V = logits.shape
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and targets.dtype == torch.int64 and targets.is_contiguous()
    loss = torch.empty(n, device=logits.device, dtype=torch.float32)
    Vp = (V + 7) // 8 * 8
    dl = torch.empty(n, Vp, device=logits.device, dtype=torch.bfloat16)[:, :V] if want_grad else None
    lib.call("b200fm_cross_entropy", _ptr(logits), logits.stride(0), _ptr(targets), _ptr(loss), _ptr(dl), dl.stride(0) if want_grad else 0,
             n, V, _stream())
    return loss, dl


def cross_entropy_dyn(logits, targets, n_dev, want_grad=True):
    """cross_entropy over the first *n_dev rows (device int32 [1]); later rows: loss 0, dlogits zero up to the next multiple of 64."""
    n, V = logits.shape
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and targets.dtype == torch.int64 and targets.is_contiguous()
    loss = torch.empty(n, device=logits.device, dtype=torch.float32)
    Vp = (V + 7) // 8 * 8
    dl = torch.empty(n, Vp, device=logits.device, dtype=torch.bfloat16)[:, :V] if want_grad else None
    lib.call("b200fm_cross_entropy_dyn", _ptr(logits), logits.stride(0), _ptr(targets), _ptr(loss), _ptr(dl), dl.stride(0) if want_grad else 0,
             n, V, n_dev.data_ptr(), _stream())
    return loss, dl


def sample_top_p(logits, top_p, temperature, u):
    """One nucleus-sampled token per row: logits fp32 [R, V] (unit inner stride), u fp32 [R] uniform in [0, 1) -> int64 [R]."""
    _need_cuda(logits, u)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1 and u.dtype == torch.float32 and u.is_contiguous()
    R, V = logits.shape
    out = torch.empty(R, device=logits.device, dtype=torch.int64)
    lib.call("b200fm_sample_top_p", _ptr(logits), logits.stride(0), R, V, float(top_p), float(temperature), _ptr(u), _ptr(out), _stream())
    return out


def kv_append(src, cache, pos_dev, col0=0):
    """cache bf16 [B, L, W]; cache[b, pos, col0 : col0 + src.shape[1]] = src[b] with pos = *pos_dev (device int64 [1])."""
    _need_cuda(src, cache, pos_dev)
    B, L, W = cache.shape
    assert src.dtype == torch.bfloat16 and cache.dtype == torch.bfloat16 and cache.is_contiguous() and src.stride(1) == 1 and src.shape[0] == B
    assert pos_dev.dtype == torch.int64
    lib.call("b200fm_kv_append", _ptr(src), src.stride(0), _ptr(cache), L, W, _ptr(pos_dev), B, src.shape[1], col0, _stream())
    return cache


def head_ce(h, wb, V, targets, n_dev=None, want_grad=True):
    """Fused masked-token head: loss_rows fp32 [n] = CE(h W^T, targets) and dlogits bf16 [n, V] = softmax - onehot, the fp32 logits
    never leave the GEMM's accumulator (b200fm_head_ce: statistics pass, row reduction, gradient pass).  h bf16 [n, D]; wb bf16
    [>=V, D]; n_dev: optional device row count (rows behind it: loss 0, dlogits zero up to the next multiple of 64)."""
    _need_cuda(h, wb, targets)
    n, D = h.shape
    assert h.dtype == torch.bfloat16 and wb.dtype == torch.bfloat16 and h.stride(1) == 1 and wb.stride(1) == 1 and wb.shape[0] >= V
    assert targets.dtype == torch.int64 and targets.is_contiguous() and targets.numel() >= n
    slots = 2 * ((V + 255) // 256)                                                # == b200fm_head_ce_ws_slots(V) float2 partials per row
    ws = torch.empty(n, 2 * slots, device=h.device, dtype=torch.float32)
    aux = torch.empty(3, n, device=h.device, dtype=torch.float32)                 # target logit, lse, loss
    Vp = (V + 7) // 8 * 8
    dl = torch.empty(n, Vp, device=h.device, dtype=torch.bfloat16)[:, :V] if want_grad else None
    if RECORD is not None:        # two tcgen05 GEMM passes over the same operands (statistics, gradient)
        RECORD.append((dict(_fn="head_ce", h=h, wb=wb, V=V, targets=targets, n_dev=n_dev, want_grad=want_grad), (4.0 if want_grad else 2.0) * n * V * D))
    if PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    lib.call("b200fm_head_ce", _ptr(h), h.stride(0), _ptr(wb), wb.stride(0), _ptr(targets), n_dev.data_ptr() if n_dev is not None else None,
             n, V, D, _ptr(ws), _ptr(aux[0]), _ptr(aux[1]), _ptr(aux[2]), _ptr(dl), dl.stride(0) if want_grad else 0, _stream())
    if PROFILE is not None:
        ev1.record()
        PROFILE.append((ev0, ev1, (4.0 if want_grad else 2.0) * n * V * D, ("head_ce", n, V, D)))
    return aux[2], dl


def masked_mean(x, n_dev):
    """(sum(x[:n]) / max(n, 1), 1 / max(n, 1)) as 0-d device tensors, n = *n_dev."""
    out = torch.empty(2, device=x.device, dtype=torch.float32)
    lib.call("b200fm_masked_mean", _ptr(x), n_dev.data_ptr(), x.numel(), out.data_ptr(), out.data_ptr() + 4, _stream())
    return out[0], out[1]


def colsum_bf16(x, out=None):
    _need_cuda(x, out)
    R, N = x.shape
    if out is None:
        out = torch.zeros(N, device=x.device, dtype=torch.float32)
    lib.call("b200fm_colsum_bf16", _ptr(x), x.stride(0), _ptr(out), R, N, _stream())
    return out


def cast_bf16(x, out=None):
    _need_cuda(x, out)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lib.call("b200fm_cast_f32_bf16", _ptr(x), _ptr(out), x.numel(), _stream())
    return out


def patchify(img, patch):
    _need_cuda(img)
    assert img.dtype == torch.float32 and img.is_contiguous()
    B, C, H, W = img.shape
    if H % patch or W % patch:
        raise AssertionError(f"Image sizes {H}x{W} must be divisible by patch sizes {patch}x{patch}")
    out = torch.empty(B * (H // patch) * (W // patch), patch * patch * C, device=img.device, dtype=torch.bfloat16)
    lib.call("b200fm_patchify", _ptr(img), _ptr(out), B, C, H, W, patch, _stream())
    return out


def split_limbs(x, terms, role, out=None):
    """fp32 [rows, K] (unit inner stride) -> bf16 [rows, terms*K] limb layout of a GEMM operand (role 0 = A, 1 = B)."""
    _need_cuda(x, out)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, K = x.shape
    if out is None:
        out = torch.empty(rows, terms * K, device=x.device, dtype=torch.bfloat16)
    lib.call("b200fm_split_limbs", _ptr(x), x.stride(0), _ptr(out), rows, K, terms, role, _stream())
    return out


def attention_f32(q, k, v, B, H, Nq, Nk, mask=None, scale=None):
    """fp32 attention (head_dim 64) on row views q [B*Nq, >=H*64], k / v [B*Nk, >=H*64]; returns fp32 [B*Nq, H*64]."""
    _need_cuda(q, k, v, mask)
    for t in (q, k, v):
        assert t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1
    scale = 64 ** -0.5 if scale is None else scale
    out = torch.empty(B * Nq, H * 64, device=q.device, dtype=torch.float32)
    mk, mp, mbs, mqs = _mask_args(mask, B, Nq, Nk)
    lib.call("b200fm_attention_f32", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), mp, mbs, mqs, _ptr(out), out.stride(0),
             B, H, Nq, Nk, float(scale), _stream())
    return out


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # fourm/utils/data_constants.py


def patchify_u8(img, patch, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """uint8 RGB [B,3,H,W] -> normalised bf16 patches (same layout as patchify)."""
    import ctypes
    _need_cuda(img)
    assert img.dtype == torch.uint8 and img.is_contiguous()
    B, C, H, W = img.shape
    if H % patch or W % patch:
        raise AssertionError(f"Image sizes {H}x{W} must be divisible by patch sizes {patch}x{patch}")
    out = torch.empty(B * (H // patch) * (W // patch), patch * patch * C, device=img.device, dtype=torch.bfloat16)
    m3, s3 = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    lib.call("b200fm_patchify_u8", _ptr(img), _ptr(out), B, C, H, W, patch, m3, s3, _stream())
    return out


def mask_images(noise, in_budget, tgt_budget=None):
    """noise fp32 [..., L], budgets int32 [...] -> (input_mask bool [..., L], target_mask bool, decoder_attention_mask int32)."""
    _need_cuda(noise, in_budget, tgt_budget)
    assert noise.dtype == torch.float32 and noise.is_contiguous() and in_budget.dtype == torch.int32 and in_budget.is_contiguous()
    L = noise.shape[-1]
    rows = noise.numel() // L
    im = torch.empty(noise.shape, device=noise.device, dtype=torch.bool)
    tm = torch.empty(noise.shape, device=noise.device, dtype=torch.bool)
    dam = torch.empty(noise.shape, device=noise.device, dtype=torch.int32)
    lib.call("b200fm_mask_images", _ptr(noise), _ptr(in_budget), _ptr(tgt_budget), _ptr(im), _ptr(tm), _ptr(dam), rows, L, _stream())
    return im, tm, dam


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, shadow=None):
    _need_cuda(p, g, m, v, shadow)
    lib.call("b200fm_adamw", _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(shadow), p.numel(), float(lr), float(beta1), float(beta2),
             float(eps), float(weight_decay), int(step), float(grad_scale), _stream())


# ----------------------------------------------------------------------------------------------------------------------
# selection plan / embedding gather-scatter
# ----------------------------------------------------------------------------------------------------------------------

class SelectionPlan:
    """Device-side result of b200fm_select_plan for one side (encoder / decoder) of one forward call."""
    __slots__ = ("segs", "n_seg", "decoder", "B", "n_keep", "src_seg", "src_pos", "pos_id", "pad_mask", "mod_mask", "mod_raw",
                 "target_ids", "dam", "keepalive")


def make_segments(seg_dicts):
    """seg_dicts: list of dicts with tensors/ints per include/b200fm.h b200fm_segment.  Returns (ctypes array, keepalive list)."""
    n = len(seg_dicts)
    if not 1 <= n <= lib.MAX_SEGMENTS:
        raise ValueError(f"{n} modalities: supported range is 1..{lib.MAX_SEGMENTS}")
    arr = (lib.Segment * n)()
    keep = []
    for i, d in enumerate(seg_dicts):
        s = arr[i]
        for name in ("mask", "ids", "dam", "token_emb", "pos_emb", "mod_emb", "x_rows", "d_token_emb", "d_mod_emb", "dx_rows", "d_pos_emb"):
            t = d.get(name)
            if t is not None:
                _need_cuda(t)
                keep.append(t)
            setattr(s, name, _ptr(t))
        s.padding_idx = d.get("padding_idx", -1) if d.get("padding_idx") is not None else -1
        s.L, s.kind, s.mod_id = int(d["L"]), int(d["kind"]), int(d["mod_id"])
        s.max_length = int(d.get("max_length", 0))
        ids = d.get("ids")
        s.ids_is_i64 = int(ids is not None and ids.dtype == torch.int64)
        s.reserved = 0
    return arr, keep


MODE_DECODER, MODE_IDENTITY, MODE_NO_SUM = 1, 2, 4


def select_plan(seg_dicts, mode, B, n_keep, device, order_dev=None):
    """mode: bit flags MODE_DECODER | MODE_IDENTITY | MODE_NO_SUM (include/b200fm.h).  order_dev: optional device int32 [n_seg]
    concatenation order (segment indices), see b200fm_select_plan_ordered."""
    import ctypes
    arr, keep = make_segments(seg_dicts)
    p = SelectionPlan()
    decoder = mode & MODE_DECODER
    p.segs, p.n_seg, p.decoder, p.B, p.n_keep, p.keepalive = seg_dicts, len(seg_dicts), int(mode), B, n_keep, keep
    i32 = dict(device=device, dtype=torch.int32)
    p.src_seg = torch.empty(B, n_keep, **i32)
    p.src_pos = torch.empty(B, n_keep, **i32)
    p.pos_id = torch.empty(B, n_keep, **i32)
    p.pad_mask = torch.empty(B, n_keep, device=device, dtype=torch.bool)
    p.mod_mask = torch.empty(B, n_keep, device=device, dtype=torch.int16)
    p.mod_raw = torch.empty(B, n_keep, device=device, dtype=torch.int16)
    p.target_ids = torch.empty(B, n_keep, device=device, dtype=torch.int64) if decoder else None
    p.dam = torch.empty(B, n_keep, **i32) if decoder else None
    lib.call("b200fm_select_plan_ordered", ctypes.addressof(arr), p.n_seg, p.decoder, B, n_keep, _ptr(p.src_seg), _ptr(p.src_pos),
             _ptr(p.pos_id), _ptr(p.pad_mask), _ptr(p.mod_mask), _ptr(p.mod_raw), _ptr(p.target_ids), _ptr(p.dam), _ptr(order_dev), _stream())
    return p


def decoder_attention_mask(dam, mod_raw, causal=False, sep=True):
    _need_cuda(dam, mod_raw)
    B, M = dam.shape
    out = torch.empty(B, M, M, device=dam.device, dtype=torch.bool)
    lib.call("b200fm_decoder_attention_mask", _ptr(dam), _ptr(mod_raw), _ptr(out), B, M, int(causal), int(sep), _stream())
    return out


def embed_rows(plan, seg_dicts, mask_token, D, want_emb):
    import ctypes
    arr, keep = make_segments(seg_dicts)
    dev = plan.src_seg.device
    x0 = torch.empty(plan.B, plan.n_keep, D, device=dev, dtype=torch.float32)
    emb = torch.empty(plan.B, plan.n_keep, D, device=dev, dtype=torch.float32) if want_emb else None
    lib.call("b200fm_embed_rows", ctypes.addressof(arr), len(seg_dicts), plan.decoder, _ptr(plan.src_seg), _ptr(plan.src_pos),
             _ptr(plan.pos_id), _ptr(plan.pad_mask), _ptr(mask_token), _ptr(x0), _ptr(emb), plan.B, plan.n_keep, D, _stream())
    return x0, emb


def embed_rows_bwd(plan, seg_dicts, dx0, demb, d_mask_token, D):
    import ctypes
    arr, keep = make_segments(seg_dicts)
    assert dx0.is_contiguous() and dx0.dtype == torch.float32 and (demb is None or (demb.is_contiguous() and demb.dtype == torch.float32))
    lib.call("b200fm_embed_rows_bwd", ctypes.addressof(arr), len(seg_dicts), plan.decoder, _ptr(plan.src_seg), _ptr(plan.src_pos),
             _ptr(plan.pos_id), _ptr(plan.pad_mask), _ptr(dx0), _ptr(demb), _ptr(d_mask_token), plan.B, plan.n_keep, D, _stream())


def head_rows(mod_mask, mod_ids_dev):
    """-> (rows int32 [n_mods, n_rows], counts int32 [n_mods]) on device."""
    _need_cuda(mod_mask, mod_ids_dev)
    n_rows = mod_mask.numel()
    n_mods = mod_ids_dev.numel()
    rows = torch.empty(n_mods, n_rows, device=mod_mask.device, dtype=torch.int32)
    counts = torch.empty(n_mods, device=mod_mask.device, dtype=torch.int32)
    lib.call("b200fm_head_rows", _ptr(mod_mask), n_rows, _ptr(mod_ids_dev), n_mods, _ptr(rows), _ptr(counts), _stream())
    return rows, counts


def gather_rows_bf16(src, rows, n, n_dev=None):
    """out[i] = src[rows[i]] for i < n (n_dev: device int32 [1] count <= n; rows [count, roundup(count, 128)) are zero-filled)."""
    _need_cuda(src, rows)
    D = src.shape[-1]
    out = torch.empty(n, D, device=src.device, dtype=torch.bfloat16)
    lib.call("b200fm_gather_rows_bf16_dyn", _ptr(src), _ptr(rows), _ptr(out), n, D, _ptr(n_dev), _stream())
    return out


def gather_i64(src, rows, n, n_dev=None):
    out = torch.empty(n, device=src.device, dtype=torch.int64) if n_dev is None else torch.zeros(n, device=src.device, dtype=torch.int64)
    lib.call("b200fm_gather_i64_dyn", _ptr(src), _ptr(rows), _ptr(out), n, _ptr(n_dev), _stream())
    return out


def scatter_add_rows(src_bf16, rows, dst, n):
    lib.call("b200fm_scatter_add_rows", _ptr(src_bf16), _ptr(rows), _ptr(dst), n, dst.shape[-1], _stream())


def scatter_rows_bf16(src, rows, dst, n, n_dev=None):
    lib.call("b200fm_scatter_rows_bf16_dyn", _ptr(src), _ptr(rows), _ptr(dst), n, dst.shape[-1], _ptr(n_dev), _stream())


def vq_ema_stats(flat, idx, K, cosine, stats=None):
    """Packed codebook statistics [K * (d + 1)] fp32 = [bins | embed_sum (K x d)] of one batch (quantize_lucid.py:409-419)."""
    _need_cuda(flat, idx)
    n, d = flat.shape
    assert flat.dtype == torch.float32 and flat.is_contiguous() and idx.dtype == torch.int64 and idx.numel() == n
    if stats is None:
        stats = torch.zeros(K * (d + 1), device=flat.device, dtype=torch.float32)
    lib.call("b200fm_vq_ema_stats", _ptr(flat), _ptr(idx), n, K, d, int(bool(cosine)), _ptr(stats), stats.data_ptr() + 4 * K, _stream())
    return stats


def vq_ema_update_cosine(embed, cluster_size, stats, decay):
    """In-place EMA of the cosine codebook buffers from the (all-reduced) packed statistics."""
    K, d = embed.shape
    assert embed.dtype == torch.float32 and embed.is_contiguous() and cluster_size.is_contiguous() and stats.numel() == K * (d + 1)
    lib.call("b200fm_vq_ema_update_cosine", _ptr(embed), _ptr(cluster_size), _ptr(stats), stats.data_ptr() + 4 * K, K, d, float(decay),
             _stream())
