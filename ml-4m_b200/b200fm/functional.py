"""Autograd functions over the C-ABI kernels: the differentiable building blocks of the 4M block stack.

Numerical contract (mirrors the reference under torch.autocast(bf16)): the residual stream is fp32;
LayerNorm statistics are fp32 and its output is consumed as bf16; every Linear / attention contraction takes bf16
operands with fp32 accumulation and produces bf16; weight gradients are produced in fp32 directly (the reference rounds
them to bf16 first).  Master weights stay fp32 `nn.Parameter`s; their bf16 shadows are cached per parameter version.
"""
import os
import weakref

import torch

from . import lib, ops

# ----------------------------------------------------------------------------------------------------------------------
# bf16 weight shadows
# ----------------------------------------------------------------------------------------------------------------------
_shadow = {}
_shadow_views = {}      # id(param) -> list of bf16 row-slices that mirror it (lets the fused optimizer refresh them in its own pass)


def _pad_rows(n, mult=8):
    return (n + mult - 1) // mult * mult


def _root(p):
    """The tensor whose identity / version counter stands for `p`: a reshaped view of a parameter (the 1x1 Conv2d weights used as
    Linear weights are `.reshape(out, in)` views created per call) is represented by its base, so repeated views share one cache
    entry instead of adding a new one on every call."""
    b = p._base
    return b if b is not None and b.numel() == p.numel() and p.is_contiguous() and b.is_contiguous() else p


def _evict(key_id):
    """weakref callback: the tensor that owned cache entries keyed on `key_id` died -> drop them (re-created models, EMA / eval
    copies and test suites would otherwise leak their bf16 mirrors, and a recycled id() could alias a dead entry)."""
    _shadow_views.pop(key_id, None)
    for k in [k for k in _shadow if (k[0] == key_id if not isinstance(k[0], tuple) else any(e[0] == key_id for e in k))]:
        _shadow.pop(k, None)


def _wref(t):
    kid = id(t)
    return weakref.ref(t, lambda _r, kid=kid: _evict(kid))


def weight_bf16(*params):
    """bf16 copy of one fp32 weight, or of several concatenated along dim 0 (e.g. [fc1; fc3]), refreshed when any
    source tensor's version counter changes (optimizer steps, load_state_dict).  Rows are zero-padded to a multiple of 8."""
    roots = [_root(p) for p in params]
    key = tuple((id(r), tuple(p.shape)) for r, p in zip(roots, params))
    ver = tuple(r._version for r in roots) + tuple(p.data_ptr() for p in params)
    hit = _shadow.get(key)
    if hit is not None and hit[0] == ver and all(w() is r for w, r in zip(hit[2], roots)):     # id() may be recycled: check identity
        return hit[1]
    with torch.no_grad():
        rows = [p.shape[0] for p in params]
        cols = params[0].shape[1]
        prow = [_pad_rows(r) for r in rows] if len(params) > 1 else rows
        buf = hit[1] if hit is not None and hit[1].shape == (sum(prow), cols) else None
        if buf is None:
            buf = torch.zeros(sum(prow), cols, device=params[0].device, dtype=torch.bfloat16)
        off = 0
        for p, root, r, pr in zip(params, roots, rows, prow):
            src = p.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            view = buf[off:off + r]
            ops.cast_bf16(src, view)
            ent = _shadow_views.get(id(root))
            if ent is None or ent[0]() is not root:
                ent = _shadow_views[id(root)] = (_wref(root), [])
            if hit is not None and buf is not hit[1]:
                # the operand buffer was re-allocated (shape change): forget the mirrors that lived in the old one, or the
                # fused optimizer would keep refreshing dead memory
                lo, hi = hit[1].data_ptr(), hit[1].data_ptr() + hit[1].numel() * 2
                ent[1][:] = [v for v in ent[1] if not lo <= v.data_ptr() < hi]
            if not any(v.data_ptr() == view.data_ptr() for v in ent[1]):
                ent[1].append(view)
            off += pr
    _shadow[key] = (ver, buf, tuple(_wref(r) for r in roots))
    return buf


def mark_updated(params):
    """`params` were rewritten IN PLACE through raw pointers (fused AdamW), together with every bf16 mirror registered for them
    in `_shadow_views`.  Bump their version counters so that any other cache keyed on the version (K-padded shadows, the conv
    weight re-layouts of the ViT tokenizers, user code) sees the change, and re-stamp the mirrors that were refreshed in the same
    pass so they are not cast again."""
    params = list(params)
    if not params:
        return
    torch._C._autograd._unsafe_set_version_counter(params, [p._version + 1 for p in params])
    ids = {id(p) for p in params}
    for key, hit in list(_shadow.items()):
        if len(key) and isinstance(key[0], tuple) and all(k[0] in ids for k in key):
            roots = [w() for w in hit[2]]
            if all(r is not None for r in roots):
                _shadow[key] = (tuple(r._version for r in roots) + hit[0][len(roots):], hit[1], hit[2])


def weight_bf16_padk(param, k_pad):
    """bf16 copy of a [N, K] weight with its K (column) dimension zero-padded to k_pad: operands whose K is not a multiple of 8
    (SwiGLU widths 2730 / 5461 of 4M-L / XL) need 16-byte rows for TMA.  Cached per parameter version."""
    key = (id(param), "padk", k_pad)
    ver = (param._version, param.data_ptr())
    hit = _shadow.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is param:
        return hit[1]
    with torch.no_grad():
        buf = hit[1] if hit is not None else torch.zeros(param.shape[0], k_pad, device=param.device, dtype=torch.bfloat16)
        buf[:, :param.shape[1]].copy_(param.detach())
    _shadow[key] = (ver, buf, _wref(param))
    return buf


# ----------------------------------------------------------------------------------------------------------------------
# precise mode: fp32-faithful inference for callers that run the reference WITHOUT bf16 autocast (VQ tokenization)
# ----------------------------------------------------------------------------------------------------------------------
_precise = {"on": False, "terms": 3}


class precise:
    """`with BF.precise():` -- inference-only (no autograd) code paths that support it (the ViT tokenizers) compute their contractions
    from bf16 limbs of the fp32 operands (csrc/precise.cu): terms=3 ~ 2^-16 relative per product, terms=6 fp32 class."""

    def __init__(self, terms=3):
        assert terms in (3, 6)
        self.terms = terms

    def __enter__(self):
        self.prev = dict(_precise)
        _precise.update(on=True, terms=self.terms)
        return self

    def __exit__(self, *exc):
        _precise.update(self.prev)


def is_precise():
    return _precise["on"] and not torch.is_grad_enabled()


def weight_limbs(param):
    """Limb layout (B operand) of an fp32 [N, K] weight for the current precise mode, cached per parameter version."""
    terms = _precise["terms"]
    root = _root(param)
    key = (id(root), "limbs", terms, tuple(param.shape))
    ver = (root._version, param.data_ptr())
    hit = _shadow.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is root:
        return hit[1]
    with torch.no_grad():
        w = param.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
        buf = ops.split_limbs(w, terms, 1)
    _shadow[key] = (ver, buf, _wref(root))
    return buf


def linear_f32(x, weight, bias=None, cache=True):
    """y (fp32) = x W^T + b with fp32-faithful arithmetic on the tcgen05 bf16 GEMM (limb products, fp32 accumulation).
    cache=False: `weight` is a derived tensor (not a parameter or a view of one): its limbs are rebuilt per call."""
    x2 = _as2d(x)
    if x2.dtype != torch.float32:
        x2 = x2.float()
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    a = ops.split_limbs(x2, _precise["terms"], 0)
    wl = weight_limbs(weight) if cache else ops.split_limbs(weight.detach().float().contiguous(), _precise["terms"], 1)
    y = ops.gemm(a, wl, epilogue=ops.EPI_F32, bias=bias, n_out=weight.shape[0])
    return y.view(*x.shape[:-1], weight.shape[0])


def shadow_views(param):
    """bf16 mirrors of `param` currently cached (the fused AdamW kernel writes the first one itself)."""
    ent = _shadow_views.get(id(param))
    return ent[1] if ent is not None and ent[0]() is param else []


def clear_weight_cache():
    _shadow.clear()
    _shadow_views.clear()


class _ZeroArena:
    """Hands out small zero-initialised fp32 vectors (LayerNorm dgamma / dbeta, embedding-sum gradients -- buffers the
    kernels ACCUMULATE into) as slices of one pre-zeroed chunk per device: one fill per ~1000 requests instead of one
    per request.  A slice is handed out once; the chunk lives as long as any slice of it does."""
    CHUNK = 1 << 20

    def __init__(self):
        self._chunks = {}

    def take(self, n, device):
        n_pad = (n + 3) & ~3                                  # keep every slice 16 B aligned
        if n_pad > self.CHUNK // 8:
            return torch.zeros(n, device=device, dtype=torch.float32)
        key = (device.type, device.index)
        ent = self._chunks.get(key)
        if ent is None or ent[1] + n_pad > self.CHUNK:
            ent = self._chunks[key] = [torch.zeros(self.CHUNK, device=device, dtype=torch.float32), 0]
        out = ent[0][ent[1]:ent[1] + n]
        ent[1] += n_pad
        return out


_zero_arena = _ZeroArena()


def zeros_f32(n, device):
    return _zero_arena.take(n, device)


def reset_zero_arena():
    """Forget the current pre-zeroed chunks.  Called around CUDA-graph capture: the first request INSIDE the capture then allocates
    (and memsets) a chunk as part of the graph, so every replay starts from zeros; after the capture eager code must not hand out
    slices of the graph's chunk."""
    _zero_arena._chunks.clear()


# ----------------------------------------------------------------------------------------------------------------------
# gradient destinations: with data-parallel training (b200fm.parallel.GradSync) every parameter has a slot in the flat gradient
# arena that is all-reduced over NVLink; the backward kernels write there directly instead of into a fresh tensor.
# ----------------------------------------------------------------------------------------------------------------------
def _claim(param, zeroed=False):
    """Arena view to write `param`'s gradient into (first producer of the step only), or None."""
    s = getattr(param, "_b200fm_slot", None)
    return s.sync.claim(param, zeroed) if s is not None else None


def _acc_f32(param, n, device):
    """Zero-initialised fp32 [n] buffer a kernel ACCUMULATES a small parameter's gradient into (dgamma, dbeta, mod_emb, ...)."""
    v = _claim(param, zeroed=True)
    return v.view(-1) if v is not None else zeros_f32(n, device)


# Weight-gradient GEMMs are LEAVES of the backward graph: nothing in backward consumes them.  Under CUDA-graph capture they can run
# on a second captured stream, concurrently with the dgrad / attention / LayerNorm chain: the persistent GEMM grids are statically
# scheduled, so the last partial wave of a kernel leaves SMs idle (2.6 waves of CTA-pair tiles for the K = N = 768 shapes) -- CTAs of
# a wgrad kernel from the other stream fill them.  Enabled by GraphedTrainStep (B200FM_SIDE_WGRAD=1); operands are kept alive until
# the join so the graph's allocator cannot reuse their blocks while the side stream still reads them.
_side = {"stream": None, "keep": [], "used": False}


def enable_side_wgrad(stream):
    _side.update(stream=stream, keep=[], used=False)


def side_stream():
    return _side["stream"] if _side["used"] else None


def join_side_wgrad(disable=False):
    """Make the current stream wait for all side-stream weight gradients issued so far (before the optimizer reads them)."""
    st = _side["stream"]
    if st is not None and _side["used"]:
        torch.cuda.current_stream().wait_stream(st)
        _side["keep"].clear()
        _side["used"] = False
    if disable:
        _side["stream"] = None


def _tn_gemm(a, b, out, alpha_dev=None, dyn=None):
    st = _side["stream"]
    if st is None:
        return ops.gemm(a, b, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32, out=out, alpha_dev=alpha_dev, dyn=dyn)
    st.wait_stream(torch.cuda.current_stream())               # the producers of a / b (and of the output's previous contents)
    with torch.cuda.stream(st):
        res = ops.gemm(a, b, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32, out=out, alpha_dev=alpha_dev, dyn=dyn)
    _side["keep"].append((a, b, alpha_dev, dyn))
    _side["used"] = True
    return res


def _wgrad(dy2, x2, weight, alpha_dev=None, dyn=None):
    """dW (fp32) = dy^T x, written into the parameter's arena slot when there is one."""
    return _tn_gemm(dy2, x2, _claim(weight), alpha_dev, dyn)


def _wgrad13(dab, h, w1, w3, H, Hp):
    """[dW1; dW3] of a SwiGLU block: ONE [2Hp, D] GEMM output, placed over the adjacent fc1 / fc3 arena slots when possible."""
    s = getattr(w1, "_b200fm_slot", None)
    out = s.sync.claim_pair(w1, w3, Hp) if s is not None else None
    dw13 = _tn_gemm(dab, h, out)
    return dw13[:H], dw13[Hp:Hp + H]


def _as2d(x):
    return x.reshape(-1, x.shape[-1])


# ----------------------------------------------------------------------------------------------------------------------
# LayerNorm
# ----------------------------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x) with fp32 statistics; y is bf16 (feeds a GEMM) or fp32.  (fm_utils.py:93-108)"""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_bf16):
        x2 = _as2d(x).contiguous()
        if x2.dtype != torch.float32:
            x2 = x2.float()
        y, mean, rstd = ops.layernorm_fwd(x2, weight, bias, eps, out_bf16=out_bf16)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.has_bias = bias is not None and bias.requires_grad
        ctx.bias_param = bias if ctx.has_bias else None
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        dy2 = _as2d(dy).contiguous()
        D = x2.shape[1]
        dgamma = _acc_f32(weight, D, x2.device) if weight.requires_grad else None
        dbeta = _acc_f32(ctx.bias_param, D, x2.device) if ctx.has_bias else None
        dx, _ = ops.layernorm_bwd(dy2, x2, weight, mean, rstd, dgamma=dgamma, dbeta=dbeta)
        return dx.view(ctx.shape), dgamma, dbeta, None, None


def layer_norm(x, weight, bias, eps=1e-6, out_bf16=True):
    return LayerNormFn.apply(x, weight, bias, eps, out_bf16)


# ----------------------------------------------------------------------------------------------------------------------
# Linear family
# ----------------------------------------------------------------------------------------------------------------------
def _to_bf16_2d(x):
    x2 = _as2d(x)
    if x2.dtype == torch.float32:
        return ops.cast_bf16(x2.contiguous())
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    return x2


def _linear_bwd(ctx, dy2, x2, weight, want_dx=True):
    """dx = dy W (NN), dW = dy^T x (TN, fp32), db = colsum(dy)."""
    wb = weight_bf16(weight)[:weight.shape[0]]
    dx = ops.gemm(dy2, wb, layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16) if want_dx else None
    dw = _wgrad(dy2, x2, weight) if weight.requires_grad else None
    return dx, dw


class LinearFn(torch.autograd.Function):
    """y(bf16) = x W^T (+ b).  nn.Linear under autocast (fm_utils.py:155-157 etc.)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = _to_bf16_2d(x)
        y = ops.gemm(x2, weight_bf16(weight), epilogue=ops.EPI_BF16, bias=bias, n_out=weight.shape[0])
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = _to_bf16_2d(dy)
        dx, dw = _linear_bwd(ctx, dy2, x2, weight, ctx.needs_input_grad[0])
        db = ops.colsum_bf16(dy2) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return (dx.view(ctx.in_shape) if dx is not None else None), dw, db


class PatchEmbedFn(torch.autograd.Function):
    """tokens(fp32) = pos + bf16(patches W_lin^T + b): the k = s = P Conv2d patch projection of the ViT tokenizers
    (vq/models/vit_models.py:482-489) as patchify + one GEMM with the positional rows in the epilogue.  conv_weight is the
    module's [O, C, P, P] parameter (state_dict contract); patches are laid out '(ph pw c)', so W_lin = W.permute(0, 2, 3, 1).
    Backward: dW (TN GEMM on the saved bf16 patches, mapped back to the conv layout), db, and the positional rows' gradient
    (= dout: a learnable / un-frozen pos_emb reaches it through the caller's reshape); images get no gradient."""

    @staticmethod
    def forward(ctx, img, conv_weight, conv_bias, pos_rows, w_lin_bf16, patch):
        patches = ops.patchify(img.float().contiguous(), patch)                       # [B*N, P*P*C] bf16
        bias = conv_bias.detach().float() if conv_bias is not None else None
        out = ops.gemm(patches, w_lin_bf16, epilogue=ops.EPI_RESID, bias=bias, resid=pos_rows)
        ctx.save_for_backward(patches)
        ctx.wshape, ctx.has_bias, ctx.pshape = conv_weight.shape, conv_bias is not None, pos_rows.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        (patches,) = ctx.saved_tensors
        d2 = _to_bf16_2d(dout)
        O, C, P, _ = ctx.wshape
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = ops.gemm(d2, patches, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32).view(O, P, P, C).permute(0, 3, 1, 2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum_bf16(d2)
        dpos = dout.reshape(ctx.pshape).float() if ctx.needs_input_grad[3] else None
        return None, dw, db, dpos, None, None


class LinearF32Fn(torch.autograd.Function):
    """y(fp32) = x W^T : logits (decoder_embeddings.py:141-152) when the caller wants them materialised."""

    @staticmethod
    def forward(ctx, x, weight):
        x2 = _to_bf16_2d(x)
        y = ops.gemm(x2, weight_bf16(weight), epilogue=ops.EPI_F32, n_out=weight.shape[0])
        ctx.save_for_backward(x2, weight)
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = _to_bf16_2d(dy)
        dx, dw = _linear_bwd(ctx, dy2, x2, weight, ctx.needs_input_grad[0])
        return (dx.view(ctx.in_shape) if dx is not None else None), dw


class LinearResidualFn(torch.autograd.Function):
    """out(fp32) = resid(fp32) + bf16(x W^T + b):   x = x + proj(...) / x = x + fc2(...) (fm_utils.py:332-334)."""

    @staticmethod
    def forward(ctx, x, weight, bias, resid):
        x2 = _to_bf16_2d(x)
        r2 = _as2d(resid).contiguous()
        out = ops.gemm(x2, weight_bf16(weight), epilogue=ops.EPI_RESID, bias=bias, resid=r2, n_out=weight.shape[0])
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.in_shape = x.shape
        return out.view(resid.shape)

    @staticmethod
    def backward(ctx, dout):
        x2, weight = ctx.saved_tensors
        d2 = _as2d(dout).contiguous()
        dy2 = ops.cast_bf16(d2) if d2.dtype == torch.float32 else d2
        dx, dw = _linear_bwd(ctx, dy2, x2, weight, ctx.needs_input_grad[0])
        db = ops.colsum_bf16(dy2) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return (dx.view(ctx.in_shape) if dx is not None else None), dw, db, dout


class SwiGLUFn(torch.autograd.Function):
    """g(bf16) = silu(x W1^T) * (x W3^T) with [W1; W3] streamed as one B operand (fm_utils.py:129-144)."""

    @staticmethod
    def forward(ctx, x, w1, w3, b1, b3):
        x2 = _to_bf16_2d(x)
        H = w1.shape[0]
        Hp = _pad_rows(H)
        w13 = weight_bf16(w1, w3)                                   # [2*Hp, D], rows zero-padded to a multiple of 8
        bias = None
        if b1 is not None:
            bias = torch.zeros(2 * Hp, device=x2.device, dtype=torch.float32)
            bias[:H] = b1
            bias[Hp:Hp + H] = b3
        ab, g = ops.gemm(x2, w13, epilogue=ops.EPI_SWIGLU, bias=bias)
        ctx.save_for_backward(x2, w1, w3, ab)
        ctx.H, ctx.Hp, ctx.has_bias, ctx.in_shape = H, Hp, b1 is not None, x.shape
        return g[:, :H].view(*x.shape[:-1], H) if Hp != H else g.view(*x.shape[:-1], H)

    @staticmethod
    def backward(ctx, dg):
        x2, w1, w3, ab = ctx.saved_tensors
        H, Hp = ctx.H, ctx.Hp
        dg2 = _as2d(dg)
        if Hp != H:
            pad = torch.zeros(dg2.shape[0], Hp, device=dg2.device, dtype=torch.bfloat16)
            pad[:, :H] = dg2
            dg2 = pad
        elif dg2.stride(-1) != 1 or dg2.dtype != torch.bfloat16:
            dg2 = _to_bf16_2d(dg2)
        dab = ops.swiglu_bwd(ab, dg2)                               # [R, 2*Hp] = [da | db]
        w13 = weight_bf16(w1, w3)
        dx = ops.gemm(dab, w13, layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16) if ctx.needs_input_grad[0] else None
        dw1 = dw3 = db1 = db3 = None
        if w1.requires_grad or w3.requires_grad:
            dw1, dw3 = _wgrad13(dab, x2, w1, w3, H, Hp)
        if ctx.has_bias:
            dbias = ops.colsum_bf16(dab)
            db1, db3 = dbias[:H], dbias[Hp:Hp + H]
        return (dx.view(ctx.in_shape) if dx is not None else None), dw1, dw3, db1, db3


class MlpActFn(torch.autograd.Function):
    """act(bf16) = GELU|tanh(x W^T + b) with the pre-activation saved by the same GEMM epilogue (fm_utils.py:111-126)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act_name="gelu"):
        x2 = _to_bf16_2d(x)
        epi = ops.EPI_GELU if act_name == "gelu" else ops.EPI_TANH
        pre, act = ops.gemm(x2, weight_bf16(weight), epilogue=epi, bias=bias, n_out=weight.shape[0])
        ctx.save_for_backward(x2, weight, pre)
        ctx.has_bias, ctx.in_shape, ctx.act_name = bias is not None, x.shape, act_name
        return act.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dact):
        x2, weight, pre = ctx.saved_tensors
        dpre = ops.act_bwd(pre, _to_bf16_2d(dact).contiguous(), ctx.act_name)
        dx, dw = _linear_bwd(ctx, dpre, x2, weight, ctx.needs_input_grad[0])
        db = ops.colsum_bf16(dpre) if ctx.has_bias else None
        return (dx.view(ctx.in_shape) if dx is not None else None), dw, db, None


# ----------------------------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------------------------
ATTN_BWD_MAX_TOKENS = 256      # b200fm_attention_bwd: query / key tiles of one (batch, head) item live in TMEM for the whole item


def _check_trainable_length(ctx, Nq, Nk):
    """The forward accepts longer sequences (attention_fwd_long, generation contexts) than the backward kernel.  A forward that records
    gradients over such a sequence is legal as long as backward() is never called (generation code that forgets torch.no_grad()), so this
    only WARNS -- once -- that a later backward() will be refused (b200fm_attention_bwd reports the supported range)."""
    global _warned_long
    if not _warned_long and any(ctx.needs_input_grad) and max(Nq, Nk) > ATTN_BWD_MAX_TOKENS:
        _warned_long = True
        import warnings
        warnings.warn(f"attention over {Nq} queries x {Nk} keys is recorded for autograd, but the B200 backward kernel supports at most "
                      f"{ATTN_BWD_MAX_TOKENS} tokens per side (encoder tokens + register tokens, decoder tokens): backward() through it "
                      f"will raise.  Use torch.no_grad() for inference, or lower num_encoder_tokens / num_decoder_tokens for training.",
                      RuntimeWarning, stacklevel=3)


_warned_long = False


class AttentionFn(torch.autograd.Function):
    """softmax(q k^T * scale, masked) v per head on 2-D row views (fm_utils.py:160-180 / 197-219).
    q [B*Nq, H*64], k / v [B*Nk, H*64] bf16 (column slices of packed qkv / kv buffers are fine)."""

    @staticmethod
    def forward(ctx, q, k, v, mask, B, H, Nq, Nk, scale):
        _check_trainable_length(ctx, Nq, Nk)
        out, stats = ops.attention_fwd(q, k, v, B, H, Nq, Nk, mask, scale)
        ctx.save_for_backward(q, k, v, out, stats, mask)
        ctx.dims = (B, H, Nq, Nk, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, stats, mask = ctx.saved_tensors
        B, H, Nq, Nk, scale = ctx.dims
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        D = H * 64
        # q/k/v that are slices of one packed buffer get their gradients written into one packed buffer too, so the
        # following dgrad / wgrad GEMMs read a single [rows, 3D] (or [rows, 2D]) operand.
        packed_qkv = (q.data_ptr() + 2 * D == k.data_ptr() and k.data_ptr() + 2 * D == v.data_ptr() and q.stride(0) == 3 * D)
        packed_kv = (not packed_qkv) and (k.data_ptr() + 2 * D == v.data_ptr() and k.stride(0) == 2 * D)
        if packed_qkv:
            buf = torch.empty(q.shape[0], 3 * D, device=q.device, dtype=torch.bfloat16)
            dq, dk, dv = buf[:, :D], buf[:, D:2 * D], buf[:, 2 * D:]
        elif packed_kv:
            dq = torch.empty(q.shape[0], D, device=q.device, dtype=torch.bfloat16)
            buf = torch.empty(k.shape[0], 2 * D, device=q.device, dtype=torch.bfloat16)
            dk, dv = buf[:, :D], buf[:, D:]
        else:
            dq = dk = dv = None
        dq, dk, dv = ops.attention_bwd(q, k, v, out, dout, stats, B, H, Nq, Nk, mask, scale, dq, dk, dv)
        return dq, dk, dv, None, None, None, None, None, None


class HeadNormFn(torch.autograd.Function):
    """LayerNorm over head_dim on every (row, head) of q or k (fm_utils.py:244-245, 290-291; qk_norm presets)."""

    @staticmethod
    def forward(ctx, x, H, weight, bias, eps):
        y, stats = ops.headnorm_fwd(x, H, weight, bias, eps)
        ctx.save_for_backward(x, stats, weight)
        ctx.H = H
        ctx.bias_grad = bias is not None and bias.requires_grad
        ctx.bias_param = bias if ctx.bias_grad else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, weight = ctx.saved_tensors
        if dy.dtype != torch.bfloat16 or dy.stride(1) != 1:
            dy = dy.to(torch.bfloat16).contiguous()
        dgamma = _acc_f32(weight, 64, x.device) if weight.requires_grad else None
        dbeta = _acc_f32(ctx.bias_param, 64, x.device) if ctx.bias_grad else None
        dx = ops.headnorm_bwd(dy, x, weight, stats, ctx.H, dgamma, dbeta)
        return dx, None, dgamma, dbeta, None


def head_norm(x, H, norm):
    """x bf16 [R, H*64] (may be a column slice of a packed qkv) through `norm` = LayerNorm(head_dim)."""
    if tuple(norm.weight.shape) != (64,):
        raise NotImplementedError("q/k LayerNorm is implemented for head_dim 64 (every reference preset)")
    return HeadNormFn.apply(x, H, norm.weight, getattr(norm, "bias", None), float(norm.eps))


def attention(q, k, v, mask, B, H, Nq, Nk, scale):
    return AttentionFn.apply(q, k, v, mask, B, H, Nq, Nk, scale)


# ----------------------------------------------------------------------------------------------------------------------
# masked-token head: logits GEMM + cross-entropy, gradient kept as bf16 (softmax - onehot)
# ----------------------------------------------------------------------------------------------------------------------
# The masked-token head keeps the logits inside the GEMM (b200fm_head_ce: statistics epilogue, row reduction, gradient epilogue) instead
# of writing fp32 [rows, V] logits and re-reading them in a cross-entropy kernel.  B200FM_FUSED_HEAD=0 selects the two-kernel path.
FUSED_HEAD = os.environ.get("B200FM_FUSED_HEAD", "1") != "0"


class LinearCrossEntropyFn(torch.autograd.Function):
    """mean_i CE(h_i W^T, t_i)  (fm.py:592-598).  h bf16 [n, D], W fp32 master [V, D], targets int64 [n]."""

    @staticmethod
    def forward(ctx, h, weight, targets):
        wb = weight_bf16(weight)
        if FUSED_HEAD:
            loss_rows, dlogits = ops.head_ce(h, wb, weight.shape[0], targets, want_grad=True)
        else:
            logits = ops.gemm(h, wb, epilogue=ops.EPI_F32, n_out=weight.shape[0])
            loss_rows, dlogits = ops.cross_entropy(logits, targets, want_grad=True)
            del logits
        ctx.save_for_backward(h, weight, dlogits)
        return loss_rows.mean()

    @staticmethod
    def backward(ctx, dloss):
        h, weight, dlogits = ctx.saved_tensors
        n = h.shape[0]
        coef = (dloss.float() / n).reshape(1).contiguous()             # stays on the device
        wb = weight_bf16(weight)[:weight.shape[0]]
        dh = ops.gemm(dlogits, wb, layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16, alpha_dev=coef) if ctx.needs_input_grad[0] else None
        dw = _wgrad(dlogits, h, weight, alpha_dev=coef) if weight.requires_grad else None
        return dh, dw, None


class LinearCrossEntropyStaticFn(torch.autograd.Function):
    """As LinearCrossEntropyFn with a DEVICE-side row count: h bf16 [R, D] holds the modality's rows first (zero rows behind them),
    targets int64 [R], n_dev int32 [1].  Every launch has a host-independent shape, so the head can be captured in a CUDA graph and
    costs no host synchronisation; tiles / rows beyond the count are skipped on the device.  Empty modality -> loss 0 (fm.py:593-595)."""

    @staticmethod
    def forward(ctx, h, weight, targets, n_dev):
        wb = weight_bf16(weight)
        if FUSED_HEAD:
            loss_rows, dlogits = ops.head_ce(h, wb, weight.shape[0], targets, n_dev, want_grad=True)
        else:
            logits = ops.gemm(h, wb, epilogue=ops.EPI_F32, n_out=weight.shape[0], dyn=n_dev)
            loss_rows, dlogits = ops.cross_entropy_dyn(logits, targets, n_dev, want_grad=True)
            del logits
        loss, inv_n = ops.masked_mean(loss_rows, n_dev)
        ctx.save_for_backward(h, weight, dlogits, n_dev, inv_n)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        h, weight, dlogits, n_dev, inv_n = ctx.saved_tensors
        coef = (dloss.float() * inv_n).reshape(1).contiguous()
        wb = weight_bf16(weight)[:weight.shape[0]]
        dh = ops.gemm(dlogits, wb, layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16, alpha_dev=coef, dyn=n_dev) if ctx.needs_input_grad[0] else None
        dw = None
        if weight.requires_grad:
            dw = _wgrad(dlogits, h, weight, alpha_dev=coef, dyn=n_dev)
        return dh, dw, None, None


class HeadGatherStaticFn(torch.autograd.Function):
    """HeadGatherFn with device-side counts: every part is a full-height [R, D] buffer whose first counts[i] rows are the
    modality's rows (rows up to the next multiple of 128 are zero)."""

    @staticmethod
    def forward(ctx, src, rows, counts_dev):
        ctx.save_for_backward(rows, counts_dev)
        ctx.shape = src.shape
        R = src.shape[0]
        return tuple(ops.gather_rows_bf16(src, rows[i], R, counts_dev[i:i + 1]) for i in range(rows.shape[0]))

    @staticmethod
    def backward(ctx, *douts):
        rows, counts_dev = ctx.saved_tensors
        d = torch.zeros(ctx.shape, device=rows.device, dtype=torch.bfloat16)
        for i, g in enumerate(douts):
            if g is not None:
                ops.scatter_rows_bf16(g.contiguous(), rows[i], d, ctx.shape[0], counts_dev[i:i + 1])
        return d, None, None


class HeadGatherFn(torch.autograd.Function):
    """Per-modality row sets of the decoder output (y[decoder_mod_mask == idx], fm.py:591) in one autograd node:
    src bf16 [R, D]; rows int32 [n_mods, R] device index lists; counts: python ints -> tuple of bf16 [n_m, D]."""

    @staticmethod
    def forward(ctx, src, rows, counts):
        ctx.save_for_backward(rows)
        ctx.counts, ctx.shape = counts, src.shape
        return tuple(ops.gather_rows_bf16(src, rows[i], n) for i, n in enumerate(counts))

    @staticmethod
    def backward(ctx, *douts):
        (rows,) = ctx.saved_tensors
        d = torch.zeros(ctx.shape, device=rows.device, dtype=torch.bfloat16)
        for i, (n, g) in enumerate(zip(ctx.counts, douts)):
            if n > 0 and g is not None:
                ops.scatter_rows_bf16(g.contiguous(), rows[i], d, n)
        return d, None, None


# ----------------------------------------------------------------------------------------------------------------------
# embedding gather (forward) / scatter (backward) over a selection plan
# ----------------------------------------------------------------------------------------------------------------------
class EmbedRowsFn(torch.autograd.Function):
    """x0 = x + emb and emb for the kept rows of one side.  `seg_static` holds the non-differentiable part of every
    segment; the differentiable tensors arrive flat: mask_token, then per segment (token_emb | x_rows, mod_emb, pos_emb).
    pos_emb gets a gradient only when it is a learnable table (sincos_pos_emb=False: tok_dinov2_global / tok_imagebind_global of
    the reference's MODALITY_INFO, modality_info.py:280-297); the sincos buffers never require grad."""

    @staticmethod
    def forward(ctx, plan, seg_static, D, want_emb, mask_token, *tensors):
        segs = []
        for i, st in enumerate(seg_static):
            d = dict(st)
            main, mod = tensors[3 * i], tensors[3 * i + 1]
            if st["kind"] in (lib.KIND_IMG, lib.KIND_SEQ_EMB):
                d["x_rows"] = main
            else:
                d["token_emb"] = main
            d["mod_emb"] = mod.reshape(-1)
            segs.append(d)
        x0, emb = ops.embed_rows(plan, segs, None if mask_token is None else mask_token.reshape(-1), D, want_emb)
        ctx.plan, ctx.seg_static, ctx.D = plan, seg_static, D
        ctx.shapes = [(t.shape, t.dtype) for t in tensors]
        ctx.mask_token_shape = None if mask_token is None else mask_token.shape
        ctx.params, ctx.mask_token_param = tensors, mask_token       # leaf parameters / activations: destinations of their own gradients
        return (x0, emb) if want_emb else (x0, None)

    @staticmethod
    def backward(ctx, dx0, demb):
        plan, D = ctx.plan, ctx.D
        dev = dx0.device
        grads, segs = [], []
        for i, st in enumerate(ctx.seg_static):
            d = dict(st)
            (mshape, _), (eshape, _), (pshape, _) = ctx.shapes[3 * i], ctx.shapes[3 * i + 1], ctx.shapes[3 * i + 2]
            need_main, need_mod, need_pos = ctx.needs_input_grad[5 + 3 * i: 8 + 3 * i]
            gm = None
            decoder_side = bool(plan.decoder & ops.MODE_DECODER)
            if need_main and decoder_side and st["kind"] == lib.KIND_TOK_IMG:
                need_main = False          # decoder image tokens enter as the mask token (fm.py:322): their table gets no gradient here
            if need_main:
                if st["kind"] in (lib.KIND_IMG, lib.KIND_SEQ_EMB):
                    gm = torch.zeros(mshape, device=dev, dtype=torch.bfloat16)
                    d["dx_rows"] = gm
                else:
                    gm = _claim(ctx.params[3 * i])
                    gm = gm.zero_() if gm is not None else torch.zeros(mshape, device=dev, dtype=torch.float32)
                    d["d_token_emb"] = gm
            ge = _acc_f32(ctx.params[3 * i + 1], D, dev) if need_mod else None
            d["d_mod_emb"] = ge
            gp = torch.zeros(pshape, device=dev, dtype=torch.float32) if need_pos else None
            d["d_pos_emb"] = gp
            grads += [gm, None if ge is None else ge.view(eshape), gp]
            segs.append(d)
        dmt = _acc_f32(ctx.mask_token_param, D, dev) if (ctx.mask_token_shape is not None and ctx.needs_input_grad[4]) else None
        ops.embed_rows_bwd(plan, segs, dx0.contiguous(), None if demb is None else demb.contiguous(), dmt, D)
        return (None, None, None, None, None if dmt is None else dmt.view(ctx.mask_token_shape), *grads)


# ----------------------------------------------------------------------------------------------------------------------
# fused sub-layers: one autograd node per residual branch.  The backward chains the kernels by hand so that
#   * the residual-stream gradient add is done inside the LayerNorm-backward kernel (dres),
#   * that kernel also emits the bf16 copy of the new residual gradient that the next (earlier) sub-layer's dgrad / wgrad
#     GEMMs consume (handed over as an attribute of the fp32 gradient tensor; absent -> one cast kernel),
# which removes every framework element-wise kernel from the block stack's backward.
# ----------------------------------------------------------------------------------------------------------------------
def _grad_bf16(dout2, dout):
    b = getattr(dout, "_b200fm_bf16", None)
    if b is not None and b.shape == dout2.shape:
        return b
    return ops.cast_bf16(dout2) if dout2.dtype == torch.float32 else dout2


def _hand_over(dx, dxb, shape):
    out = dx.view(shape)
    out._b200fm_bf16 = dxb
    return out


def _prep_stream(x, ypend, D):
    x2 = x.reshape(-1, D)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    y2 = None
    if ypend is not None:
        y2 = ypend.reshape(-1, D)
        if not y2.is_contiguous():
            y2 = y2.contiguous()
    return x2, y2


def _stream_grads(g_s, g_y, rows, D):
    """(fp32 residual-stream gradient | None, bf16 branch-output gradient) as contiguous 2-D tensors."""
    d2 = None
    if g_s is not None:
        d2 = g_s.reshape(rows, D)
        if not d2.is_contiguous():
            d2 = d2.contiguous()
    db = g_y.reshape(rows, D)
    if db.dtype != torch.bfloat16:
        db = ops.cast_bf16(db.float().contiguous())
    elif not db.is_contiguous():
        db = db.contiguous()
    return d2, db


# Sub-layer convention ("pending add"): every node takes the fp32 stream x and the previous sub-layer's bf16 branch output
# `ypend` (not yet added), does  s = x + ypend  inside its LayerNorm kernel, and returns (s, y) with y its own bf16 branch
# output.  The GEMMs therefore always use the plain bf16 epilogue; the fp32 stream is touched only by the norm kernels.
# Backward receives (grad wrt s: fp32, grad wrt y: bf16) and returns (grad wrt x: fp32, grad wrt ypend: its bf16 copy).
class SelfAttnSubLayerFn(torch.autograd.Function):
    """(x, ypend) -> (s = x + ypend, y = proj(attention(qkv(LN(s)))))   fm_utils.py:332 / 363 with Attention.forward (:160-180)."""

    @staticmethod
    def forward(ctx, x, ypend, mask, nw, nb, qkv_w, qkv_b, proj_w, proj_b, eps, heads, scale):
        B, N, D = x.shape
        _check_trainable_length(ctx, N, N)
        x2, y2 = _prep_stream(x, ypend, D)
        s2, h, mean, rstd = ops.add_layernorm_fwd(x2, y2, nw, nb, eps)
        qkv = ops.gemm(h, weight_bf16(qkv_w), epilogue=ops.EPI_BF16, bias=qkv_b, n_out=3 * D)
        o, stats = ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, heads, N, N, mask, scale)
        y = ops.gemm(o, weight_bf16(proj_w), epilogue=ops.EPI_BF16, bias=proj_b, n_out=D)
        ctx.save_for_backward(s2, mean, rstd, h, qkv, o, stats, mask, nw, qkv_w, proj_w)
        ctx.cfg = (B, N, D, heads, scale, qkv_b is not None, proj_b is not None, nb is not None and nb.requires_grad, ypend is not None)
        ctx.nb = nb
        return s2.view(B, N, D), y.view(B, N, D)

    @staticmethod
    def backward(ctx, g_s, g_y):
        s2, mean, rstd, h, qkv, o, stats, mask, nw, qkv_w, proj_w = ctx.saved_tensors
        B, N, D, heads, scale, has_qb, has_pb, nb_grad, has_pend = ctx.cfg
        d2, db = _stream_grads(g_s, g_y, B * N, D)
        do = ops.gemm(db, weight_bf16(proj_w)[:D], layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)
        dproj_w = _wgrad(db, o, proj_w) if proj_w.requires_grad else None
        dproj_b = ops.colsum_bf16(db) if has_pb else None
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, stats, B, heads, N, N, mask, scale,
                          dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:])
        dh = ops.gemm(dqkv, weight_bf16(qkv_w)[:3 * D], layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)
        dqkv_w = _wgrad(dqkv, h, qkv_w) if qkv_w.requires_grad else None
        dqkv_b = ops.colsum_bf16(dqkv) if has_qb else None
        dgamma = _acc_f32(nw, D, s2.device) if nw.requires_grad else None
        dbeta = _acc_f32(ctx.nb, D, s2.device) if nb_grad else None
        dx, dxb = ops.layernorm_bwd(dh, s2, nw, mean, rstd, dres=d2, want_bf16=has_pend, dgamma=dgamma, dbeta=dbeta)
        return (dx.view(B, N, D), dxb.view(B, N, D) if has_pend else None, None, dgamma, dbeta, dqkv_w, dqkv_b, dproj_w, dproj_b,
                None, None, None)


class CrossAttnSubLayerFn(torch.autograd.Function):
    """(x, ypend, context) -> (s, y = proj(attention(q(LNq(s)), kv(LNc(context)))))   fm_utils.py:364, CrossAttention.forward (:197-219)."""

    @staticmethod
    def forward(ctx, x, ypend, context, mask, qnw, qnb, cnw, cnb, q_w, q_b, kv_w, kv_b, proj_w, proj_b, eps_q, eps_c, heads, scale):
        B, N, D = x.shape
        M = context.shape[1]
        _check_trainable_length(ctx, N, M)
        x2, y2 = _prep_stream(x, ypend, D)
        c2 = context.reshape(B * M, D)
        if not c2.is_contiguous():
            c2 = c2.contiguous()
        s2, hq, qmean, qrstd = ops.add_layernorm_fwd(x2, y2, qnw, qnb, eps_q)
        hc, cmean, crstd = ops.layernorm_fwd(c2, cnw, cnb, eps_c, out_bf16=True)
        q = ops.gemm(hq, weight_bf16(q_w), epilogue=ops.EPI_BF16, bias=q_b, n_out=D)
        kv = ops.gemm(hc, weight_bf16(kv_w), epilogue=ops.EPI_BF16, bias=kv_b, n_out=2 * D)
        o, stats = ops.attention_fwd(q, kv[:, :D], kv[:, D:], B, heads, N, M, mask, scale)
        y = ops.gemm(o, weight_bf16(proj_w), epilogue=ops.EPI_BF16, bias=proj_b, n_out=D)
        ctx.save_for_backward(s2, c2, qmean, qrstd, cmean, crstd, hq, hc, q, kv, o, stats, mask, qnw, cnw, q_w, kv_w, proj_w)
        ctx.cfg = (B, N, M, D, heads, scale, q_b is not None, kv_b is not None, proj_b is not None,
                   qnb is not None and qnb.requires_grad, cnb is not None and cnb.requires_grad, ypend is not None)
        ctx.qnb, ctx.cnb = qnb, cnb
        return s2.view(B, N, D), y.view(B, N, D)

    @staticmethod
    def backward(ctx, g_s, g_y):
        (s2, c2, qmean, qrstd, cmean, crstd, hq, hc, q, kv, o, stats, mask, qnw, cnw, q_w, kv_w, proj_w) = ctx.saved_tensors
        B, N, M, D, heads, scale, has_qb, has_kvb, has_pb, qnb_grad, cnb_grad, has_pend = ctx.cfg
        dev = s2.device
        d2, db = _stream_grads(g_s, g_y, B * N, D)
        do = ops.gemm(db, weight_bf16(proj_w)[:D], layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)
        dproj_w = _wgrad(db, o, proj_w) if proj_w.requires_grad else None
        dproj_b = ops.colsum_bf16(db) if has_pb else None
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        ops.attention_bwd(q, kv[:, :D], kv[:, D:], o, do, stats, B, heads, N, M, mask, scale, dq, dkv[:, :D], dkv[:, D:])
        dhq = ops.gemm(dq, weight_bf16(q_w)[:D], layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)
        dq_w = _wgrad(dq, hq, q_w) if q_w.requires_grad else None
        dq_b = ops.colsum_bf16(dq) if has_qb else None
        dkv_w = _wgrad(dkv, hc, kv_w) if kv_w.requires_grad else None
        dkv_b = ops.colsum_bf16(dkv) if has_kvb else None
        dqg = _acc_f32(qnw, D, dev) if qnw.requires_grad else None
        dqb = _acc_f32(ctx.qnb, D, dev) if qnb_grad else None
        dx, dxb = ops.layernorm_bwd(dhq, s2, qnw, qmean, qrstd, dres=d2, want_bf16=has_pend, dgamma=dqg, dbeta=dqb)
        dctx = dcg = dcb = None
        # context_norm's own weight / bias train even when the context itself is detached (frozen-encoder fine-tuning,
        # forward_decoder on a no_grad encoder output): run the dgrad + LayerNorm backward whenever any of the three is wanted
        if ctx.needs_input_grad[2] or cnw.requires_grad or cnb_grad:
            dhc = ops.gemm(dkv, weight_bf16(kv_w)[:2 * D], layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)
            dcg = _acc_f32(cnw, D, dev) if cnw.requires_grad else None
            dcb = _acc_f32(ctx.cnb, D, dev) if cnb_grad else None
            dctx, _ = ops.layernorm_bwd(dhc, c2, cnw, cmean, crstd, dgamma=dcg, dbeta=dcb)
            dctx = dctx.view(B, M, D) if ctx.needs_input_grad[2] else None
        return (dx.view(B, N, D), dxb.view(B, N, D) if has_pend else None, dctx, None, dqg, dqb, dcg, dcb, dq_w, dq_b, dkv_w, dkv_b,
                dproj_w, dproj_b, None, None, None, None)


class GatedMlpSubLayerFn(torch.autograd.Function):
    """(x, ypend) -> (s, y = fc2(silu(fc1 LN(s)) * fc3 LN(s)))   fm_utils.py:333 / 365 with GatedMlp.forward (:142-144)."""

    @staticmethod
    def forward(ctx, x, ypend, nw, nb, w1, w3, w2, b1, b3, b2, eps):
        shape = x.shape
        D = shape[-1]
        x2, y2 = _prep_stream(x, ypend, D)
        H, Hp = w1.shape[0], _pad_rows(w1.shape[0])
        s2, h, mean, rstd = ops.add_layernorm_fwd(x2, y2, nw, nb, eps)
        bias13 = None
        if b1 is not None:
            bias13 = torch.zeros(2 * Hp, device=x2.device, dtype=torch.float32)
            bias13[:H] = b1
            bias13[Hp:Hp + H] = b3
        ab, g = ops.gemm(h, weight_bf16(w1, w3), epilogue=ops.EPI_SWIGLU, bias=bias13)           # [R, 2Hp], [R, Hp]
        w2b = weight_bf16(w2) if Hp == H else weight_bf16_padk(w2, Hp)      # padded gate columns are exactly zero
        y = ops.gemm(g, w2b, epilogue=ops.EPI_BF16, bias=b2, n_out=D)
        ctx.save_for_backward(s2, mean, rstd, h, ab, g, nw, w1, w3, w2)
        ctx.cfg = (shape, D, H, Hp, b1 is not None, b2 is not None, nb is not None and nb.requires_grad, ypend is not None)
        ctx.nb = nb
        return s2.view(shape), y.view(shape)

    @staticmethod
    def backward(ctx, g_s, g_y):
        s2, mean, rstd, h, ab, g, nw, w1, w3, w2 = ctx.saved_tensors
        shape, D, H, Hp, has_b13, has_b2, nb_grad, has_pend = ctx.cfg
        d2, db = _stream_grads(g_s, g_y, s2.shape[0], D)
        w2b = weight_bf16(w2)[:D] if Hp == H else weight_bf16_padk(w2, Hp)
        dg = ops.gemm(db, w2b, layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)              # [R, Hp]; padded columns are zero
        dw2 = None
        if w2.requires_grad:
            if Hp == H:
                dw2 = _wgrad(db, g, w2)                                                     # [D, H], into the gradient arena when there is one
            else:
                dw2 = _tn_gemm(db, g, None)[:, :H]                                         # [D, Hp] -> strided view (copied by autograd)
        db2 = ops.colsum_bf16(db) if has_b2 else None
        dab = ops.swiglu_bwd(ab, dg)
        dh = ops.gemm(dab, weight_bf16(w1, w3), layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)
        dw1 = dw3 = db1 = db3 = None
        if w1.requires_grad or w3.requires_grad:
            dw1, dw3 = _wgrad13(dab, h, w1, w3, H, Hp)
        if has_b13:
            dbias = ops.colsum_bf16(dab)
            db1, db3 = dbias[:H], dbias[Hp:Hp + H]
        dgamma = _acc_f32(nw, D, s2.device) if nw.requires_grad else None
        dbeta = _acc_f32(ctx.nb, D, s2.device) if nb_grad else None
        dx, dxb = ops.layernorm_bwd(dh, s2, nw, mean, rstd, dres=d2, want_bf16=has_pend, dgamma=dgamma, dbeta=dbeta)
        return dx.view(shape), dxb.view(shape) if has_pend else None, dgamma, dbeta, dw1, dw3, dw2, db1, db3, db2, None


class NormLinearResidualFn(torch.autograd.Function):
    """resid + bf16(LN(x + ypend) W^T + b):  decoder_proj_context(encoder_norm(x)) + encoder_emb  (fm.py:678-679)."""

    @staticmethod
    def forward(ctx, x, ypend, nw, nb, w, b, resid, eps):
        D = x.shape[-1]
        x2, y2 = _prep_stream(x, ypend, D)
        r2 = resid.reshape(-1, w.shape[0])
        if not r2.is_contiguous():
            r2 = r2.contiguous()
        s2, h, mean, rstd = ops.add_layernorm_fwd(x2, y2, nw, nb, eps)
        out = ops.gemm(h, weight_bf16(w), epilogue=ops.EPI_RESID, bias=b, resid=r2, n_out=w.shape[0])
        ctx.save_for_backward(s2, mean, rstd, h, nw, w)
        ctx.cfg = (x.shape, resid.shape, b is not None, nb is not None and nb.requires_grad, ypend is not None)
        ctx.nb = nb
        return out.view(resid.shape)

    @staticmethod
    def backward(ctx, dout):
        s2, mean, rstd, h, nw, w = ctx.saved_tensors
        xshape, rshape, has_b, nb_grad, has_pend = ctx.cfg
        D = s2.shape[1]
        d2 = dout.reshape(-1, w.shape[0])
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        db = ops.cast_bf16(d2) if d2.dtype == torch.float32 else d2
        dh = ops.gemm(db, weight_bf16(w)[:w.shape[0]], layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16)
        dw = _wgrad(db, h, w) if w.requires_grad else None
        dbias = ops.colsum_bf16(db) if has_b else None
        dgamma = _acc_f32(nw, D, s2.device) if nw.requires_grad else None
        dbeta = _acc_f32(ctx.nb, D, s2.device) if nb_grad else None
        dx, dxb = ops.layernorm_bwd(dh, s2, nw, mean, rstd, want_bf16=has_pend, dgamma=dgamma, dbeta=dbeta)
        return dx.view(xshape), dxb.view(xshape) if has_pend else None, dgamma, dbeta, dw, dbias, dout.view(rshape), None


class AddLayerNormFn(torch.autograd.Function):
    """LN(x + ypend) as bf16: the stack's final norm (fm.py:494 / 517) consuming the last pending branch."""

    @staticmethod
    def forward(ctx, x, ypend, nw, nb, eps):
        D = x.shape[-1]
        x2, y2 = _prep_stream(x, ypend, D)
        s2, h, mean, rstd = ops.add_layernorm_fwd(x2, y2, nw, nb, eps)
        ctx.save_for_backward(s2, mean, rstd, nw)
        ctx.cfg = (x.shape, nb is not None and nb.requires_grad, ypend is not None)
        ctx.nb = nb
        return h.view(x.shape)

    @staticmethod
    def backward(ctx, dh):
        s2, mean, rstd, nw = ctx.saved_tensors
        shape, nb_grad, has_pend = ctx.cfg
        D = s2.shape[1]
        dh2 = dh.reshape(-1, D)
        if not dh2.is_contiguous():
            dh2 = dh2.contiguous()
        dgamma = _acc_f32(nw, D, s2.device) if nw.requires_grad else None
        dbeta = _acc_f32(ctx.nb, D, s2.device) if nb_grad else None
        dx, dxb = ops.layernorm_bwd(dh2, s2, nw, mean, rstd, want_bf16=has_pend, dgamma=dgamma, dbeta=dbeta)
        return dx.view(shape), dxb.view(shape) if has_pend else None, dgamma, dbeta, None
