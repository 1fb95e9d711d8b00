"""B200-native building blocks of the 4M transformer, behind the reference's module surface.

Drop-in for `fourm/models/fm_utils.py` of apple/ml-4m (class names, constructor signatures, parameter names and
forward signatures are the reference's, file:line cited per class); the arithmetic runs in hand-written sm_100a
kernels through the b200fm C ABI (tcgen05 GEMMs with fused epilogues, fused attention, LayerNorm).  There is no
PyTorch/CPU fallback: CPU tensors or a missing libb200fm.so raise.

Sub-modules stay real `nn.Linear`s named qkv / q / kv / proj / fc1 / fc2 / fc3 (LoRA injection and `init_weights` key on
those).  When a child has been swapped for something else (e.g. a LoRA wrapper) the fused path calls the
child module instead of reading `.weight`.
"""
import torch
import torch.nn as nn

from b200fm import functional as BF


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


def softmax1(tensor):
    raise NotImplementedError("allow_zero_attn (softmax1) is not used by any shipped 4M config and has no B200 kernel yet")


def build_1d_sincos_posemb(max_len, embed_dim=1024, temperature=10000.):
    """Sine-cosine positional embeddings (reference fm_utils.py:32-44) -> [1, max_len, embed_dim]."""
    assert embed_dim % 2 == 0, 'Embed dimension must be divisible by 2 for 1D sin-cos position embedding'
    half = embed_dim // 2
    omega = 1. / (temperature ** (torch.arange(half, dtype=torch.float32) / half))
    ang = torch.einsum('n,d->nd', [torch.arange(max_len, dtype=torch.float32), omega])
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).unsqueeze(0)


def build_2d_sincos_posemb(h, w, embed_dim=1024, temperature=10000.0):
    """2D sine-cosine positional embeddings (reference fm_utils.py:46-61) -> [1, h*w, embed_dim]."""
    assert embed_dim % 4 == 0, 'Embed dimension must be divisible by 4 for 2D sin-cos position embedding'
    grid_w, grid_h = torch.meshgrid(torch.arange(w, dtype=torch.float32), torch.arange(h, dtype=torch.float32), indexing='ij')
    q = embed_dim // 4
    omega = 1. / (temperature ** (torch.arange(q, dtype=torch.float32) / q))
    out_w = torch.einsum('n,d->nd', [grid_w.reshape(-1), omega])
    out_h = torch.einsum('n,d->nd', [grid_h.reshape(-1), omega])
    return torch.cat([torch.sin(out_w), torch.cos(out_w), torch.sin(out_h), torch.cos(out_h)], dim=1).unsqueeze(0)


def drop_path(x, drop_prob: float = 0., training: bool = False):
    if drop_prob == 0. or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    random_tensor = (keep_prob + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
    return x.div(keep_prob) * random_tensor


class DropPath(nn.Module):
    """Stochastic depth (reference fm_utils.py:64-90); identity at rate 0, which is what every shipped config uses."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)

    def extra_repr(self) -> str:
        return 'p={}'.format(self.drop_prob)


def _plain_linear(m):
    return type(m) is nn.Linear


def _linear(m, x):
    """y = m(x) through the tcgen05 GEMM when m is a plain nn.Linear, else through the (wrapped) module itself."""
    if _plain_linear(m):
        return BF.LinearFn.apply(x, m.weight, m.bias)
    return m(x)


def _linear_residual(m, x, resid):
    if _plain_linear(m):
        return BF.LinearResidualFn.apply(x, m.weight, m.bias, resid)
    return resid + m(x)


class LayerNorm(nn.Module):
    """LayerNorm with optional bias (reference fm_utils.py:93-108; bias-free variants keep a zero `bias` buffer in the
    state_dict).  fp32 in -> fp32 out when called as a module; the fused blocks ask the kernel for bf16 directly."""

    def __init__(self, normalized_shape: int, eps=1e-5, bias=True):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        if bias:
            self.bias = nn.Parameter(torch.zeros(normalized_shape))
        else:
            self.register_buffer("bias", torch.zeros(normalized_shape))
        self.normalized_shape = (normalized_shape,)

    def forward(self, x, out_bf16: bool = False):
        return BF.layer_norm(x, self.weight, self.bias, self.eps, out_bf16)


def _norm_bf16(norm, x):
    """LayerNorm whose consumer is a GEMM: ask for the bf16 output (same rounding point as autocast's cast)."""
    if isinstance(norm, LayerNorm):
        return norm(x, out_bf16=True)
    if type(norm) is nn.LayerNorm and norm.elementwise_affine:
        return BF.layer_norm(x, norm.weight, norm.bias, norm.eps, True)
    return norm(x)


class Mlp(nn.Module):
    """fc2(act(fc1 x)) (reference fm_utils.py:111-126)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0., bias=True):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop = nn.Dropout(drop)

    def hidden(self, x):
        if _plain_linear(self.fc1) and type(self.act) is nn.GELU and getattr(self.act, "approximate", "none") == "none":
            return BF.MlpActFn.apply(x, self.fc1.weight, self.fc1.bias, "gelu")
        return self.act(_linear(self.fc1, x))

    def forward(self, x):
        return self.drop(_linear(self.fc2, self.hidden(x)))

    def forward_residual(self, x, resid):
        if self.drop.p == 0. or not self.training:
            return _linear_residual(self.fc2, self.hidden(x), resid)
        return resid + self.forward(x)


class GatedMlp(nn.Module):
    """SwiGLU feed-forward fc2(act(fc1 x) * fc3 x) (reference fm_utils.py:129-144)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.SiLU, bias=True):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = int(2 * (hidden_features or in_features) / 3)
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.fc3 = nn.Linear(in_features, hidden_features, bias=bias)

    def hidden(self, x):
        if _plain_linear(self.fc1) and _plain_linear(self.fc3) and type(self.act) is nn.SiLU:
            return BF.SwiGLUFn.apply(x, self.fc1.weight, self.fc3.weight, self.fc1.bias, self.fc3.bias)
        return self.act(_linear(self.fc1, x)) * _linear(self.fc3, x)

    def forward(self, x):
        return _linear(self.fc2, self.hidden(x))

    def forward_residual(self, x, resid):
        return _linear_residual(self.fc2, self.hidden(x), resid)


def _prep_mask(mask, B, Nq, Nk):
    """Reference masks are bool, True = masked, shape [B, 1|Nq, Nk] (fm_utils.py:167-169, 206-208)."""
    if mask is None:
        return None
    if mask.dim() == 2:
        mask = mask[:, None, :]
    return mask


class Attention(nn.Module):
    """Multi-head self-attention (reference fm_utils.py:147-180) on the fused sm_100a attention kernel."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, proj_bias=True, attn_drop=0., proj_drop=0., allow_zero_attn=False):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.allow_zero_attn = allow_zero_attn
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.proj_drop = nn.Dropout(proj_drop)

    def attend(self, x, mask=None):
        B, N, C = x.shape
        _check_attn_support(self, C)
        qkv = _linear(self.qkv, x).reshape(B * N, 3 * C)
        return BF.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], _prep_mask(mask, B, N, N), B, self.num_heads, N, N,
                            self.scale).view(B, N, C)

    def forward(self, x, mask=None):
        return self.proj_drop(_linear(self.proj, self.attend(x, mask)))

    def forward_residual(self, x, resid, mask=None):
        if self.proj_drop.p == 0. or not self.training:
            return _linear_residual(self.proj, self.attend(x, mask), resid)
        return resid + self.forward(x, mask)


class CrossAttention(nn.Module):
    """Multi-head cross-attention (reference fm_utils.py:182-219)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, proj_bias=True, attn_drop=0., proj_drop=0., allow_zero_attn=False):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.allow_zero_attn = allow_zero_attn
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.proj_drop = nn.Dropout(proj_drop)

    def attend(self, x, context, mask=None):
        B, N, C = x.shape
        M = context.shape[1]
        _check_attn_support(self, C)
        q = _linear(self.q, x).reshape(B * N, C)
        kv = _linear(self.kv, context).reshape(B * M, 2 * C)
        return BF.attention(q, kv[:, :C], kv[:, C:], _prep_mask(mask, B, N, M), B, self.num_heads, N, M, self.scale).view(B, N, C)

    def forward(self, x, context, mask=None):
        return self.proj_drop(_linear(self.proj, self.attend(x, context, mask)))

    def forward_residual(self, x, context, resid, mask=None):
        if self.proj_drop.p == 0. or not self.training:
            return _linear_residual(self.proj, self.attend(x, context, mask), resid)
        return resid + self.forward(x, context, mask)


def _check_attn_support(mod, C):
    if C // mod.num_heads != 64:
        raise NotImplementedError(f"b200fm attention kernels are specialised for head_dim 64 (got {C // mod.num_heads}); "
                                  "every reference preset uses 64 (fm.py:840-1130)")
    if mod.allow_zero_attn:
        raise NotImplementedError("allow_zero_attn=True (softmax1) has no B200 kernel yet")
    if mod.attn_drop.p != 0. and mod.training:
        raise NotImplementedError("attention dropout > 0 is not supported by the fused kernel (reference default is 0)")


class NormAttention(Attention):
    """Self-attention with LayerNorm on q and k per head (reference fm_utils.py:222-261; qk_norm presets): the packed qkv GEMM,
    a per-head LayerNorm kernel on the q and k column slices, then the fused attention kernel."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, proj_bias=True, norm_layer=nn.LayerNorm, attn_drop=0., proj_drop=0.,
                 allow_zero_attn=False):
        super().__init__(dim, num_heads, qkv_bias, proj_bias, attn_drop, proj_drop, allow_zero_attn)
        head_dim = dim // num_heads
        self.q_norm = norm_layer(head_dim)
        self.k_norm = norm_layer(head_dim)

    def attend(self, x, mask=None):
        B, N, C = x.shape
        _check_attn_support(self, C)
        qkv = _linear(self.qkv, x).reshape(B * N, 3 * C)
        q = BF.head_norm(qkv[:, :C], self.num_heads, self.q_norm)
        k = BF.head_norm(qkv[:, C:2 * C], self.num_heads, self.k_norm)
        return BF.attention(q, k, qkv[:, 2 * C:], _prep_mask(mask, B, N, N), B, self.num_heads, N, N, self.scale).view(B, N, C)


class NormCrossAttention(CrossAttention):
    """Cross-attention with q/k LayerNorm (reference fm_utils.py:264-307), see NormAttention."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, proj_bias=True, norm_layer=nn.LayerNorm, attn_drop=0., proj_drop=0.,
                 allow_zero_attn=False):
        super().__init__(dim, num_heads, qkv_bias, proj_bias, attn_drop, proj_drop, allow_zero_attn)
        head_dim = dim // num_heads
        self.q_norm = norm_layer(head_dim)
        self.k_norm = norm_layer(head_dim)

    def attend(self, x, context, mask=None):
        B, N, C = x.shape
        M = context.shape[1]
        _check_attn_support(self, C)
        q = BF.head_norm(_linear(self.q, x).reshape(B * N, C), self.num_heads, self.q_norm)
        kv = _linear(self.kv, context).reshape(B * M, 2 * C)
        k = BF.head_norm(kv[:, :C], self.num_heads, self.k_norm)
        return BF.attention(q, k, kv[:, C:], _prep_mask(mask, B, N, M), B, self.num_heads, N, M, self.scale).view(B, N, C)


def _residual_ok(block):
    return isinstance(block.drop_path, nn.Identity) or not block.training


def _norm_params(norm):
    """(weight, bias, eps) of a LayerNorm the kernels can run, else None."""
    if isinstance(norm, LayerNorm) or (type(norm) is nn.LayerNorm and norm.elementwise_affine):
        return norm.weight, norm.bias, norm.eps
    return None


def _plain_attention(a):
    return (type(a) is Attention and _plain_linear(a.qkv) and _plain_linear(a.proj) and not a.allow_zero_attn
            and (a.proj_drop.p == 0. or not a.training) and (a.attn_drop.p == 0. or not a.training))


def _plain_cross_attention(a):
    return (type(a) is CrossAttention and _plain_linear(a.q) and _plain_linear(a.kv) and _plain_linear(a.proj)
            and not a.allow_zero_attn and (a.proj_drop.p == 0. or not a.training) and (a.attn_drop.p == 0. or not a.training))


def _plain_gated_mlp(m):
    return (type(m) is GatedMlp and _plain_linear(m.fc1) and _plain_linear(m.fc2) and _plain_linear(m.fc3) and type(m.act) is nn.SiLU)


def _self_attn_sublayer(norm, attn, x, ypend, mask):
    """(x, pending bf16 branch) -> (s = x + pending, new pending branch) as ONE autograd node when every piece is a plain
    module; None otherwise."""
    np_ = _norm_params(norm)
    if np_ is None or not _plain_attention(attn) or x.shape[-1] // attn.num_heads != 64 or x.dtype != torch.float32:
        return None
    B, N, _ = x.shape
    return BF.SelfAttnSubLayerFn.apply(x, ypend, _prep_mask(mask, B, N, N), np_[0], np_[1], attn.qkv.weight, attn.qkv.bias,
                                       attn.proj.weight, attn.proj.bias, np_[2], attn.num_heads, attn.scale)


def _cross_attn_sublayer(qnorm, cnorm, attn, x, ypend, context, mask):
    qp, cp = _norm_params(qnorm), _norm_params(cnorm)
    if qp is None or cp is None or not _plain_cross_attention(attn) or x.shape[-1] // attn.num_heads != 64:
        return None
    if x.dtype != torch.float32 or context.dtype != torch.float32:
        return None
    B, N, _ = x.shape
    return BF.CrossAttnSubLayerFn.apply(x, ypend, context, _prep_mask(mask, B, N, context.shape[1]), qp[0], qp[1], cp[0], cp[1],
                                        attn.q.weight, attn.q.bias, attn.kv.weight, attn.kv.bias, attn.proj.weight, attn.proj.bias,
                                        qp[2], cp[2], attn.num_heads, attn.scale)


def _mlp_sublayer(norm, mlp, x, ypend):
    np_ = _norm_params(norm)
    if np_ is None or not _plain_gated_mlp(mlp) or x.dtype != torch.float32:
        return None
    return BF.GatedMlpSubLayerFn.apply(x, ypend, np_[0], np_[1], mlp.fc1.weight, mlp.fc3.weight, mlp.fc2.weight, mlp.fc1.bias,
                                       mlp.fc3.bias, mlp.fc2.bias, np_[2])


def _settle(x, ypend):
    """Materialise the stream: x + pending branch (only where a caller needs the plain tensor)."""
    return x if ypend is None else x + ypend.float()


class Block(nn.Module):
    """Pre-norm encoder block x += attn(norm1 x); x += mlp(norm2 x) (reference fm_utils.py:310-334).
    The residual adds are fused into the proj / fc2 GEMM epilogues, the norms emit the bf16 GEMM operand directly."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=True, proj_bias=True, mlp_bias=True, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, gated_mlp=False, qk_norm=False, allow_zero_attn=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        attn_cls = NormAttention if qk_norm else Attention
        kw = dict(norm_layer=norm_layer) if qk_norm else {}
        self.attn = attn_cls(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop, proj_drop=drop,
                             allow_zero_attn=allow_zero_attn, **kw)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        if not gated_mlp:
            self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, bias=mlp_bias, drop=drop)
        else:
            self.mlp = GatedMlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, bias=mlp_bias)

    def forward_pending(self, x, ypend, mask=None):
        """(stream, pending bf16 branch output) -> (stream, pending): the residual adds are deferred into the next norm
        kernel.  FourM's stacks chain blocks through this; `forward` is the reference-shaped wrapper."""
        if not _residual_ok(self):
            return self.forward(_settle(x, ypend), mask), None
        r = _self_attn_sublayer(self.norm1, self.attn, x, ypend, mask)
        if r is None:
            x = _settle(x, ypend)
            x, ypend = self.attn.forward_residual(_norm_bf16(self.norm1, x), x, mask), None
        else:
            x, ypend = r
        r = _mlp_sublayer(self.norm2, self.mlp, x, ypend)
        if r is None:
            x = _settle(x, ypend)
            x, ypend = self.mlp.forward_residual(_norm_bf16(self.norm2, x), x), None
        else:
            x, ypend = r
        return x, ypend

    def forward(self, x, mask=None):
        if _residual_ok(self):
            return _settle(*self.forward_pending(x, None, mask))
        x = x + self.drop_path(self.attn(_norm_bf16(self.norm1, x), mask))
        x = x + self.drop_path(self.mlp(_norm_bf16(self.norm2, x)))
        return x


class DecoderBlock(nn.Module):
    """Decoder block: self-attention, cross-attention over the (per-layer re-normalised) context, MLP
    (reference fm_utils.py:337-366)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=True, proj_bias=True, mlp_bias=True, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, gated_mlp=False, qk_norm=False, allow_zero_attn=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        kw = dict(norm_layer=norm_layer) if qk_norm else {}
        sa_cls, xa_cls = (NormAttention, NormCrossAttention) if qk_norm else (Attention, CrossAttention)
        self.self_attn = sa_cls(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                                proj_drop=drop, allow_zero_attn=allow_zero_attn, **kw)
        self.cross_attn = xa_cls(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                                 proj_drop=drop, allow_zero_attn=allow_zero_attn, **kw)
        self.query_norm = norm_layer(dim)
        self.context_norm = norm_layer(dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        if not gated_mlp:
            self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, bias=mlp_bias, drop=drop)
        else:
            self.mlp = GatedMlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, bias=mlp_bias)

    def forward_pending(self, x, ypend, context, sa_mask=None, xa_mask=None):
        """See Block.forward_pending."""
        if not _residual_ok(self):
            return self.forward(_settle(x, ypend), context, sa_mask, xa_mask), None
        r = _self_attn_sublayer(self.norm1, self.self_attn, x, ypend, sa_mask)
        if r is None:
            x = _settle(x, ypend)
            x, ypend = self.self_attn.forward_residual(_norm_bf16(self.norm1, x), x, sa_mask), None
        else:
            x, ypend = r
        r = _cross_attn_sublayer(self.query_norm, self.context_norm, self.cross_attn, x, ypend, context, xa_mask)
        if r is None:
            x = _settle(x, ypend)
            x, ypend = self.cross_attn.forward_residual(_norm_bf16(self.query_norm, x), _norm_bf16(self.context_norm, context), x, xa_mask), None
        else:
            x, ypend = r
        r = _mlp_sublayer(self.norm2, self.mlp, x, ypend)
        if r is None:
            x = _settle(x, ypend)
            x, ypend = self.mlp.forward_residual(_norm_bf16(self.norm2, x), x), None
        else:
            x, ypend = r
        return x, ypend

    def forward(self, x, context, sa_mask=None, xa_mask=None):
        if _residual_ok(self):
            return _settle(*self.forward_pending(x, None, context, sa_mask, xa_mask))
        x = x + self.drop_path(self.self_attn(_norm_bf16(self.norm1, x), sa_mask))
        x = x + self.drop_path(self.cross_attn(_norm_bf16(self.query_norm, x), _norm_bf16(self.context_norm, context), xa_mask))
        x = x + self.drop_path(self.mlp(_norm_bf16(self.norm2, x)))
        return x


class CrossAttentionBlock(nn.Module):
    """Cross-attention + MLP block (reference fm_utils.py:369-387)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, gated_mlp=False, allow_zero_attn=False):
        super().__init__()
        self.cross_attn = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop,
                                         allow_zero_attn=allow_zero_attn)
        self.query_norm = norm_layer(dim)
        self.context_norm = norm_layer(dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        if not gated_mlp:
            self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        else:
            self.mlp = GatedMlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer)

    def forward(self, x, context, xa_mask=None, **kwargs):
        x = x + self.drop_path(self.cross_attn(_norm_bf16(self.query_norm, x), _norm_bf16(self.context_norm, context), xa_mask))
        x = x + self.drop_path(self.mlp(_norm_bf16(self.norm2, x)))
        return x
