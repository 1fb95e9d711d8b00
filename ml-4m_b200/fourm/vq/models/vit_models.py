"""ViT encoder / decoder of the 4M VQ tokenizers on the B200 kernels.

Drop-in for the inference surface of `fourm/vq/models/vit_models.py` (apple/ml-4m): `ViTEncoder`, `ViTDecoder`, `Block`,
`Attention`, `Mlp` and the `vit_{s,b,l}_{enc,dec}` factories keep their constructor arguments and parameter names
(`pos_emb`, `proj`, `blocks.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}`, `norm_mlp`, `post_mlp`, `out_proj`), so
tokenizer checkpoints load unchanged.  Patchify is a gather + tcgen05 GEMM (the k=16, s=16 Conv2d of the reference is
exactly that), blocks run LayerNorm -> bf16 GEMMs (bias + GELU in the epilogue) -> fused attention (no mask).
Contractions are bf16 with fp32 accumulation; the reference runs this path in fp32 -- tolerance stated in the tests."""
import math
import weakref
from functools import partial
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from b200fm import functional as BF
from b200fm import ops

XFORMERS_AVAILABLE = False      # the fused sm_100a attention kernel is always used


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


def build_2d_sincos_posemb(h, w, embed_dim=1024, temperature=10000.):
    """Reference vit_models.py:38-52 -> [1, embed_dim, h, w]."""
    assert embed_dim % 4 == 0, 'Embed dimension must be divisible by 4 for 2D sin-cos position embedding'
    grid_w, grid_h = torch.meshgrid(torch.arange(w, dtype=torch.float32), torch.arange(h, dtype=torch.float32), indexing='ij')
    q = embed_dim // 4
    omega = 1. / (temperature ** (torch.arange(q, dtype=torch.float32) / q))
    ow = grid_w.flatten()[:, None] * omega[None]
    oh = grid_h.flatten()[:, None] * omega[None]
    pe = torch.cat([ow.sin(), ow.cos(), oh.sin(), oh.cos()], dim=1)
    return pe.reshape(1, h, w, embed_dim).permute(0, 3, 1, 2).contiguous()


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class Mlp(nn.Module):
    """Reference vit_models.py:145-163."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def hidden(self, x):
        if type(self.act) is nn.GELU:
            return BF.MlpActFn.apply(x, self.fc1.weight, self.fc1.bias, "gelu")
        if type(self.act) is nn.Tanh:
            return BF.MlpActFn.apply(x, self.fc1.weight, self.fc1.bias, "tanh")
        return self.act(BF.LinearFn.apply(x, self.fc1.weight, self.fc1.bias))

    def forward(self, x):
        return self.drop(BF.LinearFn.apply(self.hidden(x), self.fc2.weight, self.fc2.bias))

    def forward_residual(self, x, resid):
        return BF.LinearResidualFn.apply(self.hidden(x), self.fc2.weight, self.fc2.bias, resid)


class Attention(nn.Module):
    """Reference vit_models.py:165-197 (no mask)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def attend(self, x):
        B, N, C = x.shape
        if C // self.num_heads != 64:
            raise NotImplementedError("b200fm attention kernels are specialised for head_dim 64")
        qkv = BF.LinearFn.apply(x, self.qkv.weight, self.qkv.bias).reshape(B * N, 3 * C)
        return BF.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], None, B, self.num_heads, N, N, self.scale).view(B, N, C)

    def forward(self, x):
        return self.proj_drop(BF.LinearFn.apply(self.attend(x), self.proj.weight, self.proj.bias))

    def forward_residual(self, x, resid):
        return BF.LinearResidualFn.apply(self.attend(x), self.proj.weight, self.proj.bias, resid)


class Block(nn.Module):
    """Reference vit_models.py:232-246."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.norm2 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        if drop_path > 0.:
            raise NotImplementedError("drop_path > 0 is not supported by the B200 ViT blocks")
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, **kwargs):
        if BF.is_precise():
            return self._forward_precise(x)
        x = self.attn.forward_residual(BF.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, True), x)
        x = self.mlp.forward_residual(BF.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps, True), x)
        return x

    def _forward_precise(self, x):
        """The reference's fp32 block (vit_models.py:243-246 without autocast): fp32 LayerNorm, fp32-faithful linears, fp32 attention."""
        B, N, C = x.shape
        a = self.attn
        h = BF.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, False)
        qkv = BF.linear_f32(h, a.qkv.weight, a.qkv.bias).reshape(B * N, 3 * C)
        o = ops.attention_f32(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, a.num_heads, N, N, None, a.scale).view(B, N, C)
        x = x + BF.linear_f32(o, a.proj.weight, a.proj.bias)
        h = BF.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps, False)
        return x + BF.linear_f32(self.mlp.act(BF.linear_f32(h, self.mlp.fc1.weight, self.mlp.fc1.bias)), self.mlp.fc2.weight, self.mlp.fc2.bias)


def _init_vit(module):
    """Reference vit_models.py:430-459: xavier-uniform Linear (per-Q/K/V fan-out for qkv), LN 1/0, conv proj like a Linear."""
    for name, m in module.named_modules():
        if isinstance(m, nn.Linear):
            if 'qkv' in name:
                val = math.sqrt(6. / float(m.weight.shape[0] // 3 + m.weight.shape[1]))
                nn.init.uniform_(m.weight, -val, val)
            elif 'kv' in name:
                val = math.sqrt(6. / float(m.weight.shape[0] // 2 + m.weight.shape[1]))
                nn.init.uniform_(m.weight, -val, val)
            else:
                nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Conv2d) and (name == 'proj' or name.endswith('.proj')):
            w = m.weight.data
            nn.init.xavier_uniform_(w.view([w.shape[0], -1]))


class _ConvAsLinear:
    """k = s = P Conv2d weight [O, C, P, P] viewed as a Linear over '(ph pw c)'-ordered patches (cached per parameter version;
    the entry remembers WHICH parameter it mirrors: id() values and device addresses are recycled when models are re-created)."""
    _cache = {}

    @classmethod
    def weight(cls, conv):
        p = conv.weight
        key = id(p)
        ver = (p._version, p.data_ptr())
        hit = cls._cache.get(key)
        if hit is None or hit[0] != ver or hit[2]() is not p:
            w = p.detach().permute(0, 2, 3, 1).reshape(p.shape[0], -1).float().contiguous()
            hit = (ver, ops.cast_bf16(w), weakref.ref(p, lambda _r, key=key: cls._cache.pop(key, None)))   # evicted with its parameter
            cls._cache[key] = hit
        return hit[1]


class ViTEncoder(nn.Module):
    """Image / feature map -> latent feature map [B, dim_tokens, N_H, N_W] (reference vit_models.py:338-501)."""

    def __init__(self, *, in_channels: int = 3, patch_size: int = 16, resolution: int = 256, dim_tokens: int = 768, depth: int = 12,
                 num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True, drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6), sincos_pos_emb: bool = True,
                 learnable_pos_emb: bool = False, patch_proj: bool = True, post_mlp: bool = False, ckpt_path: Optional[str] = None,
                 **ignore_kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.P_H, self.P_W = pair(patch_size)
        self.H, self.W = pair(resolution)
        self.dim_tokens, self.patch_proj = dim_tokens, patch_proj
        assert (self.H % self.P_H == 0) and (self.W % self.P_W == 0), \
            f'Image sizes {self.H}x{self.W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
        N_H, N_W = self.H // self.P_H, self.W // self.P_W
        if sincos_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=N_H, w=N_W, embed_dim=dim_tokens), requires_grad=learnable_pos_emb)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, dim_tokens, N_H, N_W))
            trunc_normal_(self.pos_emb, std=0.02)
        k = (self.P_H, self.P_W) if patch_proj else 1
        self.proj = nn.Conv2d(in_channels=in_channels, out_channels=dim_tokens, kernel_size=k, stride=k)
        self.blocks = nn.Sequential(*[Block(dim=dim_tokens, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                                            attn_drop=attn_drop_rate, drop_path=0., norm_layer=norm_layer) for _ in range(depth)])
        if post_mlp:
            self.norm_mlp = norm_layer(dim_tokens)
            self.post_mlp = Mlp(dim_tokens, int(mlp_ratio * dim_tokens), act_layer=nn.Tanh)
        _init_vit(self)
        if ckpt_path is not None:
            raise NotImplementedError("MAE checkpoint initialisation (ckpt_path) is a training-time feature of the reference")

    def get_num_layers(self) -> int:
        return len(self.blocks)

    def tokens(self, x: torch.Tensor):
        """[B, C, H, W] -> fp32 token stream [B, N, D] after the transformer (and the post-MLP), plus (N_H, N_W)."""
        B, C, H, W = x.shape
        if BF.is_precise():
            return self._tokens_precise(x)
        train_proj = self.patch_proj and torch.is_grad_enabled() and self.proj.weight.requires_grad
        if self.patch_proj:
            assert (H % self.P_H == 0) and (W % self.P_W == 0), f'Image sizes {H}x{W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
            assert self.P_H == self.P_W, "square patches only on the B200 path"
            N_H, N_W = H // self.P_H, W // self.P_W
            patches = None if train_proj else ops.patchify(x.float().contiguous(), self.P_H)     # [B*N, P*P*C] bf16
        else:
            N_H, N_W = H, W
            patches = ops.cast_bf16(x.float().permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous())
        pe = self.pos_emb
        if pe.shape[-2:] != (N_H, N_W):
            pe = F.interpolate(pe, size=(N_H, N_W), mode='bicubic', align_corners=False)   # reference :486 (identity when sizes match)
        pe = pe.flatten(2).transpose(1, 2).float().expand(B, -1, -1).reshape(B * N_H * N_W, self.dim_tokens).contiguous()
        w = _ConvAsLinear.weight(self.proj)
        if train_proj:
            # tokenizer training (VQVAE.forward): same kernels, plus the weight / bias gradients of the patch projection
            t = BF.PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, pe, w, self.P_H).view(B, N_H * N_W, self.dim_tokens)
        else:
            bias = self.proj.bias.detach().float() if self.proj.bias is not None else None
            t = ops.gemm(patches, w, epilogue=ops.EPI_RESID, bias=bias, resid=pe).view(B, N_H * N_W, self.dim_tokens)
        t = self.blocks(t)
        if hasattr(self, 'post_mlp'):
            h = BF.layer_norm(t, self.norm_mlp.weight, self.norm_mlp.bias, self.norm_mlp.eps, True)
            t = self.post_mlp.forward_residual(h, t)
        return t, (N_H, N_W)

    def _tokens_precise(self, x):
        """`tokens` with fp32-faithful arithmetic (what the reference computes when it runs without autocast, save_vq_tokens.py:288)."""
        B, C, H, W = x.shape
        x = x.float()
        if self.patch_proj:
            assert (H % self.P_H == 0) and (W % self.P_W == 0), f'Image sizes {H}x{W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
            N_H, N_W = H // self.P_H, W // self.P_W
            patches = x.reshape(B, C, N_H, self.P_H, N_W, self.P_W).permute(0, 2, 4, 3, 5, 1).reshape(B * N_H * N_W, self.P_H * self.P_W * C)
            w_lin = self.proj.weight.detach().permute(0, 2, 3, 1).reshape(self.proj.weight.shape[0], -1)
        else:
            N_H, N_W = H, W
            patches = x.permute(0, 2, 3, 1).reshape(B * H * W, C)
            w_lin = self.proj.weight.detach().reshape(self.proj.weight.shape[0], -1)
        Kp = (patches.shape[1] + 7) // 8 * 8                     # the GEMM wants 16-byte operand rows
        if Kp != patches.shape[1]:
            patches, w_lin = F.pad(patches, (0, Kp - patches.shape[1])), F.pad(w_lin, (0, Kp - w_lin.shape[1]))
        pe = self.pos_emb
        if pe.shape[-2:] != (N_H, N_W):
            pe = F.interpolate(pe, size=(N_H, N_W), mode='bicubic', align_corners=False)
        t = BF.linear_f32(patches.contiguous(), w_lin, self.proj.bias, cache=False).view(B, N_H * N_W, self.dim_tokens)
        t = t + pe.flatten(2).transpose(1, 2).float()
        t = self.blocks(t.contiguous())
        if hasattr(self, 'post_mlp'):
            h = BF.layer_norm(t, self.norm_mlp.weight, self.norm_mlp.bias, self.norm_mlp.eps, False)
            m = self.post_mlp
            t = t + BF.linear_f32(m.act(BF.linear_f32(h, m.fc1.weight, m.fc1.bias)), m.fc2.weight, m.fc2.bias)
        return t, (N_H, N_W)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        t, (N_H, N_W) = self.tokens(x)
        B = t.shape[0]
        return t.transpose(1, 2).reshape(B, self.dim_tokens, N_H, N_W)


def _vit_factory(dim, depth, heads):
    def make(in_channels, patch_size=16, resolution=256, patch_proj=True, post_mlp=False, **kw):
        return ViTEncoder(in_channels=in_channels, patch_size=patch_size, resolution=resolution, dim_tokens=dim, depth=depth,
                          num_heads=heads, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                          patch_proj=patch_proj, post_mlp=post_mlp, **kw)
    return make


vit_s_enc = _vit_factory(512, 8, 8)        # reference vit_models.py:664-692
vit_b_enc = _vit_factory(768, 12, 12)      # :695-725
vit_l_enc = _vit_factory(1024, 24, 16)     # :728-759


class ViTDecoder(nn.Module):
    """Latent feature map [B, dim_tokens, N_H, N_W] -> image [B, C, H, W] (reference vit_models.py:504-659): positional
    embedding, the same transformer blocks as the encoder, optional Tanh post-MLP, a Linear to P*P*C per token and the
    '(c ph pw)' un-patchify.  Tokenizer-training side (VQVAE.forward / decode_quant)."""

    def __init__(self, *, out_channels: int = 3, patch_size: int = 16, resolution: int = 256, dim_tokens: int = 768, depth: int = 12,
                 num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True, drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6), sincos_pos_emb: bool = True,
                 learnable_pos_emb: bool = False, patch_proj: bool = True, post_mlp: bool = False, out_conv: bool = False, **ignore_kwargs):
        super().__init__()
        self.out_channels = out_channels
        self.P_H, self.P_W = pair(patch_size)
        self.H, self.W = pair(resolution)
        self.dim_tokens, self.patch_proj = dim_tokens, patch_proj
        assert (self.H % self.P_H == 0) and (self.W % self.P_W == 0), \
            f'Image sizes {self.H}x{self.W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
        N_H, N_W = self.H // self.P_H, self.W // self.P_W
        if sincos_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=N_H, w=N_W, embed_dim=dim_tokens), requires_grad=learnable_pos_emb)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, dim_tokens, N_H, N_W))
            trunc_normal_(self.pos_emb, std=0.02)
        if drop_path_rate > 0.:
            raise NotImplementedError("drop_path > 0 is not supported by the B200 ViT blocks")
        self.blocks = nn.Sequential(*[Block(dim=dim_tokens, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                                            attn_drop=attn_drop_rate, drop_path=0., norm_layer=norm_layer) for _ in range(depth)])
        if post_mlp:
            self.norm_mlp = norm_layer(dim_tokens)
            self.post_mlp = Mlp(dim_tokens, int(mlp_ratio * dim_tokens), act_layer=nn.Tanh)
        self.out_proj = nn.Linear(dim_tokens, out_channels * self.P_H * self.P_W if patch_proj else out_channels)
        if out_conv:
            raise NotImplementedError("out_conv (ConvNeXt blocks after the decoder) is not on the B200 path")
        _init_vit(self)

    def get_num_layers(self) -> int:
        return len(self.blocks)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, D, N_H, N_W = x.shape
        t = x.flatten(2).transpose(1, 2).float()
        pe = self.pos_emb
        if pe.shape[-2:] != (N_H, N_W):
            pe = F.interpolate(pe, size=(N_H, N_W), mode='bicubic', align_corners=False)
        t = t + pe.flatten(2).transpose(1, 2).float()
        t = self.blocks(t.contiguous())
        if hasattr(self, 'post_mlp'):
            h = BF.layer_norm(t, self.norm_mlp.weight, self.norm_mlp.bias, self.norm_mlp.eps, True)
            t = self.post_mlp.forward_residual(h, t)
        y = BF.LinearFn.apply(t, self.out_proj.weight, self.out_proj.bias)                # bf16, like nn.Linear under autocast
        ph, pw = (self.P_H, self.P_W) if self.patch_proj else (1, 1)
        y = y.view(B, N_H, N_W, self.out_channels, ph, pw).permute(0, 3, 1, 4, 2, 5)         # 'b (nh nw) (c ph pw) -> b c (nh ph) (nw pw)'
        return y.reshape(B, self.out_channels, N_H * ph, N_W * pw)


def _vit_dec_factory(dim, depth, heads):
    def make(out_channels, patch_size=16, resolution=256, patch_proj=True, post_mlp=False, out_conv=False, **kw):
        return ViTDecoder(out_channels=out_channels, patch_size=patch_size, resolution=resolution, dim_tokens=dim, depth=depth,
                          num_heads=heads, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                          patch_proj=patch_proj, post_mlp=post_mlp, out_conv=out_conv, **kw)
    return make


vit_s_dec = _vit_dec_factory(512, 8, 8)        # reference vit_models.py:762-792
vit_b_dec = _vit_dec_factory(768, 12, 12)      # :795-825
vit_l_dec = _vit_dec_factory(1024, 24, 16)     # :828-858
