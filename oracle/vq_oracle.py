"""CPU oracle for the VQ tokenizer forward  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional torch restatement of `fourm.vq.VQ.encode/tokenize` (ViT encoder -> 1x1 conv -> codebook arg-max).
Pinned against the unmodified reference via tests/golden/vq_golden.pt (see make_golden.py).  The codebook
scan also has a plain-C restatement in oracle/vq_argmax.c (bit-level check of the tie-break and fp32 order).
Citations are relative to /root/reference.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

VIT_PRESETS = {   # fourm/vq/models/vit_models.py:664-759
    "vit_s_enc": dict(dim=512, depth=8, heads=8),
    "vit_b_enc": dict(dim=768, depth=12, heads=12),
    "vit_l_enc": dict(dim=1024, depth=24, heads=16),
}


def sincos_2d_grid(h: int, w: int, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """fourm/vq/models/vit_models.py:38-52 build_2d_sincos_posemb -> [1, dim, h, w]."""
    q = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(q, dtype=torch.float32) / q))
    p = torch.arange(h * w)
    gw = (p // h).to(torch.float32)     # meshgrid(w, h, indexing='ij').flatten()
    gh = (p % h).to(torch.float32)
    aw, ah = gw[:, None] * omega[None], gh[:, None] * omega[None]
    pe = torch.cat([aw.sin(), aw.cos(), ah.sin(), ah.cos()], dim=1)        # [(h w), dim] flat index read as (h w)
    return pe.reshape(1, h, w, dim).permute(0, 3, 1, 2).contiguous()


def vit_block(x, sd, pfx, heads):
    """vit_models.py:232-246 Block + :165-197 Attention (naive branch) + :145-163 Mlp (GELU)."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], 1e-6)
    qkv = F.linear(h, sd[pfx + "attn.qkv.weight"], sd.get(pfx + "attn.qkv.bias"))
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * (C // heads) ** -0.5).softmax(dim=-1)
    o = (attn @ qkv[2]).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(o, sd[pfx + "attn.proj.weight"], sd[pfx + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], 1e-6)
    h = F.gelu(F.linear(h, sd[pfx + "mlp.fc1.weight"], sd[pfx + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[pfx + "mlp.fc2.weight"], sd[pfx + "mlp.fc2.bias"])


def vit_encoder(img, sd, pfx, cfg, patch: int, post_mlp: bool):
    """vit_models.py:465-501 ViTEncoder.forward -> [B, dim, Hq, Wq]."""
    B, C, H, W = img.shape
    if H % patch or W % patch:
        raise AssertionError(f"Image sizes {H}x{W} must be divisible by patch sizes {patch}x{patch}")
    nh, nw = H // patch, W // patch
    x = F.conv2d(img, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"], stride=patch)        # :482
    x = x.flatten(2).transpose(1, 2)
    pe = F.interpolate(sd[pfx + "pos_emb"], size=(nh, nw), mode="bicubic", align_corners=False)   # :486
    x = x + pe.flatten(2).transpose(1, 2)
    for i in range(cfg["depth"]):
        x = vit_block(x, sd, f"{pfx}blocks.{i}.", cfg["heads"])
    if post_mlp:   # :494-496, fp32 with autocast disabled, Tanh MLP
        with torch.autocast("cpu", enabled=False):
            xf = x.float()
            h = F.layer_norm(xf, (xf.shape[-1],), sd[pfx + "norm_mlp.weight"], sd[pfx + "norm_mlp.bias"], 1e-6)
            h = torch.tanh(F.linear(h, sd[pfx + "post_mlp.fc1.weight"], sd[pfx + "post_mlp.fc1.bias"]))
            x = xf + F.linear(h, sd[pfx + "post_mlp.fc2.weight"], sd[pfx + "post_mlp.fc2.bias"])
    return x.transpose(1, 2).reshape(B, -1, nh, nw)


def cosine_scan(z: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """quantize_lucid.py:388-403 CosineSimCodebook.forward (eval): fp32, l2norm both sides, argmax of
    z_n @ e_n^T; ties -> lowest index (torch.argmax).  z [n,d], embed [K,d] -> int64 [n]."""
    zn = F.normalize(z.float(), p=2, dim=-1)
    en = F.normalize(embed.float(), p=2, dim=-1)
    return (zn @ en.t()).argmax(dim=-1)


def euclidean_scan(z: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """quantize_lucid.py:263-280 EuclideanCodebook.forward (eval): argmax of -(|z|^2 - 2 z.e + |e|^2)."""
    z = z.float()
    e = embed.float().t()
    dist = -(z.pow(2).sum(1, keepdim=True) - 2 * z @ e + e.pow(2).sum(0, keepdim=True))
    return dist.argmax(dim=-1)


def scan_scores(z, embed, cosine: bool) -> torch.Tensor:
    """fp64 score matrix used by tests to accept index differences only on genuine near-ties."""
    z, e = z.double(), embed.double()
    if cosine:
        return F.normalize(z, dim=-1) @ F.normalize(e, dim=-1).t()
    return -(z.pow(2).sum(1, keepdim=True) - 2 * z @ e.t() + e.pow(2).sum(1)[None])


def vq_encode(img, sd: Dict[str, torch.Tensor], enc_type: str, patch: int, norm_codes: bool, post_mlp: bool,
              norm_latents: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """vqvae.py:302-318 VQ.encode (eval) + quantize_lucid.py:504-568 VectorQuantize.forward (heads=1, no
    projection).  Returns (quant [B,d,Hq,Wq], tokens int64 [B,Hq,Wq], latents [B,d,Hq,Wq])."""
    h = vit_encoder(img, sd, "encoder.", VIT_PRESETS[enc_type], patch, post_mlp)
    h = F.conv2d(h, sd["quant_proj.weight"], sd["quant_proj.bias"])                      # vqvae.py:316
    B, d, Hq, Wq = h.shape
    z = h.flatten(2).transpose(1, 2).reshape(-1, d)
    if norm_latents:
        z = F.normalize(z, dim=-1)
    embed = sd["quantize._codebook.embed"]
    with torch.autocast("cpu", enabled=False):
        idx = cosine_scan(z, embed) if norm_codes else euclidean_scan(z, embed)
    quant = F.embedding(idx, embed).reshape(B, Hq * Wq, d).transpose(1, 2).reshape(B, d, Hq, Wq)
    return quant, idx.reshape(B, Hq, Wq), h
