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


# ----------------------------------------------------------------------------------------------------------------------
# training side (a25): codebook EMA update, straight-through + commitment loss, ViT decoder, VQVAE.forward
# ----------------------------------------------------------------------------------------------------------------------
def cosine_codebook_train_step(x, embed, cluster_size, decay):
    """quantize_lucid.py:388-426 CosineSimCodebook.forward in training mode WITHOUT dead-code expiry (threshold 0) and without
    DDP (all_reduce_fn = noop).  x [..., d]; returns (quantize, idx, new_embed, new_cluster_size); quantize uses the OLD
    un-normalised buffer (F.embedding(embed_ind, self.embed), :407)."""
    shape = x.shape
    flatten = F.normalize(x.float().reshape(-1, shape[-1]), p=2, dim=-1)
    en = F.normalize(embed.float(), p=2, dim=-1)
    idx = (flatten @ en.t()).argmax(dim=-1)
    onehot = F.one_hot(idx, embed.shape[0]).float()
    quantize = F.embedding(idx.view(shape[:-1]), embed)
    bins = onehot.sum(0)
    new_cs = cluster_size * decay + bins * (1 - decay)
    zero = bins == 0
    bins = bins.masked_fill(zero, 1.0)
    embed_sum = flatten.t() @ onehot
    en_new = F.normalize((embed_sum / bins.unsqueeze(0)).t(), p=2, dim=-1)
    en_new = torch.where(zero[..., None], en, en_new)
    new_embed = embed * decay + en_new * (1 - decay)
    return quantize, idx.view(shape[:-1]), new_embed, new_cs


def euclid_codebook_train_step(x, embed, embed_avg, cluster_size, decay, eps=1e-5):
    """quantize_lucid.py:263-301 EuclideanCodebook.forward in training mode (no expiry, no DDP).
    Returns (quantize, idx, new_embed, new_embed_avg, new_cluster_size)."""
    shape = x.shape
    flatten = x.float().reshape(-1, shape[-1])
    e = embed.float().t()
    dist = -(flatten.pow(2).sum(1, keepdim=True) - 2 * flatten @ e + e.pow(2).sum(0, keepdim=True))
    idx = dist.argmax(dim=-1)
    onehot = F.one_hot(idx, embed.shape[0]).float()
    quantize = F.embedding(idx.view(shape[:-1]), embed)
    new_cs = cluster_size * decay + onehot.sum(0) * (1 - decay)
    embed_sum = flatten.t() @ onehot
    new_avg = embed_avg * decay + embed_sum.t() * (1 - decay)
    smoothed = (new_cs + eps) / (new_cs.sum() + embed.shape[0] * eps) * new_cs.sum()
    new_embed = new_avg / smoothed.unsqueeze(1)
    return quantize, idx.view(shape[:-1]), new_embed, new_avg, new_cs


def vector_quantize_train(fmap, quantize_fn, commitment_weight=1.0, norm_latents=False):
    """quantize_lucid.py:504-568 VectorQuantize.forward, training, heads = 1, Identity projections, image feature map in/out.
    quantize_fn(x [B, n, d]) -> (quantize, idx).  Returns (quantize_st [B, d, h, w], loss [1], idx [B, h, w])."""
    B, C, Hq, Wq = fmap.shape
    x = fmap.permute(0, 2, 3, 1).reshape(B, Hq * Wq, C)
    if norm_latents:
        x = F.normalize(x, p=2, dim=-1)
    q, idx = quantize_fn(x)
    q_st = x + (q - x).detach()                                                   # :532
    loss = torch.zeros(1) + F.mse_loss(q_st.detach(), x) * commitment_weight     # :537-539
    return q_st.reshape(B, Hq, Wq, C).permute(0, 3, 1, 2), loss, idx.reshape(B, Hq, Wq)


def vit_decoder(fmap, sd, pfx, cfg, patch: int, out_channels: int, post_mlp: bool):
    """vit_models.py:618-659 ViTDecoder.forward (patch_proj=True, no out_conv): [B, D, nh, nw] -> [B, C, H, W]."""
    B, D, nh, nw = fmap.shape
    x = fmap.flatten(2).transpose(1, 2)
    pe = F.interpolate(sd[pfx + "pos_emb"], size=(nh, nw), mode="bicubic", align_corners=False)
    x = x + pe.flatten(2).transpose(1, 2)
    for i in range(cfg["depth"]):
        x = vit_block(x, sd, f"{pfx}blocks.{i}.", cfg["heads"])
    if post_mlp:                                                                   # :645-646 (no autocast guard on this side)
        h = F.layer_norm(x, (D,), sd[pfx + "norm_mlp.weight"], sd[pfx + "norm_mlp.bias"], 1e-6)
        h = torch.tanh(F.linear(h, sd[pfx + "post_mlp.fc1.weight"], sd[pfx + "post_mlp.fc1.bias"]))
        x = x + F.linear(h, sd[pfx + "post_mlp.fc2.weight"], sd[pfx + "post_mlp.fc2.bias"])
    x = F.linear(x, sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"])
    x = x.reshape(B, nh, nw, out_channels, patch, patch).permute(0, 3, 1, 4, 2, 5)   # 'b (nh nw) (c ph pw) -> b c (nh ph) (nw pw)'
    return x.reshape(B, out_channels, nh * patch, nw * patch)


def vqvae_forward_train(img, sd, kw):
    """vqvae.py:454-471 VQVAE.forward in training mode (cosine codebook, no expiry): returns (dec, code_loss, idx, new_embed,
    new_cluster_size).  kw: enc_type / dec_type presets, patch_size, post_mlp, ema_decay, commitment_weight, norm_latents."""
    enc_cfg, dec_cfg = VIT_PRESETS[kw["enc_type"]], VIT_PRESETS[kw["dec_type"].replace("_dec", "_enc")]
    h = vit_encoder(img, sd, "encoder.", enc_cfg, kw["patch_size"], kw["post_mlp"])
    h = F.conv2d(h, sd["quant_proj.weight"], sd["quant_proj.bias"])
    state = {}

    def qfn(x):
        q, idx, ne, ncs = cosine_codebook_train_step(x, sd["quantize._codebook.embed"], sd["quantize._codebook.cluster_size"], kw["ema_decay"])
        state["embed"], state["cluster_size"] = ne, ncs
        return q, idx

    quant, loss, idx = vector_quantize_train(h, qfn, kw.get("commitment_weight", 1.0), kw.get("norm_latents", False))
    d = F.conv2d(quant, sd["post_quant_proj.weight"], sd["post_quant_proj.bias"])
    dec = vit_decoder(d, sd, "decoder.", dec_cfg, kw["patch_size"], img.shape[1], kw["post_mlp"])
    return dec, loss, idx, state["embed"], state["cluster_size"]
