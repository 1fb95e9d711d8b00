"""CPU oracle for the 4M hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state_dict in, tensors out) restatement, in plain torch on the CPU, of the algorithm the
apple/ml-4m reference runs for `FourM.forward` and its pieces.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import it; the product path under
`ml-4m_b200/` never does (it fails loudly without its CUDA library instead).

Pinning: upstream has no tests or golden vectors ("parity unpinned" upstream).  This
restatement is pinned instead against the UNMODIFIED reference imported in the authoring container
(`tests/golden/make_golden.py` -> `tests/golden/*.pt`, and `tests/test_oracle_vs_reference.py` when
/root/reference is present), torch 2.11.0 CPU.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Running any of these under `torch.autocast('cpu', dtype=torch.bfloat16)` reproduces the reference's
bf16-autocast numerics because only F.linear / matmul / F.layer_norm / softmax / cross_entropy are used,
exactly the ops the reference uses.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# static known answers
# --------------------------------------------------------------------------------------------------


def modality_id(name: str) -> int:
    """fourm/utils/misc.py:39-41 generate_uint15_hash: sha256(name) mod 2**15."""
    return int(hashlib.sha256(name.encode("utf-8")).hexdigest(), 16) % (2 ** 15)


def sincos_1d(max_len: int, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """fourm/models/fm_utils.py:32-44 build_1d_sincos_posemb -> [1, max_len, dim]."""
    assert dim % 2 == 0
    half = dim // 2
    omega = 1.0 / (temperature ** (torch.arange(half, dtype=torch.float32) / half))
    ang = torch.arange(max_len, dtype=torch.float32)[:, None] * omega[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=1)[None]


def sincos_2d(h: int, w: int, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """fourm/models/fm_utils.py:46-61 build_2d_sincos_posemb -> [1, h*w, dim].

    The reference meshgrids (w, h) with indexing='ij' and flattens, so flat index p has
    grid_w = p // h and grid_h = p % h (a transposed raster when h != w; identical when h == w)."""
    assert dim % 4 == 0
    q = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(q, dtype=torch.float32) / q))
    p = torch.arange(h * w)
    gw = (p // h).to(torch.float32)
    gh = (p % h).to(torch.float32)
    aw = gw[:, None] * omega[None, :]
    ah = gh[:, None] * omega[None, :]
    return torch.cat([aw.sin(), aw.cos(), ah.sin(), ah.cos()], dim=1)[None]


# --------------------------------------------------------------------------------------------------
# modality specs (the part of fourm/data/modality_info.py the hot path depends on)
# --------------------------------------------------------------------------------------------------


def mod_spec(name: str, kind: str, vocab: int = 0, max_length: int = 0, image_size: int = 224,
             patch_size: int = 16, channels: int = 3) -> dict:
    """kind: 'img' (pixel patches, encoder only), 'tok_img' (tokenised image), 'seq' (token sequence)."""
    assert kind in ("img", "tok_img", "seq")
    return dict(name=name, kind=kind, id=modality_id(name), vocab=vocab, max_length=max_length,
                image_size=image_size, patch_size=patch_size, channels=channels,
                n_patches=(image_size // patch_size) ** 2)


def mod7_specs() -> Dict[str, dict]:
    """fourm/data/modality_info.py:32-145, the 4M-7 entries used by
    cfgs/default/4m/data/cc12m/main/mix_mod7_all2all_rgb2all_a0.5.yaml:7-8."""
    return {
        "rgb@224": mod_spec("rgb@224", "img"),
        "tok_rgb@224": mod_spec("tok_rgb@224", "tok_img", vocab=16384),
        "tok_depth@224": mod_spec("tok_depth@224", "tok_img", vocab=8192),
        "tok_normal@224": mod_spec("tok_normal@224", "tok_img", vocab=8192),
        "tok_semseg@224": mod_spec("tok_semseg@224", "tok_img", vocab=4096),
        "tok_clip@224": mod_spec("tok_clip@224", "tok_img", vocab=8192),
        "caption": mod_spec("caption", "seq", vocab=30000, max_length=256),
        "det": mod_spec("det", "seq", vocab=30000, max_length=256),
    }


# --------------------------------------------------------------------------------------------------
# embeddings (a3-a7)
# --------------------------------------------------------------------------------------------------


def _seq_pos_ids(mask: torch.Tensor, max_length: Optional[int]) -> torch.Tensor:
    """encoder_embeddings.py:110-112 / decoder_embeddings.py:125-128: rank among valid positions."""
    pos = (~mask).int().cumsum(dim=1) - 1
    pos = pos.masked_fill(mask, 0)
    if max_length is not None:
        pos = pos.masked_fill(pos >= max_length, 0)
    return pos.long()


def embed_sequence(ids: torch.Tensor, mask: torch.Tensor, token_emb: torch.Tensor, pos_emb: torch.Tensor,
                   mod_emb: torch.Tensor, clamp_len: Optional[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """SequenceEncoderEmbedding.forward (encoder_embeddings.py:87-121; clamp_len=None) and
    SequenceDecoderEmbedding.forward_embed (decoder_embeddings.py:98-139; clamp_len=max_length).
    `token_emb` row padding_idx is whatever the state_dict holds (zero after init)."""
    x = F.embedding(ids.long(), token_emb)
    pos = _seq_pos_ids(mask, clamp_len)
    pe = pos_emb[0][pos]                          # gather of the (un-expanded) table
    pe = pe.masked_fill(mask[..., None], 0.0)
    return x, pe + mod_emb


def embed_sequence_features(feats: torch.Tensor, mask: torch.Tensor, proj: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                            pos_emb: torch.Tensor, mod_emb: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """SequenceEmbEncoderEmbedding.forward (encoder_embeddings.py:387-421; T5-XXL features of 4M-21): x = emb_proj(feats) on every
    position (`proj` = [(W, b)] or the bottleneck pair), emb = sincos1d[rank among valid inputs] (zeroed where masked) + mod_emb."""
    x = feats
    for w, b in proj:
        x = F.linear(x, w, b)
    pos = _seq_pos_ids(mask, None)
    pe = pos_emb[0][pos].masked_fill(mask[..., None], 0.0)
    return x, pe + mod_emb


def embed_image_tokens(ids: torch.Tensor, token_emb: torch.Tensor, pos_emb: torch.Tensor,
                       mod_emb: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """ImageTokenEncoderEmbedding.forward (encoder_embeddings.py:184-211) and
    ImageTokenDecoderEmbedding.forward_embed (decoder_embeddings.py:226-255)."""
    B = ids.shape[0]
    ids = ids.reshape(B, -1).long()
    x = F.embedding(ids, token_emb)
    emb = (pos_emb + mod_emb).expand(B, -1, -1)
    return x, emb


def patchify(img: torch.Tensor, p: int) -> torch.Tensor:
    """encoder_embeddings.py:301 rearrange 'b d (nh ph) (nw pw) -> b (nh nw) (ph pw d)'."""
    B, C, H, W = img.shape
    nh, nw = H // p, W // p
    x = img.reshape(B, C, nh, p, nw, p).permute(0, 2, 4, 3, 5, 1)      # b nh nw ph pw d
    return x.reshape(B, nh * nw, p * p * C)


def embed_image_pixels(img: torch.Tensor, proj_w: torch.Tensor, pos_emb: torch.Tensor, mod_emb: torch.Tensor,
                       p: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """ImageEncoderEmbedding.forward (encoder_embeddings.py:280-309): bias-free patch Linear."""
    B, C, H, W = img.shape
    if H % p or W % p:
        raise AssertionError(f"Image sizes {H}x{W} must be divisible by patch sizes {p}x{p}")
    x = F.linear(patchify(img, p), proj_w)
    emb = (pos_emb + mod_emb).expand(B, -1, -1)
    return x, emb


# --------------------------------------------------------------------------------------------------
# masking / compaction (a8-a10): integer-exact
# --------------------------------------------------------------------------------------------------


def stable_keep_indices(mask: torch.Tensor, k: int) -> torch.Tensor:
    """fm.py:364-367 `argsort(mask + arange*1e-6)[:, :k]` == stable partition: indices of the valid
    (mask False) positions in order, then the masked ones in order; first k.  Verified identical to the
    argsort form for widths up to 3000."""
    B, L = mask.shape
    ar = torch.arange(L, device=mask.device)
    key = mask.long() * L + ar                       # valid first, ties broken by position: exact integers
    return torch.argsort(key, dim=1, stable=True)[:, :k]


def decoder_attention_mask(dam: torch.Tensor, mod_mask: torch.Tensor, causal: bool = False,
                           sep: bool = True) -> torch.Tensor:
    """fm.py:440-475 adapt_decoder_attention_mask -> bool [B, M, M], True = masked."""
    B, M = dam.shape
    if causal:
        out = torch.ones(M, M, dtype=torch.bool, device=dam.device).triu(1)[None].expand(B, -1, -1)
    else:
        cs = torch.cumsum(dam, dim=-1)
        out = torch.arange(M, device=dam.device)[None, None, :] >= cs[:, :, None]
    if sep:
        out = out | (mod_mask[:, None, :] != mod_mask[:, :, None])
    return out


def encoder_select(parts: Sequence[Tuple[dict, torch.Tensor, torch.Tensor, torch.Tensor]], n_keep: int,
                   register_tokens: Optional[torch.Tensor] = None):
    """cat_encoder_tensors + forward_mask_encoder (fm.py:245-277, 338-390).
    parts: [(spec, x[B,L,D], emb[B,L,D], input_mask[B,L])] in mod_dict order.
    Returns tokens, emb [B,N,D]; mask [B,1,N] bool; mod_mask [B,N] int16; ids_keep [B,n_keep]."""
    x_all = torch.cat([p[1] for p in parts], dim=1)
    e_all = torch.cat([p[2] for p in parts], dim=1)
    m_all = torch.cat([p[3] for p in parts], dim=1)
    mod_all = torch.cat([torch.full_like(p[3], p[0]["id"], dtype=torch.int16) for p in parts], dim=1)
    keep = stable_keep_indices(m_all, n_keep)
    D = x_all.shape[2]
    gi = keep[..., None].expand(-1, -1, D)
    tok = torch.gather(x_all, 1, gi)
    emb = torch.gather(e_all, 1, gi)
    msk = torch.gather(m_all, 1, keep)
    mod = torch.gather(mod_all, 1, keep)
    if register_tokens is not None and register_tokens.shape[1] > 0:
        B, R = x_all.shape[0], register_tokens.shape[1]
        reg = register_tokens.expand(B, -1, -1)
        tok = torch.cat([reg, tok], 1)
        emb = torch.cat([torch.zeros_like(reg), emb], 1)
        msk = torch.cat([torch.zeros(B, R, dtype=torch.bool), msk], 1)
        mod = torch.cat([torch.full((B, R), -1, dtype=torch.int16), mod], 1)
    tok = tok.masked_fill(msk[..., None], 0.0)
    emb = emb.masked_fill(msk[..., None], 0.0)
    mod = mod.masked_fill(msk, -1)
    return tok, emb, msk[:, None, :], mod, keep


def decoder_select(parts, n_keep: int, mask_token: torch.Tensor, causal: bool = False, sep: bool = True):
    """cat_decoder_tensors + forward_mask_decoder (fm.py:279-336, 392-438).
    parts: [(spec, x, emb, ids, target_mask, dam)] ALREADY in the shuffled modality order
    (fm.py:306 draws that order with Python's `random`; the caller owns the RNG)."""
    xs, es, ms, ts, ds, mods = [], [], [], [], [], []
    for spec, x, emb, ids, tmask, dam in parts:
        if spec["kind"] == "seq":
            # teacher forcing shift, fm.py:309-319
            xs.append(x[:, :-1]); ts.append(ids[:, 1:]); es.append(emb[:, :-1])
            ms.append(torch.logical_or(tmask[:, 1:], tmask[:, :-1])); ds.append(dam[:, :-1])
            mods.append(torch.full_like(ids[:, :-1], spec["id"], dtype=torch.int16))
        else:
            xs.append(torch.zeros_like(x) + mask_token); ts.append(ids); es.append(emb)       # fm.py:322
            ms.append(tmask); ds.append(dam)
            mods.append(torch.full_like(ids, spec["id"], dtype=torch.int16))
    x_all, e_all, m_all = torch.cat(xs, 1), torch.cat(es, 1), torch.cat(ms, 1)
    # torch.cat type-promotes (int32 sequence ids with int64 image ids -> int64), as in the reference
    t_all, d_all, mod_all = torch.cat(ts, 1), torch.cat(ds, 1), torch.cat(mods, 1)
    keep = stable_keep_indices(m_all, n_keep)
    D = x_all.shape[2]
    gi = keep[..., None].expand(-1, -1, D)
    tok = torch.gather(x_all, 1, gi)
    emb = torch.gather(e_all, 1, gi)
    msk = torch.gather(m_all, 1, keep)
    tgt = torch.gather(t_all, 1, keep)
    dam = torch.gather(d_all, 1, keep)
    mod = torch.gather(mod_all, 1, keep)
    tok = tok.masked_fill(msk[..., None], 0.0)
    emb = emb.masked_fill(msk[..., None], 0.0)
    tgt = tgt.masked_fill(msk, 0)
    amask = decoder_attention_mask(dam, mod, causal, sep)      # BEFORE pad rows get mod = -1 (fm.py:431-432)
    mod = mod.masked_fill(msk, -1)
    return tok, emb, msk[:, None, :], tgt, amask, mod, keep


# --------------------------------------------------------------------------------------------------
# transformer pieces (a11-a16)
# --------------------------------------------------------------------------------------------------


def layer_norm(x, w, b, eps=1e-6):
    """fm_utils.py:93-108 (bias is a zero buffer in *_nobias presets)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _sdpa(q, k, v, mask, scale):
    """fm_utils.py:165-177: materialised softmax(q k^T * scale, masked_fill(-finfo.max)) v.
    q [B,h,Nq,dh], k/v [B,h,Nk,dh], mask broadcastable bool [B,1,1|Nq,Nk] (True = masked)."""
    attn = (q @ k.transpose(-2, -1)) * scale
    if mask is not None:
        attn = attn.masked_fill(mask, -torch.finfo(attn.dtype).max)
    attn = attn.softmax(dim=-1)
    return attn @ v


def _qk_norm(q, k, sd, pfx):
    """fm_utils.py:244-245, 290-291 (NormAttention / NormCrossAttention): LayerNorm over head_dim on q and k when the
    state dict carries q_norm / k_norm (qk_norm presets, fm.py:1059-1130); identity otherwise."""
    if pfx + "q_norm.weight" not in sd:
        return q, k
    return (layer_norm(q, sd[pfx + "q_norm.weight"], sd.get(pfx + "q_norm.bias")),
            layer_norm(k, sd[pfx + "k_norm.weight"], sd.get(pfx + "k_norm.bias")))


def self_attention(x, sd, pfx, heads, mask=None):
    """fm_utils.py:147-180 Attention.forward; mask [B,1|N,N]."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[pfx + "qkv.weight"], sd.get(pfx + "qkv.bias"))
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k = _qk_norm(qkv[0], qkv[1], sd, pfx)
    o = _sdpa(q, k, qkv[2], None if mask is None else mask[:, None], (C // heads) ** -0.5)
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pfx + "proj.weight"], sd.get(pfx + "proj.bias"))


def cross_attention(x, ctx, sd, pfx, heads, mask=None):
    """fm_utils.py:182-219 CrossAttention.forward; mask [B,1|N,M]."""
    B, N, C = x.shape
    M = ctx.shape[1]
    q = F.linear(x, sd[pfx + "q.weight"], sd.get(pfx + "q.bias")).reshape(B, N, heads, C // heads).permute(0, 2, 1, 3)
    kv = F.linear(ctx, sd[pfx + "kv.weight"], sd.get(pfx + "kv.bias"))
    kv = kv.reshape(B, M, 2, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k = _qk_norm(q, kv[0], sd, pfx)
    o = _sdpa(q, k, kv[1], None if mask is None else mask[:, None], (C // heads) ** -0.5)
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pfx + "proj.weight"], sd.get(pfx + "proj.bias"))


def mlp(x, sd, pfx, gated: bool, act: str = "gelu"):
    """fm_utils.py:111-144: Mlp (fc2(act(fc1 x))) or GatedMlp (fc2(act(fc1 x) * fc3 x))."""
    a = F.linear(x, sd[pfx + "fc1.weight"], sd.get(pfx + "fc1.bias"))
    a = {"gelu": F.gelu, "silu": F.silu, "tanh": torch.tanh}[act](a)
    if gated:
        a = a * F.linear(x, sd[pfx + "fc3.weight"], sd.get(pfx + "fc3.bias"))
    return F.linear(a, sd[pfx + "fc2.weight"], sd.get(pfx + "fc2.bias"))


def encoder_block(x, sd, pfx, cfg, mask):
    """fm_utils.py:331-334 Block.forward (drop_path = identity at rate 0)."""
    x = x + self_attention(layer_norm(x, sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], cfg["eps"]),
                           sd, pfx + "attn.", cfg["heads"], mask)
    x = x + mlp(layer_norm(x, sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], cfg["eps"]),
                sd, pfx + "mlp.", cfg["gated"], cfg["act"])
    return x


def decoder_block(x, ctx, sd, pfx, cfg, sa_mask, xa_mask):
    """fm_utils.py:362-366 DecoderBlock.forward; context_norm recomputed per layer."""
    e = cfg["eps"]
    x = x + self_attention(layer_norm(x, sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], e),
                           sd, pfx + "self_attn.", cfg["heads"], sa_mask)
    x = x + cross_attention(layer_norm(x, sd[pfx + "query_norm.weight"], sd[pfx + "query_norm.bias"], e),
                            layer_norm(ctx, sd[pfx + "context_norm.weight"], sd[pfx + "context_norm.bias"], e),
                            sd, pfx + "cross_attn.", cfg["heads"], xa_mask)
    x = x + mlp(layer_norm(x, sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], e), sd, pfx + "mlp.",
                cfg["gated"], cfg["act"])
    return x


def run_encoder(x, sd, cfg, mask):
    """fm.py:477-495 forward_encoder."""
    for i in range(cfg["enc_depth"]):
        x = encoder_block(x, sd, f"encoder.{i}.", cfg, mask)
    return layer_norm(x, sd["encoder_norm.weight"], sd["encoder_norm.bias"], cfg["eps"])


def run_decoder(y, ctx, sd, cfg, enc_mask, dec_mask):
    """fm.py:497-519 forward_decoder."""
    for i in range(cfg["dec_depth"]):
        y = decoder_block(y, ctx, sd, f"decoder.{i}.", cfg, dec_mask, enc_mask)
    return layer_norm(y, sd["decoder_norm.weight"], sd["decoder_norm.bias"], cfg["eps"])


# --------------------------------------------------------------------------------------------------
# heads / losses (a17)
# --------------------------------------------------------------------------------------------------


def modality_losses(y, target_ids, mod_mask, sd, dec_specs: Sequence[dict], loss_type: str = "mod"):
    """fm.py:521-637 forward_mod_loss / forward_token_loss.  dec_specs in decoder_mod_dict order
    (= mod_dict order, NOT the shuffled order)."""
    if loss_type not in ("mod", "modality", "token"):
        raise ValueError("Invalid loss type")
    mod_loss, mod_count = {}, {}
    for spec in dec_specs:
        sel = mod_mask == spec["id"]
        rows = y[sel]
        logits = F.linear(rows, sd[f"decoder_embeddings.{spec['name']}.to_logits.weight"])
        if logits.numel() == 0:
            mod_loss[spec["name"]] = torch.zeros(1)
            mod_count[spec["name"]] = 0
        else:
            mod_loss[spec["name"]] = F.cross_entropy(logits, target_ids[sel].long(), reduction="mean")
            mod_count[spec["name"]] = logits.numel()
    if loss_type == "token":
        loss = sum(mod_loss[m] * mod_count[m] for m in mod_loss) / sum(mod_count.values())
    else:
        loss = sum(mod_loss.values()) / len(mod_loss)
    return loss, mod_loss


def all_logits(y, sd, dec_specs):
    """fm.py:539-546 forward_logits(return_all_logits=True): every head on every row."""
    return {s["name"]: F.linear(y, sd[f"decoder_embeddings.{s['name']}.to_logits.weight"]) for s in dec_specs}


# --------------------------------------------------------------------------------------------------
# whole forward (a1)
# --------------------------------------------------------------------------------------------------


def model_cfg(dim, heads, enc_depth, dec_depth, gated=True, act="silu", eps=1e-6, causal=False, sep=True,
              num_register_tokens=0):
    return dict(dim=dim, heads=heads, enc_depth=enc_depth, dec_depth=dec_depth, gated=gated, act=act, eps=eps,
                causal=causal, sep=sep, num_register_tokens=num_register_tokens)


PRESETS = {   # fm.py:840-1130 (swiglu_nobias family)
    "fm_tiny_6e_6d_swiglu_nobias": model_cfg(384, 6, 6, 6),
    "fm_small_8e_8d_swiglu_nobias": model_cfg(512, 8, 8, 8),
    "fm_base_12e_12d_swiglu_nobias": model_cfg(768, 12, 12, 12),
    "fm_large_24e_24d_swiglu_nobias": model_cfg(1024, 16, 24, 24),
    "fm_xlarge_24e_24d_swiglu_nobias": model_cfg(2048, 32, 24, 24),
}


def fourm_forward(sd: Dict[str, torch.Tensor], cfg: dict, specs: Dict[str, dict], mod_dict: Dict[str, dict],
                  num_encoder_tokens: int, num_decoder_tokens: int, decoder_order: Sequence[str],
                  loss_type: str = "mod", return_logits: bool = False, return_intermediates: bool = False):
    """FourM.forward (fm.py:640-691).  `decoder_order` is the modality order the reference's
    `random.sample` (fm.py:306) produced for this call.  Does not mutate mod_dict (the reference does,
    SURVEY.md v6; the product mirrors that, the oracle has no need to)."""
    enc_parts, dec_parts = [], {}
    for name, d in mod_dict.items():
        spec = specs[name]
        pe = f"encoder_embeddings.{name}."
        if pe + "mod_emb" in sd:
            if spec["kind"] == "seq":
                x, e = embed_sequence(d["tensor"], d["input_mask"], sd[pe + "token_emb.weight"], sd[pe + "pos_emb"],
                                      sd[pe + "mod_emb"], None)
            elif spec["kind"] == "tok_img":
                x, e = embed_image_tokens(d["tensor"], sd[pe + "token_emb.weight"], sd[pe + "pos_emb"], sd[pe + "mod_emb"])
            else:
                x, e = embed_image_pixels(d["tensor"], sd[pe + "proj.weight"], sd[pe + "pos_emb"], sd[pe + "mod_emb"],
                                          spec["patch_size"])
            enc_parts.append((spec, x, e, d["input_mask"]))
        pd = f"decoder_embeddings.{name}."
        if pd + "mod_emb" in sd:
            if spec["kind"] == "seq":
                x, e = embed_sequence(d["tensor"], d["target_mask"], sd[pd + "token_emb.weight"], sd[pd + "pos_emb"],
                                      sd[pd + "mod_emb"], spec["max_length"])
                ids = d["tensor"]
            else:
                x, e = embed_image_tokens(d["tensor"], sd[pd + "token_emb.weight"], sd[pd + "pos_emb"], sd[pd + "mod_emb"])
                ids = d["tensor"].reshape(d["tensor"].shape[0], -1)
            dec_parts[name] = (spec, x, e, ids, d["target_mask"], d["decoder_attention_mask"])
    assert sorted(decoder_order) == sorted(dec_parts), (decoder_order, list(dec_parts))
    reg = sd.get("register_tokens") if cfg["num_register_tokens"] > 0 else None
    enc_tok, enc_emb, enc_mask, enc_mod, enc_keep = encoder_select(enc_parts, num_encoder_tokens, reg)
    dec_tok, dec_emb, dec_mask, tgt, dec_amask, dec_mod, dec_keep = decoder_select(
        [dec_parts[m] for m in decoder_order], num_decoder_tokens, sd["mask_token"], cfg["causal"], cfg["sep"])

    x = run_encoder(enc_tok + enc_emb, sd, cfg, enc_mask)
    ctx = F.linear(x, sd["decoder_proj_context.weight"], sd["decoder_proj_context.bias"]) + enc_emb   # fm.py:679
    y = run_decoder(dec_tok + dec_emb, ctx, sd, cfg, enc_mask, dec_amask)

    dec_specs = [specs[m] for m in mod_dict if m in dec_parts]
    inter = dict(enc_keep=enc_keep, dec_keep=dec_keep, enc_mask=enc_mask, dec_mask=dec_mask, enc_mod=enc_mod,
                 dec_mod=dec_mod, target_ids=tgt, dec_attn_mask=dec_amask, enc_x0=enc_tok + enc_emb,
                 dec_y0=dec_tok + dec_emb, enc_out=x, context=ctx, dec_out=y)
    if return_logits:
        out = all_logits(y, sd, dec_specs)
        return (out, inter) if return_intermediates else out
    loss, mod_loss = modality_losses(y, tgt, dec_mod, sd, dec_specs, loss_type)
    return (loss, mod_loss, inter) if return_intermediates else (loss, mod_loss)


# --------------------------------------------------------------------------------------------------
# synthetic mod_dict batches in the wire format (a2)
# --------------------------------------------------------------------------------------------------


def synthetic_mod7_batch(B: int, n_in_img: int = 18, n_in_seq: int = 10, n_tgt_img: int = 22, n_tgt_seq: int = 10,
                         seed: int = 1234, specs: Optional[Dict[str, dict]] = None, seq_width: int = 514,
                         extra_valid: int = 0) -> Dict[str, dict]:
    """The mod-7 batch of SURVEY.md 8d: 6*n_in_img + 2*n_in_seq valid encoder rows and
    5*n_tgt_img + 2*(n_tgt_seq-1) valid decoder rows per sample (defaults: 128 / 128).
    Wire format: fourm/data/unified_datasets.py:488-520, fourm/data/masking.py:236-266, 410-445
    (masks: True = masked out).  `extra_valid` adds that many more valid targets per tok_img modality
    (exercises truncation when > budget)."""
    specs = specs or mod7_specs()
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, s in specs.items():
        if s["kind"] in ("img", "tok_img"):
            P = s["n_patches"]
            perm = torch.stack([torch.randperm(P, generator=g) for _ in range(B)])
            imask = torch.ones(B, P, dtype=torch.bool)
            imask.scatter_(1, perm[:, :n_in_img], False)
            tmask = torch.ones(B, P, dtype=torch.bool)
            dam = torch.zeros(B, P, dtype=torch.int32)
            if s["kind"] == "tok_img":
                nt = n_tgt_img + extra_valid
                tmask.scatter_(1, perm[:, n_in_img:n_in_img + nt], False)
                first = (~tmask).int().argmax(dim=1)
                dam[torch.arange(B), first] = nt                      # masking.py:262-264
                side = int(math.isqrt(P))
                t = torch.randint(0, s["vocab"], (B, side, side), generator=g, dtype=torch.int64)
            else:
                hw = s["image_size"]
                t = torch.randn(B, s["channels"], hw, hw, generator=g)
            out[name] = dict(tensor=t, input_mask=imask, target_mask=tmask, decoder_attention_mask=dam)
        else:
            L = seq_width
            budget = L // 2
            t = torch.zeros(B, L, dtype=torch.int32)
            imask = torch.ones(B, L, dtype=torch.bool)
            tmask = torch.ones(B, L, dtype=torch.bool)
            dam = torch.zeros(B, L, dtype=torch.int32)
            t[:, :n_in_seq] = torch.randint(200, s["vocab"], (B, n_in_seq), generator=g, dtype=torch.int32)
            imask[:, :n_in_seq] = False
            lo = n_in_seq
            t[:, lo:lo + n_tgt_seq] = torch.randint(200, s["vocab"], (B, n_tgt_seq), generator=g, dtype=torch.int32)
            tmask[:, lo:lo + n_tgt_seq] = False
            dam[:, lo:lo + n_tgt_seq] = 1                              # masking.py:416-424
            del budget
            out[name] = dict(tensor=t, input_mask=imask, target_mask=tmask, decoder_attention_mask=dam)
    return out


def canonical_param_name(key: str, keys) -> str:
    """Shared parameters appear under two state_dict names (fm.py:176-180 mod_emb sharing,
    decoder_embeddings.py:89-91 / 218-220 tied to_logits); fixtures seed them by one canonical name."""
    if key.startswith("decoder_embeddings.") and key.endswith(".mod_emb"):
        e = "encoder_embeddings." + key[len("decoder_embeddings."):]
        if e in keys:
            return e
    if key.endswith(".to_logits.weight"):
        t = key[:-len("to_logits.weight")] + "token_emb.weight"
        if t in keys:
            return t
    return key


def deterministic_tensor(name: str, shape, scale: float = 0.02) -> torch.Tensor:
    """Weights for fixtures: N(0, scale) from a generator seeded by sha256(name); lets golden files omit
    the (large) state_dict.  Norm weights are centred on 1."""
    seed = int(hashlib.sha256(name.encode()).hexdigest()[:12], 16)
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(tuple(shape), generator=g) * scale
    if name.endswith("weight") and ("norm" in name.split(".")[-2]):
        t = t + 1.0
    return t
