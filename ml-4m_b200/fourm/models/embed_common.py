"""Shared plumbing of the B200 embedding modules: each module describes itself as one `b200fm_segment` so that
`FourM.forward` can select + embed all modalities with two kernel launches per side, and can also materialise its own
[B, L, D] outputs (the reference's `forward(d)` contract, used by the generation code)."""
import torch

from b200fm import functional as BF
from b200fm import lib, ops

KIND_IMG, KIND_TOK_IMG, KIND_SEQ, KIND_SEQ_EMB = lib.KIND_IMG, lib.KIND_TOK_IMG, lib.KIND_SEQ, lib.KIND_SEQ_EMB


def as_mask_u8(mask):
    """Reference masks are bool with True = masked; the kernels read them as bytes."""
    if mask.dtype != torch.bool:
        mask = mask != 0
    return mask.contiguous()


def materialise(module, d, mask_key, decoder_clamp):
    """Write d['x'] (token / patch embedding) and d['emb'] (pos + mod) for EVERY position of one modality with the
    plan + embed kernels in identity mode."""
    seg_static, main, mod = module.segment(d, mask_key, decoder_side=decoder_clamp)
    seg_static.setdefault("mod_id", 0)               # a lone modality: the id only feeds mod_mask, which is not returned here
    B = d['tensor'].shape[0]
    L = seg_static["L"]
    dev = mod.device
    if not mod.is_cuda:
        raise lib.B200FMError("b200fm embedding modules need CUDA tensors (there is no CPU fallback)")
    plan_seg = dict(seg_static)
    mode = ops.MODE_IDENTITY | ops.MODE_NO_SUM
    plan = ops.select_plan([plan_seg], mode, B, L, dev)
    x, emb = BF.EmbedRowsFn.apply(plan, [seg_static], module.dim_tokens, True, None, main, mod, module.pos_emb)
    return x, emb
