"""Synthetic mod-7 batches in the reference's mod_dict wire format (fourm/data/unified_datasets.py:488-520,
fourm/data/masking.py:236-266, 410-445): exactly `6*n_in_img + 2*n_in_seq` valid encoder tokens and
`5*n_tgt_img + 2*(n_tgt_seq - 1)` valid decoder tokens per sample (128 / 128 with the defaults)."""
import torch

MOD7 = {   # name: (kind, vocab)  -- fourm/data/modality_info.py:32-145
    'rgb@224': ('img', 0), 'tok_rgb@224': ('tok_img', 16384), 'tok_depth@224': ('tok_img', 8192), 'tok_normal@224': ('tok_img', 8192),
    'tok_semseg@224': ('tok_img', 4096), 'tok_clip@224': ('tok_img', 8192), 'caption': ('seq', 30000), 'det': ('seq', 30000),
}


def mod7_batch(B, n_in_img=18, n_in_seq=10, n_tgt_img=22, n_tgt_seq=10, seed=1234, image_size=224, patch=16, seq_width=514,
               pin_memory=False):
    g = torch.Generator().manual_seed(seed)
    P = (image_size // patch) ** 2
    side = image_size // patch
    out = {}
    for name, (kind, vocab) in MOD7.items():
        if kind in ('img', 'tok_img'):
            perm = torch.rand(B, P, generator=g).argsort(dim=1)
            imask = torch.ones(B, P, dtype=torch.bool)
            imask.scatter_(1, perm[:, :n_in_img], False)
            tmask = torch.ones(B, P, dtype=torch.bool)
            dam = torch.zeros(B, P, dtype=torch.int32)
            if kind == 'tok_img':
                tmask.scatter_(1, perm[:, n_in_img:n_in_img + n_tgt_img], False)
                first = (~tmask).int().argmax(dim=1)
                dam[torch.arange(B), first] = n_tgt_img
                t = torch.randint(0, vocab, (B, side, side), generator=g, dtype=torch.int64)
            else:
                t = torch.randn(B, 3, image_size, image_size, generator=g)
        else:
            t = torch.zeros(B, seq_width, dtype=torch.int32)
            imask = torch.ones(B, seq_width, dtype=torch.bool)
            tmask = torch.ones(B, seq_width, dtype=torch.bool)
            dam = torch.zeros(B, seq_width, dtype=torch.int32)
            t[:, :n_in_seq + n_tgt_seq] = torch.randint(200, vocab, (B, n_in_seq + n_tgt_seq), generator=g, dtype=torch.int32)
            imask[:, :n_in_seq] = False
            tmask[:, n_in_seq:n_in_seq + n_tgt_seq] = False
            dam[:, n_in_seq:n_in_seq + n_tgt_seq] = 1
        d = dict(tensor=t, input_mask=imask, target_mask=tmask, decoder_attention_mask=dam)
        if pin_memory:
            d = {k: v.pin_memory() for k, v in d.items()}
        out[name] = d
    return out


def budgets_for(n_tokens):
    """(n_in_img, n_in_seq, n_tgt_img, n_tgt_seq) giving exactly n_tokens encoder and n_tokens decoder tokens."""
    if n_tokens == 128:
        return 18, 10, 22, 10
    if n_tokens == 256:
        return 36, 20, 44, 19
    raise ValueError("synthetic budgets are defined for 128 or 256 tokens per side")


def batch_bytes(batch):
    return sum(v.numel() * v.element_size() for d in batch.values() for v in d.values())
