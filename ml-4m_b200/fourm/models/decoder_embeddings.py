"""Decoder-side modality embeddings of 4M behind the reference's class surface (fourm/models/decoder_embeddings.py of
apple/ml-4m).  `forward_embed` / `forward_logits` keep the reference contract; `segment` feeds the fused kernels."""
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from b200fm import functional as BF

from .embed_common import KIND_SEQ, KIND_TOK_IMG, as_mask_u8, materialise
from .fm_utils import build_1d_sincos_posemb, build_2d_sincos_posemb, pair


def _logits(module, x):
    """to_logits (reference decoder_embeddings.py:141-152 / 257-268): fp32 logits from the tcgen05 GEMM."""
    if type(module.to_logits) is nn.Linear and module.to_logits.bias is None:
        return BF.LinearF32Fn.apply(x, module.to_logits.weight)
    return module.to_logits(x)


class SequenceDecoderEmbedding(nn.Module):
    """Token-sequence targets (reference decoder_embeddings.py:24-152)."""

    def __init__(self, vocab_size: int, max_length: int, dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True,
                 max_sincos_pos_emb: int = 512, padding_idx: int = 0, share_embedding: bool = True, **kwargs):
        super().__init__()
        self.vocab_size, self.max_length, self.dim_tokens = vocab_size, max_length, dim_tokens
        self.sincos_pos_emb, self.padding_idx, self.max_sincos_pos_emb = sincos_pos_emb, padding_idx, max_sincos_pos_emb
        self.share_embedding = share_embedding
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768, init_std=0.02):
        self.dim_tokens = dim_tokens
        if self.sincos_pos_emb:
            if self.max_length > self.max_sincos_pos_emb:
                raise ValueError(f"Max length ({self.max_length}) is greater than the number of posembs ({self.max_sincos_pos_emb}")
            self.register_buffer("pos_emb", build_1d_sincos_posemb(max_len=self.max_sincos_pos_emb, embed_dim=dim_tokens)[:self.max_length])
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, self.max_length, dim_tokens))
            nn.init.normal_(self.pos_emb, std=init_std)
        self.mod_emb = nn.Parameter(torch.zeros(1, 1, dim_tokens))
        nn.init.normal_(self.mod_emb, std=init_std)
        self.token_emb = nn.Embedding(num_embeddings=self.vocab_size, embedding_dim=dim_tokens, padding_idx=self.padding_idx)
        self.to_logits = nn.Linear(dim_tokens, self.vocab_size, bias=False)
        if self.share_embedding:
            self.to_logits.weight = self.token_emb.weight          # tied, reference :89-91

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def segment(self, d, mask_key='target_mask', decoder_side=True):
        ids = d['tensor']
        st = dict(mask=as_mask_u8(d[mask_key]), ids=ids.contiguous(), L=ids.shape[1], kind=KIND_SEQ, pos_emb=self.pos_emb,
                  padding_idx=self.padding_idx if self.padding_idx is not None else -1, max_length=self.max_length)
        if 'decoder_attention_mask' in d:
            st["dam"] = d['decoder_attention_mask'].to(torch.int32).contiguous()
        return st, self.token_emb.weight, self.mod_emb

    def forward_embed(self, d: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        d['x'], d['emb'] = materialise(self, d, 'target_mask', True)
        d['ids'] = d['tensor']
        return d

    def forward_logits(self, x: torch.Tensor) -> torch.Tensor:
        return _logits(self, x)


class ImageTokenDecoderEmbedding(nn.Module):
    """Tokenised image-like targets (reference decoder_embeddings.py:156-268)."""

    def __init__(self, vocab_size: int, patch_size: Union[int, Tuple[int, int]] = 16, dim_tokens: Optional[int] = None,
                 sincos_pos_emb: bool = True, image_size: Union[int, Tuple[int]] = 224, share_embedding: bool = True, **kwargs):
        super().__init__()
        self.vocab_size, self.patch_size, self.dim_tokens = vocab_size, pair(patch_size), dim_tokens
        self.sincos_pos_emb, self.image_size = sincos_pos_emb, pair(image_size)
        self.num_patches = (self.image_size[0] // self.patch_size[0]) * (self.image_size[1] // self.patch_size[1])
        self.share_embedding = share_embedding
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768, init_std=0.02):
        self.dim_tokens = dim_tokens
        h_posemb, w_posemb = self.image_size[0] // self.patch_size[0], self.image_size[1] // self.patch_size[1]
        if self.sincos_pos_emb:
            self.register_buffer("pos_emb", build_2d_sincos_posemb(h=h_posemb, w=w_posemb, embed_dim=dim_tokens))
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, h_posemb * w_posemb, dim_tokens))
            nn.init.normal_(self.pos_emb, std=init_std)
        self.mod_emb = nn.Parameter(torch.zeros(1, 1, dim_tokens))
        nn.init.normal_(self.mod_emb, std=init_std)
        self.token_emb = nn.Embedding(num_embeddings=self.vocab_size, embedding_dim=dim_tokens)
        self.to_logits = nn.Linear(dim_tokens, self.vocab_size, bias=False)
        if self.share_embedding:
            self.to_logits.weight = self.token_emb.weight          # tied, reference :218-220

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def segment(self, d, mask_key='target_mask', decoder_side=True):
        ids = d['tensor']
        ids = ids.reshape(ids.shape[0], -1).contiguous()
        mask = d.get(mask_key)
        if mask is None:
            mask = torch.zeros(ids.shape, dtype=torch.bool, device=ids.device)
        st = dict(mask=as_mask_u8(mask), ids=ids, L=ids.shape[1], kind=KIND_TOK_IMG, pos_emb=self.pos_emb, padding_idx=-1, max_length=0)
        if 'decoder_attention_mask' in d:
            st["dam"] = d['decoder_attention_mask'].to(torch.int32).contiguous()
        return st, self.token_emb.weight, self.mod_emb

    def forward_embed(self, d: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        d['x'], d['emb'] = materialise(self, d, 'target_mask', True)
        d['ids'] = d['tensor'].reshape(d['tensor'].shape[0], -1)
        return d

    def forward_logits(self, x: torch.Tensor) -> torch.Tensor:
        return _logits(self, x)
