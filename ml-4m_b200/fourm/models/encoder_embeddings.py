"""Encoder-side modality embeddings of 4M behind the reference's class surface (fourm/models/encoder_embeddings.py of
apple/ml-4m; constructor signatures, parameter / buffer names and `init` are the reference's).

Each module exposes `segment(d, ...)`: its description as one gather segment (ids, mask, tables) for the fused
select + embed kernels that `FourM.forward` runs over all modalities at once; `forward(d)` (reference contract: adds the
full-length 'x' and 'emb' to the dict) materialises the same values through those kernels."""
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from b200fm import functional as BF
from b200fm import ops

from .embed_common import KIND_IMG, KIND_SEQ, KIND_SEQ_EMB, KIND_TOK_IMG, as_mask_u8, materialise
from .fm_utils import build_1d_sincos_posemb, build_2d_sincos_posemb, pair


class SequenceEncoderEmbedding(nn.Module):
    """Token-sequence inputs such as captions or detection strings (reference encoder_embeddings.py:22-121)."""

    def __init__(self, vocab_size: int, max_length: int, dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True,
                 max_sincos_pos_emb: int = 512, padding_idx: int = 0):
        super().__init__()
        self.vocab_size, self.max_length, self.dim_tokens = vocab_size, max_length, dim_tokens
        self.sincos_pos_emb, self.padding_idx, self.max_sincos_pos_emb = sincos_pos_emb, padding_idx, max_sincos_pos_emb
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768, init_std=0.02):
        self.dim_tokens = dim_tokens
        if self.sincos_pos_emb:
            if self.max_length > self.max_sincos_pos_emb:
                raise ValueError(f"Max length ({self.max_length}) is greater than the number of posembs ({self.max_sincos_pos_emb}")
            # reference quirk kept for state_dict compatibility: the slice hits the size-1 batch dim -> [1, 512, D]
            self.register_buffer("pos_emb", build_1d_sincos_posemb(max_len=self.max_sincos_pos_emb, embed_dim=dim_tokens)[:self.max_length])
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, self.max_length, dim_tokens))
            nn.init.normal_(self.pos_emb, std=init_std)
        self.mod_emb = nn.Parameter(torch.zeros(1, 1, dim_tokens))
        nn.init.normal_(self.mod_emb, std=init_std)
        self.token_emb = nn.Embedding(num_embeddings=self.vocab_size, embedding_dim=dim_tokens, padding_idx=self.padding_idx)

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def segment(self, d, mask_key='input_mask', decoder_side=False):
        ids = d['tensor']
        st = dict(mask=as_mask_u8(d[mask_key]), ids=ids.contiguous(), L=ids.shape[1], kind=KIND_SEQ, pos_emb=self.pos_emb,
                  padding_idx=self.padding_idx if self.padding_idx is not None else -1, max_length=0)
        return st, self.token_emb.weight, self.mod_emb

    def forward(self, d: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        d['x'], d['emb'] = materialise(self, d, 'input_mask', False)
        return d


class ImageTokenEncoderEmbedding(nn.Module):
    """Tokenised image-like inputs (reference encoder_embeddings.py:123-211)."""

    def __init__(self, vocab_size: int, patch_size: Union[int, Tuple[int, int]] = 16, dim_tokens: Optional[int] = None,
                 sincos_pos_emb: bool = True, image_size: Union[int, Tuple[int]] = 224, **kwargs):
        super().__init__()
        self.vocab_size, self.patch_size, self.dim_tokens = vocab_size, pair(patch_size), dim_tokens
        self.sincos_pos_emb, self.image_size = sincos_pos_emb, pair(image_size)
        self.num_patches = (self.image_size[0] // patch_size) * (self.image_size[1] // patch_size)
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

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def segment(self, d, mask_key='input_mask', decoder_side=False):
        ids = d['tensor']
        ids = ids.reshape(ids.shape[0], -1).contiguous()
        mask = d.get(mask_key)
        if mask is None:
            mask = torch.zeros(ids.shape, dtype=torch.bool, device=ids.device)
        st = dict(mask=as_mask_u8(mask), ids=ids, L=ids.shape[1], kind=KIND_TOK_IMG, pos_emb=self.pos_emb, padding_idx=-1, max_length=0)
        return st, self.token_emb.weight, self.mod_emb

    def forward(self, d: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        d['x'], d['emb'] = materialise(self, d, 'input_mask', False)
        return d


class ImageEncoderEmbedding(nn.Module):
    """Pixel / feature-map inputs, patchified and linearly projected without bias (reference encoder_embeddings.py:214-309)."""

    def __init__(self, num_channels: int, patch_size: Union[int, Tuple[int, int]], dim_tokens: Optional[int] = None,
                 sincos_pos_emb: bool = True, image_size: Union[int, Tuple[int]] = 224):
        super().__init__()
        self.num_channels, self.patch_size, self.dim_tokens = num_channels, pair(patch_size), dim_tokens
        self.sincos_pos_emb, self.image_size = sincos_pos_emb, pair(image_size)
        self.num_patches = (self.image_size[0] // patch_size) * (self.image_size[1] // patch_size)
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
        self.proj = nn.Linear(self.num_channels * self.patch_size[0] * self.patch_size[1], dim_tokens, bias=False)

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def project_patches(self, img):
        """rearrange 'b d (nh ph) (nw pw) -> b (nh nw) (ph pw d)' + Linear (reference :301) -> bf16 [B*nh*nw, D].
        All patches are projected (as in the reference); the gather kernel then reads only the kept rows."""
        B, C, H, W = img.shape
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        assert (H % self.patch_size[0] == 0) and (W % self.patch_size[1] == 0), \
            f'Image sizes {H}x{W} must be divisible by patch sizes {self.patch_size[0]}x{self.patch_size[1]}'
        assert self.patch_size[0] == self.patch_size[1], "square patches only on the B200 path"
        if img.dtype == torch.uint8:
            # raw 8-bit pixels: the loader's ToTensor + Normalize (ImageNet mean / std, fourm/data/modality_transforms.py RGBTransform) is
            # applied inside the patchify kernel, so the batch crosses PCIe as 1 byte per value (b200fm.masking)
            patches = ops.patchify_u8(img.contiguous(), self.patch_size[0])
        else:
            patches = ops.patchify(img.float().contiguous(), self.patch_size[0])
        if type(self.proj) is nn.Linear:
            return BF.LinearFn.apply(patches, self.proj.weight, self.proj.bias)
        return self.proj(patches)

    def segment(self, d, mask_key='input_mask', decoder_side=False):
        img = d['tensor']
        B = img.shape[0]
        x_rows = self.project_patches(img)
        L = x_rows.shape[0] // B
        mask = d.get(mask_key)
        if mask is None:
            mask = torch.zeros(B, L, dtype=torch.bool, device=img.device)
        st = dict(mask=as_mask_u8(mask), ids=None, L=L, kind=KIND_IMG, pos_emb=self.pos_emb, padding_idx=-1, max_length=0)
        return st, x_rows, self.mod_emb

    def forward(self, d: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        d['x'], d['emb'] = materialise(self, d, 'input_mask', False)
        return d


class SequenceEmbEncoderEmbedding(nn.Module):
    """Pre-computed sequence features such as T5-XXL embeddings (reference encoder_embeddings.py:312-421; 4M-21 only).
    The projection is one GEMM over all positions; the selection / embedding kernels read its rows (segment kind SEQ_EMB:
    feature rows like the pixel patches, positions ranked among the valid inputs like token sequences)."""

    def __init__(self, max_length: int, dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True, max_sincos_pos_emb: int = 512,
                 padding_idx: int = 0, orig_emb_dim: int = 4096, bottleneck_dim: int = 64, use_bottleneck: bool = False):
        super().__init__()
        self.max_length, self.dim_tokens, self.sincos_pos_emb = max_length, dim_tokens, sincos_pos_emb
        self.padding_idx, self.max_sincos_pos_emb, self.orig_emb_dim = padding_idx, max_sincos_pos_emb, orig_emb_dim
        self.use_bottleneck = use_bottleneck
        if self.use_bottleneck:
            self.bottleneck_dim = bottleneck_dim
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
        if self.use_bottleneck:
            self.emb_proj = nn.Sequential(nn.Linear(self.orig_emb_dim, self.bottleneck_dim), nn.Linear(self.bottleneck_dim, dim_tokens))
        else:
            self.emb_proj = nn.Linear(self.orig_emb_dim, dim_tokens)

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def project(self, feats):
        """emb_proj (Linear orig_emb_dim -> D, or the bottleneck pair) on every position -> bf16 [B*L, D] (reference :403)."""
        x = feats.reshape(-1, feats.shape[-1])
        layers = list(self.emb_proj) if isinstance(self.emb_proj, nn.Sequential) else [self.emb_proj]
        for lin in layers:
            x = BF.LinearFn.apply(x, lin.weight, lin.bias)
        return x

    def segment(self, d, mask_key='input_mask', decoder_side=False):
        feats = d['tensor']
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        B, L = feats.shape[0], feats.shape[1]
        x_rows = self.project(feats)
        pos = self.pos_emb if self.pos_emb.dim() == 2 else self.pos_emb.reshape(-1, self.dim_tokens)
        st = dict(mask=as_mask_u8(d[mask_key]), ids=None, L=L, kind=KIND_SEQ_EMB, pos_emb=pos, padding_idx=-1, max_length=0)
        return st, x_rows, self.mod_emb

    def forward(self, d):
        d['x'], d['emb'] = materialise(self, d, 'input_mask', False)
        return d
