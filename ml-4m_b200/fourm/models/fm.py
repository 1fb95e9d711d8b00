"""4M encoder-decoder (FourM / FM) on the B200-native kernels, behind the reference's API.

Drop-in for `fourm/models/fm.py` of apple/ml-4m: same class names, constructor arguments, state_dict keys, registered
model names and public methods (`forward`, `forward_encoder`, `forward_decoder`, `forward_logits`, `cat_*`,
`forward_mask_*`, freeze helpers), so `run_training_4m.py` (through `fourm.utils.create_model`) and the generation code
use it unchanged.  The training forward (reference fm.py:640-691) is restructured for the hardware:

  * all modalities are selected and embedded by two kernels per side (stable-partition PLAN + row EMBED) that write only
    the kept [B, N, D] rows -- the reference materialises and gathers [B, 2204, D] twice;
  * the block stack runs on tcgen05 GEMMs with fused epilogues and a fused attention kernel (fm_utils.py of this overlay);
  * the masked-token head gathers the per-modality row sets on the device, fuses logits + cross-entropy, and needs ONE
    host read (the 7 row counts) where the reference syncs ~15 times per step.

Python-`random` is consumed exactly like the reference (one `random.sample` over the decoder modalities per call) so
seeded runs select the same tokens.
"""
import copy
import math
import random
from functools import partial
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import nn

from b200fm import functional as BF
from b200fm import lib, ops
from b200fm.compat import MODALITY_INFO, PyTorchModelHubMixin, register_model

from .fm_utils import Block, DecoderBlock, LayerNorm, _linear_residual, _norm_bf16, _norm_params, _settle

# the 13 registered model names of the reference (fm.py:33-50); `register_model` appends them to __all__
_PRESET_NAMES = (
    'fm_tiny_6e_6d_gelu', 'fm_small_8e_8d_gelu', 'fm_base_12e_12d_gelu', 'fm_large_24e_24d_gelu', 'fm_xlarge_24e_24d_gelu',
    'fm_tiny_6e_6d_swiglu_nobias', 'fm_small_8e_8d_swiglu_nobias', 'fm_base_12e_12d_swiglu_nobias',
    'fm_large_24e_24d_swiglu_nobias', 'fm_xlarge_24e_24d_swiglu_nobias',
    'fm_base_12e_12d_swiglu_qknorm_nobias', 'fm_large_24e_24d_swiglu_qknorm_nobias', 'fm_xlarge_24e_24d_swiglu_qknorm_nobias',
)
__all__ = []

_SEQ_TYPES = ('seq', 'seq_emb', 'seq_token')


class FourM(nn.Module):
    """4M model (reference fm.py:54-691).  Constructor arguments are the reference's."""

    def __init__(self,
                 encoder_embeddings: Dict[str, nn.Module],
                 decoder_embeddings: Dict[str, nn.Module],
                 modality_info: Dict[str, Any],
                 dim: int = 768,
                 encoder_depth: int = 12,
                 decoder_depth: int = 12,
                 num_heads: int = 12,
                 mlp_ratio: float = 4.0,
                 qkv_bias: bool = True,
                 proj_bias: bool = True,
                 mlp_bias: bool = True,
                 drop_path_rate_encoder: float = 0.0,
                 drop_path_rate_decoder: float = 0.0,
                 shared_drop_path: bool = False,
                 act_layer: nn.Module = nn.GELU,
                 norm_layer: Union[partial, nn.Module] = partial(LayerNorm, eps=1e-6),
                 gated_mlp: bool = False,
                 qk_norm: bool = False,
                 decoder_causal_mask: bool = False,
                 decoder_sep_mask: bool = True,
                 num_register_tokens: int = 0,
                 use_act_checkpoint: bool = False,
                 share_modality_embeddings: bool = True,
                 ):
        super().__init__()
        self.modality_info = modality_info
        self.dim = dim
        self.decoder_causal_mask = decoder_causal_mask
        self.decoder_sep_mask = decoder_sep_mask
        self.init_std = 0.02
        self.use_act_checkpoint = use_act_checkpoint
        self.num_register_tokens = num_register_tokens
        # static_head: per-modality row counts of the masked-token head stay ON THE DEVICE (no host synchronisation, every launch has
        # a batch-independent shape -> the whole step can be captured in a CUDA graph, b200fm.graph.GraphedTrainStep).
        # Differences to the default path: an empty modality's loss is a 0-d zero instead of the reference's `zeros(1)`.
        self.static_head = False
        self._decoder_order_dev = None      # device int32 [n_decoder_mods]: decoder shuffle as data (set by GraphedTrainStep)

        self.encoder_modalities = set(encoder_embeddings.keys())
        for emb in encoder_embeddings.values():
            emb.init(dim_tokens=dim, init_std=self.init_std)
        self.encoder_embeddings = nn.ModuleDict(encoder_embeddings)

        self.decoder_modalities = set(decoder_embeddings.keys())
        for emb in decoder_embeddings.values():
            emb.init(dim_tokens=dim, init_std=self.init_std)
        self.decoder_embeddings = nn.ModuleDict(decoder_embeddings)

        if share_modality_embeddings:
            self.share_modality_embeddings()

        total = encoder_depth + decoder_depth
        if shared_drop_path:
            dpr_encoder = [x.item() for x in torch.linspace(0, drop_path_rate_encoder, total)][:encoder_depth]
            dpr_decoder = [x.item() for x in torch.linspace(0, drop_path_rate_decoder, total)][encoder_depth:]
        else:
            dpr_encoder = [x.item() for x in torch.linspace(0, drop_path_rate_encoder, encoder_depth)]
            dpr_decoder = [x.item() for x in torch.linspace(0, drop_path_rate_decoder, decoder_depth)]
        blk = dict(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, proj_bias=proj_bias, mlp_bias=mlp_bias,
                   act_layer=act_layer, norm_layer=norm_layer, gated_mlp=gated_mlp, qk_norm=qk_norm)
        self.encoder = nn.ModuleList([Block(drop_path=dpr_encoder[i], **blk) for i in range(encoder_depth)])
        self.encoder_norm = norm_layer(dim)
        self.decoder_proj_context = nn.Linear(dim, dim)
        self.decoder = nn.ModuleList([DecoderBlock(drop_path=dpr_decoder[i], **blk) for i in range(decoder_depth)])
        self.decoder_norm = norm_layer(dim)

        self.mask_token = nn.Parameter(torch.zeros(1, 1, dim))
        nn.init.normal_(self.mask_token, std=self.init_std)
        if self.num_register_tokens > 0:
            self.register_tokens = nn.Parameter(torch.zeros(1, self.num_register_tokens, dim))
            nn.init.normal_(self.register_tokens, std=self.init_std)
        else:
            self.register_tokens = None
        self.init_weights()

    # ------------------------------------------------------------------ parameters / bookkeeping
    def share_modality_embeddings(self):
        """One mod_emb per modality shared by its encoder and decoder embedding (reference fm.py:176-180)."""
        for mod in self.encoder_modalities & self.decoder_modalities:
            self.decoder_embeddings[mod].mod_emb = self.encoder_embeddings[mod].mod_emb

    def init_weights(self):
        """MAE-style init (reference fm.py:182-216): xavier-uniform Linear with per-Q/K/V fan-out for fused qkv / kv."""
        for name, m in self.named_modules():
            if "tokenizer" in name:
                continue
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
            elif isinstance(m, (nn.LayerNorm, LayerNorm)):
                nn.init.constant_(m.weight, 1.0)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, std=self.init_std)
            elif isinstance(m, nn.Conv2d) and '.proj' in name:
                w = m.weight.data
                nn.init.xavier_uniform_(w.view([w.shape[0], -1]))

    def get_num_layers_encoder(self):
        return len(self.encoder)

    def get_num_layers_decoder(self):
        return len(self.decoder)

    def get_num_layers(self):
        return self.get_num_layers_encoder() + self.get_num_layers_decoder()

    @torch.jit.ignore
    def no_weight_decay(self):
        skip = set()
        for side, embs in (("encoder_embeddings", self.encoder_embeddings), ("decoder_embeddings", self.decoder_embeddings)):
            for mod, emb_module in embs.items():
                if hasattr(emb_module, 'no_weight_decay'):
                    skip |= {f'{side}.{mod}.{name}' for name in emb_module.no_weight_decay()}
        return skip

    def _mod_type(self, mod):
        return self.modality_info[mod]['type']

    def _mod_id(self, mod):
        return int(self.modality_info[mod]['id'])

    # ------------------------------------------------------------------ fused selection + embedding (training path)
    def _embed_side(self, mod_dict, decoder: bool, n_keep: int, order, order_dev=None):
        """PLAN + EMBED kernels over the modalities in `order` (or, with order_dev, in the device-side permutation of `order`).
        Returns (x0, emb, plan)."""
        embs = self.decoder_embeddings if decoder else self.encoder_embeddings
        seg_static, tensors = [], []
        B = None
        for mod in order:
            d = mod_dict[mod]
            st, main, mod_emb = embs[mod].segment(d, 'target_mask' if decoder else 'input_mask')
            st["mod_id"] = self._mod_id(mod)
            if decoder:
                if self._mod_type(mod) in _SEQ_TYPES:
                    st["kind"] = lib.KIND_SEQ
                if "dam" not in st:
                    raise KeyError(f"modality {mod}: 'decoder_attention_mask' missing from mod_dict")
            seg_static.append(st)
            tensors += [main, mod_emb, embs[mod].pos_emb]
            B = d['tensor'].shape[0]
        dev = tensors[1].device
        if dev.type != "cuda":
            raise lib.B200FMError("FourM on the B200 path needs CUDA tensors (there is no CPU fallback)")
        plan = ops.select_plan(seg_static, ops.MODE_DECODER if decoder else 0, B, n_keep, dev, order_dev=order_dev)
        x0, emb = BF.EmbedRowsFn.apply(plan, seg_static, self.dim, not decoder, self.mask_token if decoder else None, *tensors)
        return x0, emb, plan

    # ------------------------------------------------------------------ reference-shaped helpers (generation callers)
    def cat_encoder_tensors(self, mod_dict):
        """Reference fm.py:245-277 on already materialised per-modality dicts ('x', 'emb', 'input_mask')."""
        toks, embs, masks, mods = [], [], [], []
        for mod, d in mod_dict.items():
            toks.append(d['x']); embs.append(d['emb']); masks.append(d['input_mask'])
            mods.append(torch.full_like(d['input_mask'], self._mod_id(mod), dtype=torch.int16))
        return torch.cat(toks, dim=1), torch.cat(embs, dim=1), torch.cat(masks, dim=1), torch.cat(mods, dim=1)

    def cat_decoder_tensors(self, mod_dict):
        """Reference fm.py:279-336 (consumes one `random.sample`, like the reference)."""
        toks, tgts, embs, masks, dams, mods = [], [], [], [], [], []
        items = list(mod_dict.items())
        for mod, d in random.sample(items, len(items)):
            mid = self._mod_id(mod)
            if self._mod_type(mod) in _SEQ_TYPES:
                toks.append(d['x'][:, :-1]); tgts.append(d['ids'][:, 1:]); embs.append(d['emb'][:, :-1])
                masks.append(torch.logical_or(d['target_mask'][:, 1:], d['target_mask'][:, :-1]))
                dams.append(d['decoder_attention_mask'][:, :-1])
                mods.append(torch.full_like(d['ids'][:, :-1], mid, dtype=torch.int16))
            else:
                toks.append(torch.zeros_like(d['x']) + self.mask_token); tgts.append(d['ids']); embs.append(d['emb'])
                masks.append(d['target_mask']); dams.append(d['decoder_attention_mask'])
                mods.append(torch.full_like(d['ids'], mid, dtype=torch.int16))
        return (torch.cat(toks, dim=1), torch.cat(embs, dim=1), torch.cat(masks, dim=1), torch.cat(tgts, dim=1),
                torch.cat(dams, dim=1), torch.cat(mods, dim=1))

    @staticmethod
    def _stable_keep(mask, k):
        L = mask.shape[1]
        key = mask.long() * L + torch.arange(L, device=mask.device)[None]
        return torch.argsort(key, dim=1, stable=True)[:, :k]

    def forward_mask_encoder(self, mod_dict, num_encoder_tokens):
        """Reference fm.py:338-390 on materialised dicts (the training path uses the fused kernels instead)."""
        B = list(mod_dict.values())[0]['tensor'].shape[0]
        toks, embs, masks, mods = self.cat_encoder_tensors(mod_dict)
        keep = self._stable_keep(masks, num_encoder_tokens)
        gi = keep[..., None].expand(-1, -1, toks.shape[2])
        tok, emb = torch.gather(toks, 1, gi), torch.gather(embs, 1, gi)
        msk, mod = torch.gather(masks, 1, keep), torch.gather(mods, 1, keep)
        if self.num_register_tokens > 0:
            reg = self.register_tokens.expand(B, -1, -1)
            tok = torch.cat([reg.to(tok.dtype), tok], dim=1)
            emb = torch.cat([torch.zeros_like(reg).to(emb.dtype), emb], dim=1)
            msk = torch.cat([torch.zeros((B, reg.shape[1]), dtype=torch.bool, device=msk.device), msk], dim=1)
            mod = torch.cat([torch.full((B, reg.shape[1]), -1, dtype=torch.int16, device=mod.device), mod], dim=1)
        tok = tok.masked_fill(msk[..., None], 0.)
        emb = emb.masked_fill(msk[..., None], 0.)
        mod = mod.masked_fill(msk, -1)
        return tok, emb, msk[:, None, :], mod

    def forward_mask_decoder(self, mod_dict, num_decoder_tokens):
        """Reference fm.py:392-438 on materialised dicts."""
        toks, embs, masks, tgts, dams, mods = self.cat_decoder_tensors(mod_dict)
        keep = self._stable_keep(masks, num_decoder_tokens)
        gi = keep[..., None].expand(-1, -1, toks.shape[2])
        tok, emb = torch.gather(toks, 1, gi), torch.gather(embs, 1, gi)
        msk, tgt = torch.gather(masks, 1, keep), torch.gather(tgts, 1, keep)
        dam, mod = torch.gather(dams, 1, keep), torch.gather(mods, 1, keep)
        tok = tok.masked_fill(msk[..., None], 0.)
        emb = emb.masked_fill(msk[..., None], 0.)
        tgt = tgt.masked_fill(msk, 0)
        amask = self.adapt_decoder_attention_mask(dam, mod)
        mod = mod.masked_fill(msk, -1)
        return tok, emb, msk[:, None, :], tgt, amask, mod

    def adapt_decoder_attention_mask(self, decoder_attention_mask, mod_mask=None):
        """Compressed -> dense [B, M, M] decoder self-attention mask (reference fm.py:440-475), one kernel."""
        dam = decoder_attention_mask.to(torch.int32).contiguous()
        if mod_mask is None:
            mod_mask = torch.zeros_like(dam, dtype=torch.int16)
        return ops.decoder_attention_mask(dam, mod_mask.to(torch.int16).contiguous(), self.decoder_causal_mask, self.decoder_sep_mask)

    # ------------------------------------------------------------------ transformer stacks
    def forward_encoder(self, x: torch.Tensor, encoder_mask: torch.Tensor) -> torch.Tensor:
        """Reference fm.py:477-495.  x fp32 [B, N, D], encoder_mask bool [B, 1, N] -> fp32 [B, N, D]."""
        for blk in self.encoder:
            x = blk(x, mask=encoder_mask)
        return self.encoder_norm(x)

    def forward_decoder(self, y, context, encoder_mask, decoder_attention_mask):
        """Reference fm.py:497-519."""
        for blk in self.decoder:
            y = blk(y, context, sa_mask=decoder_attention_mask, xa_mask=encoder_mask)
        return self.decoder_norm(y)

    def _block_pending(self, blk, *args, **kw):
        """blk.forward_pending, recomputed in backward when `use_act_checkpoint` is set (the reference stores the flag, fm.py:113,
        and leaves the wrapping to its FSDP launcher; here the flag alone trades the block's saved activations for a second forward)."""
        if self.use_act_checkpoint and torch.is_grad_enabled():
            from torch.utils.checkpoint import checkpoint
            return checkpoint(lambda *a: blk.forward_pending(*a, **kw), *args, use_reentrant=False)
        return blk.forward_pending(*args, **kw)

    def _encoder_to_context(self, x, encoder_mask, encoder_emb):
        """encoder blocks -> encoder_norm -> decoder_proj_context(x) + encoder_emb (reference fm.py:678-679); the norm
        emits the bf16 GEMM operand and the `+ encoder_emb` rides in the GEMM epilogue."""
        ypend = None
        for blk in self.encoder:
            x, ypend = self._block_pending(blk, x, ypend, mask=encoder_mask)
        np_ = _norm_params(self.encoder_norm)
        if np_ is not None and type(self.decoder_proj_context) is nn.Linear and x.dtype == torch.float32:
            return BF.NormLinearResidualFn.apply(x, ypend, np_[0], np_[1], self.decoder_proj_context.weight,
                                                 self.decoder_proj_context.bias, encoder_emb, np_[2])
        x = _settle(x, ypend)
        return _linear_residual(self.decoder_proj_context, _norm_bf16(self.encoder_norm, x), encoder_emb)

    # ------------------------------------------------------------------ heads
    def forward_logits(self, y, decoder_mod_dict, decoder_mod_mask, return_all_logits: bool = False):
        """Reference fm.py:521-546."""
        mod_logits = {}
        for mod in decoder_mod_dict:
            if return_all_logits:
                mod_logits[mod] = self.decoder_embeddings[mod].forward_logits(y)
            else:
                mod_logits[mod] = self.decoder_embeddings[mod].forward_logits(y[decoder_mod_mask == self._mod_id(mod)])
        return mod_logits

    def _head_index_sets(self, decoder_mods, decoder_mod_mask):
        """Row-index lists per modality + their counts.  Launched right after the decoder selection, long before the head needs
        them: the counts travel to pinned host memory asynchronously while the block stack runs, so reading them later does not
        drain the GPU queue (the reference syncs ~15 times per step here, fm.py:591-596 + run_training_4m.py:726-727)."""
        dev = decoder_mod_mask.device
        key = (tuple(decoder_mods), str(dev))
        cache = self.__dict__.setdefault("_head_ids_cache", {})
        if key not in cache:
            cache[key] = torch.tensor([self._mod_id(m) for m in decoder_mods], device=dev, dtype=torch.int32)
        rows, counts = ops.head_rows(decoder_mod_mask.reshape(-1), cache[key])
        pin = self.__dict__.setdefault("_head_pinned", {})
        slot = pin.get("i", 0) ^ 1                               # two pinned buffers, alternated (cudaHostAlloc is slow)
        pin["i"] = slot
        host = pin.get(slot)
        if host is None or host.numel() != len(decoder_mods):
            host = pin[slot] = torch.empty(len(decoder_mods), dtype=torch.int32, pin_memory=True)
        host.copy_(counts, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return rows, host, ev

    def _head_losses(self, y_bf16, target_ids, decoder_mods, decoder_mod_mask, index_sets=None):
        """Per-modality mean cross-entropy (reference fm.py:589-600) with device-side index sets and one host read."""
        dev = y_bf16.device
        rows, host, ev = index_sets if index_sets is not None else self._head_index_sets(decoder_mods, decoder_mod_mask)
        ev.synchronize()                                   # the single device->host hand-over of the head (already complete)
        counts_host = host.tolist()
        y2 = y_bf16.reshape(-1, y_bf16.shape[-1])
        parts = BF.HeadGatherFn.apply(y2, rows, tuple(counts_host))
        tflat = target_ids.reshape(-1)
        mod_loss, mod_count = {}, {}
        for i, mod in enumerate(decoder_mods):
            n = counts_host[i]
            emb = self.decoder_embeddings[mod]
            V = emb.to_logits.weight.shape[0]
            if n == 0:
                mod_loss[mod] = torch.zeros(1, device=dev)                     # reference fm.py:593-595
                mod_count[mod] = 0
                continue
            tgt = ops.gather_i64(tflat, rows[i], n)
            if type(emb.to_logits) is nn.Linear and emb.to_logits.bias is None:
                mod_loss[mod] = BF.LinearCrossEntropyFn.apply(parts[i], emb.to_logits.weight, tgt)
            else:
                mod_loss[mod] = F.cross_entropy(emb.forward_logits(parts[i]).float(), tgt, reduction='mean')
            mod_count[mod] = n * V                                                # logits.numel() in the reference
        return mod_loss, mod_count

    def _head_losses_static(self, y_bf16, target_ids, decoder_mods, decoder_mod_mask):
        """_head_losses without any host read: device-side counts drive the per-modality logits / cross-entropy / gradient launches
        (b200fm_gemm_bf16_dyn).  Returns ({mod: 0-d loss}, {mod: 0-d logits.numel() as fp32})."""
        dev = y_bf16.device
        cache = self.__dict__.setdefault("_head_ids_cache", {})
        key = (tuple(decoder_mods), str(dev))
        if key not in cache:
            cache[key] = torch.tensor([self._mod_id(m) for m in decoder_mods], device=dev, dtype=torch.int32)
        rows, counts = ops.head_rows(decoder_mod_mask.reshape(-1), cache[key])
        y2 = y_bf16.reshape(-1, y_bf16.shape[-1])
        R = y2.shape[0]
        parts = BF.HeadGatherStaticFn.apply(y2, rows, counts)
        tflat = target_ids.reshape(-1)
        countf = counts.float()
        mod_loss, mod_count = {}, {}
        for i, mod in enumerate(decoder_mods):
            emb = self.decoder_embeddings[mod]
            if not (type(emb.to_logits) is nn.Linear and emb.to_logits.bias is None):
                raise NotImplementedError("static_head needs plain bias-free to_logits layers (every reference preset)")
            n_dev = counts[i:i + 1]
            tgt = ops.gather_i64(tflat, rows[i], R, n_dev)
            mod_loss[mod] = BF.LinearCrossEntropyStaticFn.apply(parts[i], emb.to_logits.weight, tgt, n_dev)
            mod_count[mod] = countf[i] * emb.to_logits.weight.shape[0]
        return mod_loss, mod_count

    def forward_loss(self, y, target_ids, decoder_mod_dict, decoder_mod_mask, loss_type):
        """Reference fm.py:548-637 (forward_mod_loss / forward_token_loss)."""
        if loss_type not in ('mod', 'modality', 'token'):
            raise ValueError("Invalid loss type")
        y_bf16 = y if y.dtype == torch.bfloat16 else ops.cast_bf16(y.float().contiguous())
        mod_loss, mod_count = self._head_losses(y_bf16, target_ids, list(decoder_mod_dict.keys()), decoder_mod_mask)
        if loss_type == 'token':
            loss = sum(mod_loss[m] * mod_count[m] for m in mod_loss) / sum(mod_count.values())
        else:
            loss = sum(mod_loss.values()) / len(mod_loss)
        return loss, mod_loss

    def forward_mod_loss(self, y, target_ids, decoder_mod_dict, decoder_mod_mask):
        return self.forward_loss(y, target_ids, decoder_mod_dict, decoder_mod_mask, 'mod')

    def forward_token_loss(self, y, target_ids, decoder_mod_dict, decoder_mod_mask):
        return self.forward_loss(y, target_ids, decoder_mod_dict, decoder_mod_mask, 'token')

    # ------------------------------------------------------------------ forward
    def forward(self, mod_dict, num_encoder_tokens: int, num_decoder_tokens: int, loss_type: str = 'mod',
                return_logits: bool = False):
        """Reference fm.py:640-691: same arguments, same returns (`(loss, {mod: loss})` or `{mod: logits[B, M, V]}`)."""
        if loss_type not in ('mod', 'modality', 'token'):
            raise ValueError("Invalid loss type")
        enc_mods = [m for m in mod_dict if m in self.encoder_embeddings]
        dec_mods = [m for m in mod_dict if m in self.decoder_embeddings]
        order_dev = self._decoder_order_dev
        if order_dev is None:
            # same RNG consumption as cat_decoder_tensors (reference fm.py:306)
            dec_order = random.sample(dec_mods, len(dec_mods))
        else:
            # the shuffle arrives as device data (drawn by the caller with the same random.sample): replayable launch
            assert order_dev.numel() == len(dec_mods) and order_dev.dtype == torch.int32
            dec_order = dec_mods

        n_reg = self.num_register_tokens
        x0, enc_emb, eplan = self._embed_side(mod_dict, False, num_encoder_tokens, enc_mods)
        encoder_mask = eplan.pad_mask
        if n_reg > 0:
            B = x0.shape[0]
            reg = self.register_tokens.expand(B, -1, -1).float()
            x0 = torch.cat([reg, x0], dim=1)
            enc_emb = torch.cat([torch.zeros_like(reg), enc_emb], dim=1)
            encoder_mask = torch.cat([torch.zeros(B, n_reg, dtype=torch.bool, device=x0.device), encoder_mask], dim=1)
        encoder_mask = encoder_mask[:, None, :]

        y0, _, dplan = self._embed_side(mod_dict, True, num_decoder_tokens, dec_order, order_dev=order_dev)
        dec_attn_mask = ops.decoder_attention_mask(dplan.dam, dplan.mod_raw, self.decoder_causal_mask, self.decoder_sep_mask)
        index_sets = None if (return_logits or self.static_head) else self._head_index_sets(dec_mods, dplan.mod_mask)

        context = self._encoder_to_context(x0, encoder_mask, enc_emb)
        y, ypend = y0, None
        for blk in self.decoder:
            y, ypend = self._block_pending(blk, y, ypend, context, sa_mask=dec_attn_mask, xa_mask=encoder_mask)
        np_ = _norm_params(self.decoder_norm)
        if np_ is not None and y.dtype == torch.float32:
            y = BF.AddLayerNormFn.apply(y, ypend, np_[0], np_[1], np_[2])      # bf16: the head GEMMs' operand
        else:
            y = _norm_bf16(self.decoder_norm, _settle(y, ypend))

        if return_logits:
            return {mod: self.decoder_embeddings[mod].forward_logits(y) for mod in dec_mods}
        if self.static_head:
            mod_loss, mod_count = self._head_losses_static(y, dplan.target_ids, dec_mods, dplan.mod_mask)
        else:
            mod_loss, mod_count = self._head_losses(y, dplan.target_ids, dec_mods, dplan.mod_mask, index_sets)
        if loss_type == 'token':
            loss = sum(mod_loss[m] * mod_count[m] for m in mod_loss) / sum(mod_count.values())
        else:
            loss = sum(mod_loss.values()) / len(mod_loss)
        return loss, mod_loss

    # ------------------------------------------------------------------ freeze helpers (reference fm.py:694-776)
    def _set_grad(self, modules, flag):
        for m in modules:
            for p in m.parameters():
                p.requires_grad = flag

    def freeze_encoder(self, freeze_embeddings=True):
        self._set_grad([self.encoder, self.encoder_norm] + ([self.encoder_embeddings] if freeze_embeddings else []), False)

    def freeze_encoder_except_specific_embeddings(self, frozen_embedding_domain):
        doms = frozen_embedding_domain.split('-')
        self._set_grad([self.encoder, self.encoder_norm], False)
        for name, param in self.encoder_embeddings.named_parameters():
            if name.split('.')[0] in doms:
                param.requires_grad = False

    def unfreeze_encoder(self, unfreeze_embeddings=True):
        self._set_grad([self.encoder, self.encoder_norm] + ([self.encoder_embeddings] if unfreeze_embeddings else []), True)

    def freeze_decoder(self, freeze_embeddings=True):
        self._set_grad([self.decoder, self.decoder_norm] + ([self.decoder_embeddings] if freeze_embeddings else []), False)

    def freeze_decoder_except_specific_embeddings(self, frozen_embedding_domain):
        doms = frozen_embedding_domain.split('-')
        self._set_grad([self.decoder, self.decoder_norm], False)
        for name, param in self.decoder_embeddings.named_parameters():
            if name.split('.')[0] in doms:
                param.requires_grad = False

    def unfreeze_decoder(self, unfreeze_embeddings=True):
        self._set_grad([self.decoder, self.decoder_norm] + ([self.decoder_embeddings] if unfreeze_embeddings else []), True)

    def freeze_shared_params(self):
        self.freeze_encoder(freeze_embeddings=False)
        self.freeze_decoder(freeze_embeddings=False)

    def freeze_params_except_specific_embeddings(self, frozen_embedding_domain):
        self.freeze_encoder_except_specific_embeddings(frozen_embedding_domain=frozen_embedding_domain)
        self.freeze_decoder_except_specific_embeddings(frozen_embedding_domain=frozen_embedding_domain)

    def unfreeze_shared_params(self):
        self.unfreeze_encoder(unfreeze_embeddings=False)
        self.unfreeze_decoder(unfreeze_embeddings=False)

    def unfreeze_all(self):
        self.unfreeze_encoder(unfreeze_embeddings=True)
        self.unfreeze_decoder(unfreeze_embeddings=True)


class FM(FourM, PyTorchModelHubMixin):
    """Hugging Face Hub wrapper (reference fm.py:783-831): `FM(config)` with keys domains_in, domains_out, image_size,
    patch_size, norm_bias, act_layer + FourM kwargs; decoder embeddings are built with share_embedding=False."""

    def __init__(self, config: dict):
        config = copy.deepcopy(config)
        all_domains = sorted(set(config['domains_in']) | set(config['domains_out']))
        modality_info = {mod: MODALITY_INFO[mod] for mod in all_domains}
        encoder_embeddings, decoder_embeddings = {}, {}
        for side, doms, out in (("encoder_embedding", config['domains_in'], encoder_embeddings),
                                ("decoder_embedding", config['domains_out'], decoder_embeddings)):
            for mod in doms:
                info = modality_info[mod]
                if info.get(side) is None:
                    continue
                kw = {} if side == "encoder_embedding" else dict(share_embedding=False)
                if info["type"] == "img":
                    kw.update(patch_size=info.get('patch_size', config['patch_size']), image_size=info.get('input_size', config['image_size']))
                out[mod] = info[side](**kw)
        config['norm_layer'] = partial(LayerNorm, eps=1e-6, bias=config['norm_bias'])
        config['act_layer'] = getattr(torch.nn, config['act_layer'])
        for k in ('norm_bias', 'domains_in', 'domains_out', 'image_size', 'patch_size'):
            del config[k]
        super().__init__(encoder_embeddings=encoder_embeddings, decoder_embeddings=decoder_embeddings,
                         modality_info=modality_info, **config)


# ---------------------------------------------------------------------------------------------------------------------
# registered presets (reference fm.py:839-1130): (depth_enc, depth_dec, dim, heads)
# ---------------------------------------------------------------------------------------------------------------------
_SIZES = {'tiny_6e_6d': (6, 6, 384, 6), 'small_8e_8d': (8, 8, 512, 8), 'base_12e_12d': (12, 12, 768, 12),
          'large_24e_24d': (24, 24, 1024, 16), 'xlarge_24e_24d': (24, 24, 2048, 32)}


def _make_preset(size_key: str, family: str):
    e, d, dim, heads = _SIZES[size_key]
    if family == 'gelu':
        fixed = dict(qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6))
    else:
        fixed = dict(qkv_bias=False, proj_bias=False, mlp_bias=False, norm_layer=partial(LayerNorm, eps=1e-6, bias=False),
                     act_layer=nn.SiLU, gated_mlp=True)
        if family == 'swiglu_qknorm_nobias':
            fixed['qk_norm'] = True

    def entrypoint(encoder_embeddings: Dict[str, nn.Module], decoder_embeddings: Dict[str, nn.Module], **kwargs):
        return FourM(encoder_embeddings=encoder_embeddings, decoder_embeddings=decoder_embeddings, encoder_depth=e,
                     decoder_depth=d, dim=dim, num_heads=heads, mlp_ratio=4, **fixed, **kwargs)

    entrypoint.__name__ = entrypoint.__qualname__ = f'fm_{size_key}_{family}'
    entrypoint.__module__ = __name__
    return register_model(entrypoint)


for _name in _PRESET_NAMES:
    _parts = _name.split('_')
    globals()[_name] = _make_preset('_'.join(_parts[1:4]), '_'.join(_parts[4:]))
