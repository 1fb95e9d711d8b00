"""VQ tokenizer (encoder + quantizer) on the B200 kernels, behind the reference's `fourm.vq.vqvae.VQ` surface.

Drop-in for the tokenization path of `fourm/vq/vqvae.py` (apple/ml-4m): `VQ(...)` keeps its constructor arguments and
state_dict keys (`encoder.*`, `quant_proj.{weight,bias}`, `quantize._codebook.*`), `encode` / `tokenize` /
`tokens_to_embedding` keep their signatures, so `get_image_tokenizer` checkpoints and `save_vq_tokens.py` work unchanged.
`VQVAE` / `DiVAE` / `VQControlNet` (decoders: detokenization / tokenizer training) are out of the hot path; when the
reference tree is importable they are re-exported from it by `fourm/vq/__init__.py`."""
import copy
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from einops import rearrange

from b200fm import functional as BF
from b200fm import ops
from b200fm.compat import PyTorchModelHubMixin

from .models import vit_models
from .quantizers.quantize_lucid import VectorQuantize as VectorQuantizerLucid

FREEZE_MODULES = ['encoder', 'quant_proj', 'quantize', 'cls_emb']
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def denormalize(img, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    m = torch.tensor(mean, device=img.device, dtype=img.dtype)[None, :, None, None]
    s = torch.tensor(std, device=img.device, dtype=img.dtype)[None, :, None, None]
    return img * s + m


class VQ(nn.Module, PyTorchModelHubMixin):
    """Encoder + quantizer (reference vqvae.py:39-393)."""

    def __init__(self, image_size: int = 224, image_size_enc: Optional[int] = None, n_channels: str = 3, n_labels: Optional[int] = None,
                 enc_type: str = 'vit_b_enc', patch_proj: bool = True, post_mlp: bool = False, patch_size: int = 16,
                 quant_type: str = 'lucid', codebook_size: Union[int, str] = 16384, num_codebooks: int = 1, latent_dim: int = 32,
                 norm_codes: bool = True, norm_latents: bool = False, sync_codebook: bool = True, ema_decay: float = 0.99,
                 threshold_ema_dead_code: float = 0.25, code_replacement_policy: str = 'batch_random', commitment_weight: float = 1.0,
                 kmeans_init: bool = False, ckpt_path: Optional[str] = None,
                 ignore_keys: List[str] = ['decoder', 'loss', 'post_quant_conv', 'post_quant_proj', 'encoder.pos_emb'],
                 freeze_enc: bool = False, undo_std: bool = False, config: Optional[Dict[str, Any]] = None, **kwargs):
        if config is not None:
            self.__init__(**copy.deepcopy(config))
            return
        super().__init__()
        self.image_size, self.n_channels, self.n_labels, self.enc_type = image_size, n_channels, n_labels, enc_type
        self.patch_proj, self.post_mlp, self.patch_size, self.quant_type = patch_proj, post_mlp, patch_size, quant_type
        self.codebook_size, self.num_codebooks, self.latent_dim = codebook_size, num_codebooks, latent_dim
        self.norm_codes, self.norm_latents, self.sync_codebook, self.ema_decay = norm_codes, norm_latents, sync_codebook, ema_decay
        self.threshold_ema_dead_code, self.code_replacement_policy = threshold_ema_dead_code, code_replacement_policy
        self.commitment_weight, self.kmeans_init, self.ckpt_path, self.ignore_keys = commitment_weight, kmeans_init, ckpt_path, ignore_keys
        self.freeze_enc, self.undo_std = freeze_enc, undo_std
        if n_labels is not None:
            self.cls_emb = nn.Embedding(num_embeddings=n_labels, embedding_dim=n_channels)
            self.colorize = torch.randn(3, n_labels, 1, 1)
        else:
            self.cls_emb = None
        image_size_enc = image_size_enc or image_size
        if 'vit' not in enc_type:
            raise NotImplementedError(f'{enc_type}: only the ViT encoders are on the B200 path')
        self.encoder = getattr(vit_models, enc_type)(in_channels=n_channels, patch_size=patch_size, resolution=image_size_enc,
                                                     patch_proj=patch_proj, post_mlp=post_mlp)
        self.enc_dim = self.encoder.dim_tokens
        self.quant_proj = torch.nn.Conv2d(self.enc_dim, self.latent_dim, 1)
        if quant_type != 'lucid':
            raise NotImplementedError(f'{quant_type}: only the lucid quantizer is on the B200 path (all shipped configs use it)')
        self.quantize = VectorQuantizerLucid(dim=latent_dim, codebook_size=codebook_size, codebook_dim=latent_dim, heads=num_codebooks,
                                             use_cosine_sim=norm_codes, threshold_ema_dead_code=threshold_ema_dead_code,
                                             code_replacement_policy=code_replacement_policy, sync_codebook=sync_codebook,
                                             decay=ema_decay, commitment_weight=self.commitment_weight, norm_latents=norm_latents,
                                             kmeans_init=kmeans_init)
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)
        if freeze_enc:
            for name, module in self.named_children():
                if name in FREEZE_MODULES:
                    for p in module.parameters():
                        p.requires_grad = False
                    module.eval()

    def train(self, mode: bool = True) -> 'VQ':
        self.training = mode
        for name, module in self.named_children():
            if self.freeze_enc and name in FREEZE_MODULES:
                continue
            module.train(mode)
        return self

    def init_from_ckpt(self, path: str, ignore_keys: List[str] = list()) -> 'VQ':
        """Reference vqvae.py:219-267 (renames legacy quant_conv keys, drops ignored prefixes, non-strict load)."""
        ckpt = torch.load(path, map_location="cpu")
        sd = ckpt['model'] if 'model' in ckpt else ckpt['state_dict']
        for old, new in (('quant_conv.0', 'quant_proj'), ('quant_conv', 'quant_proj'), ('post_quant_conv.0', 'post_quant_proj'),
                         ('post_quant_conv', 'post_quant_proj')):
            if f'{old}.weight' in sd and f'{old}.bias' in sd and f'{new}.weight' not in sd:
                sd[f'{new}.weight'], sd[f'{new}.bias'] = sd.pop(f'{old}.weight'), sd.pop(f'{old}.bias')
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        print(self.load_state_dict(sd, strict=False))
        return self

    def prepare_input(self, x: torch.Tensor) -> torch.Tensor:
        """Reference vqvae.py:269-287."""
        if self.undo_std:
            x = 2.0 * denormalize(x) - 1.0
        if self.cls_emb is not None:
            x = rearrange(self.cls_emb(x), 'b h w c -> b c h w')
        return x

    def latents(self, x: torch.Tensor) -> torch.Tensor:
        """prepare_input -> ViT encoder -> 1x1 quant_proj: fp32 latents [B, Hq, Wq, d] (channel-last, the scan's layout).
        Precision follows the caller like the reference's does: under `torch.autocast(bfloat16)` (run_training_vqvae.py) and whenever
        gradients are recorded the contractions take bf16 operands; an inference call WITHOUT autocast -- `save_vq_tokens.py:288`
        runs `tokenize` in fp32 -- uses the fp32-faithful limb arithmetic (b200fm.functional.precise), so the arg-min over the codebook
        sees the reference's latents.  B200FM_VQ_PRECISION = bf16 | x3 | x6 overrides."""
        import os
        mode = os.environ.get("B200FM_VQ_PRECISION", "auto")
        if mode == "auto":
            mode = "bf16" if (torch.is_grad_enabled() or torch.is_autocast_enabled()) else "x3"
        if mode != "bf16" and not torch.is_grad_enabled() and not BF.is_precise():
            with BF.precise(6 if mode == "x6" else 3):
                return self.latents(x)
        x = self.prepare_input(x)
        t, (Hq, Wq) = self.encoder.tokens(x)                                          # [B, N, D] fp32
        w = self.quant_proj.weight.reshape(self.latent_dim, self.enc_dim)
        z = BF.linear_f32(t, w) if BF.is_precise() else BF.LinearF32Fn.apply(t, w)
        if self.quant_proj.bias is not None:
            z = z + self.quant_proj.bias
        return z.view(t.shape[0], Hq, Wq, self.latent_dim)

    def encode(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.LongTensor]:
        """Reference vqvae.py:302-318 -> (quant [B, d, Hq, Wq], code_loss, tokens int64 [B, Hq, Wq])."""
        z = self.latents(x)
        quant, code_loss, tokens = self.quantize(z.permute(0, 3, 1, 2))
        return quant, code_loss, tokens

    def tokenize(self, x: torch.Tensor) -> torch.LongTensor:
        """Reference vqvae.py:320-331."""
        _, _, tokens = self.encode(x)
        return tokens

    def tokens_to_embedding(self, tokens: torch.LongTensor) -> torch.Tensor:
        return self.quantize.indices_to_embedding(tokens)

    def autoencode(self, x, **kwargs):
        pass

    def decode_quant(self, quant, **kwargs):
        pass

    def decode_tokens(self, tokens, **kwargs):
        return self.decode_quant(self.tokens_to_embedding(tokens), **kwargs)

    def forward(self, x: torch.Tensor, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        quant, code_loss, _ = self.encode(x)
        return quant, code_loss



class VQVAE(VQ):
    """VQ encoder + ViT decoder with a plain reconstruction objective (reference vqvae.py:395-481): the tokenizer TRAINING model
    (run_training_vqvae.py).  Same constructor arguments and state_dict keys (`decoder.*`, `post_quant_proj.{weight,bias}` on
    top of VQ's); `forward` -> (dec, code_loss).  Only the ViT decoders are on the B200 path."""

    def __init__(self, dec_type: str = 'vit_b_dec', out_conv: bool = False, image_size_dec: int = None, patch_size_dec: int = None,
                 config: Optional[Dict[str, Any]] = None, *args, **kwargs):
        if config is not None:
            self.__init__(**copy.deepcopy(config))
            return
        ckpt_path = kwargs.get('ckpt_path', None)
        kwargs['ckpt_path'] = None                                   # load after the decoder exists (reference :424-427)
        super().__init__(*args, **kwargs)
        self.ckpt_path = ckpt_path
        out_channels = self.n_channels if self.n_labels is None else self.n_labels
        if 'vit' not in dec_type:
            raise NotImplementedError(f'{dec_type}: only the ViT decoders are on the B200 path')
        self.decoder = getattr(vit_models, dec_type)(out_channels=out_channels, patch_size=patch_size_dec or self.patch_size,
                                                     resolution=image_size_dec or self.image_size, out_conv=out_conv,
                                                     post_mlp=self.post_mlp, patch_proj=self.patch_proj)
        self.dec_dim = self.decoder.dim_tokens
        self.post_quant_proj = torch.nn.Conv2d(self.latent_dim, self.dec_dim, 1)
        if self.ckpt_path is not None:
            self.init_from_ckpt(self.ckpt_path, ignore_keys=self.ignore_keys)

    def decode_quant(self, quant: torch.Tensor, **kwargs) -> torch.Tensor:
        """quant [B, d, Hq, Wq] -> image [B, C, H, W] (reference vqvae.py:441-452): 1x1 post_quant_proj as a GEMM, ViT decoder."""
        B, d, Hq, Wq = quant.shape
        z = quant.permute(0, 2, 3, 1).reshape(B * Hq * Wq, d)
        w = self.post_quant_proj.weight.reshape(self.dec_dim, self.latent_dim)
        t = BF.LinearFn.apply(z, w, self.post_quant_proj.bias)                              # bf16 [B*N, dec_dim]
        return self.decoder(t.view(B, Hq, Wq, self.dec_dim).permute(0, 3, 1, 2))

    def decode_tokens(self, tokens: torch.LongTensor, **kwargs) -> torch.Tensor:
        return self.decode_quant(self.tokens_to_embedding(tokens), **kwargs)

    def forward(self, x: torch.Tensor, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reference vqvae.py:454-471 -> (dec [B, C, H, W], code_loss)."""
        if self.freeze_enc:
            with torch.no_grad():
                quant, code_loss, _ = self.encode(x)
        else:
            quant, code_loss, _ = self.encode(x)
        return self.decode_quant(quant), code_loss

    def autoencode(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        dec, _ = self.forward(x)
        return dec
