"""Vector quantizer of the 4M tokenizers on the fused B200 codebook-scan kernel.

Drop-in for the inference surface of `fourm/vq/quantizers/quantize_lucid.py` (apple/ml-4m): `VectorQuantize`,
`CosineSimCodebook`, `EuclideanCodebook` keep their constructor arguments and buffer names (`embed`, `cluster_size`,
`initted`, `embed_avg`) so tokenizer checkpoints load unchanged.  The reference computes a full fp32 [n, K] similarity
matrix, an arg-max and a [n, K] one-hot (quantize_lucid.py:402-407 / 275-284); here one kernel scans the codebook with
fp32 FMAs and returns indices (+ the codebook rows).  Training-time codebook maintenance (EMA update, dead-code expiry,
k-means init: quantize_lucid.py:235-261, 286-299, 409-426) is not on the hot path and not implemented yet.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from b200fm import ops


def l2norm(t):
    return F.normalize(t, p=2, dim=-1)


def uniform_init(*shape):
    t = torch.empty(shape)
    nn.init.kaiming_uniform_(t)
    return t


class _Codebook(nn.Module):
    cosine = True

    def __init__(self, dim, codebook_size, kmeans_init=False, kmeans_iters=10, decay=0.8, eps=1e-5, threshold_ema_dead_code=2,
                 code_replacement_policy='batch_random', use_ddp=False, learnable_codebook=False, sample_codebook_temp=0.):
        super().__init__()
        self.decay, self.codebook_size, self.kmeans_iters, self.eps = decay, codebook_size, kmeans_iters, eps
        self.threshold_ema_dead_code, self.code_replacement_policy = threshold_ema_dead_code, code_replacement_policy
        self.sample_codebook_temp, self.use_ddp, self.learnable_codebook = sample_codebook_temp, use_ddp, learnable_codebook
        if kmeans_init:
            embed = torch.zeros(codebook_size, dim)
        else:
            embed = l2norm(uniform_init(codebook_size, dim)) if self.cosine else uniform_init(codebook_size, dim)
        self.register_buffer('initted', torch.Tensor([not kmeans_init]))
        self.register_buffer('cluster_size', torch.zeros(codebook_size))
        if not self.cosine:
            self.register_buffer('embed_avg', embed.clone())
        if learnable_codebook:
            self.embed = nn.Parameter(embed)
        else:
            self.register_buffer('embed', embed)

    @torch.no_grad()
    def scan(self, x):
        """x [..., d] -> (quantize [..., d] fp32 = embed[idx], idx int64 [...])."""
        if self.sample_codebook_temp != 0:
            raise NotImplementedError("gumbel sampling of codes (sample_codebook_temp != 0) is not supported on the B200 path")
        shape = x.shape
        flat = x.reshape(-1, shape[-1]).float().contiguous()
        idx, quant = ops.vq_argmax(flat, self.embed.detach().float().contiguous(), cosine=self.cosine, want_quant=True)
        return quant.view(shape), idx.view(shape[:-1])

    def forward(self, x):
        if self.training:
            raise NotImplementedError("codebook EMA / dead-code expiry (training mode) is not implemented on the B200 path yet; "
                                      "call .eval() for tokenization")
        return self.scan(x)


class CosineSimCodebook(_Codebook):
    """Reference quantize_lucid.py:303-428 (eval path :388-407)."""
    cosine = True


class EuclideanCodebook(_Codebook):
    """Reference quantize_lucid.py:181-301 (eval path :263-284)."""
    cosine = False


class VectorQuantize(nn.Module):
    """Reference quantize_lucid.py:432-568 (heads = 1, Identity projections when codebook_dim * heads == dim)."""

    def __init__(self, dim, codebook_size, codebook_dim=None, heads=1, decay=0.8, eps=1e-5, kmeans_init=False, kmeans_iters=10,
                 use_cosine_sim=False, threshold_ema_dead_code=0, code_replacement_policy='batch_random', channel_last=False,
                 accept_image_fmap=True, commitment_weight=1., orthogonal_reg_weight=0., orthogonal_reg_active_codes_only=False,
                 orthogonal_reg_max_codes=None, sample_codebook_temp=0., sync_codebook=False, norm_latents=False):
        super().__init__()
        self.heads = heads
        codebook_dim = codebook_dim if codebook_dim is not None else dim
        codebook_input_dim = codebook_dim * heads
        requires_projection = codebook_input_dim != dim
        self.project_in = nn.Linear(dim, codebook_input_dim) if requires_projection else nn.Identity()
        self.project_out = nn.Linear(codebook_input_dim, dim) if requires_projection else nn.Identity()
        self.eps, self.commitment_weight, self.norm_latents = eps, commitment_weight, norm_latents
        self.orthogonal_reg_weight = orthogonal_reg_weight
        cls = CosineSimCodebook if use_cosine_sim else EuclideanCodebook
        self._codebook = cls(dim=codebook_dim, codebook_size=codebook_size, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                             decay=decay, eps=eps, threshold_ema_dead_code=threshold_ema_dead_code,
                             code_replacement_policy=code_replacement_policy, use_ddp=sync_codebook,
                             learnable_codebook=orthogonal_reg_weight > 0, sample_codebook_temp=sample_codebook_temp)
        self.codebook_size, self.accept_image_fmap, self.channel_last = codebook_size, accept_image_fmap, channel_last

    @property
    def codebook(self):
        return self._codebook.embed

    def indices_to_embedding(self, indices):
        return F.embedding(indices, self.codebook).permute(0, 3, 1, 2)

    def forward(self, x):
        if self.heads != 1 or not isinstance(self.project_in, nn.Identity):
            raise NotImplementedError("multi-head / projected codebooks are not used by any shipped tokenizer config")
        if self.accept_image_fmap:
            B, C, Hq, Wq = x.shape
            z = x.permute(0, 2, 3, 1).reshape(B, Hq * Wq, C)
        elif not self.channel_last:
            z = x.transpose(1, 2)
        else:
            z = x
        if self.norm_latents:
            z = l2norm(z)
        quantize, embed_ind = self._codebook(z)
        loss = torch.tensor([0.], device=x.device)
        if self.accept_image_fmap:
            quantize = quantize.reshape(B, Hq, Wq, C).permute(0, 3, 1, 2)
            embed_ind = embed_ind.reshape(B, Hq, Wq)
        elif not self.channel_last:
            quantize = quantize.transpose(1, 2)
        return quantize, loss, embed_ind
