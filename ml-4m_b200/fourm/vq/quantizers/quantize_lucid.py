"""Vector quantizer of the 4M tokenizers on the fused B200 codebook-scan kernel.

Drop-in for the inference surface of `fourm/vq/quantizers/quantize_lucid.py` (apple/ml-4m): `VectorQuantize`,
`CosineSimCodebook`, `EuclideanCodebook` keep their constructor arguments and buffer names (`embed`, `cluster_size`,
`initted`, `embed_avg`) so tokenizer checkpoints load unchanged.  The reference computes a full fp32 [n, K] similarity
matrix, an arg-max and a [n, K] one-hot (quantize_lucid.py:402-407 / 275-284); here one kernel scans the codebook with
fp32 FMAs and returns indices (+ the codebook rows).  Training mode adds the codebook maintenance of quantize_lucid.py:286-299 /
409-426: per-code counts and latent sums by an index scatter kernel (no one-hot, no second GEMM), one packed all-reduce when
the codebook is synchronised, the EMA in one pass over the codebook, dead-code expiry; k-means init is not implemented.
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
        self._initted_seen = False
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

    # ------------------------------------------------------------------ training-side maintenance (reference :263-301, :388-426)
    def _ddp(self):
        return self.use_ddp and torch.distributed.is_available() and torch.distributed.is_initialized()

    @torch.no_grad()
    def _ema_update(self, flat, idx):
        """bins / per-code latent sums by scatter (no one-hot, no second GEMM), ONE packed all-reduce when the codebook is
        synchronised, then the EMA of `cluster_size` and `embed` (cosine: one kernel; Euclidean: via `embed_avg`)."""
        K, d = self.embed.shape
        stats = ops.vq_ema_stats(flat, idx.reshape(-1), K, self.cosine)
        if self._ddp():
            torch.distributed.all_reduce(stats)
        if self.cosine:
            ops.vq_ema_update_cosine(self.embed.data, self.cluster_size, stats, self.decay)
        else:
            bins, esum = stats[:K], stats[K:].view(K, d)
            self.cluster_size.mul_(self.decay).add_(bins, alpha=1 - self.decay)
            self.embed_avg.mul_(self.decay).add_(esum, alpha=1 - self.decay)
            total = self.cluster_size.sum()
            smoothed = (self.cluster_size + self.eps) / (total + K * self.eps) * total
            self.embed.data.copy_(self.embed_avg / smoothed.unsqueeze(1))

    def _sample(self, samples, num):
        """quantize_lucid.py:62-71 sample_vectors / :100-113 sample_vectors_distributed."""
        def local(s, k):
            n = s.shape[0]
            ind = torch.randperm(n, device=s.device)[:k] if n >= k else torch.randint(0, n, (k,), device=s.device)
            return s[ind]
        if not self._ddp():
            return local(samples, num)
        dist = torch.distributed
        world, rank = dist.get_world_size(), dist.get_rank()
        sizes = torch.zeros(world, dtype=torch.long, device=samples.device)
        sizes[rank] = samples.shape[0]
        dist.all_reduce(sizes)
        if rank == 0:       # multinomial split of `num` over the ranks, proportional to their sample counts (:75-98)
            per = torch.distributions.Multinomial(total_count=num, probs=sizes.float().cpu() / sizes.sum().item()).sample().long().to(samples.device)
        else:
            per = torch.empty(world, dtype=torch.long, device=samples.device)
        dist.broadcast(per, src=0)
        per = per.tolist()
        mine = local(samples, per[rank])
        out = []
        for r, k in enumerate(per):
            t = mine if r == rank else samples.new_empty(k, samples.shape[1])
            dist.broadcast(t, src=r)
            out.append(t)
        return torch.cat(out, dim=0)

    @torch.no_grad()
    def expire_codes_(self, batch_samples):
        """quantize_lucid.py:238-256 / 365-383: codes whose EMA cluster size fell below the threshold are re-seeded."""
        if self.threshold_ema_dead_code == 0:
            return
        expired = self.cluster_size < self.threshold_ema_dead_code
        n_exp = int(expired.sum().item())                     # host read, like the reference's torch.any / mask.sum().item()
        if n_exp == 0:
            return
        if self.code_replacement_policy == 'batch_random':
            samples = l2norm(batch_samples.reshape(-1, batch_samples.shape[-1]).float())
            self.embed.data[expired] = self._sample(samples, n_exp)
        elif self.code_replacement_policy == 'linde_buzo_gray':
            most_used = self.embed.data[self.cluster_size.argsort(descending=True)[:n_exp]]
            noise = torch.randn_like(most_used)
            if self._ddp():
                torch.distributed.broadcast(noise, src=0)
            self.embed.data[expired] = l2norm(most_used + noise * 1e-10)
        else:
            raise ValueError(f'{self.code_replacement_policy} is not a valid dead code replacement strategy.')

    def forward(self, x):
        if not self._initted_seen:                            # one host read, then cached (initted only ever goes 0 -> 1)
            if not bool(self.initted.item()):
                raise NotImplementedError("k-means codebook initialisation (kmeans_init=True) is not on the B200 path; no shipped config uses it")
            self._initted_seen = True
        quantize, embed_ind = self.scan(x)                    # quantize = embed[idx] from the codebook BEFORE this step's update
        if self.training:
            flat = x.detach().reshape(-1, x.shape[-1]).float().contiguous()
            self._ema_update(flat, embed_ind)
            self.expire_codes_(x.detach())
        return quantize, embed_ind


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
        loss = torch.zeros(1, device=x.device, requires_grad=self.training)   # a fill kernel: torch.tensor([0.], device=...) is a blocking pageable H2D copy
        if self.training:
            quantize = z + (quantize - z).detach()                               # straight-through (reference :532)
            if self.commitment_weight > 0:
                loss = loss + F.mse_loss(quantize.detach(), z) * self.commitment_weight
            if self.orthogonal_reg_weight > 0:
                raise NotImplementedError("orthogonal codebook regularisation is not used by any shipped tokenizer config")
        if self.accept_image_fmap:
            quantize = quantize.reshape(B, Hq, Wq, C).permute(0, 3, 1, 2)
            embed_ind = embed_ind.reshape(B, Hq, Wq)
        elif not self.channel_last:
            quantize = quantize.transpose(1, 2)
        return quantize, loss, embed_ind
