"""Batch assembly on the GPU for the 4M train step.

The reference masks every sample on CPU dataloader workers (`UnifiedMasking`, fourm/data/masking.py:131-564) and ships, per
modality, the token tensor plus three mask tensors -- and the RGB image as normalised fp32 (77 of the 80 MB of a mod-7 batch of 128).
Here the image-like modalities (the bulk of the tokens) are masked on the device from per-row budgets and uniform noise
(`b200fm_mask_images`, bit-exact with `UnifiedMasking.image_mask` for the same noise), and RGB can travel as uint8 with the loader's
normalisation folded into the patchify kernel (`ImageEncoderEmbedding` accepts uint8 tensors).  Sequence modalities (captions, boxes:
span masking inserts sentinel tokens, i.e. rewrites short strings) stay on the host; their tensors are a few hundred bytes per sample.
"""
import torch

from . import ops


def sample_budgets(dirichlet_alphas, num_tokens, max_tokens, generator=None, device="cuda"):
    """Per-sample token budgets [B, n_mod] the way `UnifiedMasking.input_token_budget` draws them (masking.py:183-206), batched on
    the device: floor(Dirichlet(alpha) * num_tokens), the remainder handed out by arg-max of further Dirichlet draws, clamped to
    max_tokens.  dirichlet_alphas: [B, n_mod] (one row per sample: the mixture component is chosen by the caller)."""
    alphas = torch.as_tensor(dirichlet_alphas, dtype=torch.float32, device=device).clamp(min=1e-9)
    B, n = alphas.shape

    def draw(shape_prefix):
        g = torch._standard_gamma(alphas.expand(*shape_prefix, B, n).contiguous(), generator=generator)
        return g / g.sum(-1, keepdim=True)
    budget = (draw(()) * num_tokens).floor().to(torch.int32)
    diff = num_tokens - budget.sum(-1)                                   # < n per sample
    extra = draw((n,)).argmax(-1)                                        # [n, B]: candidate recipients of the left-over tokens
    take = torch.arange(n, device=device)[:, None] < diff[None, :]
    budget = budget + torch.zeros_like(budget).scatter_add_(1, extra.t(), take.t().to(torch.int32))
    return torch.minimum(budget, torch.as_tensor(max_tokens, dtype=torch.int32, device=device)[None])


class DeviceImageMasking:
    """mod_dict entries ('tensor', 'input_mask', 'target_mask', 'decoder_attention_mask') of the image-like modalities, masks
    generated on the device.  `mods`: names in batch order; all must have the same number of positions L (196 for 224 / 16)."""

    def __init__(self, mods, num_positions):
        self.mods, self.L = list(mods), int(num_positions)

    def __call__(self, tensors, in_budget, tgt_budget=None, noise=None, generator=None):
        """tensors: {mod: device tensor [B, ...]}; in_budget / tgt_budget: int32 [B, n_mods] (device); noise: fp32 [B, n_mods, L]
        uniform(0, 1) (drawn here when None)."""
        B, n = in_budget.shape
        dev = in_budget.device
        if noise is None:
            noise = torch.rand(B, n, self.L, device=dev, generator=generator)
        im, tm, dam = ops.mask_images(noise.contiguous(), in_budget.contiguous(), None if tgt_budget is None else tgt_budget.contiguous())
        out = {}
        for i, m in enumerate(self.mods):
            out[m] = dict(tensor=tensors[m], input_mask=im[:, i], target_mask=tm[:, i], decoder_attention_mask=dam[:, i])
        return out
