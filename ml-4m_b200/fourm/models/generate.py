"""Generation with a trained 4M model: chained MaskGIT / ROAR / autoregressive decoding with classifier-free guidance, behind the
reference's `fourm.models.generate` surface (module functions, `build_chained_generation_schedules`, `GenerationSampler` and its
step methods, `generate` / `generate_iter` / `generate_multi_guided` / `generate_sam_dense`; reference fourm/models/generate.py).

What is different inside (the results are the reference's, see tests/test_gpu_generate.py):
  * guided steps run the conditional and the unconditional pass as ONE batch through the encoder and the decoder (the contexts are
    padded to a common length with masked slots) -- the reference runs two full passes (generate.py:745-764, 936-962);
  * the per-step `copy.deepcopy(mod_dict)` of guided steps (generate.py:672, 791, 946) is replaced by a shallow copy that clones
    only the tensors the `empty_*_modality` helpers overwrite;
  * the autoregressive loop keeps per-layer K/V caches and projects the encoder context through the cross-attention `kv` layers once
    per call (`b200fm.decode.CachedDecoder`); the reference re-runs the whole decoder over the whole prefix per token (:886-913);
  * top-k / top-p filtering + softmax run as torch ops on the device without the reference's sort -> gather round trip through the
    full vocabulary where a cheaper equivalent exists; the random draws themselves are the SAME torch calls in the same order
    (`torch.manual_seed(seed)`, `torch.rand` for the ROAR order, `torch.multinomial`), so a seed reproduces the reference's stream.
`GenerationSampler.rng_device = "cpu"` draws those numbers on the host generator instead (what the reference does when it runs on
CPU): used by the parity tests against the CPU goldens.
"""
from typing import List, Optional, Union

import numpy as np
import os

import torch
import torch.nn.functional as F
from torch import nn

try:      # the reference's helpers when its tree is importable, else the local restatements (same semantics)
    from fourm.utils import get_sentinel_to_id_mapping, merge_span_masking
    from fourm.utils.generation import continue_schedule, cosine_schedule, linear_schedule, linear_temp_schedule, onex_temp_schedule
except Exception:      # noqa: BLE001
    from b200fm.genutils import (continue_schedule, cosine_schedule, get_sentinel_to_id_mapping, linear_schedule, linear_temp_schedule,
                                 merge_span_masking, onex_temp_schedule)

try:
    from tqdm import tqdm
except Exception:      # noqa: BLE001
    def tqdm(it, disable=True):
        return it

_SEQ = ('seq', 'seq_token')


# ---------------------------------------------------------------------------------------------------------------------
# mod_dict construction helpers (reference generate.py:30-206)
# ---------------------------------------------------------------------------------------------------------------------
def empty_img_modality(mod_dict, key):
    """Nothing of `key` is an input, everything is a target (:30-37)."""
    mod_dict[key]['input_mask'][:] = True
    mod_dict[key]['target_mask'][:] = False
    return mod_dict


def empty_seq_modality(mod_dict, key, s1_id=5):
    """Empty sequence = input [S_1], target [S_1] ... [S_2] (:39-63)."""
    d = mod_dict[key]
    d['tensor'][:] = 0
    d['tensor'][:, [0, 1]] = s1_id
    d['tensor'][:, -1] = s1_id + 1
    d['input_mask'][:] = True
    d['input_mask'][:, 0] = False
    d['target_mask'] = ~d['input_mask']
    d['decoder_attention_mask'][:] = 1
    d['decoder_attention_mask'][:, 0] = 0
    return mod_dict


def empty_seq_emb_modality(mod_dict, key):
    """Empty pre-computed-embedding sequence: one (zero) input position so CFG has something to attend to (:65-81)."""
    d = mod_dict[key]
    d['tensor'] = torch.zeros_like(d['tensor'])
    d['input_mask'] = torch.ones_like(d['input_mask'])
    d['input_mask'][:, 0] = False
    d['target_mask'] = torch.ones_like(d['target_mask'])
    d['decoder_attention_mask'][:] = False
    return mod_dict


def init_empty_target_modality(mod_dict, modality_info, domain, batch_size, num_tokens, device):
    """Placeholder entry for a modality that is about to be generated (:84-117)."""
    kind = modality_info[domain]['type']
    if kind == 'img':
        mod_dict[domain] = dict(tensor=torch.zeros((batch_size, num_tokens), dtype=torch.int64, device=device),
                                input_mask=torch.ones((batch_size, num_tokens), dtype=torch.bool, device=device),
                                target_mask=torch.zeros((batch_size, num_tokens), dtype=torch.bool, device=device))
        return empty_img_modality(mod_dict, domain)
    if kind in ('seq', 'seq_token', 'seq_emb'):
        n = max(num_tokens, 2)
        mod_dict[domain] = dict(tensor=torch.zeros((batch_size, n), dtype=torch.int32, device=device),
                                input_mask=torch.ones((batch_size, n), dtype=torch.bool, device=device),
                                target_mask=torch.zeros((batch_size, n), dtype=torch.bool, device=device),
                                decoder_attention_mask=torch.zeros((batch_size, n), dtype=torch.bool, device=device))
        return empty_seq_emb_modality(mod_dict, domain) if kind == 'seq_emb' else empty_seq_modality(mod_dict, domain)
    raise ValueError()


def init_full_input_modality(mod_dict, modality_info, domain, device, eos_id=3):
    """Mark a given modality as a complete input (:119-156)."""
    d = mod_dict[domain]
    if domain.startswith('rgb'):
        b, _, H, W = d['tensor'].shape
        p = modality_info[domain]['patch_size']
        shape = (b, (H // p) * (W // p))
    else:
        shape = d['tensor'].shape
    d.setdefault('input_mask', torch.zeros(shape, dtype=torch.bool, device=device))
    d.setdefault('target_mask', torch.ones(shape, dtype=torch.bool, device=device))
    d.setdefault('decoder_attention_mask', torch.zeros(shape, dtype=torch.bool, device=device))
    kind = modality_info[domain]['type']
    if kind == 'img':
        d['input_mask'][:] = False
        d['target_mask'][:] = True
    elif kind in _SEQ:
        if eos_id in d['tensor']:
            eos_idx = torch.where(d['tensor'] == eos_id)[1][0].item()
        else:
            d['tensor'][:, 0] = eos_id
            eos_idx = 0
        d['input_mask'][:, :eos_idx + 1] = False
        d['input_mask'][:, eos_idx + 1:] = True
        d['target_mask'][:] = True
    elif kind == 'seq_emb':
        d['input_mask'] = ~d['mask_valid']
        d['target_mask'] = torch.ones_like(d['mask_valid'])
        d['decoder_attention_mask'] = torch.zeros_like(d['mask_valid'])
    return mod_dict


def custom_text(sample, input_text, eos_token, key, device, text_tokenizer, target_max_len=50, start_token="[S_1]"):
    """Text prompt as input ids + a padded target span [S_1] [PAD]... eos (:158-190)."""
    inp = torch.tensor(text_tokenizer.encode(input_text).ids).unsqueeze(0)
    tgt_text = " ".join([start_token] + ["[PAD]"] * (target_max_len - 2) + [eos_token])
    tgt = torch.tensor(text_tokenizer.encode(tgt_text).ids).unsqueeze(0)
    ids = torch.cat([inp, tgt], dim=1)
    is_tgt = torch.cat([torch.zeros_like(inp, dtype=torch.bool), torch.ones_like(tgt, dtype=torch.bool)], dim=1)
    sample[key] = dict(tensor=ids.to(device), input_mask=is_tgt.to(device), target_mask=(~is_tgt).to(device),
                       decoder_attention_mask=torch.zeros(ids.shape, dtype=torch.bool, device=device))
    return sample


def expand_to_batch(mod_dict, batch_size):
    """Broadcast batch-1 entries to batch_size (:192-203)."""
    for mod, d in mod_dict.items():
        for k, v in d.items():
            if k in ('tensor', 'input_mask', 'target_mask', 'decoder_attention_mask', 'mask_valid'):
                if v.shape[0] == 1:
                    d[k] = v.expand(batch_size, *v.shape[1:])
                elif v.shape[0] != batch_size:
                    raise ValueError(f"Invalid batch size: {v.shape[0]} instead of {batch_size}")
    return mod_dict


def build_chained_generation_schedules(cond_domains: List[str], target_domains: List[str], tokens_per_target: List[int],
                                       autoregression_schemes: List[str], decoding_steps: List[int], token_decoding_schedules: List[str],
                                       temps: List[float], temp_schedules: List[float], cfg_scales: List[float], cfg_schedules: List[str],
                                       cfg_grow_conditioning: bool = False, modality_info: Optional[dict] = None):
    """List of per-step dicts {target_domain, scheme, num_tokens, temperature, cfg_scale, cfg_cond_domains} (:208-318)."""
    steps, cond = [], list(cond_domains)
    for i, target in enumerate(target_domains):
        scheme, ntoks, temp, cfg_scale = autoregression_schemes[i], tokens_per_target[i], temps[i], cfg_scales[i]
        if scheme == 'autoregressive':
            steps.append(dict(target_domain=target, scheme=scheme, num_tokens=None, temperature=temp, cfg_scale=cfg_scale,
                              cfg_cond_domains=list(cond)))
            continue
        if modality_info is not None:
            assert modality_info[target]['type'] not in _SEQ, f'Illegal autoregressive scheme {scheme} for target domain {target}'
        n_steps = decoding_steps[i]
        if scheme == 'maskgit':
            kind = token_decoding_schedules[i]
            if kind == 'cosine':
                tok_sched = cosine_schedule(n_steps, ntoks)
            elif kind == 'linear':
                tok_sched = linear_schedule(n_steps, ntoks)
            else:
                raise ValueError(f'Illegal MaskGIT token schedule {kind}')
        elif scheme == 'roar':
            tok_sched = linear_schedule(n_steps, ntoks)
        else:
            raise ValueError(f'Illegal decoding scheme {scheme}')
        ts = temp_schedules[i]
        if ts == 'linear':
            temp_sched = linear_temp_schedule(temp, tok_sched)
        elif ts == 'constant':
            temp_sched = temp * np.ones(n_steps)
        elif 'onex' in ts:
            min_t, power = [float(f) for f in ts.split(':')[1:]]
            temp_sched = onex_temp_schedule(max_t=temp, min_t=min_t, token_schedule=tok_sched, power=power)
        else:
            raise ValueError(f'Illegal temperature schedule {ts}')
        if cfg_schedules[i] == 'constant':
            if isinstance(cfg_scale, float):
                cfg_sched = cfg_scale * np.ones(n_steps)
            elif isinstance(cfg_scale, list):
                cfg_sched = np.array(cfg_scale) * np.ones(n_steps).reshape(-1, 1)
        elif cfg_schedules[i] == 'cosine':
            raise NotImplementedError()
        else:
            raise ValueError(f'Illegal guidance schedule {cfg_schedules[i]}')
        steps += [dict(target_domain=target, scheme=scheme, num_tokens=t, temperature=tt, cfg_scale=c, cfg_cond_domains=list(cond))
                  for t, tt, c in zip(tok_sched, temp_sched, cfg_sched)]
        if cfg_grow_conditioning:
            cond.append(target)
    return steps


def _clone_dict(mod_dict, deep_keys=()):
    """Copy of the two-level dict; tensors of the modalities in `deep_keys` are cloned (they will be overwritten in place), all other
    tensors are shared.  Replaces copy.deepcopy(mod_dict) of the guided steps."""
    return {m: {k: (v.clone() if (m in deep_keys and torch.is_tensor(v)) else v) for k, v in d.items()} for m, d in mod_dict.items()}


def _deep_clone(mod_dict):
    return {m: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()} for m, d in mod_dict.items()}


class GenerationSampler(nn.Module):
    """Wraps a trained 4M model for generation (reference generate.py:321-1272)."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.rng_device = None       # None: random numbers on the tensors' device (the reference's behaviour); "cpu": host generator
        self.batch_cfg = True        # conditional + unconditional pass as one batch
        self.kv_cache = True         # K/V-cached autoregressive loop (falls back when the model has non-plain modules)
        # The reference tests "every sequence holds an EOS" on the host after EVERY token (generate.py:1011-1013), which drains the GPU
        # queue once per token.  Here the per-token flags stay on the device and are read every `eos_check_every` tokens; the sequence
        # is then cut at the first token where the condition held, so the returned tokens are the reference's.  (The extra tokens drawn
        # meanwhile advance the random stream; with rng_device = "cpu" -- the parity mode -- the check is per token like the reference's.)
        self.eos_check_every = 8
        # AR sampling at temperature > 0 with top_k = 0: one fused kernel (b200fm_sample_top_p: nucleus cut by bisection + inverse-CDF draw
        # from a torch.rand number) instead of sort / cumsum / scatter / 2 softmaxes / multinomial per token.  Same distribution, a different
        # use of the random stream; rng_device = "cpu" (the parity mode) keeps the reference's torch calls.
        self.fused_sampling = os.environ.get("B200FM_FUSED_SAMPLING", "1") != "0"

    # ------------------------------------------------------------------ sampling
    def top_k_top_p_filtering(self, logits, top_k=0.0, top_p=0.0):
        """In-place -inf masking of everything outside the top-k / nucleus (:332-359).  [rows, V]."""
        if top_k > 0.0:
            if isinstance(top_k, int):
                k = min(top_k, logits.shape[-1])
            elif isinstance(top_k, float):
                k = min(int(top_k * logits.shape[-1]), logits.shape[-1])
            else:
                raise ValueError(f"Invalid value for top_k: {top_k}")
            kth = torch.topk(logits, k)[0][..., -1, None]
            logits.masked_fill_(logits < kth, float("-inf"))
        if top_p > 0.0:
            # keep the smallest prefix of the descending-sorted distribution whose mass reaches top_p (first token above the
            # threshold included): per row, that is every logit >= the logit at the cut position.
            srt, idx = torch.sort(logits, dim=1, descending=True)
            cum = torch.cumsum(F.softmax(srt, dim=-1), dim=-1)
            drop_sorted = cum > top_p
            drop_sorted[..., 1:] = drop_sorted[..., :-1].clone()
            drop_sorted[..., 0] = False
            drop = torch.zeros_like(drop_sorted).scatter_(1, idx, drop_sorted)
            logits.masked_fill_(drop, float("-inf"))
        return logits

    def _multinomial(self, probs):
        if self.rng_device == "cpu" and probs.is_cuda:
            return torch.multinomial(probs.float().cpu(), 1)[:, 0].to(probs.device)
        return torch.multinomial(probs, 1)[:, 0]

    def sample_tokens(self, logits, temperature=1.0, top_k=0.0, top_p=0.0):
        """(:361-371) -> (samples [rows], probability of each sample)."""
        if np.isclose(temperature, 0, atol=1e-10):
            samples = torch.argmax(logits, dim=-1)
            return samples, torch.ones_like(samples, dtype=torch.float32)
        probs = F.softmax(self.top_k_top_p_filtering(logits, top_k, top_p) / temperature, dim=-1)
        samples = self._multinomial(probs)
        return samples, probs.gather(1, samples[:, None])[:, 0]

    def sample_tokens_batched(self, logits, temperature=1.0, top_k=0.0, top_p=0.0):
        if logits.ndim > 2:
            B, N = logits.shape[:2]
            s, p = self.sample_tokens(logits.reshape(B * N, -1), temperature, top_k, top_p)
            return s.reshape(B, N), p.reshape(B, N)
        return self.sample_tokens(logits, temperature, top_k, top_p)

    def select_tokens(self, logits, num_select, temperature=1.0, top_k=0.0, top_p=0.0, return_all_samples=False):
        samples, probs = self.sample_tokens(logits, temperature, top_k, top_p)
        top = torch.topk(probs, num_select)[1]
        return (samples[top], top, samples) if return_all_samples else (samples[top], top)

    def select_tokens_batched(self, logits, num_select, temperature=1.0, top_k=0.0, top_p=0.0, return_all_samples=False):
        if logits.ndim > 2:
            samples, probs = self.sample_tokens_batched(logits, temperature, top_k, top_p)
            top = torch.topk(probs, num_select, dim=-1)[1]
            picked = torch.gather(samples, -1, top)
            return (picked, top, samples) if return_all_samples else (picked, top)
        return self.select_tokens(logits, num_select, temperature, top_k, top_p, return_all_samples)

    # ------------------------------------------------------------------ token selection for the two stacks
    @staticmethod
    def _keep_first_valid(mask, k, jitter=None):
        """Indices of the first k unmasked positions: argsort(mask + tiny increasing / random term)[:, :k] (:430-434, 486-491)."""
        L = mask.shape[1]
        if jitter is None:
            jitter = torch.arange(L, device=mask.device).unsqueeze(0) * 1e-6
        return torch.argsort(mask + jitter, dim=1)[:, :k]

    def forward_mask_encoder_generation(self, encoder_mod_dict):
        """All visible encoder tokens, batch-padded to the largest count (:407-445)."""
        B = list(encoder_mod_dict.values())[0]['tensor'].shape[0]
        toks, embs, masks, mods = self.model.cat_encoder_tensors(encoder_mod_dict)
        n_keep = int((~masks.reshape(B, -1)).sum(dim=1).max())
        keep = self._keep_first_valid(masks, n_keep)
        gi = keep[..., None].expand(-1, -1, toks.shape[2])
        tok, emb = torch.gather(toks, 1, gi), torch.gather(embs, 1, gi)
        msk, mod = torch.gather(masks, 1, keep), torch.gather(mods, 1, keep)
        if self.model.num_register_tokens > 0:
            reg = self.prompt_tokens.expand(B, -1, -1)
            tok = torch.cat([reg, tok], dim=1)
            emb = torch.cat([torch.zeros_like(reg), emb], dim=1)
            msk = torch.cat([torch.zeros((B, reg.shape[1]), dtype=torch.bool, device=msk.device), msk], dim=1)
            mod = torch.cat([torch.full((B, reg.shape[1]), -1, dtype=torch.int16, device=mod.device), mod], dim=1)
        tok = tok.masked_fill(msk[..., None], 0.)
        emb = emb.masked_fill(msk[..., None], 0.)
        mod = mod.masked_fill(msk, -1)
        return tok, emb, msk[:, None, :], mod

    def _mask_decoder(self, mod_dict, target_mod, scheme, num_select=None, seed=None):
        """Decoder-side selection of the three schemes (:448-560): returns (x | ids, emb, mask, mod_mask, positions)."""
        if seed is not None:
            torch.manual_seed(seed)
        d = mod_dict[target_mod]
        B, L = d['target_mask'].shape[0], d['x'].shape[1]
        dev = d['x'].device
        n_valid = int((~d['target_mask'][0]).sum())
        jitter = None
        if scheme == 'roar':
            n_valid = min(num_select, n_valid)
            rdev = "cpu" if self.rng_device == "cpu" else dev
            jitter = (torch.rand(L, device=rdev).to(dev).unsqueeze(0) * 1e-6)
        keep = self._keep_first_valid(d['target_mask'], n_valid, jitter)
        gi = keep[..., None].expand(-1, -1, d['emb'].shape[2])
        emb = torch.gather(d['emb'], 1, gi)
        msk = torch.gather(d['target_mask'], 1, keep)
        mod = torch.full_like(d['ids'], self.model.modality_info[target_mod]['id'], dtype=torch.int16).gather(1, keep).masked_fill(msk, -1)
        pos = torch.arange(L, device=dev).unsqueeze(0).expand(B, -1).gather(1, keep)
        emb = emb.masked_fill(msk[..., None], 0.)
        if scheme == 'autoregressive':
            first = torch.gather(d['ids'], 1, keep).masked_fill(msk, 0)
        else:
            first = (torch.zeros_like(emb) + self.model.mask_token).masked_fill(msk[..., None], 0.)
        return first, emb, msk, mod, pos

    def forward_mask_decoder_maskgit(self, mod_dict, target_mod, seed=None):
        return self._mask_decoder(mod_dict, target_mod, 'maskgit', seed=seed)

    def forward_mask_decoder_roar(self, mod_dict, target_mod, num_select, seed=None):
        return self._mask_decoder(mod_dict, target_mod, 'roar', num_select, seed)

    def forward_mask_decoder_autoregressive(self, mod_dict, target_mod, seed=None):
        return self._mask_decoder(mod_dict, target_mod, 'autoregressive', seed=seed)

    # ------------------------------------------------------------------ sequence merging (host side, :562-640)
    def merge_sequences(self, mod_dict, pred_ids, target_mod, text_tokenizer, default_sentinel="[S_1]"):
        device = mod_dict[target_mod]['tensor'].device
        ids = mod_dict[target_mod]['tensor'].squeeze().detach().cpu()
        ids = ids[mod_dict[target_mod]['input_mask'].squeeze().detach().cpu() == 0].tolist()
        if len(ids) == 0:
            ids = [text_tokenizer.get_vocab()[default_sentinel]]
        pred = pred_ids.squeeze().detach().cpu().tolist()
        if isinstance(pred, int):
            pred = [pred]
        sentinels = set(get_sentinel_to_id_mapping(text_tokenizer).values())
        merged = torch.tensor(merge_span_masking(ids, pred, sentinels)).unsqueeze(0)
        mod_dict[target_mod] = dict(tensor=merged.to(device), input_mask=torch.zeros_like(merged, dtype=torch.bool).to(device),
                                    target_mask=torch.ones_like(merged, dtype=torch.bool).to(device),
                                    decoder_attention_mask=torch.zeros_like(merged, dtype=torch.bool).to(device))
        return mod_dict

    def merge_sequences_batched(self, mod_dict, pred_ids, target_mod, text_tokenizer, default_sentinel="[S_1]"):
        pad_id = text_tokenizer.token_to_id("[PAD]")
        device = mod_dict[target_mod]['tensor'].device
        rows = []
        for t, im, pi in zip(torch.split(mod_dict[target_mod]['tensor'], 1), torch.split(mod_dict[target_mod]['input_mask'], 1), torch.split(pred_ids, 1)):
            rows.append(self.merge_sequences({target_mod: dict(tensor=t, input_mask=im)}, pi, target_mod, text_tokenizer, default_sentinel)[target_mod])
        width = max(r['tensor'].shape[1] for r in rows)
        tens = torch.cat([F.pad(r['tensor'], (0, width - r['tensor'].shape[1]), "constant", pad_id) for r in rows], dim=0).to(device)
        # (the reference pads BOTH masks from the merged input_mask, :618-631)
        masks = torch.cat([F.pad(r['input_mask'], (0, width - r['input_mask'].shape[1]), "constant", True) for r in rows], dim=0).to(device)
        mod_dict[target_mod] = dict(tensor=tens, input_mask=masks, target_mask=masks.clone(),
                                    decoder_attention_mask=torch.zeros_like(masks, dtype=torch.bool))
        return mod_dict

    # ------------------------------------------------------------------ encoder / decoder passes
    def _encode(self, mod_dict):
        enc = {mod: self.model.encoder_embeddings[mod](dict(d)) for mod, d in mod_dict.items() if mod in self.model.encoder_embeddings}
        tok, emb, mask, _ = self.forward_mask_encoder_generation(enc)
        return tok + emb, emb, mask

    def _contexts(self, mod_dicts):
        """Encoder + context projection for several variants of the batch (conditional, unconditional, ...) in ONE pass: the variants
        are stacked along the batch axis, shorter ones padded with masked slots.  Returns ([k*B, N, D] context, [k*B, 1, N] mask)."""
        parts = [self._encode(md) for md in mod_dicts]
        if len(parts) > 1 and not self.batch_cfg:
            outs = [(self.model.decoder_proj_context(self.model.forward_encoder(x, m)) + e, m) for x, e, m in parts]
        else:
            n = max(p[0].shape[1] for p in parts)

            def pad(t, value):
                return t if t.shape[-2 if t.dim() == 3 and t.dtype != torch.bool else -1] == n else None
            xs, es, ms = [], [], []
            for x, e, m in parts:
                extra = n - x.shape[1]
                if extra:
                    x, e = F.pad(x, (0, 0, 0, extra)), F.pad(e, (0, 0, 0, extra))
                    m = F.pad(m, (0, extra), value=True)
                xs.append(x); es.append(e); ms.append(m)
            x, e, m = torch.cat(xs), torch.cat(es), torch.cat(ms)
            ctx = self.model.decoder_proj_context(self.model.forward_encoder(x, m)) + e
            return ctx, m
        n = max(c.shape[1] for c, _ in outs)
        ctx = torch.cat([F.pad(c, (0, 0, 0, n - c.shape[1])) for c, _ in outs])
        msk = torch.cat([F.pad(m, (0, n - m.shape[-1]), value=True) for _, m in outs])
        return ctx, msk

    def _img_logits(self, mod_dicts, target_mod, scheme, num_select=None, seed=None):
        """Logits of the decoder slots of `target_mod` for every variant in mod_dicts -> ([k, B, n, V], positions [B, n])."""
        ctx, enc_mask = self._contexts(mod_dicts)
        k = len(mod_dicts)
        dec = {target_mod: self.model.decoder_embeddings[target_mod].forward_embed(dict(mod_dicts[-1][target_mod]))}
        tok, emb, _, mod_mask, pos = self._mask_decoder(dec, target_mod, scheme, num_select, seed)
        y = (tok + emb).repeat(k, 1, 1)
        y = self.model.forward_decoder(y, ctx, enc_mask, None)
        kB, n, _ = y.shape
        logits = self.model.forward_logits(y, dec, mod_mask.repeat(k, 1))[target_mod].float()
        return logits.reshape(k, kB // k, n, -1), pos

    def forward_enc_dec_maskgit_batched(self, mod_dict, target_mod, seed=None):
        logits, pos = self._img_logits([mod_dict], target_mod, 'maskgit', seed=seed)
        return logits[0], pos

    def forward_enc_dec_roar_batched(self, mod_dict, target_mod, num_select, seed=None):
        logits, pos = self._img_logits([mod_dict], target_mod, 'roar', num_select, seed)
        return logits[0], pos

    # ------------------------------------------------------------------ guidance plumbing
    def _unconditional(self, mod_dict, conditioning):
        un = _clone_dict(mod_dict, deep_keys=set(conditioning))
        for mod in conditioning:
            kind = self.model.modality_info[mod]['type']
            if kind in _SEQ:
                empty_seq_modality(un, mod)
            elif kind == 'seq_emb':
                empty_seq_emb_modality(un, mod)
            else:
                empty_img_modality(un, mod)
        return un

    @staticmethod
    def _write(d, pos, tokens, as_input=True):
        d['tensor'] = torch.scatter(d['tensor'], -1, pos, tokens.to(d['tensor'].dtype))
        d['input_mask'] = torch.scatter(d['input_mask'], -1, pos, torch.full_like(tokens, not as_input, dtype=torch.bool))
        d['target_mask'] = torch.scatter(d['target_mask'], -1, pos, torch.full_like(tokens, as_input, dtype=torch.bool))

    # ------------------------------------------------------------------ MaskGIT (:642-743)
    def maskgit_step_batched(self, mod_dict, target_mod, num_select, temperature, top_k, top_p, seed=None):
        logits, pos = self.forward_enc_dec_maskgit_batched(mod_dict, target_mod, seed=seed)
        picked, top = self.select_tokens_batched(logits, num_select, temperature=temperature, top_k=top_k, top_p=top_p)
        self._write(mod_dict[target_mod], torch.gather(pos, -1, top), picked)
        return mod_dict

    def guided_maskgit_step_batched(self, mod_dict, target_mod, num_select, temperature, top_k, top_p, conditioning=[], guidance_scale=1.0,
                                    seed=None, write_all_predictions=False):
        logits, pos = self._img_logits([mod_dict, self._unconditional(mod_dict, conditioning)], target_mod, 'maskgit', seed=seed)
        guided = logits[1] + (logits[0] - logits[1]) * guidance_scale
        picked, top, samples = self.select_tokens_batched(guided, num_select, temperature=temperature, top_k=top_k, top_p=top_p,
                                                          return_all_samples=True)
        top_pos = torch.gather(pos, -1, top)
        if write_all_predictions:
            mod_dict[target_mod]['tensor'][:, pos] = samples
            d = mod_dict[target_mod]
            d['input_mask'] = torch.scatter(d['input_mask'], -1, top_pos, torch.zeros_like(picked, dtype=torch.bool))
            d['target_mask'] = torch.scatter(d['target_mask'], -1, top_pos, torch.ones_like(picked, dtype=torch.bool))
        else:
            self._write(mod_dict[target_mod], top_pos, picked)
        return mod_dict

    def multi_guided_maskgit_step_batched(self, uncond_dict, cond_dicts, cond_weights, target_mod, num_select, temperature, top_k, top_p,
                                          seed=None, write_all_predictions=False):
        logits, pos = self._img_logits(list(cond_dicts) + [uncond_dict], target_mod, 'maskgit', seed=seed)
        guided = logits[-1] + sum(w * (logits[i] - logits[-1]) for i, w in enumerate(cond_weights))
        picked, top, samples = self.select_tokens_batched(guided, num_select, temperature=temperature, top_k=top_k, top_p=top_p,
                                                          return_all_samples=True)
        top_pos = torch.gather(pos, -1, top)
        if write_all_predictions:
            uncond_dict[target_mod]['tensor'][:, pos] = samples
            d = uncond_dict[target_mod]
            d['input_mask'] = torch.scatter(d['input_mask'], -1, top_pos, torch.zeros_like(picked, dtype=torch.bool))
            d['target_mask'] = torch.scatter(d['target_mask'], -1, top_pos, torch.ones_like(picked, dtype=torch.bool))
        else:
            self._write(uncond_dict[target_mod], top_pos, picked)
        for cd in cond_dicts:
            self._write(cd[target_mod], top_pos, picked)
        return uncond_dict, cond_dicts

    # ------------------------------------------------------------------ ROAR (:745-848)
    def roar_step_batched(self, mod_dict, target_mod, num_select, temperature, top_k, top_p, seed=None):
        logits, pos = self.forward_enc_dec_roar_batched(mod_dict, target_mod, num_select, seed=seed)
        samples, _ = self.sample_tokens_batched(logits, temperature, top_k=top_k, top_p=top_p)
        self._write(mod_dict[target_mod], pos, samples)
        return mod_dict

    def guided_roar_step_batched(self, mod_dict, target_mod, num_select, temperature, top_k, top_p, conditioning=[], guidance_scale=1.0,
                                 seed=None):
        logits, pos = self._img_logits([mod_dict, self._unconditional(mod_dict, conditioning)], target_mod, 'roar', num_select, seed)
        guided = logits[1] + (logits[0] - logits[1]) * guidance_scale
        samples, _ = self.sample_tokens_batched(guided, temperature, top_k=top_k, top_p=top_p)
        self._write(mod_dict[target_mod], pos, samples)
        return mod_dict

    def multi_guided_roar_step_batched(self, uncond_dict, cond_dicts, cond_weights, target_mod, num_select, temperature, top_k, top_p,
                                       seed=None):
        logits, pos = self._img_logits(list(cond_dicts) + [uncond_dict], target_mod, 'roar', num_select, seed)
        guided = logits[-1] + sum(w * (logits[i] - logits[-1]) for i, w in enumerate(cond_weights))
        samples, _ = self.sample_tokens_batched(guided, temperature, top_k=top_k, top_p=top_p)
        self._write(uncond_dict[target_mod], pos, samples)
        for cd in cond_dicts:
            self._write(cd[target_mod], pos, samples, as_input=False)          # (reference :840-843 keeps them as targets here)
        return uncond_dict, cond_dicts

    # ------------------------------------------------------------------ autoregressive (:850-1025)
    def _autoregress(self, mod_dicts, weights, target_mod, temperature, top_k, top_p, use_eos, eos_token, start_tokens, seed):
        """mod_dicts = [conditional, (unconditional)]; logits = l_uncond + w * (l_cond - l_uncond) when guided."""
        model = self.model
        ctx, enc_mask = self._contexts(mod_dicts)
        k = len(mod_dicts)
        dec = {target_mod: model.decoder_embeddings[target_mod].forward_embed(dict(mod_dicts[-1][target_mod]))}
        ids, emb, dmask, mod_mask, _ = self._mask_decoder(dec, target_mod, 'autoregressive', seed=seed)
        B, dev = ids.shape[0], ids.device
        seq_len = model.modality_info[target_mod]['max_tokens']
        if use_eos and eos_token is None:
            eos_token = ids[0][dmask[0] == 0][-1]
        if use_eos:
            eos_token = eos_token.to(dev)
        out = ids[:, :1] if start_tokens is None else start_tokens.to(dev)
        if use_eos and (out == eos_token).any(dim=-1).all():
            return out, True
        y_emb = emb[:, :seq_len]
        seq_len = y_emb.shape[1]
        emb_mod = model.decoder_embeddings[target_mod]
        from b200fm import decode
        cached = self.kv_cache and decode.supported(model) and not model.training
        dec_state = None
        if cached:
            dec_state = decode.CachedDecoder(model, ctx, enc_mask, seq_len + out.shape[1])
            for t in range(out.shape[1] - 1):                      # prime the cache with all but the last given token
                dec_state.step((emb_mod.token_emb(out[:, t]) + y_emb[:, t]).repeat(k, 1))
        check_every = 1 if self.rng_device == "cpu" else max(1, int(self.eos_check_every))
        flags, flag_len = [], []
        for step_i in range(seq_len):
            cur = out.shape[1]
            if cached:
                h = dec_state.step((emb_mod.token_emb(out[:, -1]) + y_emb[:, cur - 1]).repeat(k, 1))
                logits = emb_mod.forward_logits(h).float().reshape(k, B, -1)
            else:
                y = (emb_mod.token_emb(out) + y_emb[:, :cur]).repeat(k, 1, 1)
                causal = torch.ones((cur, cur), dtype=torch.bool, device=dev).triu(1).unsqueeze(0).expand(k * B, -1, -1)
                y = model.forward_decoder(y, ctx, enc_mask, causal)
                logits = model.forward_logits(y, dec, mod_mask[:, :cur].repeat(k, 1))[target_mod].float().reshape(k, B, cur, -1)[:, :, -1]
            last = logits[0] if k == 1 else logits[-1] + sum(w * (logits[i] - logits[-1]) for i, w in enumerate(weights))
            if np.isclose(temperature, 0, atol=1e-10):
                nxt = torch.argmax(last, dim=-1, keepdim=True)
            elif self.fused_sampling and last.is_cuda and self.rng_device != "cpu" and not top_k > 0.0 and last.shape[-1] <= 51200:
                from b200fm import ops
                nxt = ops.sample_top_p(last.float().contiguous(), top_p, temperature, torch.rand(last.shape[0], device=last.device))[:, None]
            else:
                probs = F.softmax(self.top_k_top_p_filtering(last, top_k, top_p) / temperature, dim=-1)
                nxt = self._multinomial(probs)[:, None]
            out = torch.cat((out, nxt.to(out.dtype)), dim=-1)
            if use_eos:
                flags.append((out == eos_token).any(dim=-1).all())
                flag_len.append(out.shape[1])
                if len(flags) >= check_every or step_i == seq_len - 1:
                    hit = torch.stack(flags).nonzero()             # one host read for the whole window
                    if hit.numel() > 0:
                        out = out[:, :flag_len[int(hit[0])]]       # the length at which the reference's per-token test fires
                        break
                    flags, flag_len = [], []
        return out, False

    def autoregressive_step_batched(self, mod_dict, target_mod, temperature, top_k: Union[float, int], top_p: float, use_eos=True,
                                    eos_token=None, start_tokens=None, text_tokenizer=None, seed=None):
        out, early = self._autoregress([mod_dict], None, target_mod, temperature, top_k, top_p, use_eos, eos_token, start_tokens, seed)
        if early:
            return out
        return self.merge_sequences_batched(mod_dict, out, target_mod, text_tokenizer)

    def guided_autoregressive_step_batched(self, mod_dict, target_mod, temperature, top_k: Union[float, int], top_p: float, use_eos=True,
                                           eos_token=None, start_tokens=None, text_tokenizer=None, conditioning=[], guidance_scale=1.0,
                                           seed=None):
        pair = [mod_dict, self._unconditional(mod_dict, conditioning)]
        out, early = self._autoregress(pair, [guidance_scale], target_mod, temperature, top_k, top_p, use_eos, eos_token, start_tokens, seed)
        if early:
            return out
        return self.merge_sequences_batched(mod_dict, out, target_mod, text_tokenizer)

    # ------------------------------------------------------------------ drivers (:1028-1272)
    def _one_step(self, mod_dict, info, step, top_k, top_p, text_tokenizer, seed, write_all):
        target, temp = info['target_domain'], info['temperature']
        scale, cond = info.get('cfg_scale', 1.0), info.get('cfg_cond_domains', [])
        seed_i = seed + step if seed is not None else None
        guided = not (scale == 1.0 or len(cond) == 0)
        kind = self.model.modality_info[target]['type']
        if kind == 'img':
            scheme, n = info['scheme'].lower(), info['num_tokens']
            if scheme == 'maskgit':
                if guided:
                    kw = dict(write_all_predictions=True) if write_all else {}
                    return self.guided_maskgit_step_batched(mod_dict, target, n, temperature=temp, top_k=top_k, top_p=top_p, conditioning=cond,
                                                            guidance_scale=scale, seed=seed_i, **kw)
                return self.maskgit_step_batched(mod_dict, target, n, temperature=temp, top_k=top_k, top_p=top_p, seed=seed_i)
            if scheme == 'roar':
                if guided:
                    return self.guided_roar_step_batched(mod_dict, target, n, temperature=temp, top_k=top_k, top_p=top_p, conditioning=cond,
                                                         guidance_scale=scale, seed=seed_i)
                return self.roar_step_batched(mod_dict, target, n, temperature=temp, top_k=top_k, top_p=top_p, seed=seed_i)
            raise ValueError("Invalid sampling scheme")
        if kind in _SEQ:
            if guided:
                return self.guided_autoregressive_step_batched(mod_dict, target, temperature=temp, top_k=top_k, top_p=top_p,
                                                               text_tokenizer=text_tokenizer, conditioning=cond, guidance_scale=scale, seed=seed_i)
            return self.autoregressive_step_batched(mod_dict, target, temperature=temp, top_k=top_k, top_p=top_p,
                                                    text_tokenizer=text_tokenizer, seed=seed_i)
        raise ValueError("Invalid schedule")

    @torch.no_grad()
    def generate(self, mod_dict, schedule, top_k=0.0, top_p=0.0, text_tokenizer=None, verbose=False, seed=None):
        """Run the whole schedule; returns the completed mod_dict (the input dict is not modified)."""
        mod_dict = _deep_clone(mod_dict)
        for step, info in tqdm(enumerate(schedule), disable=not verbose):
            mod_dict = self._one_step(mod_dict, info, step, top_k, top_p, text_tokenizer, seed, write_all=False)
        return mod_dict

    @torch.no_grad()
    def generate_iter(self, mod_dict, schedule, top_k=0.0, top_p=0.0, text_tokenizer=None, verbose=False, seed=None):
        """As generate, yielding the mod_dict after every step (guided MaskGIT steps also write their provisional predictions)."""
        mod_dict = _deep_clone(mod_dict)
        for step, info in tqdm(enumerate(schedule), disable=not verbose):
            mod_dict = self._one_step(mod_dict, info, step, top_k, top_p, text_tokenizer, seed, write_all=True)
            yield mod_dict

    @torch.no_grad()
    def generate_multi_guided(self, uncond_dict, cond_dicts, schedule, top_k=0.0, top_p=0.0, text_tokenizer=None, verbose=False, seed=None):
        """Several weighted conditions (conjunction of guidance terms); image modalities only (:1161-1219)."""
        cur = schedule[0]['target_domain']
        uncond_dict = _deep_clone(uncond_dict)
        cond_dicts = [_deep_clone(c) for c in cond_dicts]
        for c in cond_dicts:
            c[cur] = _deep_clone({cur: uncond_dict[cur]})[cur]
        for step, info in tqdm(enumerate(schedule), disable=not verbose):
            target, temp, n, weights = info['target_domain'], info['temperature'], info['num_tokens'], info['cfg_scale']
            if cur != target:                                     # the previous modality is complete: it becomes one more condition
                for c in cond_dicts:
                    del c[cur]
                    c[target] = _deep_clone({target: uncond_dict[target]})[target]
                uncond_dict[cur]['input_mask'][:] = True
                new = {cur: _deep_clone({cur: uncond_dict[cur]})[cur], target: _deep_clone({target: uncond_dict[target]})[target]}
                new[cur]['input_mask'][:] = False
                new[cur]['target_mask'][:] = True
                cond_dicts.append(new)
                cur = target
            if self.model.modality_info[target]['type'] != 'img':
                raise NotImplementedError("Only image modalities are supported for now")
            scheme = info['scheme'].lower()
            if scheme == 'maskgit':
                uncond_dict, cond_dicts = self.multi_guided_maskgit_step_batched(uncond_dict, cond_dicts, weights, target, n, temp, top_k, top_p, seed=seed)
            elif scheme == 'roar':
                uncond_dict, cond_dicts = self.multi_guided_roar_step_batched(uncond_dict, cond_dicts, weights, target, n, temp, top_k, top_p, seed=seed)
            else:
                raise ValueError("Invalid sampling scheme")
        return uncond_dict

    @torch.no_grad()
    def generate_sam_dense(self, mod_dict, schedule, text_tokenizer, batch_size=16, key='sam_instance', top_k=0.0, top_p=0.0, seed=None,
                           verbose=False):
        """Dense SAM instances: the same prompt batch_size times, generated sequences merged into one (:1221-1272)."""
        device = mod_dict[list(mod_dict.keys())[0]]['tensor'].device
        mod_dict = _deep_clone(mod_dict)
        batch = expand_to_batch(_deep_clone(mod_dict), batch_size=batch_size)
        out = self.generate(batch, [s for s in schedule if s['target_domain'] == key], text_tokenizer=text_tokenizer, verbose=verbose, seed=seed,
                            top_p=top_p, top_k=top_k)
        sentinels = set(get_sentinel_to_id_mapping(text_tokenizer).values())
        merged = []
        for i in range(batch_size):
            row = out[key]['tensor'][i]
            merged.extend(merge_span_masking(row[out[key]['input_mask'][i] == 0].tolist(), row[out[key]['target_mask'][i] == 0].tolist(),
                                             sentinel_ids=sentinels))
        merged = torch.tensor(merged, device=device).unsqueeze(0)
        mod_dict[key] = dict(tensor=merged, input_mask=torch.zeros(merged.shape, dtype=torch.bool, device=device),
                             target_mask=torch.ones(merged.shape, dtype=torch.bool, device=device),
                             decoder_attention_mask=torch.zeros(merged.shape, dtype=torch.bool, device=device))
        return mod_dict
