"""Small host-side helpers of the generation path: decoding schedules (fourm/utils/generation.py:49-104) and the span-masking
sentinel bookkeeping (fourm/utils/tokenizer/text_tokenizer.py:108-136).  Pure numpy / Python; restated here so that
`fourm.models.generate` works on a box without the reference tree."""
import math
from collections import defaultdict

import numpy as np


def cosine_schedule(num_steps, total_tokens):
    """Tokens decoded per step following a half cosine from 1 to 0 (generation.py:49-58); the last step takes the remainder."""
    levels = [0.5 * (1.0 + math.cos(math.pi * i / num_steps)) for i in range(num_steps)]
    per_step = [round(total_tokens * (a - b)) for a, b in zip(levels[:-1], levels[1:])]
    per_step.append(total_tokens - sum(per_step))
    return np.array(per_step)


def linear_schedule(num_steps, total_tokens):
    """Equal shares, larger ones first, zero-sized steps dropped (generation.py:61-66)."""
    edges = np.linspace(0, total_tokens, num_steps + 1, dtype=int)
    shares = np.sort(np.diff(edges))[::-1]
    return np.trim_zeros(shares, 'b')


def continue_schedule(schedule, num_current_tokens):
    """Remainder of a schedule after num_current_tokens have been decoded (generation.py:69-75)."""
    done = np.cumsum(schedule)
    keep = done > num_current_tokens
    rest = schedule[keep]
    rest[0] = done[keep][0] - num_current_tokens
    return rest


def onex_temp_schedule(max_t, min_t, token_schedule, power=0.5, min_linspace=1, max_linspace=100):
    """1 / x^power temperature decay over the decoded fraction (generation.py:84-96)."""
    x = np.linspace(min_linspace, max_linspace, num=sum(token_schedule))
    y = 1.0 / (x ** power)
    y = y - min(y)
    y = y / max(y)
    frac = np.cumsum(token_schedule) / np.sum(token_schedule)
    shaped = [(1.0 - f) * v for v, f in zip(y, frac)]
    return np.array([min_t + (max_t - min_t) * s for s in shaped]).clip(min=1e-9)


def linear_temp_schedule(temp, token_schedule):
    """Temperature proportional to the share of tokens still to decode (generation.py:99-101)."""
    total = token_schedule.sum()
    tail = (temp * (total - token_schedule.cumsum()) / total)[:-1]
    return np.concatenate([np.array([temp * 1.0]), tail]).clip(min=1e-9)


def get_sentinel_to_id_mapping(tokenizer, match_str="[S_"):
    """{k: token id of "[S_k]"} ordered by token id (text_tokenizer.py:108-112)."""
    hits = sorted(((v, k) for k, v in tokenizer.get_vocab().items() if k.startswith(match_str)))
    return {int(name.split("_")[1][:-1]): tid for tid, name in hits}


def merge_span_masking(input_seq, decoder_seq, sentinel_ids):
    """Replace every sentinel of the input by the tokens the decoder produced after that sentinel (text_tokenizer.py:115-136)."""
    spans = defaultdict(list)
    current = None
    for tok in decoder_seq:
        if tok in sentinel_ids:
            current = tok
        else:
            spans[current].append(tok)
    merged = []
    for tok in input_seq:
        if tok in sentinel_ids:
            merged.extend(spans[tok])
        else:
            merged.append(tok)
    return merged
