"""Host-side generation helpers (b200fm.genutils) and the local MODALITY_INFO restatement (b200fm.compat) against known answers of
the unmodified reference (tests/golden/make_golden_genutils.py).  CPU only."""
import json
import os

import numpy as np

from tests import helpers as H


def _gold():
    with open(os.path.join(H.GOLDEN, "genutils_golden.json")) as f:
        return json.load(f)


def test_schedules_match_reference():
    from b200fm import genutils as G
    for key, ref in _gold()["schedules"].items():
        steps, total = (int(x) for x in key.split("_"))
        ts = G.linear_schedule(steps, total)
        assert G.cosine_schedule(steps, total).tolist() == ref["cosine"], key
        assert ts.tolist() == ref["linear"], key
        assert np.allclose(G.linear_temp_schedule(1.5, ts), ref["linear_temp"], rtol=0, atol=1e-12), key
        assert np.allclose(G.onex_temp_schedule(2.0, 0.1, ts, power=0.7), ref["onex"], rtol=0, atol=1e-12), key
        assert G.continue_schedule(ts.copy(), min(17, total - 1)).tolist() == ref["cont"], key


def test_sentinel_merge_matches_reference():
    from b200fm import genutils as G
    sent = set(range(4, 104))
    for case in _gold()["merge"]:
        assert G.merge_span_masking(case["inp"], case["dec"], sent) == case["merged"]
    tok = H.StubTextTokenizer()
    m = G.get_sentinel_to_id_mapping(tok)
    assert m[1] == 5 and m[2] == 6 and len(m) == 100


def test_local_modality_info_matches_reference_fields():
    """Every modality the local table restates carries the reference's type / vocabulary / token budget / patch geometry / id and builds
    the same embedding class with the same keyword arguments."""
    from b200fm.compat import MOD21_IN, MOD21_OUT, local_modality_info
    ref = _gold()["modality_info"]
    local = local_modality_info()
    assert set(MOD21_IN) <= set(local) and set(MOD21_OUT) <= set(local)
    for name, d in local.items():
        r = ref[name]
        for k in ("type", "vocab_size", "max_tokens", "min_tokens", "patch_size", "input_size", "id"):
            if k in r:
                assert d.get(k) == r[k], (name, k, d.get(k), r[k])
        for side in ("encoder_embedding", "decoder_embedding"):
            f = d.get(side)
            if r[side] is None:
                assert f is None, (name, side)
            else:
                assert f.func.__name__ == r[side]["cls"], (name, side)
                kw = {k: v for k, v in f.keywords.items()}
                want = dict(r[side]["kw"])
                want.pop("sincos_pos_emb", None) if want.get("sincos_pos_emb") is True else None      # the default
                kw.pop("sincos_pos_emb", None) if kw.get("sincos_pos_emb") is True else None
                assert kw == want, (name, side, kw, want)


def test_sampling_rule_matches_reference():
    """`top_k_top_p_filtering` of the overlay (torch, the path `rng_device="cpu"` and the fallback use) and the oracle's sort-free statement
    of the nucleus rule (the one csrc/sampling.cu evaluates by bisection) keep exactly the tokens the UNMODIFIED reference keeps
    (tests/golden/sampling_golden.pt: 9 cases incl. top-k as int / fraction, top-k + top-p, top_p = 1)."""
    import os
    import sys
    import numpy as np
    import torch
    from oracle import sampling_oracle as S
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden_sampling as G
    from fourm.models.generate import GenerationSampler
    gold = torch.load(os.path.join(here, "golden", "sampling_golden.pt"), weights_only=False)
    assert len(gold["cases"]) == len(G.CASES)
    for c in gold["cases"]:
        logits = G.case_logits(c["rows"], c["V"], c["scale"], c["seed"])
        ours = torch.isfinite(GenerationSampler.top_k_top_p_filtering(None, logits.clone(), c["top_k"], c["top_p"]))
        assert torch.equal(ours, c["kept"]), (c["V"], c["top_k"], c["top_p"])
        x = logits.numpy()
        keep = S.top_k_keep(x, c["top_k"])
        keep &= S.nucleus_keep(np.where(keep, x, -np.inf), c["top_p"])
        diff = int((keep != c["kept"].numpy()).sum())
        # the sort-free statement may differ from the sorted cumulative sum only by float32 summation order at the cut: at most one token a row
        assert diff <= c["rows"] and (diff == 0 or c["V"] >= 8192), (c["V"], c["top_k"], c["top_p"], diff)
