"""Golden trajectory of the UNMODIFIED reference's GenerationSampler (fourm/models/generate.py) on 4M-Tiny, CPU, fp32.

    python tests/golden/make_golden_gen.py       -> tests/golden/gen_tiny_golden.pt

RGB -> depth (MaskGIT) -> normals (ROAR) -> caption (autoregressive), all with classifier-free guidance (tests/helpers.py
`generation_case`).  Stored: the target modalities' (tensor, input_mask, target_mask) after EVERY schedule step, and for every arg-max
decision (temperature-0 steps) the gap between the best and the second-best logit, so the GPU test can tell a genuine near-tie from
an error.  Random numbers come from the CPU generator (torch.manual_seed(seed + step) inside the reference's step functions)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402
from make_golden import build_reference_fourm, det_state_dict  # noqa: E402
from oracle import fourm_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

AR_MAX_TOKENS = 40


def main():
    fm, fm_utils, MODALITY_INFO = ref_import.import_reference_models()
    from fourm.models.generate import GenerationSampler
    specs = O.mod7_specs()
    torch.manual_seed(0)
    model = build_reference_fourm("fm_tiny_6e_6d_swiglu_nobias", specs, MODALITY_INFO).eval()
    model.load_state_dict(det_state_dict(model))
    model.modality_info = {k: dict(v) for k, v in model.modality_info.items()}
    model.modality_info['caption']['max_tokens'] = AR_MAX_TOKENS          # keep the O(L^2) reference loop short
    sampler = GenerationSampler(model)
    tok = H.StubTextTokenizer()
    sample, schedule = H.generation_case(model.modality_info)
    seed = 0
    gaps, picks, states = [], [], []
    real_argmax = torch.argmax

    def spy(x, *a, **k):
        if x.dim() >= 2 and x.shape[-1] > 1000:
            top2 = torch.topk(x.float(), 2, dim=-1)[0]
            gaps[-1].append((top2[..., 0] - top2[..., 1]).reshape(-1).clone())
            picks[-1].append(real_argmax(x, dim=-1).reshape(-1).clone())
        return real_argmax(x, *a, **k)

    state = sample
    for step, info in enumerate(schedule):
        gaps.append([])
        picks.append([])
        torch.argmax = spy
        try:
            state = sampler.generate(state, [info], top_k=0.0, top_p=0.8, text_tokenizer=tok, seed=seed + step)
        finally:
            torch.argmax = real_argmax
        states.append({m: {k: v.clone() for k, v in state[m].items() if k in ('tensor', 'input_mask', 'target_mask')} for m in H.GEN_TARGETS})
        print(step, info['target_domain'], info['scheme'], info['num_tokens'], float(info['temperature']), "argmax calls", len(gaps[-1]))
    gold = dict(meta=dict(torch=torch.__version__, reference_commit="cda590f"), model="fm_tiny_6e_6d_swiglu_nobias", seed=seed, top_p=0.8,
                ar_max_tokens=AR_MAX_TOKENS, schedule=[{k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in s.items()} for s in schedule],
                states=states, gaps=gaps, picks=picks)
    path = os.path.join(HERE, "gen_tiny_golden.pt")
    torch.save(gold, path)
    print(path, os.path.getsize(path) // 1024, "KiB", "caption:", states[-1]['caption']['tensor'][0, :12].tolist())


if __name__ == "__main__":
    main()
