"""Known answers from the UNMODIFIED reference for the token-sampling rule of the generation path:
`GenerationSampler.top_k_top_p_filtering` (fourm/models/generate.py:332-359) on seeded logit rows for several (top_k, top_p).
-> tests/golden/sampling_golden.pt (the kept-token masks; the logits are regenerated from the stored seeds).

    python tests/golden/make_golden_sampling.py          (authoring container only)"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_import  # noqa: E402

CASES = [  # (rows, V, logit scale, top_k, top_p, seed)
    (6, 1000, 2.5, 0.0, 0.8, 11), (4, 30000, 2.5, 0.0, 0.8, 12), (5, 133, 1.0, 0.0, 0.5, 13), (3, 8192, 4.0, 0.0, 0.95, 14),
    (4, 1000, 2.5, 50, 0.0, 15), (4, 1000, 2.5, 0.1, 0.0, 16), (4, 1000, 2.5, 50, 0.8, 17), (3, 500, 0.2, 0.0, 0.3, 18), (2, 64, 3.0, 0.0, 1.0, 19),
]


def case_logits(rows, V, scale, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows, V, generator=g) * scale


def main():
    ref_import.import_reference_models()
    import fourm.models.generate as gen
    assert gen.__file__.startswith(ref_import.REFERENCE_ROOT), gen.__file__
    flt = gen.GenerationSampler.top_k_top_p_filtering
    out = []
    for rows, V, scale, top_k, top_p, seed in CASES:
        kept = torch.isfinite(flt(None, case_logits(rows, V, scale, seed).clone(), top_k, top_p))
        out.append(dict(rows=rows, V=V, scale=scale, top_k=top_k, top_p=top_p, seed=seed, kept=kept, n_kept=kept.sum(-1).tolist()))
        print(rows, V, top_k, top_p, "kept per row", out[-1]["n_kept"])
    path = os.path.join(HERE, "sampling_golden.pt")
    torch.save(dict(meta=dict(torch=str(torch.__version__), reference_commit="cda590f"), cases=out), path)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
