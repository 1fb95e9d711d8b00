"""Goldens for on-GPU masking from the UNMODIFIED reference: `UnifiedMasking.image_mask` (fourm/data/masking.py:236-266) with
torch.rand replaced by recorded noise, and torchvision-free restatement check of the RGB normalisation order.
    python tests/golden/make_golden_masking.py -> tests/golden/masking_golden.pt"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_import  # noqa: E402


def main():
    ref_import.import_reference_models()
    from fourm.data.masking import UnifiedMasking
    g = torch.Generator().manual_seed(0)
    cases = []
    real_rand = torch.rand
    for L, n_in, n_tgt in ((196, 18, 22), (196, 0, 40), (196, 196, 0), (196, 30, None), (256, 100, 200), (196, 0, 0), (16, 5, 6), (196, 1, 195)):
        for rep in range(3):
            noise = real_rand(L, generator=g)
            torch.rand = lambda *a, **k: noise.clone()
            try:
                out = UnifiedMasking.image_mask(None, torch.zeros(L), L, n_in, n_tgt)
            finally:
                torch.rand = real_rand
            cases.append(dict(L=L, n_in=n_in, n_tgt=n_tgt, noise=noise, input_mask=out["input_mask"].clone(),
                              target_mask=out["target_mask"].clone(), dam=out["decoder_attention_mask"].clone()))
    path = os.path.join(HERE, "masking_golden.pt")
    torch.save(dict(meta=dict(torch=torch.__version__, reference_commit="cda590f"), cases=cases), path)
    print(path, os.path.getsize(path) // 1024, "KiB", len(cases), "cases")


if __name__ == "__main__":
    main()
