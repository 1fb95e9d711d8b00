"""CPU oracle for on-GPU batch assembly -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/ may import it).

Restates `UnifiedMasking.image_mask` (fourm/data/masking.py:236-266) as a function of explicit noise, and the RGB loader's
normalisation (fourm/data/modality_transforms.py:206-219: ToTensor then Normalize).  Pinned against the unmodified reference by
tests/golden/make_golden_masking.py -> tests/golden/masking_golden.pt."""
import numpy as np


def image_mask(noise, input_budget, target_budget=None):
    """noise fp32 [L] -> (input_mask bool [L], target_mask bool [L], decoder_attention_mask int32 [L]); True = masked."""
    L = noise.shape[0]
    pi = np.argsort(noise, kind="stable")
    base_in = np.ones(L, dtype=bool)
    base_in[:input_budget] = False
    input_mask = base_in[pi]
    if target_budget is None:
        target_mask = ~input_mask
    else:
        base_t = np.ones(L, dtype=bool)
        base_t[input_budget:input_budget + target_budget] = False
        target_mask = base_t[pi]
    dam = np.zeros(L, dtype=np.int32)
    first = int(np.argmin(target_mask.astype(np.float64) + np.arange(L) * 1e-6))
    dam[first] = int((~target_mask).sum())
    return input_mask, target_mask, dam


def normalise_rgb_u8(img_u8, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """uint8 [B,3,H,W] -> fp32, ToTensor (x / 255) then Normalize ((x - mean) / std), fp32 arithmetic in that order."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[None, :, None, None]
    s = np.asarray(std, dtype=np.float32)[None, :, None, None]
    return (x - m) / s
