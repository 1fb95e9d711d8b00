"""Golden fixture for the qk_norm presets (fm.py:1059-1130; NormAttention / NormCrossAttention, fm_utils.py:222-307): the
UNMODIFIED reference 4M-Tiny with qk_norm=True on the same deterministic weights / synthetic mod7 batch as make_golden.py.

Run in the authoring container only:   python tests/golden/make_golden_qknorm.py   -> tests/golden/fourm_tiny_qknorm_golden.pt
"""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402
import make_golden as MG  # noqa: E402
from oracle import fourm_oracle as O  # noqa: E402

CASES = {"fp32_128": (False, 128, 128, 0, 1234, 0), "bf16_128": (True, 128, 128, 0, 1234, 0)}
SLICE_KEYS = ["encoder.0.attn.q_norm.weight", "encoder.3.attn.k_norm.weight", "decoder.0.cross_attn.q_norm.weight",
              "decoder.5.cross_attn.k_norm.weight", "decoder.2.self_attn.q_norm.weight", "encoder.0.attn.qkv.weight",
              "decoder.0.cross_attn.kv.weight", "mask_token"]


def main():
    fm, fm_utils, MODALITY_INFO = ref_import.import_reference_models()
    torch.manual_seed(0)
    specs = O.mod7_specs()
    model = MG.build_reference_fourm("fm_tiny_6e_6d_swiglu_nobias", specs, MODALITY_INFO, qk_norm=True)
    assert type(model.encoder[0].attn).__name__ == "NormAttention"
    sd = MG.det_state_dict(model)
    model.load_state_dict(sd)
    gold = dict(meta=dict(torch=torch.__version__, reference_commit="cda590f"), model="fm_tiny_6e_6d_swiglu_nobias", qk_norm=True,
                weight_checksums={k: float(v.double().sum()) for k, v in sd.items()},
                shapes={k: tuple(v.shape) for k, v in sd.items()},
                param_names=[k for k, _ in model.named_parameters(remove_duplicate=False)], cases={})
    for tag, (amp, N, M, seed, bseed, extra) in CASES.items():
        b = O.synthetic_mod7_batch(2, seed=bseed, extra_valid=extra)
        loss, mod_loss, grads, logits, tl = MG.run_fourm_case(model, b, N, M, seed, amp)
        random.seed(seed)
        dec_names = [m for m in b if m in model.decoder_embeddings]
        order = random.sample(dec_names, len(dec_names))
        gold["cases"][tag] = dict(
            amp=amp, N=N, M=M, py_seed=seed, batch_seed=bseed, extra_valid=extra, decoder_order=order,
            loss=loss, mod_loss=mod_loss, token_loss=tl, grads=MG.grad_summary(grads, SLICE_KEYS),
            logits_slices={m: v[:, :4, :32].float().clone() for m, v in logits.items()},
            logits_norm={m: v.float().norm().item() for m, v in logits.items()})
        print(tag, float(loss), {k: round(float(v), 5) for k, v in mod_loss.items()})
    torch.save(gold, os.path.join(HERE, "fourm_tiny_qknorm_golden.pt"))


if __name__ == "__main__":
    main()
