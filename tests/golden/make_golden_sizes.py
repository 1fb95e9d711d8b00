"""Golden fixtures at the BENCHMARKED sizes, from the UNMODIFIED reference (apple/ml-4m @ /root/reference).

    python tests/golden/make_golden_sizes.py        (authoring container only; ~10 min of CPU)

Writes tests/golden/fourm_{base,large}_golden.pt:
  * 4M-B mod7 (`fm_base_12e_12d_swiglu_nobias`), B=2, 128+128 tokens   -- BASELINE.json configs[1] per-sample shape
  * 4M-L mod7 (`fm_large_24e_24d_swiglu_nobias`), B=1, 256+256 tokens  -- BASELINE.json configs[2] per-sample shape
Each is run three times through the reference's own `FourM.forward` + backward (fm.py:640-691): fp32, bf16-autocast
(what `run_training_4m.py --dtype bfloat16` does) and fp64 (`model.double()`, the yardstick).  The fixture stores the
loss, per-modality losses, logits slices and the norm of EVERY parameter gradient for all three, so the GPU tests can set
their tolerance to a multiple of the reference's own bf16-vs-fp64 error instead of a guess.
Weights: `oracle.fourm_oracle.deterministic_tensor` (regenerated in the tests, checksummed here).
"""
import os
import random
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402
from make_golden import build_reference_fourm, clone_batch, det_state_dict  # noqa: E402
from oracle import fourm_oracle as O  # noqa: E402

CASES = {   # file tag: (model, B, tokens per side, (n_in_img, n_in_seq, n_tgt_img, n_tgt_seq), python seed, batch seed)
    "base": ("fm_base_12e_12d_swiglu_nobias", 2, 128, (18, 10, 22, 10), 0, 1234),
    "large": ("fm_large_24e_24d_swiglu_nobias", 1, 256, (36, 20, 44, 19), 0, 4321),
}
SLICE_KEYS = ["mask_token", "encoder.0.attn.qkv.weight", "encoder.{last}.mlp.fc2.weight", "decoder.0.cross_attn.kv.weight",
              "decoder.{last}.mlp.fc1.weight", "decoder_proj_context.weight", "encoder_embeddings.rgb@224.proj.weight",
              "encoder_norm.weight", "decoder.3.query_norm.weight", "decoder_embeddings.tok_rgb@224.token_emb.weight",
              "encoder_embeddings.caption.token_emb.weight"]


def run(model, batch, N, seed, mode):
    dt = dict(fp32=torch.float32, bf16=torch.float32, fp64=torch.float64)[mode]
    model = model.to(dt)
    b = clone_batch(batch)
    if mode == "fp64":
        b["rgb@224"]["tensor"] = b["rgb@224"]["tensor"].double()
    model.zero_grad(set_to_none=True)
    random.seed(seed)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        loss, mod_loss = model(b, num_encoder_tokens=N, num_decoder_tokens=N, loss_type="mod")
    loss.backward()
    grads = {k: p.grad.detach().double() for k, p in model.named_parameters() if p.grad is not None}
    random.seed(seed)
    b = clone_batch(batch)
    if mode == "fp64":
        b["rgb@224"]["tensor"] = b["rgb@224"]["tensor"].double()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        logits = model(b, num_encoder_tokens=N, num_decoder_tokens=N, return_logits=True)
    return dict(loss=float(loss), mod_loss={k: float(v) for k, v in mod_loss.items()},
                grad_norm={k: float(g.norm()) for k, g in grads.items()},
                logits_slices={m: v[:, :4, :32].double().clone() for m, v in logits.items()},
                logits_norm={m: float(v.double().norm()) for m, v in logits.items()}), grads


def main():
    fm, fm_utils, MODALITY_INFO = ref_import.import_reference_models()
    specs = O.mod7_specs()
    for tag, (name, B, N, budgets, seed, bseed) in CASES.items():
        t0 = time.time()
        torch.manual_seed(0)
        model = build_reference_fourm(name, specs, MODALITY_INFO)
        sd = det_state_dict(model)
        model.load_state_dict(sd)
        batch = O.synthetic_mod7_batch(B, *budgets, seed=bseed)
        gold = dict(meta=dict(torch=torch.__version__, reference_commit="cda590f"), model=name, B=B, N=N, budgets=budgets,
                    py_seed=seed, batch_seed=bseed,
                    weight_checksums={k: float(v.double().sum()) for k, v in sd.items()},
                    shapes={k: tuple(v.shape) for k, v in sd.items()},
                    param_names=[k for k, _ in model.named_parameters(remove_duplicate=False)], runs={})
        random.seed(seed)
        dec_names = [m for m in batch if m in model.decoder_embeddings]
        gold["decoder_order"] = random.sample(dec_names, len(dec_names))
        last = len(model.encoder) - 1
        keys = [k.format(last=last) for k in SLICE_KEYS]
        for mode in ("fp32", "bf16", "fp64"):
            res, grads = run(model, batch, N, seed, mode)
            if mode in ("fp64", "bf16"):
                res["grad_slices"] = {k: grads[k].flatten()[:4096].clone() for k in keys if k in grads}
            gold["runs"][mode] = res
            print(tag, mode, res["loss"], f"{time.time() - t0:.0f}s", flush=True)
        r = gold["runs"]
        print(tag, "bf16-vs-fp64 loss err", abs(r["bf16"]["loss"] - r["fp64"]["loss"]), "fp32-vs-fp64", abs(r["fp32"]["loss"] - r["fp64"]["loss"]))
        rel = sorted(abs(r["bf16"]["grad_norm"][k] - v) / max(v, 1e-30) for k, v in r["fp64"]["grad_norm"].items())
        print(tag, "bf16 grad-norm rel err: median", rel[len(rel) // 2], "p90", rel[int(len(rel) * .9)], "max", rel[-1])
        path = os.path.join(HERE, f"fourm_{tag}_golden.pt")
        torch.save(gold, path)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
