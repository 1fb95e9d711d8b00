"""A/B of runtime options inside ONE process on ONE box (pool boxes differ by ~3 % in clocks): alternates the option value between
blocks of timed train steps and reports the per-value mean step time.

    python tools/ab_options.py ln_bwd_v2 0 1            # option name, then the values to compare
    python tools/ab_options.py pdl 1 0 --steps 10 --rounds 4
"""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import lib
from b200fm.compat import build_mod7_embeddings, create_model
from b200fm.optim import FusedAdamW, param_groups_like_reference
from b200fm.synthetic import budgets_for, mod7_batch

ap = argparse.ArgumentParser()
ap.add_argument("option")
ap.add_argument("values", nargs="+", type=int)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--model", default="fm_base_12e_12d_swiglu_nobias")
ap.add_argument("--batch", type=int, default=128)
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc, dec, info = build_mod7_embeddings()
model = create_model(args.model, encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).to(dev)
opt = FusedAdamW(param_groups_like_reference(model, 0.05), lr=1e-4, betas=(0.9, 0.95))
a, b, c, d = budgets_for(128)
batches = [{m: {k: v.to(dev) for k, v in dd.items()} for m, dd in mod7_batch(args.batch, a, b, c, d, seed=s).items()} for s in (1, 2)]
random.seed(0)


def step(i):
    loss, _ = model(batches[i % 2], num_encoder_tokens=128, num_decoder_tokens=128)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1e9)
    opt.step()
    opt.zero_grad(set_to_none=True)


for i in range(4):
    step(i)
times = {v: [] for v in args.values}
for r in range(args.rounds):
    for v in args.values:
        lib.set_option(args.option, v)
        step(0)                                             # one untimed step under the new setting
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        times[v].append(e0.elapsed_time(e1) / args.steps)
for v, ts in times.items():
    print(f"{args.option}={v}: {sum(ts) / len(ts):.3f} ms/step  (blocks: {', '.join(f'{t:.3f}' for t in ts)})")
