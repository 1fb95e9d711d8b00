"""Host-side (Python) profile of one training step's issue path: cProfile over a few steps, top functions by own time."""
import cProfile
import os
import pstats
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm.compat import build_mod7_embeddings, create_model
from b200fm.optim import FusedAdamW, param_groups_like_reference
from b200fm.synthetic import mod7_batch

dev = torch.device("cuda")
torch.manual_seed(0)
enc, dec, info = build_mod7_embeddings()
model = create_model("fm_base_12e_12d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).to(dev)
opt = FusedAdamW(param_groups_like_reference(model, 0.05), lr=1e-4, betas=(0.9, 0.95))
batch = {m: {k: v.to(dev) for k, v in d.items()} for m, d in mod7_batch(128).items()}
random.seed(0)


def step():
    loss, _ = model(batch, num_encoder_tokens=128, num_decoder_tokens=128)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
