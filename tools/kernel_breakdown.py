"""GPU time per kernel family for one steady-state training step (CUPTI via torch.profiler; complements the ncu launch list)."""
import collections, os, random, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
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


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type.name != "CUDA":
        continue
    name = re.sub(r"\(.*", "", e.name)[:70]
    if "gemm_kernel" in e.name:
        name = "gemm (all)"
    agg[name][0] += 1
    agg[name][1] += e.device_time / 1e3 if hasattr(e, "device_time") else e.cuda_time / 1e3
tot = sum(v[1] for v in agg.values())
print(f"GPU busy per step: {tot / 2:.2f} ms")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{ms / 2:8.3f} ms {100 * ms / tot:5.1f}%  n={n // 2:4d}  {k}")
