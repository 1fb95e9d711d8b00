"""Host-bound step time: the same train step at a tiny batch (GPU work negligible, launch count identical), wall-clocked, then
cProfile'd for the split between the autograd Functions, the ctypes calls and torch's own ops."""
import cProfile, os, pstats, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import lib
from b200fm.compat import build_mod7_embeddings, create_model
from b200fm.optim import FusedAdamW, param_groups_like_reference
from b200fm.synthetic import mod7_batch

dev = torch.device("cuda")
torch.manual_seed(0)
enc, dec, info = build_mod7_embeddings()
model = create_model("fm_base_12e_12d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).to(dev)
opt = FusedAdamW(param_groups_like_reference(model, 0.05), lr=1e-4, betas=(0.9, 0.95))
B = int(os.environ.get("HB", "2"))
batch = {m: {k: v.to(dev) for k, v in d.items()} for m, d in mod7_batch(B).items()}
random.seed(0)


def step():
    loss, _ = model(batch, num_encoder_tokens=128, num_decoder_tokens=128)
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1e9)
    opt.step()
    opt.zero_grad(set_to_none=True)


def phases():
    t0 = time.perf_counter(); loss, _ = model(batch, num_encoder_tokens=128, num_decoder_tokens=128)
    t1 = time.perf_counter(); loss.backward()
    t2 = time.perf_counter(); torch.nn.utils.clip_grad_norm_(model.parameters(), 1e9)
    t3 = time.perf_counter(); opt.step(); opt.zero_grad(set_to_none=True)
    t4 = time.perf_counter()
    return [t1 - t0, t2 - t1, t3 - t2, t4 - t3]


for _ in range(5):
    step()
torch.cuda.synchronize()
c0 = lib.CALLS["n"]
t = time.perf_counter()
N = 10
for _ in range(N):
    step()
torch.cuda.synchronize()
print(f"B={B}: wall {1e3 * (time.perf_counter() - t) / N:.2f} ms/step, {(lib.CALLS['n'] - c0) / N:.0f} C-ABI calls/step")
acc = [0.0] * 4
for _ in range(N):
    for i, v in enumerate(phases()):
        acc[i] += v
torch.cuda.synchronize()
print("host ms: fwd %.2f  bwd %.2f  clip %.2f  opt %.2f" % tuple(1e3 * a / N for a in acc))
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
