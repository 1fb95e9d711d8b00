"""LayerNorm forward / backward variants stand-alone at the 4M-B shape (16384 x 768 fp32 rows), rotating over 6 buffer sets (1.2 GB > L2):
time per launch and algorithmic GB/s (fwd: read 4D, write 2D; bwd: read 4D + 2D + 4D, write 4D + 2D per row) for every value of
both values of the option "ln_bwd_v2".  (Round 2 also measured a two-rows-in-flight backward and
shared-memory-gamma variants of both kernels: no gain, removed; numbers in profiles/r2_ln_bench.txt.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import lib, ops

rows, D, S = int(os.environ.get("ROWS", "16384")), int(os.environ.get("D", "768")), 6
torch.manual_seed(0)
xs = [torch.randn(rows, D, device="cuda") for _ in range(S)]
dys = [torch.randn(rows, D, device="cuda").bfloat16() for _ in range(S)]
drs = [torch.randn(rows, D, device="cuda") for _ in range(S)]
w = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
_, mean, rstd = ops.layernorm_fwd(xs[0], w, b, 1e-6)
dgamma = torch.zeros(D, device="cuda"); dbeta = torch.zeros(D, device="cuda")


def timeit(fn, n=60):
    for i in range(6):
        fn(i % S)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(n):
        fn(i % S)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


us = timeit(lambda i: ops.layernorm_fwd(xs[i], w, b, 1e-6))
print(f"fwd: {us:7.2f} us  {rows * D * 6 / us / 1e3:7.1f} GB/s")
for v in (0, 1):
    lib.set_option("ln_bwd_v2", v)
    us = timeit(lambda i: ops.layernorm_bwd(dys[i], xs[i], w, mean, rstd, dres=drs[i], want_bf16=True, dgamma=dgamma, dbeta=dbeta))
    print(f"bwd ln_bwd_v2={v}: {us:7.2f} us  {rows * D * 16 / us / 1e3:7.1f} GB/s")
lib.set_option("ln_bwd_v2", 1)
