"""Weight-streaming rate of a chain of dependent small-M linears shaped like the 4M-XL decoder (24 layers x {qkv, proj, q, proj, fc1|fc3
SwiGLU, fc2} at M = 2 rows), replayed from a CUDA graph: GB/s of weight bytes per second, with the L2 prefetch ahead of the
programmatic-dependency wait on and off (option "gemv_prefetch"), and with the tcgen05 tile path (option "gemv" = 0) for reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import lib, ops

D, H, L, M = 2048, 5464, 24, int(os.environ.get("M", "2"))
torch.manual_seed(0)
dev = "cuda"
layers = []
for _ in range(L):
    layers.append(dict(qkv=(torch.randn(3 * D, D, device=dev) * 0.02).bfloat16(), proj=(torch.randn(D, D, device=dev) * 0.02).bfloat16(),
                       q=(torch.randn(D, D, device=dev) * 0.02).bfloat16(), proj2=(torch.randn(D, D, device=dev) * 0.02).bfloat16(),
                       fc13=(torch.randn(2 * H, D, device=dev) * 0.02).bfloat16(), fc2=(torch.randn(D, H, device=dev) * 0.02).bfloat16()))
wbytes = sum(w.numel() * 2 for l in layers for w in l.values())
x0 = torch.randn(M, D, device=dev).bfloat16()


def chain(x):
    for l in layers:
        y = ops.gemm(x, l["qkv"], epilogue=ops.EPI_BF16)
        x = ops.gemm(y[:, :D], l["proj"], epilogue=ops.EPI_BF16)
        x = ops.gemm(x, l["q"], epilogue=ops.EPI_BF16)
        x = ops.gemm(x, l["proj2"], epilogue=ops.EPI_BF16)
        _, g = ops.gemm(x, l["fc13"], epilogue=ops.EPI_SWIGLU)
        x = ops.gemm(g, l["fc2"], epilogue=ops.EPI_BF16)
    return x


def measure(tag):
    for _ in range(2):
        chain(x0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        out = chain(x0)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{tag:28s} {ms:7.3f} ms per chain of {6 * L} linears  {wbytes / ms / 1e6:8.1f} GB/s of weights ({wbytes / 1e9:.2f} GB)  checksum {float(out.float().abs().sum()):.4f}")


print("M =", M)
for tag, opts in (("gemv, prefetch on", dict(gemv=1, gemv_prefetch=1)), ("gemv, prefetch off", dict(gemv=1, gemv_prefetch=0)),
                  ("gemv, prefetch on, no PDL", dict(gemv=1, gemv_prefetch=1, pdl=0)), ("tcgen05 tiles", dict(gemv=0))):
    for k, v in opts.items():
        lib.set_option(k, v)
    measure(tag)
    lib.set_option("pdl", 1); lib.set_option("gemv", 1); lib.set_option("gemv_prefetch", 1)
