"""Host-side cost per C-ABI call (issue only, GPU kept busy but never waited on inside the loop)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import ops, lib

dev = "cuda"
a = torch.randn(256, 256, device=dev).to(torch.bfloat16)
b = torch.randn(256, 256, device=dev).to(torch.bfloat16)
out = torch.empty(256, 256, device=dev, dtype=torch.bfloat16)
x = torch.randn(256, 768, device=dev)
g = torch.ones(768, device=dev)


def bench(name, fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t) / n * 1e6
    torch.cuda.synchronize()
    print(f"{name:40s} {dt:7.2f} us/call")


bench("ops.gemm (alloc out)", lambda: ops.gemm(a, b))
bench("ops.gemm (out given)", lambda: ops.gemm(a, b, out=out))
args = (0, 0, 256, 256, 256, a.data_ptr(), 256, b.data_ptr(), 256, out.data_ptr(), 256, 0, 0, 0, 0, 0, 1.0, 0, ops._stream())
bench("lib.call gemm (prebuilt args)", lambda: lib.call("b200fm_gemm_bf16", *args))
bench("ops.layernorm_fwd", lambda: ops.layernorm_fwd(x, g, None, 1e-6))
bench("torch.empty bf16", lambda: torch.empty(256, 256, device=dev, dtype=torch.bfloat16))
bench("torch add (eager kernel)", lambda: torch.add(x, x))
bench("ops._stream()", lambda: ops._stream())
