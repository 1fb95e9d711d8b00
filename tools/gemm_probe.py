"""Where does the tcgen05 GEMM lose time at the 4M-B shapes?  Times each shape with the normal epilogue, with an epilogue that reads the
accumulator but stores nothing (option gemm_debug = 1) and with the epilogue skipped entirely (= 2): the difference separates the cost of
the mainloop (TMA + MMA) from TMEM read-out and from the staged global stores; = 3 keeps the shared-memory staging but drops the global
stores, = 4 stores straight from registers (bf16 epilogues only).  Rotating operand sets (> L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import lib, ops

M = 16384
SHAPES = [("qkv NT bf16", ops.LAYOUT_NT, ops.EPI_BF16, M, 2304, 768), ("proj NT bf16", ops.LAYOUT_NT, ops.EPI_BF16, M, 768, 768),
          ("fc13 NT swiglu", ops.LAYOUT_NT, ops.EPI_SWIGLU, M, 2048, 768), ("fc2 NT bf16", ops.LAYOUT_NT, ops.EPI_BF16, M, 768, 2048),
          ("dgrad NN bf16", ops.LAYOUT_NN, ops.EPI_BF16, M, 768, 2304), ("wgrad TN f32", ops.LAYOUT_TN, ops.EPI_F32, 2304, 768, M),
          ("big NT bf16", ops.LAYOUT_NT, ops.EPI_BF16, 8192, 8192, 8192)]
S = 4
torch.manual_seed(0)
for name, layout, epi, m, n, k in SHAPES:
    if layout == ops.LAYOUT_NT:
        As = [torch.randn(m, k, device="cuda").bfloat16() for _ in range(S)]
        Bs = [torch.randn((2 * n if epi == ops.EPI_SWIGLU else n), k, device="cuda").bfloat16() for _ in range(S)]
    elif layout == ops.LAYOUT_NN:
        As = [torch.randn(m, k, device="cuda").bfloat16() for _ in range(S)]
        Bs = [torch.randn(k, n, device="cuda").bfloat16() for _ in range(S)]
    else:
        As = [torch.randn(k, m, device="cuda").bfloat16() for _ in range(S)]
        Bs = [torch.randn(k, n, device="cuda").bfloat16() for _ in range(S)]
    flops = 2.0 * m * k * (2 * n if epi == ops.EPI_SWIGLU else n)
    outs = None
    res = []
    for dbg in (0, 1, 2, 3, 4):
        lib.set_option("gemm_debug", dbg)
        r = ops.gemm(As[0], Bs[0], layout=layout, epilogue=epi)
        o0, o1 = (r if isinstance(r, tuple) else (r, None))
        for i in range(4):
            ops.gemm(As[i % S], Bs[i % S], layout=layout, epilogue=epi, out=o0, out1=o1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(40):
            ops.gemm(As[i % S], Bs[i % S], layout=layout, epilogue=epi, out=o0, out1=o1)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 40 * 1e3)
    lib.set_option("gemm_debug", 0)
    print(f"{name:16s} {m:6d}x{n:5d}x{k:5d}  normal {res[0]:7.1f} us ({flops / res[0] / 1e6:7.1f} TF/s)   no-store {res[1]:7.1f} us ({flops / res[1] / 1e6:7.1f})   "
          f"no-epilogue {res[2]:7.1f} us ({flops / res[2] / 1e6:7.1f})   staging-only {res[3]:7.1f} us   direct-stg {res[4]:7.1f} us")
