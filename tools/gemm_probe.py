"""Back-to-back timing of the 4M-B GEMM shapes in isolation (hot L2, boost clocks): the upper bound the in-step numbers of
profiles/r1_gemm_shapes_*.json are compared with."""
import os, sys, torch
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ml-4m_b200"))
from b200fm import ops
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
R = 16384
x = torch.randn(R, 768, device="cuda").bfloat16(); xk = torch.randn(R, 2048, device="cuda").bfloat16()
w_qkv = torch.randn(2304, 768, device="cuda").bfloat16(); w_p = torch.randn(768, 768, device="cuda").bfloat16()
w2 = torch.randn(768, 2048, device="cuda").bfloat16(); w13 = torch.randn(4096, 768, device="cuda").bfloat16()
dy = torch.randn(R, 2304, device="cuda").bfloat16()
for name, fn, fl in [("NT 16384x2304x768", lambda: ops.gemm(x, w_qkv), 2 * R * 2304 * 768),
                     ("NT 16384x768x768", lambda: ops.gemm(x, w_p), 2 * R * 768 * 768),
                     ("NT 16384x768x2048", lambda: ops.gemm(xk, w2), 2 * R * 768 * 2048),
                     ("NT swiglu 16384x2048x768", lambda: ops.gemm(x, w13, epilogue=ops.EPI_SWIGLU), 2 * R * 4096 * 768),
                     ("TN 2304x768x16384", lambda: ops.gemm(dy, x, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32), 2 * R * 2304 * 768),
                     ("TN 768x768x16384", lambda: ops.gemm(x, x, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32), 2 * R * 768 * 768),
                     ("TN 768x2048x16384", lambda: ops.gemm(x, xk, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32), 2 * R * 768 * 2048)]:
    us = t(fn)
    print(f"{name:28s} {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s")
