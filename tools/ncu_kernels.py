"""Stand-alone launches of the hot kernels at 4M-B (cfg-2) shapes for `ncu --set full` captures (kept tiny on purpose: ncu
replays every kernel ~40 times).  Usage: ncu --set full ... python tools/ncu_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
R, D, H = 16384, 768, 2048


def rnd(*s):
    return (torch.randn(*s, device=dev, generator=g) * 0.5).to(torch.bfloat16)


x, wqkv, w13, w2 = rnd(R, D), rnd(3 * D, D), rnd(2 * H, D), rnd(D, H)
for _ in range(2):
    qkv = ops.gemm(x, wqkv)                                             # NT bf16   16384 x 2304 x 768
    ab, gate = ops.gemm(x, w13, epilogue=ops.EPI_SWIGLU)                # NT swiglu 16384 x 2048(x2) x 768
    y = ops.gemm(gate, w2)                                              # NT bf16   16384 x 768 x 2048
    dx = ops.gemm(ab, w13, layout=ops.LAYOUT_NN)                        # NN bf16   16384 x 768 x 4096
    dw = ops.gemm(ab, x, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32)    # TN f32    4096 x 768 x 16384 (split-K)
    B, Hh, N = 128, 12, 128
    mask = (torch.rand(B, 1, N, device=dev, generator=g) < 0.2)
    o, st = ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, Hh, N, N, mask)
    dq, dk, dv = ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, rnd(R, D), st, B, Hh, N, N, mask)
    xs = torch.randn(R, D, device=dev, generator=g)
    s, hh, mean, rstd = ops.add_layernorm_fwd(xs, y, torch.ones(D, device=dev), None, 1e-6)
    ops.layernorm_bwd(hh, s, torch.ones(D, device=dev), mean, rstd, dres=xs, want_bf16=True)
torch.cuda.synchronize()
print("ok")
