"""Attention forward / backward stand-alone at the 4M-B shape (B=128 samples x 12 heads x 128 tokens, dense mask), rotating over 6
operand sets (> L2): microseconds per launch and the HBM roofline fraction (fwd: q,k,v read + o written = 64 KB per (b,h); bwd: q,k,v,o,dO
read + dq,dk,dv written = 128 KB) for the backward with 8 and with 16 math warps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import lib, ops

B, H, N, S = int(os.environ.get("B", "128")), 12, int(os.environ.get("N", "128")), 6
torch.manual_seed(0)
sets = []
for _ in range(S):
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    sets.append((qkv, torch.randn(B * N, H * 64, device="cuda").bfloat16()))
mask = (torch.rand(B, N, N, device="cuda") < 0.1)
C = H * 64


def t(fn, n=60):
    for i in range(6): fn(i % S)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(n): fn(i % S)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


outs = [ops.attention_fwd(q[:, :C], q[:, C:2 * C], q[:, 2 * C:], B, H, N, N, mask) for q, _ in sets]
us = t(lambda i: ops.attention_fwd(sets[i][0][:, :C], sets[i][0][:, C:2 * C], sets[i][0][:, 2 * C:], B, H, N, N, mask))
print(f"fwd: {us:7.2f} us  {B * H * N * 64 * 2 * 4 / us / 1e3:7.1f} GB/s")
for mw in (8, 16):
    lib.set_option("attn_bwd_warps", mw)
    us = t(lambda i: ops.attention_bwd(sets[i][0][:, :C], sets[i][0][:, C:2 * C], sets[i][0][:, 2 * C:], outs[i][0], sets[i][1], outs[i][1], B, H, N, N, mask))
    print(f"bwd math_warps={mw:2d}: {us:7.2f} us  {B * H * N * 64 * 2 * 8 / us / 1e3:7.1f} GB/s")
lib.set_option("attn_bwd_warps", 8)
