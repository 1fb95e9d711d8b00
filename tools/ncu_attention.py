"""Stand-alone attention launches at 4M-B shapes with the three mask forms a train step uses, for `ncu --set full` captures:
   0 encoder self-attention  [B,1,N] padding mask (all False for the synthetic batch)
   1 decoder self-attention  [B,M,M] per-modality block mask (the dense mask adapt_decoder_attention_mask builds)
   2 decoder cross-attention [B,1,N]
each forward + backward.  ncu ... -k regex:attention python tools/ncu_attention.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from b200fm import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, H, N, D = 128, 12, 128, 768
R = B * N


def rnd(*s):
    return (torch.randn(*s, device=dev, generator=g) * 0.5).to(torch.bfloat16)


qkv, do = rnd(R, 3 * D), rnd(R, D)
pad = torch.zeros(B, 1, N, dtype=torch.bool, device=dev)
# decoder block mask: 5 image modalities x 22 tokens + 2 sequences x 9 tokens, tokens attend inside their modality; sequences causal
mod = torch.cat([torch.full((22,), i) for i in range(5)] + [torch.full((9,), 5 + i) for i in range(2)]).to(dev)
blk = (mod[:, None] != mod[None, :])
caus = torch.ones(N, N, dtype=torch.bool, device=dev).triu(1)
blk = blk | ((mod[:, None] >= 5) & caus)
dmask = blk[None].expand(B, N, N).contiguous()
for rep in range(int(os.environ.get("REPS", "2"))):
    for mask in (pad, dmask, pad):
        o, st = ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, N, N, mask)
        ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, st, B, H, N, N, mask)
torch.cuda.synchronize()
# plain CUDA-event timing of each variant (not under ncu)
if not os.environ.get("NCU"):
    for name, mask in (("enc self [B,1,N]", pad), ("dec self [B,M,M]", dmask)):
        for fn_name in ("fwd", "bwd"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o, st = ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, N, N, mask)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20):
                if fn_name == "fwd":
                    ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, N, N, mask)
                else:
                    ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, st, B, H, N, N, mask)
            e1.record(); torch.cuda.synchronize()
            print(f"{name} {fn_name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
print("ok")
