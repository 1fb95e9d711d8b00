"""GPU time per kernel of one VQ.tokenize call (ViT-B, 256x256, K=16384, batch 64): fp32-faithful (default call, no autocast) and bf16
(B200FM_VQ_PRECISION=bf16).  CUPTI via torch.profiler."""
import collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
import fourm.vq as vq

torch.manual_seed(0)
model = vq.VQ(enc_type="vit_b_enc", image_size=256, patch_size=16, codebook_size=16384, latent_dim=32, norm_codes=True, post_mlp=True,
              sync_codebook=False).cuda().eval()
x = torch.randn(64, 3, 256, 256, device="cuda")
for mode in ("x3", "bf16"):
    os.environ["B200FM_VQ_PRECISION"] = mode
    with torch.no_grad():
        for _ in range(3):
            model.tokenize(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model.tokenize(x)
        e1.record(); torch.cuda.synchronize()
        print(f"== {mode}: {e0.elapsed_time(e1) / 5:.2f} ms per batch of 64 ({64 / (e0.elapsed_time(e1) / 5e3):.0f} img/s)")
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            model.tokenize(x)
            torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type.name != "CUDA":
            continue
        name = re.sub(r"\(.*", "", e.name)[:80]
        agg[name][0] += 1
        agg[name][1] += (e.device_time if hasattr(e, "device_time") else e.cuda_time) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"GPU busy: {tot:.2f} ms")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  n={n:4d}  {k}")
