"""VQ tokenizer measurements (secondary metric of BASELINE.md): codebook scan micro-benchmark (cfg-5 size) and ViT-B tokenize."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
import torch.nn.functional as F
from b200fm import ops, lib

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
out = {}
# --- scan: z = l2norm(N(0,1)) [131072, 32], codebook = l2norm(U) [16384, 32]  (SURVEY 8d micro-benchmark)
n, K, d = 131072, 16384, 32
z = F.normalize(torch.randn(n, d, device=dev, generator=g), dim=-1)
cb = F.normalize(torch.rand(K, d, device=dev, generator=g) * 2 - 1, dim=-1)
for _ in range(3):
    idx = ops.vq_argmax(z, cb, cosine=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    idx = ops.vq_argmax(z, cb, cosine=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out["scan"] = dict(n=n, K=K, d=d, ms=ms, tflops_fp32=2.0 * n * K * d / (ms * 1e-3) / 1e12, latents_per_s=n / (ms * 1e-3),
                   algorithmic_bytes=n * d * 4 + K * d * 4 + n * 8, gbs=(n * d * 4 + K * d * 4 + n * 8) / (ms * 1e-3) / 1e9)
# reference-style scan on the same GPU (materialised [n, K] fp32 sim + argmax), for context
t_ref = None
try:
    zz, cc = z[:32768], cb
    for _ in range(2):
        r = (zz @ cc.t()).argmax(-1)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        r = (zz @ cc.t()).argmax(-1)
    e1.record(); torch.cuda.synchronize()
    t_ref = e0.elapsed_time(e1) / 5 * (n / 32768)
except Exception as ex:      # pragma: no cover
    t_ref = str(ex)
out["scan"]["torch_eager_same_gpu_ms_extrapolated"] = t_ref
# --- host-buffer path (what save_vq_tokens.py would call): H2D of z + scan + D2H of idx
zh = z.cpu().pin_memory(); ih = torch.empty(n, dtype=torch.int64).pin_memory()
stream = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    lib.call("b200fm_vq_argmax_host", zh.data_ptr(), cb.data_ptr(), ih.data_ptr(), n, K, d, 1, stream)
t = time.perf_counter()
for _ in range(5):
    lib.call("b200fm_vq_argmax_host", zh.data_ptr(), cb.data_ptr(), ih.data_ptr(), n, K, d, 1, stream)
out["scan"]["e2e_host_ms"] = (time.perf_counter() - t) / 5 * 1e3
assert torch.equal(ih.cuda(), idx)
# --- ViT-B tokenizer forward: 256x256, K = 16384 (cfg-5 / save_vq_tokens shapes)
import fourm.vq as vq
torch.manual_seed(0)
model = vq.VQ(enc_type="vit_b_enc", image_size=256, patch_size=16, codebook_size=16384, latent_dim=32, norm_codes=True, post_mlp=True,
              sync_codebook=False).to(dev).eval()
B = 64
x = torch.randn(B, 3, 256, 256, device=dev)
with torch.no_grad():
    for _ in range(3):
        tok = model.tokenize(x)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        tok = model.tokenize(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out["tokenize_vit_b_256"] = dict(batch=B, ms=ms, img_per_s=B / (ms * 1e-3), tokens_per_s=B * 256 / (ms * 1e-3),
                                 tflops=48.6e9 * B / (ms * 1e-3) / 1e12)
# --- cfg-5: one VQ-VAE training step (ViT-B encoder + ViT-B decoder @256^2, K = 16384, d = 32, EMA codebook, MSE loss, fused AdamW), B = 64
from b200fm.optim import FusedAdamW
vae = vq.VQVAE(enc_type="vit_b_enc", dec_type="vit_b_dec", image_size=256, patch_size=16, codebook_size=16384, latent_dim=32, norm_codes=True,
               post_mlp=True, sync_codebook=False, ema_decay=0.99).to(dev).train()
opt = FusedAdamW([p for p in vae.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99), weight_decay=0.0)


def vae_step():
    dec, code_loss = vae(x)
    loss = F.mse_loss(dec.float(), x) + code_loss.sum()
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


for _ in range(3):
    vae_step()
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    l = vae_step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
n_par = sum(p.numel() for p in vae.parameters())
# fwd+bwd FLOPs: encoder + decoder ViT-B at 256 tokens = 2 x 48.6 GFLOP fwd per image, x3 for fwd+bwd
out["vqvae_train_vit_b_256"] = dict(batch=B, ms=ms, img_per_s=B / (ms * 1e-3), params_m=n_par / 1e6, loss=float(l),
                                    model_tflops=3 * 2 * 48.6e9 * B / (ms * 1e-3) / 1e12)
print(json.dumps(out))
