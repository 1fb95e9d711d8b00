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
    # the K = 768 / N = 768 family (proj / q forward, their dgrad and wgrad): 2.6 waves of CTA-pair tiles
    wp = rnd(D, D)
    yp = ops.gemm(x, wp)                                                # NT bf16   16384 x 768 x 768
    dxp = ops.gemm(yp, wp, layout=ops.LAYOUT_NN)                        # NN bf16   16384 x 768 x 768
    dwp = ops.gemm(yp, x, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32)   # TN f32    768 x 768 x 16384 (split-K)
    dab = ops.swiglu_bwd(ab, gate)                                      # SwiGLU backward, 16384 x 2048
    # fused AdamW, one 4.7 M-element tensor
    from b200fm.optim import FusedAdamW
    pw = torch.nn.Parameter(torch.randn(2304, 2048, device=dev))
    pw.grad = torch.randn_like(pw)
    FusedAdamW([pw], lr=1e-3).step()
    # selection plan + embedding gather of one mod-7 batch (encoder side) through the model
    if _ == 0:
        sys.path.insert(0, ROOT)
        from b200fm.compat import build_mod7_embeddings, create_model
        from b200fm.synthetic import mod7_batch
        enc, dec, info = build_mod7_embeddings()
        model = create_model("fm_base_12e_12d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).to(dev)
        batch = {m: {k: v.to(dev) for k, v in d.items()} for m, d in mod7_batch(128).items()}
    with torch.no_grad():
        model._embed_side(batch, False, 128, [m for m in batch if m in model.encoder_embeddings])
        model._embed_side(batch, True, 128, [m for m in batch if m in model.decoder_embeddings])
torch.cuda.synchronize()
print("ok")
