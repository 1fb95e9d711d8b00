"""Golden fixtures for the VQ-VAE training side (a25): the UNMODIFIED reference's codebook EMA update
(quantize_lucid.py:263-301, 388-426) and one VQVAE training step (vqvae.py:454-471 + ViTDecoder) on deterministic
weights / inputs.  Dead-code expiry is disabled (threshold 0) because it draws from the device RNG.

Run in the authoring container only:   python tests/golden/make_golden_vqtrain.py   -> tests/golden/vq_train_golden.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402
from oracle import fourm_oracle as O  # noqa: E402

VQVAE_KW = dict(enc_type="vit_s_enc", dec_type="vit_s_dec", image_size=64, patch_size=16, codebook_size=256, latent_dim=32,
                norm_codes=True, post_mlp=True, sync_codebook=False, ema_decay=0.9, threshold_ema_dead_code=0.0)


def vqvae_state_dict(model):
    sd = {}
    for k, v in model.state_dict().items():
        if k.endswith("pos_emb") or k.endswith("initted") or k.endswith("cluster_size"):
            sd[k] = v.clone()
        elif k.endswith("_codebook.embed"):
            sd[k] = torch.nn.functional.normalize(O.deterministic_tensor("quantize._codebook.embed", v.shape, 1.0), dim=-1)
        else:
            sd[k] = O.deterministic_tensor(k, v.shape, 0.05 if v.ndim > 1 else 0.02)
    return sd


def main():
    ref_import.install()
    from fourm.vq.quantizers.quantize_lucid import CosineSimCodebook, EuclideanCodebook
    import fourm.vq.vqvae as vqvae
    assert vqvae.__file__.startswith("/root/reference")
    gold = dict(meta=dict(torch=torch.__version__, reference_commit="cda590f"))

    # ---- stand-alone codebook updates: two consecutive training steps ----
    g = torch.Generator().manual_seed(21)
    z = [torch.randn(2, 400, 32, generator=g), torch.randn(2, 400, 32, generator=g) * 0.5 + 0.2]
    cb = CosineSimCodebook(dim=32, codebook_size=512, decay=0.9, threshold_ema_dead_code=0).train()
    e0 = torch.nn.functional.normalize(torch.randn(512, 32, generator=g), dim=-1)
    cb.embed.copy_(e0)
    cos = dict(z=[t.clone() for t in z], embed0=e0.clone(), steps=[])
    for t in z:
        q, idx = cb(t)
        cos["steps"].append(dict(quant_sum=q.double().sum(-1), idx=idx.clone(), embed=cb.embed.clone(), cluster_size=cb.cluster_size.clone()))
    eb = EuclideanCodebook(dim=32, codebook_size=300, decay=0.8, threshold_ema_dead_code=0).train()
    e1 = torch.randn(300, 32, generator=g)
    eb.embed.copy_(e1); eb.embed_avg.copy_(e1)
    l2 = dict(embed0=e1.clone(), steps=[])
    for t in z:
        q, idx = eb(t)
        l2["steps"].append(dict(quant_sum=q.double().sum(-1), idx=idx.clone(), embed=eb.embed.clone(), embed_avg=eb.embed_avg.clone(),
                                cluster_size=eb.cluster_size.clone()))
    gold["cosine"], gold["euclid"] = cos, l2

    # ---- one VQVAE training step (fp32, no autocast): loss = mse(dec, x) + code_loss ----
    model = vqvae.VQVAE(**VQVAE_KW).train()
    sd = vqvae_state_dict(model)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 3, 64, 64, generator=g)
    dec, code_loss = model(x)
    rec = torch.nn.functional.mse_loss(dec, x)
    (rec + code_loss.sum()).backward()
    grads = {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None}
    keys = ["encoder.proj.weight", "encoder.proj.bias", "encoder.blocks.0.attn.qkv.weight", "encoder.post_mlp.fc2.weight", "quant_proj.weight",
            "quant_proj.bias", "post_quant_proj.weight", "decoder.blocks.3.mlp.fc1.weight", "decoder.out_proj.weight", "decoder.out_proj.bias",
            "decoder.norm_mlp.weight"]
    gold["vqvae"] = dict(kw=VQVAE_KW, shapes={k: tuple(v.shape) for k, v in sd.items()},
                         weight_checksums={k: float(v.double().sum()) for k, v in sd.items()},
                         param_names=[k for k, _ in model.named_parameters()],
                         rec_loss=rec.detach().clone(), code_loss=code_loss.detach().clone(), dec_slice=dec[:, :, :8, :8].detach().clone(),
                         dec_norm=float(dec.norm()), embed_after=model.quantize._codebook.embed.clone(),
                         cluster_size_after=model.quantize._codebook.cluster_size.clone(),
                         grad_norm={k: float(v.norm()) for k, v in grads.items()},
                         grad_slices={k: grads[k].flatten()[:64].clone() for k in keys})
    with torch.no_grad():
        _, _, tokens = model.eval().encode(x)
    gold["vqvae"]["tokens_after"] = tokens.clone()
    print("rec", float(rec), "code", float(code_loss), "grads", len(grads))
    torch.save(gold, os.path.join(HERE, "vq_train_golden.pt"))
    print(os.path.getsize(os.path.join(HERE, "vq_train_golden.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
