"""Golden tokens of the UNMODIFIED reference tokenizer at the size save_vq_tokens.py / cfg-5 use: ViT-B encoder, 256x256, K = 16384,
d = 32, cosine codebook, fp32 (no autocast, like save_vq_tokens.py:288).  -> tests/golden/vq_b_golden.pt (tokens, latents, margins)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_import  # noqa: E402
from oracle import fourm_oracle as O  # noqa: E402

KW = dict(enc_type="vit_b_enc", image_size=256, codebook_size=16384, latent_dim=32, norm_codes=True, post_mlp=True)


def main():
    ref_import.import_reference_models()
    import fourm.vq as vq
    m = vq.VQ(patch_size=16, sync_codebook=False, **KW).eval()
    sd = {}
    for k, v in m.state_dict().items():
        if k.endswith("pos_emb") or k.endswith("initted") or k.endswith("cluster_size"):
            sd[k] = v.clone()
        elif k.endswith("_codebook.embed") or k.endswith("embed_avg"):
            sd[k] = torch.nn.functional.normalize(O.deterministic_tensor("quantize._codebook.embed", v.shape, 1.0), dim=-1)
        else:
            sd[k] = O.deterministic_tensor(k, v.shape, 0.05 if v.ndim > 1 else 0.02)
    m.load_state_dict(sd)
    x = torch.randn(4, 3, 256, 256, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        tokens = m.tokenize(x)
        lat = m.quant_proj(m.encoder(x))
        z = torch.nn.functional.normalize(lat.permute(0, 2, 3, 1).reshape(-1, 32), dim=-1)
        top2 = (z @ sd["quantize._codebook.embed"].t()).topk(2, dim=1).values
    gold = dict(meta=dict(torch=torch.__version__, reference_commit="cda590f"), kw=KW, shapes={k: tuple(v.shape) for k, v in sd.items()},
                tokens=tokens.clone(), latents=lat.clone(), margin=(top2[:, 0] - top2[:, 1]).clone())
    path = os.path.join(HERE, "vq_b_golden.pt")
    torch.save(gold, path)
    print(path, os.path.getsize(path) // 1024, "KiB", tokens.flatten()[:8].tolist(), "median margin", float(gold["margin"].median()))


if __name__ == "__main__":
    main()
