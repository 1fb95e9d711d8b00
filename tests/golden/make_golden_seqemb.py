"""Golden fixture for SequenceEmbEncoderEmbedding (a6; encoder_embeddings.py:312-421, T5-XXL features of 4M-21): the UNMODIFIED
reference module, plain and bottleneck variants, forward + gradients.

Run in the authoring container only:   python tests/golden/make_golden_seqemb.py   -> tests/golden/seqemb_golden.pt
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


def inputs(B=3, L=77, E=4096):
    g = torch.Generator().manual_seed(41)
    feats = torch.randn(B, L, E, generator=g)
    mask = torch.rand(B, L, generator=g) < 0.35
    mask[0] = False
    mask[1, 5:] = True
    wx = torch.randn(B, L, 384, generator=g)
    we = torch.randn(B, L, 384, generator=g)
    return feats, mask, wx, we


def main():
    ref_import.install()
    import fourm.models.encoder_embeddings as ee
    assert ee.__file__.startswith("/root/reference")
    gold = dict(meta=dict(torch=torch.__version__, reference_commit="cda590f"), cases={})
    feats, mask, wx, we = inputs()
    for tag, kw in {"plain": dict(use_bottleneck=False), "bottleneck": dict(use_bottleneck=True, bottleneck_dim=64)}.items():
        m = ee.SequenceEmbEncoderEmbedding(max_length=77, dim_tokens=384, orig_emb_dim=4096, **kw)
        sd = {}
        for k, v in m.state_dict().items():
            sd[k] = v.clone() if k == "pos_emb" else O.deterministic_tensor("seqemb." + k, v.shape, 0.02)
        m.load_state_dict(sd)
        d = m(dict(tensor=feats.clone(), input_mask=mask.clone()))
        (d["x"] * wx).sum().add((d["emb"] * we).sum()).backward()
        gold["cases"][tag] = dict(kw=kw, shapes={k: tuple(v.shape) for k, v in sd.items()},
                                  x_sum=d["x"].detach().double().sum(-1), x_slice=d["x"].detach()[:, :6, :48].clone(),
                                  emb_sum=d["emb"].detach().double().sum(-1), emb_slice=d["emb"].detach()[:, :6, :48].clone(),
                                  grad_norm={k: float(p.grad.norm()) for k, p in m.named_parameters()},
                                  grad_slices={k: p.grad.flatten()[:64].clone() for k, p in m.named_parameters()})
        print(tag, float(d["x"].norm()), float(d["emb"].norm()))
    torch.save(gold, os.path.join(HERE, "seqemb_golden.pt"))
    print(os.path.getsize(os.path.join(HERE, "seqemb_golden.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
