"""Golden fixture for the tokenizer constructors: state_dict digests of the UNMODIFIED reference `VQ` / `VQVAE` for the ViT-S/B/L encoders
and decoders at the shipped option combinations.  Run in the authoring container only -> tests/golden/vq_presets_golden.json"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden_presets as MGP  # noqa: E402

CASES = {
    "vq_s_224": ("VQ", dict(enc_type="vit_s_enc", image_size=224, patch_size=16, codebook_size=8192, latent_dim=32, post_mlp=True)),
    "vq_b_224_nopost": ("VQ", dict(enc_type="vit_b_enc", image_size=224, patch_size=16, codebook_size=16384, latent_dim=32, post_mlp=False)),
    "vq_b_256_l2": ("VQ", dict(enc_type="vit_b_enc", image_size=256, patch_size=16, codebook_size=4096, latent_dim=16, norm_codes=False, post_mlp=True)),
    "vq_l_224_p8": ("VQ", dict(enc_type="vit_l_enc", image_size=224, patch_size=8, codebook_size=16384, latent_dim=32, post_mlp=True)),
    "vq_b_semseg": ("VQ", dict(enc_type="vit_b_enc", image_size=224, patch_size=16, n_labels=134, n_channels=64, codebook_size=4096, latent_dim=32, post_mlp=True)),
    "vqvae_b_b_256": ("VQVAE", dict(enc_type="vit_b_enc", dec_type="vit_b_dec", image_size=256, patch_size=16, codebook_size=16384, latent_dim=32, post_mlp=True)),
    "vqvae_s_l_224": ("VQVAE", dict(enc_type="vit_s_enc", dec_type="vit_l_dec", image_size=224, patch_size=16, codebook_size=8192, latent_dim=32, post_mlp=False)),
}


def main():
    import ref_import
    ref_import.install()
    import fourm.vq.vqvae as vqvae
    assert vqvae.__file__.startswith("/root/reference")
    out = {}
    for tag, (cls, kw) in CASES.items():
        with MGP.no_init():
            m = getattr(vqvae, cls)(sync_codebook=False, **kw)
        out[tag] = MGP.digest(m)
        print(tag, out[tag]["n_params"], out[tag]["n_keys"])
    json.dump(out, open(os.path.join(HERE, "vq_presets_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
