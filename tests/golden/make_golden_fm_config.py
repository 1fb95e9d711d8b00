"""Golden fixture for the Hugging Face wrapper `FM(config)` (fm.py:783-831): state_dict digest of the UNMODIFIED reference for a
4M-7-style config (untied decoder heads: share_embedding=False).  Authoring container only -> tests/golden/fm_config_golden.json"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden_presets as MGP  # noqa: E402

MOD7 = ["rgb@224", "caption", "det", "tok_rgb@224", "tok_depth@224", "tok_normal@224", "tok_semseg@224", "tok_clip@224"]
CONFIGS = {
    "4m7_b_like": dict(domains_in=MOD7, domains_out=[m for m in MOD7 if m != "rgb@224"], encoder_depth=12, decoder_depth=12, dim=768, num_heads=12,
                       mlp_ratio=4, qkv_bias=False, proj_bias=False, mlp_bias=False, norm_bias=False, act_layer="SiLU", gated_mlp=True,
                       qk_norm=False, image_size=224, patch_size=16, share_modality_embeddings=True),
    "tiny_gelu_bias": dict(domains_in=MOD7[:4], domains_out=MOD7[1:4], encoder_depth=2, decoder_depth=3, dim=384, num_heads=6, mlp_ratio=4,
                           qkv_bias=True, proj_bias=True, mlp_bias=True, norm_bias=True, act_layer="GELU", gated_mlp=False, qk_norm=False,
                           image_size=224, patch_size=16, share_modality_embeddings=False),
}


def main():
    import ref_import
    fm, fm_utils, MODALITY_INFO = ref_import.import_reference_models()
    out = {}
    for tag, cfg in CONFIGS.items():
        with MGP.no_init():
            m = fm.FM(cfg)
        out[tag] = MGP.digest(m)
        print(tag, out[tag]["n_params"], out[tag]["n_keys"])
    json.dump(out, open(os.path.join(HERE, "fm_config_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
