"""Golden fixture for the 13 registered model presets (fm.py:840-1130): state_dict keys + shapes (as a digest) and parameter counts of
the UNMODIFIED reference constructors on the mod7 embeddings.  Weight initialisation is patched out (shapes only), so the 2.8 B
parameter XL presets build in seconds.

Run in the authoring container only:   python tests/golden/make_golden_presets.py   -> tests/golden/presets_golden.json
"""
import contextlib
import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


@contextlib.contextmanager
def no_init():
    """nn.init.* become no-ops: parameters stay torch.empty (untouched virtual memory)."""
    names = ["trunc_normal_", "normal_", "uniform_", "kaiming_uniform_", "xavier_uniform_", "constant_", "zeros_", "ones_"]
    saved = {n: getattr(torch.nn.init, n) for n in names}
    try:
        for n in names:
            setattr(torch.nn.init, n, lambda t, *a, **k: t)
        yield
    finally:
        for n, f in saved.items():
            setattr(torch.nn.init, n, f)


def digest(model):
    sd = model.state_dict()
    lines = [f"{k}:{tuple(v.shape)}:{str(v.dtype)}" for k, v in sd.items()]
    return dict(n_keys=len(lines), n_params=sum(p.numel() for p in model.parameters()),
                n_param_tensors=len(list(model.parameters())), sha256=hashlib.sha256("\n".join(lines).encode()).hexdigest())


def main():
    import ref_import
    import make_golden as MG
    from oracle import fourm_oracle as O
    fm, fm_utils, MODALITY_INFO = ref_import.import_reference_models()
    out = {}
    for name in fm.__all__:
        with no_init():
            model = MG.build_reference_fourm(name, O.mod7_specs(), MODALITY_INFO)
        out[name] = digest(model)
        print(name, out[name]["n_params"], out[name]["n_keys"])
        del model
    json.dump(out, open(os.path.join(HERE, "presets_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
