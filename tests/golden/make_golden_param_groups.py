"""Golden fixture for the optimizer parameter groups (a18): the UNMODIFIED reference's `get_parameter_groups`
(fourm/utils/optim_factory.py:111-168, as called by create_optimizer with filter_bias_and_bn=True and the model's
no_weight_decay() skip list, :188-199) on the reference 4M-Tiny mod7 model.

Run in the authoring container only:   python tests/golden/make_golden_param_groups.py   -> tests/golden/param_groups_golden.json
"""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402
import make_golden as MG  # noqa: E402
from oracle import fourm_oracle as O  # noqa: E402


def main():
    fm, fm_utils, MODALITY_INFO = ref_import.import_reference_models()
    from fourm.utils import optim_factory
    assert optim_factory.__file__.startswith("/root/reference")
    out = {}
    for tag, kw in {"tiny": {}, "tiny_qknorm": dict(qk_norm=True)}.items():
        model = MG.build_reference_fourm("fm_tiny_6e_6d_swiglu_nobias", O.mod7_specs(), MODALITY_INFO, **kw)
        skip = model.no_weight_decay()
        names = {id(p): n for n, p in model.named_parameters()}
        with contextlib.redirect_stdout(io.StringIO()):
            groups = optim_factory.get_parameter_groups(model, weight_decay=0.05, skip_list=skip)
        out[tag] = {("decay" if g["weight_decay"] > 0 else "no_decay"): sorted(names[id(p)] for p in g["params"]) for g in groups}
        out[tag]["skip_list"] = sorted(skip)
        print(tag, {k: len(v) for k, v in out[tag].items()})
    json.dump(out, open(os.path.join(HERE, "param_groups_golden.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
