"""Known answers from the UNMODIFIED reference for the host-side helpers of the generation path:
decoding schedules (fourm/utils/generation.py), sentinel merging (fourm/utils/tokenizer/text_tokenizer.py) and the hot-path fields of
MODALITY_INFO (fourm/data/modality_info.py) that `b200fm.compat.local_modality_info` restates.  -> tests/golden/genutils_golden.json"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_import  # noqa: E402


def main():
    fm, fm_utils, MODALITY_INFO = ref_import.import_reference_models()
    from fourm.utils import generation as G
    from fourm.utils.tokenizer.text_tokenizer import merge_span_masking
    out = dict(schedules={}, merge=[], modality_info={})
    for steps, total in ((1, 196), (4, 196), (8, 196), (12, 256), (50, 196), (196, 196), (300, 196)):
        ts = G.linear_schedule(steps, total)
        out["schedules"][f"{steps}_{total}"] = dict(
            cosine=G.cosine_schedule(steps, total).tolist(), linear=ts.tolist(),
            linear_temp=G.linear_temp_schedule(1.5, ts).tolist(), onex=G.onex_temp_schedule(2.0, 0.1, ts, power=0.7).tolist(),
            cont=G.continue_schedule(ts.copy(), min(17, total - 1)).tolist())
    rng = np.random.RandomState(0)
    sent = set(range(4, 104))
    for _ in range(6):
        inp = [int(x) for x in rng.randint(4, 300, size=20)]
        dec = [int(x) for x in rng.randint(4, 300, size=40)]
        out["merge"].append(dict(inp=inp, dec=dec, merged=merge_span_masking(inp, dec, sent)))
    for name, d in MODALITY_INFO.items():
        ent = {k: d[k] for k in ("type", "vocab_size", "max_tokens", "min_tokens", "patch_size", "input_size", "id", "num_channels") if k in d}
        for side in ("encoder_embedding", "decoder_embedding"):
            f = d.get(side)
            ent[side] = None if f is None else dict(cls=f.func.__name__, kw={k: v for k, v in f.keywords.items()})
        out["modality_info"][name] = ent
    path = os.path.join(HERE, "genutils_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(path, os.path.getsize(path) // 1024, "KiB", len(out["modality_info"]), "modalities")


if __name__ == "__main__":
    main()
