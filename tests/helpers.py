"""Shared helpers for the tests (fixtures, deterministic weights, golden loading)."""
import ctypes
import os
import random

import torch

from oracle import fourm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def golden_state_dict(gold):
    """Rebuild the fixture weights from names+shapes and verify their checksums."""
    params = set(gold.get("param_names", []))
    sd = {}
    for k, shape in gold["shapes"].items():
        if k in params:
            sd[k] = O.deterministic_tensor(O.canonical_param_name(k, gold['shapes']), shape)
        elif k.endswith("pos_emb"):
            sd[k] = None    # filled by caller (sincos tables)
        else:
            sd[k] = torch.zeros(shape)
    return sd


def fill_fourm_buffers(sd, specs, dim):
    for name, s in specs.items():
        for side in ("encoder_embeddings", "decoder_embeddings"):
            k = f"{side}.{name}.pos_emb"
            if k in sd and sd[k] is None:
                if s["kind"] == "seq":
                    sd[k] = O.sincos_1d(512, dim)       # quirk: [1,512,D] (slice hits the batch dim)
                else:
                    side_len = s["image_size"] // s["patch_size"]
                    sd[k] = O.sincos_2d(side_len, side_len, dim)
    return sd


def decoder_order(seed, names):
    random.seed(seed)
    return random.sample(list(names), len(names))


def load_c_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "libvq_argmax_oracle.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    for fn in (lib.vq_cosine_argmax_oracle, lib.vq_euclid_argmax_oracle):
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                       ctypes.c_void_p, ctypes.c_void_p]
        fn.restype = None
    return lib
