"""Golden fixture for the model EMA (reference `ModelEmaV2`, fourm/utils/timm/model_ema.py:84-127, as run_training_vqvae.py:683, 1171
uses it): the UNMODIFIED reference class on CPU over 3 updates of a small Linear / BatchNorm / Linear model whose weights move by
seeded perturbations.  Stores the EMA state_dict after every update; the GPU test replays the same perturbations through
b200fm.optim.FusedModelEma and demands bit equality (IEEE fp32 mul / mul / add on both sides).

    python tests/golden/make_golden_ema.py          (authoring container only)"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

DECAY = 0.9999
SEED = 7


def build_model():
    torch.manual_seed(SEED)
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.BatchNorm1d(64), torch.nn.Linear(64, 1031))


def perturbations(model, step):
    g = torch.Generator().manual_seed(1000 + step)
    return [torch.randn(p.shape, generator=g) * 0.1 for p in model.parameters()]


def main():
    # the class lives in a leaf module without package-level dependencies: load the reference's file directly
    path = os.path.join(ref_import.REFERENCE_ROOT, "fourm", "utils", "timm", "model_ema.py")
    spec = importlib.util.spec_from_file_location("ref_model_ema", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = build_model()
    ema = mod.ModelEmaV2(model, decay=DECAY)
    states = []
    for step in range(3):
        with torch.no_grad():
            for p, d in zip(model.parameters(), perturbations(model, step)):
                p.add_(d)
            model[1].running_mean.add_(0.5)
            model[1].num_batches_tracked.add_(1)
        ema.update(model)
        states.append({k: v.clone() for k, v in ema.module.state_dict().items()})
    out = os.path.join(HERE, "ema_golden.pt")
    torch.save(dict(meta=dict(torch=str(torch.__version__), reference_commit="cda590f", source=path.replace(ref_import.REFERENCE_ROOT, "")),
                    decay=DECAY, seed=SEED, states=states), out)
    print(out, os.path.getsize(out) // 1024, "KiB", {k: tuple(v.shape) for k, v in states[-1].items()})


if __name__ == "__main__":
    main()
