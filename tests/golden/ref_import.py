"""Import the UNMODIFIED apple/ml-4m reference (read-only tree at /root/reference) in-process.

Only usable in the authoring container (the GPU box has no /root/reference); it is used by
`make_golden.py` to generate the committed fixtures and by `tests/test_oracle_vs_reference.py`
(auto-skipped when the tree is absent) to pin the oracle restatement against the real thing.

Two shims:
  1. stub modules for optional deps that are not installed (boto3, diffusers, timm, webdataset, ...);
     xformers is deliberately NOT stubbed so fourm/vq/models/vit_models.py:26-31 takes its naive path.
  2. `random.sample(dict_items, n)` (fourm/models/fm.py:306) raises on Python >= 3.11; wrap it so the
     population is listified first (consumes the RNG identically to the <= 3.10 behaviour).
"""
import importlib.abc
import importlib.machinery
import os
import random
import sys
import types

REFERENCE_ROOT = os.environ.get("ML4M_REFERENCE", "/root/reference")
_STUBBED = ("boto3", "botocore", "diffusers", "timm", "webdataset", "albumentations", "braceexpand",
            "ftfy", "torchmetrics", "matplotlib", "wandb", "cv2", "torchvision")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fourm", "models"))


class _IterMeta(type):
    def __iter__(cls):
        return iter(())

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_dummy(name)


def _make_dummy(name):
    def _identity_decorator(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f
    ns = {"__init__": lambda self, *a, **k: None, "register_to_config": staticmethod(_identity_decorator)}
    return _IterMeta(name, (), ns)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name == "register_to_config":
            return lambda f: f
        val = _make_dummy(name)
        setattr(self, name, val)
        return val


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, names):
        self.names = set(names)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        return None


_installed = False


def install(extra_first_paths=()):
    """Put the reference on sys.path (after `extra_first_paths`) and install both shims."""
    global _installed
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _installed:
        missing = []
        for name in _STUBBED:
            try:
                __import__(name)
            except Exception:
                for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                    del sys.modules[k]
                missing.append(name)
        sys.meta_path.append(_StubFinder(missing))
        _orig_sample = random.sample

        def _sample(population, k, **kw):
            if not isinstance(population, (list, tuple, range, str)):
                population = list(population)
            return _orig_sample(population, k, **kw)
        random.sample = _sample
        _installed = True
    for p in list(extra_first_paths) + [REFERENCE_ROOT]:
        if p not in sys.path:
            sys.path.append(p)


def import_reference_models():
    """Return the reference's own modules (never the overlay): must be called in a process where the
    overlay `ml-4m_b200/` directory is NOT ahead of the reference on sys.path."""
    install()
    import fourm.models.fm as fm
    import fourm.models.fm_utils as fm_utils
    from fourm.data.modality_info import MODALITY_INFO
    assert fm.__file__.startswith(REFERENCE_ROOT), fm.__file__
    return fm, fm_utils, MODALITY_INFO
