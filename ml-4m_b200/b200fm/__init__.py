"""b200fm -- Python host side of the B200-native 4M hot path.

`b200fm.lib`  loads libb200fm.so (the C-ABI CUDA library, sm_100a only) through ctypes; there is NO fallback:
              a missing library or a non-CUDA tensor raises.
`b200fm.ops`  thin tensor-level wrappers + autograd.Functions built on the C ABI.
The reference-facing module surface lives in the sibling overlay package `fourm/` (fourm.models.fm, fourm.vq, ...).
"""
from . import lib  # noqa: F401
