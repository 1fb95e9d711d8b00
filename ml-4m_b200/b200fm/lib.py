"""ctypes binding of include/b200fm.h.  Fails loudly when the CUDA library is missing (no CPU fallback)."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200fm.so")

c_void_p, c_int, c_ll, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t

# name -> argtypes  (restype is int unless noted); must list every symbol declared in include/b200fm.h
SIGNATURES = {
    "b200fm_abi_version": [],
    "b200fm_device_info": [c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "b200fm_gemm_bf16": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll,
                         c_void_p, c_void_p, c_ll, c_float, c_void_p, c_void_p],
    "b200fm_layernorm_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p],
    "b200fm_layernorm_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_int, c_void_p],
    "b200fm_vq_argmax": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    "b200fm_vq_argmax_host": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
}

_lib = None
_lock = threading.Lock()


class B200FMError(RuntimeError):
    pass


def load():
    """Return the loaded CDLL; raises B200FMError if libb200fm.so has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200FMError(f"{LIB_PATH} not found: build it with `python -m b200fm.build` (nvcc, sm_100a). "
                              "b200fm has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.b200fm_last_error.restype = ctypes.c_char_p
        lib.b200fm_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        if lib.b200fm_abi_version() != 1:
            raise B200FMError("libb200fm.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        raise B200FMError(f"{what} failed (rc={rc}): {load().b200fm_last_error().decode(errors='replace')}")


def call(name, *args):
    check(getattr(load(), name)(*args), name)
