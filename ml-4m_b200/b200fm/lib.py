"""ctypes binding of include/b200fm.h.  Fails loudly when the CUDA library is missing (no CPU fallback)."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200fm.so")

c_void_p, c_int, c_ll, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t



class Segment(ctypes.Structure):
    """Mirror of `b200fm_segment` (include/b200fm.h)."""
    _fields_ = [("mask", c_void_p), ("ids", c_void_p), ("dam", c_void_p), ("token_emb", c_void_p), ("pos_emb", c_void_p),
                ("mod_emb", c_void_p), ("x_rows", c_void_p), ("d_token_emb", c_void_p), ("d_mod_emb", c_void_p),
                ("dx_rows", c_void_p), ("padding_idx", c_ll), ("L", c_int), ("kind", c_int), ("mod_id", c_int),
                ("max_length", c_int), ("ids_is_i64", c_int), ("reserved", c_int), ("d_pos_emb", c_void_p)]


MAX_SEGMENTS = 24
KIND_IMG, KIND_TOK_IMG, KIND_SEQ, KIND_SEQ_EMB = 0, 1, 2, 3

# name -> argtypes  (restype is int unless noted); must list every symbol declared in include/b200fm.h
SIGNATURES = {
    "b200fm_abi_version": [],
    "b200fm_device_info": [c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "b200fm_set_option": [ctypes.c_char_p, c_int],
    "b200fm_get_option": [ctypes.c_char_p, c_void_p],
    "b200fm_gemm_bf16": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll,
                         c_void_p, c_void_p, c_ll, c_float, c_void_p, c_void_p],
    "b200fm_layernorm_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p],
    "b200fm_add_layernorm_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                 c_float, c_void_p],
    "b200fm_layernorm_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_int, c_void_p],
    "b200fm_attention_fwd": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_void_p,
                             c_int, c_int, c_int, c_int, c_float, c_void_p],
    "b200fm_attention_bwd": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_void_p, c_ll,
                             c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_float,
                             c_void_p],
    "b200fm_swiglu_bwd": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p],
    "b200fm_act_bwd": [c_int, c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "b200fm_cross_entropy": [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_ll, c_int, c_void_p],
    "b200fm_colsum_bf16": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p],
    "b200fm_cast_f32_bf16": [c_void_p, c_void_p, c_ll, c_void_p],
    "b200fm_patchify": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200fm_adamw": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_float, c_int,
                     c_float, c_void_p],
    "b200fm_select_plan": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p],
    "b200fm_decoder_attention_mask": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "b200fm_embed_rows": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                          c_int, c_int, c_void_p],
    "b200fm_embed_rows_bwd": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_int, c_int, c_void_p],
    "b200fm_head_rows": [c_void_p, c_ll, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "b200fm_gather_rows_bf16": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "b200fm_gather_i64": [c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "b200fm_scatter_rows_bf16": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "b200fm_scatter_add_rows": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "b200fm_vq_ema_stats": [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "b200fm_vq_ema_update_cosine": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p],
    "b200fm_headnorm_fwd": [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_int, c_float, c_void_p],
    "b200fm_headnorm_bwd": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "b200fm_adamw_multi": [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_float, c_float, c_int, c_float, c_void_p],
    "b200fm_adamw_chunk_elems": [],
    "b200fm_vq_argmax": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    "b200fm_vq_argmax_host": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    "b200fm_gemm_bf16_dyn": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll,
                             c_void_p, c_void_p, c_ll, c_float, c_void_p, c_void_p, c_int, c_void_p],
    "b200fm_cross_entropy_dyn": [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p],
    "b200fm_masked_mean": [c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p],
    "b200fm_adamw_multi_gnorm": [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_float, c_float, c_int, c_float, c_void_p, c_void_p, c_void_p],
    "b200fm_ema_multi": [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p],
    "b200fm_adamw_multi_dev": [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p],
    "b200fm_select_plan_ordered": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p],
    "b200fm_gather_rows_bf16_dyn": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p],
    "b200fm_gather_i64_dyn": [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p],
    "b200fm_scatter_rows_bf16_dyn": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p],
    "b200fm_allreduce_f32_seq": [c_void_p, c_void_p, c_int, c_int, c_ll, c_ll, c_float, ctypes.c_uint, c_void_p, c_int, c_void_p],
    "b200fm_split_limbs": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    "b200fm_attention_f32": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int,
                             c_float, c_void_p],
    "b200fm_mask_images": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "b200fm_patchify_u8": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "b200fm_attention_decode": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_float, c_void_p],
    "b200fm_sample_top_p": [c_void_p, c_ll, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p],
    "b200fm_kv_append": [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_int, c_int, c_int, c_void_p],
    "b200fm_head_ce_ws_slots": [c_int],
    "b200fm_head_ce": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                       c_void_p, c_ll, c_void_p],
    "b200fm_comm_flag_bytes": [],
    "b200fm_comm_alloc": [c_ll, c_void_p],
    "b200fm_comm_free": [c_void_p],
    "b200fm_comm_ipc_export": [c_void_p, c_void_p],
    "b200fm_comm_ipc_open": [c_void_p, c_void_p],
    "b200fm_comm_ipc_close": [c_void_p],
    "b200fm_allreduce_f32": [c_void_p, c_void_p, c_int, c_int, c_ll, c_ll, c_float, ctypes.c_uint, c_int, c_void_p],
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
            if not hasattr(lib, name):      # tests/test_abi.py enforces that every declared symbol is exported
                continue
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


CALLS = {"n": 0}      # kernel-launching C-ABI calls issued by this process (bench.py reports it as gpu_launches)


_bound = {}


def call(name, *args):
    CALLS["n"] += 1
    fn = _bound.get(name)
    if fn is None:
        fn = _bound[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        check(rc, name)


def set_option(name: str, value: int) -> None:
    """Runtime option of the kernel library (include/b200fm.h: "pdl", "gemm_cta_pairs", "ln_bwd_v2", "sm_reserve", "gemv", "gemv_prefetch", "gemm_tma_store", "comm_slim", "gemm_debug"); for in-process A/B measurements."""
    check(load().b200fm_set_option(name.encode(), int(value)), "b200fm_set_option")


def get_option(name: str) -> int:
    v = ctypes.c_int(0)
    check(load().b200fm_get_option(name.encode(), ctypes.byref(v)), "b200fm_get_option")
    return v.value
