"""ctypes binding of libvjepa2_b200.so (include/vjepa2_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call fails, a
RuntimeError is raised -- the product path never silently runs anything but the sm_100a kernels.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvjepa2_b200.so")

ABI_VERSION = 3          # vj_abi_version() of the library this binding (struct layouts, prototypes) was written for
VJ_BF16, VJ_F32 = 0, 1
EPI_BIAS, EPI_GELU, EPI_DGELU, EPI_RESIDUAL = 1, 2, 4, 8
EPI_OUT_F32, EPI_RES_F32, EPI_ROUND_BF16, EPI_AUX_OUT, EPI_ROPE, EPI_BIAS_GRAD = 16, 32, 64, 128, 256, 512


class GemmArgs(Structure):
    _fields_ = [
        ("a", c_void_p), ("b", c_void_p), ("out", c_void_p),
        ("M", c_int64), ("N", c_int64), ("K", c_int64),
        ("lda", c_int64), ("ldb", c_int64), ("ldo", c_int64),
        ("a_mn_major", c_int32), ("b_mn_major", c_int32), ("flags", c_int32),
        ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_int64),
        ("aux_out", c_void_p), ("aux_in", c_void_p), ("ld_aux", c_int64),
        ("rope_table", c_void_p), ("rope_hd", c_int32), ("rope_D", c_int32),
        ("bias_grad", c_void_p),
    ]


class MaskSpec(Structure):
    _fields_ = [
        ("frames", c_int32), ("rows", c_int32), ("cols", c_int32), ("num_blocks", c_int32),
        ("context_frames", c_int32), ("max_keep", c_int32), ("full_complement", c_int32),
        ("pred_full_complement", c_int32),
        ("temporal_lo", c_double), ("temporal_hi", c_double), ("spatial_lo", c_double), ("spatial_hi", c_double),
        ("aspect_lo", c_double), ("aspect_hi", c_double),
    ]


MASK_RNG_WORDS = 626
MAX_PEERS = 16


class PtrList(Structure):
    _fields_ = [("ptr", c_void_p * MAX_PEERS)]


# name -> (restype, argtypes); every symbol include/vjepa2_b200.h declares
SIGNATURES = {
    "vj_last_error": (c_char_p, []),
    "vj_abi_version": (c_int, []),
    "vj_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "vj_gemm": (c_int, [POINTER(GemmArgs), c_void_p]),
    "vj_layernorm_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                 c_int64, c_int64, c_int64, c_float, c_void_p]),
    "vj_layernorm_bwd_scratch": (c_size_t, [c_int64, c_int64]),
    "vj_layernorm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "vj_rope_table": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vj_rope_apply": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    "vj_attn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vj_attn_bwd_scratch": (c_size_t, [c_int, c_int, c_int, c_int]),
    "vj_attn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                            c_int, c_int, c_void_p]),
    "vj_gather_rows": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "vj_scatter_add_rows": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "vj_mask_to_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "vj_im2col_tubelets": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int64, c_int, c_void_p]),
    "vj_colsum_scratch": (c_size_t, [c_int64, c_int64]),
    "vj_colsum": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p]),
    "vj_l1_scratch": (c_size_t, [c_int64, c_int64, c_int64]),
    "vj_l1_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                           c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "vj_argsort_rank": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "vj_pred_indices": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "vj_ema_update": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "vj_grad_check": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "vj_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                              c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vj_adam_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_int, ctypes.c_double, ctypes.c_double, c_void_p]),
    "vj_scaler_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_float,
                                 c_void_p]),
    "vj_fill_f32": (c_int, [c_void_p, c_int64, c_float, c_void_p]),
    "vj_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vj_mask_collate_scratch": (c_size_t, [POINTER(MaskSpec), c_int64]),
    "vj_mask_collate": (c_int, [c_void_p, POINTER(MaskSpec), ctypes.c_uint32, c_int64, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "vj_peer_barrier": (c_int, [POINTER(PtrList), c_int, c_int, ctypes.c_uint32, c_void_p]),
    "vj_sum_into": (c_int, [c_void_p, POINTER(PtrList), c_int, c_int64, c_void_p]),
    "vj_gemm_set_pair_mode": (c_int, [c_int]),
    "vj_gemm_set_epi16_mode": (c_int, [c_int]),
}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"vjepa2_b200: {LIB_PATH} not found.  Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C vjepa2_b200/csrc`).  There is no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.vj_abi_version()
    if got != ABI_VERSION:               # a stale .so with another vj_gemm_args layout must not load silently
        raise RuntimeError(f"vjepa2_b200: {LIB_PATH} has ABI version {got}, this binding expects {ABI_VERSION}; "
                           "rebuild it (`make -C vjepa2_b200/csrc`)")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().vj_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"vjepa2_b200 {what} failed (rc={rc}): {last_error()}")
