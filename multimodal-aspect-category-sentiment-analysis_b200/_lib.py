"""ctypes binding of libfcmf_b200.so (the C ABI declared in include/fcmf_b200.h).

There is no CPU fallback and no PyTorch fallback: if the library is missing or a call fails, a RuntimeError is
raised with the library's own message (fcmf_last_error)."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfcmf_b200.so")

ABI_VERSION = 4               # FCMF_ABI_VERSION of include/fcmf_b200.h
F32, BF16 = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
EPI_NONE, EPI_GELU, EPI_TANH, EPI_DGELU = 0, 1, 2, 3

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float


class Dropout(C.Structure):
    """fcmf_dropout (include/fcmf_b200.h): p, 64-bit seed, optional DEVICE pointer to a uint64 added to the seed."""
    _fields_ = [("p", _f32), ("seed", C.c_uint64), ("seed_dev", _vp)]


class OptTensor(C.Structure):
    """fcmf_opt_tensor (include/fcmf_b200.h)."""
    _fields_ = [("p", _vp), ("g", _vp), ("m", _vp), ("v", _vp), ("n", _i64), ("lr", _f32), ("wd", _f32)]


class Seg(C.Structure):
    _fields_ = [("ptr", _vp), ("ld", _i64), ("rows", _i32), ("groups", _i32), ("idx", _vp)]


class AttnDesc(C.Structure):
    _fields_ = [("q", Seg * 2), ("k", Seg * 2), ("v", Seg * 2),
                ("mask_add", _vp), ("ld_mask", _i64), ("mask_div", _i32),
                ("bias", _vp), ("NP", _i32), ("heads", _i32), ("dh", _i32), ("scale", _f32), ("causal", _i32),
                ("drop", Dropout), ("engine", _i32)]


# name -> argtypes (every entry point returns int); must list EVERY symbol include/fcmf_b200.h declares.
PROTOTYPES = {
    "fcmf_abi_version": [],
    "fcmf_abi_layout": [C.c_int],
    "fcmf_dropout_keep": [_f32, C.c_uint64, C.c_uint64, C.c_uint32],
    "fcmf_device_info": [C.POINTER(C.c_int)] * 3,
    "fcmf_gemm_tn": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int, _vp],
    "fcmf_gemm_tn_f32": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, C.c_int, _vp],
    "fcmf_gemm_wgrad": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int, _vp],
    "fcmf_gemm_wgrad_plan": [_i64, _i64, _i64, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)],
    "fcmf_ln_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, C.POINTER(Dropout), C.c_int, _vp],
    "fcmf_ln_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(Dropout), _i64, C.c_int, _vp],
    "fcmf_mask_additive": [_vp, _i64, _vp, _i64, _i64, _vp],
    "fcmf_gather_sum_rows": [_vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, C.c_int, C.c_int, _vp],
    "fcmf_dtanh": [_vp, _vp, _vp, _i64, C.c_int, _vp],
    "fcmf_cast_matrix": [_vp, _vp, _i64, _i64, C.c_int, C.c_int, _vp],
    "fcmf_cast_matrix_ld": [_vp, _vp, _i64, _i64, _i64, C.c_int, C.c_int, _vp],
    "fcmf_cast_to_f32": [_vp, _vp, _i64, C.c_int, _vp],
    "fcmf_set_attn_engine": [C.c_int],
    "fcmf_attn_fwd": [C.POINTER(AttnDesc), _vp, _i64, _vp, C.c_int, _vp],
    "fcmf_attn_bwd": [C.POINTER(AttnDesc), _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp],
    "fcmf_box_geometry_fwd": [_vp, _vp, _vp, C.POINTER(_f32), _vp, _vp, _i64, _i32, _i32, _vp],
    "fcmf_box_geometry_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp],
    "fcmf_cls_ce_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, C.POINTER(Dropout), C.c_int, _vp],
    "fcmf_vocab_ce_fwd": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, C.c_int, _vp],
    "fcmf_vocab_ce_bwd": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i64, C.c_int, _vp],
    "fcmf_opt_sumsq": [_vp, _vp, _vp, _i64, _vp, _vp],
    "fcmf_opt_clip_coef": [_vp, _f32, _vp, _vp, _vp],
    "fcmf_opt_adamw": [_vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _i64, C.c_int, _vp],
    "fcmf_cls_ce_bwd": [_vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _i64, _i64, _i32, C.POINTER(Dropout), C.c_int, _vp],
}

_lock = threading.Lock()
_lib = None
launches = 0          # number of C-ABI compute calls issued by this process (bench.py reports it)


def load() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built -- the product path never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (nvcc, sm_100a). There is no CPU or PyTorch fallback for the fusion path.")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError => ABI mismatch, also loud
            fn.argtypes = argtypes
            fn.restype = C.c_int
        lib.fcmf_last_error.argtypes = []
        lib.fcmf_last_error.restype = C.c_char_p
        lib.fcmf_kernel_launches.argtypes = []
        lib.fcmf_kernel_launches.restype = C.c_longlong
        if lib.fcmf_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libfcmf_b200.so ABI version {lib.fcmf_abi_version()} != {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


# FCMF_NVTX=1: every C-ABI call and every stage of the folded path (fusion.py) is wrapped in an NVTX range, so an nsys /
# ncu timeline reads "fcmf_gemm_tn", "fcmf_attn_bwd", "fusion/text->image", ... (tracing hook, off by default).
NVTX = os.environ.get("FCMF_NVTX", "0") not in ("", "0")


class trace:
    """`with trace("fusion/text->image"):` -- an NVTX range when FCMF_NVTX is set, nothing otherwise."""
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


def mark(name: str) -> None:
    """Instantaneous NVTX marker (stage boundaries of the folded path) when FCMF_NVTX is set."""
    if NVTX:
        import torch
        torch.cuda.nvtx.mark(name)


def call(name: str, *args) -> None:
    global launches
    lib = load()
    if NVTX:
        with trace(name):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.fcmf_last_error().decode(errors='replace')}")
    launches += 1


def kernel_launches() -> int:
    """Kernels launched by libfcmf_b200.so in this process so far."""
    return int(load().fcmf_kernel_launches())


def exported_symbols():
    return list(PROTOTYPES) + ["fcmf_last_error", "fcmf_kernel_launches"]
