"""Loader of the CUDA engine's C-ABI library.  Fails loudly: there is no CPU or PyTorch fallback."""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libga_b200.so")
_LIB = None

# every entry point declared in include/ga_b200.h
EXPORTS = ["ga_abi_version", "ga_status_string", "ga_engine_create", "ga_engine_destroy", "ga_last_error",
           "ga_upload_reference", "ga_run", "ga_launch_count", "ga_last_kernel_ms"]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing - build it with `python -m genomeanonymizer_b200.build` "
                           "(nvcc, sm_100a). The masking path has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.ga_abi_version.restype = C.c_int
    if L.ga_abi_version() != _abi.GA_ABI_VERSION:
        raise RuntimeError("libga_b200.so ABI version does not match the Python wrapper")
    L.ga_status_string.restype = C.c_char_p
    L.ga_status_string.argtypes = [C.c_int]
    L.ga_engine_create.restype = C.c_int
    L.ga_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ga_engine_destroy.restype = None
    L.ga_engine_destroy.argtypes = [C.c_void_p]
    L.ga_last_error.restype = C.c_char_p
    L.ga_last_error.argtypes = [C.c_void_p]
    L.ga_upload_reference.restype = C.c_int
    L.ga_upload_reference.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    L.ga_run.restype = C.c_int
    L.ga_run.argtypes = [C.c_void_p, C.POINTER(_abi.GaReads), C.POINTER(_abi.GaSessions), C.POINTER(_abi.GaResult), C.c_void_p]
    L.ga_launch_count.restype = C.c_int64
    L.ga_launch_count.argtypes = [C.c_void_p]
    L.ga_last_kernel_ms.restype = C.c_float
    L.ga_last_kernel_ms.argtypes = [C.c_void_p]
    _LIB = L
    return L
