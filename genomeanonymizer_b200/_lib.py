"""Loader of the CUDA engine's C-ABI library.  Fails loudly: there is no CPU or PyTorch fallback."""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GA_B200_LIB") or os.path.join(_HERE, "libga_b200.so")   # override: A/B timing of kernel variants
_LIB = None

# every entry point declared in include/ga_b200.h
EXPORTS = ["ga_abi_version", "ga_status_string", "ga_engine_create", "ga_engine_destroy", "ga_last_error",
           "ga_upload_reference", "ga_run", "ga_run_host", "ga_last_host_traffic", "ga_launch_count", "ga_last_kernel_ms", "ga_kernel_ms_history", "ga_stage_ms_history", "ga_last_fallback_sessions", "ga_engine_keep_edits", "ga_record_edits", "ga_fastq_layout", "ga_fastq_render", "ga_result_digest"]
# include/ga_wire.h
WIRE_EXPORTS = ["ga_wire_pack_sizes", "ga_wire_pack", "ga_run_wire"]
# include/ga_synth.h - the synthetic-input generator lives in its own library (never needed by the masking path)
SYNTH_LIB_PATH = os.path.join(_HERE, "libga_synth.so")
SYNTH_EXPORTS = ["ga_synth_plan_sizes", "ga_synth_reference", "ga_synth_sessions", "ga_synth_reads_count", "ga_synth_reads_fill",
                 "ga_synth_reference_host", "ga_synth_sessions_host", "ga_synth_reads_count_host", "ga_synth_reads_fill_host"]
_SYNTH = None


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
    L.ga_kernel_ms_history.restype = C.c_int
    L.ga_kernel_ms_history.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
    L.ga_stage_ms_history.restype = C.c_int
    L.ga_stage_ms_history.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]
    L.ga_engine_keep_edits.restype = C.c_int
    L.ga_engine_keep_edits.argtypes = [C.c_void_p, C.c_int]
    L.ga_record_edits.restype = C.c_int
    L.ga_record_edits.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_uint32)]
    L.ga_last_fallback_sessions.restype = C.c_int
    L.ga_last_fallback_sessions.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int]
    L.ga_fastq_layout.restype = C.c_int
    L.ga_fastq_layout.argtypes = [C.c_void_p, C.POINTER(_abi.GaReads), C.POINTER(_abi.GaResult), C.c_int64, C.POINTER(_abi.GaFastqItems),
                                  C.c_void_p, C.c_void_p]
    L.ga_fastq_render.restype = C.c_int
    L.ga_fastq_render.argtypes = [C.c_void_p, C.POINTER(_abi.GaReads), C.POINTER(_abi.GaResult), C.c_int64, C.POINTER(_abi.GaFastqItems),
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.ga_run_host.restype = C.c_int
    L.ga_run_host.argtypes = [C.c_void_p, C.POINTER(_abi.GaReads), C.POINTER(_abi.GaSessions), C.POINTER(_abi.GaResult), C.c_int64]
    L.ga_last_host_traffic.restype = None
    L.ga_last_host_traffic.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.ga_result_digest.restype = C.c_int
    L.ga_result_digest.argtypes = [C.c_void_p, C.POINTER(_abi.GaResult), C.c_int64, C.POINTER(_abi.GaDigestIds), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    L.ga_wire_pack_sizes.restype = C.c_int
    L.ga_wire_pack_sizes.argtypes = [C.POINTER(_abi.GaReads), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.ga_wire_pack.restype = C.c_int
    L.ga_wire_pack.argtypes = [C.POINTER(_abi.GaReads), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]
    L.ga_run_wire.restype = C.c_int
    L.ga_run_wire.argtypes = [C.c_void_p, C.POINTER(_abi.GaReadsWire), C.POINTER(_abi.GaSessions), C.POINTER(_abi.GaResult), C.c_int64]
    _LIB = L
    return L


def synth_lib():
    """libga_synth.so: benchmark / test input generator (include/ga_synth.h).  Loading it does not map the engine."""
    global _SYNTH
    if _SYNTH is not None:
        return _SYNTH
    if not os.path.exists(SYNTH_LIB_PATH):
        raise RuntimeError(f"{SYNTH_LIB_PATH} is missing - build it with `python -m genomeanonymizer_b200.build`")
    L = C.CDLL(SYNTH_LIB_PATH)
    P = C.POINTER(_abi.GaSynthParams)
    vp = C.c_void_p
    L.ga_synth_plan_sizes.restype = C.c_int
    L.ga_synth_plan_sizes.argtypes = [P, C.POINTER(_abi.GaSynthPlan)]
    L.ga_synth_reference.restype = C.c_int
    L.ga_synth_reference.argtypes = [P, vp, C.c_int64, C.c_int64, vp]
    L.ga_synth_sessions.restype = C.c_int
    L.ga_synth_sessions.argtypes = [P] + [vp] * 8 + [vp]
    L.ga_synth_reads_count.restype = C.c_int
    L.ga_synth_reads_count.argtypes = [P, vp, vp, vp, vp]
    L.ga_synth_reads_fill.restype = C.c_int
    L.ga_synth_reads_fill.argtypes = [P, C.POINTER(_abi.GaReads), vp, vp]
    L.ga_synth_reference_host.restype = C.c_int
    L.ga_synth_reference_host.argtypes = [P, vp, C.c_int64, C.c_int64]
    L.ga_synth_sessions_host.restype = C.c_int
    L.ga_synth_sessions_host.argtypes = [P] + [vp] * 8
    L.ga_synth_reads_count_host.restype = C.c_int
    L.ga_synth_reads_count_host.argtypes = [P, vp, vp, vp]
    L.ga_synth_reads_fill_host.restype = C.c_int
    L.ga_synth_reads_fill_host.argtypes = [P, C.POINTER(_abi.GaReads), vp]
    _SYNTH = L
    return L
