"""ctypes wrapper of oracle/libga_oracle.so - TEST INFRASTRUCTURE (see ga_oracle.c header).

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from genomeanonymizer_b200 import _abi
from genomeanonymizer_b200.batch import MaskResult, ReadBatch, SessionTable, decode_result

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libga_oracle.so")
    src = os.path.join(_HERE, "ga_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ga_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libga_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libga_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.ga_oracle_run.restype = C.c_int
        _LIB.ga_oracle_run.argtypes = [C.POINTER(_abi.GaReads), C.POINTER(_abi.GaSessions), C.c_void_p, C.c_int64,
                                       C.POINTER(_abi.GaResult), C.c_int]
        _LIB.ga_oracle_threads.restype = C.c_int
    return _LIB


def n_threads():
    return int(lib().ga_oracle_threads())


def run(batch: ReadBatch, sessions: SessionTable, reference: bytes, threads: int = 0, decode: bool = True,
        cap_frac: float = 1.0):
    """Returns (MaskResult | raw dict, status)."""
    keep = []
    R = batch.as_struct(keep)
    S = sessions.as_struct(keep)
    n = batch.n_reads
    cap_rec = max(16, int(2 * n * cap_frac) + 16)
    cap16 = int(batch.seq4.shape[0] // 16 * cap_frac) * 4 + 4096
    mod_sess = np.zeros(cap_rec, np.int32)
    mod_read = np.zeros(cap_rec, np.int32); mod_len = np.zeros(cap_rec, np.uint32)
    mod_so = np.zeros(cap_rec, np.uint32); mod_qo = np.zeros(cap_rec, np.uint32)
    out_seq = np.zeros(cap16 * 16, np.uint8); out_qual = np.zeros(cap16 * 32, np.uint8)
    counts = np.zeros(max(1, sessions.n_sessions) * 4, np.uint32)
    totals = _abi.GaTotals()
    res = _abi.GaResult()
    res.cap_records, res.cap_seq16, res.cap_qual16 = cap_rec, cap16, cap16
    res.mod_session = mod_sess.ctypes.data
    res.mod_read, res.mod_len = mod_read.ctypes.data, mod_len.ctypes.data
    res.mod_seq_off16, res.mod_qual_off16 = mod_so.ctypes.data, mod_qo.ctypes.data
    res.out_seq4, res.out_qual = out_seq.ctypes.data, out_qual.ctypes.data
    res.sess_counts = counts.ctypes.data
    res.totals = C.addressof(totals)
    refb = np.frombuffer(reference if isinstance(reference, (bytes, bytearray)) else reference.encode("ascii"), dtype=np.uint8)
    st = lib().ga_oracle_run(C.byref(R), C.byref(S), refb.ctypes.data, len(refb), C.byref(res), int(threads))
    if not decode:
        return {"totals": totals, "counts": counts.reshape(-1, 4)}, st
    if st != _abi.GA_OK:
        return None, st
    return decode_result(sessions.n_sessions, totals, mod_sess, mod_read, mod_len, mod_so, mod_qo, out_seq, out_qual,
                         counts[:sessions.n_sessions * 4]), st
