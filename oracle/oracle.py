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
    hdrs = [os.path.join(_HERE, "..", "include", h) for h in ("ga_b200.h", "ga_digest.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
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
        _LIB.ga_oracle_set_reapply.restype = None
        _LIB.ga_oracle_set_reapply.argtypes = [C.c_void_p, C.c_int64]
        _LIB.ga_oracle_set_edit_sink.restype = None
        _LIB.ga_oracle_set_edit_sink.argtypes = [C.c_void_p, C.c_int64]
        _LIB.ga_oracle_digest.restype = C.c_int
        _LIB.ga_oracle_digest.argtypes = [C.POINTER(_abi.GaResult), C.c_int64, C.POINTER(_abi.GaDigestIds), C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def n_threads():
    return int(lib().ga_oracle_threads())


def run(batch: ReadBatch, sessions: SessionTable, reference: bytes, threads: int = 0, decode: bool = True,
        cap_frac: float = 1.0, reapply=(), edits: bool = False):
    """Returns (MaskResult | raw dict, status).  reapply: (session, read) pairs whose left-over indels the reference
    applies twice (quirk Q12 of DESIGN.md; the whole-sample checkers pass the pairs the plan flags).  edits: records carry
    "edits" as Engine.run(edits=True) returns them."""
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
    keys = np.sort(np.asarray([(int(a) << 32) | int(b) for a, b in reapply], np.int64))
    lib().ga_oracle_set_reapply(keys.ctypes.data if len(keys) else None, len(keys))
    aux = np.full((cap_rec, 8), 0xFFFFFFFF, np.uint32) if edits else None
    lib().ga_oracle_set_edit_sink(aux.ctypes.data if edits else None, cap_rec if edits else 0)
    try:
        st = lib().ga_oracle_run(C.byref(R), C.byref(S), refb.ctypes.data, len(refb), C.byref(res), int(threads))
    finally:
        lib().ga_oracle_set_reapply(None, 0)
        lib().ga_oracle_set_edit_sink(None, 0)
    if not decode:
        return {"totals": totals, "counts": counts.reshape(-1, 4), "result": res,
                "arrays": (mod_sess, mod_read, mod_len, mod_so, mod_qo, out_seq, out_qual, counts)}, st
    if st != _abi.GA_OK:
        return None, st
    return decode_result(sessions.n_sessions, totals, mod_sess, mod_read, mod_len, mod_so, mod_qo, out_seq, out_qual,
                         counts[:sessions.n_sessions * 4], edits=aux), st


def digest(result_struct, n_records, session_base=0, tumor_base=0, normal_base=0, n_tumor=0, contig=0, records=False, accumulate=None):
    """Host twin of ga_result_digest over a HOST ga_result (an _abi.GaResult whose arrays the caller keeps alive):
    uint64[4] digest {sum lo, sum hi, records, sum of new lengths} and, with records=True, (keys[n,2], hashes[n,2])."""
    dig = accumulate if accumulate is not None else np.zeros(4, np.uint64)
    keys = np.zeros((max(1, n_records), 2), np.uint64) if records else None
    hashes = np.zeros((max(1, n_records), 2), np.uint64) if records else None
    ids = _abi.GaDigestIds(int(session_base), int(tumor_base), int(normal_base), int(n_tumor), int(contig))
    st = lib().ga_oracle_digest(C.byref(result_struct), int(n_records), C.byref(ids), keys.ctypes.data if records else None,
                                hashes.ctypes.data if records else None, dig.ctypes.data)
    if st != _abi.GA_OK:
        raise ValueError(f"ga_oracle_digest failed with status {st}")
    if records:
        return dig, keys[:n_records], hashes[:n_records]
    return dig


def compare_records(keys_a, hash_a, keys_b, hash_b):
    """Number of records that differ between two (keys, hashes) sets: records present on one side only, plus records
    with the same key and different hashes.  Inputs are [n,2] integer arrays (any 64-bit dtype)."""
    def canon(k, h):
        k = np.ascontiguousarray(k).view(np.uint64).reshape(-1, 2)
        h = np.ascontiguousarray(h).view(np.uint64).reshape(-1, 2)
        o = np.lexsort((k[:, 1], k[:, 0]))
        return k[o], h[o]
    ka, ha = canon(keys_a, hash_a)
    kb, hb = canon(keys_b, hash_b)
    if len(ka) == len(kb) and np.array_equal(ka, kb):
        return int(np.count_nonzero((ha != hb).any(axis=1)))
    sa = {(int(a), int(b)): (int(c), int(d)) for (a, b), (c, d) in zip(ka, ha)}
    sb = {(int(a), int(b)): (int(c), int(d)) for (a, b), (c, d) in zip(kb, hb)}
    return sum(1 for k in sa.keys() | sb.keys() if sa.get(k) != sb.get(k))
