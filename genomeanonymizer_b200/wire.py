"""Wire form of a read batch (include/ga_wire.h): what the host hands to ga_run_wire instead of the plain arrays.

`pack_wire` calls the library's host packer (csrc/ga_wire.cu, plain C++ threads - no device work); the qualities stay
the sparse records of ga_reads.  The packer runs where the reference's own packer runs: when the BAM records are
decoded, outside the masking call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _abi, _lib
from .batch import ReadBatch


@dataclass
class WireBatch:
    n_reads: int
    n_tumor: int
    n_blocks: int
    n_tumor_blocks: int
    blob: np.ndarray                 # uint8, 16-byte aligned
    dir: np.ndarray                  # structured view of ga_wire_dir, n_blocks + 1 entries
    qual: Optional[np.ndarray]
    qual_reads: Optional[np.ndarray]
    qual_off16: Optional[np.ndarray]
    qual_units: int
    max_ref_span: int
    contig_id: int

    @property
    def nbytes(self) -> int:
        return int(self.blob.nbytes + self.dir.nbytes + sum(a.nbytes for a in (self.qual, self.qual_reads, self.qual_off16) if a is not None))


DIR_DTYPE = np.dtype([("byte", "<u8"), ("read", "<u4"), ("unit", "<u4"), ("ops", "<u4"), ("pos", "<i4"), ("reserved", "<u4", (2,))])
assert DIR_DTYPE.itemsize == C.sizeof(_abi.GaWireDir)


def aligned_bytes(n: int, align: int = 64) -> np.ndarray:
    raw = np.zeros(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n]


def sparse_qualities(batch: ReadBatch):
    """(qual, qual_reads, qual_off16, units) with one contiguous record per read that has an I or D op (the only reads
    whose qualities can change, anonymizer_methods.py:170-176 vs 183-195); a batch that is sparse already is returned
    as it is."""
    if batch.qual is None:
        return None, None, None, 0
    if batch.qual_reads is not None:
        qr = np.asarray(batch.qual_reads, np.int32)
        units = 0
        if len(qr):
            L = int(batch.len_flag[qr[-1]] & 0xFFFF)
            units = int(batch.qual_off16[-1]) + max(1, (L + 31) // 32)
        return batch.qual, qr, np.asarray(batch.qual_off16, np.uint32), units
    ops = batch.cigar & 15
    is_id = ((ops == 1) | (ops == 2)).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(is_id)])
    has = (cum[batch.cigar_off[1:].astype(np.int64)] - cum[batch.cigar_off[:-1].astype(np.int64)]) > 0
    qr = np.nonzero(has)[0].astype(np.int32)
    L = (batch.len_flag[qr] & 0xFFFF).astype(np.int64)
    u = np.maximum(1, (L + 31) // 32)
    off = np.concatenate([[0], np.cumsum(u)])
    qual = np.zeros(int(off[-1]) * 32 + 64, np.uint8)
    for k, r in enumerate(qr):
        o = int(batch.seq_off16[r]) * 32
        qual[int(off[k]) * 32:int(off[k]) * 32 + int(L[k])] = batch.qual[o:o + int(L[k])]
    return qual, qr, off[:-1].astype(np.uint32), int(off[-1])


def pack_wire(batch: ReadBatch, threads: int = 0) -> WireBatch:
    L = _lib.lib()
    keep = []
    R = batch.as_struct(keep)
    nb, ntb, nbytes, span = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int32(0)
    st = L.ga_wire_pack_sizes(C.byref(R), C.byref(nb), C.byref(ntb), C.byref(nbytes), C.byref(span))
    _abi.raise_for_status(st, "ga_wire_pack_sizes: reads of a dataset must be in coordinate order")
    blob = aligned_bytes(max(16, int(nbytes.value)))
    d = np.zeros(int(nb.value) + 1, DIR_DTYPE)
    st = L.ga_wire_pack(C.byref(R), blob.ctypes.data, int(nbytes.value), d.ctypes.data, int(nb.value), int(threads))
    _abi.raise_for_status(st, "ga_wire_pack")
    qual, qr, qo, units = sparse_qualities(batch)
    return WireBatch(batch.n_reads, batch.n_tumor, int(nb.value), int(ntb.value), blob[:int(nbytes.value)] if nbytes.value else blob[:0], d,
                     qual, qr, qo, units, int(span.value), int(batch.contig_id))
