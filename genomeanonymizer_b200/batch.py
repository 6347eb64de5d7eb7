"""Host-side packer: reads / windows -> the structure-of-arrays batch of include/ga_b200.h.

Replaces what crosses the reference's method boundary as pysam objects (query_sequence,
query_qualities, cigarstring, reference_start, flag) with packed arrays:
4-bit base codes (low nibble first), phred bytes, BAM-encoded CIGAR words, positions, flags.
Arrays are numpy on the host; `to_device()` gives torch tensors (pinned staging optional).
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _abi

_CIG = re.compile(r"(\d+)([MIDNSHP=XB])")
_OPCODE = {c: i for i, c in enumerate(_abi.CIGAR_OPS)}
_ASC2CODE = np.full(256, 15, dtype=np.uint8)
for _i, _ch in enumerate(_abi.CODE2ASC):
    _ASC2CODE[ord(_ch)] = _i
    if _ch.isalpha():
        _ASC2CODE[ord(_ch.lower())] = _i


def encode_bases(seq: str) -> np.ndarray:
    return _ASC2CODE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]


def decode_bases(codes: np.ndarray) -> str:
    return "".join(_abi.CODE2ASC[int(c)] for c in codes)


def parse_cigar(cigar: str) -> np.ndarray:
    return np.array([(int(n) << 4) | _OPCODE[op] for n, op in _CIG.findall(cigar)], dtype=np.uint32)


def pack_nibbles(codes: np.ndarray, cap_bases: int) -> np.ndarray:
    buf = np.zeros(cap_bases, dtype=np.uint8)
    buf[:len(codes)] = codes
    return (buf[0::2] | (buf[1::2] << 4)).astype(np.uint8)


def unpack_nibbles(rec: np.ndarray, n: int) -> np.ndarray:
    out = np.empty(len(rec) * 2, dtype=np.uint8)
    out[0::2] = rec & 15
    out[1::2] = rec >> 4
    return out[:n]


@dataclass
class ReadBatch:
    """Host arrays of one contig's session reads (tumor first, then normal; each coordinate sorted)."""
    n_tumor: int
    pos: np.ndarray
    len_flag: np.ndarray
    seq_off16: np.ndarray
    cigar_off: np.ndarray
    cigar: np.ndarray
    seq4: np.ndarray
    qual: Optional[np.ndarray]
    max_ref_span: int = 0
    contig_id: int = 0
    qual_reads: Optional[np.ndarray] = None
    qual_off16: Optional[np.ndarray] = None
    names: Optional[List[str]] = None          # host only; never uploaded

    @property
    def n_reads(self) -> int:
        return int(self.pos.shape[0])

    def read_len(self, r: int) -> int:
        return int(self.len_flag[r] & 0xFFFF)

    def flag(self, r: int) -> int:
        return int(self.len_flag[r] >> 16)

    def sequence_codes(self, r: int) -> np.ndarray:
        o = int(self.seq_off16[r]) * 16
        n = self.read_len(r)
        return unpack_nibbles(self.seq4[o:o + (n + 1) // 2], n)

    def qualities(self, r: int) -> np.ndarray:
        o = int(self.seq_off16[r]) * 32
        return self.qual[o:o + self.read_len(r)]

    def as_struct(self, keepalive: list) -> _abi.GaReads:
        def p(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a)
            keepalive.append(a)
            return a.ctypes.data
        s = _abi.GaReads()
        s.n_reads, s.n_tumor = self.n_reads, self.n_tumor
        s.pos, s.len_flag, s.seq_off16 = p(self.pos), p(self.len_flag), p(self.seq_off16)
        s.cigar_off, s.cigar, s.seq4, s.qual = p(self.cigar_off), p(self.cigar), p(self.seq4), p(self.qual)
        s.seq4_bytes = int(self.seq4.shape[0])
        s.n_qual = 0 if self.qual_reads is None else int(self.qual_reads.shape[0])
        s.qual_reads, s.qual_off16 = p(self.qual_reads), p(self.qual_off16)
        s.max_ref_span, s.contig_id = int(self.max_ref_span), int(self.contig_id)
        return s


@dataclass
class SessionTable:
    first: np.ndarray
    last: np.ndarray
    keep_type: np.ndarray
    keep_pos: np.ndarray
    keep_end: np.ndarray
    keep_len: np.ndarray
    keep_allele_off: np.ndarray
    keep_alleles: np.ndarray

    @property
    def n_sessions(self) -> int:
        return int(self.first.shape[0])

    def as_struct(self, keepalive: list) -> _abi.GaSessions:
        def p(a):
            a = np.ascontiguousarray(a)
            keepalive.append(a)
            return a.ctypes.data
        s = _abi.GaSessions()
        s.n_sessions = self.n_sessions
        s.first, s.last = p(self.first), p(self.last)
        s.keep_type, s.keep_pos, s.keep_end, s.keep_len = p(self.keep_type), p(self.keep_pos), p(self.keep_end), p(self.keep_len)
        s.keep_allele_off, s.keep_alleles = p(self.keep_allele_off), p(self.keep_alleles)
        return s


def ref_span(cigar_words: np.ndarray) -> int:
    ops = cigar_words & 15
    lens = cigar_words >> 4
    return int(lens[(ops == 0) | (ops == 2) | (ops == 3) | (ops == 7) | (ops == 8)].sum())


def pack_reads(reads: Sequence[dict], contig_id: int = 0, sparse_qual: bool = False) -> ReadBatch:
    """reads: dicts with name, flag, pos, cigar (string), seq (ASCII), qual (ints), dataset (0 T / 1 N).
    Order inside each dataset is preserved (must be coordinate order, as in the BAM)."""
    order = [i for i, r in enumerate(reads) if r["dataset"] == 0] + [i for i, r in enumerate(reads) if r["dataset"] == 1]
    n = len(order)
    n_tumor = sum(1 for r in reads if r["dataset"] == 0)
    pos = np.zeros(n, np.int32)
    len_flag = np.zeros(n, np.uint32)
    seq_off16 = np.zeros(n, np.uint32)
    cigar_off = np.zeros(n + 1, np.uint32)
    cig_chunks, seq_chunks, qual_chunks, names = [], [], [], []
    qual_reads, qual_off16 = [], []
    off16 = 0
    qoff16 = 0
    max_span = 0
    for k, i in enumerate(order):
        r = reads[i]
        L = len(r["seq"])
        if L > 0xFFFF:
            raise ValueError("reads longer than 65535 bases are not supported")
        if len(r["qual"]) != L:
            raise ValueError("Length of the qualities does not match the length of the sequence")
        cw = parse_cigar(r["cigar"])
        qlen = int((cw >> 4)[np.isin(cw & 15, (0, 1, 4, 7, 8))].sum())
        if qlen != L and not (len(cw) == 0 and int(r["flag"]) & 0x4):      # an unmapped read has no CIGAR
            raise ValueError(f"CIGAR {r['cigar']} consumes {qlen} query bases, sequence has {L}")
        cap = max(32, (L + 31) // 32 * 32)
        pos[k] = r["pos"]
        len_flag[k] = (int(r["flag"]) << 16) | L
        seq_off16[k] = off16
        cigar_off[k + 1] = cigar_off[k] + len(cw)
        cig_chunks.append(cw)
        seq_chunks.append(pack_nibbles(encode_bases(r["seq"]), cap))
        q = np.zeros(cap, np.uint8)
        q[:L] = np.asarray(r["qual"], dtype=np.uint8)
        has_indel = bool(np.isin(cw & 15, (1, 2)).any())
        if not sparse_qual:
            qual_chunks.append(q)
        elif has_indel:
            qual_reads.append(k)
            qual_off16.append(qoff16)
            qual_chunks.append(q)
            qoff16 += cap // 32
        off16 += cap // 32
        max_span = max(max_span, ref_span(cw))
        names.append(r["name"])
    cat = lambda ch, dt: (np.concatenate(ch).astype(dt) if ch else np.zeros(0, dt))
    b = ReadBatch(n_tumor=n_tumor, pos=pos, len_flag=len_flag, seq_off16=seq_off16, cigar_off=cigar_off,
                  cigar=cat(cig_chunks, np.uint32), seq4=cat(seq_chunks, np.uint8), qual=cat(qual_chunks, np.uint8),
                  max_ref_span=max_span, contig_id=contig_id, names=names)
    if sparse_qual:
        b.qual_reads = np.asarray(qual_reads, np.int32)
        b.qual_off16 = np.asarray(qual_off16, np.uint32)
    for lo, hi in ((0, n_tumor), (n_tumor, n)):
        if hi - lo > 1 and np.any(np.diff(pos[lo:hi].astype(np.int64)) < 0):
            raise ValueError("reads of a dataset must be in coordinate order")
    return b


def pack_sessions(windows: Sequence[dict]) -> SessionTable:
    """windows: dicts with first, last, keep (None or dict type/pos/end/length/allele) in genome order."""
    S = len(windows)
    first = np.array([w["first"] for w in windows], np.int32)
    last = np.array([w["last"] for w in windows], np.int32)
    if S > 1 and (np.any(np.diff(first.astype(np.int64)) < 0)):
        raise ValueError("sessions must be sorted by (first, last)")
    kt = np.zeros(S, np.int32); kp = np.zeros(S, np.int32); ke = np.zeros(S, np.int32); kl = np.zeros(S, np.int32)
    off = np.zeros(S + 1, np.uint32)
    alle = bytearray()
    for i, w in enumerate(windows):
        k = w.get("keep")
        if k is not None:
            kt[i] = _abi.VT_BY_NAME.get(k["type"], 99) if isinstance(k["type"], str) else int(k["type"])
            kp[i], ke[i], kl[i] = k["pos"], k["end"], k["length"]
            alle += k["allele"].encode("ascii")
        off[i + 1] = len(alle)
    return SessionTable(first, last, kt, kp, ke, kl, off, np.frombuffer(bytes(alle) + b"\0", dtype=np.uint8).copy())


@dataclass
class MaskResult:
    """Host view of a ga_result: modified (session, read) pairs."""
    records: Dict[tuple, dict] = field(default_factory=dict)   # (session, read) -> {"seq": codes, "qual": bytes|None}
    sess_counts: Optional[np.ndarray] = None                   # [S,4]
    totals: Optional[dict] = None

    def counts8(self) -> np.ndarray:
        """Per-session counters in the reference's 8-column VariantType layout (SR.py:183-204)."""
        out = np.zeros((self.sess_counts.shape[0], 8), np.int64)
        out[:, 0:3] = self.sess_counts[:, 0:3]
        return out


def parse_edits(aux) -> Optional[list]:
    """The 8 words of ga_record_edits (include/ga_b200.h) as [(in_read_pos, reference position, length, is_insertion)] in
    application order; None when the description was not kept."""
    if int(aux[6]) == 0xFFFFFFFF:
        return None
    s32 = lambda w: int(w) - (1 << 32) if int(w) >> 31 else int(w)
    return [(s32(aux[3 * k]), s32(aux[3 * k + 1]), int(aux[3 * k + 2]) & 0x7FFFFFFF, bool(int(aux[3 * k + 2]) >> 31)) for k in range(int(aux[6]) & 0xFF)]


def decode_result(n_sessions, totals: _abi.GaTotals, mod_session, mod_read, mod_len, mod_seq_off16, mod_qual_off16,
                  out_seq4, out_qual, sess_counts, edits=None) -> MaskResult:
    """edits: [n, 8] words of ga_record_edits - every indel-masked record then carries "edits" (parse_edits)."""
    res = MaskResult()
    n = int(totals.n_modified)
    for k in range(n):
        r = (int(mod_session[k]), int(mod_read[k])); L = int(mod_len[k])
        so = int(mod_seq_off16[k]) * 16
        codes = unpack_nibbles(np.asarray(out_seq4[so:so + (L + 1) // 2]), L)
        q = None
        if int(mod_qual_off16[k]) != 0xFFFFFFFF:
            qo = int(mod_qual_off16[k]) * 32
            q = np.asarray(out_qual[qo:qo + L]).copy()
        if r in res.records:
            raise ValueError(f"(session, read) {r} emitted twice")
        res.records[r] = {"seq": codes, "qual": q}
        if edits is not None and q is not None:
            res.records[r]["edits"] = parse_edits(edits[k])
    res.sess_counts = np.asarray(sess_counts).reshape(n_sessions, 4).copy()
    res.totals = {"n_modified": n, "seq16_used": int(totals.seq16_used), "qual16_used": int(totals.qual16_used),
                  "session_reads": int(totals.session_reads), "session_bases": int(totals.session_bases),
                  "indel_records": int(totals.indel_records), "masked": [int(x) for x in totals.masked],
                  "error": int(totals.error), "error_detail": int(totals.error_detail)}
    return res
