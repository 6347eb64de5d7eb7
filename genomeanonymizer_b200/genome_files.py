"""Genome file readers that feed the engine - thin ctypes wrappers over include/ga_genome_io.h.

The reference opens its inputs through pysam / htslib and variant-extractor, neither of which exists in this image:
    pysam.AlignmentFile(tumor_bam_file) ...     short_read_tumor_normal_anonymizer.py:661-664
    pysam.FastaFile(ref_genome_file)            short_read_tumor_normal_anonymizer.py:915-916
    VariantExtractor(sample_vcf_variants)       short_read_tumor_normal_anonymizer.py:920-921
Here a whole contig of a BAM file is decoded by the C++ reader (all host threads, zlib only) straight into the
structure-of-arrays batch the engine takes; no per-read object is built.  The VCF reader restates what the path
needs from variant-extractor ^4.0.6 (not in /root/reference, not installed - unpinned): contig, 1-based pos, end,
length, REF, ALT and the variant type of plain SNV / DEL / INS / MNV records and of symbolic <DEL> <INS> <DUP> <INV>
<CNV> records (END / SVLEN as the VCF states them); breakend records are rejected loudly rather than guessed.
"""
import ctypes as C
import gzip
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .batch import ReadBatch


class GenomeFileError(OSError):
    pass


class _BamSizes(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("seq16_units", C.c_int64), ("n_cigar", C.c_int64), ("name_bytes", C.c_int64),
                ("max_ref_span", C.c_int32), ("sorted", C.c_int32)]


class _BamDest(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("len_flag", C.c_void_p), ("seq_off16", C.c_void_p), ("cigar_off", C.c_void_p),
                ("ref_end", C.c_void_p), ("name_off", C.c_void_p), ("cigar", C.c_void_p), ("seq4", C.c_void_p),
                ("qual", C.c_void_p), ("names", C.c_void_p), ("seq16_base", C.c_uint32), ("cigar_base", C.c_uint32),
                ("name_base", C.c_uint64)]


IO_EXPORTS = ["ga_io_last_error", "ga_io_set_error", "ga_bam_open", "ga_bam_close", "ga_bam_n_references", "ga_bam_reference_name",
              "ga_bam_reference_length", "ga_bam_n_records", "ga_bam_inflated_bytes", "ga_bam_contig_sizes",
              "ga_bam_pack_contig", "ga_fasta_open", "ga_fasta_close", "ga_fasta_n_references", "ga_fasta_reference_name",
              "ga_fasta_reference_length", "ga_fasta_fetch"]
_BOUND = False


def _io():
    """libga_b200.so with the ga_genome_io.h entry points typed."""
    global _BOUND
    L = _lib.lib()
    if _BOUND:
        return L
    vp = C.c_void_p
    L.ga_io_last_error.restype = C.c_char_p
    L.ga_io_last_error.argtypes = []
    L.ga_bam_open.restype = C.c_int
    L.ga_bam_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.ga_bam_close.restype = None
    L.ga_bam_close.argtypes = [vp]
    L.ga_bam_n_references.restype = C.c_int
    L.ga_bam_n_references.argtypes = [vp]
    L.ga_bam_reference_name.restype = C.c_char_p
    L.ga_bam_reference_name.argtypes = [vp, C.c_int]
    L.ga_bam_reference_length.restype = C.c_int64
    L.ga_bam_reference_length.argtypes = [vp, C.c_int]
    L.ga_bam_n_records.restype = C.c_int64
    L.ga_bam_n_records.argtypes = [vp]
    L.ga_bam_inflated_bytes.restype = C.c_int64
    L.ga_bam_inflated_bytes.argtypes = [vp]
    L.ga_bam_contig_sizes.restype = C.c_int
    L.ga_bam_contig_sizes.argtypes = [vp, C.c_int, C.c_uint32, C.POINTER(_BamSizes)]
    L.ga_bam_pack_contig.restype = C.c_int
    L.ga_bam_pack_contig.argtypes = [vp, C.c_int, C.c_uint32, C.POINTER(_BamDest), C.c_int]
    L.ga_fasta_open.restype = C.c_int
    L.ga_fasta_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.ga_fasta_close.restype = None
    L.ga_fasta_close.argtypes = [vp]
    L.ga_fasta_n_references.restype = C.c_int
    L.ga_fasta_n_references.argtypes = [vp]
    L.ga_fasta_reference_name.restype = C.c_char_p
    L.ga_fasta_reference_name.argtypes = [vp, C.c_int]
    L.ga_fasta_reference_length.restype = C.c_int64
    L.ga_fasta_reference_length.argtypes = [vp, C.c_int]
    L.ga_fasta_fetch.restype = C.c_int64
    L.ga_fasta_fetch.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, vp]
    L.ga_plan_sample.restype = C.c_int
    L.ga_plan_sample.argtypes = [C.c_int64, C.c_int64, vp, vp, vp, vp, vp, C.c_int32, vp, vp, C.c_int64, C.POINTER(vp)]
    L.ga_plan_free.restype = None
    L.ga_plan_free.argtypes = [vp]
    for fn in ("ga_plan_n_sessions", "ga_plan_n_pairs", "ga_plan_n_singles"):
        getattr(L, fn).restype = C.c_int64
        getattr(L, fn).argtypes = [vp]
    L.ga_plan_sessions.restype = None
    L.ga_plan_sessions.argtypes = [vp, vp, vp, vp]
    L.ga_plan_pairs.restype = None
    L.ga_plan_pairs.argtypes = [vp, vp]
    L.ga_plan_singles.restype = None
    L.ga_plan_singles.argtypes = [vp, vp]
    _BOUND = True
    return L


PLAN_EXPORTS = ["ga_plan_sample", "ga_plan_free", "ga_plan_n_sessions", "ga_plan_n_pairs", "ga_plan_n_singles", "ga_plan_sessions",
                "ga_plan_pairs", "ga_plan_singles"]


def plan_contig(cb: "ContigBatch", windows, contig_len: int):
    """driver.plan_sample in native code (include/ga_plan.h) over the packed arrays of one contig: no per-read Python
    object.  Returns a driver.Plan whose pairs / singles are int32 arrays of 5 / 3 columns."""
    from .driver import Plan
    L = _io()
    b = cb.batch
    pos = np.ascontiguousarray(b.pos, np.int32)
    end = np.ascontiguousarray(cb.ref_end, np.int32)
    lf = np.ascontiguousarray(b.len_flag, np.uint32)
    noff = np.ascontiguousarray(cb.name_off, np.int64)
    blob = np.ascontiguousarray(cb.name_blob, np.uint8) if cb.name_blob.size else np.zeros(1, np.uint8)
    wf = np.asarray([w["first"] for w in windows], np.int32)
    wl = np.asarray([w["last"] for w in windows], np.int32)
    h = C.c_void_p()
    _check(L, L.ga_plan_sample(b.n_reads, b.n_tumor, pos.ctypes.data, end.ctypes.data, lf.ctypes.data, noff.ctypes.data, blob.ctypes.data,
                               len(windows), wf.ctypes.data if len(windows) else None, wl.ctypes.data if len(windows) else None,
                               int(contig_len), C.byref(h)))
    try:
        ns, npairs, nsing = int(L.ga_plan_n_sessions(h)), int(L.ga_plan_n_pairs(h)), int(L.ga_plan_n_singles(h))
        first, last, win = np.zeros(ns, np.int32), np.zeros(ns, np.int32), np.zeros(ns, np.int32)
        pairs, singles = np.zeros((npairs, 5), np.int32), np.zeros((nsing, 4), np.int32)
        L.ga_plan_sessions(h, first.ctypes.data, last.ctypes.data, win.ctypes.data)
        L.ga_plan_pairs(h, pairs.ctypes.data)
        L.ga_plan_singles(h, singles.ctypes.data)
    finally:
        L.ga_plan_free(h)
    plan = Plan()
    plan.sessions = [{"first": int(f), "last": int(l), "keep": windows[int(k)].get("keep") if k >= 0 else None,
                      "window": int(k) if k >= 0 else None} for f, l, k in zip(first, last, win)]
    plan.pairs, plan.singles = pairs, singles
    return plan


def _check(L, rc):
    if rc < 0:
        msg = (L.ga_io_last_error() or b"").decode("utf-8", "replace")
        if rc == -3:
            raise ValueError(msg)
        raise GenomeFileError(msg)
    return rc


class BamFile:
    """One BAM file, inflated and indexed per reference (stands where the reference has pysam.AlignmentFile)."""

    def __init__(self, path: str, threads: int = 0):
        self._L = _io()
        self._h = C.c_void_p()
        self.filename = path
        self.threads = int(threads)
        _check(self._L, self._L.ga_bam_open(str(path).encode(), self.threads, C.byref(self._h)))
        n = self._L.ga_bam_n_references(self._h)
        self.references = tuple(self._L.ga_bam_reference_name(self._h, k).decode() for k in range(n))
        self.lengths = tuple(int(self._L.ga_bam_reference_length(self._h, k)) for k in range(n))

    def close(self):
        if self._h:
            self._L.ga_bam_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_records(self) -> int:
        return int(self._L.ga_bam_n_records(self._h))

    @property
    def inflated_bytes(self) -> int:
        return int(self._L.ga_bam_inflated_bytes(self._h))

    def ref_id(self, contig: str) -> int:
        try:
            return self.references.index(contig)
        except ValueError:
            return -1

    def contig_sizes(self, contig: str, flag_exclude: int = 0) -> _BamSizes:
        s = _BamSizes()
        rid = self.ref_id(contig)
        if rid < 0:
            return s
        _check(self._L, self._L.ga_bam_contig_sizes(self._h, rid, flag_exclude, C.byref(s)))
        return s


@dataclass
class ContigBatch:
    """Reads of one contig of a tumor-normal pair: the engine's batch plus what only the host needs."""
    batch: ReadBatch
    ref_end: np.ndarray            # [n] reference_end
    name_blob: np.ndarray          # uint8, names back to back
    name_off: np.ndarray           # [n + 1] int64

    def name(self, r: int) -> str:
        return self.name_blob[self.name_off[r]:self.name_off[r + 1]].tobytes().decode("ascii")

    def read_table(self) -> List[dict]:
        """(name, flag, dataset, pos, end) rows for driver.plan_sample."""
        import gc
        b = self.batch
        blob = self.name_blob.tobytes()
        off = self.name_off.tolist()
        flags = (b.len_flag >> 16).tolist()
        pos, end = b.pos.tolist(), self.ref_end.tolist()
        nt = b.n_tumor
        was_enabled = gc.isenabled()
        gc.disable()                                               # one dict per read: see driver.plan_sample
        try:
            return [{"name": blob[off[k]:off[k + 1]].decode("ascii"), "flag": flags[k], "dataset": 0 if k < nt else 1,
                     "pos": pos[k], "end": end[k]} for k in range(b.n_reads)]
        finally:
            if was_enabled:
                gc.enable()


def _aligned(n_bytes: int, align: int = 64) -> np.ndarray:
    raw = np.zeros(n_bytes + align, np.uint8)
    o = (-raw.ctypes.data) % align
    return raw[o:o + n_bytes]


def pack_tumor_normal(tumor: BamFile, normal: BamFile, contig: str, contig_id: int = 0, flag_exclude: int = 0,
                      qualities: bool = True) -> ContigBatch:
    """Every read of `contig` of both files as one batch (tumor reads first, file order inside each dataset)."""
    L = tumor._L
    sizes = [f.contig_sizes(contig, flag_exclude) for f in (tumor, normal)]
    for f, s in zip((tumor, normal), sizes):
        if s.n_reads and not s.sorted:
            raise ValueError(f"{f.filename}: records of {contig} are not in coordinate order")
    n = sum(int(s.n_reads) for s in sizes)
    units = sum(int(s.seq16_units) for s in sizes)
    n_cig = sum(int(s.n_cigar) for s in sizes)
    n_name = sum(int(s.name_bytes) for s in sizes)
    pos = np.zeros(n, np.int32)
    len_flag = np.zeros(n, np.uint32)
    seq_off16 = np.zeros(n, np.uint32)
    cigar_off = np.zeros(n + 1, np.uint32)
    ref_end = np.zeros(n, np.int32)
    name_off = np.zeros(n + 1, np.uint64)
    cigar = np.zeros(max(1, n_cig), np.uint32)
    seq4 = _aligned(16 * units)
    qual = _aligned(32 * units) if qualities else None
    names = np.zeros(max(1, n_name), np.uint8)
    r0 = u0 = c0 = m0 = 0
    for f, s in zip((tumor, normal), sizes):
        rid = f.ref_id(contig)
        if rid >= 0 and s.n_reads:
            d = _BamDest()
            d.pos = pos[r0:].ctypes.data
            d.len_flag = len_flag[r0:].ctypes.data
            d.seq_off16 = seq_off16[r0:].ctypes.data
            d.cigar_off = cigar_off[r0:].ctypes.data
            d.ref_end = ref_end[r0:].ctypes.data
            d.name_off = name_off[r0:].ctypes.data
            d.cigar, d.seq4, d.names = cigar.ctypes.data, seq4.ctypes.data, names.ctypes.data
            d.qual = qual.ctypes.data if qual is not None else None
            d.seq16_base, d.cigar_base, d.name_base = u0, c0, m0
            _check(L, L.ga_bam_pack_contig(f._h, rid, flag_exclude, C.byref(d), f.threads))
        r0 += int(s.n_reads); u0 += int(s.seq16_units); c0 += int(s.n_cigar); m0 += int(s.name_bytes)
        cigar_off[r0] = c0
        name_off[r0] = m0
    b = ReadBatch(n_tumor=int(sizes[0].n_reads), pos=pos, len_flag=len_flag, seq_off16=seq_off16, cigar_off=cigar_off,
                  cigar=cigar[:n_cig], seq4=seq4, qual=qual, max_ref_span=max(int(s.max_ref_span) for s in sizes),
                  contig_id=contig_id)
    return ContigBatch(batch=b, ref_end=ref_end, name_blob=names[:n_name], name_off=name_off.astype(np.int64))


class FastaFile:
    """FASTA, plain text or gzip / bgzip (stands where the reference has pysam.FastaFile: .references, .lengths, .fetch)."""

    def __init__(self, path: str):
        self._L = _io()
        self._h = C.c_void_p()
        self.filename = path
        _check(self._L, self._L.ga_fasta_open(str(path).encode(), C.byref(self._h)))
        n = self._L.ga_fasta_n_references(self._h)
        self.references = tuple(self._L.ga_fasta_reference_name(self._h, k).decode() for k in range(n))
        self.lengths = tuple(int(self._L.ga_fasta_reference_length(self._h, k)) for k in range(n))

    def close(self):
        if self._h:
            self._L.ga_fasta_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_reference_length(self, contig: str) -> int:
        return self.lengths[self.references.index(contig)]

    def fetch_bytes(self, contig: str, start: int = 0, end: Optional[int] = None) -> np.ndarray:
        if contig not in self.references:
            raise KeyError(f"sequence '{contig}' not present in {self.filename}")
        k = self.references.index(contig)
        end = self.lengths[k] if end is None else min(int(end), self.lengths[k])
        start = max(0, int(start))
        out = np.zeros(max(0, end - start), np.uint8)
        if out.size:
            got = _check(self._L, self._L.ga_fasta_fetch(self._h, k, start, end, out.ctypes.data))
            out = out[:got]
        return out

    def fetch(self, contig: str, start: int = 0, end: Optional[int] = None) -> str:
        return self.fetch_bytes(contig, start, end).tobytes().decode("ascii")


# ---------------------------------------------------------------------------------------------------- VCF
@dataclass
class VcfVariant:
    """The fields of variant_extractor's VariantRecord this path reads (short_read_tumor_normal_anonymizer.py:80-128,
    variants.py:59-62): 1-based pos / end, length, REF, ALT, variant type name."""
    contig: str
    pos: int
    end: int
    length: int
    ref: str
    alt: str
    variant_type: str


_SYMBOLIC_TYPES = {"DEL": "DEL", "INS": "INS", "DUP": "DUP", "INV": "INV", "CNV": "CNV"}


def classify_vcf_alleles(pos: int, ref: str, alt: str, info: Optional[Dict[str, str]] = None) -> Tuple[int, int, str]:
    """(end, length, type) of one REF/ALT pair.
    Plain alleles - SNV: end = pos, length 1.  DEL: length = len(REF) - len(ALT), end = pos + length.  INS: length =
    len(ALT) - len(REF), end = pos + 1.  Same-length multi-base records are kept as SNV-typed records of that length
    (window geometry only; they can never equal a pileup SNV key, variants.py:83-96).
    Symbolic alleles <DEL> <INS> <DUP> <INV> <CNV> (with :subtype) - end = INFO END (INS: pos + 1 when absent), length =
    |INFO SVLEN| or end - pos.  Breakend (BND) and '*' alleles are not supported: variant-extractor pairs breakends into
    TRA / INV / DEL / DUP records by rules this reader does not restate."""
    info = info or {}
    if alt.startswith("<") and alt.endswith(">"):
        vt = _SYMBOLIC_TYPES.get(alt[1:-1].split(":")[0])
        if vt is None:
            raise ValueError(f"symbolic ALT allele {alt!r} at position {pos} is not supported by this reader")
        svlen = abs(int(info["SVLEN"].split(",")[0])) if "SVLEN" in info else None
        if "END" in info:
            end = int(info["END"])
        elif vt == "INS":
            end = pos + 1
        elif svlen is not None:
            end = pos + svlen
        else:
            raise ValueError(f"{alt} record at position {pos} has neither END nor SVLEN")
        return end, (svlen if svlen is not None else end - pos), vt
    if "[" in alt or "]" in alt or alt == "*" or alt == "." or alt.startswith("."):
        raise ValueError(f"breakend / missing ALT allele {alt!r} at position {pos} is not supported by this reader")
    if len(ref) == len(alt):
        return (pos, 1, "SNV") if len(ref) == 1 else (pos + len(ref) - 1, len(ref), "SNV")
    if len(ref) > len(alt):
        ln = len(ref) - len(alt)
        return pos + ln, ln, "DEL"
    return pos + 1, len(alt) - len(ref), "INS"


def read_vcf(path: str) -> List[VcfVariant]:
    """Records of a (plain or gzip / bgzip compressed) VCF, one VcfVariant per ALT allele, in file order."""
    with open(path, "rb") as fh:
        magic = fh.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    out: List[VcfVariant] = []
    with opener(path, "rt") as fh:
        for line in fh:
            if not line.strip() or line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 5:
                raise ValueError(f"{path}: malformed VCF record: {line!r}")
            pos = int(f[1])
            info = dict(kv.partition("=")[::2] for kv in f[7].split(";")) if len(f) > 7 and f[7] not in (".", "") else {}
            for alt in f[4].split(","):
                end, length, vt = classify_vcf_alleles(pos, f[3], alt, info)
                out.append(VcfVariant(f[0], pos, end, length, f[3], alt, vt))
    return out


def windows_by_contig(variants: List[VcfVariant], contig_order: Dict[str, int]) -> Dict[str, List[dict]]:
    """get_windows (short_read_tumor_normal_anonymizer.py:71-131) without the breakend branches: SNV -> one window
    [pos - 1000, pos + 1001); INV -> one window around both ends when they are closer than a window, else one around
    each end; every other type -> [pos - 1000, end + 1001) below 100 kb, else one window around each end.  Sorted by
    (contig, first, last), each with the 0-based key of the variant to keep (variants.py:59-62)."""
    from .driver import WINDOW_HALF, window_of_variant
    rows = []
    for v in variants:
        if v.contig not in contig_order:
            raise KeyError(f"VCF contig {v.contig!r} is not in the reference genome")
        if v.variant_type == "SNV":
            spans = [window_of_variant(v.pos, v.pos)]
        elif v.variant_type == "INV":
            near = v.pos + WINDOW_HALF > v.end - WINDOW_HALF
            spans = [window_of_variant(v.pos, v.end)] if near else [window_of_variant(v.pos, v.pos), window_of_variant(v.end, v.end)]
        elif v.length < 100_000:
            spans = [window_of_variant(v.pos, v.end)]
        else:
            spans = [window_of_variant(v.pos, v.pos), window_of_variant(v.end, v.end)]
        for first, last in spans:
            rows.append((contig_order[v.contig], first, last, v))
    rows.sort(key=lambda t: t[:3])
    out: Dict[str, List[dict]] = {}
    for _, first, last, v in rows:
        out.setdefault(v.contig, []).append({"first": first, "last": last,
                                             "keep": {"type": v.variant_type, "pos": v.pos - 1, "end": v.end - 1,
                                                      "length": v.length, "allele": v.alt}})
    return out
