"""Host-side mirror of the reference's anonymization-method interface for the masking hot path.

Drop-in for `CompleteGermlineAnonymizer` (reference src/GenomeAnonymizer/anonymizer_methods.py:422-556):
same constructor, `anonymize(validated_source_variant, tumor_normal_pileup, ref_genome, stats_recorder=None)`
generator and `reset()`, fed with the same duck-typed pysam objects (pileup columns, alignments, FASTA) and
yielding two-element lists of `AnonymizedRead`-compatible objects in the reference's order, so
`anonymize_window` (short_read_tumor_normal_anonymizer.py:279-372) can use it unchanged.

What differs is where the work happens: the pileup stream is only used to collect the session's reads;
discovery, germline resolution and masking run in the CUDA engine (libga_b200.so) on one packed batch.
Nothing in `anonymize()` masks on the CPU; without the CUDA library / a GPU the class raises.  The one piece of the
reference's read object that edits bases on the host is kept because the reference's DRIVER calls it after the path:
`AnonymizedRead.mask_or_anonymize_left_over_variants`, which the driver's writers invoke a second time on a read that was
parked unpaired and met again (quirk Q12 of DESIGN.md) - the engine hands over what it applied (ga_record_edits) and the
object repeats it exactly as the reference's object would.
"""
from __future__ import annotations

import array
import bisect
import enum
from typing import Dict, Generator, List, Optional

import numpy as np

from . import _abi
from . import batch as B

DATASET_IDX_TUMOR = 0        # variation_classifier.py:13
DATASET_IDX_NORMAL = 1       # variation_classifier.py:14
PAIR_1_IDX = 0
PAIR_2_IDX = 1

_COMPLEMENT = np.zeros(256, dtype=np.uint8)
for _a, _b in zip("ACGTN", "TGCAN"):                       # anonymizer_methods.py:22 (any other base: TypeError there)
    _COMPLEMENT[ord(_a)] = ord(_b)


class VariantType(enum.Enum):
    """variant_extractor.variants.VariantType values (statistics column = value - 1, SR.py:198-204)."""
    SNV = 1
    DEL = 2
    INS = 3
    DUP = 4
    INV = 5
    CNV = 6
    TRA = 7
    SGL = 8


class MaskedVariant:
    """What `stats_recorder.count_variant` receives: only `.variant_type` is read (SR.py:198-204)."""

    def __init__(self, variant_type, length: int = 0, ref_allele: str = ""):
        self.variant_type = variant_type
        self.length = length
        self.ref_allele = ref_allele


def generate_anonymized_read(name: str, sequence: str, quality: str) -> str:
    return f"@{name}\n{sequence}\n+\n{quality}"              # anonymizer_methods.py:57-58


class AnonymizedRead:
    """Attribute-compatible with the reference's AnonymizedRead (anonymizer_methods.py:84-288) for everything
    the session driver and writer touch (SR.py:134-165, 304-360)."""

    def __init__(self, query_name, is_read1, is_read2, is_reverse, dataset_idx, sequence: np.ndarray, forward_qualities):
        self.query_name = query_name
        self.is_read1 = is_read1
        self.is_read2 = is_read2
        self.is_reverse = is_reverse
        self.dataset_idx = dataset_idx
        self.anonymized_sequence_array = sequence                 # uint8 ASCII, alignment orientation
        self.anonymized_qualities_array = forward_qualities       # original-read orientation (AM.py:95)
        self.is_supplementary = False
        self.has_supplementary = False
        self.supplementary_hashes = set()
        self.n_supplementaries = 0
        self.left_over_variants_to_mask: list = []
        self.has_left_overs_to_mask = False

    def get_pair_idx(self):
        if self.is_read1:
            return PAIR_1_IDX
        if self.is_read2:
            return PAIR_2_IDX

    def anonymized_read_is_complete(self) -> bool:
        return True                                               # no supplementary bookkeeping on this path

    def reverse_complement(self):
        self.anonymized_sequence_array = np.flip(_COMPLEMENT[self.anonymized_sequence_array])
        self.anonymized_qualities_array = list(reversed(self.anonymized_qualities_array))

    def get_anonymized_fastq_record(self) -> str:
        if self.is_reverse:
            self.reverse_complement()
        name = f"{self.query_name}/{PAIR_1_IDX + 1}" if self.is_read1 else f"{self.query_name}/{PAIR_2_IDX + 1}"
        seq = bytes(np.asarray(self.anonymized_sequence_array, dtype=np.uint8)).decode("ascii")
        qual = "".join(chr(int(x) + 33) for x in self.anonymized_qualities_array)
        return generate_anonymized_read(name, seq, qual)

    def update_anonymized_read_from_other(self, other):
        """anonymizer_methods.py:281-287 - the list of an already masked read is never emptied, so meeting the read again
        switches the flag back on and the next writer masks its indels a second time (quirk Q12)."""
        if other.has_left_overs_to_mask:
            self.left_over_variants_to_mask.extend(other.left_over_variants_to_mask)
        if self.left_over_variants_to_mask:
            self.has_left_overs_to_mask = True

    def mask_or_anonymize_left_over_variants(self):
        """anonymizer_methods.py:254-270 over mask_or_modify_indel (:178-203), for the writers of the reference's driver
        (SR.py:354-357, 401-404, 615-616).  The first application happened on the device."""
        self.left_over_variants_to_mask.sort(key=lambda e: e[1].variant_type.value)
        seq = [int(x) for x in self.anonymized_sequence_array]
        qual = [int(x) for x in self.anonymized_qualities_array]
        for at, v in self.left_over_variants_to_mask:
            if v.variant_type == VariantType.INS:                 # the inserted bases go
                seq, qual = seq[:at] + seq[at + v.length:], qual[:at] + qual[at + v.length:]
            elif v.variant_type == VariantType.DEL:               # the deleted reference bases come back, floor(mean quality)
                fill = int(np.mean(qual))
                seq, qual = seq[:at] + [ord(c) for c in v.ref_allele] + seq[at:], qual[:at] + [fill] * v.length + qual[at:]
        self.anonymized_sequence_array = np.asarray(seq, dtype=np.uint8)
        self.anonymized_qualities_array = array.array("B", qual)
        self.has_left_overs_to_mask = False


def anonymized_read_pair_is_writeable(r1, r2) -> bool:
    return r1 is not None and r2 is not None and r1.anonymized_read_is_complete() and r2.anonymized_read_is_complete()


def _variant_type_value(vt) -> int:
    if vt is None:
        return 0
    name = getattr(vt, "name", None)
    if name in _abi.VT_BY_NAME:
        return _abi.VT_BY_NAME[name]
    return 99                                                     # a type this path never calls: never equal


def keep_from_variant(v, shift: int) -> Optional[dict]:
    """`variant_to_keep` (AM.py:546-547) as a session-table row; coordinates shifted into the uploaded region."""
    if v is None:
        return None
    return {"type": _variant_type_value(v.variant_type), "pos": int(v.pos) - shift, "end": int(v.end) - shift,
            "length": int(v.length), "allele": str(v.allele)}


class SessionReads:
    """Collects the reads of one pileup stream in first-appearance order (per dataset = file order)."""

    def __init__(self):
        self.alns = [[], []]
        self.seen = [set(), set()]
        self.appearance: Dict[str, int] = {}
        self.normal_cols: List[int] = []
        self.contig = None
        self.first_col = None
        self.last_col = None

    def consume(self, tumor_normal_pileup):
        for pair in tumor_normal_pileup:
            for ds, col in enumerate(pair):
                if col is None:
                    continue
                pos = col.reference_pos
                if self.contig is None:
                    self.contig = col.reference_name
                self.first_col = pos if self.first_col is None else min(self.first_col, pos)
                self.last_col = pos if self.last_col is None else max(self.last_col, pos)
                if ds == DATASET_IDX_NORMAL:
                    self.normal_cols.append(pos)
                for pr in col.pileups:
                    aln = pr.alignment
                    # identity = (name, mate), as in the reference (seen_read_alns, VC.py:30-31,201-207; registry
                    # AM.py:323-335): the first alignment of a (name, mate) is the read
                    key = (aln.query_name, bool(aln.is_read1))
                    if key in self.seen[ds]:
                        continue
                    self.seen[ds].add(key)
                    self.alns[ds].append(aln)
                    self.appearance.setdefault(aln.query_name, len(self.appearance))


class B200GermlineAnonymizer:
    """`CompleteGermlineAnonymizer` on the B200 engine.  Picklable (the engine handle is created lazily per
    process, as the reference pickles its anonymizer into pool workers, SR.py:953-959)."""

    def __init__(self, device: int = 0, engine=None):
        """engine: an engine.Engine that already exists (e.g. one shared by several samples of a process); by default
        the engine is created on first use, on CUDA device `device`."""
        self.device = device
        self.anonymized_reads: Dict[str, list] = dict()
        self._engine = engine

    def __getstate__(self):
        return {"device": self.device}

    def __setstate__(self, st):
        self.device = st["device"]
        self.anonymized_reads = dict()
        self._engine = None

    def reset(self):
        self.anonymized_reads = dict()

    def _get_engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self.device)                    # raises without CUDA: there is no CPU fallback
        return self._engine

    # -- the reference's call shape (AM.py:431-432, call site SR.py:292-293)
    def anonymize(self, validated_source_variant, tumor_normal_pileup, ref_genome,
                  stats_recorder=None) -> Generator[list, None, None]:
        reads = SessionReads()
        reads.consume(tumor_normal_pileup)
        if reads.first_col is None:
            self.reset()
            return
        alns = reads.alns[0] + reads.alns[1]
        n_t = len(reads.alns[0])
        # region of the contig the session can touch: every covered column (truncate=False, PIO.pyx:12-17)
        lo = min(a.reference_start for a in alns)
        hi = max(a.reference_end for a in alns)
        ref = ref_genome.fetch(reads.contig, lo, hi)               # upper-cased on the device (VC.py:89,194)
        rows = []
        for ds, lst in enumerate(reads.alns):
            for a in lst:
                if not (a.is_read1 or a.is_read2):
                    raise TypeError("read with neither READ1 nor READ2 flag")          # AM.py:119-123 returns None -> TypeError
                rows.append({"name": a.query_name, "flag": int(a.flag), "pos": int(a.reference_start) - lo,
                             "cigar": a.cigarstring, "seq": a.query_sequence, "qual": list(a.query_qualities), "dataset": ds})
        batch = B.pack_reads(rows, sparse_qual=True)
        sessions = B.pack_sessions([{"first": reads.first_col - lo, "last": reads.last_col + 1 - lo,
                                     "keep": keep_from_variant(validated_source_variant, lo)}])
        eng = self._get_engine()
        eng.upload_reference(0, ref)
        result = eng.run(batch, sessions, edits=True)
        if stats_recorder is not None:
            for col, vt in enumerate((VariantType.SNV, VariantType.DEL, VariantType.INS)):
                for _ in range(int(result.sess_counts[0, col])):
                    stats_recorder.count_variant(MaskedVariant(vt))
        # ---- rebuild the reference's objects
        pairs: Dict[str, list] = {}
        max_end: Dict[str, int] = {}
        for r, a in enumerate(alns):
            ds = 0 if r < n_t else 1
            rec = result.records.get((0, r))
            if rec is None:
                seq = np.frombuffer(a.query_sequence.upper().encode("ascii"), dtype=np.uint8).copy()
                fq = a.get_forward_qualities()
            else:
                seq = np.frombuffer(B.decode_bases(rec["seq"]).encode("ascii"), dtype=np.uint8).copy()
                if rec["qual"] is None:
                    fq = a.get_forward_qualities()
                else:                                            # engine qualities are in printed order (after AM.py:213)
                    q = [int(x) for x in rec["qual"]]
                    fq = array.array("B", reversed(q) if a.is_reverse else q)
            ar = AnonymizedRead(a.query_name, a.is_read1, a.is_read2, a.is_reverse, ds, seq, fq)
            for at, pos, ln, ins in (rec.get("edits") or ()) if rec is not None else ():
                ar.left_over_variants_to_mask.append(
                    (at, MaskedVariant(VariantType.INS if ins else VariantType.DEL, ln, "" if ins else ref[pos:pos + ln].upper())))
            slot = pairs.setdefault(a.query_name, [None, None])
            if slot[ar.get_pair_idx()] is None:                   # first alignment of (name, mate) wins (AM.py:331-335)
                slot[ar.get_pair_idx()] = ar
            max_end[a.query_name] = max(max_end.get(a.query_name, -1), int(a.reference_end))
        self.anonymized_reads = pairs
        # ---- the reference's yield order (AM.py:472-476, 489-512, 521-532): a complete pair leaves at the first
        # normal column right of both mates, in first-appearance order; the rest at the end in registry order
        ncols = sorted(set(reads.normal_cols))
        early, late = [], []
        for name, order in reads.appearance.items():
            p = pairs[name]
            if anonymized_read_pair_is_writeable(p[0], p[1]):
                k = bisect.bisect_right(ncols, max_end[name])
                if k < len(ncols):
                    early.append((ncols[k], order, name))
                    continue
            late.append((order, name))
        for _, _, name in sorted(early):
            yield pairs.pop(name)
        for _, name in sorted(late):
            yield pairs[name]
        self.reset()


# name the reference's CLI looks up (genome_anonymizer.py:10-13)
CompleteGermlineAnonymizer = B200GermlineAnonymizer
