"""Duck-typed in-memory stand-in for the parts of pysam the reference touches.

TEST INFRASTRUCTURE ONLY.  pysam/htslib is not installed in this image, so the
golden-vector generator (tests/golden/make_golden.py) imports the reference's
unmodified Python modules from /root/reference with this module on sys.path.
Only what the reference resolves at import or run time is provided
(SURVEY.md Appendix A).  The pileup emulation follows the documented htslib
behaviour:

    reads(region)  = mapped reads with reference_start < stop and reference_end > start, file order
    columns        = every covered position of those reads, ascending (truncate=False)
    qpos(r, p)     = number of M/=/X/I/S bases before p if p lies in an M/=/X op, else None
"""
import array
import re

_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=XB])")
_REGISTRY = {}          # path -> list[AlignedSegment] (coordinate sorted) or FASTA dict


def register_alignment_file(path, reads, references):
    _REGISTRY[path] = ("bam", list(reads), list(references))


def register_fasta(path, contigs):
    """contigs: ordered dict name -> sequence string"""
    _REGISTRY[path] = ("fasta", dict(contigs))


class AlignedSegment:
    def __init__(self, query_name, flag, reference_name, reference_start, cigarstring,
                 query_sequence, query_qualities, reference_id=0, tags=None):
        self.query_name = query_name
        self.flag = flag
        self.reference_name = reference_name
        self.reference_id = reference_id
        self.reference_start = reference_start
        self.cigarstring = cigarstring
        self.query_sequence = query_sequence
        self.query_qualities = array.array('B', query_qualities)
        self._tags = dict(tags or {})
        self.cigartuples = [(int(n), op) for n, op in _CIGAR_RE.findall(cigarstring or "")]

    # flag accessors -------------------------------------------------------
    @property
    def is_paired(self): return bool(self.flag & 0x1)
    @property
    def is_unmapped(self): return bool(self.flag & 0x4)
    @property
    def is_mapped(self): return not self.is_unmapped
    @property
    def is_reverse(self): return bool(self.flag & 0x10)
    @property
    def is_read1(self): return bool(self.flag & 0x40)
    @property
    def is_read2(self): return bool(self.flag & 0x80)
    @property
    def is_secondary(self): return bool(self.flag & 0x100)
    @property
    def is_supplementary(self): return bool(self.flag & 0x800)

    @property
    def reference_end(self):
        if self.is_unmapped or not self.cigartuples:
            return None
        return self.reference_start + sum(n for n, op in self.cigartuples if op in "MDN=X")

    def get_forward_qualities(self):
        q = array.array('B', self.query_qualities)
        if self.is_reverse:
            q.reverse()
        return q

    def has_tag(self, tag): return tag in self._tags
    def get_tag(self, tag): return self._tags[tag]

    def to_string(self):
        return "\t".join(map(str, (self.query_name, self.flag, self.reference_name,
                                   self.reference_start + 1, self.cigarstring, self.query_sequence)))

    def qpos_at(self, p):
        ref, q = self.reference_start, 0
        for n, op in self.cigartuples:
            if op in "M=X":
                if ref <= p < ref + n:
                    return q + (p - ref)
                ref += n
                q += n
            elif op in "DN":
                if ref <= p < ref + n:
                    return None
                ref += n
            elif op in "IS":
                q += n
        return None


class PileupRead:
    def __init__(self, alignment, query_position):
        self.alignment = alignment
        self.query_position = query_position
        self.is_del = query_position is None


class PileupColumn:
    def __init__(self, reference_name, reference_pos, pileups):
        self.reference_name = reference_name
        self.reference_pos = reference_pos
        self.pileups = pileups
        self.nsegments = len(pileups)

    def __len__(self):
        return len(self.pileups)


def emulate_pileup(reads, reference_name, start, stop):
    """Generator of PileupColumn over in-memory reads (file order), htslib rules."""
    sel = [r for r in reads if r.is_mapped and r.reference_name == reference_name
           and (stop is None or r.reference_start < stop) and (start is None or r.reference_end > start)]
    if not sel:
        return
    lo = min(r.reference_start for r in sel)
    hi = max(r.reference_end for r in sel)
    # sweep: bucket reads by start for speed
    by_start = {}
    for r in sel:
        by_start.setdefault(r.reference_start, []).append(r)
    active = []
    for p in range(lo, hi):
        if p in by_start:
            active.extend(by_start[p])
            # keep file order
            order = {id(r): i for i, r in enumerate(sel)}
            active.sort(key=lambda r: order[id(r)])
        active = [r for r in active if r.reference_end > p]
        if active:
            yield PileupColumn(reference_name, p, [PileupRead(r, r.qpos_at(p)) for r in active])


class AlignmentFile:
    def __init__(self, path, mode="r", reference_filename=None, threads=1, header=None, **kw):
        kind, reads, refs = _REGISTRY[path]
        assert kind == "bam"
        self._reads = reads
        self.references = refs
        self.filename = path

    def __enter__(self): return self
    def __exit__(self, *a): return False
    def close(self): pass

    def pileup(self, reference=None, start=None, end=None, **kw):
        return emulate_pileup(self._reads, reference, start, end)

    def fetch(self, reference=None, start=None, stop=None, until_eof=False, **kw):
        if start is not None and stop is not None and (start < 0 or start > stop):
            raise ValueError(f"invalid coordinates: start ({start}) > stop ({stop})")
        if start is not None and start < 0:
            raise ValueError("start out of range")
        for r in self._reads:
            if reference is not None and r.reference_name != reference:
                continue
            s = r.reference_start
            e = r.reference_end if r.is_mapped else s + 1
            if (stop is None or s < stop) and (start is None or e > start):
                yield r


class FastaFile:
    def __init__(self, path):
        kind, contigs = _REGISTRY[path]
        assert kind == "fasta"
        self._c = contigs
        self.references = list(contigs.keys())
        self.lengths = [len(v) for v in contigs.values()]

    def fetch(self, reference=None, start=None, end=None):
        s = self._c[reference]
        return s[start:end]

    def close(self): pass


class FastxRecord:  # referenced only in comments of the reference
    pass
