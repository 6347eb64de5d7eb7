#!/usr/bin/env python
"""BASELINE config 1 through the REFERENCE'S OWN PYTHON (build container only: needs /root/reference).

Every window of the synthetic workload `chr22-1k` (1,000 somatic SNVs, 84,000 session reads) is run through the
reference's unmodified `CompleteGermlineAnonymizer.anonymize` (anonymizer_methods.py:431-535) over
`pileup_io.iter_pileups` (pileup_io.pyx:8-41) under the stub pysam of tests/ref_stub - one core, as the reference runs
one tumor-normal pair.  The masked reads are reduced to the record digest of include/ga_digest.h and compared with the
oracle's digest of the same workload (tests/golden/workload_digests.json), so the timing is of a run whose output is
the one the engine is checked against.  Writes tests/golden/ref_timing.json (bench.py prints it as
`cpu_baseline_reference_python`).

    python tools/ref_python_timing.py [n_windows]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GA_REFERENCE_ROOT", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "tests", "ref_stub"), REF, ROOT]

import numpy as np  # noqa: E402
import pysam  # noqa: E402  (the stub)
import pileup_io  # noqa: E402
from variant_extractor.variants import VariantType  # noqa: E402
from src.GenomeAnonymizer.anonymizer_methods import CompleteGermlineAnonymizer  # noqa: E402
from src.GenomeAnonymizer.variants import CalledGenomicVariant  # noqa: E402

from genomeanonymizer_b200 import _abi  # noqa: E402
from genomeanonymizer_b200 import batch as B  # noqa: E402
from genomeanonymizer_b200 import synthdev as SD  # noqa: E402
from tests.test_digest import python_record_hash  # noqa: E402

M64 = (1 << 64) - 1
OPS = "MIDNSHP=X"


def main():
    name = "chr22-1k"
    cfg = SD.WORKLOADS[name]
    n_w = int(sys.argv[1]) if len(sys.argv) > 1 else cfg.total_windows
    b, s, ref = SD.generate_host(cfg, 0, n_w)
    ref_text = ref.decode("ascii")
    pl = cfg.plan()
    per_t, per_n = (int(x) for x in pl.reads_per_window)
    contig = "chr22"
    pysam.register_fasta("ref.fa", {contig: ref_text})
    fasta = pysam.FastaFile("ref.fa")
    asc = np.frombuffer(_abi.CODE2ASC.encode(), np.uint8)

    def segment(r, nm):
        L = b.read_len(r)
        cig = "".join(f"{int(w) >> 4}{OPS[int(w) & 15]}" for w in b.cigar[int(b.cigar_off[r]):int(b.cigar_off[r + 1])])
        seq = bytes(asc[b.sequence_codes(r)]).decode()
        k = np.searchsorted(b.qual_reads, r)
        if k < len(b.qual_reads) and b.qual_reads[k] == r:
            o = int(b.qual_off16[k]) * 32
            q = [int(x) for x in b.qual[o:o + L]]
        else:
            q = [30] * L                                    # reads without an I/D op never have their qualities changed
        return pysam.AlignedSegment(nm, b.flag(r), contig, int(b.pos[r]), cig, seq, q)

    digest = [0, 0, 0, 0]
    t_ref = 0.0
    reads = 0
    masked = [0, 0, 0]
    for w in range(n_w):
        t_idx = list(range(w * per_t, (w + 1) * per_t))
        n_idx = list(range(b.n_tumor + w * per_n, b.n_tumor + (w + 1) * per_n))
        # names: one per read (every synthetic read is its own pair half; T / N names differ, Appendix B)
        t_al = [segment(r, f"T{r}") for r in t_idx]
        n_al = [segment(r, f"N{r}") for r in n_idx]
        pysam.register_alignment_file("T.bam", t_al, [contig])
        pysam.register_alignment_file("N.bam", n_al, [contig])
        keep = CalledGenomicVariant(contig, int(s.keep_pos[w]), int(s.keep_end[w]), VariantType.SNV, 1,
                                    allele=chr(int(s.keep_alleles[int(s.keep_allele_off[w])])), ref_allele="")
        orig = {a.query_name: a for a in t_al + n_al}

        class Rec:
            def count_variant(self, v):
                masked[v.variant_type.value - 1] += 1
        t0 = time.perf_counter()
        pile = pileup_io.iter_pileups(pysam.AlignmentFile("T.bam"), pysam.AlignmentFile("N.bam"), fasta, seq_name=contig,
                                      start=int(s.first[w]), stop=int(s.last[w]))
        out = []
        for pair in CompleteGermlineAnonymizer().anonymize(keep, pile, fasta, stats_recorder=Rec()):
            out.extend(a for a in pair if a is not None)
        t_ref += time.perf_counter() - t0
        reads += len(out)
        for a in out:
            o = orig[a.query_name]
            seq = bytes(bytearray(int(x) for x in a.anonymized_sequence_array)).decode()
            fq = [int(x) for x in a.anonymized_qualities_array]
            printed = list(reversed(fq)) if a.is_reverse else fq            # anonymizer_methods.py:213
            q_changed = printed != [int(x) for x in o.query_qualities]
            if seq == o.query_sequence.upper() and not q_changed:
                continue
            r = int(a.query_name[1:])
            gid = r if a.query_name[0] == "T" else (1 << 40) | (r - b.n_tumor)
            codes = B.encode_bases(seq)
            lo, hi = python_record_hash(0, w, gid, codes, printed if q_changed else None)
            digest = [(digest[0] + lo) & M64, (digest[1] + hi) & M64, digest[2] + 1, digest[3] + len(seq)]
        if w % 100 == 99:
            print(f"{w + 1} windows, {reads} session reads, {t_ref:.1f} s in the reference", flush=True)
    want = None
    try:
        want = json.load(open(os.path.join(ROOT, "tests", "golden", "workload_digests.json")))[name]["total"] if n_w == cfg.total_windows else None
    except Exception:
        pass
    res = {"workload": name, "windows": n_w, "session_reads": reads, "seconds": t_ref, "value": reads / t_ref, "unit": "reads/s", "cores": 1,
           "kind": "reference", "masked_snv_del_ins": masked, "digest": digest, "oracle_digest": want,
           "records_equal_oracle": (digest == want) if want else None,
           "how": "reference CompleteGermlineAnonymizer.anonymize + pileup_io.iter_pileups (unmodified sources) under tests/ref_stub pysam, "
                  "one core of the build container, timed around the anonymize() generator of every window; tools/ref_python_timing.py"}
    print(json.dumps(res))
    if n_w == cfg.total_windows:
        with open(os.path.join(ROOT, "tests", "golden", "ref_timing.json"), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
