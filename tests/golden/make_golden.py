#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the reference's unmodified Python modules.

Runs only in the build container (needs /root/reference).  pysam and variant_extractor are not
installed, so the reference is imported with tests/ref_stub on sys.path (duck-typed fakes plus a
run-time de-cythonised pileup_io); only htslib's pileup arithmetic is emulated (rules in
tests/ref_stub/pysam.py).  Everything else - allele discovery, the T/N state
machine, masking, left-over handling, FASTQ rendering, statistics, window/section arithmetic - is
the reference's own code:

    src/GenomeAnonymizer/anonymizer_methods.py      CompleteGermlineAnonymizer.anonymize (431-535)
    src/GenomeAnonymizer/variation_classifier.py    classify_variation_in_pileup_column (185-215)
    src/GenomeAnonymizer/short_read_tumor_normal_anonymizer.py   get_windows, anonymize_genome

Outputs (committed): tests/golden/session_cases.json, tests/golden/genome_cases.json.
Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GA_REFERENCE_ROOT", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "tests", "ref_stub"), REF, ROOT]

import pysam  # noqa: E402  (the stub)
import pileup_io  # noqa: E402
import variant_extractor  # noqa: E402
from variant_extractor.variants import VariantRecord, VariantType  # noqa: E402
from src.GenomeAnonymizer.anonymizer_methods import CompleteGermlineAnonymizer  # noqa: E402
from src.GenomeAnonymizer.variants import CalledGenomicVariant  # noqa: E402
from src.GenomeAnonymizer import short_read_tumor_normal_anonymizer as SR  # noqa: E402

from genomeanonymizer_b200 import synth  # noqa: E402


def seg(r, contig):
    return pysam.AlignedSegment(r["name"], r["flag"], contig, r["pos"], r["cigar"], r["seq"], r["qual"])


def keep_variant(contig, k):
    if k is None:
        return None
    return CalledGenomicVariant(contig, k["pos"], k["end"], VariantType[k["type"]], k["length"],
                                allele=k["allele"], ref_allele=k.get("ref", ""))


def run_session(case, window):
    contig = case["contig"]
    t = [seg(r, contig) for r in case["reads"] if r["dataset"] == 0]
    n = [seg(r, contig) for r in case["reads"] if r["dataset"] == 1]
    pysam.register_alignment_file("T.bam", t, [contig])
    pysam.register_alignment_file("N.bam", n, [contig])
    pysam.register_fasta("ref.fa", {contig: case["reference"]})
    fasta = pysam.FastaFile("ref.fa")
    rec = SR.AnonymizedVariantsStatistics("/dev/null")
    w = SR.Window(sequence=contig, first=window["first"], last=window["last"],
                  variant=keep_variant(contig, window.get("keep")))
    rec.add_window(w)
    pile = pileup_io.iter_pileups(pysam.AlignmentFile("T.bam"), pysam.AlignmentFile("N.bam"), fasta,
                                  seq_name=contig, start=window["first"], stop=window["last"])
    anonymizer = CompleteGermlineAnonymizer()
    order = []
    out = {}
    for pair in anonymizer.anonymize(w.variant, pile, fasta, stats_recorder=rec):
        names = [a.query_name for a in pair if a is not None]
        order.append(names[0])
        for a in pair:
            if a is None:
                continue
            mate = 1 if a.is_read1 else 2
            s = bytes(bytearray(int(x) for x in a.anonymized_sequence_array)).decode()
            q = [int(x) for x in a.anonymized_qualities_array]
            fq = a.get_anonymized_fastq_record()
            out[f"{a.dataset_idx}|{a.query_name}|{mate}"] = {"seq": s, "qual_internal": q, "fastq": fq}
    return {"yield_order": order, "counts": rec.window_var_counts[str(w)], "reads": out}


def R(name, flag, pos, cigar, seq, qual, ds):
    return {"name": name, "flag": flag, "pos": pos, "cigar": cigar, "seq": seq, "qual": qual, "dataset": ds}


F1, R1, F2, R2 = 0x1 | 0x2 | 0x20 | 0x40, 0x1 | 0x2 | 0x10 | 0x40, 0x1 | 0x2 | 0x20 | 0x80, 0x1 | 0x2 | 0x10 | 0x80


def q(n, start=10):
    return list(range(start, start + n))


def kat_cases():
    ref = "ACGT" * 12
    cases = []
    # K-A: SNV + DEL + INS shared by T and N, tumor-only SNV kept by state machine, reverse reads
    cases.append({"name": "K-A", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 40, "keep": None}],
                  "reads": [R("t1", F1, 2, "10M", "GTAAGTACGT", q(10), 0),
                            R("t2", F1, 4, "4M2D6M", "ACGTGTACGT", q(10), 0),
                            R("t2", R2, 6, "3M2I5M", "TTATTCGTAC", q(10), 0),
                            R("t1", R2, 8, "10M", "CCGTACGTAC", q(10), 0),
                            R("n1", F1, 0, "12M", "ACGTAAGTCCGT", q(12), 1),
                            R("n1", R2, 4, "4M2D6M", "ACGTGTACGT", q(10), 1),
                            R("n2", F1, 5, "4M2I6M", "CGTATTCGTACG", q(12), 1),
                            R("n2", R2, 20, "8M", "ACGTACGT", q(8), 1)]})
    # K-B two indels + soft clip in one read (Q5/Q6)
    cases.append({"name": "K-B", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 40, "keep": None}],
                  "reads": [R("t1", F1, 4, "2S4M2I4M2D6M", "GGACGTTTACGTGTACGT", q(18), 0),
                            R("t1", R2, 24, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 4, "2S4M2I4M2D6M", "GGACGTTTACGTGTACGT", q(18), 1),
                            R("n1", R2, 24, "8M", "ACGTACGT", q(8), 1)]})
    # K-C keep variant equal / different alt
    base = [R("t1", F1, 0, "12M", "ACGTAAGTACGT", q(12), 0), R("t1", R2, 20, "8M", "ACGTACGT", q(8), 0),
            R("n1", F1, 0, "12M", "ACGTAAGTACGT", q(12), 1), R("n1", R2, 20, "8M", "ACGTACGT", q(8), 1)]
    cases.append({"name": "K-C-keep-equal", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 40,
                               "keep": {"type": "SNV", "pos": 5, "end": 5, "length": 1, "allele": "A"}}],
                  "reads": base})
    cases.append({"name": "K-C-keep-other-alt", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 40,
                               "keep": {"type": "SNV", "pos": 5, "end": 5, "length": 1, "allele": "G"}}],
                  "reads": base})
    cases.append({"name": "K-C-keep-end-mismatch", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 40,
                               "keep": {"type": "SNV", "pos": 5, "end": 6, "length": 1, "allele": "A"}}],
                  "reads": base})
    # K-D hard clip shifts the indel offset (Q7): INS not matched between T and N
    cases.append({"name": "K-D", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 40, "keep": None}],
                  "reads": [R("t1", F1, 4, "3H4M2I6M", "ACGTTTACGTAC", q(12), 0),
                            R("t1", R2, 24, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 4, "4M2I6M", "ACGTTTACGTAC", q(12), 1),
                            R("n1", R2, 24, "8M", "ACGTACGT", q(8), 1)]})
    # Q7 on both sides: identical hard-clipped reads -> matched with shifted offsets (clamped slices)
    cases.append({"name": "K-D2-hardclip-both", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 40, "keep": None}],
                  "reads": [R("t1", F1, 4, "3H4M2I6M", "ACGTTTACGTAC", q(12), 0),
                            R("t1", R2, 24, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 4, "3H4M2I6M", "ACGTTTACGTAC", q(12), 1),
                            R("n1", R2, 24, "8M", "ACGTACGT", q(8), 1),
                            R("t3", F1, 8, "9H4M2D6M", "ACGTGTACGT", q(10), 0),
                            R("n3", F1, 8, "9H4M2D6M", "ACGTGTACGT", q(10), 1),
                            R("t3", R2, 30, "6M", "GTACGT", q(6), 0),
                            R("n3", R2, 30, "6M", "GTACGT", q(6), 1)]})
    # K-E lower-case read base upper-cased and masked; N never a variant; lower-case reference
    ref_l = "ACGTacgtACGT" * 4
    cases.append({"name": "K-E", "contig": "c", "reference": ref_l,
                  "windows": [{"first": 0, "last": 48, "keep": None}],
                  "reads": [R("t1", F1, 0, "12M", "ACGTAaGTNCGT", q(12), 0), R("t1", R2, 20, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 0, "12M", "ACGTAAGTNCGT", q(12), 1), R("n1", R2, 20, "8M", "ACGTACGT", q(8), 1)]})
    # reference N disables calling; IUPAC read base on forward read
    ref_n = "ACGTANGTACGT" + "ACGT" * 8
    cases.append({"name": "K-refN", "contig": "c", "reference": ref_n,
                  "windows": [{"first": 0, "last": 44, "keep": None}],
                  "reads": [R("t1", F1, 0, "12M", "ACGTAAGTACTT", q(12), 0), R("t1", F2, 20, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 0, "12M", "ACGTAAGTACTT", q(12), 1), R("n1", F2, 20, "8M", "ACGTACGT", q(8), 1)]})
    # reverse read with DEL + INS: forward-orientation quality quirk (Q1/Q2), distinct qualities
    cases.append({"name": "K-rev-indel", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 44, "keep": None}],
                  "reads": [R("t1", R1, 4, "4M2D3M3I5M", "ACGTGTAGGGCGTAC", [3, 9, 27, 40, 5, 6, 7, 30, 31, 32, 11, 12, 13, 14, 2], 0),
                            R("t1", F2, 24, "8M", "ACGTACGT", q(8), 0),
                            R("n1", R1, 4, "4M2D3M3I5M", "ACGTGTAGGGCGTAC", [13, 19, 7, 4, 15, 16, 17, 3, 1, 2, 21, 22, 23, 24, 12], 1),
                            R("n1", F2, 24, "8M", "ACGTACGT", q(8), 1)]})
    # tumor-only / normal-only evidence is never masked; empty normal dataset
    cases.append({"name": "K-tumor-only", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 44, "keep": None}],
                  "reads": [R("t1", F1, 0, "12M", "ACGTAAGTACGT", q(12), 0), R("t1", R2, 20, "8M", "ACGTACGT", q(8), 0),
                            R("t2", F1, 1, "11M", "CGTAAGTACGT", q(11), 0), R("t2", R2, 21, "7M", "CGTACGT", q(7), 0)]})
    # mate outside the window -> pair yielded incomplete at session end; insertion at read end + soft clip
    cases.append({"name": "K-partial-pair", "contig": "c", "reference": "ACGT" * 30,
                  "windows": [{"first": 10, "last": 30, "keep": None}],
                  "reads": [R("t1", F1, 8, "12M", "ACTTACGTACGT", q(12), 0), R("t1", R2, 80, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 8, "12M", "ACTTACGTACGT", q(12), 1), R("n1", R2, 16, "4M2I4M3S", "ACGTCCACGTGGG", q(13), 1),
                            R("t2", F1, 16, "4M2I4M3S", "ACGTCCACGTGGG", q(13), 0)]})
    # insertion whose allele is truncated by the read end (python slice clamp), DEL allele of <2 bases
    cases.append({"name": "K-trunc-allele", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 44, "keep": None}],
                  "reads": [R("t1", F1, 4, "6M3D1M", "ACGTACT", q(7), 0), R("t1", R2, 24, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 4, "6M3D1M", "ACGTACT", q(7), 1), R("n1", R2, 24, "8M", "ACGTACGT", q(8), 1),
                            R("t2", F1, 6, "4M3D1M", "GTACT", q(5), 0), R("n2", F1, 6, "4M3D1M", "GTACG", q(5), 1),
                            R("t2", R2, 26, "6M", "GTACGT", q(6), 0), R("n2", R2, 26, "6M", "GTACGT", q(6), 1)]})
    # same position I and D in one read, adjacent indels, three DELs (mean recomputation)
    cases.append({"name": "K-multi-indel", "contig": "c", "reference": "ACGT" * 16,
                  "windows": [{"first": 0, "last": 64, "keep": None}],
                  "reads": [R("t1", F1, 2, "4M1D4M2D4M3D4M2I2M", "GTACTACGCGTAACGTTTAC", [40, 2, 30, 4, 25, 6, 20, 8, 15, 10, 12, 33, 14, 3, 16, 37, 18, 9, 20, 1], 0),
                            R("t1", R2, 40, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 2, "4M1D4M2D4M3D4M2I2M", "GTACTACGCGTAACGTTTAC", q(20, 5), 1),
                            R("n1", R2, 40, "8M", "ACGTACGT", q(8), 1),
                            R("t2", F1, 4, "4M2I2D6M", "ACGTGGGTACGT", q(12), 0), R("n2", F1, 4, "4M2I2D6M", "ACGTGGGTACGT", q(12, 3), 1),
                            R("t2", R2, 44, "8M", "ACGTACGT", q(8), 0), R("n2", R2, 44, "8M", "ACGTACGT", q(8), 1)]})
    # the same on the reverse strand, plus a forward mate with five indels: edits index the forward-orientation
    # quality array while the bases stay in alignment order (quirks Q1/Q2 with more than two edits per read)
    cases.append({"name": "K-multi-indel-rev", "contig": "c", "reference": "ACGT" * 20,
                  "windows": [{"first": 0, "last": 80, "keep": None}],
                  "reads": [R("t1", F1, 2, "3M1I3M2D4M1D3M2I3M1D2M", "GTATCGTGTACTACTTGTAGT", [3, 40, 5, 38, 7, 36, 9, 34, 11, 32, 13, 30, 15, 28, 17, 26, 19, 24, 21, 22, 2], 0),
                            R("t1", R2, 38, "4M2D4M1I4M3D4M", "GTACACGTTACGTCGTA", [40, 2, 30, 4, 25, 6, 20, 8, 15, 10, 12, 33, 14, 3, 16, 37, 18], 0),
                            R("n1", F1, 2, "3M1I3M2D4M1D3M2I3M1D2M", "GTATCGTGTACTACTTGTAGT", q(21, 4), 1),
                            R("n1", R2, 38, "4M2D4M1I4M3D4M", "GTACACGTTACGTCGTA", q(17, 7), 1),
                            R("t2", F1, 60, "8M", "ACGTACGT", q(8), 0), R("n2", F1, 60, "8M", "ACGTACGT", q(8), 1),
                            R("t2", R2, 66, "8M", "GTACGTAC", q(8), 0), R("n2", R2, 66, "8M", "GTACGTAC", q(8), 1)]})
    # indel keep-variant that matches a discovered indel key exactly
    cases.append({"name": "K-keep-indel", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 44,
                               "keep": {"type": "INS", "pos": 8, "end": 9, "length": 2, "allele": "TT"}}],
                  "reads": [R("t1", F1, 4, "4M2I6M", "ACGTTTACGTAC", q(12), 0), R("t1", R2, 24, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 4, "4M2I6M", "ACGTTTACGTAC", q(12), 1), R("n1", R2, 24, "8M", "ACGTACGT", q(8), 1)]})
    # an insertion as the last op of a read (its position is the read's reference_end), with and without a soft clip behind
    # it; seen by both datasets (t1 / n1, t3 / n3) and by the tumor alone where no normal read reaches (t2)  [ADVICE r01]
    cases.append({"name": "K-trailing-ins", "contig": "c", "reference": ref,
                  "windows": [{"first": 0, "last": 44, "keep": None}],
                  "reads": [R("t1", F1, 4, "8M2I", "ACGTACGTTT", q(10), 0), R("t1", R2, 24, "8M", "ACGTACGT", q(8), 0),
                            R("n1", F1, 4, "8M2I", "ACGTACGTTT", q(10, 3), 1), R("n1", R2, 24, "8M", "ACGTACGT", q(8), 1),
                            R("t2", F1, 14, "6M2I", "GTACGTCC", q(8), 0), R("t2", R2, 32, "8M", "ACGTACGT", q(8), 0),
                            R("t3", R1, 26, "6M2I2S", "GTACGTAAGG", q(10), 0), R("n3", R1, 26, "6M2I2S", "GTACGTAACC", q(10, 7), 1),
                            R("t3", F2, 34, "6M", "GTACGT", q(6), 0), R("n3", F2, 34, "6M", "GTACGT", q(6), 1)]})
    # NOT a golden case: the same (name, mate) in the tumor and in the normal file.  The reference keys seen_read_alns and its
    # read registry by name and mate alone, so the normal read is taken for the tumor read already seen - its indels are
    # never discovered and it is never written; the engine keeps the datasets apart.  Read names of two sequencing runs do
    # not collide in practice; stated as a limit in DESIGN.md (section 9), not reproduced.
    for c in cases:      # a BAM is coordinate sorted; keep hand-written order among equal positions
        c["reads"] = sorted(c["reads"], key=lambda r: (r["dataset"], r["pos"]))
    return cases


def random_cases():
    cases = []
    cases.append(synth.make_case(11, contig_len=2600, n_pairs=(150, 130), read_len=60, name="rand-11-dense-indel",
                                 snp_rate=4e-3, indel_rate=3e-3, err_rate=4e-3, n_rate=1e-3, clip_frac=0.25))
    cases.append(synth.make_case(12, contig_len=2800, n_pairs=(140, 90), read_len=75, name="rand-12-refN",
                                 snp_rate=3e-3, indel_rate=1.5e-3, ref_n_runs=0.3, ref_lower=0.1))
    cases.append(synth.make_case(13, contig_len=5000, n_pairs=(150, 140), read_len=100, name="rand-13-two-windows",
                                 somatic_positions=[1500, 3600], snp_rate=3e-3, indel_rate=1e-3))
    cases.append(synth.make_case(14, contig_len=2400, n_pairs=(220, 200), read_len=40, name="rand-14-short-nokeep",
                                 snp_rate=5e-3, indel_rate=4e-3, keep_somatic=False, max_indel=12, clip_frac=0.4))
    cases.append(synth.make_case(15, contig_len=3000, n_pairs=(120, 0), read_len=50, name="rand-15-no-normal"))
    cases.append(synth.make_case(16, contig_len=2600, n_pairs=(45, 40), read_len=150, name="rand-16-many-indels",
                                 snp_rate=3e-3, indel_rate=2e-2, max_indel=4, clip_frac=0.2))
    return cases


def run_genome(case, vcf_records, more_cases=()):
    """Whole-sample orchestration of the reference (anonymize_genome) under functional fakes; more_cases = further
    contigs of the same sample (each a case with its own contig name), in genome order."""
    cases = [case] + list(more_cases)
    t = [seg(r, c["contig"]) for c in cases for r in c["reads"] if r["dataset"] == 0]
    n = [seg(r, c["contig"]) for c in cases for r in c["reads"] if r["dataset"] == 1]
    pysam.register_alignment_file("T.bam", t, [c["contig"] for c in cases])
    pysam.register_alignment_file("N.bam", n, [c["contig"] for c in cases])
    pysam.register_fasta("ref.fa", {c["contig"]: c["reference"] for c in cases})
    variant_extractor.register_vcf("s.vcf", vcf_records)
    fasta = pysam.FastaFile("ref.fa")
    windows = SR.get_windows(variant_extractor.VariantExtractor("s.vcf"), SR.get_ref_idxs(fasta))
    sections = SR.get_genome_sections(windows, fasta)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            SR.anonymize_genome(windows, "T.bam", "N.bam", "ref.fa", CompleteGermlineAnonymizer(),
                                os.path.join(td, "T.anonymized"), os.path.join(td, "N.anonymized"),
                                None, None, True, 1)
            files = {}
            for nm in ("T.anonymized.1.fastq", "T.anonymized.2.fastq", "N.anonymized.1.fastq",
                       "N.anonymized.2.fastq", "T.anonymized.single_end.fastq", "N.anonymized.single_end.fastq",
                       "N.bam.statistics.txt"):
                pth = os.path.join(td, nm)
                files[nm] = open(pth).read() if os.path.exists(pth) else None
        finally:
            os.chdir(cwd)
    return {"windows": [[w.sequence, w.first, w.last] for w in windows],
            "sections": [[s.sequence, s.first, s.last, s.variant is not None] for s in sections],
            "files": files}


def unmap_mates(case, every):
    """Every `every`-th pair of each dataset: one mate (alternating) becomes an unmapped read placed at its mate's
    position, as aligners write it: flag 0x4, no CIGAR, RNAME / POS of the mate, stored sequence as read; the mapped mate
    gets 0x8 and loses 0x2.  File order: by position, the unmapped read behind mapped reads of the same position."""
    reads = case["reads"]
    by_name = {}
    for k, r in enumerate(reads):
        by_name.setdefault((r["dataset"], r["name"]), []).append(k)
    n = 0
    for key in sorted(by_name):
        idx = by_name[key]
        if len(idx) != 2:
            continue
        n += 1
        if n % every:
            continue
        keep, unm = (idx[0], idx[1]) if (n // every) % 2 else (idx[1], idx[0])
        rk, ru = reads[keep], reads[unm]
        ru["pos"], ru["cigar"] = rk["pos"], "*"
        ru["flag"] = (ru["flag"] & (0x40 | 0x80)) | 0x1 | 0x4 | (0x20 if rk["flag"] & 0x10 else 0)
        rk["flag"] = (rk["flag"] & ~0x2 & ~0x20) | 0x8
    order = sorted(range(len(reads)), key=lambda k: (reads[k]["dataset"], reads[k]["pos"], 1 if reads[k]["flag"] & 4 else 0))
    case["reads"] = [reads[k] for k in order]


def genome_cases():
    out = []
    specs = [
        # two windows 2,150 bp apart: reads of length 80 can overlap both (gap of 149 bp)
        (21, dict(contig_len=9000, n_pairs=(260, 240), read_len=80, somatic_positions=[2500, 4650], snp_rate=2e-3, indel_rate=0.0,
                  clip_frac=0.0), "snv-only", None),
        (22, dict(contig_len=9000, n_pairs=(260, 240), read_len=80, somatic_positions=[2500, 4650], snp_rate=2e-3, indel_rate=1.5e-3,
                  clip_frac=0.1), "with-indels", None),
        # ~2x coverage: many read islands between the windows, tumor / normal island pairs become extra sessions
        (23, dict(contig_len=9000, n_pairs=(70, 60), read_len=80, somatic_positions=[2500, 4650], snp_rate=4e-3, indel_rate=1e-3,
                  clip_frac=0.1), "sparse-islands", None),
        # reads whose mate is missing: the single_end files and pairs completed across sections
        (24, dict(contig_len=9000, n_pairs=(200, 180), read_len=100, somatic_positions=[2300, 4600, 6900], snp_rate=3e-3,
                  indel_rate=8e-4, clip_frac=0.1), "orphans-three-windows", 7),
        # no somatic variant at all: the whole contig is one inter-window region
        (25, dict(contig_len=5000, n_pairs=(40, 35), read_len=70, somatic_positions=[], snp_rate=4e-3, indel_rate=1e-3), "no-variants", None),
        # placed-unmapped mates (flag 0x4 at the mate's position, mate flagged 0x8), inside windows and between them:
        # pileups never see them (htslib), regions collect them for their end (pileup_io.pyx:93-96, 298), and what is
        # left is paired after the last section (pair_unmapped_mates, short_read_tumor_normal_anonymizer.py:561-600)
        (41, dict(contig_len=9000, n_pairs=(120, 110), read_len=80, somatic_positions=[2500, 6000], snp_rate=3e-3, indel_rate=8e-4,
                  clip_frac=0.1), "unmapped-mates", -6),
        # quirk Q12: thin coverage, orphans and many indels - reads that one session masks and parks unpaired are met again by
        # an overlapping later session, and the reference applies their left-over indels a second time when it writes them
        (54, dict(contig_len=9000, n_pairs=(70, 60), read_len=80, somatic_positions=[2500, 4650, 6800], snp_rate=4e-3, indel_rate=3e-3,
                  clip_frac=0.1), "twice-masked", 5),
    ]
    for seed, kw, label, drop in specs:
        case = synth.make_case(seed, name=f"genome-{label}", **kw)
        if drop and drop < 0:
            unmap_mates(case, -drop)
        elif drop:                                                # deterministic orphans: every `drop`-th read disappears
            case["reads"] = [r for k, r in enumerate(case["reads"]) if k % drop != 3]
        vcf = []
        for w in case["windows"]:
            k = w["keep"]
            r = case["reference"][k["pos"]].upper()
            vcf.append(VariantRecord("c", k["pos"] + 1, k["pos"] + 1, 1, r, k["allele"], VariantType.SNV))
        res = run_genome(case, vcf)
        out.append({"case": case, "vcf": [[v.contig, v.pos, v.end, v.length, v.ref, v.alt, v.variant_type.name] for v in vcf],
                    "expected": res})
    return out


def two_contig_case():
    """One sample over two contigs with pairs whose mates lie on different contigs (mate 1 on c1, mate 2 on c2): the
    reference keeps unpaired reads across contigs and writes such a pair when the second mate is processed."""
    a = synth.make_case(31, name="genome-two-contigs-c1", contig_len=9000, n_pairs=(260, 240), read_len=70,
                        somatic_positions=[2500, 6000], snp_rate=4e-3, indel_rate=1e-3)
    b = synth.make_case(32, name="genome-two-contigs-c2", contig_len=8000, n_pairs=(240, 230), read_len=70,
                        somatic_positions=[3200], snp_rate=4e-3, indel_rate=1e-3)
    a["contig"], b["contig"] = "c1", "c2"
    b["reads"] = [dict(r, name=r["name"] + "b") for r in b["reads"]]      # read names are unique within a sample
    for ds in (0, 1):
        # every 9th pair of c1 keeps only its first mate, every 9th pair of c2 only its second, renamed to the c1 pair
        names_a = sorted({r["name"] for r in a["reads"] if r["dataset"] == ds})[::9]
        names_b = sorted({r["name"] for r in b["reads"] if r["dataset"] == ds})[::9]
        pairs = list(zip(names_a, names_b))
        ren = {nb: na for na, nb in pairs}
        keep_a = {na for na, _ in pairs}
        a["reads"] = [r for r in a["reads"] if not (r["dataset"] == ds and r["name"] in keep_a and not (r["flag"] & 0x40))]
        out = []
        for r in b["reads"]:
            if r["dataset"] == ds and r["name"] in ren:
                if r["flag"] & 0x40:
                    continue                                     # its first mate goes
                r = dict(r, name=ren[r["name"]])
            out.append(r)
        b["reads"] = out
    vcf = []
    for c in (a, b):
        for w in c["windows"]:
            k = w["keep"]
            vcf.append(VariantRecord(c["contig"], k["pos"] + 1, k["pos"] + 1, 1, c["reference"][k["pos"]].upper(), k["allele"], VariantType.SNV))
    res = run_genome(a, vcf, [b])
    return {"cases": [a, b], "vcf": [[v.contig, v.pos, v.end, v.length, v.ref, v.alt, v.variant_type.name] for v in vcf], "expected": res}


def window_kats():
    """K-F: window and section arithmetic (SR.get_windows / get_genome_sections)."""
    pysam.register_fasta("w.fa", {"c1": "A" * 100000, "c2": "A" * 5000})
    fasta = pysam.FastaFile("w.fa")
    recs = [VariantRecord("c1", 500, 500, 1, "A", "C", VariantType.SNV),
            VariantRecord("c1", 30000, 30000, 1, "A", "C", VariantType.SNV),
            VariantRecord("c1", 31500, 31500, 1, "A", "G", VariantType.SNV),
            VariantRecord("c1", 60000, 60010, 10, "AAAAAAAAAAA", "A", VariantType.DEL),
            VariantRecord("c1", 70000, 70001, 4, "A", "ACCCC", VariantType.INS)]
    ws = SR.get_windows(recs, SR.get_ref_idxs(fasta))
    secs = SR.get_genome_sections(ws, fasta)
    return {"contigs": {"c1": 100000, "c2": 5000},
            "vcf": [[v.contig, v.pos, v.end, v.length, v.ref, v.alt, v.variant_type.name] for v in recs],
            "windows": [[w.sequence, w.first, w.last, w.variant.pos, w.variant.end, w.variant.variant_type.name,
                         w.variant.length, w.variant.allele] for w in ws],
            "sections": [[s.sequence, s.first, s.last, s.variant is not None] for s in secs]}


def sv_window_kats():
    """Window arithmetic of SR.get_windows for structural records: INV (one or two windows), DUP / CNV / symbolic DEL
    (one window, or two from 100 kb on).  The records are built by hand (variant-extractor is not installed); what is
    pinned here is the reference's geometry given pos / end / length / type."""
    pysam.register_fasta("sv.fa", {"c1": "A" * 600000, "c2": "A" * 5000})
    fasta = pysam.FastaFile("sv.fa")
    recs = [VariantRecord("c1", 10000, 15000, 5000, "A", "<DEL>", VariantType.DEL),
            VariantRecord("c1", 20000, 22000, 2000, "A", "<DUP>", VariantType.DUP),
            VariantRecord("c1", 40000, 40500, 500, "A", "<INV>", VariantType.INV),
            VariantRecord("c1", 200000, 260000, 60000, "A", "<INV>", VariantType.INV),
            VariantRecord("c1", 300000, 450000, 150000, "A", "<DEL>", VariantType.DEL),
            VariantRecord("c1", 500000, 503000, 3000, "A", "<CNV>", VariantType.CNV),
            VariantRecord("c1", 550000, 550001, 300, "A", "<INS>", VariantType.INS)]
    ws = SR.get_windows(recs, SR.get_ref_idxs(fasta))
    return {"contigs": {"c1": 600000, "c2": 5000},
            "vcf": [[v.contig, v.pos, v.end, v.length, v.ref, v.alt, v.variant_type.name] for v in recs],
            "windows": [[w.sequence, w.first, w.last, w.variant.pos, w.variant.end, w.variant.variant_type.name,
                         w.variant.length, w.variant.allele] for w in ws]}


def main():
    import logging
    logging.disable(logging.CRITICAL)
    sess = []
    for case in kat_cases() + random_cases():
        exp = [run_session(case, w) for w in case["windows"]]
        sess.append({"case": case, "expected": exp})
        nmod = 0
        for e in exp:
            print(f"{case['name']:28s} reads={len(e['reads']):4d} counts={e['counts'][:3]} yield={len(e['yield_order'])}")
    with open(os.path.join(HERE, "session_cases.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": sess}, f, separators=(",", ":"))
    gen = genome_cases()
    for g in gen:
        print(g["case"]["name"], {k: (len(v) if v else v) for k, v in g["expected"]["files"].items()})
    with open(os.path.join(HERE, "genome_cases.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": gen, "windows_kat": window_kats(),
                   "sv_windows_kat": sv_window_kats(), "two_contigs": two_contig_case()}, f,
                  separators=(",", ":"))


if __name__ == "__main__":
    main()
