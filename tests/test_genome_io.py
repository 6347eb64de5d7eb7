"""N4: BGZF/BAM, FASTA and VCF readers feeding the engine, and the file-level entry point.

CPU: the C++ readers (include/ga_genome_io.h) against files written by the test-side writers of tests/helpers.py -
the packed batch must equal batch.pack_reads of the same reads array for array, so everything proven for packed
batches (tests/test_genome_files.py, the golden samples of the reference) carries over to BAM input.
GPU: run_short_read_tumor_normal_anonymizer file to file = the files the reference's anonymize_genome wrote
(tests/golden/genome_cases.json), byte for byte."""
import gzip
import os

import numpy as np
import pytest

from genomeanonymizer_b200 import batch as B
from genomeanonymizer_b200 import genome_files as GF
from genomeanonymizer_b200 import _lib
from tests import helpers as H

GOLD = H.load_golden("genome_cases.json")
GENOME = GOLD["cases"]
IDS = [e["case"]["name"] for e in GENOME]
ARRAYS = ["pos", "len_flag", "seq_off16", "cigar_off", "cigar", "seq4", "qual"]


def test_library_exports_the_genome_io_abi():
    L = _lib.lib()
    for name in GF.IO_EXPORTS:
        assert hasattr(L, name), name
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "ga_genome_io.h")).read()
    for name in GF.IO_EXPORTS:
        assert name + "(" in hdr, name


@pytest.mark.parametrize("entry", GENOME, ids=IDS)
def test_bam_reader_packs_the_same_batch_as_the_array_packer(entry, tmp_path):
    case = entry["case"]
    t, n, fa, vc = H.write_sample_files(str(tmp_path), case, entry["vcf"])
    assert gzip.open(t).read(4) == b"BAM\1"                          # the test writer makes real gzip members
    with GF.BamFile(t, 3) as T, GF.BamFile(n, 1) as N:
        assert T.references == (case["contig"],) and T.lengths == (len(case["reference"]),)
        assert T.n_records + N.n_records == len(case["reads"])
        cb = GF.pack_tumor_normal(T, N, case["contig"])
    ref = B.pack_reads(H.ordered_reads(case))
    for f in ARRAYS:
        assert np.array_equal(getattr(cb.batch, f), getattr(ref, f)), f
    assert cb.batch.n_tumor == ref.n_tumor and cb.batch.max_ref_span == ref.max_ref_span
    assert cb.batch.seq4.ctypes.data % 16 == 0 and cb.batch.qual.ctypes.data % 16 == 0
    assert [cb.name(k) for k in range(cb.batch.n_reads)] == ref.names
    rows = cb.read_table()
    want = [dict(name=r["name"], flag=r["flag"], dataset=r["dataset"], pos=r["pos"]) for r in H.ordered_reads(case)]
    assert [{k: r[k] for k in ("name", "flag", "dataset", "pos")} for r in rows] == want
    spans = [B.ref_span(B.parse_cigar(r["cigar"])) for r in H.ordered_reads(case)]
    assert [r["end"] - r["pos"] for r in rows] == spans
    F = GF.FastaFile(fa)
    assert F.references == (case["contig"],) and F.fetch(case["contig"]) == case["reference"]
    assert F.fetch(case["contig"], 7, 131) == case["reference"][7:131]
    assert F.fetch(case["contig"], len(case["reference"]) - 3, len(case["reference"]) + 50) == case["reference"][-3:]
    assert GF.windows_by_contig(GF.read_vcf(vc), {case["contig"]: 0}).get(case["contig"], []) == case["windows"]


def test_windows_of_indel_records_match_the_reference_geometry(tmp_path):
    kat = GOLD["windows_kat"]
    vc = str(tmp_path / "k.vcf")
    H.write_vcf(vc, kat["vcf"])
    got = GF.windows_by_contig(GF.read_vcf(vc), {c: k for k, c in enumerate(kat["contigs"])})
    rows = [[c, w["first"], w["last"], w["keep"]["pos"], w["keep"]["end"], w["keep"]["type"], w["keep"]["length"], w["keep"]["allele"]]
            for c, ws in got.items() for w in ws]
    assert rows == kat["windows"]
    gz = str(tmp_path / "k.vcf.gz")
    with gzip.open(gz, "wt") as fh:
        fh.write(open(vc).read())
    assert GF.read_vcf(gz) == GF.read_vcf(vc)
    open(vc, "w").write("##fileformat=VCFv4.2\nc1\t10\t.\tA\tA[c2:77[\t.\tPASS\tSVTYPE=BND\n")
    with pytest.raises(ValueError):                                  # breakends are not restated: loud, not guessed
        GF.read_vcf(vc)
    open(vc, "w").write("c1\t10\t.\tA\t<DEL>\t.\tPASS\t.\n")
    with pytest.raises(ValueError):                                  # a symbolic allele needs END or SVLEN
        GF.read_vcf(vc)


def test_windows_of_structural_records_match_the_reference_geometry(tmp_path):
    """tests/golden: the reference's get_windows on DEL / DUP / INV (near and far ends) / 150 kb DEL / CNV / INS records."""
    kat = GOLD["sv_windows_kat"]
    vc = str(tmp_path / "sv.vcf")
    H.write_vcf(vc, kat["vcf"])
    got = GF.windows_by_contig(GF.read_vcf(vc), {c: k for k, c in enumerate(kat["contigs"])})
    rows = [[c, w["first"], w["last"], w["keep"]["pos"], w["keep"]["end"], w["keep"]["type"], w["keep"]["length"], w["keep"]["allele"]]
            for c, ws in got.items() for w in ws]
    assert rows == kat["windows"]
    from genomeanonymizer_b200 import batch as B2
    t = B2.pack_sessions(got["c1"])                                  # structural keep types never equal a called allele
    assert set(int(x) for x in t.keep_type) <= {2, 3, 99}


def test_multi_contig_bam_flag_filter_small_blocks_and_sort_check(tmp_path):
    case = GENOME[1]["case"]
    reads = [r for r in case["reads"] if r["dataset"] == 0]
    two = [dict(r, contig="a") for r in reads[:200]] + [dict(r, contig="b") for r in reads[200:300]]
    two[5] = dict(two[5], flag=two[5]["flag"] | 0x400)
    p = str(tmp_path / "two.bam")
    H.write_bam(p, [("a", 9000), ("b", 9000), ("empty", 10)], two, block_bytes=777)    # records straddle BGZF blocks
    with GF.BamFile(p) as f:
        assert f.references == ("a", "b", "empty")
        assert [int(f.contig_sizes(c).n_reads) for c in f.references] == [200, 100, 0]
        assert int(f.contig_sizes("a", 0x400).n_reads) == 199
        cb = GF.pack_tumor_normal(f, f, "b")
        ref = B.pack_reads([dict(r, dataset=0) for r in reads[200:300]] + [dict(r, dataset=1) for r in reads[200:300]])
        for k in ARRAYS:
            assert np.array_equal(getattr(cb.batch, k), getattr(ref, k)), k
        cb = GF.pack_tumor_normal(f, f, "a", flag_exclude=0x400)
        assert cb.batch.n_reads == 398 and cb.batch.n_tumor == 199
        assert GF.pack_tumor_normal(f, f, "empty").batch.n_reads == 0
        assert GF.pack_tumor_normal(f, f, "not-there").batch.n_reads == 0
    H.write_bam(p, [("a", 9000)], list(reversed(reads[:50])))
    with GF.BamFile(p) as f:
        assert not f.contig_sizes("a").sorted
        with pytest.raises(ValueError):
            GF.pack_tumor_normal(f, f, "a")


def test_corrupt_and_missing_files_fail_loudly(tmp_path):
    case = GENOME[0]["case"]
    p = str(tmp_path / "x.bam")
    H.write_bam(p, [("c", 100)], [r for r in case["reads"] if r["dataset"] == 0][:20])
    raw = bytearray(open(p, "rb").read())
    raw[40] ^= 0x55                                                     # inside the first block's deflate stream
    open(p, "wb").write(bytes(raw))
    with pytest.raises(GF.GenomeFileError):
        GF.BamFile(p)
    open(p, "wb").write(b"not a bam file at all, just text")
    with pytest.raises(GF.GenomeFileError):
        GF.BamFile(p)
    with pytest.raises(GF.GenomeFileError):
        GF.BamFile(str(tmp_path / "missing.bam"))
    with pytest.raises(GF.GenomeFileError):
        GF.FastaFile(str(tmp_path / "missing.fa"))
    fa = str(tmp_path / "two.fa")
    open(fa, "w").write(">s1 first\r\nACGT\r\nac\r\n>s2\nNNNN\n\nGG")
    F = GF.FastaFile(fa)
    assert F.references == ("s1", "s2") and F.lengths == (6, 6)
    assert F.fetch("s1") == "ACGTac" and F.fetch("s2", 3) == "NGG"
    with pytest.raises(KeyError):
        F.fetch("s3")


# --------------------------------------------------------------------------------------------------------- GPU
def _run_files(tmp, entry, eng):
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import name_output, run_short_read_tumor_normal_anonymizer
    case = entry["case"]
    t, n, fa, vc = H.write_sample_files(tmp, case, entry["vcf"])
    outs = (name_output(t), name_output(n))
    assert outs[0].endswith("T.anonymized")
    run_short_read_tumor_normal_anonymizer([vc], [(t, n)], fa, eng, [outs], True, 2, False)
    return t, n, outs


@pytest.mark.gpu
@pytest.mark.parametrize("entry", GENOME, ids=IDS)
def test_entry_point_writes_the_reference_files_from_bam_input(entry, tmp_path):
    from genomeanonymizer_b200.engine import Engine
    eng = Engine(0)
    try:
        t, n, outs = _run_files(str(tmp_path), entry, eng)
    finally:
        eng.close()
    gold = entry["expected"]["files"]
    for name, text in gold.items():
        path = os.path.join(str(tmp_path), name)
        if text is None:
            assert not os.path.exists(path), name                      # the reference writes no single-end files then
        else:
            assert open(path).read() == text, (entry["case"]["name"], name)


@pytest.mark.gpu
def test_entry_point_with_the_method_object_two_contigs_and_two_samples(tmp_path):
    """Sample 0: contig c = golden case 0, contig d = golden case 1 (one BAM pair, one VCF, one FASTA): every file is
    the concatenation of the two golden samples' files, contig by contig.  Sample 1: case 0 alone against the same
    two-contig FASTA."""
    from genomeanonymizer_b200.anonymizer_methods import B200GermlineAnonymizer
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import run_short_read_tumor_normal_anonymizer
    c0, c1 = GENOME[0]["case"], GENOME[1]["case"]
    contigs = [("c", len(c0["reference"])), ("d", len(c1["reference"]))]
    fa = str(tmp_path / "ref.fa")
    H.write_fasta(fa, [("c", c0["reference"]), ("d", c1["reference"])], width=70)
    reads = [dict(r, contig="c") for r in c0["reads"]] + [dict(r, contig="d") for r in c1["reads"]]
    t, n = str(tmp_path / "T.bam"), str(tmp_path / "N.bam")
    H.write_bam(t, contigs, [r for r in reads if r["dataset"] == 0])
    H.write_bam(n, contigs, [r for r in reads if r["dataset"] == 1])
    vc = str(tmp_path / "s.vcf")
    H.write_vcf(vc, GENOME[0]["vcf"] + [["d"] + v[1:] for v in GENOME[1]["vcf"]])
    d1 = tmp_path / "one"
    d1.mkdir()
    t1, n1, _, vc1 = H.write_sample_files(str(d1), c0, GENOME[0]["vcf"])
    res = run_short_read_tumor_normal_anonymizer([vc, vc1], [(t, n), (t1, n1)], fa, B200GermlineAnonymizer(),
                                                 [(str(tmp_path / "T.out"), str(tmp_path / "N.out")), (str(d1 / "T.out"), str(d1 / "N.out"))],
                                                 True, 1, False)
    assert [r["reads"] for r in res] == [len(reads), len(c0["reads"])]
    g0, g1 = GENOME[0]["expected"]["files"], GENOME[1]["expected"]["files"]
    for p in "TN":
        for m in "12":
            assert open(str(tmp_path / f"{p}.out.{m}.fastq")).read() == g0[f"{p}.anonymized.{m}.fastq"] + g1[f"{p}.anonymized.{m}.fastq"]
            assert open(str(d1 / f"{p}.out.{m}.fastq")).read() == g0[f"{p}.anonymized.{m}.fastq"]
    assert open(n1 + ".statistics.txt").read() == g0["N.bam.statistics.txt"]
    rows = lambda text: [ln for ln in text.split("\n") if ln and ln[0] != "#" and not ln.startswith("outside")]
    stats = open(n + ".statistics.txt").read()
    assert rows(stats) == rows(g0["N.bam.statistics.txt"]) + [ln.replace("c\t", "d\t", 1) for ln in rows(g1["N.bam.statistics.txt"])]


@pytest.mark.gpu
def test_pairs_split_over_two_contigs_are_written_where_the_reference_writes_them(tmp_path):
    """tests/golden "two_contigs": the reference's anonymize_genome on a two-contig sample in which every ninth pair has
    its first mate on c1 and its second on c2.  The reference keeps unpaired reads across contigs and writes the pair
    when the second mate is processed; the entry point (per-contig plans + carried reads) must write the same files."""
    from genomeanonymizer_b200.engine import Engine
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import name_output, run_short_read_tumor_normal_anonymizer
    two = GOLD["two_contigs"]
    a, b = two["cases"]
    contigs = [(a["contig"], len(a["reference"])), (b["contig"], len(b["reference"]))]
    reads = [dict(r, contig=a["contig"]) for r in a["reads"]] + [dict(r, contig=b["contig"]) for r in b["reads"]]
    t, n = str(tmp_path / "T.bam"), str(tmp_path / "N.bam")
    H.write_bam(t, contigs, [r for r in reads if r["dataset"] == 0])
    H.write_bam(n, contigs, [r for r in reads if r["dataset"] == 1])
    fa, vc = str(tmp_path / "ref.fa"), str(tmp_path / "s.vcf")
    H.write_fasta(fa, [(a["contig"], a["reference"]), (b["contig"], b["reference"])])
    H.write_vcf(vc, two["vcf"])
    eng = Engine(0)
    try:
        run_short_read_tumor_normal_anonymizer([vc], [(t, n)], fa, eng, [(name_output(t), name_output(n))], True, 2, False)
    finally:
        eng.close()
    for name, text in two["expected"]["files"].items():
        path = os.path.join(str(tmp_path), name)
        if text is None:
            assert not os.path.exists(path), name
        else:
            assert open(path).read() == text, name


def test_large_files_are_read_one_contig_at_a_time(tmp_path, monkeypatch):
    """Above GA_BAM_EAGER_BYTES the reader keeps the file mapped and inflates the records of one reference when they
    are asked for (found by one windowed walk at open time); here the threshold is 0 and the window smaller than the
    file, with records straddling both BGZF blocks and windows.  Same arrays as the eager reader; a file whose
    references are interleaved falls back to the eager mode."""
    case = GENOME[1]["case"]
    reads = [r for r in case["reads"] if r["dataset"] == 0]
    three = [dict(r, contig="a") for r in reads[:220]] + [dict(r, contig="b") for r in reads[220:300]] + [dict(r, contig="d") for r in reads[300:420]]
    p = str(tmp_path / "three.bam")
    H.write_bam(p, [("a", 9000), ("b", 9000), ("empty", 10), ("d", 9000)], three, block_bytes=3001)
    with GF.BamFile(p) as f:
        eager = {c: GF.pack_tumor_normal(f, f, c) for c in ("a", "b", "empty", "d")}
        n_all = f.n_records
    monkeypatch.setenv("GA_BAM_EAGER_BYTES", "0")
    monkeypatch.setenv("GA_BAM_WINDOW_BYTES", "65536")
    with GF.BamFile(p, 2) as f:
        assert f.references == ("a", "b", "empty", "d") and f.n_records == n_all
        assert [int(f.contig_sizes(c).n_reads) for c in f.references] == [220, 80, 0, 120]
        for c in ("d", "a", "empty", "b", "a"):                      # any order, twice
            lazy = GF.pack_tumor_normal(f, f, c)
            for k in ARRAYS:
                assert np.array_equal(getattr(lazy.batch, k), getattr(eager[c].batch, k)), (c, k)
            assert np.array_equal(lazy.name_blob, eager[c].name_blob) and np.array_equal(lazy.ref_end, eager[c].ref_end)
    mixed = [dict(r, contig="a") for r in reads[:50]] + [dict(r, contig="b") for r in reads[50:90]] + [dict(r, contig="a") for r in reads[90:120]]
    H.write_bam(p, [("a", 9000), ("b", 9000)], mixed, block_bytes=3001)
    with GF.BamFile(p) as f:                                         # references interleaved: the whole stream is kept
        assert [int(f.contig_sizes(c).n_reads) for c in f.references] == [80, 40]
        assert GF.pack_tumor_normal(f, f, "b").batch.n_reads == 80


@pytest.mark.parametrize("flag", [0x800])
def test_entry_point_refuses_supplementary_records(tmp_path, flag):
    """The reference's supplementary-alignment bookkeeping (anonymizer_methods.py:98-149) is not implemented: the entry
    point says so instead of treating such records as ordinary alignments (no engine is needed to find out - the check
    runs on the packed flags before any device work)."""
    from genomeanonymizer_b200 import genome_files as GF
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import anonymize_genome
    case = dict(GENOME[0]["case"])
    reads = [dict(r) for r in case["reads"]]
    reads[5]["flag"] |= flag
    case["reads"] = reads
    t, n, fa, vc = H.write_sample_files(str(tmp_path), case, GENOME[0]["vcf"])
    f = GF.FastaFile(fa)
    order = {name: k for k, name in enumerate(f.references)}
    f.close()
    windows = GF.windows_by_contig(GF.read_vcf(vc), order)
    with pytest.raises(ValueError, match="supplementary"):
        anonymize_genome(windows, t, n, fa, None, str(tmp_path / "T.out"), str(tmp_path / "N.out"))
    # secondary records are first-wins in the reference too: not refused by the flag check
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import _refuse_unsupported_records
    import numpy as np
    _refuse_unsupported_records(np.array([(0x100 | 0x1 | 0x40) << 16 | 150, (0x4 | 0x1 | 0x80) << 16 | 150], np.uint32), "c")


# ------------------------------------------------------------------------------------- format conformance (CPU)
def _spec_bam_record(ref_id, pos, mapq, flag, name, cigar_ops, seq, qual, aux=b"", next_ref=-1, next_pos=-1, tlen=0, bin_=None):
    """One BAM alignment record laid out field by field from SAM v1.6 section 4.2 (written separately from
    tests/helpers.bam_record on purpose): block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq,
    next_refID, next_pos, tlen, read_name, cigar, seq (high nibble first), qual, auxiliary fields."""
    import struct
    rn = name.encode() + b"\0"
    cig = b"".join(struct.pack("<I", (n << 4) | "MIDNSHP=X".index(op)) for n, op in cigar_ops)
    codes = ["=ACMGRSVTWYHKDBN".index(c) for c in seq]
    sq = bytes(((codes[k] << 4) | (codes[k + 1] if k + 1 < len(codes) else 0)) for k in range(0, len(codes), 2))
    span = sum(n for n, op in cigar_ops if op in "MDN=X")
    if bin_ is None:
        bin_ = H._reg2bin(pos, pos + max(span, 1))
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(rn), mapq, bin_, len(cigar_ops), flag, len(seq), next_ref, next_pos, tlen)
    body += rn + cig + sq + bytes(qual) + aux
    return struct.pack("<I", len(body)) + body


def _gzip_member(data: bytes, level=6, extra_before=b"", extra_after=b"") -> bytes:
    """A BGZF member: gzip header with FEXTRA, the BC subfield (total member size - 1) among other subfields."""
    import struct, zlib
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = co.compress(data) + co.flush()
    xlen = len(extra_before) + 6 + len(extra_after)
    bsize = 12 + xlen + len(cdata) + 8
    extra = extra_before + b"BC" + struct.pack("<HH", 2, bsize - 1) + extra_after
    return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", xlen) + extra + cdata +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def test_bam_features_htslib_writes_are_read(tmp_path):
    """Auxiliary fields of every type behind the qualities, mate fields, bin 0, '*' qualities (0xff), other gzip extra
    subfields beside BC, an empty member in mid-file, stored (level 0) members, records and header cut anywhere by
    member boundaries, @PG / @RG / @CO header lines: none of it may change what the reader hands to the engine."""
    import struct
    rng = np.random.default_rng(5)
    reads = []
    for k in range(60):
        L = int(rng.integers(30, 120))
        seq = "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, L))
        shape = k % 4
        if shape == 0: ops = [(L, "M")]
        elif shape == 1: ops = [(5, "S"), (L - 5, "M")]
        elif shape == 2: ops = [(10, "M"), (3, "I"), (L - 13, "M")]
        else: ops = [(3, "H"), (12, "="), (4, "D"), (L - 12, "X"), (2, "H")]
        reads.append(dict(name=f"r{k:03d}/x", flag=[99, 147, 83, 163, 1024 + 99][k % 5], pos=100 + 7 * k, ops=ops, seq=seq,
                          qual=[0xff] * L if k % 7 == 3 else [int(q) for q in rng.integers(2, 41, L)]))
    aux_all = (b"NMC\x03" + b"MDZ10A5^AC6\0" + b"RGZgrp1\0" + b"XSc\xfe" + b"XTs" + struct.pack("<h", -300) + b"XUS" + struct.pack("<H", 60000) +
               b"XIi" + struct.pack("<i", -70000) + b"XVI" + struct.pack("<I", 4000000000) + b"XFf" + struct.pack("<f", 1.5) + b"XAA!" +
               b"XHH1AE301\0" + b"XBBc" + struct.pack("<I", 3) + b"\x01\x02\xff" + b"XCBS" + struct.pack("<I", 2) + struct.pack("<HH", 1, 65535) +
               b"XDBf" + struct.pack("<I", 1) + struct.pack("<f", 2.25) + b"XEBi" + struct.pack("<I", 0))
    text = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ctgA\tLN:5000\tM5:abc\n@SQ\tSN:ctgB\tLN:300\n@RG\tID:grp1\tSM:s\n"
            "@PG\tID:bwa\tPN:bwa\tCL:bwa mem -Y x y\n@CO\tfree text\twith tabs\n")
    stream = bytearray(b"BAM\1" + struct.pack("<I", len(text)) + text.encode() + struct.pack("<I", 2))
    for nm, ln in (("ctgA", 5000), ("ctgB", 300)):
        stream += struct.pack("<I", len(nm) + 1) + nm.encode() + b"\0" + struct.pack("<I", ln)
    for k, r in enumerate(reads):
        stream += _spec_bam_record(0, r["pos"], 255 if k % 3 else 0, r["flag"], r["name"], r["ops"], r["seq"], r["qual"],
                                   aux=aux_all if k % 2 else b"", next_ref=0 if k % 2 else -1, next_pos=r["pos"] + 200, tlen=(-1) ** k * 350,
                                   bin_=0 if k % 4 == 0 else None)
    p = str(tmp_path / "spec.bam")
    with open(p, "wb") as fh:
        cuts = [0, 7, 64, 65, 900, 901, 2500, 2501, 2502, len(stream) // 2, len(stream) - 3, len(stream)]   # header and records cut anywhere
        for i, (a, b) in enumerate(zip(cuts, cuts[1:])):
            fh.write(_gzip_member(bytes(stream[a:b]), level=0 if i % 3 == 0 else 9,
                                  extra_before=b"RA\x04\x00abcd" if i % 2 else b"", extra_after=b"ZZ\x01\x00q" if i % 4 == 1 else b""))
            if i == 4:
                fh.write(_gzip_member(b""))                              # an empty member in mid-file is legal
        fh.write(H._BAM_EOF)
    with GF.BamFile(p, 2) as f:
        assert f.references == ("ctgA", "ctgB") and f.lengths == (5000, 300)
        assert f.n_records == len(reads)
        cb = GF.pack_tumor_normal(f, f, "ctgA")
    as_dict = [dict(name=r["name"], flag=r["flag"], pos=r["pos"], cigar="".join(f"{n}{op}" for n, op in r["ops"]), seq=r["seq"], qual=r["qual"], dataset=0)
               for r in reads]
    ref = B.pack_reads(as_dict + [dict(r, dataset=1) for r in as_dict])
    for k in ARRAYS:
        assert np.array_equal(getattr(cb.batch, k), getattr(ref, k)), k
    assert [cb.name(k) for k in range(len(reads))] == [r["name"] for r in reads]


def test_bam_without_sequence_and_long_cigar_records(tmp_path):
    """A record without a stored sequence ('*': l_seq = 0, as secondary alignments have) is a read of length zero; a
    CIGAR moved to the CG tag (more than 65,535 ops, SAM v1.6 4.2.2) is refused by name, not mis-read."""
    import struct
    hdr = bytearray(b"BAM\1" + struct.pack("<I", 0) + struct.pack("<I", 1) + struct.pack("<I", 2) + b"c\0" + struct.pack("<I", 100000))
    ok = hdr + _spec_bam_record(0, 10, 30, 256, "nosq", [(50, "M")], "", []) + _spec_bam_record(0, 20, 30, 0, "plain", [(4, "M")], "ACGT", [30] * 4)
    p = str(tmp_path / "nosq.bam")
    open(p, "wb").write(_gzip_member(bytes(ok)) + H._BAM_EOF)
    with GF.BamFile(p) as f:
        rows = GF.pack_tumor_normal(f, f, "c").read_table()
    assert [(r["name"], r["pos"], r["end"]) for r in rows[:2]] == [("nosq", 10, 60), ("plain", 20, 24)]
    cg = hdr + _spec_bam_record(0, 10, 30, 0, "long", [(4, "S"), (70000, "N")], "ACGT", [30] * 4,
                                aux=b"CGBI" + struct.pack("<I", 2) + struct.pack("<II", (2 << 4) | 0, (2 << 4) | 4))
    open(p, "wb").write(_gzip_member(bytes(cg)) + H._BAM_EOF)
    with pytest.raises((GF.GenomeFileError, ValueError)):
        with GF.BamFile(p) as f:
            GF.pack_tumor_normal(f, f, "c")


def test_fasta_reader_takes_gzip_and_bgzip_files(tmp_path):
    """pysam.FastaFile reads bgzip-compressed FASTA as well as plain text; the reader takes any sequence of gzip members
    (one member: gzip, many 64 KiB members + the empty EOF member: bgzip) and refuses a truncated or corrupt stream."""
    import gzip
    rng = np.random.default_rng(5)
    contigs = [("chrA", "".join(rng.choice(list("ACGTNacgt"), size=150001))), ("chrB some description", "".join(rng.choice(list("ACGT"), size=70))),
               ("chrC", "")]
    plain = str(tmp_path / "ref.fa")
    H.write_fasta(plain, contigs, width=61)
    raw = open(plain, "rb").read()
    one = str(tmp_path / "ref.fa.gz")
    with open(one, "wb") as f:
        f.write(gzip.compress(raw))
    many = str(tmp_path / "ref.bgz.fa.gz")
    with open(many, "wb") as f:
        for k in range(0, len(raw), 0xff00):
            f.write(H._bgzf_block(raw[k:k + 0xff00]))
        f.write(H._bgzf_block(b""))
    want = GF.FastaFile(plain)
    for path in (one, many):
        got = GF.FastaFile(path)
        assert got.references == want.references == ("chrA", "chrB", "chrC") and got.lengths == want.lengths
        for name in got.references:
            assert got.fetch(name) == want.fetch(name)
        assert got.fetch("chrA", 149990, 150010) == want.fetch("chrA", 149990, 150010)
        got.close()
    cut = str(tmp_path / "cut.fa.gz")
    with open(cut, "wb") as f:
        f.write(open(many, "rb").read()[:40000])
    with pytest.raises(GF.GenomeFileError):
        GF.FastaFile(cut)
    bad = bytearray(open(one, "rb").read())
    bad[len(bad) // 2] ^= 0x55
    with open(cut, "wb") as f:
        f.write(bytes(bad))
    with pytest.raises(GF.GenomeFileError):
        GF.FastaFile(cut)


@pytest.mark.parametrize("entry", [GENOME[1], GENOME[5]], ids=[IDS[1], IDS[5]])
@pytest.mark.parametrize("crlf", [False, True], ids=["lf", "crlf"])
def test_sam_text_packs_the_same_batch_as_the_bam(entry, crlf, tmp_path):
    """pysam.AlignmentFile opens SAM text as it opens BAM (the reference's name_output knows '.sam'); the reader turns the
    text into the uncompressed BAM stream and everything behind it is shared."""
    case = entry["case"]
    contigs = [(case["contig"], len(case["reference"]))]
    t, n, _, _ = H.write_sample_files(str(tmp_path), case, entry["vcf"])
    ts, ns = str(tmp_path / "T.sam"), str(tmp_path / "N.sam")
    H.write_sam(ts, contigs, [r for r in case["reads"] if r["dataset"] == 0], crlf)
    H.write_sam(ns, contigs, [r for r in case["reads"] if r["dataset"] == 1], crlf)
    with GF.BamFile(t, 2) as T, GF.BamFile(n, 2) as N, GF.BamFile(ts, 2) as TS, GF.BamFile(ns, 2) as NS:
        assert TS.references == T.references and TS.lengths == T.lengths and (TS.n_records, NS.n_records) == (T.n_records, N.n_records)
        a, b = GF.pack_tumor_normal(T, N, case["contig"]), GF.pack_tumor_normal(TS, NS, case["contig"])
    for f in ARRAYS:
        assert np.array_equal(getattr(a.batch, f), getattr(b.batch, f)), f
    assert a.batch.n_tumor == b.batch.n_tumor and a.batch.max_ref_span == b.batch.max_ref_span
    assert np.array_equal(a.ref_end, b.ref_end) and np.array_equal(a.name_off, b.name_off) and np.array_equal(a.name_blob, b.name_blob)


def test_sam_text_edge_cases_and_errors(tmp_path):
    contigs = [("c", 1000), ("d", 500)]
    reads = [dict(name="r1", flag=0x41, pos=10, cigar="3S5M2I4M1D3M2H", seq="ACGTNACGTTACGTACG", qual=list(range(17)), contig="c"),
             dict(name="r2", flag=0x91, pos=20, cigar="4M", seq="", qual=[], contig="c"),                      # no stored sequence
             dict(name="r3", flag=0x81, pos=5, cigar="6M", seq="acgtRY", qual=[40] * 6, contig="d")]            # lower case, ambiguity codes
    sam, bam = str(tmp_path / "x.sam"), str(tmp_path / "x.bam")
    H.write_sam(sam, contigs, reads)
    H.write_bam(bam, contigs, [dict(r, seq=r["seq"].upper()) for r in reads])
    with GF.BamFile(sam, 1) as S, GF.BamFile(bam, 1) as Bm:
        assert S.references == ("c", "d") and S.lengths == (1000, 500) and S.n_records == 3
        for contig in ("c", "d"):
            a, b = GF.pack_tumor_normal(S, S, contig), GF.pack_tumor_normal(Bm, Bm, contig)
            for f in ARRAYS:
                assert np.array_equal(getattr(a.batch, f), getattr(b.batch, f)), (contig, f)
    head = "@SQ\tSN:c\tLN:100\n"
    for body in ("r\t0\tzz\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n",            # a reference without @SQ
                 "r\t0\tc\t1\t60\t4M\t*\t0\t0\tACGT\n",                   # ten fields
                 "r\t0\tc\t1\t60\t4Q\t*\t0\t0\tACGT\tIIII\n",             # CIGAR
                 "r\t0\tc\t1\t60\tM\t*\t0\t0\tACGT\tIIII\n",
                 "r\t0\tc\t1\t60\t4M\t*\t0\t0\tACGT\tIII\n"):             # SEQ / QUAL lengths
        path = str(tmp_path / "bad.sam")
        with open(path, "w") as fh:
            fh.write(head + body)
        with pytest.raises(GF.GenomeFileError):
            GF.BamFile(path, 1)
    with open(str(tmp_path / "only_header.sam"), "w") as fh:
        fh.write(head)
    with GF.BamFile(str(tmp_path / "only_header.sam"), 1) as E:
        assert E.n_records == 0 and E.references == ("c",)


@pytest.mark.gpu
def test_entry_point_reads_text_alignment_files(tmp_path):      # no "sam" in the test name: name_output's unescaped pattern (SR.py:55-58) would rewrite the temporary directory
    from genomeanonymizer_b200.engine import Engine
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import name_output, run_short_read_tumor_normal_anonymizer
    entry = GENOME[1]
    case = entry["case"]
    contigs = [(case["contig"], len(case["reference"]))]
    _, _, fa, vc = H.write_sample_files(str(tmp_path), case, entry["vcf"])
    ts, ns = str(tmp_path / "T.sam"), str(tmp_path / "N.sam")
    H.write_sam(ts, contigs, [r for r in case["reads"] if r["dataset"] == 0])
    H.write_sam(ns, contigs, [r for r in case["reads"] if r["dataset"] == 1])
    outs = (name_output(ts), name_output(ns))
    assert outs[0].endswith("T.anonymized")
    eng = Engine(0)
    try:
        run_short_read_tumor_normal_anonymizer([vc], [(ts, ns)], fa, eng, [outs], True, 2, False)
    finally:
        eng.close()
    for name, text in entry["expected"]["files"].items():
        path = os.path.join(str(tmp_path), name.replace("N.bam.statistics", "N.sam.statistics"))
        if text is None:
            assert not os.path.exists(path), name
        else:
            assert open(path).read() == text, name
