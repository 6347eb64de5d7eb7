"""Shared test helpers: golden loading and result comparison."""
import json
import os

import numpy as np

from genomeanonymizer_b200 import batch as B

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def case_key(read):
    mate = 1 if read["flag"] & 0x40 else 2
    return f"{read['dataset']}|{read['name']}|{mate}"


def ordered_reads(case):
    """Reads in batch order (tumor then normal, file order inside each)."""
    rs = case["reads"]
    return [r for r in rs if r["dataset"] == 0] + [r for r in rs if r["dataset"] == 1]


def final_read(batch, result, r, session=0):
    """(sequence string in alignment orientation, printed-order qualities list) of batch read r as masked by `session`."""
    rec = result.records.get((session, r))
    if rec is None:
        return B.decode_bases(batch.sequence_codes(r)), [int(x) for x in batch.qualities(r)]
    seq = B.decode_bases(rec["seq"])
    q = rec["qual"] if rec["qual"] is not None else batch.qualities(r)
    return seq, [int(x) for x in q]


def check_session_against_golden(case, widx, expected, batch, result):
    """expected: golden dict of ONE session; result: MaskResult of a run whose table held only that window."""
    reads = ordered_reads(case)
    keys = [case_key(r) for r in reads]
    exp_reads = expected["reads"]
    seen = 0
    for i, k in enumerate(keys):
        seq, qual = final_read(batch, result, i)
        if k in exp_reads:
            seen += 1
            e = exp_reads[k]
            assert seq == e["seq"].upper(), (case["name"], widx, k, seq, e["seq"])
            printed = [ord(c) - 33 for c in e["fastq"].split("\n")[3]]
            assert qual == printed, (case["name"], widx, k, qual, printed)
        else:
            assert (0, i) not in result.records, (case["name"], widx, k, "read outside the session was modified")
    assert seen == len(exp_reads), (case["name"], "golden reads missing from the batch")
    got = [int(x) for x in result.sess_counts[0, :3]]
    assert got == expected["counts"][:3], (case["name"], widx, got, expected["counts"])
    assert int(result.sess_counts[0, 3]) == len(exp_reads), (case["name"], widx, "session read count")


def load_pysam_stub():
    """The duck-typed pysam stand-in (tests/ref_stub/pysam.py) as a module, without touching sys.modules['pysam']."""
    import importlib.util
    path = os.path.join(HERE, "ref_stub", "pysam.py")
    spec = importlib.util.spec_from_file_location("ga_pysam_stub", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def iter_pileups_stub(stub, tumor_reads, normal_reads, contig, start, stop):
    """Two pileups merged by reference_pos into (colT|None, colN|None) pairs - restatement of the reference's
    pileup_io.iter_pileups (pileup_io.pyx:8-41) over the stub's emulated pileups."""
    it_t = stub.emulate_pileup(tumor_reads, contig, start, stop)
    it_n = stub.emulate_pileup(normal_reads, contig, start, stop)
    ct, cn = next(it_t, None), next(it_n, None)
    while ct is not None or cn is not None:
        if cn is None or (ct is not None and ct.reference_pos < cn.reference_pos):
            yield ct, None
            ct = next(it_t, None)
        elif ct is None or cn.reference_pos < ct.reference_pos:
            yield None, cn
            cn = next(it_n, None)
        else:
            yield ct, cn
            ct, cn = next(it_t, None), next(it_n, None)


# ---------------------------------------------------------------------------------------------------------------
# Genome files for the N4 tests: a minimal BAM / FASTA / VCF WRITER.  Test infrastructure only - the
# product reads these formats (csrc/ga_genome_io.cpp), it never writes them.  Written from the SAM/BAM specification
# (BGZF members of <= 64 KiB with the 'BC' subfield, the 28-byte EOF marker, little-endian alignment records).
import re as _re
import struct as _struct
import zlib as _zlib

_CIGAR_RE = _re.compile(r"(\d+)([MIDNSHP=X])")
_BAM_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data: bytes, level: int = 6) -> bytes:
    co = _zlib.compressobj(level, _zlib.DEFLATED, -15)
    cdata = co.compress(data) + co.flush()
    bsize = len(cdata) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + _struct.pack("<H", bsize) + cdata +
            _struct.pack("<II", _zlib.crc32(data) & 0xffffffff, len(data)))


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def bam_record(r: dict, ref_id: int) -> bytes:
    ops = [(int(n), "MIDNSHP=X".index(op)) for n, op in _CIGAR_RE.findall(r["cigar"])]
    span = sum(n for n, op in ops if op in (0, 2, 3, 7, 8))
    name = r["name"].encode("ascii") + b"\0"
    seq = r["seq"]
    codes = ["=ACMGRSVTWYHKDBN".index(c) for c in seq.upper()]
    if len(codes) & 1:
        codes.append(0)
    packed = bytes((codes[k] << 4) | codes[k + 1] for k in range(0, len(codes), 2))
    qual = bytes(int(q) for q in r["qual"])
    body = _struct.pack("<iiBBHHHIiii", ref_id, r["pos"], len(name), 60, _reg2bin(r["pos"], r["pos"] + max(span, 1)), len(ops),
                        r["flag"], len(seq), -1, -1, 0)
    body += name + b"".join(_struct.pack("<I", (n << 4) | op) for n, op in ops) + packed + qual
    return _struct.pack("<I", len(body)) + body


def write_bam(path, contigs, reads, block_bytes: int = 0xff00, level: int = 6):
    """contigs: [(name, length)]; reads: dicts with name / flag / pos / cigar / seq / qual and optionally contig
    (default: the first), written in the order given."""
    names = [c for c, _ in contigs]
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{c}\tLN:{n}\n" for c, n in contigs)
    stream = bytearray(b"BAM\1" + _struct.pack("<I", len(text)) + text.encode() + _struct.pack("<I", len(contigs)))
    for c, n in contigs:
        stream += _struct.pack("<I", len(c) + 1) + c.encode() + b"\0" + _struct.pack("<I", n)
    for r in reads:
        stream += bam_record(r, names.index(r.get("contig", names[0])))
    with open(path, "wb") as fh:
        for o in range(0, len(stream), block_bytes):
            fh.write(_bgzf_block(bytes(stream[o:o + block_bytes]), level))
        fh.write(_BAM_EOF)


def write_sam(path, contigs, reads, crlf: bool = False):
    """The same reads as SAM text (SAM v1.6 section 1): header dictionary, eleven mandatory fields, one optional field."""
    names = [c for c, _ in contigs]
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate" + nl + "".join(f"@SQ\tSN:{c}\tLN:{n}{nl}" for c, n in contigs) + "@PG\tID:test" + nl)
        for r in reads:
            unmapped = bool(r["flag"] & 4) and r["cigar"] == "*"
            qual = "".join(chr(int(q) + 33) for q in r["qual"]) or "*"
            fh.write("\t".join([r["name"], str(r["flag"]), r.get("contig", names[0]), str(r["pos"] + 1), "0" if unmapped else "60", r["cigar"], "=", str(r["pos"] + 1), "0",
                                r["seq"] or "*", qual, "NM:i:0"]) + nl)


def write_fasta(path, contigs, width: int = 60):
    """contigs: [(name, bases)]"""
    with open(path, "w") as fh:
        for name, seq in contigs:
            fh.write(f">{name} test contig\n")
            for o in range(0, len(seq), width):
                fh.write(seq[o:o + width] + "\n")


def write_vcf(path, records):
    """records: [contig, pos, end, length, ref, alt, type] rows (the layout of genome_cases.json's "vcf")."""
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for c, pos, end, length, ref, alt, _t in records:
            info = f"END={end};SVLEN={length}" if alt.startswith("<") else "."
            fh.write(f"{c}\t{pos}\t.\t{ref}\t{alt}\t.\tPASS\t{info}\n")


def write_sample_files(tmp, case, vcf, contig_len=None):
    """Tumor / normal BAM, FASTA and VCF of one genome_cases.json case; returns the four paths."""
    import os
    contig = case["contig"]
    contigs = [(contig, contig_len or len(case["reference"]))]
    t, n = os.path.join(tmp, "T.bam"), os.path.join(tmp, "N.bam")
    write_bam(t, contigs, [r for r in case["reads"] if r["dataset"] == 0])
    write_bam(n, contigs, [r for r in case["reads"] if r["dataset"] == 1])
    fa, vc = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "somatic.vcf")
    write_fasta(fa, [(contig, case["reference"])])
    write_vcf(vc, vcf)
    return t, n, fa, vc


_TWIST_OPS = _re.compile(r"(\d+)([MIDNSHP=X])")


def twist_reads(reads):
    """Edge shapes the generator never makes, chosen by reference coordinates so that tumor and normal reads agree on them:
    an insertion as the last / first aligned op, hard clips around the read, a reference skip in the middle of a match."""
    out = []
    for r in reads:
        ops = [(int(n), op) for n, op in _TWIST_OPS.findall(r["cigar"])]
        seq, qual, pos = r["seq"], list(r["qual"]), r["pos"]
        end = pos + sum(n for n, op in ops if op in "MDN=X")
        if ops and ops[-1][1] == "M" and ops[-1][0] > 6 and end % 5 == 0:          # ... kM -> (k-j)M jI : the insertion sits at reference_end
            j = 1 + end % 3
            ops[-1:] = [(ops[-1][0] - j, "M"), (j, "I")]
        elif ops and ops[0][1] == "M" and ops[0][0] > 6 and pos % 7 == 0:          # kM ... -> jI (k-j)M : at reference_start
            j = 1 + pos % 2
            ops[:1] = [(j, "I"), (ops[0][0] - j, "M")]
            pos += j
        elif len(ops) == 1 and ops[0][1] == "M" and ops[0][0] > 20 and pos % 11 == 0:   # aM bN cM : the bases under the skip leave the read
            a, b = 8, 3 + pos % 4
            c = ops[0][0] - a - b
            ops = [(a, "M"), (b, "N"), (c, "M")]
            seq, qual = seq[:a] + seq[a + b:], qual[:a] + qual[a + b:]
        if pos % 3 == 0:
            ops = [(2, "H")] + ops + [(1, "H")]
        out.append(dict(r, pos=pos, cigar="".join(f"{n}{op}" for n, op in ops), seq=seq, qual=qual))
    out.sort(key=lambda r: (r["dataset"], r["pos"]))
    return out
