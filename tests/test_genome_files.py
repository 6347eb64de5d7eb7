"""N2 + N3, file level: the reference's four FASTQ files, byte for byte.

tests/golden/genome_cases.json holds what the reference's own anonymize_genome (short_read_tumor_normal_anonymizer.py:
625-760, executed under functional pysam / variant_extractor fakes) wrote for whole samples.  The batched path -
genomeanonymizer_b200.driver.plan_sample (sections, island sessions, mate pairing, first write wins) + one masking
pass over all sessions + FASTQ rendering - must write the same files and count the same variants.
CPU: masking by the oracle, rendering by oracle/fastq.py.  GPU: ga_run + ga_fastq_render through the C ABI."""
import re

import numpy as np
import pytest

from genomeanonymizer_b200 import batch as B
from genomeanonymizer_b200 import driver as D
from genomeanonymizer_b200 import synth
from oracle import fastq as OF
from oracle import oracle
from tests import helpers as H

GENOME = H.load_golden("genome_cases.json")["cases"]
_CIG = re.compile(r"(\d+)([MIDNSHP=X])")


def with_ends(reads):
    out = []
    for r in reads:
        span = sum(int(n) for n, op in _CIG.findall(r["cigar"]) if op in "MDN=X")
        out.append(dict(r, end=r["pos"] + span))
    return out


def golden_counts(files):
    rows = []
    for ln in files["N.bam.statistics.txt"].split("\n"):
        if not ln or ln.startswith("#"):
            continue
        f = ln.split("\t")
        rows.append((f[0], [int(x) for x in f[3:6]]))
    return rows


def assemble(plan, reads, text_of):
    """The six output files from the write plan; text_of(read index, version) -> record text.  The version reaches text_of
    without the plan's REAPPLY flag (quirk Q12): the checker gets the flagged pairs through driver.reapply_pairs."""
    text_of = (lambda f: lambda r, v: f(int(r), D.version_of(v)))(text_of)
    files = {f"{p}.anonymized.{s}.fastq": [] for p in "TN" for s in ("1", "2", "single_end")}
    for ds, r1, v1, r2, v2 in plan.pairs:
        p = "TN"[ds]
        files[f"{p}.anonymized.1.fastq"].append(text_of(r1, v1))
        files[f"{p}.anonymized.2.fastq"].append(text_of(r2, v2))
    for ds, r, v in plan.singles:
        files[f"{'TN'[ds]}.anonymized.single_end.fastq"].append(text_of(r, v))
    return {k: "".join(v) for k, v in files.items()}


def check_files(entry, files, plan, sess_counts):
    gold = entry["expected"]["files"]
    for name, text in files.items():
        assert text == (gold.get(name) or ""), (entry["case"]["name"], name)
    rows = golden_counts(gold)
    mine_windows = [[int(x) for x in sess_counts[s][:3]] for s, ses in enumerate(plan.sessions) if ses["window"] is not None]
    outside = np.sum([sess_counts[s][:3] for s, ses in enumerate(plan.sessions) if ses["window"] is None], axis=0) if any(
        ses["window"] is None for ses in plan.sessions) else np.zeros(3, int)
    assert mine_windows == [c for n, c in rows if n != "outside_windows"], entry["case"]["name"]
    assert [int(x) for x in outside] == [c for n, c in rows if n == "outside_windows"][0], entry["case"]["name"]


def test_window_and_section_arithmetic_matches_reference():
    kat = H.load_golden("genome_cases.json")["windows_kat"]
    ws = [w for w in kat["windows"] if w[0] == "c1"]
    for rec, w in zip([v for v in kat["vcf"] if v[0] == "c1"], ws):
        assert list(D.window_of_variant(rec[1], rec[2])) == w[1:3]
    secs = D.genome_sections([{"first": w[1], "last": w[2]} for w in ws], kat["contigs"]["c1"])
    assert [[a, b, k is not None] for a, b, k in secs] == [s[1:] for s in kat["sections"] if s[0] == "c1"]
    assert D.genome_sections([], 5000) == [(0, 0, None)]


@pytest.mark.parametrize("entry", GENOME, ids=[e["case"]["name"] for e in GENOME])
def test_oracle_writes_the_reference_files(entry):
    case = entry["case"]
    reads = with_ends(H.ordered_reads(case))
    plan = D.plan_sample(reads, case["windows"], len(case["reference"]))
    assert [[s["first"], s["last"]] for s in plan.sessions if s["window"] is not None] == [w[1:] for w in entry["expected"]["windows"]]
    batch = B.pack_reads(reads)
    res, st = oracle.run(batch, B.pack_sessions(plan.sessions), case["reference"], reapply=D.reapply_pairs(plan))
    assert st == 0

    def text_of(i, version):
        if version >= 0 and (version, i) in res.records:
            seq, qual = H.final_read(batch, res, i, session=version)
        else:
            seq, qual = B.decode_bases(batch.sequence_codes(i)), [int(x) for x in batch.qualities(i)]
        return OF.render(reads[i]["name"], reads[i]["flag"], seq, qual)

    check_files(entry, assemble(plan, reads, text_of), plan, res.sess_counts)
    assert D.statistics_text(case["contig"], plan, res.sess_counts) == entry["expected"]["files"]["N.bam.statistics.txt"]


@pytest.mark.gpu
@pytest.mark.parametrize("entry", GENOME, ids=[e["case"]["name"] for e in GENOME])
def test_device_writes_the_reference_files(entry):
    """driver.anonymize_sample: plan + ga_run + ga_fastq_render = the reference's six files and its statistics file."""
    from genomeanonymizer_b200.engine import Engine
    case = entry["case"]
    eng = Engine(0)
    try:
        got = D.anonymize_sample(eng, H.ordered_reads(case), case["windows"], case["reference"], contig=case["contig"])
    finally:
        eng.close()
    gold = entry["expected"]["files"]
    for p in "TN":
        for s in ("1", "2", "single_end"):
            assert got[f"{p}.{s}"] == (gold.get(f"{p}.anonymized.{s}.fastq") or ""), (case["name"], p, s)
    assert got["statistics"] == gold["N.bam.statistics.txt"], case["name"]


def device_sample_equals_plan_and_oracle(eng, case):
    """driver.anonymize_sample on the device against plan_sample + oracle (flagged reads masked twice) + the oracle's
    renderer; returns the number of flagged (session, read) pairs."""
    reads = with_ends(H.ordered_reads(case))
    plan = D.plan_sample(reads, case["windows"], len(case["reference"]))
    flagged = D.reapply_pairs(plan)
    batch = B.pack_reads(reads)
    res, st = oracle.run(batch, B.pack_sessions(plan.sessions), case["reference"], reapply=flagged)
    assert st == 0

    def text_of(i, version):
        if version >= 0 and (version, i) in res.records:
            seq, qual = H.final_read(batch, res, i, session=version)
        else:
            seq, qual = B.decode_bases(batch.sequence_codes(i)), [int(x) for x in batch.qualities(i)]
        return OF.render(reads[i]["name"], reads[i]["flag"], seq, qual)
    want = assemble(plan, reads, text_of)
    got = D.anonymize_sample(eng, H.ordered_reads(case), case["windows"], case["reference"], contig=case["contig"])
    for p in "TN":
        for s in ("1", "2", "single_end"):
            assert got[f"{p}.{s}"] == want[f"{p}.anonymized.{s}.fastq"], (case["name"], p, s)
    assert got["statistics"] == D.statistics_text(case["contig"], plan, res.sess_counts), case["name"]
    return len(flagged)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [50, 54, 57])
def test_device_applies_left_over_indels_twice_where_the_reference_does(seed):
    """Quirk Q12 on the device path: thin, orphan-rich samples in which sessions meet reads that an earlier session parked
    unpaired; driver.anonymize_sample must print what plan + oracle print with the flagged reads masked twice (the pair is
    pinned to the reference's own files by the golden sample genome-twice-masked and by tools/fuzz_genome.py)."""
    from genomeanonymizer_b200.engine import Engine
    case = synth.make_case(seed, name=f"twice-{seed}", contig_len=9000, n_pairs=(70, 60), read_len=80, somatic_positions=[2500, 4650, 6800],
                           snp_rate=4e-3, indel_rate=3e-3, clip_frac=0.1)
    case["reads"] = [r for k, r in enumerate(case["reads"]) if k % 5 != 3]
    eng = Engine(0)
    try:
        assert device_sample_equals_plan_and_oracle(eng, case) > 0
    finally:
        eng.close()
