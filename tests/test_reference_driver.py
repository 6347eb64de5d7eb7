"""Boundary proof inside the reference's OWN driver: the reference's unmodified `anonymize_genome`
(short_read_tumor_normal_anonymizer.py:625-760) - its window / section loop, `anonymize_window`, pairing, first-write-wins
and the FASTQ / statistics writers - is executed with `B200GermlineAnonymizer` passed as its `anonymizer`, and must write
the same seven files it wrote with its own `CompleteGermlineAnonymizer` (tests/golden/genome_cases.json).

This needs the reference sources, which exist in the build container only (no GPU there) and cannot travel to the GPU box
(the package's build backend, poetry-core, and pysam are absent, so it cannot be installed into baseline/_ref).  The proof
is therefore split: HERE the plugin's host side (pileup collection, packing, object reconstruction, yield order, counter
calls) runs inside the reference's driver with the CPU oracle standing in for the device behind the engine's two calls;
on the GPU box tests/test_gpu_plugin.py runs the same plugin class on the CUDA engine against the reference's sessions.
"""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GA_REFERENCE_ROOT", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "GenomeAnonymizer", "pileup_io.pyx")),
                                reason="the reference sources are only mounted in the build container")


class OracleEngine:
    """The two engine calls the plugin makes, answered by the CPU oracle (test infrastructure)."""

    def __init__(self):
        self.reference = None
        self.runs = 0

    def upload_reference(self, contig_id, bases):
        self.reference = bases if isinstance(bases, (bytes, bytearray)) else str(bases).encode("ascii")

    def run(self, batch, sessions, edits=False):
        from oracle import oracle
        res, st = oracle.run(batch, sessions, self.reference, edits=edits)
        assert st == 0
        self.runs += 1
        return res


@pytest.fixture(scope="module")
def ref_modules():
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("pysam", "pileup_io", "variant_extractor", "variant_extractor.variants")}
    sys.path[:0] = [os.path.join(HERE, "ref_stub"), REF]
    for k in saved_mods:
        sys.modules.pop(k, None)
    import pysam, variant_extractor                                                     # noqa: E401  (the stubs)
    from variant_extractor.variants import VariantRecord, VariantType
    from src.GenomeAnonymizer import short_read_tumor_normal_anonymizer as SR
    yield {"pysam": pysam, "vx": variant_extractor, "VariantRecord": VariantRecord, "VariantType": VariantType, "SR": SR}
    sys.path[:] = saved_path
    for k, v in saved_mods.items():
        sys.modules.pop(k, None)
        if v is not None:
            sys.modules[k] = v
    for k in [m for m in sys.modules if m.startswith("src.GenomeAnonymizer") or m == "src"]:
        sys.modules.pop(k, None)


GENOME = json.load(open(os.path.join(HERE, "golden", "genome_cases.json")))


def run_reference_driver(mods, cases, vcf_rows, anonymizer, tmp):
    pysam, vx, SR = mods["pysam"], mods["vx"], mods["SR"]
    seg = lambda r, contig: pysam.AlignedSegment(r["name"], r["flag"], contig, r["pos"], r["cigar"], r["seq"], r["qual"])
    t = [seg(r, c["contig"]) for c in cases for r in c["reads"] if r["dataset"] == 0]
    n = [seg(r, c["contig"]) for c in cases for r in c["reads"] if r["dataset"] == 1]
    pysam.register_alignment_file("T.bam", t, [c["contig"] for c in cases])
    pysam.register_alignment_file("N.bam", n, [c["contig"] for c in cases])
    pysam.register_fasta("ref.fa", {c["contig"]: c["reference"] for c in cases})
    vcf = [mods["VariantRecord"](v[0], v[1], v[2], v[3], v[4], v[5], mods["VariantType"][v[6]]) for v in vcf_rows]
    vx.register_vcf("s.vcf", vcf)
    fasta = pysam.FastaFile("ref.fa")
    windows = SR.get_windows(vx.VariantExtractor("s.vcf"), SR.get_ref_idxs(fasta))
    cwd = os.getcwd()
    os.chdir(tmp)                                                                        # the driver writes <T>_<N>.mem_debug into the cwd (SR.py:633)
    try:
        SR.anonymize_genome(windows, "T.bam", "N.bam", "ref.fa", anonymizer, os.path.join(tmp, "T.anonymized"), os.path.join(tmp, "N.anonymized"),
                            None, None, True, 1)
    finally:
        os.chdir(cwd)
    names = ("T.anonymized.1.fastq", "T.anonymized.2.fastq", "N.anonymized.1.fastq", "N.anonymized.2.fastq",
             "T.anonymized.single_end.fastq", "N.anonymized.single_end.fastq", "N.bam.statistics.txt")
    return {nm: (open(os.path.join(tmp, nm)).read() if os.path.exists(os.path.join(tmp, nm)) else None) for nm in names}


@pytest.mark.parametrize("entry", GENOME["cases"], ids=[e["case"]["name"] for e in GENOME["cases"]])
def test_reference_driver_with_the_plugin_writes_the_reference_files(ref_modules, entry, tmp_path):
    from genomeanonymizer_b200.anonymizer_methods import B200GermlineAnonymizer
    eng = OracleEngine()
    files = run_reference_driver(ref_modules, [entry["case"]], entry["vcf"], B200GermlineAnonymizer(engine=eng), str(tmp_path))
    assert eng.runs >= len(entry["expected"]["windows"])                                 # one engine call per pileup region
    for name, text in entry["expected"]["files"].items():
        assert files[name] == text, (entry["case"]["name"], name)


def test_reference_driver_with_the_plugin_on_two_contigs(ref_modules, tmp_path):
    from genomeanonymizer_b200.anonymizer_methods import B200GermlineAnonymizer
    two = GENOME["two_contigs"]
    files = run_reference_driver(ref_modules, two["cases"], two["vcf"], B200GermlineAnonymizer(engine=OracleEngine()), str(tmp_path))
    for name, text in two["expected"]["files"].items():
        assert files[name] == text, name


@pytest.mark.parametrize("seed,unmap", [(811, 3), (812, 5), (813, 2)])
def test_plan_matches_the_reference_on_samples_with_unmapped_mates(ref_modules, seed, unmap, tmp_path):
    """Live golden: seeded samples whose placed-unmapped mates sit behind OR in front of their mapped partner (an unmapped
    read can then be the first read a region fetches and seeds an island, pileup_io.pyx:150-158) run through the
    reference's own anonymize_genome with its own CompleteGermlineAnonymizer; driver.plan_sample + oracle must write the
    same files."""
    from genomeanonymizer_b200 import batch as B
    from genomeanonymizer_b200 import driver as D
    from genomeanonymizer_b200 import synth
    from oracle import fastq as OF
    from oracle import oracle
    from src.GenomeAnonymizer.anonymizer_methods import CompleteGermlineAnonymizer
    from tests import helpers as H
    from tests.test_genome_files import assemble, with_ends
    from tests.test_plan_native import unmap_some
    case = synth.make_case(seed, name=f"live-{seed}", contig_len=9000, n_pairs=(90, 80), read_len=80, somatic_positions=[2400, 6100],
                           snp_rate=3e-3, indel_rate=8e-4, clip_frac=0.1)
    case["reads"] = unmap_some([dict(r) for r in case["reads"]], unmap)
    vcf = [["c", w["keep"]["pos"] + 1, w["keep"]["pos"] + 1, 1, case["reference"][w["keep"]["pos"]].upper(), w["keep"]["allele"], "SNV"] for w in case["windows"]]
    gold = run_reference_driver(ref_modules, [case], vcf, CompleteGermlineAnonymizer(), str(tmp_path))
    reads = with_ends(H.ordered_reads(case))
    plan = D.plan_sample(reads, case["windows"], len(case["reference"]))
    batch = B.pack_reads(reads)
    res, st = oracle.run(batch, B.pack_sessions(plan.sessions), case["reference"], reapply=D.reapply_pairs(plan))
    assert st == 0

    def text_of(i, version):
        if version >= 0 and (version, i) in res.records:
            seq, qual = H.final_read(batch, res, i, session=version)
        else:
            seq, qual = B.decode_bases(batch.sequence_codes(i)), [int(x) for x in batch.qualities(i)]
        return OF.render(reads[i]["name"], reads[i]["flag"], seq, qual)
    mine = assemble(plan, reads, text_of)
    assert sum(1 for r in reads if r["flag"] & 4) > 5
    for name, text in mine.items():
        assert text == (gold.get(name) or ""), name
    assert D.statistics_text(case["contig"], plan, res.sess_counts) == gold["N.bam.statistics.txt"]


def test_oracle_equals_the_reference_on_random_and_twisted_samples():
    """A slice of tools/fuzz_reference.py: seeded samples with random shape parameters - and with the edge shapes the
    generator never makes (an insertion as the last / first aligned op, hard clips, reference skips: tests/helpers.twist_reads)
    - go through the reference's unmodified CompleteGermlineAnonymizer.anonymize and through the oracle.  (The whole tool:
    8,000 samples without a difference after the normal-column rule of anonymizer_methods.py:474-481 was restated.)"""
    import subprocess
    for extra in ([], ["--twist"]):
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_reference.py"), "910000", "25"] + extra, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0 and "25 cases, 0 mismatches" in res.stdout, res.stdout[-600:] + res.stderr[-600:]
