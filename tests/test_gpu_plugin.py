"""The reference-facing plugin (B200GermlineAnonymizer, same call shape as CompleteGermlineAnonymizer.anonymize,
anonymizer_methods.py:431-535) against the golden vectors produced by the reference itself: yield order,
internal sequence / quality arrays, FASTQ records and the statistics counters."""
import pickle

import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

GOLD = H.load_golden("session_cases.json")["cases"]


class Recorder:
    """Stand-in for AnonymizedVariantsStatistics.count_variant (SR.py:198-204)."""

    def __init__(self):
        self.counts = [0] * 8

    def count_variant(self, v):
        self.counts[v.variant_type.value - 1] += 1


class Keep:
    """Duck-typed CalledGenomicVariant (variants.py:42-62) for variant_to_keep."""

    def __init__(self, contig, k):
        from genomeanonymizer_b200.anonymizer_methods import VariantType
        self.seq_name, self.pos, self.end, self.length, self.allele = contig, k["pos"], k["end"], k["length"], k["allele"]
        self.variant_type = VariantType[k["type"]]


@pytest.fixture(scope="module")
def anonymizer():
    from genomeanonymizer_b200.anonymizer_methods import B200GermlineAnonymizer
    return B200GermlineAnonymizer()


@pytest.mark.parametrize("entry", GOLD, ids=[e["case"]["name"] for e in GOLD])
def test_plugin_reproduces_reference_sessions(anonymizer, entry):
    stub = H.load_pysam_stub()
    case = entry["case"]
    contig = case["contig"]
    seg = lambda r: stub.AlignedSegment(r["name"], r["flag"], contig, r["pos"], r["cigar"], r["seq"], r["qual"])
    t = [seg(r) for r in case["reads"] if r["dataset"] == 0]
    n = [seg(r) for r in case["reads"] if r["dataset"] == 1]
    stub.register_fasta("ref.fa", {contig: case["reference"]})
    fasta = stub.FastaFile("ref.fa")
    for w, exp in zip(case["windows"], entry["expected"]):
        keep = Keep(contig, w["keep"]) if w.get("keep") else None
        rec = Recorder()
        pile = H.iter_pileups_stub(stub, t, n, contig, w["first"], w["last"])
        order, got = [], {}
        for pair in anonymizer.anonymize(keep, pile, fasta, stats_recorder=rec):
            assert len(pair) == 2
            order.append(next(a.query_name for a in pair if a is not None))
            for a in pair:
                if a is None:
                    continue
                mate = 1 if a.is_read1 else 2
                assert a.get_pair_idx() == mate - 1 and a.anonymized_read_is_complete() and not a.has_left_overs_to_mask
                seq = bytes(bytearray(int(x) for x in a.anonymized_sequence_array)).decode()
                qual = [int(x) for x in a.anonymized_qualities_array]
                got[f"{a.dataset_idx}|{a.query_name}|{mate}"] = {"seq": seq, "qual_internal": qual,
                                                                 "fastq": a.get_anonymized_fastq_record()}
        assert order == exp["yield_order"], case["name"]
        assert sorted(got) == sorted(exp["reads"]), case["name"]
        for k, e in exp["reads"].items():
            assert got[k]["seq"] == e["seq"].upper(), (case["name"], k)
            assert got[k]["qual_internal"] == e["qual_internal"], (case["name"], k)
            assert got[k]["fastq"] == e["fastq"], (case["name"], k)
        assert rec.counts == exp["counts"], case["name"]
        assert anonymizer.anonymized_reads == {}                    # reset() at session end (AM.py:534)


def test_plugin_is_picklable_and_lazy(anonymizer):
    clone = pickle.loads(pickle.dumps(anonymizer))               # the reference pickles it into pool workers (SR.py:953-959)
    assert clone._engine is None and clone.device == anonymizer.device
