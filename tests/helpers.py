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
