"""The CPU oracle (oracle/ga_oracle.c) against every golden vector produced by the reference itself."""
import pytest

from genomeanonymizer_b200 import batch as B
from oracle import oracle
from tests import helpers as H

GOLD = H.load_golden("session_cases.json")["cases"]


@pytest.mark.parametrize("entry", GOLD, ids=[e["case"]["name"] for e in GOLD])
@pytest.mark.parametrize("sparse_qual", [False, True], ids=["dense-qual", "sparse-qual"])
def test_oracle_matches_reference_sessions(entry, sparse_qual):
    case = entry["case"]
    batch = B.pack_reads(H.ordered_reads(case), sparse_qual=sparse_qual)
    if sparse_qual:
        # helpers read qualities through the dense layout; keep a dense twin for decoding
        dense = B.pack_reads(H.ordered_reads(case))
    for widx, (w, exp) in enumerate(zip(case["windows"], entry["expected"])):
        sessions = B.pack_sessions([w])
        res, st = oracle.run(batch, sessions, case["reference"])
        assert st == 0
        H.check_session_against_golden(case, widx, exp, dense if sparse_qual else batch, res)


def test_oracle_thread_count_invariant():
    entry = next(e for e in GOLD if e["case"]["name"].startswith("rand-13"))
    case = entry["case"]
    batch = B.pack_reads(H.ordered_reads(case))
    sessions = B.pack_sessions(case["windows"])
    r1, s1 = oracle.run(batch, sessions, case["reference"], threads=1)
    r4, s4 = oracle.run(batch, sessions, case["reference"], threads=4)
    assert s1 == 0 and s4 == 0
    assert r1.totals == r4.totals
    assert sorted(r1.records) == sorted(r4.records)
    for k in r1.records:
        assert (r1.records[k]["seq"] == r4.records[k]["seq"]).all()


def test_multi_session_table_equals_single_sessions():
    """Sessions are independent: a table with both windows gives the same per-session records as two runs."""
    entry = next(e for e in GOLD if e["case"]["name"].startswith("rand-13"))
    case = entry["case"]
    batch = B.pack_reads(H.ordered_reads(case))
    both, st = oracle.run(batch, B.pack_sessions(case["windows"]), case["reference"])
    assert st == 0
    for widx, w in enumerate(case["windows"]):
        one, st = oracle.run(batch, B.pack_sessions([w]), case["reference"])
        assert st == 0
        mine = {r: v for (s, r), v in both.records.items() if s == widx}
        assert sorted(mine) == sorted(r for (_, r) in one.records)
        for r, v in mine.items():
            assert (v["seq"] == one.records[(0, r)]["seq"]).all()
        assert (both.sess_counts[widx] == one.sess_counts[0]).all()
