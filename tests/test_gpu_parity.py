"""Parity of the CUDA engine (through the C ABI of include/ga_b200.h) with the reference's golden
vectors and with the CPU oracle on seeded random sessions.  Bit-exact: sequence codes, printed
qualities, per-session counters."""
import numpy as np
import pytest

from genomeanonymizer_b200 import batch as B
from genomeanonymizer_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

GOLD = H.load_golden("session_cases.json")["cases"]


@pytest.fixture(scope="module")
def engine():
    from genomeanonymizer_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("entry", GOLD, ids=[e["case"]["name"] for e in GOLD])
@pytest.mark.parametrize("sparse_qual", [False, True], ids=["dense-qual", "sparse-qual"])
def test_engine_matches_reference_sessions(engine, entry, sparse_qual):
    case = entry["case"]
    batch = B.pack_reads(H.ordered_reads(case), sparse_qual=sparse_qual)
    dense = B.pack_reads(H.ordered_reads(case)) if sparse_qual else batch
    engine.upload_reference(0, case["reference"])
    for widx, (w, exp) in enumerate(zip(case["windows"], entry["expected"])):
        res = engine.run(batch, B.pack_sessions([w]))
        H.check_session_against_golden(case, widx, exp, dense, res)


def assert_same_result(a, b, what):
    assert a.totals == b.totals, (what, a.totals, b.totals)
    assert sorted(a.records) == sorted(b.records), what
    for k, v in a.records.items():
        w = b.records[k]
        assert np.array_equal(v["seq"], w["seq"]), (what, k)
        assert (v["qual"] is None) == (w["qual"] is None), (what, k)
        if v["qual"] is not None:
            assert np.array_equal(v["qual"], w["qual"]), (what, k)
    assert np.array_equal(a.sess_counts, b.sess_counts), what


RANDOM_CASES = [
    dict(seed=101, contig_len=6000, n_pairs=(200, 200), read_len=100),
    dict(seed=102, contig_len=9000, n_pairs=(400, 250), read_len=150, somatic_positions=[2500, 4700, 6900]),
    dict(seed=103, contig_len=5000, n_pairs=(300, 300), read_len=75, indel_rate=3e-3, clip_frac=0.5),
    dict(seed=104, contig_len=7000, n_pairs=(250, 250), read_len=151, ref_n_runs=0.05, ref_lower=0.2),
    dict(seed=105, contig_len=8000, n_pairs=(500, 60), read_len=120, somatic_positions=[3000, 3900]),   # overlapping windows
    dict(seed=106, contig_len=4000, n_pairs=(100, 0), read_len=100),                                   # no normal reads
    dict(seed=107, contig_len=12000, n_pairs=(900, 900), read_len=250, snp_rate=5e-3, indel_rate=1e-3,
         somatic_positions=[3000, 6000, 9000], max_indel=20),
    # read lengths around the limits of the emission kernels (156 bases: lane per record; 248: staged by a lane group)
    dict(seed=108, contig_len=7000, n_pairs=(300, 300), read_len=156, snp_rate=3e-3, indel_rate=3e-3, clip_frac=0.3),
    dict(seed=109, contig_len=7000, n_pairs=(300, 300), read_len=158, snp_rate=3e-3, indel_rate=3e-3, clip_frac=0.3),
    dict(seed=110, contig_len=8000, n_pairs=(250, 250), read_len=200, snp_rate=3e-3, indel_rate=3e-3, clip_frac=0.3, max_indel=12),
    # germline indels longer than the lane-per-record kernel takes (32 bases) and records longer than its rows (188)
    dict(seed=111, contig_len=7000, n_pairs=(300, 300), read_len=150, snp_rate=2e-3, indel_rate=2e-3, clip_frac=0.2, max_indel=60),
]


@pytest.mark.parametrize("kw", RANDOM_CASES, ids=[f"seed{k['seed']}" for k in RANDOM_CASES])
@pytest.mark.parametrize("sparse_qual", [False, True], ids=["dense-qual", "sparse-qual"])
def test_engine_matches_oracle_random(engine, kw, sparse_qual):
    from oracle import oracle
    case = synth.make_case(**kw)
    batch = B.pack_reads(H.ordered_reads(case), sparse_qual=sparse_qual)
    sessions = B.pack_sessions(case["windows"])
    exp, st = oracle.run(batch, sessions, case["reference"])
    assert st == 0
    engine.upload_reference(0, case["reference"])
    got = engine.run(batch, sessions)
    assert_same_result(got, exp, kw["seed"])
    assert got.totals["n_modified"] > 0 or kw["n_pairs"][1] == 0


@pytest.mark.parametrize("kw", [k for k in RANDOM_CASES if k.get("indel_rate", 1e-3) > 0 and k["n_pairs"][1] > 0], ids=lambda k: f"seed{k['seed']}")
def test_engine_edit_descriptions_match_oracle(engine, kw):
    """ga_record_edits (what the per-sample driver and the plugin's read objects need for quirk Q12): every indel-masked
    record carries the edits the oracle applied, in application order; records and counters are unchanged by keeping them."""
    from oracle import oracle
    case = synth.make_case(**kw)
    batch = B.pack_reads(H.ordered_reads(case))
    sessions = B.pack_sessions(case["windows"])
    exp, st = oracle.run(batch, sessions, case["reference"], edits=True)
    assert st == 0
    engine.upload_reference(0, case["reference"])
    try:
        got = engine.run(batch, sessions, edits=True)
    finally:
        engine.keep_edits(False)
    assert_same_result(got, exp, kw["seed"])
    n = 0
    for key, rec in exp.records.items():
        if rec["qual"] is None:
            assert "edits" not in got.records[key]
            continue
        assert got.records[key]["edits"] == rec["edits"], (kw["seed"], key)
        n += rec["edits"] is not None
    assert n > 0


def _trim_clean_reads(case, seed, keep_min):
    """Mixed read lengths (trimmed reads): every other clean read loses a random tail."""
    rng = np.random.default_rng(seed)
    for r in case["reads"]:
        L = len(r["seq"])
        if r["cigar"] == f"{L}M" and rng.random() < 0.5:
            n = int(rng.integers(keep_min, L + 1))
            r["seq"], r["qual"], r["cigar"] = r["seq"][:n], r["qual"][:n], f"{n}M"
    return case


@pytest.mark.parametrize("kw,keep_min", [(dict(seed=401, contig_len=7000, n_pairs=(400, 400), read_len=150, indel_rate=1e-3, clip_frac=0.2), 1),
                                         (dict(seed=402, contig_len=9000, n_pairs=(300, 300), read_len=300, snp_rate=3e-3), 20),
                                         (dict(seed=403, contig_len=4000, n_pairs=(2500, 2000), read_len=36, snp_rate=3e-3), 30)],
                         ids=["mixed-150", "mixed-300", "short-36"])
def test_engine_mixed_read_lengths(engine, kw, keep_min):
    """Tiles whose records differ in length (variable TMA tile sizes, reads above the 256-base clean path, very
    short reads) against the oracle."""
    from oracle import oracle
    case = _trim_clean_reads(synth.make_case(**kw), kw["seed"], keep_min)
    batch = B.pack_reads(H.ordered_reads(case), sparse_qual=True)
    sessions = B.pack_sessions(case["windows"])
    exp, st = oracle.run(batch, sessions, case["reference"])
    assert st == 0
    engine.upload_reference(0, case["reference"])
    got = engine.run(batch, sessions)
    assert_same_result(got, exp, kw["seed"])
    assert got.totals["n_modified"] > 0


def test_engine_reports_fallback_sessions_and_stage_times(engine):
    """The diagnostic entry points: a read with three germline indels sends its session to the fallback kernel."""
    from oracle import oracle
    case = synth.make_case(seed=103, contig_len=5000, n_pairs=(300, 300), read_len=75, indel_rate=3e-3, clip_frac=0.5)
    batch = B.pack_reads(H.ordered_reads(case), sparse_qual=True)
    sessions = B.pack_sessions(case["windows"])
    engine.upload_reference(0, case["reference"])
    got = engine.run(batch, sessions)
    exp, st = oracle.run(batch, sessions, case["reference"])
    assert st == 0
    assert_same_result(got, exp, "diag")
    n, reasons = engine.fallback_sessions()
    assert n >= 0 and len(reasons) >= 9 and sum(reasons[:5]) == n
    for stage in range(4):
        h = engine.stage_ms_history(stage, 4)
        assert len(h) >= 1 and all(x >= 0 for x in h)
    assert engine.kernel_ms_history(1)[0] > 0


def test_engine_deep_session_uses_big_path(engine):
    """A session deeper than the shared-memory tables (SmemLayout caps) takes the global-scratch path."""
    from oracle import oracle
    case = synth.make_case(seed=201, contig_len=3000, n_pairs=(2600, 2600), read_len=100, somatic_positions=[1500])
    batch = B.pack_reads(H.ordered_reads(case), sparse_qual=True)
    sessions = B.pack_sessions(case["windows"])
    exp, st = oracle.run(batch, sessions, case["reference"])
    assert st == 0
    engine.upload_reference(0, case["reference"])
    got = engine.run(batch, sessions)
    assert int(got.sess_counts[0, 3]) > 4096
    assert_same_result(got, exp, "deep")


def test_engine_empty_inputs(engine):
    case = synth.make_case(seed=301, contig_len=3000, n_pairs=(20, 20), read_len=100)
    engine.upload_reference(0, case["reference"])
    batch = B.pack_reads(H.ordered_reads(case))
    none = B.pack_sessions([])
    res = engine.run(batch, none)
    assert res.totals["n_modified"] == 0 and res.totals["session_reads"] == 0
    empty = B.pack_reads([])
    res = engine.run(empty, B.pack_sessions(case["windows"]))
    assert res.totals["n_modified"] == 0
    far = B.pack_sessions([{"first": 100000, "last": 102001, "keep": None}])
    res = engine.run(batch, far)
    assert res.totals["session_reads"] == 0


def test_engine_capacity_error_reports_needed_sizes(engine):
    from genomeanonymizer_b200 import _abi
    from genomeanonymizer_b200.engine import DeviceBatch, DeviceResult, DeviceSessions
    import torch
    case = synth.make_case(seed=102, contig_len=9000, n_pairs=(400, 250), read_len=150, somatic_positions=[2500, 4700, 6900])
    batch = B.pack_reads(H.ordered_reads(case))
    sessions = B.pack_sessions(case["windows"])
    engine.upload_reference(0, case["reference"])
    db, ds = DeviceBatch(batch, engine.device), DeviceSessions(sessions, engine.device)
    small = DeviceResult(sessions.n_sessions, 4, 8, 8, engine.device)
    engine.run_device(db, ds, small)
    torch.cuda.synchronize()
    t = small.read_totals()
    assert t.error == _abi.GA_ERR_CAPACITY
    with pytest.raises(_abi.GaError):
        engine.check_device_status(small)
    # the totals say how much is needed: a second run with exactly that capacity succeeds
    exact = DeviceResult(sessions.n_sessions, int(t.n_modified), int(t.seq16_used), max(1, int(t.qual16_used)), engine.device)
    engine.run_device(db, ds, exact)
    torch.cuda.synchronize()
    engine.check_device_status(exact)


def test_engine_rejects_unknown_contig(engine):
    case = synth.make_case(seed=301, contig_len=3000, n_pairs=(20, 20), read_len=100)
    batch = B.pack_reads(H.ordered_reads(case), contig_id=77)
    with pytest.raises(ValueError):
        engine.run(batch, B.pack_sessions(case["windows"]))


def test_engine_read_past_reference_end_is_index_error(engine):
    """A read whose alignment runs off the contig raises like the reference's IndexError path."""
    case = synth.make_case(seed=301, contig_len=3000, n_pairs=(20, 20), read_len=100)
    engine.upload_reference(0, case["reference"][:1600])
    batch = B.pack_reads(H.ordered_reads(case))
    with pytest.raises(IndexError):
        engine.run(batch, B.pack_sessions(case["windows"]))


@pytest.mark.parametrize("kw", [dict(seed=501, contig_len=5000, n_pairs=(30, 30), read_len=250, indel_rate=1.2e-2, max_indel=6),
                                dict(seed=502, contig_len=5000, n_pairs=(40, 25), read_len=150, indel_rate=2.5e-2, max_indel=3, clip_frac=0.3),
                                dict(seed=503, contig_len=6000, n_pairs=(25, 40), read_len=300, indel_rate=1.5e-2, max_indel=12, snp_rate=4e-3)],
                         ids=["many-250", "many-150-clips", "many-300-snps"])
@pytest.mark.parametrize("sparse_qual", [False, True], ids=["dense-qual", "sparse-qual"])
def test_reads_with_many_germline_indels_stay_in_the_streaming_pipeline(engine, kw, sparse_qual):
    """Reads with three and more germline indels (edit lists in the side buffer: collect_many / emit_many_group)
    against the oracle; none of these sessions may need the fallback kernel for that reason."""
    from oracle import oracle
    case = synth.make_case(**kw)
    batch = B.pack_reads(H.ordered_reads(case), sparse_qual=sparse_qual)
    sessions = B.pack_sessions(case["windows"])
    exp, st = oracle.run(batch, sessions, case["reference"])
    assert st == 0
    engine.upload_reference(0, case["reference"])
    got = engine.run(batch, sessions)
    assert_same_result(got, exp, kw["seed"])
    n_many = sum(1 for r in H.ordered_reads(case) if sum(c in "ID" for c in r["cigar"]) > 2)
    assert n_many > 10 and got.totals["masked"][1] + got.totals["masked"][2] > 10, (n_many, got.totals)
    n, reasons = engine.fallback_sessions()
    assert reasons[4] == 0, reasons


def test_records_that_are_not_contiguous_in_seq4_take_the_unstaged_path(engine):
    """The scan kernel stages a tile with one TMA bulk copy only when its records are contiguous and ascending in
    seq4; any other layout the C ABI allows (here: records in reverse order with gaps between them) goes read by read
    through the whole-warp walk.  Same results either way."""
    from oracle import oracle
    case = synth.make_case(seed=601, contig_len=7000, n_pairs=(300, 280), read_len=150, indel_rate=1e-3, clip_frac=0.2,
                           somatic_positions=[2500, 4600])
    batch = B.pack_reads(H.ordered_reads(case))
    n = batch.n_reads
    units = np.maximum(1, ((batch.len_flag & 0xFFFF).astype(np.int64) + 31) // 32)
    new_off = np.zeros(n, np.int64)
    cur = 3
    for r in range(n - 1, -1, -1):                                    # last read first, one spare unit between records
        new_off[r] = cur
        cur += int(units[r]) + 1
    seq4 = np.zeros(16 * cur, np.uint8)
    qual = np.zeros(32 * cur, np.uint8)
    for r in range(n):
        o, u, s = int(batch.seq_off16[r]), int(units[r]), int(new_off[r])
        seq4[16 * s:16 * (s + u)] = batch.seq4[16 * o:16 * (o + u)]
        qual[32 * s:32 * (s + u)] = batch.qual[32 * o:32 * (o + u)]
    scattered = B.ReadBatch(n_tumor=batch.n_tumor, pos=batch.pos, len_flag=batch.len_flag, seq_off16=new_off.astype(np.uint32),
                            cigar_off=batch.cigar_off, cigar=batch.cigar, seq4=seq4, qual=qual, max_ref_span=batch.max_ref_span,
                            contig_id=0, names=batch.names)
    sessions = B.pack_sessions(case["windows"])
    exp, st = oracle.run(batch, sessions, case["reference"])
    assert st == 0
    engine.upload_reference(0, case["reference"])
    got = engine.run(scattered, sessions)
    assert_same_result(got, exp, "scattered records")
    assert got.totals["n_modified"] > 0


@pytest.mark.gpu
def test_engine_equals_oracle_on_random_and_twisted_samples():
    """A slice of tools/fuzz_parity.py (engine through the C ABI against the oracle, random shape parameters, dense and
    sparse qualities), plain and with trailing / leading insertions, hard clips and reference skips (helpers.twist_reads)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ([], ["--twist"]):
        res = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "920000", "20"] + extra, capture_output=True, text=True, timeout=900)
        assert res.returncode == 0 and "20 cases, 0 mismatches" in res.stdout, res.stdout[-600:] + res.stderr[-600:]
