"""Device-side synthetic generator, the device entry ga_run() and the host entry ga_run_host() on the same
workload: all three must agree with the CPU oracle bit for bit, for every chunking of the session table."""
import numpy as np
import pytest

from genomeanonymizer_b200 import synthdev as SD
from tests.test_gpu_parity import assert_same_result

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from genomeanonymizer_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("name", ["tiny", "tiny-stress", "tiny-varied"])
def test_device_generator_equals_host_twin(name):
    import torch
    cfg = SD.WORKLOADS[name]
    hb, hs, ref = SD.generate_host(cfg)
    db, ds = SD.generate_device(cfg, torch.device("cuda", 0))
    gb, gs = db.to_host(), ds.to_host()
    for k in ("pos", "len_flag", "seq_off16", "cigar_off", "cigar", "seq4", "qual_reads", "qual_off16"):
        assert np.array_equal(getattr(hb, k), getattr(gb, k)), k
    assert np.array_equal(hb.qual, gb.qual[:len(hb.qual)])
    assert hb.max_ref_span == gb.max_ref_span and hb.n_tumor == gb.n_tumor
    for k in ("first", "last", "keep_type", "keep_pos", "keep_end", "keep_len", "keep_allele_off"):
        assert np.array_equal(getattr(hs, k), getattr(gs, k)), k
    assert bytes(hs.keep_alleles[:hs.n_sessions]) == bytes(gs.keep_alleles[:gs.n_sessions])
    dref = SD.reference_device(cfg, torch.device("cuda", 0)).cpu().numpy().tobytes()
    assert dref == ref


@pytest.mark.parametrize("name", ["tiny", "tiny-stress", "tiny-varied"])
def test_device_entry_matches_oracle_on_synthetic_workload(engine, name):
    from oracle import oracle
    cfg = SD.WORKLOADS[name]
    hb, hs, ref = SD.generate_host(cfg)
    exp, st = oracle.run(hb, hs, ref)
    assert st == 0 and exp.totals["n_modified"] > 0 and sum(exp.totals["masked"]) > 0
    engine.upload_reference(0, ref)
    got = engine.run(hb, hs)
    assert_same_result(got, exp, name)


@pytest.mark.parametrize("name,chunk", [("tiny", 0), ("tiny", 1), ("tiny", 7), ("tiny-stress", 5), ("tiny-stress", 24), ("tiny-varied", 9)])
def test_host_entry_matches_oracle_for_every_chunking(engine, name, chunk):
    from genomeanonymizer_b200.engine import HostBatch, HostResult
    from oracle import oracle
    cfg = SD.WORKLOADS[name]
    hb, hs, ref = SD.generate_host(cfg)
    exp, st = oracle.run(hb, hs, ref)
    assert st == 0
    engine.upload_reference(0, ref)
    host = HostBatch(hb, hs)
    units = hb.seq4.shape[0] // 16
    out = HostResult(hs.n_sessions, hb.n_reads, units, units)
    engine.run_host(host, out, chunk)
    got = out.decode()
    assert_same_result(got, exp, (name, chunk))
    h2d, d2h = engine.host_traffic()
    assert h2d > hb.seq4.shape[0] and d2h > 0


def test_host_entry_retries_a_chunk_that_overflows_its_first_guess(engine):
    """A chunk whose modified records exceed the engine's first capacity guess is re-run with the exact need."""
    from genomeanonymizer_b200.engine import HostBatch, HostResult
    from dataclasses import replace
    from oracle import oracle
    cfg = replace(SD.WORKLOADS["tiny-stress"], snp_rate=2e-2, name="dense-snp")     # most reads get modified
    hb, hs, ref = SD.generate_host(cfg)
    exp, st = oracle.run(hb, hs, ref)
    assert st == 0 and exp.totals["n_modified"] > hb.n_reads // 3 + 1024
    engine.upload_reference(0, ref)
    units = hb.seq4.shape[0] // 16
    out = HostResult(hs.n_sessions, 2 * hb.n_reads, 2 * units, 2 * units)
    engine.run_host(HostBatch(hb, hs), out, 0)
    assert_same_result(out.decode(), exp, "retry")


def test_host_entry_reports_caller_capacity(engine):
    from genomeanonymizer_b200 import _abi
    from genomeanonymizer_b200.engine import HostBatch, HostResult
    cfg = SD.WORKLOADS["tiny"]
    hb, hs, ref = SD.generate_host(cfg)
    engine.upload_reference(0, ref)
    out = HostResult(hs.n_sessions, 8, 8, 8)
    with pytest.raises(_abi.GaError):
        engine.run_host(HostBatch(hb, hs), out, 10)
    assert out.totals.n_modified > 8            # the totals say how much is needed
