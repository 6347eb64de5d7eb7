"""Record digest (include/ga_digest.h): the arithmetic both sides of a parity check at scale agree on.

CPU: the oracle's digest against a line-by-line Python restatement of the header, independence of record order and
of how the sessions are cut into shards, sensitivity to a single changed base / quality / length.
GPU (-m gpu): ga_result_digest of the engine's result equals the oracle's digest of the oracle's result, record by
record, on slices of the BASELINE workloads of >= 2,000 windows through both ga_run and ga_run_host."""
import ctypes as C

import numpy as np
import pytest

from genomeanonymizer_b200 import _abi
from genomeanonymizer_b200 import synthdev as SD

M64 = (1 << 64) - 1
SEEDS = (0x243F6A8885A308D3, 0x13198A2E03707344)
MULS = (0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F)


def _mix(h, w, mul):
    h ^= w
    h = (h * mul) & M64
    return h ^ (h >> 29)


def _fin(h):
    h ^= h >> 33; h = (h * 0xff51afd7ed558ccd) & M64
    h ^= h >> 33; h = (h * 0xc4ceb9fe1a85ec53) & M64
    return h ^ (h >> 33)


def python_record_hash(contig, session, gid, codes, qual):
    """include/ga_digest.h, restated: codes = base codes (one per element), qual = printed qualities or None."""
    out = []
    n = len(codes)
    for seed, mul in zip(SEEDS, MULS):
        h = seed
        for w in (contig, session, gid, n | ((1 if qual is not None else 0) << 32)):
            h = _mix(h, w, mul)
        for b in range(0, n, 8):
            h = _mix(h, sum(int(c) << (4 * i) for i, c in enumerate(codes[b:b + 8])), mul)
        if qual is not None:
            for b in range(0, n, 4):
                h = _mix(h, sum(int(q) << (8 * i) for i, q in enumerate(qual[b:b + 4])), mul)
        out.append(_fin(h))
    return out


def oracle_digest_of(cfg, w0=0, nw=None, contig=0, records=True):
    from oracle import oracle
    b, s, ref = SD.generate_host(cfg, w0, nw)
    raw, st = oracle.run(b, s, ref, decode=False)
    assert st == 0
    pl = cfg.plan()
    per_t, per_n = pl.reads_per_window
    n = int(raw["totals"].n_modified)
    return oracle.digest(raw["result"], n, session_base=w0, tumor_base=w0 * per_t, normal_base=w0 * per_n, n_tumor=b.n_tumor,
                         contig=contig, records=records), raw, b


def test_oracle_digest_matches_the_python_restatement_of_the_header():
    from oracle import oracle
    cfg = SD.WORKLOADS["tiny-stress"]
    b, s, ref = SD.generate_host(cfg, 3, 4)
    res, st = oracle.run(b, s, ref)
    raw, st2 = oracle.run(b, s, ref, decode=False)
    assert st == 0 and st2 == 0 and res.totals["n_modified"] > 50 and res.totals["indel_records"] > 5
    n = int(raw["totals"].n_modified)
    dig, keys, hashes = oracle.digest(raw["result"], n, session_base=1000, tumor_base=77, normal_base=99, n_tumor=b.n_tumor, contig=5, records=True)
    ms, mr = raw["arrays"][0], raw["arrays"][1]
    sums = [0, 0, 0, 0]
    for k in range(n):
        rec = res.records[(int(ms[k]), int(mr[k]))]
        r = int(mr[k])
        gid = 77 + r if r < b.n_tumor else (1 << 40) | (99 + r - b.n_tumor)
        assert [int(keys[k, 0]), int(keys[k, 1])] == [1000 + int(ms[k]), gid]
        lo, hi = python_record_hash(5, 1000 + int(ms[k]), gid, rec["seq"], rec["qual"])
        assert [int(hashes[k, 0]), int(hashes[k, 1])] == [lo, hi], k
        sums = [(sums[0] + lo) & M64, (sums[1] + hi) & M64, sums[2] + 1, sums[3] + len(rec["seq"])]
    assert [int(x) for x in dig] == sums


@pytest.mark.parametrize("name", ["tiny", "tiny-stress"])
def test_digest_is_independent_of_the_shard_cut(name):
    from oracle import oracle
    cfg = SD.WORKLOADS[name]
    (full, fk, fh), _, _ = oracle_digest_of(cfg)
    acc = np.zeros(4, np.uint64)
    keys, hashes = [], []
    for w0, nw in ((0, 7), (7, 1), (8, cfg.total_windows - 8)):
        (d, k, h), _, _ = oracle_digest_of(cfg, w0, nw)
        with np.errstate(over="ignore"):
            acc += d
        keys.append(k); hashes.append(h)
    assert np.array_equal(acc, full) and int(full[2]) > 100
    assert oracle.compare_records(np.concatenate(keys), np.concatenate(hashes), fk, fh) == 0
    perm = np.random.default_rng(0).permutation(len(fk))
    assert oracle.compare_records(fk[perm], fh[perm], fk, fh) == 0


def test_digest_sees_a_single_changed_nibble_quality_or_length():
    from oracle import oracle
    cfg = SD.WORKLOADS["tiny-stress"]
    (full, fk, fh), raw, b = oracle_digest_of(cfg)
    n = int(raw["totals"].n_modified)
    ms, mr, ml, mso, mqo, oseq, oqual, _ = raw["arrays"]
    kq = int(np.nonzero(mqo[:n] != 0xFFFFFFFF)[0][0])

    def again():
        return oracle.digest(raw["result"], n, 0, 0, 0, b.n_tumor, 0, records=True)

    oseq[16 * int(mso[3])] ^= 0x10                       # second base of record 3
    d, k, h = again()
    assert oracle.compare_records(k, h, fk, fh) == 1 and not np.array_equal(d[:2], full[:2])
    oseq[16 * int(mso[3])] ^= 0x10
    oqual[32 * int(mqo[kq]) + 1] += 1                    # one quality of the first record that carries qualities
    d, k, h = again()
    assert oracle.compare_records(k, h, fk, fh) == 1
    oqual[32 * int(mqo[kq]) + 1] -= 1
    ml[5] -= 1                                           # one record one base shorter
    d, k, h = again()
    assert oracle.compare_records(k, h, fk, fh) == 1 and int(d[3]) == int(full[3]) - 1
    ml[5] += 1
    # padding behind the record is not observable
    L = int(ml[3])
    base, units = 16 * int(mso[3]), max(1, (L + 31) // 32)
    if L % 2:
        oseq[base + L // 2] ^= 0xf0                      # the unused high nibble of the last byte
    for p in range(base + (L + 1) // 2, base + 16 * units):
        oseq[p] ^= 0xff
    d, k, h = again()
    assert oracle.compare_records(k, h, fk, fh) == 0 and np.array_equal(d, full)
    assert oracle.compare_records(k[1:], h[1:], fk, fh) == 1     # a missing record counts


# ---------------------------------------------------------------------------------------------------------------- GPU

@pytest.fixture(scope="module")
def engine():
    from genomeanonymizer_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _device_run(engine, cfg, w0, nw):
    import torch
    from genomeanonymizer_b200.engine import DeviceResult
    dev = torch.device("cuda", 0)
    db, ds = SD.generate_device(cfg, dev, w0, nw)
    ref = SD.reference_device(cfg, dev)
    engine.upload_reference(0, ref)
    units = db.seq4_bytes // 16
    cap = db.n_reads // 2 + 1024
    upr = max(1, units // max(1, db.n_reads))
    dres = DeviceResult(nw, cap, cap * (upr + 1), cap * (upr + 1), dev)
    engine.run_device(db, ds, dres)
    torch.cuda.synchronize()
    tot = engine.check_device_status(dres)
    return db, ds, dres, tot, ref


SCALE_SLICES = [("tiny", 0, 40), ("tiny-stress", 0, 24), ("chr1-30x-50k", 21000, 2200), ("cigar-stress", 4000, 2000),
                ("dense-60x30x", 50000, 2048), ("varied-depth", 17000, 2000)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,w0,nw", SCALE_SLICES)
def test_engine_records_equal_oracle_records_at_scale(engine, name, w0, nw):
    """Every modified record of a >= 2,000-window slice of each BASELINE workload, compared by key and 128-bit hash:
    ga_run (device buffers) and ga_run_host (host buffers, chunked) against the oracle."""
    import torch
    from genomeanonymizer_b200.engine import HostBatch, HostResult
    from oracle import oracle
    cfg = SD.WORKLOADS[name]
    db, ds, dres, tot, ref = _device_run(engine, cfg, w0, nw)
    pl = cfg.plan()
    per_t, per_n = pl.reads_per_window
    ids = dict(session_base=w0, tumor_base=w0 * per_t, normal_base=w0 * per_n, n_tumor=db.n_tumor, contig=3)
    n = int(tot.n_modified)
    dig, keys, hashes = engine.digest(dres, n, records=True, **ids)
    torch.cuda.synchronize()
    hb, hs = db.to_host(), ds.to_host()
    raw, st = oracle.run(hb, hs, ref.cpu().numpy().tobytes(), decode=False)
    assert st == 0
    n_exp = int(raw["totals"].n_modified)
    edig, ekeys, ehashes = oracle.digest(raw["result"], n_exp, records=True, **ids)
    assert n == n_exp and n > nw
    assert oracle.compare_records(keys.cpu().numpy(), hashes.cpu().numpy(), ekeys, ehashes) == 0
    assert np.array_equal(dig.cpu().numpy().view(np.uint64), edig)
    assert np.array_equal(dres.sess_counts.view(-1, 4)[:nw].cpu().numpy().view(np.uint32), raw["counts"][:nw])
    for f in ("session_reads", "session_bases", "indel_records", "seq16_used", "qual16_used"):
        assert int(getattr(tot, f)) == int(getattr(raw["totals"], f)), f
    # the host entry, several chunks on the three lanes
    out = HostResult(nw, n + 64, int(tot.seq16_used) + 64, int(tot.qual16_used) + 64)
    th = engine.run_host(HostBatch(hb, hs), out, max(1, nw // 5))
    assert int(th.n_modified) == n
    hdig, hkeys, hhashes = oracle.digest(out.as_struct(), n, records=True, **ids)
    assert oracle.compare_records(hkeys, hhashes, ekeys, ehashes) == 0
    assert np.array_equal(hdig, edig)
    assert np.array_equal(out.sess_counts.numpy().view(np.uint32)[:4 * nw].reshape(-1, 4), raw["counts"][:nw])
    # the wire entry (include/ga_wire.h): same records from ~44 bytes per read
    from genomeanonymizer_b200.engine import HostWire
    from genomeanonymizer_b200.wire import pack_wire
    out2 = HostResult(nw, n + 64, int(tot.seq16_used) + 64, int(tot.qual16_used) + 64)
    tw = engine.run_wire(HostWire(pack_wire(hb), hs), out2, max(1, nw // 4))
    assert int(tw.n_modified) == n
    wdig, wkeys, whashes = oracle.digest(out2.as_struct(), n, records=True, **ids)
    assert oracle.compare_records(wkeys, whashes, ekeys, ehashes) == 0
    assert np.array_equal(wdig, edig)
    assert np.array_equal(out2.sess_counts.numpy().view(np.uint32)[:4 * nw].reshape(-1, 4), raw["counts"][:nw])


@pytest.mark.gpu
def test_device_digest_accumulates_over_shards(engine):
    import torch
    cfg = SD.WORKLOADS["tiny"]
    pl = cfg.plan()
    per_t, per_n = pl.reads_per_window
    db, ds, dres, tot, _ = _device_run(engine, cfg, 0, cfg.total_windows)
    full = engine.digest(dres, int(tot.n_modified), n_tumor=db.n_tumor).cpu().numpy()
    acc = torch.zeros(4, dtype=torch.int64, device="cuda:0")
    for w0, nw in ((0, 13), (13, 27)):
        db, ds, dres, tot, _ = _device_run(engine, cfg, w0, nw)
        engine.digest(dres, int(tot.n_modified), session_base=w0, tumor_base=w0 * per_t, normal_base=w0 * per_n, n_tumor=db.n_tumor,
                      accumulate=acc)
    assert np.array_equal(acc.cpu().numpy(), full)


@pytest.mark.gpu
def test_runs_on_several_streams_overlap_on_the_engine_lanes_and_stay_correct(engine):
    """ga_run gives every stream its own lane (scratch + side stream); with more streams than lanes a lane is reused
    behind the run that held it.  Five shards on five streams, three rounds, in flight together: every round's
    records must equal the one-stream result."""
    import torch
    from genomeanonymizer_b200.engine import DeviceResult
    cfg = SD.WORKLOADS["tiny-stress"]
    dev = torch.device("cuda", 0)
    engine.upload_reference(0, SD.reference_device(cfg, dev))
    pl = cfg.plan()
    per_t, per_n = pl.reads_per_window
    cuts = [(0, 5), (5, 4), (9, 6), (15, 2), (17, 7)]
    shards = []
    for w0, nw in cuts:
        db, ds = SD.generate_device(cfg, dev, w0, nw)
        units = db.seq4_bytes // 16
        shards.append((db, ds, DeviceResult(nw, db.n_reads + 16, 2 * units + 64, 2 * units + 64, dev), w0))
    want = torch.zeros(4, dtype=torch.int64, device=dev)
    for db, ds, dres, w0 in shards:
        engine.run_device(db, ds, dres)
        torch.cuda.synchronize()
        t = engine.check_device_status(dres)
        engine.digest(dres, int(t.n_modified), session_base=w0, tumor_base=w0 * per_t, normal_base=w0 * per_n, n_tumor=db.n_tumor, accumulate=want)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev) for _ in shards]
    for _ in range(3):
        for (db, ds, dres, w0), st in zip(shards, streams):
            dres.out_seq4.zero_(); dres.mod_len.zero_()
        torch.cuda.synchronize()
        for (db, ds, dres, w0), st in zip(shards, streams):
            engine.run_device(db, ds, dres, stream=st)
        torch.cuda.synchronize()
        got = torch.zeros(4, dtype=torch.int64, device=dev)
        for db, ds, dres, w0 in shards:
            t = engine.check_device_status(dres)
            engine.digest(dres, int(t.n_modified), session_base=w0, tumor_base=w0 * per_t, normal_base=w0 * per_n, n_tumor=db.n_tumor, accumulate=got)
        assert np.array_equal(got.cpu().numpy(), want.cpu().numpy()) and int(want[2]) > 100
