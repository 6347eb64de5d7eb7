"""Wire form of a read batch (include/ga_wire.h).

CPU: the host packer against a Python decoder written from the header's format description - every read's position,
length / flag, CIGAR and bases must come back; block cuts at 1,024 reads and at position gaps beyond 16 bits.
GPU (-m gpu): ga_run_wire (expansion on the device) gives the records of ga_run / the oracle for every chunking."""
import numpy as np
import pytest

from genomeanonymizer_b200 import batch as B
from genomeanonymizer_b200 import synthdev as SD
from genomeanonymizer_b200 import wire as W


def decode_wire(w: W.WireBatch):
    """include/ga_wire.h, restated: returns per read (pos, len_flag, cigar words, base codes)."""
    out = []
    blob = w.blob
    al4 = lambda n: (n + 3) & ~3
    for b in range(w.n_blocks):
        o = int(w.dir["byte"][b])
        n, n_words, n_gen, n_gen_ops, n_exc, lc_nf, n_lenx, n_posx = (int(x) for x in blob[o:o + 32].view("<u4"))
        len_common, n_flags = lc_nf & 0xFFFF, lc_nf >> 16
        assert int(w.dir["read"][b]) == len(out) and o % 16 == 0
        p = o + 32
        if n_flags:
            fdict = blob[p:p + 32].view("<u2"); p += 32
            fidx = blob[p:p + (n + 1) // 2]; p += al4((n + 1) // 2)
            flag = [int(fdict[(int(fidx[i >> 1]) >> (4 * (i & 1))) & 15]) for i in range(n)]
            assert n_flags <= 16 and len(set(flag)) == n_flags
        else:
            flag = [int(x) for x in blob[p:p + 2 * n].view("<u2")]; p += al4(2 * n)
            assert len(set(flag)) > 16
        dpos = [int(x) for x in blob[p:p + n]]; p += al4(n)
        posx = blob[p:p + 4 * n_posx].view("<u4"); p += 4 * n_posx
        assert all(d < 255 or i in {int(x) >> 16 for x in posx} for i, d in enumerate(dpos))
        for x in posx:
            assert dpos[int(x) >> 16] == 255 and (int(x) & 0xFFFF) >= 255
            dpos[int(x) >> 16] = int(x) & 0xFFFF
        lenx = blob[p:p + 4 * n_lenx].view("<u4"); p += 4 * n_lenx
        gen_idx = blob[p:p + 2 * n_gen].view("<u2"); p += al4(2 * n_gen)
        gen_off = blob[p:p + 4 * (n_gen + 1)].view("<u4"); p += 4 * (n_gen + 1)
        gen_cig = blob[p:p + 4 * n_gen_ops].view("<u4"); p += 4 * n_gen_ops
        exc = blob[p:p + 4 * n_exc].view("<u4"); p += 4 * n_exc
        bases = blob[p:p + 4 * n_words].view("<u4"); p += 4 * n_words
        assert (p + 15) // 16 * 16 - o == int(w.dir["byte"][b + 1]) - o
        gen = {int(i): gen_cig[int(gen_off[j]):int(gen_off[j + 1])] for j, i in enumerate(gen_idx)}
        lengths = [len_common] * n
        for x in lenx:
            lengths[int(x) >> 16] = int(x) & 0xFFFF
        assert all(int(a) < int(b2) for a, b2 in zip(lenx[:-1] >> 16, lenx[1:] >> 16)) and all(lengths[int(x) >> 16] != len_common for x in lenx)
        patches = {}
        for e in exc:
            patches.setdefault(int(e) >> 20, []).append(((int(e) >> 4) & 0xFFFF, int(e) & 15))
        two_all = (np.repeat(bases, 16) >> (2 * np.tile(np.arange(16, dtype=np.uint32), n_words))) & 3      # the block's bases, back to back
        pos, at = int(w.dir["pos"][b]), 0
        for i in range(n):
            L = lengths[i]
            pos += int(dpos[i])
            codes = (1 << two_all[at:at + L]).astype(np.uint8)
            for q, c in patches.get(i, []):
                codes[q] = c
            at += L
            cig = gen[i] if i in gen else np.array([L << 4], np.uint32)
            out.append((pos, (int(flag[i]) << 16) | L, cig, codes))
        assert n_words == (at + 15) // 16 + 3 and not two_all[at:].any()
    return out


def assert_wire_equals_batch(w, b):
    reads = decode_wire(w)
    assert len(reads) == b.n_reads == w.n_reads
    for r, (pos, lf, cig, codes) in enumerate(reads):
        assert pos == int(b.pos[r]) and lf == int(b.len_flag[r]), r
        assert np.array_equal(cig, b.cigar[int(b.cigar_off[r]):int(b.cigar_off[r + 1])]), r
        assert np.array_equal(codes, b.sequence_codes(r)), r
    assert int(w.dir["unit"][-1]) == sum(max(1, (int(x & 0xFFFF) + 31) // 32) for x in b.len_flag)
    assert int(w.dir["ops"][-1]) == int(b.cigar_off[-1]) and int(w.dir["read"][w.n_tumor_blocks]) == b.n_tumor


@pytest.mark.parametrize("name", ["tiny", "tiny-stress"])
def test_packer_round_trips_through_the_documented_format(name):
    b, s, ref = SD.generate_host(SD.WORKLOADS[name], 0, 6)
    w = W.pack_wire(b)
    assert_wire_equals_batch(w, b)
    assert w.blob.nbytes < 0.6 * (b.seq4.nbytes + 16 * b.n_reads)        # about 36 instead of 100 bytes per 100-150 bp read
    assert w.max_ref_span == b.max_ref_span


def _reads(n, pos, seqs=None, cigars=None, L=50, dataset=0, rng=None):
    rng = rng or np.random.default_rng(1)
    out = []
    for k in range(n):
        seq = seqs[k] if seqs else "".join("ACGT"[x] for x in rng.integers(0, 4, L))
        out.append({"name": f"r{dataset}_{k}", "flag": 99 if k % 2 else 147, "pos": int(pos[k]), "cigar": cigars[k] if cigars else f"{len(seq)}M",
                    "seq": seq, "qual": [30] * len(seq), "dataset": dataset})
    return out


def test_blocks_are_cut_at_1024_reads_at_wide_gaps_and_between_datasets():
    pos_t = np.sort(np.concatenate([np.arange(0, 3000 * 3, 3), [80000, 80001, 200000]]))
    pos_n = np.arange(5, 1500 * 7, 7)
    b = B.pack_reads(_reads(len(pos_t), pos_t) + _reads(len(pos_n), pos_n, dataset=1), sparse_qual=True)
    w = W.pack_wire(b, threads=3)
    starts = [int(x) for x in w.dir["read"]]
    # 3000 dense reads -> blocks at 0, 1024, 2048; the gap 8997 -> 80000 is beyond 16 bits -> new block at read 3000;
    # 80001 -> 200000 again -> new block at 3002; then the normal dataset
    assert starts == [0, 1024, 2048, 3000, 3002, 3003, 3003 + 1024, len(pos_t) + len(pos_n)]
    assert w.n_tumor_blocks == 5 and w.n_blocks == 7
    assert_wire_equals_batch(w, b)


def test_other_base_codes_generic_cigars_and_odd_lengths_survive():
    seqs = ["ACGTNNACGTRYKM" * 3 + "A", "N", "ACGT" * 8, "acgtnACGT", "ACGTACGTACGTACGTA", "=ACGT"]
    cig = ["43M", "1M", "10M2I20M", "3S6M", "8M1D9M", "5="]
    pos = [10, 10, 11, 500, 67000, 67001]
    b = B.pack_reads(_reads(len(seqs), pos, seqs, cig), sparse_qual=True)
    w = W.pack_wire(b)
    assert_wire_equals_batch(w, b)
    assert w.n_blocks == 2                                                # 500 -> 67000 does not fit 16 bits


def test_more_than_sixteen_distinct_flags_and_every_position_step_survive():
    rng = np.random.default_rng(3)
    steps = np.concatenate([[0], rng.integers(0, 4, 300), [254, 255, 256, 1000, 65535], rng.integers(0, 600, 200)])
    pos = np.cumsum(steps)
    reads = _reads(len(pos), pos, L=37)
    for k, r in enumerate(reads):
        r["flag"] = 1 + 2 * (k % 23) + (0x10 if k % 3 else 0x20)          # 40-odd distinct values: the dictionary does not hold them
    many = B.pack_reads(reads, sparse_qual=True)
    w = W.pack_wire(many)
    assert_wire_equals_batch(w, many)
    for k, r in enumerate(reads):
        r["flag"] = [99, 147, 83, 163, 1123, 65, 129, 73, 133, 89, 153, 97, 145, 81, 161, 1187][k % 16]      # exactly sixteen: the dictionary is full
    w = W.pack_wire(B.pack_reads(reads, sparse_qual=True))
    assert_wire_equals_batch(w, B.pack_reads(reads, sparse_qual=True))


def test_unsorted_reads_are_refused():
    b = B.pack_reads(_reads(3, [5, 6, 7]), sparse_qual=True)
    b.pos = np.array([5, 9, 7], np.int32)
    with pytest.raises(ValueError):
        W.pack_wire(b)


# ---------------------------------------------------------------------------------------------------------------- GPU

@pytest.fixture(scope="module")
def engine():
    from genomeanonymizer_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,chunk", [("tiny", 0), ("tiny", 1), ("tiny", 7), ("tiny-stress", 5), ("tiny-stress", 24), ("tiny-varied", 6)])
def test_wire_entry_matches_oracle_for_every_chunking(engine, name, chunk):
    from genomeanonymizer_b200.engine import HostResult, HostWire
    from oracle import oracle
    from tests.test_gpu_parity import assert_same_result
    cfg = SD.WORKLOADS[name]
    hb, hs, ref = SD.generate_host(cfg)
    exp, st = oracle.run(hb, hs, ref)
    assert st == 0
    engine.upload_reference(0, ref)
    w = W.pack_wire(hb)
    units = hb.seq4.shape[0] // 16
    out = HostResult(hs.n_sessions, hb.n_reads, units, units)
    engine.run_wire(HostWire(w, hs), out, chunk)
    assert_same_result(out.decode(), exp, (name, chunk))
    h2d, d2h = engine.host_traffic()
    assert h2d > 0 and d2h > 0
    if chunk == 0:                                                        # one-session chunks re-send whole blocks
        assert h2d < 0.62 * (hb.seq4.nbytes + 20 * hb.n_reads + hb.qual.nbytes)


@pytest.mark.gpu
def test_wire_entry_with_n_bases_iupac_generic_reads_and_dense_source_qualities(engine):
    """A hand-made session: N and IUPAC bases (exception list), soft clips and indels (generic list), reads of odd
    lengths, a batch whose qualities were dense (the wire form carries the sparse records)."""
    from genomeanonymizer_b200 import synth
    from genomeanonymizer_b200.engine import HostResult, HostWire
    from oracle import oracle
    from tests.test_gpu_parity import assert_same_result
    case = synth.make_case(seed=21, contig_len=9000, n_pairs=(260, 240), read_len=101, somatic_positions=[3000, 6200], indel_rate=2e-3,
                           clip_frac=0.2, n_rate=5e-3)
    reads = [r for r in case["reads"] if r["dataset"] == 0] + [r for r in case["reads"] if r["dataset"] == 1]
    hb = B.pack_reads(reads, sparse_qual=False)
    hs = B.pack_sessions(case["windows"])
    exp, st = oracle.run(hb, hs, case["reference"])
    assert st == 0 and exp.totals["n_modified"] > 10
    engine.upload_reference(0, case["reference"])
    w = W.pack_wire(hb)
    assert_wire_equals_batch(w, hb)
    units = hb.seq4.shape[0] // 16
    out = HostResult(hs.n_sessions, hb.n_reads, units, units)
    engine.run_wire(HostWire(w, hs), out, 1)
    assert_same_result(out.decode(), exp, "hand-made")
