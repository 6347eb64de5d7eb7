"""The native plan (include/ga_plan.h, csrc/ga_plan.cpp) against its checker, driver.plan_sample: same sessions, same
write plan, same single-end spill on the reference's golden samples and on seeded samples with sparse coverage (island
sessions), orphans, duplicated names and no windows at all.  The Python plan itself is pinned to the reference's
output files (tests/test_genome_files.py)."""
import numpy as np
import pytest

from genomeanonymizer_b200 import batch as B
from genomeanonymizer_b200 import driver as D
from genomeanonymizer_b200 import genome_files as GF
from genomeanonymizer_b200 import synth
from genomeanonymizer_b200 import _lib
from tests import helpers as H

GENOME = H.load_golden("genome_cases.json")["cases"]


def contig_batch(reads):
    """ContigBatch of read dicts (tumor first) without going through a BAM file."""
    rs = [r for r in reads if r["dataset"] == 0] + [r for r in reads if r["dataset"] == 1]
    b = B.pack_reads(rs)
    ends = np.array([r["pos"] + B.ref_span(B.parse_cigar(r["cigar"])) for r in rs], np.int32)
    enc = [r["name"].encode() for r in rs]
    off = np.zeros(len(enc) + 1, np.int64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    blob = np.frombuffer(b"".join(enc) or b"\0", np.uint8).copy()
    table = [dict(name=r["name"], flag=r["flag"], dataset=r["dataset"], pos=r["pos"], end=int(e)) for r, e in zip(rs, ends)]
    return GF.ContigBatch(batch=b, ref_end=ends, name_blob=blob[:off[-1]], name_off=off), table


def same_plan(native, py):
    assert [(s["first"], s["last"], s["window"], s["keep"]) for s in native.sessions] == [(s["first"], s["last"], s["window"], s["keep"]) for s in py.sessions]
    assert np.asarray(native.pairs).reshape(-1, 5).tolist() == [list(p) for p in py.pairs]
    assert np.asarray(native.singles).reshape(-1, 4)[:, :3].tolist() == [list(p) for p in py.singles]


def test_library_exports_the_plan_abi():
    L = _lib.lib()
    hdr = open(__file__.replace("tests/test_plan_native.py", "include/ga_plan.h")).read()
    for name in GF.PLAN_EXPORTS:
        assert hasattr(L, name) and name + "(" in hdr, name


@pytest.mark.parametrize("entry", GENOME, ids=[e["case"]["name"] for e in GENOME])
def test_native_plan_equals_python_plan_on_golden_samples(entry):
    case = entry["case"]
    cb, table = contig_batch(case["reads"])
    same_plan(GF.plan_contig(cb, case["windows"], len(case["reference"])), D.plan_sample(table, case["windows"], len(case["reference"])))


RANDOM = [
    dict(seed=701, contig_len=12000, n_pairs=(60, 50), read_len=100, somatic_positions=[2500, 6200, 9800]),          # ~1x: islands
    dict(seed=702, contig_len=9000, n_pairs=(400, 380), read_len=150, somatic_positions=[2500, 4700, 6900]),
    dict(seed=703, contig_len=20000, n_pairs=(150, 20), read_len=75, somatic_positions=[5000, 15000]),               # thin normal
    dict(seed=704, contig_len=6000, n_pairs=(80, 80), read_len=60, somatic_positions=[]),                             # no windows
    dict(seed=705, contig_len=15000, n_pairs=(25, 30), read_len=120, somatic_positions=[4000, 11000], clip_frac=0.5),
]


def unmap_some(reads, every):
    """Every `every`-th pair: one mate becomes an unmapped read placed at its mate's position - behind it in the file,
    or (every other time) in front of it, which can make it the first read a region fetches."""
    by_name = {}
    for k, r in enumerate(reads):
        by_name.setdefault((r["dataset"], r["name"]), []).append(k)
    front = set()
    n = 0
    for key in sorted(by_name):
        idx = by_name[key]
        if len(idx) != 2:
            continue
        n += 1
        if n % every:
            continue
        keep, unm = (idx[0], idx[1]) if (n // every) % 2 else (idx[1], idx[0])
        reads[unm] = dict(reads[unm], pos=reads[keep]["pos"], cigar="*", flag=(reads[unm]["flag"] & 0xC0) | 0x5)
        reads[keep] = dict(reads[keep], flag=(reads[keep]["flag"] & ~0x22) | 0x8)
        if (n // every) % 3 == 0:
            front.add(unm)
    order = sorted(range(len(reads)), key=lambda k: (reads[k]["dataset"], reads[k]["pos"], 0 if k in front else (2 if reads[k]["flag"] & 4 else 1)))
    return [reads[k] for k in order]


@pytest.mark.parametrize("kw", RANDOM, ids=[f"seed{k['seed']}" for k in RANDOM])
@pytest.mark.parametrize("drop", [0, 5, 2])
@pytest.mark.parametrize("unmap", [0, 4])
def test_native_plan_equals_python_plan_on_random_samples(kw, drop, unmap):
    case = synth.make_case(**kw)
    reads = [dict(r) for r in case["reads"]]
    if unmap:
        reads = unmap_some(reads, unmap)
    if drop:                                                         # orphans: every `drop`-th read disappears
        reads = [r for k, r in enumerate(reads) if k % drop != 1]
    if kw["seed"] == 705:                                            # a duplicated alignment of the same (name, mate)
        reads = reads + [dict(reads[3]), dict(reads[11])]
        reads = sorted(reads, key=lambda r: (r["dataset"], r["pos"]))
    cb, table = contig_batch(reads)
    same_plan(GF.plan_contig(cb, case["windows"], kw["contig_len"]), D.plan_sample(table, case["windows"], kw["contig_len"]))


def test_native_plan_rejects_what_the_python_plan_rejects():
    case = GENOME[0]["case"]
    cb, table = contig_batch(case["reads"])
    close = [{"first": 1000, "last": 3001, "keep": None}, {"first": 2000, "last": 4001, "keep": None}]   # closer than a window
    with pytest.raises(ValueError):
        D.plan_sample(table, close, len(case["reference"]))
    with pytest.raises(ValueError):
        GF.plan_contig(cb, close, len(case["reference"]))
    empty, _ = contig_batch([])
    p = GF.plan_contig(empty, case["windows"], len(case["reference"]))
    assert len(p.sessions) == len(case["windows"]) and len(p.pairs) == 0 and len(p.singles) == 0
