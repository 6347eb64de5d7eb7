"""N1: FASTQ rendering.  CPU: the oracle restatement (oracle/fastq.py) against the "fastq" strings the
reference itself produced (tests/golden/session_cases.json).  GPU: ga_fastq_layout / ga_fastq_render through the C ABI
against the same strings, and against the oracle on seeded random sessions (masked and pass-through reads)."""
import numpy as np
import pytest

from genomeanonymizer_b200 import batch as B
from genomeanonymizer_b200 import synth
from oracle import fastq as OF
from oracle import oracle
from tests import helpers as H

GOLD = H.load_golden("session_cases.json")["cases"]


def golden_fastq(case, exp, batch, result):
    """(reads in batch order that the session holds, their golden record text) for ONE session."""
    reads = H.ordered_reads(case)
    out = []
    for i, r in enumerate(reads):
        e = exp["reads"].get(H.case_key(r))
        if e is not None:
            out.append((i, r, e["fastq"] + "\n"))
    return out


@pytest.mark.parametrize("entry", GOLD, ids=[e["case"]["name"] for e in GOLD])
def test_oracle_fastq_matches_reference(entry):
    case = entry["case"]
    batch = B.pack_reads(H.ordered_reads(case))
    for w, exp in zip(case["windows"], entry["expected"]):
        res, st = oracle.run(batch, B.pack_sessions([w]), case["reference"])
        assert st == 0
        for i, r, text in golden_fastq(case, exp, batch, res):
            seq, qual = H.final_read(batch, res, i)
            assert OF.render(r["name"], r["flag"], seq, qual) == text, (case["name"], r["name"])


def test_oracle_reverse_complement_table():
    assert OF.reverse_complement("ACGTN") == "NACGT"
    assert OF.render("q", 0x40, "ACGTN", [1, 2, 3, 4, 5]) == "@q/1\nACGTN\n+\n\"#$%&\n"
    assert OF.render("q", 0x80 | 0x10, "AACGT", [1, 2, 3, 4, 5]) == "@q/2\nACGTT\n+\n\"#$%&\n"     # Q1: qualities stay in BAM order


@pytest.fixture(scope="module")
def engine():
    from genomeanonymizer_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def device_render(engine, batch, sessions, names, reads_idx):
    """Masks on the device, then renders `reads_idx` of session 0 (modified records where there are any)."""
    import torch
    from genomeanonymizer_b200.engine import DeviceBatch, DeviceResult, DeviceSessions
    db, ds = DeviceBatch(batch, engine.device), DeviceSessions(sessions, engine.device)
    units = batch.seq4.shape[0] // 16
    dres = DeviceResult(sessions.n_sessions, 2 * batch.n_reads + 16, 2 * units + 64, 2 * units + 64, engine.device)
    engine.run_device(db, ds, dres)
    torch.cuda.synchronize()
    tot = engine.check_device_status(dres)
    n = int(tot.n_modified)
    ms, mr = dres.mod_session[:n].cpu().numpy(), dres.mod_read[:n].cpu().numpy()
    rec_of = {int(r): k for k, (s, r) in enumerate(zip(ms, mr)) if s == 0}
    recs = [rec_of.get(int(i), -1) for i in reads_idx]
    text, off = engine.render_fastq(db, names, reads_idx, recs, dres, n)
    return [text[off[k]:off[k + 1]].decode("ascii") for k in range(len(reads_idx))], sum(1 for x in recs if x >= 0)


@pytest.mark.gpu
@pytest.mark.parametrize("entry", GOLD, ids=[e["case"]["name"] for e in GOLD])
def test_device_fastq_matches_reference(engine, entry):
    case = entry["case"]
    reads = H.ordered_reads(case)
    batch = B.pack_reads(reads)                                      # dense qualities: every read can be printed
    names = [r["name"] for r in reads]
    engine.upload_reference(0, case["reference"])
    for w, exp in zip(case["windows"], entry["expected"]):
        gold = golden_fastq(case, exp, batch, None)
        got, _ = device_render(engine, batch, B.pack_sessions([w]), names, [i for i, _, _ in gold])
        assert got == [t for _, _, t in gold], case["name"]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(seed=301, contig_len=6000, n_pairs=(300, 300), read_len=150),
                                dict(seed=302, contig_len=5000, n_pairs=(250, 250), read_len=101, indel_rate=3e-3, clip_frac=0.4)],
                         ids=["snv", "indel-clip"])
def test_device_fastq_matches_oracle_on_random_sessions(engine, kw):
    case = synth.make_case(**kw)
    reads = [r for r in case["reads"] if r["dataset"] == 0] + [r for r in case["reads"] if r["dataset"] == 1]
    batch = B.pack_reads(reads)
    sessions = B.pack_sessions(case["windows"][:1])
    exp, st = oracle.run(batch, sessions, case["reference"])
    assert st == 0
    engine.upload_reference(0, case["reference"])
    idx = list(range(len(reads)))                                     # masked and untouched reads alike
    got, n_masked = device_render(engine, batch, sessions, [r["name"] for r in reads], idx)
    assert n_masked == sum(1 for (s, _) in exp.records if s == 0) and n_masked > 0
    for i, r in enumerate(reads):
        seq, qual = H.final_read(batch, exp, i)
        assert got[i] == OF.render(r["name"], r["flag"], seq, qual), (kw["seed"], i)


@pytest.mark.gpu
def test_device_fastq_every_alignment_of_every_section(engine):
    """Read lengths from 1 to beyond a staging slice, names of 1 to 60 characters, both strands: the header, base and
    quality sections of a record start at every byte alignment of the text (the renderer produces aligned words and
    shifts), records chain into spans or not, and the longest take the general path."""
    rng = np.random.default_rng(77)
    lengths = [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 101, 127, 128, 129, 149, 150, 151, 200, 241, 247, 248, 249, 250,
               255, 256, 257, 300]
    ref_len = 4000
    reference = "".join(rng.choice(list("ACGT"), size=ref_len))
    reads = []
    for ds in (0, 1):
        for k in range(400):
            L = int(lengths[(k * 7 + ds) % len(lengths)]) if k % 3 else int(rng.integers(1, 260))
            pos = int(rng.integers(0, ref_len - 320))
            nl = 1 + (k * 5 + ds) % 60
            name = "".join(rng.choice(list("abcXYZ019:_"), size=nl)) + f"{ds}{k}"
            flag = (0x40 if k % 2 else 0x80) | (0x10 if (k // 2) % 2 else 0) | 1
            seq = "".join(rng.choice(list("ACGTN"), size=L, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
            reads.append(dict(name=name, flag=flag, pos=pos, cigar=f"{L}M", seq=seq, qual=[int(x) for x in rng.integers(0, 94, size=L)], dataset=ds))
    reads.sort(key=lambda r: (r["dataset"], r["pos"]))
    batch = B.pack_reads(reads)
    sessions = B.pack_sessions([{"first": 1000, "last": 3001, "keep": None}])
    exp, st = oracle.run(batch, sessions, reference)
    assert st == 0
    engine.upload_reference(0, reference)
    for idx in (list(range(len(reads))), list(range(len(reads) - 1, -1, -3))):          # file order (records chain into spans) and scattered
        got, n_masked = device_render(engine, batch, sessions, [r["name"] for r in reads], idx)
        for k, i in enumerate(idx):
            seq, qual = H.final_read(batch, exp, i)
            assert got[k] == OF.render(reads[i]["name"], reads[i]["flag"], seq, qual), (i, len(reads[i]["seq"]))
