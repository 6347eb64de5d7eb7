#!/usr/bin/env python
"""Randomised parity: engine (ga_run through the C ABI) against the CPU oracle on many small seeded samples whose shape
parameters are drawn at random (read length 30-300, indel / clip / SNP / error rates, indel lengths up to 60, coverage,
window count).  On the dense-quality pass the edit descriptions (ga_record_edits) and the FASTQ text of every read of the
sample (ga_fastq_layout + ga_fastq_render over the masked records of every session) are compared as well.
usage: tools/fuzz_parity.py [first seed] [cases] [--twist]     exit code 1 when anything differed."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeanonymizer_b200 import batch as B           # noqa: E402
from genomeanonymizer_b200 import synth                # noqa: E402
from genomeanonymizer_b200.engine import Engine        # noqa: E402
from oracle import oracle                              # noqa: E402
from tests import helpers as H                         # noqa: E402


def fastq_equal(eng, batch, sessions, reads, exp, seed):
    """Every (session, read) record and every read as it came in, rendered on the device, against the oracle's renderer."""
    import torch
    from genomeanonymizer_b200.engine import DeviceBatch, DeviceResult, DeviceSessions
    from oracle import fastq as OF
    db, ds = DeviceBatch(batch, eng.device), DeviceSessions(sessions, eng.device)
    units = batch.seq4.shape[0] // 16
    dres = DeviceResult(sessions.n_sessions, 2 * batch.n_reads + 16, 2 * units + 64, 2 * units + 64, eng.device)
    eng.run_device(db, ds, dres)
    torch.cuda.synchronize()
    n = int(eng.check_device_status(dres).n_modified)
    ms, mr = dres.mod_session[:n].cpu().numpy(), dres.mod_read[:n].cpu().numpy()
    items = [(int(r), k, int(s)) for k, (s, r) in enumerate(zip(ms, mr))] + [(i, -1, -1) for i in range(len(reads))]
    text, off = eng.render_fastq(db, [r["name"] for r in reads], [i for i, _, _ in items], [k for _, k, _ in items], dres, n)
    for j, (i, k, s) in enumerate(items):
        if k >= 0:
            seq, qual = H.final_read(batch, exp, i, session=s)
        else:
            seq, qual = B.decode_bases(batch.sequence_codes(i)), [int(x) for x in batch.qualities(i)]
        if text[off[j]:off[j + 1]].decode("ascii") != OF.render(reads[i]["name"], reads[i]["flag"], seq, qual):
            print("FASTQ text differs", seed, "read", i, "record", k, flush=True)
            return False
    return True


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 100
    eng = Engine(0)
    bad = 0
    for seed in range(seed0, seed0 + n):
        rng = np.random.default_rng(seed)
        read_len = int(rng.choice([30, 50, 75, 100, 101, 125, 150, 151, 156, 157, 180, 200, 249, 250, 300]))
        n_t, n_n = int(rng.integers(20, 500)), int(rng.integers(0, 500))
        kw = dict(seed=seed, contig_len=int(rng.integers(3000, 9000)), n_pairs=(n_t, n_n), read_len=read_len,
                  snp_rate=float(rng.choice([1e-3, 3e-3, 1e-2])), indel_rate=float(rng.choice([0, 1e-4, 1e-3, 4e-3, 1e-2])),
                  clip_frac=float(rng.choice([0, 0.1, 0.5])), max_indel=int(rng.choice([3, 10, 20, 40, 60])))
        case = synth.make_case(**kw)
        if "--twist" in sys.argv:                                            # trailing / leading insertions, hard clips, reference skips
            case["reads"] = H.twist_reads(case["reads"])
        reads = [r for r in case["reads"] if r["dataset"] == 0] + [r for r in case["reads"] if r["dataset"] == 1]
        for sparse in (False, True):
            batch = B.pack_reads(reads, sparse_qual=sparse)
            sessions = B.pack_sessions(case["windows"])
            exp, st = oracle.run(batch, sessions, case["reference"], edits=not sparse)
            assert st == 0, (seed, st)
            eng.upload_reference(0, case["reference"])
            got = eng.run(batch, sessions, edits=not sparse)
            eng.keep_edits(False)
            ok = got.totals == exp.totals and sorted(got.records) == sorted(exp.records) and np.array_equal(got.sess_counts, exp.sess_counts)
            if ok:
                for k, v in got.records.items():
                    e = exp.records[k]
                    if not np.array_equal(v["seq"], e["seq"]) or (v["qual"] is None) != (e["qual"] is None) or (v["qual"] is not None and not np.array_equal(v["qual"], e["qual"])):
                        ok = False
                        print("record differs", seed, sparse, k, flush=True)
                        break
                    if not sparse and v.get("edits") != e.get("edits"):
                        ok = False
                        print("edit description differs", seed, k, v.get("edits"), e.get("edits"), flush=True)
                        break
            if ok and not sparse:
                ok = fastq_equal(eng, batch, sessions, reads, exp, seed)
            if not ok:
                bad += 1
                print("MISMATCH", seed, sparse, kw, got.totals, exp.totals, flush=True)
        if (seed - seed0) % 20 == 19:
            print(f"{seed - seed0 + 1} cases, {bad} mismatches", flush=True)
    eng.close()
    print(f"done: {n} cases, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
